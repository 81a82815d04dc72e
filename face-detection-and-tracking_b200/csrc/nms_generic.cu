// Greedy NMS for the cases k_sort_nms (detect.cu) does not take: more than FDT_MAX_NMS_TOP_K candidates (layers/box_utils.nms has no
// cap, box_utils.py:296-298) and float64 inputs (MTCNN's `nms` runs in the dtype of `dets`, float64 in its pipeline,
// MTCNN/mtcnn/core/utils.py:62-113, callers core/detect.py:314, 326, 431, 579).  The textbook three-step formulation, any n:
//
//   sort     (order-preserving key of the score, index) descending with a bitonic network over global memory (one launch per
//            step: these are the rare paths, n is a few thousand boxes in MTCNN) -- ties: higher index first, the library's rule;
//   mask     grid of 64 x 64 tiles over the sorted boxes: bit j of row i  <=>  box i (higher score) suppresses box j, by the
//            reference's overlap rule in the input dtype (FDT_NMS_* flags, operand order and NaN behaviour as in detect.cu);
//   reduce   one block walks the rows in order: a row that is not yet removed is kept and ORs its bits into the removed set.
#include "fdt_common.cuh"

namespace {

constexpr unsigned PAD_IDX = 0xffffffffu;

__device__ __forceinline__ unsigned long long order_key(float f) { return (unsigned long long)fdt_float_key(f) << 32; }
__device__ __forceinline__ unsigned long long order_key(double d)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);       // larger double -> larger key; +NaN above +inf
}

template <typename T>
__global__ void k_gen_keys(const T *__restrict__ scores, int64_t n, int64_t n2, unsigned long long *__restrict__ keys, unsigned *__restrict__ idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    keys[i] = i < n ? order_key(scores[i]) : 0ull;
    idx[i] = i < n ? (unsigned)i : PAD_IDX;
}

// a sorts before b (descending): real entries before padding, larger key first, higher index first among equal keys
__device__ __forceinline__ bool sorts_before(unsigned long long ka, unsigned ia, unsigned long long kb, unsigned ib)
{
    if (ia == PAD_IDX || ib == PAD_IDX) return ib == PAD_IDX && ia != PAD_IDX;
    return ka > kb || (ka == kb && ia > ib);
}

__global__ void k_bitonic_step(unsigned long long *__restrict__ keys, unsigned *__restrict__ idx, int64_t n2, int64_t j, int64_t k)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const int64_t p = i ^ j;
    if (p <= i) return;
    const unsigned long long ka = keys[i], kb = keys[p];
    const unsigned ia = idx[i], ib = idx[p];
    const bool first_half = (i & k) == 0;                      // this pair sorts "descending" (our order) in the first half
    const bool swap = first_half ? sorts_before(kb, ib, ka, ia) : sorts_before(ka, ia, kb, ib);
    if (swap) { keys[i] = kb; keys[p] = ka; idx[i] = ib; idx[p] = ia; }
}

template <typename T> __device__ __forceinline__ T tmax(T a, T b);
template <> __device__ __forceinline__ float tmax<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ double tmax<double>(double a, double b) { return fmax(a, b); }
template <typename T> __device__ __forceinline__ T tmin(T a, T b);
template <> __device__ __forceinline__ float tmin<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double tmin<double>(double a, double b) { return fmin(a, b); }

template <typename T>
__device__ __forceinline__ T box_area(const T *b, const int variant)
{
    if (variant & FDT_NMS_PLUS1) return ((b[2] - b[0]) + (T)1) * ((b[3] - b[1]) + (T)1);
    return (b[2] - b[0]) * (b[3] - b[1]);
}

// "i (higher score, kept) suppresses j": box_utils.py:322-339 for variant 0 (union = (area_j - inter) + area_i, survive iff
// IoU < overlap is false for NaN -> suppressed); the sibling rules exactly as fdt_suppresses_v in detect.cu.
template <typename T>
__device__ __forceinline__ bool suppresses_t(const T *bi, const T ai, const T *bj, const T aj, const T thr, const int variant)
{
    const T xx1 = tmax(bi[0], bj[0]), yy1 = tmax(bi[1], bj[1]);
    const T xx2 = tmin(bi[2], bj[2]), yy2 = tmin(bi[3], bj[3]);
    T dw = xx2 - xx1, dh = yy2 - yy1;
    if (variant & FDT_NMS_PLUS1) { dw += (T)1; dh += (T)1; }
    const T inter = tmax((T)0, dw) * tmax((T)0, dh);
    T den;
    if (variant & FDT_NMS_MINIMUM) den = (ai != ai || aj != aj) ? (T)NAN : tmin(ai, aj);
    else if (variant & FDT_NMS_SUMFIRST) den = (ai + aj) - inter;
    else den = (aj - inter) + ai;
    const T ovr = inter / den;
    return (variant & FDT_NMS_LE) ? !(ovr <= thr) : !(ovr < thr);
}

// one block of 64 threads per (row tile, column tile) with column tile >= row tile; thread = one row
template <typename T>
__global__ void __launch_bounds__(64)
k_gen_mask(const T *__restrict__ boxes, const unsigned *__restrict__ order, int k, T thr, int variant, unsigned long long *__restrict__ mask)
{
    const int rt = blockIdx.y, ct = blockIdx.x;
    if (ct < rt) return;
    __shared__ T s_box[64][4];
    __shared__ T s_area[64];
    const int W = (k + 63) / 64;
    const int cj = ct * 64 + threadIdx.x;
    if (cj < k) {
        const T *b = boxes + 4 * (int64_t)order[cj];
        s_box[threadIdx.x][0] = b[0]; s_box[threadIdx.x][1] = b[1]; s_box[threadIdx.x][2] = b[2]; s_box[threadIdx.x][3] = b[3];
        s_area[threadIdx.x] = box_area(b, variant);
    }
    __syncthreads();
    const int i = rt * 64 + threadIdx.x;
    if (i >= k) return;
    const T *bsrc = boxes + 4 * (int64_t)order[i];
    const T bi[4] = {bsrc[0], bsrc[1], bsrc[2], bsrc[3]};
    const T ai = box_area(bi, variant);
    unsigned long long bits = 0ull;
    const int ncol = min(64, k - ct * 64);
    for (int c = (ct == rt ? threadIdx.x + 1 : 0); c < ncol; ++c)
        if (suppresses_t<T>(bi, ai, s_box[c], s_area[c], thr, variant)) bits |= 1ull << c;
    mask[(int64_t)i * W + ct] = bits;
}

__global__ void __launch_bounds__(1024)
k_gen_reduce(const unsigned long long *__restrict__ mask, const unsigned *__restrict__ order, int k, int64_t n_out,
             unsigned long long *__restrict__ removed, int64_t *__restrict__ keep, int64_t *__restrict__ count)
{
    const int W = (k + 63) / 64;
    for (int w = threadIdx.x; w < W; w += blockDim.x) removed[w] = 0ull;
    for (int64_t t = threadIdx.x; t < n_out; t += blockDim.x) keep[t] = 0;       // box_utils.py:289 zero-initialised
    __shared__ int s_kept;
    if (threadIdx.x == 0) s_kept = 0;
    __syncthreads();
    for (int i = 0; i < k; ++i) {
        const bool dead = (removed[i >> 6] >> (i & 63)) & 1ull;                  // uniform: every thread reads the same word
        __syncthreads();
        if (!dead) {
            if (threadIdx.x == 0) { keep[s_kept] = (int64_t)order[i]; ++s_kept; }
            const unsigned long long *row = mask + (int64_t)i * W;
            for (int w = (i >> 6) + threadIdx.x; w < W; w += blockDim.x) removed[w] |= row[w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = s_kept;
}

struct GenWs { unsigned long long *keys; unsigned *idx; unsigned long long *mask; unsigned long long *removed; size_t bytes; };
GenWs plan_gen_ws(void *ws, int64_t n)
{
    GenWs g;
    int64_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    const int64_t W = (n + 63) / 64;
    char *p = (char *)ws;
    size_t o = 0;
    g.keys = (unsigned long long *)(p + o); o += fdt_align256((size_t)n2 * 8);
    g.idx = (unsigned *)(p + o); o += fdt_align256((size_t)n2 * 4);
    g.removed = (unsigned long long *)(p + o); o += fdt_align256((size_t)W * 8);
    g.mask = (unsigned long long *)(p + o); o += fdt_align256((size_t)n * (size_t)W * 8);
    g.bytes = o;
    return g;
}

}  // namespace

size_t fdt_nms_generic_workspace_bytes(int64_t n) { return plan_gen_ws(nullptr, n > 0 ? n : 1).bytes; }

template <typename T>
int fdt_nms_generic(const T *boxes, const T *scores, int64_t n, T thresh, int64_t top_k, int variant,
                    int64_t *keep, int64_t *count, void *ws, size_t ws_bytes, cudaStream_t st)
{
    FDT_REQUIRE(n >= 1 && n <= FDT_MAX_NMS_GENERIC, FDT_E_UNSUPPORTED, "nms: n=%lld outside [1,%d] for the mask formulation", (long long)n, FDT_MAX_NMS_GENERIC);
    GenWs g = plan_gen_ws(ws, n);
    FDT_REQUIRE(ws_bytes >= g.bytes, FDT_E_WORKSPACE, "nms: workspace %zu < %zu bytes", ws_bytes, g.bytes);
    int64_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    const unsigned blocks = (unsigned)((n2 + 255) / 256);
    k_gen_keys<T><<<blocks, 256, 0, st>>>(scores, n, n2, g.keys, g.idx);
    FDT_LAUNCH_CHECK();
    for (int64_t k = 2; k <= n2; k <<= 1)
        for (int64_t j = k >> 1; j > 0; j >>= 1) {
            k_bitonic_step<<<blocks, 256, 0, st>>>(g.keys, g.idx, n2, j, k);
            FDT_LAUNCH_CHECK();
        }
    const int kk = (int)((top_k <= 0 || top_k > n) ? n : top_k);                 // idx[-top_k:]; idx[-0:] is the whole list
    const int tiles = (kk + 63) / 64;
    k_gen_mask<T><<<dim3((unsigned)tiles, (unsigned)tiles), 64, 0, st>>>(boxes, g.idx, kk, thresh, variant, g.mask);
    FDT_LAUNCH_CHECK();
    k_gen_reduce<<<1, 1024, 0, st>>>(g.mask, g.idx, kk, n, g.removed, keep, count);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}
template int fdt_nms_generic<float>(const float *, const float *, int64_t, float, int64_t, int, int64_t *, int64_t *, void *, size_t, cudaStream_t);
template int fdt_nms_generic<double>(const double *, const double *, int64_t, double, int64_t, int, int64_t *, int64_t *, void *, size_t, cudaStream_t);

// ---- float64 sibling NMS (MTCNN/mtcnn/core/utils.py:62-113 in the dtype of its float64 `dets`; core/nms.py:4-40; FaceBoxes nms_np)
FDT_API size_t fdt_nms_f64_workspace_bytes(int64_t n) { return fdt_nms_generic_workspace_bytes(n); }

FDT_API int fdt_nms_variant_f64(const double *boxes, const double *scores, int64_t n, double thresh, int variant,
                                int64_t *keep, int64_t *count, void *ws, size_t ws_bytes, fdt_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    FDT_REQUIRE(variant >= 0 && variant < 16, FDT_E_INVALID, "fdt_nms_variant_f64: unknown variant flags %d", variant);
    FDT_REQUIRE(n >= 0 && count != nullptr, FDT_E_INVALID, "fdt_nms_variant_f64: bad arguments");
    if (n == 0) { FDT_CUDA(cudaMemsetAsync(count, 0, sizeof(int64_t), st)); return FDT_OK; }
    FDT_REQUIRE(boxes && scores && keep && ws && fdt_aligned(ws, 256), FDT_E_INVALID, "fdt_nms_variant_f64: null / misaligned pointer");
    return fdt_nms_generic<double>(boxes, scores, n, thresh, 0, variant, keep, count, ws, ws_bytes, st);
}
