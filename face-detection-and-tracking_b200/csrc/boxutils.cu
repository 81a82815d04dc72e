// Elementwise / pairwise box math of layers/box_utils.py and utils/calc_performance.py, plus PriorBox.
#include "fdt_common.cuh"

namespace {

constexpr int EW_THREADS = 256;
inline unsigned ew_blocks(int64_t n) { return (unsigned)((n + EW_THREADS - 1) / EW_THREADS); }

// prior_box.py:28-44 -- fp64 arithmetic in python operand order, one rounding to fp32
struct PriorParams {
    double width, height, stride, box;
    int n_scales, n_ar, f_w, f_h;
    double box_scale[8];
    double sqrt_ar[8];
};
__global__ void k_priorbox(const PriorParams P, float4 *__restrict__ out)
{
    const int per_cell = P.n_scales * (1 + P.n_ar);
    const int64_t total = (int64_t)P.f_h * P.f_w * per_cell;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int v = (int)(t % per_cell);
    int64_t cell = t / per_cell;
    int j = (int)(cell % P.f_w), i = (int)(cell / P.f_w);
    int s = v / (1 + P.n_ar), a = v % (1 + P.n_ar);
    double cx = (j + 0.5) * P.stride / P.width;            // :34
    double cy = (i + 0.5) * P.stride / P.height;           // :35
    double sx = P.box * P.box_scale[s] / P.width;          // :36
    double sy = P.box * P.box_scale[s] / P.height;         // :37
    if (a > 0) { sx = sx / P.sqrt_ar[a - 1]; sy = sy * P.sqrt_ar[a - 1]; }   // :41
    out[t] = make_float4((float)cx, (float)cy, (float)sx, (float)sy);
}

__global__ void k_point_form(const float4 *__restrict__ b, int64_t n, float4 *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = b[i];
    float hw = p.z / 2.0f, hh = p.w / 2.0f;
    out[i] = make_float4(p.x - hw, p.y - hh, p.x + hw, p.y + hh);          // box_utils.py:15-16
}
__global__ void k_center_size(const float4 *__restrict__ b, int64_t n, float4 *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = b[i];
    out[i] = make_float4((p.z + p.x) / 2.0f, (p.w + p.y) / 2.0f, p.z - p.x, p.w - p.y);   // box_utils.py:27-28
}
__global__ void k_encode(const float4 *__restrict__ m, const float4 *__restrict__ p, int64_t n, float v0, float v1, float4 *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fdt_encode1(m[i], p[i], v0, v1);
}
__global__ void k_decode(const float4 *__restrict__ l, const float4 *__restrict__ p, int64_t n, float v0, float v1, float4 *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fdt_decode1(l[i], p[i], v0, v1);
}

// torch.min/max and np.minimum/maximum propagate NaN; fminf/fmaxf do not
template <typename T> __device__ __forceinline__ T nmin(T a, T b) { return (a != a || b != b) ? (T)NAN : (a < b ? a : b); }
template <typename T> __device__ __forceinline__ T nmax(T a, T b) { return (a != a || b != b) ? (T)NAN : (a > b ? a : b); }

// box_utils.py:58-66 / :94-100 and calc_performance.py:20-31 / :65-74 ; out[A,B], one thread per pair,
// B (priors) is the fast axis so box_b loads and the store are coalesced, box_a is a broadcast.
template <typename T, bool IOU>
__global__ void k_pairwise(const T *__restrict__ a, int64_t A, const T *__restrict__ b, int64_t B, T *__restrict__ out)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = blockIdx.y;
    if (j >= B) return;
    const T ax1 = a[4 * i], ay1 = a[4 * i + 1], ax2 = a[4 * i + 2], ay2 = a[4 * i + 3];
    const T bx1 = b[4 * j], by1 = b[4 * j + 1], bx2 = b[4 * j + 2], by2 = b[4 * j + 3];
    T w = nmin(ax2, bx2) - nmax(ax1, bx1);
    T h = nmin(ay2, by2) - nmax(ay1, by1);
    w = nmax(w, (T)0); h = nmax(h, (T)0);
    T inter = w * h;
    if (IOU) {
        T area_a = (ax2 - ax1) * (ay2 - ay1);
        T area_b = (bx2 - bx1) * (by2 - by1);
        T uni = area_a + area_b - inter;
        out[i * B + j] = inter / uni;
    } else {
        out[i * B + j] = inter;
    }
}

// box_utils.py:261-269: global max, then log(sum(exp(x - max))) + max per row
__global__ void k_global_max(const float *__restrict__ x, int64_t n, unsigned *__restrict__ gmax_key)
{
    float m = -INFINITY;
    bool nan = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = x[i];
        nan |= (v != v);
        m = fmaxf(m, v);
    }
    if (nan) m = NAN;
    unsigned k = fdt_float_key(m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k = max(k, __shfl_xor_sync(0xffffffffu, k, o));
    if ((threadIdx.x & 31) == 0) atomicMax(gmax_key, k);
}
__global__ void k_lse_rows(const float *__restrict__ x, int64_t R, int C, const unsigned *__restrict__ gmax_key, float *__restrict__ out)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float xmax = fdt_key_float(*gmax_key);
    float s = 0.0f;
    for (int c = 0; c < C; ++c) s += fdt_expf_cr(x[r * C + c] - xmax);
    out[r] = fdt_logf_cr(s) + xmax;
}

}  // namespace

FDT_API int fdt_priorbox(double width, double height, double stride, double box,
                         int n_scales, const double *box_scale_h, int n_ar, const double *sqrt_ar_h,
                         int f_w, int f_h, float *out, fdt_stream_t stream)
{
    FDT_REQUIRE(n_scales >= 0 && n_scales <= 8 && n_ar >= 0 && n_ar <= 8, FDT_E_UNSUPPORTED,
                "fdt_priorbox: n_scales=%d / n_ar=%d outside [0,8]", n_scales, n_ar);
    FDT_REQUIRE(f_w >= 0 && f_h >= 0, FDT_E_INVALID, "fdt_priorbox: negative feature-map size");
    int64_t total = (int64_t)f_h * f_w * n_scales * (1 + n_ar);
    if (total == 0) return FDT_OK;
    FDT_REQUIRE(out && fdt_aligned(out, 16), FDT_E_INVALID, "fdt_priorbox: out null or not 16-byte aligned");
    FDT_REQUIRE((n_scales == 0 || box_scale_h) && (n_ar == 0 || sqrt_ar_h), FDT_E_INVALID, "fdt_priorbox: null scale arrays");
    PriorParams P{};
    P.width = width; P.height = height; P.stride = stride; P.box = box;
    P.n_scales = n_scales; P.n_ar = n_ar; P.f_w = f_w; P.f_h = f_h;
    for (int s = 0; s < n_scales; ++s) P.box_scale[s] = box_scale_h[s];
    for (int a = 0; a < n_ar; ++a) P.sqrt_ar[a] = sqrt_ar_h[a];
    k_priorbox<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(P, (float4 *)out);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}

#define FDT_EW_PROLOGUE(name, ...)                                                                         \
    FDT_REQUIRE(n >= 0, FDT_E_INVALID, name ": negative n");                                               \
    if (n == 0) return FDT_OK;                                                                             \
    { const void *ptrs_[] = {__VA_ARGS__};                                                                 \
      for (const void *q : ptrs_) FDT_REQUIRE(q && fdt_aligned(q, 16), FDT_E_INVALID, name ": null or not 16-byte aligned pointer"); }

FDT_API int fdt_point_form(const float *boxes, int64_t n, float *out, fdt_stream_t stream)
{
    FDT_EW_PROLOGUE("fdt_point_form", boxes, out)
    k_point_form<<<ew_blocks(n), EW_THREADS, 0, (cudaStream_t)stream>>>((const float4 *)boxes, n, (float4 *)out);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}
FDT_API int fdt_center_size(const float *boxes, int64_t n, float *out, fdt_stream_t stream)
{
    FDT_EW_PROLOGUE("fdt_center_size", boxes, out)
    k_center_size<<<ew_blocks(n), EW_THREADS, 0, (cudaStream_t)stream>>>((const float4 *)boxes, n, (float4 *)out);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}
FDT_API int fdt_encode(const float *matched, const float *priors, int64_t n, float var0, float var1, float *out, fdt_stream_t stream)
{
    FDT_EW_PROLOGUE("fdt_encode", matched, priors, out)
    k_encode<<<ew_blocks(n), EW_THREADS, 0, (cudaStream_t)stream>>>((const float4 *)matched, (const float4 *)priors, n, var0, var1, (float4 *)out);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}
FDT_API int fdt_decode(const float *loc, const float *priors, int64_t n, float var0, float var1, float *out, fdt_stream_t stream)
{
    FDT_EW_PROLOGUE("fdt_decode", loc, priors, out)
    k_decode<<<ew_blocks(n), EW_THREADS, 0, (cudaStream_t)stream>>>((const float4 *)loc, (const float4 *)priors, n, var0, var1, (float4 *)out);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}

template <typename T, bool IOU>
static int pairwise(const char *name, const T *a, int64_t A, const T *b, int64_t B, T *out, fdt_stream_t stream)
{
    FDT_REQUIRE(A >= 0 && B >= 0 && A < 65536, FDT_E_INVALID, "%s: bad sizes A=%lld (max 65535) B=%lld", name, (long long)A, (long long)B);
    if (A == 0 || B == 0) return FDT_OK;
    FDT_REQUIRE(a && b && out, FDT_E_INVALID, "%s: null pointer", name);
    dim3 g(ew_blocks(B), (unsigned)A);
    k_pairwise<T, IOU><<<g, EW_THREADS, 0, (cudaStream_t)stream>>>(a, A, b, B, out);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}
FDT_API int fdt_intersect(const float *a, int64_t A, const float *b, int64_t B, float *out, fdt_stream_t s) { return pairwise<float, false>("fdt_intersect", a, A, b, B, out, s); }
FDT_API int fdt_calculate_iou(const float *a, int64_t A, const float *b, int64_t B, float *out, fdt_stream_t s) { return pairwise<float, true>("fdt_calculate_iou", a, A, b, B, out, s); }
FDT_API int fdt_calculate_iou_f64(const double *a, int64_t A, const double *b, int64_t B, double *out, fdt_stream_t s) { return pairwise<double, true>("fdt_calculate_iou_f64", a, A, b, B, out, s); }

FDT_API int fdt_log_sum_exp(const float *x, int64_t R, int C, float *out, void *ws, size_t ws_bytes, fdt_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    FDT_REQUIRE(R >= 0 && C >= 1, FDT_E_INVALID, "fdt_log_sum_exp: bad sizes");
    if (R == 0) return FDT_OK;
    FDT_REQUIRE(x && out && ws && ws_bytes >= 256, FDT_E_WORKSPACE, "fdt_log_sum_exp: needs a 256-byte workspace");
    FDT_CUDA(cudaMemsetAsync(ws, 0, 4, st));
    int64_t n = R * C;
    unsigned blocks = (unsigned)((n + EW_THREADS * 8 - 1) / (EW_THREADS * 8));
    if (blocks > FDT_NUM_SMS * 8) blocks = FDT_NUM_SMS * 8;
    k_global_max<<<blocks, EW_THREADS, 0, st>>>(x, n, (unsigned *)ws);
    FDT_LAUNCH_CHECK();
    k_lse_rows<<<ew_blocks(R), EW_THREADS, 0, st>>>(x, R, C, (const unsigned *)ws, out);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}
