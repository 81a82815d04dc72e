// Detect (layers/functions/detection.py:34-84) and nms (layers/box_utils.py:275-340) for sm_100a.
//
// Two launches per call, no host synchronisation:
//   K2  k_threshold_compact : streams conf once (HBM-bound), strict `score > conf_thresh`, warp-ballot +
//                             block-aggregated compaction into 64-bit keys (score_key << 32 | prior).
//   K3  k_sort_nms          : one CTA per (image, class) list.  Radix-select of the top nms_top_k keys when
//                             the list exceeds the shared-memory sort capacity, bitonic sort in shared
//                             memory, decode of the selected priors only (loc/priors gathered through L2),
//                             then greedy NMS evaluated LAZILY in chunks of 64 sorted candidates:
//                               phase A: chunk x kept-so-far IoU tests (16 threads per candidate),
//                               phase B: 64x64 intra-chunk suppression bitmask (warp ballots, boxes staged
//                                        in shared memory) + warp-serial mask reduction,
//                             stopping as soon as top_k boxes are kept (Detect reads only keep[:top_k],
//                             detection.py:80-81, so the output is identical to running to completion).
//
// Tie rule (unspecified in the reference, torch.sort is unstable): keys are unique, descending key order =
// descending score, higher prior index first among equal scores.
#include "fdt_common.cuh"

namespace {

constexpr int K2_THREADS = 256;
constexpr int K2_PER_THREAD = 8;
constexpr int K2_TILE = K2_THREADS * K2_PER_THREAD;

constexpr int K3_THREADS = 1024;
constexpr int SORT_CAP = FDT_MAX_NMS_TOP_K;     // 8192 keys = 64 KB
constexpr int CHUNK = 64;
constexpr int GROUP = K3_THREADS / CHUNK;       // 16 threads cooperate on one candidate in phase A

enum { MODE_DETECT = 0, MODE_NMS = 1 };

// ------------------------------------------------------------------------------------------- K2
template <bool C2>
__global__ void __launch_bounds__(K2_THREADS)
k_threshold_compact(const float *__restrict__ conf, int64_t N, int C, float thr,
                    int32_t *__restrict__ counters, uint64_t *__restrict__ keys)
{
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * K2_TILE;
    const float *cb = conf + (int64_t)b * N * C;
    __shared__ int s_warp[K2_THREADS / 32];
    __shared__ int s_base;

    for (int cl = 1; cl < C; ++cl) {
        float sc[K2_PER_THREAD];
        unsigned bal[K2_PER_THREAD];
        int wtotal = 0;
#pragma unroll
        for (int u = 0; u < K2_PER_THREAD; ++u) {
            int64_t p = base + u * K2_THREADS + tid;
            float s = 0.0f;
            bool in = p < N;
            if (in) {
                if (C2) s = __ldg(reinterpret_cast<const float2 *>(cb) + p).y;
                else    s = __ldg(cb + p * C + cl);
            }
            sc[u] = s;
            bal[u] = __ballot_sync(0xffffffffu, in && s > thr);      // detection.py:64 strict gt
            wtotal += __popc(bal[u]);
        }
        if (lane == 0) s_warp[warp] = wtotal;
        __syncthreads();
        const int list = b * (C - 1) + (cl - 1);
        if (tid == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < K2_THREADS / 32; ++w) { int c = s_warp[w]; s_warp[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(&counters[list], tot) : 0;
        }
        __syncthreads();
        int off = s_base + s_warp[warp];
        uint64_t *kl = keys + (int64_t)list * N;
#pragma unroll
        for (int u = 0; u < K2_PER_THREAD; ++u) {
            if ((bal[u] >> lane) & 1u) {
                int64_t p = base + u * K2_THREADS + tid;
                kl[off + __popc(bal[u] & ((1u << lane) - 1u))] = ((uint64_t)fdt_float_key(sc[u]) << 32) | (uint32_t)p;
            }
            off += __popc(bal[u]);
        }
        __syncthreads();
    }
}

__global__ void k_build_keys(const float *__restrict__ scores, int64_t n, uint64_t *__restrict__ keys)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((uint64_t)fdt_float_key(scores[i]) << 32) | (uint32_t)i;
}

// ------------------------------------------------------------------------------------------- K3
struct SortNmsParams {
    const uint64_t *keys;       // [lists, key_stride]
    const int32_t *counters;    // [lists] (MODE_DETECT)
    int64_t key_stride;
    const float *loc;           // [B,N,4]  (MODE_DETECT)
    const float *priors;        // [N,4]    (MODE_DETECT)
    const float *boxes;         // [n,4]    (MODE_NMS)
    int64_t N;
    int C;
    int nms_top_k;              // candidates that enter NMS
    int max_keep;               // stop after this many kept
    int top_k;                  // output rows (MODE_DETECT)
    float nms_thresh, v0, v1;
    float *out;                 // [B,C,top_k,5]
    int32_t *counts;            // [B,C] or null
    int64_t *kept_prior;        // [B,C,top_k] or null
    int64_t *keep;              // [n] (MODE_NMS)
    int64_t *count_out;         // [1] (MODE_NMS)
    int64_t n;                  // MODE_NMS list length
    int off_cbox, off_kbox, off_karea, off_kpos;   // byte offsets into dynamic shared memory
};

// "i (kept, higher score) suppresses j": box_utils.py:322-339, union = (area_j - inter) + area_i,
// survive iff IoU < overlap, so NaN suppresses.
__device__ __forceinline__ bool fdt_suppresses(const float4 bi, const float area_i, const float4 bj, const float area_j,
                                               const float thr)
{
    float xx1 = fmaxf(bj.x, bi.x), yy1 = fmaxf(bj.y, bi.y);
    float xx2 = fminf(bj.z, bi.z), yy2 = fminf(bj.w, bi.w);
    float w = fmaxf(xx2 - xx1, 0.0f), h = fmaxf(yy2 - yy1, 0.0f);
    float inter = w * h;
    float uni = (area_j - inter) + area_i;
    if (inter > 0.0f) return !(inter / uni < thr);
    return !(uni > 0.0f) && !(uni < 0.0f);          // 0/uni is NaN iff uni is 0 or NaN
}

template <int MODE, bool KEPT_COPY>
__global__ void __launch_bounds__(K3_THREADS, 1)
k_sort_nms(const SortNmsParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *skeys = reinterpret_cast<uint64_t *>(smem);
    float4 *cbox = reinterpret_cast<float4 *>(smem + P.off_cbox);
    float4 *kbox = reinterpret_cast<float4 *>(smem + P.off_kbox);
    float *karea = reinterpret_cast<float *>(smem + P.off_karea);
    int32_t *kpos = reinterpret_cast<int32_t *>(smem + P.off_kpos);
    __shared__ uint64_t s_mask[CHUNK];
    __shared__ uint64_t s_keptbits;
    __shared__ unsigned char s_flag[CHUNK];
    __shared__ int s_hist[256];
    __shared__ int s_sel[3];
    __shared__ int s_cnt;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int list = blockIdx.x;
    const int b = (MODE == MODE_DETECT) ? list / (P.C - 1) : 0;
    const int cl = (MODE == MODE_DETECT) ? 1 + list % (P.C - 1) : 0;
    const uint64_t *gkeys = P.keys + (int64_t)list * P.key_stride;

    int n_c = (MODE == MODE_DETECT) ? P.counters[list] : (int)P.n;
    if (MODE == MODE_DETECT && n_c == 1) n_c = 0;            // detection.py:66-72: one candidate -> `continue`
    int n_sorted;

    // ---------------- select: top nms_top_k keys into shared memory
    if (n_c <= SORT_CAP) {
        for (int i = tid; i < n_c; i += K3_THREADS) skeys[i] = gkeys[i];
        n_sorted = n_c;
    } else {
        // MSB-first radix select (8 bits per pass) of the nms_top_k-th largest key over the global list
        uint64_t prefix = 0, pmask = 0;
        int need = P.nms_top_k;
        for (int shift = 56; shift >= 0; shift -= 8) {
            if (tid < 256) s_hist[tid] = 0;
            __syncthreads();
            for (int base = 0; base < n_c; base += K3_THREADS) {
                int i = base + tid;
                int d = 256;
                if (i < n_c) {
                    uint64_t key = gkeys[i];
                    if ((key & pmask) == prefix) d = (int)((key >> shift) & 0xff);
                }
                unsigned peers = __match_any_sync(0xffffffffu, d);
                if (d < 256 && lane == __ffs(peers) - 1) atomicAdd(&s_hist[d], __popc(peers));
            }
            __syncthreads();
            if (warp == 0) {
                int c = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) c += s_hist[255 - 8 * lane - q];
                int cum = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += v; }
                unsigned hit = __ballot_sync(0xffffffffu, cum >= need);
                int first = __ffs(hit) - 1;
                if (lane == first) {
                    int rem = need - (cum - c);
                    for (int q = 0; q < 8; ++q) {
                        int hcount = s_hist[255 - 8 * lane - q];
                        if (hcount >= rem) { s_sel[0] = 255 - 8 * lane - q; s_sel[1] = rem; s_sel[2] = hcount; break; }
                        rem -= hcount;
                    }
                }
            }
            __syncthreads();
            prefix |= (uint64_t)s_sel[0] << shift;
            pmask |= (uint64_t)0xff << shift;
            need = s_sel[1];
            const bool done = (s_sel[2] == need);
            __syncthreads();
            if (done) break;
        }
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (int base = 0; base < n_c; base += K3_THREADS) {
            int i = base + tid;
            uint64_t key = 0;
            bool take = false;
            if (i < n_c) { key = gkeys[i]; take = key >= prefix; }
            unsigned bal = __ballot_sync(0xffffffffu, take);
            int wbase = 0;
            if (lane == 0 && bal) wbase = atomicAdd(&s_cnt, __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (take) skeys[wbase + __popc(bal & ((1u << lane) - 1u))] = key;
        }
        __syncthreads();
        n_sorted = s_cnt;           // == nms_top_k (keys are unique)
    }

    // ---------------- bitonic sort, descending, padded with the minimum key
    int P2 = 2;
    while (P2 < n_sorted) P2 <<= 1;
    for (int i = n_sorted + tid; i < P2; i += K3_THREADS) skeys[i] = 0;
    __syncthreads();
    if (n_sorted > 1) {
        for (int size = 2; size <= P2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < (P2 >> 1); t += K3_THREADS) {
                    int i = 2 * t - (t & (stride - 1));
                    int j = i + stride;
                    uint64_t a = skeys[i], c = skeys[j];
                    bool desc = (i & size) == 0;
                    if ((a < c) == desc) { skeys[i] = c; skeys[j] = a; }
                }
                __syncthreads();
            }
        }
    }
    const int k = min(n_sorted, P.nms_top_k);                 // box_utils.py:299 idx[-top_k:]

    // ---------------- boxes of the selected candidates (decode only what NMS will look at)
    for (int j = tid; j < k; j += K3_THREADS) {
        uint32_t p = (uint32_t)skeys[j];
        float4 bx;
        if (MODE == MODE_DETECT) {
            float4 l = __ldg(reinterpret_cast<const float4 *>(P.loc) + ((int64_t)b * P.N + p));
            float4 pr = __ldg(reinterpret_cast<const float4 *>(P.priors) + p);
            bx = fdt_decode1(l, pr, P.v0, P.v1);              // detection.py:55
        } else {
            bx = __ldg(reinterpret_cast<const float4 *>(P.boxes) + p);
        }
        cbox[j] = bx;
    }
    __syncthreads();

    // ---------------- lazy greedy NMS
    const float thr = P.nms_thresh;
    const int max_keep = P.max_keep;
    int nkept = 0;
    const int c = tid / GROUP, s = tid % GROUP;
    const unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << ((lane / GROUP) * GROUP));
    for (int pos = 0; pos < k && nkept < max_keep; pos += CHUNK) {
        // phase A: candidate j against every box kept so far
        {
            const int j = pos + c;
            const bool valid = j < k;
            const float4 bj = cbox[valid ? j : 0];
            const float aj = (bj.z - bj.x) * (bj.w - bj.y);                  // box_utils.py:296
            bool sup = false;
            const int nk = min(nkept, max_keep);
            for (int m0 = 0; m0 < nk; m0 += GROUP) {
                int m = m0 + s;
                if (valid && m < nk) {
                    float4 bi; float ai;
                    if (KEPT_COPY) { bi = kbox[m]; ai = karea[m]; }
                    else { bi = cbox[kpos[m]]; ai = (bi.z - bi.x) * (bi.w - bi.y); }
                    sup = fdt_suppresses(bi, ai, bj, aj, thr);
                }
                if (__any_sync(gmask, sup)) { sup = true; break; }
            }
            if (s == 0) s_flag[c] = valid && !sup;
        }
        __syncthreads();
        // phase B: 64x64 intra-chunk mask; warp w owns rows 2w, 2w+1, lanes own columns lane, lane+32
        {
            const int j0 = pos + lane, j1 = pos + lane + 32;
            const float4 b0 = cbox[j0 < k ? j0 : 0], b1 = cbox[j1 < k ? j1 : 0];
            const float a0 = (b0.z - b0.x) * (b0.w - b0.y), a1 = (b1.z - b1.x) * (b1.w - b1.y);
            const bool f0 = s_flag[lane], f1 = s_flag[lane + 32];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int i = 2 * warp + r;
                uint64_t word = 0;
                if (s_flag[i]) {                                            // warp-uniform
                    const float4 bi = cbox[pos + i];
                    const float ai = (bi.z - bi.x) * (bi.w - bi.y);
                    bool t0 = f0 && lane > i && fdt_suppresses(bi, ai, b0, a0, thr);
                    bool t1 = f1 && lane + 32 > i && fdt_suppresses(bi, ai, b1, a1, thr);
                    unsigned lo = __ballot_sync(0xffffffffu, t0), hi = __ballot_sync(0xffffffffu, t1);
                    word = (uint64_t)lo | ((uint64_t)hi << 32);
                }
                if (lane == 0) s_mask[i] = word;
            }
        }
        __syncthreads();
        // warp-serial mask reduction
        if (warp == 0) {
            unsigned lo = __ballot_sync(0xffffffffu, s_flag[lane]), hi = __ballot_sync(0xffffffffu, s_flag[lane + 32]);
            uint64_t remv = ~((uint64_t)lo | ((uint64_t)hi << 32));
            uint64_t keptbits = 0;
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                uint64_t mi = s_mask[i];
                if (!((remv >> i) & 1ull)) { keptbits |= 1ull << i; remv |= mi; }
            }
            if (lane == 0) s_keptbits = keptbits;
        }
        __syncthreads();
        const uint64_t kb = s_keptbits;
        if (tid < CHUNK && ((kb >> tid) & 1ull)) {
            int slot = nkept + __popcll(kb & ((1ull << tid) - 1ull));
            if (slot < max_keep) {
                kpos[slot] = pos + tid;
                if (KEPT_COPY) {
                    float4 bx = cbox[pos + tid];
                    kbox[slot] = bx;
                    karea[slot] = (bx.z - bx.x) * (bx.w - bx.y);
                }
            }
        }
        nkept += __popcll(kb);
        __syncthreads();
    }
    nkept = min(nkept, max_keep);

    // ---------------- outputs
    if (MODE == MODE_DETECT) {
        const int top_k = P.top_k;
        const int cnt = min(nkept, top_k);                                   // detection.py:80
        float *o = P.out + ((int64_t)(b * P.C + cl) * top_k) * 5;
        for (int t = tid; t < top_k * 5; t += K3_THREADS) {
            int r = t / 5, col = t - 5 * r;
            float v = 0.0f;
            if (r < cnt) {
                int ps = kpos[r];
                if (col == 0) v = fdt_key_float((uint32_t)(skeys[ps] >> 32));
                else {
                    const float *bx = reinterpret_cast<const float *>(cbox + ps);
                    v = bx[col - 1];
                }
            }
            o[t] = v;                                                        // detection.py:82
        }
        if (P.kept_prior) {
            int64_t *kp = P.kept_prior + (int64_t)(b * P.C + cl) * top_k;
            for (int r = tid; r < top_k; r += K3_THREADS) kp[r] = r < cnt ? (int64_t)(uint32_t)skeys[kpos[r]] : -1;
        }
        if (P.counts && tid == 0) P.counts[b * P.C + cl] = cnt;
        if (cl == 1) {                                                       // class-0 plane stays zero (:48, :63)
            float *o0 = P.out + ((int64_t)(b * P.C) * top_k) * 5;
            for (int t = tid; t < top_k * 5; t += K3_THREADS) o0[t] = 0.0f;
            if (P.kept_prior) {
                int64_t *kp = P.kept_prior + (int64_t)(b * P.C) * top_k;
                for (int r = tid; r < top_k; r += K3_THREADS) kp[r] = -1;
            }
            if (P.counts && tid == 0) P.counts[b * P.C] = 0;
        }
    } else {
        for (int64_t t = tid; t < P.n; t += K3_THREADS)
            P.keep[t] = t < nkept ? (int64_t)(uint32_t)skeys[kpos[t]] : 0;    // box_utils.py:289 zero-initialised
        if (tid == 0) *P.count_out = nkept;
    }
}

struct SmemPlan { int off_cbox, off_kbox, off_karea, off_kpos, total; };

SmemPlan plan_smem(int kcap, int max_keep, bool kept_copy)
{
    SmemPlan s;
    int off = SORT_CAP * 8;
    s.off_cbox = off; off += ((kcap * 16 + 15) / 16) * 16;
    s.off_kbox = off; if (kept_copy) off += max_keep * 16;
    s.off_karea = off; if (kept_copy) off += ((max_keep * 4 + 15) / 16) * 16;
    s.off_kpos = off; off += ((max_keep * 4 + 15) / 16) * 16;
    s.total = off;
    return s;
}

template <int MODE>
int launch_sort_nms(SortNmsParams &P, int lists, int kcap, cudaStream_t st)
{
    SmemPlan sp = plan_smem(kcap, P.max_keep, true);
    const int limit = FDT_SMEM_MAX - 2048;      // static shared memory of k_sort_nms
    bool kept_copy = sp.total <= limit;
    if (!kept_copy) sp = plan_smem(kcap, P.max_keep, false);
    FDT_REQUIRE(sp.total <= limit, FDT_E_UNSUPPORTED,
                "nms_top_k=%d / max_keep=%d need %d bytes of shared memory (limit %d)", kcap, P.max_keep, sp.total, limit);
    P.off_cbox = sp.off_cbox; P.off_kbox = sp.off_kbox; P.off_karea = sp.off_karea; P.off_kpos = sp.off_kpos;
    if (kept_copy) {
        FDT_CUDA(cudaFuncSetAttribute(k_sort_nms<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
        k_sort_nms<MODE, true><<<lists, K3_THREADS, sp.total, st>>>(P);
    } else {
        FDT_CUDA(cudaFuncSetAttribute(k_sort_nms<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
        k_sort_nms<MODE, false><<<lists, K3_THREADS, sp.total, st>>>(P);
    }
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}

}  // namespace

// =============================================================================================== C ABI
FDT_API size_t fdt_detect_workspace_bytes(int B, int64_t N, int C)
{
    if (B <= 0 || N <= 0 || C <= 1) return 256;
    size_t lists = (size_t)B * (size_t)(C - 1);
    return fdt_align256(lists * sizeof(int32_t)) + fdt_align256(lists * (size_t)N * sizeof(uint64_t));
}

static int detect_check_common(const char *who, int B, int64_t N, int C, const void *ws, size_t ws_bytes)
{
    FDT_REQUIRE(B >= 0 && N >= 0 && C >= 1, FDT_E_INVALID, "%s: bad sizes B=%d N=%lld C=%d", who, B, (long long)N, C);
    FDT_REQUIRE(N < (1ll << 31), FDT_E_UNSUPPORTED, "%s: N=%lld exceeds 2^31-1", who, (long long)N);
    if (B == 0 || C == 1 || N == 0) return FDT_OK;
    FDT_REQUIRE(ws && fdt_aligned(ws, 256), FDT_E_INVALID, "%s: workspace null or not 256-byte aligned", who);
    FDT_REQUIRE(ws_bytes >= fdt_detect_workspace_bytes(B, N, C), FDT_E_WORKSPACE,
                "%s: workspace %zu < %zu bytes", who, ws_bytes, fdt_detect_workspace_bytes(B, N, C));
    return FDT_OK;
}

FDT_API int fdt_detect_threshold_compact(const float *conf, int B, int64_t N, int C, float conf_thresh,
                                         void *ws, size_t ws_bytes, fdt_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int rc = detect_check_common("fdt_detect_threshold_compact", B, N, C, ws, ws_bytes);
    if (rc != FDT_OK) return rc;
    const int lists = B * (C - 1);
    if (lists == 0 || N == 0) return FDT_OK;
    FDT_REQUIRE(conf && fdt_aligned(conf, 8), FDT_E_INVALID, "fdt_detect_threshold_compact: conf null or not 8-byte aligned");
    int32_t *counters = (int32_t *)ws;
    uint64_t *keys = (uint64_t *)((char *)ws + fdt_align256((size_t)lists * sizeof(int32_t)));
    FDT_CUDA(cudaMemsetAsync(counters, 0, (size_t)lists * sizeof(int32_t), st));
    dim3 g2((unsigned)((N + K2_TILE - 1) / K2_TILE), (unsigned)B);
    if (C == 2) k_threshold_compact<true><<<g2, K2_THREADS, 0, st>>>(conf, N, C, conf_thresh, counters, keys);
    else        k_threshold_compact<false><<<g2, K2_THREADS, 0, st>>>(conf, N, C, conf_thresh, counters, keys);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}

FDT_API int fdt_detect_candidate_counts(const void *ws, int B, int C, int32_t *counts_out, fdt_stream_t stream)
{
    FDT_REQUIRE(ws && counts_out && B >= 0 && C >= 1, FDT_E_INVALID, "fdt_detect_candidate_counts: bad arguments");
    if (B * (C - 1) == 0) return FDT_OK;
    FDT_CUDA(cudaMemcpyAsync(counts_out, ws, (size_t)B * (C - 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return FDT_OK;
}

FDT_API int fdt_detect_sort_nms(const float *loc, const float *priors, int B, int64_t N, int C, int top_k, int nms_top_k,
                                float nms_thresh, float var0, float var1,
                                float *out, int32_t *counts, int64_t *kept_prior,
                                void *ws, size_t ws_bytes, fdt_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int rc = detect_check_common("fdt_detect_sort_nms", B, N, C, ws, ws_bytes);
    if (rc != FDT_OK) return rc;
    FDT_REQUIRE(top_k >= 1, FDT_E_INVALID, "fdt_detect_sort_nms: top_k=%d", top_k);
    FDT_REQUIRE(nms_top_k >= 1 && nms_top_k <= FDT_MAX_NMS_TOP_K, FDT_E_UNSUPPORTED,
                "fdt_detect: nms_top_k=%d outside [1,%d]", nms_top_k, FDT_MAX_NMS_TOP_K);
    FDT_REQUIRE(nms_thresh > 0.0f, FDT_E_INVALID, "fdt_detect: nms_thresh must be > 0 (detection.py:28-29)");
    if (B == 0) return FDT_OK;
    FDT_REQUIRE(out != nullptr, FDT_E_INVALID, "fdt_detect: out is null");
    const int lists = B * (C - 1);
    if (lists == 0 || N == 0) {
        FDT_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * C * top_k * 5, st));
        if (counts) FDT_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)B * C, st));
        if (kept_prior) FDT_CUDA(cudaMemsetAsync(kept_prior, 0xff, sizeof(int64_t) * (size_t)B * C * top_k, st));
        return FDT_OK;
    }
    FDT_REQUIRE(loc && priors && fdt_aligned(loc, 16) && fdt_aligned(priors, 16), FDT_E_INVALID,
                "fdt_detect: loc/priors null or not 16-byte aligned");
    int32_t *counters = (int32_t *)ws;
    uint64_t *keys = (uint64_t *)((char *)ws + fdt_align256((size_t)lists * sizeof(int32_t)));
    SortNmsParams P{};
    P.keys = keys; P.counters = counters; P.key_stride = N;
    P.loc = loc; P.priors = priors; P.N = N; P.C = C;
    P.nms_top_k = nms_top_k; P.max_keep = top_k < nms_top_k ? top_k : nms_top_k; P.top_k = top_k;
    P.nms_thresh = nms_thresh; P.v0 = var0; P.v1 = var1;
    P.out = out; P.counts = counts; P.kept_prior = kept_prior;
    int kcap = (int)((int64_t)nms_top_k < N ? nms_top_k : N);
    return launch_sort_nms<MODE_DETECT>(P, lists, kcap, st);
}

FDT_API int fdt_detect(const float *loc, const float *conf, const float *priors,
                       int B, int64_t N, int C, int top_k, int nms_top_k,
                       float conf_thresh, float nms_thresh, float var0, float var1,
                       float *out, int32_t *counts, int64_t *kept_prior,
                       void *ws, size_t ws_bytes, fdt_stream_t stream)
{
    int rc = fdt_detect_threshold_compact(conf, B, N, C, conf_thresh, ws, ws_bytes, stream);
    if (rc != FDT_OK) return rc;
    return fdt_detect_sort_nms(loc, priors, B, N, C, top_k, nms_top_k, nms_thresh, var0, var1,
                               out, counts, kept_prior, ws, ws_bytes, stream);
}

FDT_API size_t fdt_nms_workspace_bytes(int64_t n)
{
    return fdt_align256((size_t)(n > 0 ? n : 1) * sizeof(uint64_t));
}

FDT_API int fdt_nms(const float *boxes, const float *scores, int64_t n, float overlap, int64_t top_k,
                    int64_t *keep, int64_t *count, void *ws, size_t ws_bytes, fdt_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    FDT_REQUIRE(n >= 0 && n < (1ll << 31), FDT_E_INVALID, "fdt_nms: bad n=%lld", (long long)n);
    FDT_REQUIRE(count != nullptr, FDT_E_INVALID, "fdt_nms: count is null");
    if (n == 0) { FDT_CUDA(cudaMemsetAsync(count, 0, sizeof(int64_t), st)); return FDT_OK; }   // box_utils.py:290-291
    FDT_REQUIRE(boxes && scores && keep && ws, FDT_E_INVALID, "fdt_nms: null pointer argument");
    FDT_REQUIRE(fdt_aligned(boxes, 16) && fdt_aligned(ws, 256), FDT_E_INVALID, "fdt_nms: boxes need 16-byte, workspace 256-byte alignment");
    FDT_REQUIRE(ws_bytes >= fdt_nms_workspace_bytes(n), FDT_E_WORKSPACE, "fdt_nms: workspace %zu < %zu bytes", ws_bytes, fdt_nms_workspace_bytes(n));
    int64_t k = (top_k <= 0 || top_k > n) ? n : top_k;         // idx[-top_k:]; idx[-0:] is the whole list
    FDT_REQUIRE(k <= FDT_MAX_NMS_TOP_K, FDT_E_UNSUPPORTED, "fdt_nms: min(n, top_k)=%lld exceeds %d", (long long)k, FDT_MAX_NMS_TOP_K);
    uint64_t *keys = (uint64_t *)ws;
    k_build_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scores, n, keys);
    FDT_LAUNCH_CHECK();
    SortNmsParams P{};
    P.keys = keys; P.key_stride = n; P.boxes = boxes; P.n = n; P.N = n; P.C = 2;
    P.nms_top_k = (int)k; P.max_keep = (int)k; P.nms_thresh = overlap;
    P.keep = keep; P.count_out = count;
    return launch_sort_nms<MODE_NMS>(P, 1, (int)k, st);
}
