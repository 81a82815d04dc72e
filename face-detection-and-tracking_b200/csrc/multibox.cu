// MultiBoxLoss (layers/modules/multibox_loss.py:48-136): fused jaccard/argmax match + encode
// (layers/box_utils.py:103-210), per-prior confidence loss, hard-negative mining, loss reduction and
// the backward pass, for sm_100a.
//
//   k_match        one thread per (image, prior); GT boxes of the image staged through shared memory in
//                  tiles; IoU = inter / ((area_a + area_b) - inter) exactly as calculate_iou
//                  (box_utils.py:94-100), argmax with first-index ties (torch.max(0)), then label +
//                  encode written straight to loc_t / conf_t -- the [G,N] overlap matrix is never
//                  materialised.  Bipartite mode also reduces the best prior per GT (warp REDUX ->
//                  shared -> one 64-bit atomicMax per block and GT) and a second pass applies
//                  box_utils.py:150-154.
//   k_loss_prior   smooth-L1 over positives (:96-101) and the mining input loss_c (:104-110).
//   mining         neg = rank < num_neg of the descending sort (:112-116) == the num_neg largest 64-bit composites
//                  (loss key << 32 | ~prior): unique, so ties resolve to the lower prior index.  k_loss_prior (or
//                  k_mine_hist) histograms the top 12 bits chip-wide, k_mine_select (one CTA per image) finishes an MSB
//                  radix select 12 bits per row scan (normally 2 scans) -> one cutoff per image, k_mine_apply (chip-wide)
//                  writes the mask and accumulates CE over pos U neg (:119-128).
//   k_loss_final / k_multibox_backward.
#include "fdt_common.cuh"

namespace {

// Launch behind the previous kernel of the stream with programmatic stream serialization: the grid is scheduled while its
// predecessor drains and every kernel below starts with fdt_pdl_enter() (let the successor be scheduled, then wait for the
// predecessor's results).  The forward is a chain of six short kernels; the hand-overs are a sizeable part of it.
__device__ __forceinline__ void fdt_pdl_enter()
{
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

constexpr int M_THREADS = 256;
constexpr int M_WARPS = M_THREADS / 32;
constexpr int GT_TILE = 256;
constexpr int MINE_THREADS = 1024;
constexpr int MINE_BINS = 4096;             // 12-bit digits
constexpr int MINE_COLLECT = 2048;          // k_mine_select: largest level-0 bin finished by one collecting scan
constexpr int MINE_REPL = 16;               // level-0 histogram copies (by prior index) to spread same-address atomics

// key of a value for the GLOBAL maximum of conf (box_utils.py:268, x.max()): torch's max propagates NaN whatever its sign bit, the
// order-preserving key alone would rank a sign-bit NaN (x86's default QNaN) lowest and drop it.  Every NaN maps to the top key, which
// fdt_key_float turns back into a NaN.
__device__ __forceinline__ unsigned gmax_key_of(float v) { return v != v ? 0xffffffffu : fdt_float_key(v); }

// (a NaN loss ranks highest whatever its sign bit, as torch.sort does)
__device__ __forceinline__ unsigned long long mine_comp(float v, unsigned p) { return ((unsigned long long)(v != v ? 0xffffffffu : fdt_float_key(v)) << 32) | (unsigned)~p; }

// Dispatch order of the matchers: blocks are issued x-fastest, image by image, and the blocks of the coarse pyramid levels (the END of
// the prior array: a 512-pixel prior overlaps most GT boxes, so its warp walks the whole list) run several times longer than the rest
// -- in image order the last image's coarse tiles start last and the kernel ends with a dozen SMs finishing them alone (ncu: SMs busy
// 61 % of the kernel).  Remapped: linear block id -> image = id % B, tile = last - id / B, i.e. every image's last tile first.
struct MatchBlock { int b; int tile; };
__device__ __forceinline__ MatchBlock match_block()
{
    const int id = blockIdx.y * gridDim.x + blockIdx.x;
    MatchBlock m;
    m.b = id % (int)gridDim.y;
    m.tile = (int)gridDim.x - 1 - id / (int)gridDim.y;
    return m;
}

struct GtTile {
    float4 box[GT_TILE];
    float area[GT_TILE];
    int idx[GT_TILE];          // GT index inside the image (tiles are compacted, order preserved)
};

__device__ __forceinline__ float iou_match(const float4 a, const float area_a, const float4 pf, const float area_b)
{
    float w = fminf(a.z, pf.z) - fmaxf(a.x, pf.x);
    float h = fminf(a.w, pf.w) - fmaxf(a.y, pf.y);
    w = fmaxf(w, 0.0f); h = fmaxf(h, 0.0f);
    float inter = w * h;
    float uni = area_a + area_b - inter;               // box_utils.py:98
    // inter / uni, IEEE -- but only where it matters.  Most pairs do not overlap; the compiler would otherwise evaluate the
    // division speculatively for every pair (and 0 / x takes the slow-path subroutine: FCHK rejects a zero numerator).  The
    // volatile asm pins it inside the branch.
    if (inter > 0.0f || !(uni > 0.0f)) {               // overlapping, or 0/0, 0/negative, NaN: exactly what the reference computes
        float q;
        asm volatile("div.rn.f32 %0, %1, %2;" : "=f"(q) : "f"(inter), "f"(uni));
        return q;
    }
    return 0.0f;                                       // 0 / positive
}

// The default matcher only asks whether a pair's IoU beats the prior's running best (`v > best`, first index wins ties), so the IEEE
// division is skipped where it cannot: with uni > 0 and best * uni > 0 (hence best > 0, both normal numbers), inter < 0.999999 *
// RN(best * uni) implies inter / uni < best exactly (RN is within 2^-24 relative of the product), and rounding is monotone, so
// RN(inter / uni) <= best and the update would not happen.  Returns 0 in that case (0 > best is false as well).  `always` (GT 0, which
// seeds best with its value whatever it is) and every special case (uni <= 0, NaN, infinities on the wrong side) take the division.
__device__ __forceinline__ float iou_match_vs_best(const float4 a, const float area_a, const float4 pf, const float area_b,
                                                   const float best, const bool always)
{
    float w = fminf(a.z, pf.z) - fmaxf(a.x, pf.x);
    float h = fminf(a.w, pf.w) - fmaxf(a.y, pf.y);
    w = fmaxf(w, 0.0f); h = fmaxf(h, 0.0f);
    const float inter = w * h;
    const float uni = area_a + area_b - inter;         // box_utils.py:98
    if (inter > 0.0f || !(uni > 0.0f)) {
        const float pth = best * uni;
        const bool cannot_win = !always && uni > 0.0f && pth > 1e-30f && inter < pth * 0.999999f;
        if (!cannot_win) {
            float q;
            asm volatile("div.rn.f32 %0, %1, %2;" : "=f"(q) : "f"(inter), "f"(uni));
            return q;
        }
    }
    return 0.0f;
}

__device__ __forceinline__ void finalize_prior(const float *__restrict__ gt, int64_t g0, int idx, float ov, float thr,
                                               float4 pr, float v0, float v1, float4 *loc_t, int64_t *conf_t,
                                               int32_t *bti, float *bto, int64_t t, const bool encode_all)
{
    const float *row = gt + 5 * (g0 + idx);
    float4 m = make_float4(row[0], row[1], row[2], row[3]);
    float c = row[4] + 1.0f;                           // box_utils.py:205
    if (ov < thr) c = 0.0f;                            // :206
    const int64_t label = (int64_t)c;                  // :210 (float -> long)
    conf_t[t] = label;
    // :208 encodes every prior; the loss only ever reads the positives (multibox_loss.py:96-101), so the fused forward
    // (encode_all = false) spares the two fp64 logs and four divisions of the ~99 % background priors and stores zeros
    loc_t[t] = (encode_all || label > 0) ? fdt_encode1(m, pr, v0, v1) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (bti) bti[t] = idx;
    if (bto) bto[t] = ov;
}

// One thread per prior.  GT boxes are staged through shared memory in tiles of 256 and CULLED per block: a GT box whose
// clamped overlap with the bounding box of the block's 256 priors is empty has IoU exactly 0 with every one of them (fp32
// subtraction is monotone, so w_bb <= 0 implies w <= 0 for each prior), and a zero can never win `v > best` -- skipping it
// leaves best_truth_idx / best_truth_overlap bit-identical.  Consecutive priors are spatial neighbours (one or two feature
// map rows), so at the fine pyramid levels ~90 % of the GT boxes drop out.  GT 0 is never culled (it seeds the argmax,
// box_utils.py:197 returns index 0 when all overlaps are 0) and in bipartite mode block 0 culls nothing, so the first prior
// still wins an all-zero row of overlaps.max(1) (:136).
// Only the bipartite matcher (BIP = true) is instantiated: it needs the block-level structure for the per-GT best-prior
// reduction.  The default matcher is k_match_default below.
template <bool BIP>
__global__ void __launch_bounds__(M_THREADS)
k_match(const float4 *__restrict__ priors, const float *__restrict__ gt, const int64_t *__restrict__ gt_off,
        int64_t N, float thr, float v0, float v1,
        float4 *__restrict__ loc_t, int64_t *__restrict__ conf_t, int32_t *__restrict__ bti, float *__restrict__ bto,
        int32_t *__restrict__ tmp_idx, float *__restrict__ tmp_ov, unsigned long long *__restrict__ bestprior, const bool encode_all)
{
    fdt_pdl_enter();
    __shared__ GtTile tile;
    __shared__ unsigned long long s_best[BIP ? GT_TILE : 1][BIP ? M_WARPS : 1];
    __shared__ unsigned s_bb[4];
    __shared__ int s_wcnt[M_WARPS];
    const MatchBlock mb = match_block();
    const int b = mb.b, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p = (int64_t)mb.tile * M_THREADS + tid;
    const bool valid = p < N;
    const int64_t g0 = gt_off[b];
    const int G = (int)(gt_off[b + 1] - g0);
    const int64_t t = (int64_t)b * N + p;
    const float4 pr = priors[valid ? p : 0];
    if (G <= 0) {                                      // reference raises (Q3); defined: all background
        if (valid) {
            conf_t[t] = 0; loc_t[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bti) bti[t] = 0;
            if (bto) bto[t] = 0.0f;
        }
        return;
    }
    const float hw = pr.z / 2.0f, hh = pr.w / 2.0f;    // point_form, box_utils.py:15-16
    const float4 pf = make_float4(pr.x - hw, pr.y - hh, pr.x + hw, pr.y + hh);
    const float area_b = (pf.z - pf.x) * (pf.w - pf.y);
    // bounding box of the block's priors (order-preserving integer keys make float min/max an integer atomic)
    if (tid < 4) s_bb[tid] = tid < 2 ? 0xffffffffu : 0u;
    __syncthreads();
    {
        unsigned k0 = valid ? fdt_float_key(pf.x) : 0xffffffffu, k1 = valid ? fdt_float_key(pf.y) : 0xffffffffu;
        unsigned k2 = valid ? fdt_float_key(pf.z) : 0u, k3 = valid ? fdt_float_key(pf.w) : 0u;
        k0 = __reduce_min_sync(0xffffffffu, k0); k1 = __reduce_min_sync(0xffffffffu, k1);
        k2 = __reduce_max_sync(0xffffffffu, k2); k3 = __reduce_max_sync(0xffffffffu, k3);
        if (lane == 0) { atomicMin(&s_bb[0], k0); atomicMin(&s_bb[1], k1); atomicMax(&s_bb[2], k2); atomicMax(&s_bb[3], k3); }
    }
    __syncthreads();
    const float4 bb = make_float4(fdt_key_float(s_bb[0]), fdt_key_float(s_bb[1]), fdt_key_float(s_bb[2]), fdt_key_float(s_bb[3]));
    const bool cull = !(BIP && mb.tile == 0);
    float best = 0.0f;
    int bi = 0;
    for (int t0 = 0; t0 < G; t0 += GT_TILE) {
        const int tn = min(GT_TILE, G - t0);
        // ---- stage + cull + compact (order preserved)
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        bool keep = false;
        if (tid < tn) {
            const float *row = gt + 5 * (g0 + t0 + tid);
            a = make_float4(row[0], row[1], row[2], row[3]);
            const float wbb = fminf(a.z, bb.z) - fmaxf(a.x, bb.x), hbb = fminf(a.w, bb.w) - fmaxf(a.y, bb.y);
            keep = !cull || (t0 + tid == 0) || !(wbb <= 0.0f || hbb <= 0.0f);      // NaN keeps
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        __syncthreads();                               // previous tile fully consumed
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, tc = 0;
#pragma unroll
        for (int w = 0; w < M_WARPS; ++w) { const int c = s_wcnt[w]; if (w < warp) before += c; tc += c; }
        if (keep) {
            const int ps = before + __popc(bal & ((1u << lane) - 1u));
            tile.box[ps] = a; tile.area[ps] = (a.z - a.x) * (a.w - a.y); tile.idx[ps] = t0 + tid;
        }
        __syncthreads();
        for (int g = 0; g < tc; ++g) {
            const float v = iou_match(tile.box[g], tile.area[g], pf, area_b);
            const int gi = tile.idx[g];
            if (gi == 0) { best = v; bi = 0; }
            else if (v > best) { best = v; bi = gi; }                          // first index wins ties (:197)
            if (BIP) {
                unsigned key = valid ? fdt_float_key(v) : 0u;
                unsigned mx = __reduce_max_sync(0xffffffffu, key);
                unsigned cand = (valid && key == mx) ? (unsigned)p : 0xffffffffu;
                unsigned pm = __reduce_min_sync(0xffffffffu, cand);           // first prior wins ties (:136)
                if (lane == 0) s_best[g][warp] = ((unsigned long long)mx << 32) | (0xffffffffu - pm);
            }
        }
        if (BIP) {
            __syncthreads();
            if (tid < tc) {
                unsigned long long m = s_best[tid][0];
#pragma unroll
                for (int w = 1; w < M_WARPS; ++w) m = max(m, s_best[tid][w]);
                atomicMax(&bestprior[g0 + tile.idx[tid]], m);
            }
        }
    }
    if (!valid) return;
    if (BIP) { tmp_idx[t] = bi; tmp_ov[t] = best; }
    else finalize_prior(gt, g0, bi, best, thr, pr, v0, v1, loc_t, conf_t, bti, bto, t, encode_all);
}

// Default (non-bipartite) matcher, the production one (MyTrain_repo.py:113): the same per-warp culling as above but WITHOUT the
// block-level stage -- the GT boxes of a tile are staged once, then every warp lists the boxes that touch the bounding box of its
// 32 consecutive priors (ascending index, so the first-index tie rule holds; GT 0 is always listed) and runs the IoU loop over its
// own list.  Two block barriers per tile instead of six and no compaction bookkeeping: k_match<false> spent more instructions around
// the IoU loop than in it.  Optionally folds in the global maximum of `conf` that log_sum_exp needs (box_utils.py:268): one block
// reduction and one atomicMax per block instead of a separate pass over conf.
__global__ void __launch_bounds__(M_THREADS)
k_match_default(const float4 *__restrict__ priors, const float *__restrict__ gt, const int64_t *__restrict__ gt_off,
                int64_t N, float thr, float v0, float v1,
                float4 *__restrict__ loc_t, int64_t *__restrict__ conf_t, int32_t *__restrict__ bti, float *__restrict__ bto,
                const bool encode_all, const float *__restrict__ conf, int C, unsigned *__restrict__ gmax_key)
{
    fdt_pdl_enter();
    __shared__ float4 s_box[GT_TILE];
    __shared__ float s_area[GT_TILE];
    __shared__ unsigned char s_wl[M_WARPS][GT_TILE];
    __shared__ unsigned s_cmax[M_WARPS];
    const MatchBlock mb = match_block();
    const int b = mb.b, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p = (int64_t)mb.tile * M_THREADS + tid;
    const bool valid = p < N;
    const int64_t g0 = gt_off[b];
    const int G = (int)(gt_off[b + 1] - g0);
    const int64_t t = (int64_t)b * N + p;
    // global max of conf for log_sum_exp (every block, also for images without GT): the row is loaded here and reduced at the END of
    // the kernel, so that nothing waits for it
    float2 cv = make_float2(0.f, 0.f);
    if (conf && valid && C == 2) cv = __ldg(reinterpret_cast<const float2 *>(conf + t * 2));
    auto flush_conf_max = [&]() {
        if (!conf) return;
        unsigned k = 0u;
        if (valid) {
            if (C == 2) k = max(gmax_key_of(cv.x), gmax_key_of(cv.y));
            else for (int c = 0; c < C; ++c) k = max(k, gmax_key_of(__ldg(conf + t * C + c)));
        }
        k = __reduce_max_sync(0xffffffffu, k);
        if (lane == 0) s_cmax[warp] = k;
        __syncthreads();
        if (tid == 0) {
            unsigned m = s_cmax[0];
#pragma unroll
            for (int w = 1; w < M_WARPS; ++w) m = max(m, s_cmax[w]);
            atomicMax(gmax_key, m);
        }
    };
    const float4 pr = priors[valid ? p : 0];
    if (G <= 0) {                                      // reference raises (Q3); defined: all background
        if (valid) {
            conf_t[t] = 0; loc_t[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bti) bti[t] = 0;
            if (bto) bto[t] = 0.0f;
        }
        flush_conf_max();
        return;
    }
    const float hw = pr.z / 2.0f, hh = pr.w / 2.0f;    // point_form, box_utils.py:15-16
    const float4 pf = make_float4(pr.x - hw, pr.y - hh, pr.x + hw, pr.y + hh);
    const float area_b = (pf.z - pf.x) * (pf.w - pf.y);
    float4 wb;                                         // bounding box of the warp's priors
    {
        unsigned k0 = valid ? fdt_float_key(pf.x) : 0xffffffffu, k1 = valid ? fdt_float_key(pf.y) : 0xffffffffu;
        unsigned k2 = valid ? fdt_float_key(pf.z) : 0u, k3 = valid ? fdt_float_key(pf.w) : 0u;
        k0 = __reduce_min_sync(0xffffffffu, k0); k1 = __reduce_min_sync(0xffffffffu, k1);
        k2 = __reduce_max_sync(0xffffffffu, k2); k3 = __reduce_max_sync(0xffffffffu, k3);
        wb = make_float4(fdt_key_float(k0), fdt_key_float(k1), fdt_key_float(k2), fdt_key_float(k3));
    }
    float best = 0.0f;
    int bi = 0;
    for (int t0 = 0; t0 < G; t0 += GT_TILE) {
        const int tn = min(GT_TILE, G - t0);
        __syncthreads();                               // previous tile fully consumed
        if (tid < tn) {
            const float *row = gt + 5 * (g0 + t0 + tid);
            const float4 a = make_float4(row[0], row[1], row[2], row[3]);
            s_box[tid] = a; s_area[tid] = (a.z - a.x) * (a.w - a.y);
        }
        __syncthreads();
        int wn = 0;
        for (int gb = 0; gb < tn; gb += 32) {
            const int g = gb + lane;
            bool k2 = false;
            if (g < tn) {
                const float4 a2 = s_box[g];
                const float wbb = fminf(a2.z, wb.z) - fmaxf(a2.x, wb.x), hbb = fminf(a2.w, wb.w) - fmaxf(a2.y, wb.y);
                k2 = (t0 + g == 0) || !(wbb <= 0.0f || hbb <= 0.0f);           // NaN keeps
            }
            const unsigned bal2 = __ballot_sync(0xffffffffu, k2);
            if (k2) s_wl[warp][wn + __popc(bal2 & ((1u << lane) - 1u))] = (unsigned char)g;
            wn += __popc(bal2);
        }
        __syncwarp();
        for (int q = 0; q < wn; ++q) {
            const int g = s_wl[warp][q];
            const int gi = t0 + g;
            const float v = iou_match_vs_best(s_box[g], s_area[g], pf, area_b, best, gi == 0);
            if (gi == 0) { best = v; bi = 0; }
            else if (v > best) { best = v; bi = gi; }                          // first index wins ties (:197)
        }
        __syncwarp();
    }
    if (valid) finalize_prior(gt, g0, bi, best, thr, pr, v0, v1, loc_t, conf_t, bti, bto, t, encode_all);
    flush_conf_max();
}

// box_utils.py:150-154: best_truth_overlap[best_prior_idx[j]] = 2; best_truth_idx[best_prior_idx[j]] = j (last j wins)
__global__ void __launch_bounds__(M_THREADS)
k_match_bipartite_finalize(const float4 *__restrict__ priors, const float *__restrict__ gt, const int64_t *__restrict__ gt_off,
                           int64_t N, float thr, float v0, float v1,
                           float4 *__restrict__ loc_t, int64_t *__restrict__ conf_t, int32_t *__restrict__ bti, float *__restrict__ bto,
                           const int32_t *__restrict__ tmp_idx, const float *__restrict__ tmp_ov,
                           const unsigned long long *__restrict__ bestprior, const bool encode_all)
{
    fdt_pdl_enter();
    __shared__ unsigned s_bp[GT_TILE];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int64_t p = (int64_t)blockIdx.x * M_THREADS + tid;
    const bool valid = p < N;
    const int64_t g0 = gt_off[b];
    const int G = (int)(gt_off[b + 1] - g0);
    if (G <= 0) return;                                 // k_match already wrote the all-background rows
    const int64_t t = (int64_t)b * N + p;
    int idx = valid ? tmp_idx[t] : 0;
    float ov = valid ? tmp_ov[t] : 0.0f;
    for (int t0 = 0; t0 < G; t0 += GT_TILE) {
        const int tn = min(GT_TILE, G - t0);
        __syncthreads();
        if (tid < tn) s_bp[tid] = 0xffffffffu - (unsigned)(bestprior[g0 + t0 + tid] & 0xffffffffull);
        __syncthreads();
        for (int j = 0; j < tn; ++j)
            if (s_bp[j] == (unsigned)p) { idx = t0 + j; ov = 2.0f; }
    }
    if (valid) finalize_prior(gt, g0, idx, ov, thr, priors[p], v0, v1, loc_t, conf_t, bti, bto, t, encode_all);
}

// ---------------------------------------------------------------------------------------------- loss
__global__ void k_conf_global_max(const float *__restrict__ x, int64_t n, unsigned *__restrict__ gmax_key)
{
    fdt_pdl_enter();
    unsigned k = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        k = max(k, gmax_key_of(x[i]));
    k = __reduce_max_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0) atomicMax(gmax_key, k);
}

// Level-0 mining histogram, chip-wide: neighbouring priors have similar losses, so the 32 lanes of a warp hit a handful of bins --
// one atomic per distinct bin and warp (match_any) instead of one per prior, spread over MINE_REPL copies by warp.
__device__ __forceinline__ void mine_hist_add(int *__restrict__ hist, const int b, const int bin)
{
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    const int lane = threadIdx.x & 31;
    if (bin >= 0 && lane == __ffs(peers) - 1)
        atomicAdd(&hist[((size_t)b * MINE_REPL + ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (MINE_REPL - 1))) * MINE_BINS + bin], __popc(peers));
}

struct LossAcc {            // lives in the workspace, zeroed per call
    double loss_l, loss_c;
    unsigned gmax_key;
    unsigned pad;
};

template <typename T>
__device__ __forceinline__ T block_sum(T v, T *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? s_red[threadIdx.x] : (T)0;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;     // valid in thread 0
}

// multibox_loss.py:90-110
__global__ void __launch_bounds__(M_THREADS)
k_loss_prior(const float4 *__restrict__ loc, const float *__restrict__ conf, const float4 *__restrict__ loc_t,
             const int64_t *__restrict__ conf_t, int64_t N, int C, LossAcc *__restrict__ acc,
             float *__restrict__ loss_c_all, int32_t *__restrict__ num_pos, int *__restrict__ hist)
{
    fdt_pdl_enter();
    __shared__ double s_tot;
    __shared__ int s_cnt;
    const int b = blockIdx.y;
    const int64_t p = (int64_t)blockIdx.x * M_THREADS + threadIdx.x;
    if (threadIdx.x == 0) { s_tot = 0.0; s_cnt = 0; }
    __syncthreads();
    const float xmax = fdt_key_float(acc->gmax_key);
    double sl = 0.0;
    int is_pos = 0, bin = -1;
    if (p < N) {
        const int64_t t = (int64_t)b * N + p;
        const int64_t label = conf_t[t];
        is_pos = label > 0;
        if (is_pos) {                                              // :96-101 smooth L1, beta = 1, sum
            const float4 a = loc[t], g = loc_t[t];
            const float d[4] = {fabsf(a.x - g.x), fabsf(a.y - g.y), fabsf(a.z - g.z), fabsf(a.w - g.w)};
#pragma unroll
            for (int k = 0; k < 4; ++k) sl += (double)(d[k] < 1.0f ? 0.5f * d[k] * d[k] : d[k] - 0.5f);
        }
        const float *row = conf + t * C;
        float s = 0.0f;
        for (int c = 0; c < C; ++c) s += fdt_expf_cr(row[c] - xmax);                // box_utils.py:269
        const float v = (fdt_logf_cr(s) + xmax) - row[label];                      // :106
        const float lc = is_pos ? 0.0f : v;                                        // :110
        loss_c_all[t] = lc;
        bin = (int)(mine_comp(lc, (unsigned)p) >> 52);
    }
    mine_hist_add(hist, b, bin);
    // positives are ~1 % of the priors: only the warps that hold one reduce, into shared memory, and the block flushes once
    const unsigned posm = __ballot_sync(0xffffffffu, is_pos);
    if (posm) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sl += __shfl_xor_sync(0xffffffffu, sl, o);
        if ((threadIdx.x & 31) == 0) { atomicAdd(&s_tot, sl); atomicAdd(&s_cnt, __popc(posm)); }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) {
        if (s_tot != 0.0) atomicAdd(&acc->loss_l, s_tot);
        atomicAdd(&num_pos[b], s_cnt);
    }
}

// standalone mining entry: histogram of the top 12 composite bits + positives per image
__global__ void __launch_bounds__(M_THREADS)
k_mine_hist(const float *__restrict__ loss_c, const uint8_t *__restrict__ pos, int64_t N, int *__restrict__ hist, int32_t *__restrict__ num_pos)
{
    fdt_pdl_enter();
    __shared__ int s_cnt[M_WARPS];
    const int b = blockIdx.y;
    const int64_t p = (int64_t)blockIdx.x * M_THREADS + threadIdx.x;
    int is_pos = 0, bin = -1;
    if (p < N) {
        is_pos = pos[(int64_t)b * N + p] != 0;
        bin = (int)(mine_comp(loss_c[(int64_t)b * N + p], (unsigned)p) >> 52);
    }
    mine_hist_add(hist, b, bin);
    const int cnt = block_sum<int>(is_pos, s_cnt);
    if (threadIdx.x == 0 && cnt) atomicAdd(&num_pos[b], cnt);
}

// One CTA per image: cutoff[b] = the num_neg-th largest composite (selected <=> comp >= cutoff); ~0 = nothing selected.
__global__ void __launch_bounds__(MINE_THREADS, 1)
k_mine_select(const float *__restrict__ loss_c, const int *__restrict__ hist0, const int32_t *__restrict__ num_pos, int64_t N,
              int negpos_ratio, unsigned long long *__restrict__ cutoff)
{
    fdt_pdl_enter();
    __shared__ int s_h[MINE_BINS];
    __shared__ int s_warp[33];
    __shared__ int s_sel[3];
    __shared__ unsigned long long s_cand[MINE_COLLECT];      // composites of the straddling level-0 bin
    __shared__ int s_ncand;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *lrow = loss_c + (int64_t)b * N;
    long long num_neg = (long long)negpos_ratio * num_pos[b];                   // multibox_loss.py:115
    if (num_neg > N - 1) num_neg = N - 1;
    if (num_neg <= 0) { if (tid == 0) cutoff[b] = ~0ull; return; }
    int need = (int)num_neg;
    unsigned long long prefix = 0, pmask = 0;
    bool collected = false;
    for (int level = 0, shift = 52; ; ++level, shift -= 12) {
        const int sh = shift < 0 ? 0 : shift;
        const int nb = shift < 0 ? 16 : MINE_BINS;                            // last level: the 4 lowest bits
        if (level == 0) {
            for (int i = tid; i < MINE_BINS; i += MINE_THREADS) {
                int c = 0;
#pragma unroll
                for (int r = 0; r < MINE_REPL; ++r) c += hist0[((size_t)b * MINE_REPL + r) * MINE_BINS + i];
                s_h[i] = c;
            }
        } else {
            for (int i = tid; i < MINE_BINS; i += MINE_THREADS) s_h[i] = 0;
            __syncthreads();
            if (collected) {
                const int nc = s_ncand;
                for (int e = tid; e < nc; e += MINE_THREADS) {
                    const unsigned long long c = s_cand[e];
                    if ((c & pmask) == prefix) atomicAdd(&s_h[(int)((c >> sh) & (unsigned long long)(nb - 1))], 1);
                }
            } else {
                for (int64_t p = tid; p < N; p += MINE_THREADS) {
                    const unsigned long long c = mine_comp(lrow[p], (unsigned)p);
                    if ((c & pmask) == prefix) atomicAdd(&s_h[(int)((c >> sh) & (unsigned long long)(nb - 1))], 1);
                }
            }
        }
        __syncthreads();
        // descending scan over the bins: find the digit d with (count above d) < need <= (count above d) + h[d]
        int c4[MINE_BINS / MINE_THREADS], sum = 0;
#pragma unroll
        for (int q = 0; q < MINE_BINS / MINE_THREADS; ++q) { c4[q] = s_h[MINE_BINS - 1 - (tid * (MINE_BINS / MINE_THREADS) + q)]; sum += c4[q]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += v; }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        int run = inc - sum + s_warp[warp];
#pragma unroll
        for (int q = 0; q < MINE_BINS / MINE_THREADS; ++q) {
            if (run < need && run + c4[q] >= need) { s_sel[0] = MINE_BINS - 1 - (tid * (MINE_BINS / MINE_THREADS) + q); s_sel[1] = need - run; s_sel[2] = c4[q]; }
            run += c4[q];
        }
        __syncthreads();
        prefix |= (unsigned long long)s_sel[0] << sh;
        pmask |= (unsigned long long)(nb - 1) << sh;
        need = s_sel[1];
        const bool done = (s_sel[2] == need) || shift < 0;                     // whole bin taken, or all 64 bits fixed
        const int bin_count = s_sel[2];
        __syncthreads();
        if (done) break;
        if (level == 0 && bin_count <= MINE_COLLECT) {
            // The cutoff lies inside one level-0 bin of at most MINE_COLLECT composites: ONE scan of the row collects them and
            // the remaining levels histogram that list in shared memory instead of re-reading the row from L2 each time.
            if (tid == 0) s_ncand = 0;
            __syncthreads();
            for (int64_t p = tid; p < N; p += MINE_THREADS) {
                const unsigned long long c = mine_comp(lrow[p], (unsigned)p);
                if ((c & pmask) == prefix) s_cand[atomicAdd(&s_ncand, 1)] = c;
            }
            __syncthreads();
            collected = true;
        }
    }
    if (tid == 0) cutoff[b] = prefix;
}

// MODE 0: neg mask only.  MODE 1: sel = pos | neg and CE over the selection (F.cross_entropy, sum).
// A block covers APPLY_TILES consecutive tiles of one image: the fp64 atomicAdd on the single loss accumulator is the serial
// resource here (one per block), so fewer, longer blocks finish sooner than one block per 256 priors.
constexpr int APPLY_TILES = 8;
template <int MODE>
__global__ void __launch_bounds__(M_THREADS)
k_mine_apply(const float *__restrict__ loss_c, const unsigned long long *__restrict__ cutoff, const int64_t *__restrict__ conf_t,
             const float *__restrict__ conf, int64_t N, int C, uint8_t *__restrict__ out_mask, LossAcc *__restrict__ acc)
{
    fdt_pdl_enter();
    __shared__ double s_red[M_WARPS];
    __shared__ int s_list[APPLY_TILES * M_THREADS];          // selected priors of the block (MODE 1)
    __shared__ int s_n;
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const unsigned long long cut = cutoff[b];
    if (MODE == 1) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
    }
#pragma unroll 2
    for (int u = 0; u < APPLY_TILES; ++u) {
        const int64_t p = ((int64_t)blockIdx.x * APPLY_TILES + u) * M_THREADS + threadIdx.x;
        bool sel = false;
        if (p < N) {
            const int64_t t = (int64_t)b * N + p;
            const bool neg = cut != ~0ull && mine_comp(loss_c[t], (unsigned)p) >= cut;      // multibox_loss.py:116
            const int64_t label = (MODE == 1) ? conf_t[t] : 0;
            sel = neg || (MODE == 1 && label > 0);
            out_mask[t] = (uint8_t)sel;
        }
        if (MODE == 1) {
            // a few percent of the priors are selected: list them and evaluate the fp64 cross entropy over the dense list below,
            // not under a 3-lanes-in-32 branch here
            const unsigned bal = __ballot_sync(0xffffffffu, sel);
            if (bal) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_n, __popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (sel) s_list[base + __popc(bal & ((1u << lane) - 1u))] = (int)p;
            }
        }
    }
    if (MODE == 1) {
        __syncthreads();
        const int n = s_n;
        double ce = 0.0;
        for (int e = threadIdx.x; e < n; e += M_THREADS) {
            const int64_t t = (int64_t)b * N + s_list[e];
            const int64_t label = conf_t[t];
            const float *row = conf + t * C;
            float m = row[0];
            for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
            double s = 0.0;
            for (int c = 0; c < C; ++c) s += exp((double)(row[c] - m));
            ce += (log(s) + (double)m) - (double)row[label];                                 // :128
        }
        if (n) {                                                                             // (uniform: s_n is shared)
            ce = block_sum<double>(ce, s_red);
            if (threadIdx.x == 0 && ce != 0.0) atomicAdd(&acc->loss_c, ce);
        }
    }
}

__global__ void k_loss_final(const LossAcc *__restrict__ acc, const int32_t *__restrict__ num_pos, int B,
                             float *__restrict__ losses, float *__restrict__ norm)
{
    fdt_pdl_enter();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long n = 0;
    for (int b = 0; b < B; ++b) n += num_pos[b];
    double Nn = (double)n;                         // multibox_loss.py:130
    if (n == 0) Nn = (double)B;                    // :132-133
    losses[0] = (float)(acc->loss_l / Nn);
    losses[1] = (float)(acc->loss_c / Nn);
    norm[0] = (float)Nn;
}

// d loss_l / d loc = smooth-L1' at positives / N ; d loss_c / d conf = (softmax - onehot) at pos U neg / N
__global__ void __launch_bounds__(M_THREADS)
k_multibox_backward(const float4 *__restrict__ loc, const float *__restrict__ conf, const float4 *__restrict__ loc_t,
                    const int64_t *__restrict__ conf_t, const uint8_t *__restrict__ sel, const float *__restrict__ norm,
                    float g_l, float g_c, const float *__restrict__ g_l_dev, const float *__restrict__ g_c_dev,
                    int64_t total, int C, float4 *__restrict__ grad_loc, float *__restrict__ grad_conf)
{
    const int64_t t = (int64_t)blockIdx.x * M_THREADS + threadIdx.x;
    if (t >= total) return;
    if (g_l_dev) g_l = __ldg(g_l_dev);              // upstream gradients as device scalars: no host synchronisation
    if (g_c_dev) g_c = __ldg(g_c_dev);
    const float inv = 1.0f / norm[0];
    const int64_t label = conf_t[t];
    float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (label > 0) {
        const float4 a = loc[t], g = loc_t[t];
        const float d[4] = {a.x - g.x, a.y - g.y, a.z - g.z, a.w - g.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = (fabsf(d[k]) < 1.0f ? d[k] : (d[k] > 0.f ? 1.0f : -1.0f)) * g_l * inv;
        gl = make_float4(o[0], o[1], o[2], o[3]);
    }
    grad_loc[t] = gl;
    const float *row = conf + t * C;
    float *go = grad_conf + t * C;
    if (sel[t]) {
        float m = row[0];
        for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(row[c] - m);
        for (int c = 0; c < C; ++c) go[c] = (expf(row[c] - m) / s - (c == label ? 1.0f : 0.0f)) * g_c * inv;
    } else {
        for (int c = 0; c < C; ++c) go[c] = 0.0f;
    }
}

struct MatchWs { unsigned long long *bestprior; int32_t *tmp_idx; float *tmp_ov; size_t bytes; };
MatchWs plan_match_ws(void *ws, int B, int64_t N, int64_t total_gt)
{
    MatchWs m;
    char *p = (char *)ws;
    size_t o = 0;
    m.bestprior = (unsigned long long *)(p + o); o += fdt_align256((size_t)(total_gt > 0 ? total_gt : 1) * 8);
    m.tmp_idx = (int32_t *)(p + o); o += fdt_align256((size_t)B * N * 4);
    m.tmp_ov = (float *)(p + o); o += fdt_align256((size_t)B * N * 4);
    m.bytes = o;
    return m;
}

int launch_match(const float *priors, const float *gt, const int64_t *gt_off, int B, int64_t N, int64_t total_gt,
                 float thr, float v0, float v1, int bipartite, float *loc_t, int64_t *conf_t, int32_t *bti, float *bto,
                 void *ws, cudaStream_t st, bool encode_all = true, const float *conf = nullptr, int C = 0, unsigned *gmax_key = nullptr)
{
    dim3 grid((unsigned)((N + M_THREADS - 1) / M_THREADS), (unsigned)B);
    MatchWs m = plan_match_ws(ws, B, N, total_gt);
    if (!bipartite) {
        FDT_CUDA(launch_pdl(k_match_default, grid, dim3(M_THREADS), st, (const float4 *)priors, gt, gt_off, N, thr, v0, v1, (float4 *)loc_t, conf_t,
                                                    bti, bto, encode_all, conf, C, gmax_key));
        FDT_LAUNCH_CHECK();
    } else {
        FDT_CUDA(cudaMemsetAsync(m.bestprior, 0, (size_t)(total_gt > 0 ? total_gt : 1) * 8, st));
        FDT_CUDA(launch_pdl(k_match<true>, grid, dim3(M_THREADS), st, (const float4 *)priors, gt, gt_off, N, thr, v0, v1, (float4 *)loc_t, conf_t,
                                                  bti, bto, m.tmp_idx, m.tmp_ov, m.bestprior, encode_all));
        FDT_LAUNCH_CHECK();
        FDT_CUDA(launch_pdl(k_match_bipartite_finalize, grid, dim3(M_THREADS), st, (const float4 *)priors, gt, gt_off, N, thr, v0, v1, (float4 *)loc_t,
                                                               conf_t, bti, bto, (const int32_t *)m.tmp_idx, (const float *)m.tmp_ov, (const unsigned long long *)m.bestprior, encode_all));
        FDT_LAUNCH_CHECK();
    }
    return FDT_OK;
}

struct MineWs { int *hist; unsigned long long *cutoff; int32_t *num_pos; size_t bytes; };
MineWs plan_mine_ws(void *ws, int B)
{
    MineWs m;
    char *p = (char *)ws;
    size_t o = 0;
    m.hist = (int *)(p + o); o += fdt_align256((size_t)B * MINE_REPL * MINE_BINS * 4);
    m.cutoff = (unsigned long long *)(p + o); o += fdt_align256((size_t)B * 8);
    m.num_pos = (int32_t *)(p + o); o += fdt_align256((size_t)B * 4);
    m.bytes = o;
    return m;
}

template <int MODE>
int launch_mine_tail(const float *loss_c, const MineWs &m, const int64_t *conf_t, const float *conf, int B, int64_t N, int C,
                     int ratio, uint8_t *mask, LossAcc *acc, cudaStream_t st)
{
    FDT_CUDA(launch_pdl(k_mine_select, dim3(B), dim3(MINE_THREADS), st, loss_c, (const int *)m.hist, (const int32_t *)m.num_pos, N, ratio, m.cutoff));
    FDT_LAUNCH_CHECK();
    dim3 grid((unsigned)((N + M_THREADS * APPLY_TILES - 1) / (M_THREADS * APPLY_TILES)), (unsigned)B);
    FDT_CUDA(launch_pdl(k_mine_apply<MODE>, grid, dim3(M_THREADS), st, loss_c, (const unsigned long long *)m.cutoff, conf_t, conf, N, C, mask, acc));
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}

}  // namespace

// ================================================================================================ C ABI
FDT_API size_t fdt_match_workspace_bytes(int B, int64_t N, int64_t total_gt)
{
    if (B <= 0 || N <= 0) return 256;
    return plan_match_ws(nullptr, B, N, total_gt).bytes;
}

static int match_args_ok(const char *who, const float *priors, const float *gt, const int64_t *gt_off, int B, int64_t N,
                         const float *loc_t, const int64_t *conf_t, const void *ws)
{
    FDT_REQUIRE(B >= 0 && N >= 0 && N < (1ll << 31), FDT_E_INVALID, "%s: bad sizes B=%d N=%lld", who, B, (long long)N);
    if (B == 0 || N == 0) return FDT_OK;
    FDT_REQUIRE(priors && gt && gt_off && loc_t && conf_t && ws, FDT_E_INVALID, "%s: null pointer argument", who);
    FDT_REQUIRE(fdt_aligned(priors, 16) && fdt_aligned(loc_t, 16) && fdt_aligned(ws, 256), FDT_E_INVALID,
                "%s: priors/loc_t need 16-byte, workspace 256-byte alignment", who);
    return FDT_OK;
}

FDT_API int fdt_match_encode(const float *priors, const float *gt, const int64_t *gt_off, int64_t total_gt, int B, int64_t N,
                             float threshold, float var0, float var1, int bipartite,
                             float *loc_t, int64_t *conf_t, int32_t *best_truth_idx, float *best_truth_overlap,
                             void *ws, size_t ws_bytes, fdt_stream_t stream)
{
    int rc = match_args_ok("fdt_match_encode", priors, gt, gt_off, B, N, loc_t, conf_t, ws);
    if (rc != FDT_OK || B == 0 || N == 0) return rc;
    FDT_REQUIRE(total_gt >= 0, FDT_E_INVALID, "fdt_match_encode: negative total_gt");
    FDT_REQUIRE(ws_bytes >= fdt_match_workspace_bytes(B, N, total_gt), FDT_E_WORKSPACE,
                "fdt_match_encode: workspace %zu < %zu bytes", ws_bytes, fdt_match_workspace_bytes(B, N, total_gt));
    return launch_match(priors, gt, gt_off, B, N, total_gt, threshold, var0, var1, bipartite, loc_t, conf_t,
                        best_truth_idx, best_truth_overlap, ws, (cudaStream_t)stream);
}

FDT_API size_t fdt_mine_workspace_bytes(int B, int64_t) { return plan_mine_ws(nullptr, B > 0 ? B : 1).bytes; }

FDT_API int fdt_hard_negative_mine(const float *loss_c, const uint8_t *pos, int B, int64_t N, int negpos_ratio,
                                   uint8_t *neg, void *ws, size_t ws_bytes, fdt_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    FDT_REQUIRE(B >= 0 && N >= 0 && negpos_ratio >= 0, FDT_E_INVALID, "fdt_hard_negative_mine: bad sizes");
    if (B == 0 || N == 0) return FDT_OK;
    FDT_REQUIRE(loss_c && pos && neg && ws && fdt_aligned(ws, 256), FDT_E_INVALID, "fdt_hard_negative_mine: null / misaligned pointer");
    MineWs m = plan_mine_ws(ws, B);
    FDT_REQUIRE(ws_bytes >= m.bytes, FDT_E_WORKSPACE, "fdt_hard_negative_mine: workspace %zu < %zu bytes", ws_bytes, m.bytes);
    FDT_CUDA(cudaMemsetAsync(ws, 0, m.bytes, st));
    dim3 grid((unsigned)((N + M_THREADS - 1) / M_THREADS), (unsigned)B);
    FDT_CUDA(launch_pdl(k_mine_hist, grid, dim3(M_THREADS), st, loss_c, pos, N, m.hist, m.num_pos));
    FDT_LAUNCH_CHECK();
    return launch_mine_tail<0>(loss_c, m, nullptr, nullptr, B, N, 2, negpos_ratio, neg, nullptr, st);
}

struct LossWs { LossAcc *acc; MineWs mine; float *loss_c_all; void *match; size_t bytes; };
static LossWs plan_loss_ws(void *ws, int B, int64_t N, int64_t total_gt)
{
    LossWs w;
    char *p = (char *)ws;
    size_t o = 0;
    w.acc = (LossAcc *)(p + o); o += 256;
    w.mine = plan_mine_ws(p + o, B); o += w.mine.bytes;
    w.loss_c_all = (float *)(p + o); o += fdt_align256((size_t)B * N * 4);
    w.match = (void *)(p + o); o += plan_match_ws(nullptr, B, N, total_gt).bytes;
    w.bytes = o;
    return w;
}

FDT_API size_t fdt_multibox_workspace_bytes(int B, int64_t N, int, int64_t total_gt)
{
    if (B <= 0 || N <= 0) return 256;
    return plan_loss_ws(nullptr, B, N, total_gt).bytes;
}

FDT_API int fdt_multibox_loss_forward(const float *loc, const float *conf, const float *priors,
                                      const float *gt, const int64_t *gt_off, int64_t total_gt, int B, int64_t N, int C,
                                      float threshold, int negpos_ratio, int bipartite, float var0, float var1,
                                      float *losses, float *norm, float *loc_t, int64_t *conf_t, uint8_t *sel,
                                      float *loss_c_all, void *ws, size_t ws_bytes, fdt_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int rc = match_args_ok("fdt_multibox_loss_forward", priors, gt, gt_off, B, N, loc_t, conf_t, ws);
    if (rc != FDT_OK) return rc;
    FDT_REQUIRE(B > 0 && N > 0 && C >= 2, FDT_E_INVALID, "fdt_multibox_loss_forward: needs B > 0, N > 0, C >= 2");
    FDT_REQUIRE(loc && conf && losses && norm && sel, FDT_E_INVALID, "fdt_multibox_loss_forward: null pointer argument");
    FDT_REQUIRE(fdt_aligned(loc, 16), FDT_E_INVALID, "fdt_multibox_loss_forward: loc needs 16-byte alignment");
    FDT_REQUIRE(total_gt >= 0, FDT_E_INVALID, "fdt_multibox_loss_forward: negative total_gt");
    LossWs w = plan_loss_ws(ws, B, N, total_gt);
    FDT_REQUIRE(ws_bytes >= w.bytes, FDT_E_WORKSPACE, "fdt_multibox_loss_forward: workspace %zu < %zu bytes", ws_bytes, w.bytes);
    float *lca = loss_c_all ? loss_c_all : w.loss_c_all;

    FDT_CUDA(cudaMemsetAsync(w.acc, 0, 256 + w.mine.bytes, st));
    if (bipartite) {               // the default matcher folds the global max of conf (box_utils.py:268) into its own pass
        const int64_t n_conf = (int64_t)B * N * C;
        unsigned blocks = (unsigned)((n_conf + 256 * 8 - 1) / (256 * 8));
        if (blocks > FDT_NUM_SMS * 8) blocks = FDT_NUM_SMS * 8;
        k_conf_global_max<<<blocks, 256, 0, st>>>(conf, n_conf, &w.acc->gmax_key);
        FDT_LAUNCH_CHECK();
    }
    rc = launch_match(priors, gt, gt_off, B, N, total_gt, threshold, var0, var1, bipartite, loc_t, conf_t, nullptr, nullptr, w.match, st, false,
                      bipartite ? nullptr : conf, C, &w.acc->gmax_key);
    if (rc != FDT_OK) return rc;
    dim3 grid((unsigned)((N + M_THREADS - 1) / M_THREADS), (unsigned)B);
    FDT_CUDA(launch_pdl(k_loss_prior, grid, dim3(M_THREADS), st, (const float4 *)loc, conf, (const float4 *)loc_t, (const int64_t *)conf_t, N, C, w.acc, lca, w.mine.num_pos, w.mine.hist));
    FDT_LAUNCH_CHECK();
    rc = launch_mine_tail<1>(lca, w.mine, conf_t, conf, B, N, C, negpos_ratio, sel, w.acc, st);
    if (rc != FDT_OK) return rc;
    FDT_CUDA(launch_pdl(k_loss_final, dim3(1), dim3(32), st, (const LossAcc *)w.acc, (const int32_t *)w.mine.num_pos, B, losses, norm));
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}

FDT_API int fdt_multibox_loss_backward(const float *loc, const float *conf, const float *loc_t, const int64_t *conf_t,
                                       const uint8_t *sel, const float *norm, float g_l, float g_c,
                                       int B, int64_t N, int C, float *grad_loc, float *grad_conf, fdt_stream_t stream)
{
    FDT_REQUIRE(B >= 0 && N >= 0 && C >= 2, FDT_E_INVALID, "fdt_multibox_loss_backward: bad sizes");
    if (B == 0 || N == 0) return FDT_OK;
    FDT_REQUIRE(loc && conf && loc_t && conf_t && sel && norm && grad_loc && grad_conf, FDT_E_INVALID,
                "fdt_multibox_loss_backward: null pointer argument");
    FDT_REQUIRE(fdt_aligned(loc, 16) && fdt_aligned(loc_t, 16) && fdt_aligned(grad_loc, 16), FDT_E_INVALID,
                "fdt_multibox_loss_backward: loc/loc_t/grad_loc need 16-byte alignment");
    const int64_t total = (int64_t)B * N;
    k_multibox_backward<<<(unsigned)((total + M_THREADS - 1) / M_THREADS), M_THREADS, 0, (cudaStream_t)stream>>>(
        (const float4 *)loc, conf, (const float4 *)loc_t, conf_t, sel, norm, g_l, g_c, nullptr, nullptr, total, C, (float4 *)grad_loc, grad_conf);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}

FDT_API int fdt_multibox_loss_backward_dev(const float *loc, const float *conf, const float *loc_t, const int64_t *conf_t,
                                           const uint8_t *sel, const float *norm, const float *g_l_dev, const float *g_c_dev,
                                           int B, int64_t N, int C, float *grad_loc, float *grad_conf, fdt_stream_t stream)
{
    FDT_REQUIRE(B >= 0 && N >= 0 && C >= 2, FDT_E_INVALID, "fdt_multibox_loss_backward_dev: bad sizes");
    if (B == 0 || N == 0) return FDT_OK;
    FDT_REQUIRE(loc && conf && loc_t && conf_t && sel && norm && grad_loc && grad_conf && g_l_dev && g_c_dev, FDT_E_INVALID,
                "fdt_multibox_loss_backward_dev: null pointer argument");
    FDT_REQUIRE(fdt_aligned(loc, 16) && fdt_aligned(loc_t, 16) && fdt_aligned(grad_loc, 16), FDT_E_INVALID,
                "fdt_multibox_loss_backward_dev: loc/loc_t/grad_loc need 16-byte alignment");
    const int64_t total = (int64_t)B * N;
    k_multibox_backward<<<(unsigned)((total + M_THREADS - 1) / M_THREADS), M_THREADS, 0, (cudaStream_t)stream>>>(
        (const float4 *)loc, conf, (const float4 *)loc_t, conf_t, sel, norm, 0.0f, 0.0f, g_l_dev, g_c_dev, total, C, (float4 *)grad_loc, grad_conf);
    FDT_LAUNCH_CHECK();
    return FDT_OK;
}
