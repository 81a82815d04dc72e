// MultiBoxLoss (layers/modules/multibox_loss.py:48-136): fused jaccard/argmax match + encode
// (layers/box_utils.py:103-210), per-prior confidence loss, hard-negative mining, loss reduction and
// the backward pass, for sm_100a.
//
//   forward, default matcher (three kernels chained by programmatic dependent launch):
//     k_mbl_prepare   zeroes the per-call state and reduces the global maximum of conf (box_utils.py:268) to per-block partials
//     k_match_loss    one thread per (image, prior): per-warp culled IoU arg max (first-index ties, torch.max(0)), label +
//                     encode of the positives, then smooth L1 (:96-101), the mining input loss_c (:104-110) and the image's
//                     mining histogram -- the [G,N] overlap matrix is never materialised, conf_t / loc_t are not read back
//     k_mine_apply2   neg = rank < num_neg of the descending sort (:112-116) == the num_neg largest 64-bit composites
//                     (loss key << 32 | ~prior): unique, so ties resolve to the lower prior index.  Cutoff bin from the
//                     histogram, mask + cross entropy over pos U neg (:119-128), the candidates of the cutoff bin settled by the
//                     image's last block, the division by N (:130-135) by the grid's last block
//   forward, bipartite matcher: k_mbl_prepare (+ the boxes of the 32-prior tiles) -> k_best_prior (one warp per GT box: its best
//                     prior, box_utils.py:136) -> k_match_loss<BIP> (forces those priors, :150-154) -> k_mine_apply2
//   standalone entries: k_match_default / k_match<true> + k_match_bipartite_finalize (fdt_match_encode: every row encoded, arg max
//                     and overlap of every prior returned), k_mine_hist / k_mine_select / k_mine_apply (fdt_hard_negative_mine)
//   k_multibox_backward.
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

// Dispatch order of the matchers: blocks are issued x-fastest, and the blocks of the coarse pyramid levels (the END of the prior
// array: a 512-pixel prior overlaps most GT boxes) run several times longer than the rest -- in image-major order the last image's
// coarse tiles start last and the kernel ends with a dozen SMs finishing them alone (ncu: SMs busy 61 % of the kernel).  The grid
// is (images, tiles): x = image, y counts the tiles from the last one down, i.e. every image's last tile first (and no division).
struct MatchBlock { int b; int tile; };
__device__ __forceinline__ MatchBlock match_block()
{
    MatchBlock m;
    m.b = (int)blockIdx.x;
    m.tile = (int)gridDim.y - 1 - (int)blockIdx.y;
    return m;
}
__host__ inline dim3 match_grid(int B, int64_t N) { return dim3((unsigned)B, (unsigned)((N + M_THREADS - 1) / M_THREADS)); }

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

// Default (non-bipartite) matcher, the production one (MyTrain_repo.py:113): the same culling as above but per WARP and without the
// block-level stage -- the GT boxes of a tile are staged once, then every warp lists the boxes that touch the bounding box of its
// 32 consecutive priors (ascending index, so the first-index tie rule holds; GT 0 is always listed) and runs the IoU loop over its
// own list.  Two block barriers per tile and no compaction bookkeeping.  The IoU loop is the bulk of the forward's instructions:
// one 32-byte shared-memory entry per GT box addressed through 32-bit shared addresses (the generic-address form re-derived the
// shared window base inside the loop), GT 0 peeled (it seeds the arg max unconditionally, box_utils.py:197).
struct __align__(32) GtEntry { float4 box; float area; float label; float pad[2]; };

//
// AREA_CULL (the fused forward only, which outputs nothing but the labels and the encoded POSITIVES): IoU <= min(area) / max(area),
// so a GT box whose area is below thr * (smallest prior area of the warp) or above (largest prior area) / thr cannot reach the
// positive threshold with any prior of the warp and is not listed.  Every box that does reach thr with some prior is listed, so a
// positive's best overlap and arg max (first index among ties) are unchanged, and a prior whose true best is below thr still ends
// below it: conf_t and the encoded rows are bit-identical; only the arg max of BACKGROUND priors (not an output here) may differ.
// The 0.999 margin covers the few fp32 roundings between the bound and the computed IoU; boxes with a non-positive or NaN area
// and GT 0 are always listed.  At the production threshold (0.35) the 16-pixel level drops ~85 % of its candidates and the 256- and
// 512-pixel levels, whose warps used to walk every GT box of the image, usually list none.
template <bool AREA_CULL>
__device__ __forceinline__ void match_default_core(const float *__restrict__ gt, const int64_t g0, const int G, const float4 pf,
                                                   const float area_b, const bool valid, const float thr, GtEntry *s_gt,
                                                   unsigned char *s_wl_warp, float &best, int &bi)
{
    const int tid = threadIdx.x, lane = tid & 31;
    float4 wb;                                         // bounding box of the warp's priors
    float ap_lo = 0.0f, ap_hi = 0.0f;                  // thr' * smallest prior area, largest prior area / thr' (AREA_CULL)
    {
        unsigned k0 = valid ? fdt_float_key(pf.x) : 0xffffffffu, k1 = valid ? fdt_float_key(pf.y) : 0xffffffffu;
        unsigned k2 = valid ? fdt_float_key(pf.z) : 0u, k3 = valid ? fdt_float_key(pf.w) : 0u;
        k0 = __reduce_min_sync(0xffffffffu, k0); k1 = __reduce_min_sync(0xffffffffu, k1);
        k2 = __reduce_max_sync(0xffffffffu, k2); k3 = __reduce_max_sync(0xffffffffu, k3);
        wb = make_float4(fdt_key_float(k0), fdt_key_float(k1), fdt_key_float(k2), fdt_key_float(k3));
        if (AREA_CULL) {
            // any invalid lane, non-positive or NaN prior area switches the area cull off for the warp (amin <= 0)
            const bool okp = valid && area_b > 0.0f;
            const unsigned all_ok = __all_sync(0xffffffffu, okp || !valid) && __any_sync(0xffffffffu, okp);
            const unsigned amin = __reduce_min_sync(0xffffffffu, okp ? __float_as_uint(area_b) : 0xffffffffu);
            const unsigned amax = __reduce_max_sync(0xffffffffu, okp ? __float_as_uint(area_b) : 0u);
            const float kk = 0.999f * thr;
            if (all_ok && kk > 0.0f) { ap_lo = kk * __uint_as_float(amin); ap_hi = __uint_as_float(amax) / kk; }
        }
    }
    const bool area_cull = AREA_CULL && ap_lo > 0.0f;
    const unsigned a_gt = (unsigned)__cvta_generic_to_shared(s_gt);
    const unsigned a_wl = (unsigned)__cvta_generic_to_shared(s_wl_warp);
    best = 0.0f;
    bi = 0;
    for (int t0 = 0; t0 < G; t0 += GT_TILE) {
        const int tn = min(GT_TILE, G - t0);
        __syncthreads();                               // previous tile fully consumed
        if (tid < tn) {
            const float *row = gt + 5 * (g0 + t0 + tid);
            const float4 a = make_float4(row[0], row[1], row[2], row[3]);
            s_gt[tid].box = a; s_gt[tid].area = (a.z - a.x) * (a.w - a.y); s_gt[tid].label = row[4];
        }
        __syncthreads();
        int wn = 0;
        for (int gb = 0; gb < tn; gb += 32) {
            const int g = gb + lane;
            bool k2 = false;
            if (g < tn) {
                const float4 a2 = s_gt[g].box;
                const float wbb = fminf(a2.z, wb.z) - fmaxf(a2.x, wb.x), hbb = fminf(a2.w, wb.w) - fmaxf(a2.y, wb.y);
                k2 = !(wbb <= 0.0f || hbb <= 0.0f);                            // NaN keeps
                if (area_cull) {
                    const float ag = s_gt[g].area;
                    if (ag > 0.0f && (ag < ap_lo || ag > ap_hi)) k2 = false;
                }
                k2 = k2 || (t0 + g == 0);
            }
            const unsigned bal2 = __ballot_sync(0xffffffffu, k2);
            if (k2) s_wl_warp[wn + __popc(bal2 & ((1u << lane) - 1u))] = (unsigned char)g;
            wn += __popc(bal2);
        }
        __syncwarp();
        int q = 0;
        if (t0 == 0) {                                 // GT 0 heads the first list: best = its IoU whatever it is
            best = iou_match(s_gt[0].box, s_gt[0].area, pf, area_b);
            q = 1;
        }
#pragma unroll 2
        for (; q < wn; ++q) {
            unsigned g;
            float4 a;
            float ar;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(g) : "r"(a_wl + q));
            const unsigned ad = a_gt + g * 32u;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(ad));
            asm volatile("ld.shared.f32 %0, [%1+16];" : "=f"(ar) : "r"(ad));
            const float v = iou_match_vs_best(a, ar, pf, area_b, best, false);
            if (v > best) { best = v; bi = t0 + (int)g; }                      // first index wins ties (:197)
        }
        __syncwarp();
    }
}

// standalone matcher (fdt_match_encode, box_utils.match): labels + encode of every prior
__global__ void __launch_bounds__(M_THREADS)
k_match_default(const float4 *__restrict__ priors, const float *__restrict__ gt, const int64_t *__restrict__ gt_off,
                int64_t N, float thr, float v0, float v1,
                float4 *__restrict__ loc_t, int64_t *__restrict__ conf_t, int32_t *__restrict__ bti, float *__restrict__ bto,
                const bool encode_all)
{
    fdt_pdl_enter();
    __shared__ GtEntry s_gt[GT_TILE];
    __shared__ unsigned char s_wl[M_WARPS][GT_TILE];
    const MatchBlock mb = match_block();
    const int b = mb.b, tid = threadIdx.x, warp = tid >> 5;
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
    float best;
    int bi;
    match_default_core<false>(gt, g0, G, pf, area_b, valid, thr, s_gt, s_wl[warp], best, bi);
    if (valid) finalize_prior(gt, g0, bi, best, thr, pr, v0, v1, loc_t, conf_t, bti, bto, t, encode_all);
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
// Level-0 mining histogram of the standalone entry (fdt_hard_negative_mine), chip-wide: neighbouring priors have similar losses, so
// the 32 lanes of a warp hit a handful of bins -- one atomic per distinct bin and warp (match_any), spread over MINE_REPL copies.
__device__ __forceinline__ void mine_hist_add(int *__restrict__ hist, const int b, const int bin)
{
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    const int lane = threadIdx.x & 31;
    if (bin >= 0 && lane == __ffs(peers) - 1)
        atomicAdd(&hist[((size_t)b * MINE_REPL + ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (MINE_REPL - 1))) * MINE_BINS + bin], __popc(peers));
}

// Mining bins of the fused forward: 4,096 bins, monotone in the composite's loss key -- 256 per octave over [2^-8, 2^8), everything
// below (the zeros of the positives, tiny and negative losses) in bin 0, everything above (and NaN) in bin 4095.  The cutoff of an
// image falls into ONE bin; with 8 mantissa bits per octave that bin holds a few dozen of its 34 k priors, so the exact order only
// has to be settled among those (k_mine_apply2).
__device__ __forceinline__ int mine_bin2(const float lc)
{
    const unsigned key = lc != lc ? 0xffffffffu : fdt_float_key(lc);
    const unsigned lo = 0x80000000u | 0x3b800000u;                 // key of 2^-8
    if (key < lo) return 0;
    const unsigned v = (key - lo) >> 15;
    return v > (unsigned)(MINE_BINS - 1) ? MINE_BINS - 1 : (int)v;
}

struct LossAcc {            // lives in the workspace, zeroed per call (k_mbl_prepare)
    double loss_l, loss_c;
    unsigned done;          // blocks of k_mine_apply2 that have finished
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

// First kernel of the forward: zeroes the per-call state (accumulators, per-image counters, the mining histogram) and reduces the
// GLOBAL maximum of conf that log_sum_exp subtracts (box_utils.py:268) to one key per block -- the consumers take the maximum of
// the PREP_BLOCKS partials, so nothing has to be zeroed before this kernel and no memset node precedes it.  It releases its
// successor only after its own grid dependency has resolved: when the matcher starts, the previous call on this workspace is over.
constexpr int PREP_BLOCKS = FDT_NUM_SMS, PREP_THREADS = 1024;
__global__ void __launch_bounds__(PREP_THREADS)
k_mbl_prepare(const float *__restrict__ conf, const int64_t n_conf, unsigned *__restrict__ gmax_part, int4 *__restrict__ zero_base,
              const int64_t zero_int4s, const float4 *__restrict__ priors, const int64_t N, float4 *__restrict__ tile_bb, float4 *__restrict__ tile_dim,
              float4 *__restrict__ super_bb, float4 *__restrict__ super_dim)
{
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    __shared__ unsigned s_k[PREP_THREADS / 32];
    const int64_t gtid = (int64_t)blockIdx.x * PREP_THREADS + threadIdx.x, gsz = (int64_t)PREP_BLOCKS * PREP_THREADS;
    for (int64_t i = gtid; i < zero_int4s; i += gsz) zero_base[i] = make_int4(0, 0, 0, 0);
    if (tile_bb) {
        // bipartite matcher: the bounding box (point form) and the ranges (min, max) of the widths and heights of every 32
        // consecutive priors (a tile) and of every 1,024 (a super-tile = one iteration of a block), for k_best_prior's two-level
        // culling.  Ranges of -1 mark a tile with a prior whose width or height is not positive (NaN included): never pruned by size.
        __shared__ unsigned s_t[8][PREP_THREADS / 32];
        __shared__ int s_ok[PREP_THREADS / 32];
        const int64_t n1024 = (N + 1023) & ~(int64_t)1023;
        for (int64_t q = gtid; q < n1024; q += gsz) {                           // (uniform trip count within a block)
            const bool valid = q < N;
            const float4 pr = priors[valid ? q : 0];
            const float hw = pr.z / 2.0f, hh = pr.w / 2.0f;
            const float4 pf = make_float4(pr.x - hw, pr.y - hh, pr.x + hw, pr.y + hh);
            const float pw = pf.z - pf.x, ph = pf.w - pf.y;
            const bool okp = !valid || (pw > 0.0f && ph > 0.0f);
            unsigned k[8];
            k[0] = valid ? fdt_float_key(pf.x) : 0xffffffffu; k[1] = valid ? fdt_float_key(pf.y) : 0xffffffffu;
            k[2] = valid ? fdt_float_key(pf.z) : 0u; k[3] = valid ? fdt_float_key(pf.w) : 0u;
            k[4] = valid && okp ? __float_as_uint(pw) : 0xffffffffu; k[5] = valid && okp ? __float_as_uint(pw) : 0u;      // (positive floats order as integers)
            k[6] = valid && okp ? __float_as_uint(ph) : 0xffffffffu; k[7] = valid && okp ? __float_as_uint(ph) : 0u;
#pragma unroll
            for (int c = 0; c < 8; ++c) k[c] = (c < 2 || c == 4 || c == 6) ? __reduce_min_sync(0xffffffffu, k[c]) : __reduce_max_sync(0xffffffffu, k[c]);
            const bool all_ok = __all_sync(0xffffffffu, okp);
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const float4 none = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
            if (lane == 0) {
                if ((q >> 5) < ((N + 31) >> 5)) {
                    tile_bb[q >> 5] = make_float4(fdt_key_float(k[0]), fdt_key_float(k[1]), fdt_key_float(k[2]), fdt_key_float(k[3]));
                    tile_dim[q >> 5] = all_ok ? make_float4(__uint_as_float(k[4]), __uint_as_float(k[5]), __uint_as_float(k[6]), __uint_as_float(k[7])) : none;
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) s_t[c][warp] = k[c];
                s_ok[warp] = all_ok;
            }
            __syncthreads();
            if (warp == 0) {
                unsigned r[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) r[c] = (c < 2 || c == 4 || c == 6) ? __reduce_min_sync(0xffffffffu, s_t[c][lane]) : __reduce_max_sync(0xffffffffu, s_t[c][lane]);
                const bool ok = __all_sync(0xffffffffu, s_ok[lane] != 0);
                if (lane == 0) {
                    super_bb[q >> 10] = make_float4(fdt_key_float(r[0]), fdt_key_float(r[1]), fdt_key_float(r[2]), fdt_key_float(r[3]));
                    super_dim[q >> 10] = ok ? make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7])) : none;
                }
            }
            __syncthreads();
        }
    }
    unsigned k = 0u;
    if (((uintptr_t)conf & 15) == 0) {
        const float4 *c4 = reinterpret_cast<const float4 *>(conf);
        const int64_t n4 = n_conf >> 2;
        for (int64_t i = gtid; i < n4; i += gsz) {
            const float4 v = __ldg(c4 + i);
            k = max(max(k, gmax_key_of(v.x)), max(gmax_key_of(v.y), max(gmax_key_of(v.z), gmax_key_of(v.w))));
        }
        for (int64_t i = (n4 << 2) + gtid; i < n_conf; i += gsz) k = max(k, gmax_key_of(__ldg(conf + i)));
    } else {
        for (int64_t i = gtid; i < n_conf; i += gsz) k = max(k, gmax_key_of(__ldg(conf + i)));
    }
    k = __reduce_max_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0) s_k[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x < 32) {
        k = __reduce_max_sync(0xffffffffu, s_k[threadIdx.x]);
        if (threadIdx.x == 0) gmax_part[blockIdx.x] = k;
    }
}
// the global maximum from the partials of k_mbl_prepare (one warp-wide reduction; every warp of a consumer does its own)
__device__ __forceinline__ float gmax_from_partials(const unsigned *__restrict__ gmax_part)
{
    unsigned k = 0u;
    for (int i = threadIdx.x & 31; i < PREP_BLOCKS; i += 32) k = max(k, __ldcg(gmax_part + i));
    return fdt_key_float(__reduce_max_sync(0xffffffffu, k));
}

// per-prior loss terms of the fused k_match_loss (multibox_loss.py:96-110): smooth L1 of a positive
// (beta = 1, sum) in `sl`, the mining input loss_c (0 for positives) returned; cv = the row when C == 2 (loaded by the caller)
__device__ __forceinline__ float prior_loss_terms(const float4 a, const float4 g, const bool is_pos, const float2 cv,
                                                  const float *__restrict__ row, const int C, const int64_t label, const float xmax,
                                                  double &sl)
{
    if (is_pos) {
        const float d[4] = {fabsf(a.x - g.x), fabsf(a.y - g.y), fabsf(a.z - g.z), fabsf(a.w - g.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) sl += (double)(d[k] < 1.0f ? 0.5f * d[k] * d[k] : d[k] - 0.5f);
    }
    float s, xl;
    if (C == 2) {
        s = fdt_expf_cr(cv.x - xmax) + fdt_expf_cr(cv.y - xmax);                   // box_utils.py:269
        xl = label == 0 ? cv.x : (label == 1 ? cv.y : row[label]);
    } else {
        s = 0.0f;
        for (int c = 0; c < C; ++c) s += fdt_expf_cr(row[c] - xmax);
        xl = row[label];
    }
    const float v = (fdt_logf_cr(s) + xmax) - xl;                                  // :106
    return is_pos ? 0.0f : v;                                                      // :110
}
// positives are ~1 % of the priors: only the warps that hold one reduce, into shared memory, and the block flushes once
__device__ __forceinline__ void flush_positives(double sl, const int is_pos, double *s_tot, int *s_cnt, LossAcc *__restrict__ acc,
                                                int32_t *__restrict__ num_pos_b)
{
    const unsigned posm = __ballot_sync(0xffffffffu, is_pos);
    if (posm) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sl += __shfl_xor_sync(0xffffffffu, sl, o);
        if ((threadIdx.x & 31) == 0) { atomicAdd(s_tot, sl); atomicAdd(s_cnt, __popc(posm)); }
    }
    __syncthreads();
    if (threadIdx.x == 0 && *s_cnt) {
        if (*s_tot != 0.0) atomicAdd(&acc->loss_l, *s_tot);
        atomicAdd(num_pos_b, *s_cnt);
    }
}

// Bipartite matcher, per-GT half (box_utils.py:136, overlaps.max(1)): the prior with the largest IoU for every GT box, first prior
// among ties.  One block per GT box of the whole batch (the priors are the same for every image).  Two-level culling over the
// boxes k_mbl_prepare made of every 1,024 and every 32 consecutive priors: a tile whose box does not touch the GT box holds exact
// zeros only, which cannot beat an evaluated pair (tile 0 is always evaluated, so an all-zero row ends at prior 0 as torch.max
// does).  A best-first order keeps the evaluated tiles to a handful: with the width / height ranges of a tile's priors, IoU <= I /
// (smallest prior area + GT area - I), I = min(w) * min(h); the tiles with the largest bounds (the pyramid level that fits the box)
// are evaluated first and the others only while their bound still reaches the best overlap found -- normally none.  A skipped prior is strictly below the maximum, so the arg max and
// its first-index tie rule are exact.
__global__ void __launch_bounds__(M_THREADS)
k_best_prior(const float4 *__restrict__ priors, const int64_t N, const float *__restrict__ gt, const int64_t total_gt,
             const float4 *__restrict__ tile_bb, const float4 *__restrict__ tile_dim, const float4 *__restrict__ super_bb,
             const float4 *__restrict__ super_dim, unsigned *__restrict__ bestprior)
{
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();                   // the tile boxes
    constexpr int TL_CAP = 1024;                       // hit tiles listed per round (more: further rounds)
    __shared__ unsigned s_key[M_WARPS], s_p[M_WARPS];
    __shared__ int s_tl[TL_CAP];
    __shared__ int s_ntl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntile = (N + 31) >> 5, nsuper = (N + 1023) >> 10;
    // one block per GT box at a time (a single wave of blocks walks the batch's boxes); its warps share the super-tiles round robin
    for (int64_t g = blockIdx.x; g < total_gt; g += gridDim.x) {
        const float *row = gt + 5 * g;
        const float4 a = make_float4(row[0], row[1], row[2], row[3]);
        const float area_a = (a.z - a.x) * (a.w - a.y);
        const float wg = a.z - a.x, hg = a.w - a.y;
        const bool prune = wg > 0.0f && hg > 0.0f;     // (NaN: no)
        unsigned bestkey = 0u, bestp = 0xffffffffu;
        auto eval = [&](const int64_t p, const float4 pr) {
            const float hw = pr.z / 2.0f, hh = pr.w / 2.0f;                    // point_form, box_utils.py:15-16
            const float4 pf = make_float4(pr.x - hw, pr.y - hh, pr.x + hw, pr.y + hh);
            const unsigned key = fdt_float_key(iou_match(a, area_a, pf, (pf.z - pf.x) * (pf.w - pf.y)));
            if (bestp == 0xffffffffu || key > bestkey || (key == bestkey && (unsigned)p < bestp)) { bestkey = key; bestp = (unsigned)p; }   // first prior wins ties
        };
        // touches: the clamped overlap with a box is not empty (NaN keeps)
        auto touches = [&](const float4 bb) {
            const float wbb = fminf(a.z, bb.z) - fmaxf(a.x, bb.x), hbb = fminf(a.w, bb.w) - fmaxf(a.y, bb.y);
            return !(wbb <= 0.0f || hbb <= 0.0f);
        };
        // d = (wmin, wmax, hmin, hmax) of the priors: inter <= min(wmax, wg) * min(hmax, hg) =: I and union >= wmin * hmin + area - I
        auto bound_of = [&](const float4 d) {
            if (!prune || !(d.x > 0.0f)) return 1.0f;
            const float I = fminf(d.y, wg) * fminf(d.w, hg);
            const float u = d.x * d.z + area_a - I;
            return u > I ? I / u : 1.0f;
        };
        // best first: the tiles are visited in four classes of their bound, [0.5, inf), [0.25, 0.5), [0.125, 0.25), the rest; a
        // tile whose bound is below the best overlap found so far (b1) is skipped, and once b1 reaches the upper end of the next
        // class nothing that is left can beat it
        float b1 = 0.0f;
        bool have = false;
        for (int pass = 0; pass < 4; ++pass) {
            const float lo = pass == 0 ? 0.5f : pass == 1 ? 0.25f : pass == 2 ? 0.125f : -1.0f;
            const float hi = pass == 0 ? 3.0e38f : pass == 1 ? 0.5f : pass == 2 ? 0.25f : 0.125f;
            if (have && hi * 1.00001f < b1) break;                              // (block-uniform)
            for (int64_t s0 = 0; s0 < nsuper; s0 += 32 * M_WARPS) {
                // list the hit tiles of this round of super-tiles (warps own super-tiles round robin) ...
                if (threadIdx.x == 0) s_ntl = 0;
                __syncthreads();
                const int64_t sidx = s0 + warp + M_WARPS * lane;
                bool shit = false;
                if (sidx < nsuper) {
                    const float sb = bound_of(__ldg(super_dim + sidx));
                    shit = (sidx == 0 && pass == 0) || (touches(__ldg(super_bb + sidx)) && sb >= lo && !(have && sb * 1.00001f < b1));
                }
                unsigned sm = __ballot_sync(0xffffffffu, shit);
                while (sm) {
                    const int64_t t0 = (s0 + warp + M_WARPS * (__ffs(sm) - 1)) << 5;    // the 32 tiles of this super-tile, one per lane
                    sm &= sm - 1;
                    const int64_t t = t0 + lane;
                    bool hit = false;
                    if (t < ntile) {
                        const float bound = bound_of(__ldg(tile_dim + t));
                        if (t == 0) hit = pass == 0;
                        else hit = touches(__ldg(tile_bb + t)) && bound >= lo && bound < hi && !(have && bound * 1.00001f < b1);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&s_ntl, __popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        const int e = base + __popc(m & ((1u << lane) - 1u));
                        if (hit && e < TL_CAP) s_tl[e] = (int)t;
                    }
                }
                __syncthreads();
                // ... and evaluate them, all warps together, two tiles in flight per warp.  (Tiles beyond the list's capacity --
                // more than 1,024 hit tiles in one round -- are evaluated by their owner below.)
                const int ntl = min(s_ntl, TL_CAP);
                for (int e = warp; e < ntl; e += 2 * M_WARPS) {
                    const bool two = e + M_WARPS < ntl;
                    const int64_t p1 = ((int64_t)s_tl[e] << 5) + lane, p2 = two ? ((int64_t)s_tl[e + M_WARPS] << 5) + lane : p1;
                    const float4 pr1 = __ldg(priors + min(p1, N - 1)), pr2 = __ldg(priors + min(p2, N - 1));
                    // (ascending order within a lane is not kept here: ties are settled by the explicit index comparison)
                    if (p1 < N) eval(p1, pr1);
                    if (two && p2 < N) eval(p2, pr2);
                }
                if (s_ntl > TL_CAP) {                  // (block-uniform) overflow: every warp re-walks its own super-tiles
                    unsigned sm2 = __ballot_sync(0xffffffffu, shit);
                    while (sm2) {
                        const int64_t t0 = (s0 + warp + M_WARPS * (__ffs(sm2) - 1)) << 5;
                        sm2 &= sm2 - 1;
                        const int64_t t = t0 + lane;
                        bool hit = false;
                        if (t < ntile) {
                            const float bound = bound_of(__ldg(tile_dim + t));
                            if (t == 0) hit = pass == 0;
                            else hit = touches(__ldg(tile_bb + t)) && bound >= lo && bound < hi && !(have && bound * 1.00001f < b1);
                        }
                        unsigned m = __ballot_sync(0xffffffffu, hit);
                        while (m) {
                            const int64_t p1 = ((t0 + __ffs(m) - 1) << 5) + lane;
                            m &= m - 1;
                            if (p1 < N) eval(p1, __ldg(priors + p1));
                        }
                    }
                }
                __syncthreads();
            }
            // the block's best so far: its key bounds the next pass, its prior is the result after the last one
            const unsigned mx = __reduce_max_sync(0xffffffffu, bestp == 0xffffffffu ? 0u : bestkey);
            const unsigned pm = __reduce_min_sync(0xffffffffu, (bestp != 0xffffffffu && bestkey == mx) ? bestp : 0xffffffffu);
            __syncthreads();                           // (the previous pass's values have been read)
            if (lane == 0) { s_key[warp] = mx; s_p[warp] = pm; }
            __syncthreads();
            unsigned bk = 0u;
            have = false;
#pragma unroll
            for (int w = 0; w < M_WARPS; ++w) if (s_p[w] != 0xffffffffu) { bk = have ? max(bk, s_key[w]) : s_key[w]; have = true; }
            b1 = have ? fdt_key_float(bk) : 0.0f;
        }
        if (threadIdx.x == 0) {
            unsigned bk = 0u, bp = 0xffffffffu;
            for (int w = 0; w < M_WARPS; ++w) {                                 // a lower prior index wins ties
                const unsigned k = s_key[w], q = s_p[w];
                if (q != 0xffffffffu && (bp == 0xffffffffu || k > bk || (k == bk && q < bp))) { bk = k; bp = q; }
            }
            bestprior[g] = bp;
        }
        __syncthreads();                               // s_key / s_p are reused by the next box
    }
}

// The forward in ONE kernel per prior: match (k_match_default's core), label + encode of the positives, and -- once
// k_mbl_prepare's maximum is there -- the loss terms (multibox_loss.py:96-110) and the mining histogram on the values still in registers.
// The matcher is bound by instruction issue and the loss terms by fp64 latency; in one kernel the warps of both phases share every
// SM, and conf_t / loc_t are not read back.  BIP: the bipartite matcher (box_utils.py:103-162) -- the same per-prior arg max (the
// area bound stays valid: a forced match overrides whatever it found, every other prior is labelled by `best >= thr` alone), then
// the priors that k_best_prior found for the image's GT boxes are forced to their box with overlap 2 (:150-154, the last GT wins).
template <bool BIP>
__global__ void __launch_bounds__(M_THREADS, 5)
k_match_loss(const float4 *__restrict__ priors, const float *__restrict__ gt, const int64_t *__restrict__ gt_off,
             int64_t N, float thr, float v0, float v1, float4 *__restrict__ loc_t, int64_t *__restrict__ conf_t,
             const float4 *__restrict__ loc, const float *__restrict__ conf, int C, LossAcc *__restrict__ acc,
             const unsigned *__restrict__ gmax_part, float *__restrict__ loss_c_all, int32_t *__restrict__ num_pos, int *__restrict__ hist,
             const unsigned *__restrict__ bestprior)
{
    constexpr int FORCED_CAP = 64;
    __shared__ GtEntry s_gt[GT_TILE];
    __shared__ unsigned char s_wl[M_WARPS][GT_TILE];
    __shared__ double s_tot;
    __shared__ int s_cnt;
    __shared__ float s_xmax;
    __shared__ int s_nf, s_fj[BIP ? FORCED_CAP : 1], s_fp[BIP ? FORCED_CAP : 1];
    const MatchBlock mb = match_block();
    const int b = mb.b, tid = threadIdx.x, warp = tid >> 5;
    const int64_t p = (int64_t)mb.tile * M_THREADS + tid;
    const bool valid = p < N;
    const int64_t g0 = gt_off[b];
    const int G = (int)(gt_off[b + 1] - g0);
    const int64_t t = (int64_t)b * N + p;
    if (tid == 0) { s_tot = 0.0; s_cnt = 0; s_nf = 0; }
    const float4 pr = priors[valid ? p : 0];
    float2 cv = make_float2(0.f, 0.f);                 // the row of the loss phase, in flight during the match
    if (C == 2 && valid) cv = __ldg(reinterpret_cast<const float2 *>(conf + t * 2));
    int64_t label = 0;
    float4 enc = make_float4(0.f, 0.f, 0.f, 0.f);
    float best = 0.0f;
    int bi = 0;
    if (G > 0) {                                       // G <= 0: the reference raises (Q3); defined: all background
        const float hw = pr.z / 2.0f, hh = pr.w / 2.0f;    // point_form, box_utils.py:15-16
        const float4 pf = make_float4(pr.x - hw, pr.y - hh, pr.x + hw, pr.y + hh);
        const float area_b = (pf.z - pf.x) * (pf.w - pf.y);
        match_default_core<true>(gt, g0, G, pf, area_b, valid, thr, s_gt, s_wl[warp], best, bi);
    }
    auto finalize = [&]() {
        if (G > 0 && !(best < thr)) {                  // box_utils.py:205-206 (a NaN overlap stays positive, as there)
            float4 m;
            float lab;
            if (G <= GT_TILE) { m = s_gt[bi].box; lab = s_gt[bi].label; }       // the only tile is still staged
            else { const float *row = gt + 5 * (g0 + bi); m = make_float4(row[0], row[1], row[2], row[3]); lab = row[4]; }
            label = (int64_t)(lab + 1.0f);             // :205, :210 (float -> long)
            // :208 encodes every prior; the loss only reads the positives (multibox_loss.py:96-101): zeros elsewhere
            if (label > 0) enc = fdt_encode1(m, pr, v0, v1);
        }
    };
    int is_pos = 0;
    float4 a = enc;
    if (!BIP) {
        finalize();
        is_pos = valid && label > 0;
        if (is_pos) a = loc[t];
        if (valid) { conf_t[t] = label; loc_t[t] = enc; }
    }
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();                   // k_mbl_prepare: zeroed state + the partial maxima (BIP: k_best_prior too)
    if (warp == 0) {
        const float x = gmax_from_partials(gmax_part);
        if (tid == 0) s_xmax = x;
    }
    if (BIP && G > 0) {
        // the GT boxes whose best prior lies in this block (usually none, a handful at most)
        const unsigned p0 = (unsigned)((int64_t)mb.tile * M_THREADS);
        for (int j = tid; j < G; j += M_THREADS) {
            const unsigned d = __ldcg(bestprior + g0 + j) - p0;
            if (d < (unsigned)M_THREADS) {
                const int e = atomicAdd(&s_nf, 1);
                if (e < FORCED_CAP) { s_fj[e] = j; s_fp[e] = (int)d; }
            }
        }
    }
    __syncthreads();
    if (BIP) {
        if (G > 0) {
            const int nf = s_nf;
            int fj = -1;
            if (nf <= FORCED_CAP) {
                for (int e = 0; e < nf; ++e) if (s_fp[e] == tid) fj = max(fj, s_fj[e]);
            } else {
                for (int j = 0; j < G; ++j) if (__ldcg(bestprior + g0 + j) == (unsigned)p) fj = j;
            }
            if (fj >= 0) { best = 2.0f; bi = fj; }     // :150-154
        }
        finalize();
        is_pos = valid && label > 0;
        a = enc;
        if (is_pos) a = loc[t];
        if (valid) { conf_t[t] = label; loc_t[t] = enc; }
    }
    const float xmax = s_xmax;
    double sl = 0.0;
    if (valid) {
        const float lc = prior_loss_terms(a, enc, is_pos, cv, conf + t * C, C, label, xmax, sl);
        loss_c_all[t] = lc;
        atomicAdd(&hist[(size_t)b * MINE_BINS + mine_bin2(lc)], 1);
    }
    flush_positives(sl, is_pos, &s_tot, &s_cnt, acc, &num_pos[b]);
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

// standalone mining entry: the mask of the hard negatives (multibox_loss.py:116)
constexpr int APPLY_TILES = 8;
__global__ void __launch_bounds__(M_THREADS)
k_mine_apply(const float *__restrict__ loss_c, const unsigned long long *__restrict__ cutoff, int64_t N, uint8_t *__restrict__ out_mask)
{
    fdt_pdl_enter();
    const int b = blockIdx.y;
    const unsigned long long cut = cutoff[b];
#pragma unroll 2
    for (int u = 0; u < APPLY_TILES; ++u) {
        const int64_t p = ((int64_t)blockIdx.x * APPLY_TILES + u) * M_THREADS + threadIdx.x;
        if (p < N) {
            const int64_t t = (int64_t)b * N + p;
            out_mask[t] = (uint8_t)(cut != ~0ull && mine_comp(loss_c[t], (unsigned)p) >= cut);
        }
    }
}

// fp64 exp / log of the cross entropy below: the lean evaluations of fdt_common.cuh without the final rounding to fp32 (a few
// fp64 ulps of error; the loss is a sum of thousands of such terms compared at 1e-5) -- the library routines are several times
// longer, and this kernel is a chain of dependent latencies
__device__ __forceinline__ double lean_exp_d(const float x)
{
    if (x <= 0.0f && x >= -87.0f) {
        const double xd = (double)x;
        double t = fma(xd, FDT_LN2[0], FDT_LN2[1]);
        const int k = __double2loint(t);
        t -= FDT_LN2[1];
        double r = fma(t, -FDT_LN2[2], xd);
        r = fma(t, -FDT_LN2[3], r);
        double p = FDT_EXP_C[13];
#pragma unroll
        for (int n = 12; n >= 0; --n) p = fma(p, r, FDT_EXP_C[n]);
        return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    }
    return exp((double)x);
}
__device__ __forceinline__ double lean_log_d(const double s)
{
    const int hi = __double2hiint(s);
    if ((unsigned)(hi - 0x00100000) < 0x7fe00000u) {         // positive, normal, finite
        int e = (hi >> 20) - 1023;
        double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(s));
        if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
        const double f = m - 1.0, den = m + 1.0;
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
        double er = fma(-den, y, 1.0);
        y = fma(y, er, y);
        er = fma(-den, y, 1.0);
        y = fma(y, er, y);
        double u = f * y;
        u = fma(fma(-den, u, f), y, u);
        const double u2 = u * u;
        double q = FDT_LOG_C[9];
#pragma unroll
        for (int n = 8; n >= 0; --n) q = fma(q, u2, FDT_LOG_C[n]);
        const double ed = (double)e;
        double res = fma(u, u2 * q, u + u);
        res = fma(ed, FDT_LN2[3], res);
        return fma(ed, FDT_LN2[2], res);
    }
    return log(s);
}
// cross entropy of one prior row, F.cross_entropy (sum) in fp64 (multibox_loss.py:128)
__device__ __forceinline__ double ce_row(const float *__restrict__ row, const int C, const int64_t label)
{
    if (C == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(row));
        const float m = fmaxf(v.x, v.y);
        const double s = lean_exp_d(v.x - m) + lean_exp_d(v.y - m);
        return (lean_log_d(s) + (double)m) - (double)(label == 0 ? v.x : (label == 1 ? v.y : row[label]));
    }
    float m = row[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += lean_exp_d(row[c] - m);
    return (lean_log_d(s) + (double)m) - (double)row[label];
}

// Mining + selection + cross entropy + the final division of the fused forward, one kernel (multibox_loss.py:112-135).
//   1. every block finds the cutoff BIN of its image from the 4,096-bin histogram (descending scan; num_neg = min(ratio * num_pos,
//      N - 1), :115): bins above it are hard negatives, bins below are not;
//   2. it classifies its 2,048 priors (loaded before the scan), writes the selection mask (pos | neg, :119-120), lists the selected
//      rows and sums their cross entropy; the composites of the priors IN the cutoff bin go to the image's candidate list;
//   3. the LAST block of an image (a ticket) settles the candidates: the need0 largest composites (loss key, then lower prior index
//      -- the order of the reference's stable descending sort) among them are negatives too;
//   4. the last of those B blocks divides by N (:130-135).
// No separate select and final kernels, and the row is read once.
__global__ void __launch_bounds__(M_THREADS)
k_mine_apply2(const float *__restrict__ loss_c, const int *__restrict__ hist, const int32_t *__restrict__ num_pos,
              const int64_t *__restrict__ conf_t, const float *__restrict__ conf, const int64_t N, const int C, const int B,
              const int negpos_ratio, uint8_t *__restrict__ out_mask, LossAcc *__restrict__ acc, int *__restrict__ img_ticket,
              int *__restrict__ cand_cnt, unsigned long long *__restrict__ cand, float *__restrict__ losses, float *__restrict__ norm)
{
    fdt_pdl_enter();
    __shared__ double s_red[M_WARPS];
    __shared__ __align__(8) int s_list[APPLY_TILES * M_THREADS];        // selected priors of the block; composites in step 3
    __shared__ int s_h[256];
    __shared__ int s_warp[M_WARPS];
    __shared__ int s_sel[2];
    __shared__ int s_n, s_flag;
    __shared__ long long s_ntot;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the block's rows first: their latency overlaps the histogram scan
    float lcv[APPLY_TILES];
    bool posv[APPLY_TILES];
#pragma unroll
    for (int u = 0; u < APPLY_TILES; ++u) {
        const int64_t p = ((int64_t)blockIdx.x * APPLY_TILES + u) * M_THREADS + tid;
        lcv[u] = 0.0f; posv[u] = false;
        if (p < N) { lcv[u] = loss_c[(int64_t)b * N + p]; posv[u] = conf_t[(int64_t)b * N + p] > 0; }
    }
    long long num_neg = (long long)negpos_ratio * num_pos[b];                   // multibox_loss.py:115
    if (num_neg > N - 1) num_neg = N - 1;
    const bool mining = num_neg > 0;                                            // (block-uniform)
    if (tid == 0) s_n = 0;
    if (warp == M_WARPS - 1) {                                                  // the batch's positives, for whichever block ends up last
        long long n = 0;
        for (int i = lane; i < B; i += 32) n += num_pos[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) s_ntot = n;
    }
    // ---- 1. cutoff bin d0 and the number need0 of its members that are selected
    int d0 = MINE_BINS, need0 = 0;
    if (mining) {
        constexpr int PER = MINE_BINS / M_THREADS;                              // 16 bins per thread, thread 0 holds the top ones
        const int base = MINE_BINS - PER * (tid + 1);
        const int4 *hp = reinterpret_cast<const int4 *>(hist + (size_t)b * MINE_BINS + base);
        int c[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER / 4; ++q) { const int4 v = hp[q]; c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w; }
#pragma unroll
        for (int q = 0; q < PER; ++q) sum += c[q];
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int run = inc - sum;
#pragma unroll
        for (int w = 0; w < M_WARPS; ++w) if (w < warp) run += s_warp[w];
        const int need = (int)num_neg;
#pragma unroll
        for (int q = PER - 1; q >= 0; --q) {                                    // descending bins
            if (run < need && run + c[q] >= need) { s_sel[0] = base + q; s_sel[1] = need - run; }
            run += c[q];
        }
        __syncthreads();
        d0 = s_sel[0]; need0 = s_sel[1];
    } else {
        __syncthreads();
    }
    // ---- 2. classify, mask, lists
#pragma unroll
    for (int u = 0; u < APPLY_TILES; ++u) {
        const int64_t p = ((int64_t)blockIdx.x * APPLY_TILES + u) * M_THREADS + tid;
        bool sel = false, cnd = false;
        if (p < N) {
            const int bin = mine_bin2(lcv[u]);
            sel = bin > d0 || posv[u];                                          // (d0 = MINE_BINS without mining)
            cnd = bin == d0;
            // the mask byte of a candidate is written by the settling block alone (no ordering between two writers needed)
            if (!cnd || sel) out_mask[(int64_t)b * N + p] = (uint8_t)sel;
        }
        // a few percent of the priors are selected: list them and evaluate the fp64 cross entropy over the dense list below
        const unsigned bal = __ballot_sync(0xffffffffu, sel);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_n, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (sel) s_list[base + __popc(bal & ((1u << lane) - 1u))] = (int)p | (posv[u] ? (int)0x80000000 : 0);
        }
        const unsigned cbal = __ballot_sync(0xffffffffu, cnd);
        if (cbal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&cand_cnt[b], __popc(cbal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (cnd) cand[(size_t)b * N + base + __popc(cbal & ((1u << lane) - 1u))] = mine_comp(lcv[u], (unsigned)p);
            __threadfence();                                                    // the list entries, before the ticket below
        }
    }
    __syncthreads();
    // the image's ticket is taken now: the settling block works while the others still sum their cross entropy
    if (tid == 0) s_flag = atomicAdd(&img_ticket[b], 1) == (int)gridDim.x - 1;
    double ce = 0.0;
    {
        const int n = s_n;
        for (int e = tid; e < n; e += M_THREADS) {
            const int v = s_list[e];
            const int64_t t = (int64_t)b * N + (v & 0x7fffffff);
            ce += ce_row(conf + t * C, C, v < 0 ? conf_t[t] : 0);
        }
    }
    __syncthreads();
    if (s_flag && mining) {
    // ---- 3. the last block of the image settles the cutoff bin
        __threadfence();
        const int n = *(volatile int *)&cand_cnt[b];
        const unsigned long long *cl = cand + (size_t)b * N;
        unsigned long long cutoff = 0ull;                                       // need0 == n: every candidate
        if (need0 < n && n < APPLY_TILES * M_THREADS / 2) {
            // rank among the candidates in shared memory: selected <=> fewer than need0 composites are larger
            unsigned long long *s_comp = reinterpret_cast<unsigned long long *>(s_list);
            for (int e = tid; e < n; e += M_THREADS) s_comp[e] = __ldcg(cl + e);
            __syncthreads();
            for (int e = tid; e < n; e += M_THREADS) {
                const unsigned long long c = s_comp[e];
                int rank = 0;
                for (int j = 0; j < n; ++j) rank += s_comp[j] > c;
                if (rank == need0 - 1) s_comp[n] = c;                           // composites are unique: exactly one writer
            }
            __syncthreads();
            cutoff = s_comp[n];
        } else if (need0 < n) {
            // (degenerate rows, e.g. thousands of equal losses) MSB radix select over the list, 8 bits per pass
            unsigned long long prefix = 0ull, pmask = 0ull;
            int need = need0;
            for (int shift = 56; shift >= 0; shift -= 8) {
                s_h[tid] = 0;
                __syncthreads();
                for (int e = tid; e < n; e += M_THREADS) {
                    const unsigned long long c = __ldcg(cl + e);
                    if ((c & pmask) == prefix) atomicAdd(&s_h[(int)((c >> shift) & 255ull)], 1);
                }
                __syncthreads();
                if (tid == 0) {
                    int run = 0, d = 255;
                    for (; d > 0; --d) { if (run + s_h[d] >= need) break; run += s_h[d]; }
                    s_sel[0] = d; s_sel[1] = need - run;
                }
                __syncthreads();
                prefix |= (unsigned long long)s_sel[0] << shift;
                pmask |= 255ull << shift;
                need = s_sel[1];
                __syncthreads();
            }
            cutoff = prefix;
        }
        for (int e = tid; e < n; e += M_THREADS) {
            const unsigned long long c = __ldcg(cl + e);
            const int64_t t = (int64_t)b * N + (unsigned)~(unsigned)c;
            // positives are selected (and written) already; their loss is 0, so they can only be candidates when the cutoff bin is 0
            if (d0 == 0 && conf_t[t] > 0) continue;
            const bool take = c >= cutoff;
            out_mask[t] = (uint8_t)take;
            if (take) ce += ce_row(conf + t * C, C, 0);
        }
    }
    ce = block_sum<double>(ce, s_red);
    // ---- 4. the last block of the grid finishes (multibox_loss.py:130-135)
    if (tid == 0) {
        if (ce != 0.0) atomicAdd(&acc->loss_c, ce);
        __threadfence();
        s_flag = atomicAdd(&acc->done, 1u) == gridDim.x * gridDim.y - 1u;
    }
    if (tid == 0 && s_flag) {                          // (thread 0 set the flag itself)
        __threadfence();
        const long long n = s_ntot;
        double Nn = (double)n;                         // :130
        if (n == 0) Nn = (double)B;                    // :132-133
        losses[0] = (float)(*(volatile double *)&acc->loss_l / Nn);
        losses[1] = (float)(*(volatile double *)&acc->loss_c / Nn);
        norm[0] = (float)Nn;
    }
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
                 void *ws, cudaStream_t st, bool encode_all = true)
{
    dim3 grid((unsigned)((N + M_THREADS - 1) / M_THREADS), (unsigned)B);
    MatchWs m = plan_match_ws(ws, B, N, total_gt);
    if (!bipartite) {
        FDT_CUDA(launch_pdl(k_match_default, match_grid(B, N), dim3(M_THREADS), st, (const float4 *)priors, gt, gt_off, N, thr, v0, v1, (float4 *)loc_t, conf_t,
                                                    bti, bto, encode_all));
        FDT_LAUNCH_CHECK();
    } else {
        FDT_CUDA(cudaMemsetAsync(m.bestprior, 0, (size_t)(total_gt > 0 ? total_gt : 1) * 8, st));
        FDT_CUDA(launch_pdl(k_match<true>, match_grid(B, N), dim3(M_THREADS), st, (const float4 *)priors, gt, gt_off, N, thr, v0, v1, (float4 *)loc_t, conf_t,
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

int launch_mine_tail(const float *loss_c, const MineWs &m, int B, int64_t N, int ratio, uint8_t *mask, cudaStream_t st)
{
    FDT_CUDA(launch_pdl(k_mine_select, dim3(B), dim3(MINE_THREADS), st, loss_c, (const int *)m.hist, (const int32_t *)m.num_pos, N, ratio, m.cutoff));
    FDT_LAUNCH_CHECK();
    dim3 grid((unsigned)((N + M_THREADS * APPLY_TILES - 1) / (M_THREADS * APPLY_TILES)), (unsigned)B);
    FDT_CUDA(launch_pdl(k_mine_apply, grid, dim3(M_THREADS), st, loss_c, (const unsigned long long *)m.cutoff, N, mask));
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
    FDT_REQUIRE(B >= 0 && N >= 0 && N <= 65535ll * M_THREADS && B <= 65535, FDT_E_INVALID, "%s: bad sizes B=%d N=%lld", who, B, (long long)N);
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
    return launch_mine_tail(loss_c, m, B, N, negpos_ratio, neg, st);
}

// workspace of the fused forward: [zeroed per call by k_mbl_prepare: accumulators | num_pos, image tickets, candidate counts |
// mining histogram] [partial maxima] [candidate lists B x N] [loss_c] [matcher scratch]
struct LossWs {
    LossAcc *acc; int32_t *num_pos; int *img_ticket; int *cand_cnt; int *hist; size_t zero_bytes;
    unsigned *gmax_part; unsigned long long *cand; float *loss_c_all; float4 *tile_bb, *super_bb, *tile_dim, *super_dim; unsigned *bestprior; size_t bytes;
};
static LossWs plan_loss_ws(void *ws, int B, int64_t N, int64_t total_gt, int bipartite)
{
    LossWs w;
    char *p = (char *)ws;
    size_t o = 0;
    w.acc = (LossAcc *)(p + o); o += 256;
    w.num_pos = (int32_t *)(p + o); o += fdt_align256((size_t)B * 4);
    w.img_ticket = (int *)(p + o); o += fdt_align256((size_t)B * 4);
    w.cand_cnt = (int *)(p + o); o += fdt_align256((size_t)B * 4);
    w.hist = (int *)(p + o); o += fdt_align256((size_t)B * MINE_BINS * 4);
    w.zero_bytes = o;
    w.gmax_part = (unsigned *)(p + o); o += fdt_align256((size_t)PREP_BLOCKS * 4);
    w.cand = (unsigned long long *)(p + o); o += fdt_align256((size_t)B * N * 8);
    w.loss_c_all = (float *)(p + o); o += fdt_align256((size_t)B * N * 4);
    w.tile_bb = (float4 *)(p + o); o += bipartite ? fdt_align256((size_t)((N + 31) / 32) * 16) : 0;      // bipartite: 32-prior tile boxes
    w.tile_dim = (float4 *)(p + o); o += bipartite ? fdt_align256((size_t)((N + 31) / 32) * 16) : 0;
    w.super_bb = (float4 *)(p + o); o += bipartite ? fdt_align256((size_t)((N + 1023) / 1024) * 16) : 0;
    w.super_dim = (float4 *)(p + o); o += bipartite ? fdt_align256((size_t)((N + 1023) / 1024) * 16) : 0;
    w.bestprior = (unsigned *)(p + o); o += bipartite ? fdt_align256((size_t)(total_gt > 0 ? total_gt : 1) * 4) : 0;
    w.bytes = o;
    return w;
}

FDT_API size_t fdt_multibox_workspace_bytes(int B, int64_t N, int, int64_t total_gt)
{
    if (B <= 0 || N <= 0) return 256;
    return plan_loss_ws(nullptr, B, N, total_gt, 1).bytes;
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
    FDT_REQUIRE(C != 2 || fdt_aligned(conf, 8), FDT_E_INVALID, "fdt_multibox_loss_forward: conf needs 8-byte alignment");
    FDT_REQUIRE(total_gt >= 0 && negpos_ratio >= 0, FDT_E_INVALID, "fdt_multibox_loss_forward: negative total_gt / negpos_ratio");
    LossWs w = plan_loss_ws(ws, B, N, total_gt, bipartite);
    FDT_REQUIRE(ws_bytes >= w.bytes, FDT_E_WORKSPACE, "fdt_multibox_loss_forward: workspace %zu < %zu bytes", ws_bytes, w.bytes);
    float *lca = loss_c_all ? loss_c_all : w.loss_c_all;

    // three kernels (four with the bipartite matcher) chained by programmatic dependent launch:
    // prepare -> [best prior per GT box ->] match + loss terms -> mining
    FDT_CUDA(launch_pdl(k_mbl_prepare, dim3(PREP_BLOCKS), dim3(PREP_THREADS), st, conf, (int64_t)B * N * C, w.gmax_part, (int4 *)ws,
                        (int64_t)(w.zero_bytes / 16), (const float4 *)priors, N, bipartite ? w.tile_bb : (float4 *)nullptr, w.tile_dim,
                        w.super_bb, w.super_dim));
    FDT_LAUNCH_CHECK();
    if (!bipartite) {
        FDT_CUDA(launch_pdl(k_match_loss<false>, match_grid(B, N), dim3(M_THREADS), st, (const float4 *)priors, gt, gt_off, N, threshold, var0, var1,
                            (float4 *)loc_t, conf_t, (const float4 *)loc, conf, C, w.acc, (const unsigned *)w.gmax_part, lca, w.num_pos, w.hist,
                            (const unsigned *)nullptr));
        FDT_LAUNCH_CHECK();
    } else {
        if (total_gt > 0) {
            const unsigned nb = (unsigned)(total_gt < FDT_NUM_SMS * 8 ? total_gt : FDT_NUM_SMS * 8);          // one wave
            FDT_CUDA(launch_pdl(k_best_prior, dim3(nb), dim3(M_THREADS), st, (const float4 *)priors, N, gt,
                                total_gt, (const float4 *)w.tile_bb, (const float4 *)w.tile_dim, (const float4 *)w.super_bb,
                                (const float4 *)w.super_dim, w.bestprior));
            FDT_LAUNCH_CHECK();
        }
        FDT_CUDA(launch_pdl(k_match_loss<true>, match_grid(B, N), dim3(M_THREADS), st, (const float4 *)priors, gt, gt_off, N, threshold, var0, var1,
                            (float4 *)loc_t, conf_t, (const float4 *)loc, conf, C, w.acc, (const unsigned *)w.gmax_part, lca, w.num_pos, w.hist,
                            (const unsigned *)w.bestprior));
        FDT_LAUNCH_CHECK();
    }
    dim3 agrid((unsigned)((N + M_THREADS * APPLY_TILES - 1) / (M_THREADS * APPLY_TILES)), (unsigned)B);
    FDT_CUDA(launch_pdl(k_mine_apply2, agrid, dim3(M_THREADS), st, (const float *)lca, (const int *)w.hist, (const int32_t *)w.num_pos,
                        (const int64_t *)conf_t, conf, N, C, B, negpos_ratio, sel, w.acc, w.img_ticket, w.cand_cnt, w.cand, losses, norm));
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
