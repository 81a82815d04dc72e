/*
 * fdt_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the SSD box pipeline of limacv/Face-detection-and-tracking.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (face-detection-and-tracking_b200/) never does.
 *
 * Every function cites the reference file:line (relative to the reference root) it restates.
 * Parity status: the reference ships NO golden vectors or tests for this path (SURVEY.md
 * section 8c).  This oracle is pinned instead against outputs of the reference's own Python
 * code executed in the build container (oracle/make_golden.py -> tests/golden/ *.npz);
 * tests/test_oracle_golden.py re-checks that pin on every CPU run.
 *
 * Arithmetic rules that make bit-parity possible:
 *   - fp32 IEEE add/sub/mul/div/min/max in the reference's operand order; build with
 *     -ffp-contract=off so no FMA is formed (torch evaluates each op as its own kernel).
 *   - exp/log (decode/encode/log_sum_exp) are evaluated in fp64 and rounded once to fp32,
 *     i.e. the correctly-rounded fp32 value up to a 2^-29 double-rounding chance.  torch's CPU
 *     (SLEEF) and CUDA (expf) results are each within 1 ulp of it; coordinates that went through
 *     exp/log are therefore compared at 1e-5 relative, everything else bit-exact.
 *   - sort ties (unspecified in the reference: torch.sort is unstable) are DEFINED here as
 *     "stable ascending, consumed from the end" for NMS (higher candidate index first among equal
 *     scores) and "stable descending" for hard-negative mining (lower prior index first).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static inline float f_exp(float x) { return (float)exp((double)x); }
static inline float f_log(float x) { return (float)log((double)x); }
/* torch.min/max/clamp propagate NaN; C fminf/fmaxf do not. */
static inline float f_min(float a, float b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
static inline float f_max(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
static inline double d_min(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
static inline double d_max(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }

ORC_API int orc_version(void) { return 1; }

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * PriorBoxLayer.__call__  (layers/functions/prior_box.py:28-44)
 * python-float (fp64) arithmetic, one rounding to fp32 at torch.Tensor(mean) (:43).
 * box_scale[s] = (2 ** (1/3)) ** s and sqrt_ar[a] = sqrt(ar) are evaluated by the caller with
 * python floats exactly as :33 and :41 do, so libm pow differences cannot enter.
 * Layout: i (y) outer, j (x) inner, then scale, then [plain box, one box per aspect ratio].
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_priorbox(double width, double height, double stride, double box,
                          int n_scales, const double *box_scale, int n_ar, const double *sqrt_ar,
                          int f_w, int f_h, float *out)
{
    size_t o = 0;
    for (int i = 0; i < f_h; ++i)
        for (int j = 0; j < f_w; ++j)
            for (int s = 0; s < n_scales; ++s) {
                double cx = (j + 0.5) * stride / width;        /* :34 */
                double cy = (i + 0.5) * stride / height;       /* :35 */
                double sx = box * box_scale[s] / width;        /* :36 */
                double sy = box * box_scale[s] / height;       /* :37 */
                out[o++] = (float)cx; out[o++] = (float)cy; out[o++] = (float)sx; out[o++] = (float)sy;
                for (int a = 0; a < n_ar; ++a) {               /* :40-41 */
                    out[o++] = (float)cx; out[o++] = (float)cy;
                    out[o++] = (float)(sx / sqrt_ar[a]); out[o++] = (float)(sy * sqrt_ar[a]);
                }
            }
}

/* point_form (layers/box_utils.py:7-16): [cx,cy,w,h] -> [x1,y1,x2,y2] */
ORC_API void orc_point_form(const float *b, int64_t n, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        const float *p = b + 4 * i; float *o = out + 4 * i;
        float hw = p[2] / 2.0f, hh = p[3] / 2.0f;
        o[0] = p[0] - hw; o[1] = p[1] - hh; o[2] = p[0] + hw; o[3] = p[1] + hh;
    }
}

/* center_size (layers/box_utils.py:19-28): [x1,y1,x2,y2] -> [cx,cy,w,h] */
ORC_API void orc_center_size(const float *b, int64_t n, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        const float *p = b + 4 * i; float *o = out + 4 * i;
        o[0] = (p[2] + p[0]) / 2.0f; o[1] = (p[3] + p[1]) / 2.0f; o[2] = p[2] - p[0]; o[3] = p[3] - p[1];
    }
}

/* intersect, normal branch (layers/box_utils.py:58-66).  The >1000 MB CPU branch (:44-56) has a
 * bug (max_xy_cpu -= max_xy_cpu) and is outside every configured size; not restated. */
static inline float inter1(const float *a, const float *b)
{
    float w = f_min(a[2], b[2]) - f_max(a[0], b[0]);
    float h = f_min(a[3], b[3]) - f_max(a[1], b[1]);
    w = f_max(w, 0.0f); h = f_max(h, 0.0f);
    return w * h;
}
ORC_API void orc_intersect(const float *a, int64_t A, const float *b, int64_t B, float *out)
{
    for (int64_t i = 0; i < A; ++i)
        for (int64_t j = 0; j < B; ++j) out[i * B + j] = inter1(a + 4 * i, b + 4 * j);
}

/* calculate_iou (layers/box_utils.py:70-100): inter / ((area_a + area_b) - inter) */
static inline float iou1(const float *a, const float *b)
{
    float inter = inter1(a, b);
    float area_a = (a[2] - a[0]) * (a[3] - a[1]);
    float area_b = (b[2] - b[0]) * (b[3] - b[1]);
    float uni = area_a + area_b - inter;
    return inter / uni;
}
ORC_API void orc_calculate_iou(const float *a, int64_t A, const float *b, int64_t B, float *out)
{
    for (int64_t i = 0; i < A; ++i)
        for (int64_t j = 0; j < B; ++j) out[i * B + j] = iou1(a + 4 * i, b + 4 * j);
}

/* encode (layers/box_utils.py:213-234) */
static inline void encode1(const float *m, const float *p, float v0, float v1, float *o)
{
    o[0] = ((m[0] + m[2]) / 2.0f - p[0]) / (v0 * p[2]);
    o[1] = ((m[1] + m[3]) / 2.0f - p[1]) / (v0 * p[3]);
    o[2] = f_log((m[2] - m[0]) / p[2]) / v1;
    o[3] = f_log((m[3] - m[1]) / p[3]) / v1;
}
ORC_API void orc_encode(const float *matched, const float *priors, int64_t n, float v0, float v1, float *out)
{
    for (int64_t i = 0; i < n; ++i) encode1(matched + 4 * i, priors + 4 * i, v0, v1, out + 4 * i);
}

/* decode (layers/box_utils.py:238-258) */
static inline void decode1(const float *l, const float *p, float v0, float v1, float *o)
{
    float cx = p[0] + (l[0] * v0) * p[2];
    float cy = p[1] + (l[1] * v0) * p[3];
    float w = p[2] * f_exp(l[2] * v1);
    float h = p[3] * f_exp(l[3] * v1);
    float x1 = cx - w / 2.0f;        /* :256 */
    float y1 = cy - h / 2.0f;
    o[0] = x1; o[1] = y1; o[2] = w + x1; o[3] = h + y1;   /* :257 */
}
ORC_API void orc_decode(const float *loc, const float *priors, int64_t n, float v0, float v1, float *out)
{
    for (int64_t i = 0; i < n; ++i) decode1(loc + 4 * i, priors + 4 * i, v0, v1, out + 4 * i);
}

/* log_sum_exp (layers/box_utils.py:261-269): GLOBAL max over the whole [R,C] tensor */
ORC_API void orc_log_sum_exp(const float *x, int64_t R, int C, float *out)
{
    float xmax = -INFINITY;
    for (int64_t i = 0; i < R * C; ++i) xmax = f_max(xmax, x[i]);
    for (int64_t r = 0; r < R; ++r) {
        float s = 0.0f;
        for (int c = 0; c < C; ++c) s += f_exp(x[r * C + c] - xmax);
        out[r] = f_log(s) + xmax;
    }
}

/* ------------------------------------------------------------------------------------------
 * nms (layers/box_utils.py:275-340)
 * keep[] has n entries, zero-initialised (:289); count returned.
 * Sort rule: ascending by (score, index) -- the reference's unspecified tie order is defined so.
 * ------------------------------------------------------------------------------------------ */
typedef struct { float s; int64_t i; } orc_si;
static int cmp_si_asc(const void *pa, const void *pb)
{
    const orc_si *a = (const orc_si *)pa, *b = (const orc_si *)pb;
    int an = a->s != a->s, bn = b->s != b->s;            /* torch.sort puts NaN last (largest) */
    if (an != bn) return an - bn;
    if (!an) { if (a->s < b->s) return -1; if (a->s > b->s) return 1; }
    return (a->i > b->i) - (a->i < b->i);
}

/* max_keep < 0: run to completion (reference behaviour).  max_keep >= 0: stop once that many are
 * kept -- identical prefix, used by orc_detect because Detect only reads keep[:top_k]. */
static int64_t nms_core(const float *boxes, const float *scores, int64_t n, float overlap, int64_t top_k,
                        int64_t max_keep, int64_t *keep)
{
    memset(keep, 0, sizeof(int64_t) * (size_t)n);
    if (n == 0) return 0;                                             /* :290-291 */
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    orc_si *ord = (orc_si *)malloc(sizeof(orc_si) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float *b = boxes + 4 * i;
        area[i] = (b[2] - b[0]) * (b[3] - b[1]);                      /* :296 */
        ord[i].s = scores[i]; ord[i].i = i;
    }
    qsort(ord, (size_t)n, sizeof(orc_si), cmp_si_asc);                /* :297 */
    int64_t m = top_k < n ? top_k : n;                                /* :299 idx[-top_k:] */
    if (top_k <= 0) m = n;                                            /* idx[-0:] is the whole list */
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)m);
    for (int64_t t = 0; t < m; ++t) idx[t] = ord[n - m + t].i;
    int64_t count = 0;
    while (m > 0) {                                                   /* :308 */
        int64_t i = idx[m - 1];
        keep[count++] = i;                                            /* :311-312 */
        if (m == 1) break;                                            /* :313-314 */
        if (max_keep >= 0 && count >= max_keep) break;
        m -= 1;                                                       /* :315 */
        const float *bi = boxes + 4 * i;
        int64_t w = 0;
        for (int64_t t = 0; t < m; ++t) {
            int64_t j = idx[t];
            const float *bj = boxes + 4 * j;
            float xx1 = f_max(bj[0], bi[0]);                          /* :322 clamp(min=x1[i]) */
            float yy1 = f_max(bj[1], bi[1]);
            float xx2 = f_min(bj[2], bi[2]);                          /* :324 clamp(max=x2[i]) */
            float yy2 = f_min(bj[3], bi[3]);
            float ww = f_max(xx2 - xx1, 0.0f);                        /* :328-332 */
            float hh = f_max(yy2 - yy1, 0.0f);
            float inter = ww * hh;
            float uni = (area[j] - inter) + area[i];                  /* :336 asymmetric order */
            float iou = inter / uni;
            if (iou < overlap) idx[w++] = j;                          /* :339 (NaN -> dropped) */
        }
        m = w;
    }
    free(idx); free(ord); free(area);
    return count;
}

ORC_API int64_t orc_nms(const float *boxes, const float *scores, int64_t n, float overlap, int64_t top_k,
                        int64_t *keep)
{
    return nms_core(boxes, scores, n, overlap, top_k, -1, keep);
}

/* ------------------------------------------------------------------------------------------
 * Detect.__call__ (layers/functions/detection.py:34-84)
 * out[B,C,top_k,5] zero-filled (:48); class 0 never written (:63 range(1, C)).
 * counts[B,C] (may be NULL) = rows written; kept_prior[B,C,top_k] (may be NULL) = prior index of
 * each written row, -1 elsewhere.  early_exit != 0 stops each NMS at top_k kept (same output).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_detect(const float *loc, const float *conf, const float *priors,
                        int B, int64_t N, int C, int top_k, int nms_top_k,
                        float conf_thresh, float nms_thresh, float v0, float v1,
                        float *out, int32_t *counts, int64_t *kept_prior, int early_exit, int n_threads)
{
    memset(out, 0, sizeof(float) * (size_t)B * C * top_k * 5);
    if (counts) memset(counts, 0, sizeof(int32_t) * (size_t)B * C);
    if (kept_prior) for (int64_t t = 0; t < (int64_t)B * C * top_k; ++t) kept_prior[t] = -1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int b = 0; b < B; ++b) {
        float *boxes = (float *)malloc(sizeof(float) * 4 * (size_t)N);
        float *cb = (float *)malloc(sizeof(float) * 4 * (size_t)N);
        float *cs = (float *)malloc(sizeof(float) * (size_t)N);
        int64_t *cp = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
        int64_t *keep = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
        orc_decode(loc + (size_t)b * N * 4, priors, N, v0, v1, boxes);            /* :55 */
        for (int cl = 1; cl < C; ++cl) {                                          /* :63 */
            int64_t n = 0;
            for (int64_t p = 0; p < N; ++p) {
                float s = conf[((size_t)b * N + p) * C + cl];
                if (s > conf_thresh) {                                            /* :64 strict gt */
                    cs[n] = s; cp[n] = p; memcpy(cb + 4 * n, boxes + 4 * p, 16); ++n;
                }
            }
            if (n == 1) continue;           /* :66-72 nonzero().squeeze() is 0-d -> `continue` */
            int64_t k = n < nms_top_k ? n : nms_top_k;                            /* :79 */
            int64_t count = (n == 0) ? 0 : nms_core(cb, cs, n, nms_thresh, k, early_exit ? top_k : -1, keep);
            if (count > top_k) count = top_k;                                     /* :80 */
            float *o = out + (((size_t)b * C + cl) * top_k) * 5;
            for (int64_t r = 0; r < count; ++r) {                                 /* :82 */
                int64_t id = keep[r];
                o[5 * r] = cs[id]; memcpy(o + 5 * r + 1, cb + 4 * id, 16);
                if (kept_prior) kept_prior[((size_t)b * C + cl) * top_k + r] = cp[id];
            }
            if (counts) counts[b * C + cl] = (int32_t)count;
        }
        free(keep); free(cp); free(cs); free(cb); free(boxes);
    }
}

/* ------------------------------------------------------------------------------------------
 * match_default (layers/box_utils.py:165-210)  /  match_ensure_max_prior (:103-162)
 * truth[G,4] corner form, labels[G] (float, always 0.0 in the reference), priors[N,4] centre form.
 * Outputs for ONE image: loc_t[N,4], conf_t[N] int64, best_truth_idx[N] int64, best_truth_overlap[N].
 * argmax ties: first (lowest) index, as torch.max(dim) does on CPU.
 * G == 0: the reference raises (Q3).  Returns -1 and writes nothing.
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_match(int bipartite, float threshold, const float *truth, const float *labels, int64_t G,
                      const float *priors, int64_t N, float v0, float v1,
                      float *loc_t, int64_t *conf_t, int64_t *best_truth_idx, float *best_truth_overlap)
{
    if (G <= 0) return -1;
    float *pf = (float *)malloc(sizeof(float) * 4 * (size_t)N);
    orc_point_form(priors, N, pf);                                               /* :194 */
    for (int64_t p = 0; p < N; ++p) {                                            /* :197 overlaps.max(0) */
        float best = iou1(truth, pf + 4 * p); int64_t bi = 0;
        for (int64_t g = 1; g < G; ++g) {
            float v = iou1(truth + 4 * g, pf + 4 * p);
            if (v > best) { best = v; bi = g; }
        }
        best_truth_idx[p] = bi; best_truth_overlap[p] = best;
    }
    if (bipartite) {
        int64_t *bp = (int64_t *)malloc(sizeof(int64_t) * (size_t)G);
        for (int64_t g = 0; g < G; ++g) {                                        /* :136 overlaps.max(1) */
            float best = iou1(truth + 4 * g, pf); int64_t bi = 0;
            for (int64_t p = 1; p < N; ++p) {
                float v = iou1(truth + 4 * g, pf + 4 * p);
                if (v > best) { best = v; bi = p; }
            }
            bp[g] = bi;
        }
        for (int64_t g = 0; g < G; ++g) best_truth_overlap[bp[g]] = 2.0f;         /* :150 */
        for (int64_t g = 0; g < G; ++g) best_truth_idx[bp[g]] = g;               /* :153-154 last j wins */
        free(bp);
    }
    for (int64_t p = 0; p < N; ++p) {
        int64_t g = best_truth_idx[p];
        float c = labels[g] + 1.0f;                                              /* :205 */
        if (best_truth_overlap[p] < threshold) c = 0.0f;                         /* :206 */
        conf_t[p] = (int64_t)c;                                                  /* :210 float -> long */
        encode1(truth + 4 * g, priors + 4 * p, v0, v1, loc_t + 4 * p);           /* :208 */
    }
    free(pf);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Hard-negative mining (layers/modules/multibox_loss.py:112-116)
 * loss_c[B,N] already zeroed at positives (:110).  rank = position in the descending sort;
 * neg = rank < clamp(ratio * num_pos, max = N-1).  Tie rule: stable (lower index first).
 * ------------------------------------------------------------------------------------------ */
typedef struct { float v; int32_t i; } orc_vi;
static int cmp_vi_desc(const void *pa, const void *pb)
{
    const orc_vi *a = (const orc_vi *)pa, *b = (const orc_vi *)pb;
    int an = a->v != a->v, bn = b->v != b->v;             /* NaN sorts first when descending */
    if (an != bn) return bn - an;
    if (!an) { if (a->v > b->v) return -1; if (a->v < b->v) return 1; }
    return (a->i > b->i) - (a->i < b->i);
}
ORC_API void orc_hard_negative_mine(const float *loss_c, const uint8_t *pos, int B, int64_t N, int negpos_ratio,
                                    uint8_t *neg, int n_threads)
{
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int b = 0; b < B; ++b) {
        orc_vi *a = (orc_vi *)malloc(sizeof(orc_vi) * (size_t)N);
        int64_t num_pos = 0;
        for (int64_t p = 0; p < N; ++p) { a[p].v = loss_c[(size_t)b * N + p]; a[p].i = (int32_t)p; num_pos += pos[(size_t)b * N + p] != 0; }
        qsort(a, (size_t)N, sizeof(orc_vi), cmp_vi_desc);                        /* :112-113 */
        int64_t num_neg = (int64_t)negpos_ratio * num_pos;                       /* :115 */
        if (num_neg > N - 1) num_neg = N - 1;
        memset(neg + (size_t)b * N, 0, (size_t)N);
        for (int64_t r = 0; r < num_neg; ++r) neg[(size_t)b * N + a[r].i] = 1;   /* :116 */
        free(a);
    }
}

/* ------------------------------------------------------------------------------------------
 * MultiBoxLoss.forward (layers/modules/multibox_loss.py:48-136)
 * gt[sum G,5] rows [x1,y1,x2,y2,label]; gt_off[B+1].  Images with G == 0 (reference raises, Q3)
 * are DEFINED as all-background with loc_t = 0.
 * Outputs: losses[2] = {loss_l/N, loss_c/N}; optional loc_t[B,N,4], conf_t[B,N] int64,
 * loss_c_all[B,N] (the mining input, zero at positives), neg[B,N].
 * Sums are accumulated in fp64 (torch sums fp32 with its own blocking; compared at 1e-5 rel).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_multibox_loss(const float *loc, const float *conf, const float *priors,
                               const float *gt, const int64_t *gt_off, int B, int64_t N, int C,
                               float threshold, int negpos_ratio, int bipartite, float v0, float v1,
                               float *losses, float *loc_t_out, int64_t *conf_t_out,
                               float *loss_c_all_out, uint8_t *neg_out, int n_threads)
{
    size_t BN = (size_t)B * N;
    float *loc_t = loc_t_out ? loc_t_out : (float *)malloc(sizeof(float) * 4 * BN);
    int64_t *conf_t = conf_t_out ? conf_t_out : (int64_t *)malloc(sizeof(int64_t) * BN);
    float *lc = loss_c_all_out ? loss_c_all_out : (float *)malloc(sizeof(float) * BN);
    uint8_t *neg = neg_out ? neg_out : (uint8_t *)malloc(BN);
    uint8_t *pos = (uint8_t *)malloc(BN);
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int b = 0; b < B; ++b) {                                                /* :69-81 */
        int64_t G = gt_off[b + 1] - gt_off[b];
        float *lt = loc_t + (size_t)b * N * 4; int64_t *ct = conf_t + (size_t)b * N;
        if (G <= 0) { memset(lt, 0, sizeof(float) * 4 * (size_t)N); memset(ct, 0, sizeof(int64_t) * (size_t)N); continue; }
        float *truth = (float *)malloc(sizeof(float) * 4 * (size_t)G);
        float *labels = (float *)malloc(sizeof(float) * (size_t)G);
        for (int64_t g = 0; g < G; ++g) { memcpy(truth + 4 * g, gt + 5 * (gt_off[b] + g), 16); labels[g] = gt[5 * (gt_off[b] + g) + 4]; }
        int64_t *bti = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
        float *bto = (float *)malloc(sizeof(float) * (size_t)N);
        orc_match(bipartite, threshold, truth, labels, G, priors, N, v0, v1, lt, ct, bti, bto);
        free(bto); free(bti); free(labels); free(truth);
    }
    /* :90-101 smooth L1 over positives */
    double loss_l = 0.0; int64_t num_pos_total = 0;
    for (size_t t = 0; t < BN; ++t) {
        pos[t] = conf_t[t] > 0;
        if (pos[t]) {
            ++num_pos_total;
            for (int k = 0; k < 4; ++k) {
                float d = fabsf(loc[4 * t + k] - loc_t[4 * t + k]);
                loss_l += d < 1.0f ? 0.5f * d * d : d - 0.5f;
            }
        }
    }
    /* :104-110 per-prior CE via log_sum_exp with GLOBAL max */
    float xmax = -INFINITY;
    for (size_t t = 0; t < BN * (size_t)C; ++t) xmax = f_max(xmax, conf[t]);
    for (size_t t = 0; t < BN; ++t) {
        float s = 0.0f;
        for (int c = 0; c < C; ++c) s += f_exp(conf[t * C + c] - xmax);
        float v = (f_log(s) + xmax) - conf[t * C + conf_t[t]];
        lc[t] = pos[t] ? 0.0f : v;
    }
    orc_hard_negative_mine(lc, pos, B, N, negpos_ratio, neg, n_threads);         /* :112-116 */
    /* :119-128 CE(sum) over pos U neg, F.cross_entropy = per-row log-softmax */
    double loss_c = 0.0;
    for (size_t t = 0; t < BN; ++t) {
        if (!(pos[t] || neg[t])) continue;
        float m = conf[t * C];
        for (int c = 1; c < C; ++c) m = f_max(m, conf[t * C + c]);
        double s = 0.0;
        for (int c = 0; c < C; ++c) s += exp((double)(conf[t * C + c] - m));
        loss_c += (log(s) + (double)m) - (double)conf[t * C + conf_t[t]];
    }
    double Nn = (double)num_pos_total;                                           /* :130 */
    if (Nn == 0) Nn = (double)B;                                                 /* :132-133 */
    losses[0] = (float)(loss_l / Nn); losses[1] = (float)(loss_c / Nn);
    free(pos);
    if (!neg_out) free(neg);
    if (!loss_c_all_out) free(lc);
    if (!conf_t_out) free(conf_t);
    if (!loc_t_out) free(loc_t);
}

/* ------------------------------------------------------------------------------------------
 * utils.calc_performance.intersect / calculate_iou  (utils/calc_performance.py:4-31, 54-74)
 * float64; np.minimum/np.maximum propagate NaN.
 * ------------------------------------------------------------------------------------------ */
static inline double iou1_f64(const double *a, const double *b)
{
    double w = d_min(a[2], b[2]) - d_max(a[0], b[0]);
    double h = d_min(a[3], b[3]) - d_max(a[1], b[1]);
    w = d_max(w, 0.0); h = d_max(h, 0.0);
    double inter = w * h;
    double area_a = (a[2] - a[0]) * (a[3] - a[1]);
    double area_b = (b[2] - b[0]) * (b[3] - b[1]);
    double uni = area_a + area_b - inter;
    return inter / uni;
}
ORC_API void orc_calculate_iou_f64(const double *a, int64_t A, const double *b, int64_t B, double *out)
{
    for (int64_t i = 0; i < A; ++i)
        for (int64_t j = 0; j < B; ++j) out[i * B + j] = iou1_f64(a + 4 * i, b + 4 * j);
}

/* utils.calc_performance.intersect (utils/calc_performance.py:4-31), float64 */
ORC_API void orc_intersect_f64(const double *a, int64_t A, const double *b, int64_t B, double *out)
{
    for (int64_t i = 0; i < A; ++i)
        for (int64_t j = 0; j < B; ++j) {
            const double *p = a + 4 * i, *q = b + 4 * j;
            double w = d_max(d_min(p[2], q[2]) - d_max(p[0], q[0]), 0.0);
            double h = d_max(d_min(p[3], q[3]) - d_max(p[1], q[1]), 0.0);
            out[i * B + j] = w * h;
        }
}

/* utils.calc_performance.calculate_distance (utils/calc_performance.py:34-51), float64; `dis ** 0.25` is libm pow */
ORC_API void orc_calculate_distance_f64(const double *a, int64_t A, const double *b, int64_t B, double *out)
{
    for (int64_t i = 0; i < A; ++i)
        for (int64_t j = 0; j < B; ++j) {
            const double *p = a + 4 * i, *q = b + 4 * j;
            double adx = p[2] - p[0], ady = p[3] - p[1], bdx = q[2] - q[0], bdy = q[3] - q[1];      /* :41-42 */
            double cax = (p[2] + p[0]) / 2, cay = (p[3] + p[1]) / 2, cbx = (q[2] + q[0]) / 2, cby = (q[3] + q[1]) / 2;
            double dx = cbx - cax, dy = cby - cay;                                                    /* :45 */
            double dz = ((adx - bdx) + (ady - bdy)) / 2;                                              /* :46-47 */
            double dis = dz * dz + dx * dx + dy * dy;                                                 /* :48 */
            out[i * B + j] = pow(dis, 0.25);                                                          /* :49 */
        }
}

/* utils.calc_performance.calc_pr (utils/calc_performance.py:77-92): truth rows [x, y, w, h] -> tf[P] = max_t IoU > thresh
 * (np.max propagates NaN, NaN > thresh is False).  Returns truth_num. */
ORC_API int64_t orc_calc_pr(const double *predict, int64_t P, int predict_stride, const double *truth, int64_t T,
                            double iou_thresh, int32_t *tf)
{
    for (int64_t j = 0; j < P; ++j) {
        double best = -INFINITY; int any_nan = 0;
        for (int64_t t = 0; t < T; ++t) {
            double tb[4] = {truth[4 * t], truth[4 * t + 1], truth[4 * t + 2] + truth[4 * t], truth[4 * t + 3] + truth[4 * t + 1]};   /* :88 */
            double v = iou1_f64(tb, predict + predict_stride * j);
            if (v != v) any_nan = 1;
            else if (v > best) best = v;
        }
        tf[j] = (!any_nan && best > iou_thresh) ? 1 : 0;                                              /* :91 */
    }
    return T;
}

/* ------------------------------------------------------------------------------------------
 * IoU tracker loop (iouTracke_cal.py:126-155 per frame, :174-176 flush), use_iou = True (metric 0) or False (metric 1).
 * dets[total,5] float64 rows [x1,y1,x2,y2,score]; frame_off[F+1]; frame numbers are 1-based (:118).
 * Output (CSR): returns T = number of finished tracks; track_off[T+1] into track_dets (global det
 * row indices, in append order); track_start[T] (1-based), track_max[T].
 * Caller sizes track_off/start/max for total+1 entries and track_dets for total entries.
 * ------------------------------------------------------------------------------------------ */
typedef struct { int64_t *d; int64_t len, cap; double max_score; int64_t start; } orc_track_t;
static void tr_push(orc_track_t *t, int64_t g)
{
    if (t->len == t->cap) { t->cap = t->cap ? t->cap * 2 : 8; t->d = (int64_t *)realloc(t->d, sizeof(int64_t) * (size_t)t->cap); }
    t->d[t->len++] = g;
}
/* association value of detection a against a track's last box b, larger = better:
 * use_iou (:131-134) the IoU, matched iff > sigma_iou; else (:135-138) minus calculate_distance, matched iff distance < sigma_dis.
 * argmin of the distance == argmax of its negation, first index on ties, first NaN wins either way (numpy). */
static double track_value(const double *a, const double *b, int metric)
{
    if (metric == 0) return iou1_f64(a, b);
    double d;
    orc_calculate_distance_f64(a, 1, b, 1, &d);
    return -d;
}
ORC_API int64_t orc_iou_track_metric(const double *dets, const int64_t *frame_off, int64_t F,
                                     int metric, double sigma, double sigma_h, int64_t t_min,
                                     int64_t *track_off, int64_t *track_dets, int64_t *track_start, double *track_max)
{
    const double sigma_iou = metric == 0 ? sigma : -sigma;
    int64_t total = frame_off[F];
    orc_track_t *active = NULL, *updated = NULL; int64_t n_active = 0;
    int64_t T = 0, w = 0; track_off[0] = 0;
    int64_t maxd = 0;
    for (int64_t f = 0; f < F; ++f) { int64_t d = frame_off[f + 1] - frame_off[f]; if (d > maxd) maxd = d; }
    int64_t *alive = (int64_t *)malloc(sizeof(int64_t) * (size_t)(maxd + 1));
#define FINISH(tr) do { for (int64_t q = 0; q < (tr).len; ++q) track_dets[w++] = (tr).d[q]; \
        track_start[T] = (tr).start; track_max[T] = (tr).max_score; ++T; track_off[T] = w; } while (0)
    for (int64_t f = 0; f < F; ++f) {
        int64_t D = frame_off[f + 1] - frame_off[f];
        int64_t n_alive = D;                                       /* :127 dets = det0.tolist() */
        for (int64_t d = 0; d < D; ++d) alive[d] = frame_off[f] + d;
        updated = (orc_track_t *)malloc(sizeof(orc_track_t) * (size_t)(n_active + D + 1));
        int64_t n_upd = 0;
        for (int64_t t = 0; t < n_active; ++t) {                   /* :129 */
            orc_track_t *tr = &active[t];
            if (n_alive > 0) {                                     /* :130 */
                const double *last = dets + 5 * tr->d[tr->len - 1];
                int64_t best = 0; double bv = track_value(dets + 5 * alive[0], last, metric);   /* :132 / :136 */
                for (int64_t a = 1; a < n_alive && !(bv != bv); ++a) {     /* :133 argmax (:137 argmin): first extremum, first NaN wins */
                    double v = track_value(dets + 5 * alive[a], last, metric);
                    if (v != v || v > bv) { bv = v; best = a; }
                }
                if (bv > sigma_iou) {                              /* :134, :140-145 */
                    int64_t g = alive[best];
                    tr_push(tr, g);
                    if (dets[5 * g + 4] > tr->max_score) tr->max_score = dets[5 * g + 4];
                    updated[n_upd++] = *tr;
                    memmove(alive + best, alive + best + 1, sizeof(int64_t) * (size_t)(n_alive - best - 1));
                    --n_alive;
                } else {                                           /* :146-148 */
                    if (tr->max_score > sigma_h && tr->len > t_min) FINISH(*tr);
                    free(tr->d);
                }
            } else {
                free(tr->d);                                       /* Q5: silently dropped */
            }
        }
        for (int64_t a = 0; a < n_alive; ++a) {                    /* :150-154 new tracks */
            orc_track_t nt = {0};
            tr_push(&nt, alive[a]); nt.max_score = dets[5 * alive[a] + 4]; nt.start = f + 1;
            updated[n_upd++] = nt;
        }
        free(active); active = updated; n_active = n_upd;          /* :155 */
    }
    for (int64_t t = 0; t < n_active; ++t) {                       /* :174-175 flush, >= t_min */
        if (active[t].max_score > sigma_h && active[t].len >= t_min) FINISH(active[t]);
        free(active[t].d);
    }
#undef FINISH
    free(active); free(alive);
    (void)total;
    return T;
}
ORC_API int64_t orc_iou_track(const double *dets, const int64_t *frame_off, int64_t F,
                              double sigma_iou, double sigma_h, int64_t t_min,
                              int64_t *track_off, int64_t *track_dets, int64_t *track_start, double *track_max)
{
    return orc_iou_track_metric(dets, frame_off, F, 0, sigma_iou, sigma_h, t_min, track_off, track_dets, track_start, track_max);
}

/* ------------------------------------------------------------------------------------------
 * Head post-processing that feeds Detect / MultiBoxLoss  (pyramid.py:291-309, 331-332; identical
 * code in pyramid_mobile_try1.py:297-327 and pyramid_mb2_try3/4/5.py).
 * Per level l: conf map [B,4,H,W] -> (neg, pos): level with neg_max[l] != 0 has neg = max(ch0..2),
 * pos = ch3 (:293-297), the others neg = ch0, pos = max(ch1..3) (:299-304); torch.max propagates
 * NaN.  permute(0,2,3,1) + view + cat = prior index n = off[l] + y*W + x.  softmax != 0 applies
 * nn.Softmax(dim=-1) (:332): exp(x - max) / sum, exp per this file's convention (fp64, one rounding).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_heads_to_loc_conf(const float *const *loc_maps, const float *const *conf_maps,
                                   const int *f_h, const int *f_w, const int *neg_max, int n_levels, int B,
                                   int softmax, float *loc_out, float *conf_out)
{
    int64_t N = 0;
    for (int l = 0; l < n_levels; ++l) N += (int64_t)f_h[l] * f_w[l];
    int64_t off = 0;
    for (int l = 0; l < n_levels; ++l) {
        const int64_t hw = (int64_t)f_h[l] * f_w[l];
        for (int b = 0; b < B; ++b)
            for (int64_t q = 0; q < hw; ++q) {
                const int64_t n = (int64_t)b * N + off + q;
                if (conf_out) {
                    const float *c = conf_maps[l] + (int64_t)b * 4 * hw + q;
                    const float c0 = c[0], c1 = c[hw], c2 = c[2 * hw], c3 = c[3 * hw];
                    float neg, pos;
                    if (neg_max[l]) { neg = f_max(f_max(c0, c1), c2); pos = c3; }
                    else            { neg = c0; pos = f_max(f_max(c1, c2), c3); }
                    if (softmax) {
                        const float m = f_max(neg, pos);
                        const float e0 = f_exp(neg - m), e1 = f_exp(pos - m);
                        const float sum = e0 + e1;
                        neg = e0 / sum; pos = e1 / sum;
                    }
                    conf_out[2 * n] = neg; conf_out[2 * n + 1] = pos;
                }
                if (loc_out) {
                    const float *c = loc_maps[l] + (int64_t)b * 4 * hw + q;
                    loc_out[4 * n] = c[0]; loc_out[4 * n + 1] = c[hw]; loc_out[4 * n + 2] = c[2 * hw]; loc_out[4 * n + 3] = c[3 * hw];
                }
            }
        off += hw;
    }
}

/* ------------------------------------------------------------------------------------------
 * Sibling NMS implementations.  flags: 1 SUMFIRST, 2 MINIMUM, 4 PLUS1, 8 LE.
 *   FACEBOX/encoderl.py:218-266 nms_np and MTCNN/mtcnn/core/utils.py:62-113 nms (numpy):
 *     areas = (x2-x1)*(y2-y1); order = scores.argsort()[::-1]; per kept i: inter = max(0, xx2-xx1)*max(0, yy2-yy1);
 *     "Union": inter / (areas[i] + areas[rest] - inter); "Minimum": inter / min(areas[i], areas[rest]); keep ovr < thr.
 *   MTCNN/mtcnn/core/nms.py:4-40 torch_nms: the same with "+ 1" on widths/heights/areas and keep ovr <= thr.
 *   FACEBOX/encoderl.py:268-306 DataEncoder.nms (torch): Union, keep ovr <= thr.
 * np.maximum / np.minimum / clamp propagate NaN.  Sort ties (argsort / torch.sort are unstable): defined as
 * ascending by (score, index) read from the end, i.e. higher index first, like orc_nms.
 * Every box enters (no top_k); keep[] gets `count` indices.
 * ------------------------------------------------------------------------------------------ */
ORC_API int64_t orc_nms_variant(const float *boxes, const float *scores, int64_t n, float thr, int flags, int64_t *keep)
{
    memset(keep, 0, sizeof(int64_t) * (size_t)n);
    if (n == 0) return 0;
    const float p1 = (flags & 4) ? 1.0f : 0.0f;
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    orc_si *ord = (orc_si *)malloc(sizeof(orc_si) * (size_t)n);
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float *b = boxes + 4 * i;
        area[i] = (flags & 4) ? ((b[2] - b[0]) + 1.0f) * ((b[3] - b[1]) + 1.0f) : (b[2] - b[0]) * (b[3] - b[1]);
        ord[i].s = scores[i]; ord[i].i = i;
    }
    qsort(ord, (size_t)n, sizeof(orc_si), cmp_si_asc);
    int64_t m = n, count = 0;
    for (int64_t t = 0; t < n; ++t) idx[t] = ord[t].i;
    while (m > 0) {
        const int64_t i = idx[m - 1];
        keep[count++] = i;
        m -= 1;
        const float *bi = boxes + 4 * i;
        int64_t w = 0;
        for (int64_t t = 0; t < m; ++t) {
            const int64_t j = idx[t];
            const float *bj = boxes + 4 * j;
            const float xx1 = f_max(bi[0], bj[0]), yy1 = f_max(bi[1], bj[1]);
            const float xx2 = f_min(bi[2], bj[2]), yy2 = f_min(bi[3], bj[3]);
            float dw = xx2 - xx1, dh = yy2 - yy1;
            if (flags & 4) { dw += p1; dh += p1; }
            const float inter = f_max(0.0f, dw) * f_max(0.0f, dh);
            float den;
            if (flags & 2) den = f_min(area[i], area[j]);
            else if (flags & 1) den = (area[i] + area[j]) - inter;
            else den = (area[j] - inter) + area[i];
            const float ovr = inter / den;
            const int survive = (flags & 8) ? (ovr <= thr) : (ovr < thr);
            if (survive) idx[w++] = j;
        }
        m = w;
    }
    free(idx); free(ord); free(area);
    return count;
}

/* The same in float64: MTCNN's `nms` runs in the dtype of its float64 `dets` (MTCNN/mtcnn/core/utils.py:62-113). */
typedef struct { double s; int64_t i; } orc_di;
static int cmp_di_asc(const void *pa, const void *pb)
{
    const orc_di *a = (const orc_di *)pa, *b = (const orc_di *)pb;
    int an = a->s != a->s, bn = b->s != b->s;
    if (an != bn) return an - bn;
    if (!an) { if (a->s < b->s) return -1; if (a->s > b->s) return 1; }
    return (a->i > b->i) - (a->i < b->i);
}
ORC_API int64_t orc_nms_variant_f64(const double *boxes, const double *scores, int64_t n, double thr, int flags, int64_t *keep)
{
    memset(keep, 0, sizeof(int64_t) * (size_t)n);
    if (n == 0) return 0;
    const double p1 = (flags & 4) ? 1.0 : 0.0;
    double *area = (double *)malloc(sizeof(double) * (size_t)n);
    orc_di *ord = (orc_di *)malloc(sizeof(orc_di) * (size_t)n);
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const double *b = boxes + 4 * i;
        area[i] = (flags & 4) ? ((b[2] - b[0]) + 1.0) * ((b[3] - b[1]) + 1.0) : (b[2] - b[0]) * (b[3] - b[1]);
        ord[i].s = scores[i]; ord[i].i = i;
    }
    qsort(ord, (size_t)n, sizeof(orc_di), cmp_di_asc);
    int64_t m = n, count = 0;
    for (int64_t t = 0; t < n; ++t) idx[t] = ord[t].i;
    while (m > 0) {
        const int64_t i = idx[m - 1];
        keep[count++] = i;
        m -= 1;
        const double *bi = boxes + 4 * i;
        int64_t w = 0;
        for (int64_t t = 0; t < m; ++t) {
            const int64_t j = idx[t];
            const double *bj = boxes + 4 * j;
            const double xx1 = d_max(bi[0], bj[0]), yy1 = d_max(bi[1], bj[1]);
            const double xx2 = d_min(bi[2], bj[2]), yy2 = d_min(bi[3], bj[3]);
            double dw = xx2 - xx1, dh = yy2 - yy1;
            if (flags & 4) { dw += p1; dh += p1; }
            const double inter = d_max(0.0, dw) * d_max(0.0, dh);
            double den;
            if (flags & 2) den = d_min(area[i], area[j]);
            else if (flags & 1) den = (area[i] + area[j]) - inter;
            else den = (area[j] - inter) + area[i];
            const double ovr = inter / den;
            const int survive = (flags & 8) ? (ovr <= thr) : (ovr < thr);
            if (survive) idx[w++] = j;
        }
        m = w;
    }
    free(idx); free(ord); free(area);
    return count;
}

/* FaceBoxes DataEncoder.decode_np, box part (FACEBOX/encoderl.py:318-320) */
ORC_API void orc_facebox_decode(const float *loc, const float *dbox, int64_t n, float v0, float v1, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        const float *l = loc + 4 * i, *d = dbox + 4 * i;
        const float cx = (l[0] * v0) * d[2] + d[0], cy = (l[1] * v0) * d[3] + d[1];
        const float w = f_exp(l[2] * v1) * d[2], h = f_exp(l[3] * v1) * d[3];
        out[4 * i] = cx - w / 2.0f; out[4 * i + 1] = cy - h / 2.0f; out[4 * i + 2] = cx + w / 2.0f; out[4 * i + 3] = cy + h / 2.0f;
    }
}
