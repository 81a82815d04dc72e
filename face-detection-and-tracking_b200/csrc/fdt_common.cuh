// Shared helpers for libfdt_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "fdt_b200.h"

#define FDT_API extern "C" __attribute__((visibility("default")))

void fdt_set_error(const char *fmt, ...);

// library-internal form of fdt_detect (hostctx.cu): `loc` may be a device view of pinned HOST memory that k_sort_nms gathers in place
constexpr unsigned FDT_FLAG_LOC_HOST_MAPPED = 1u;
int fdt_detect_flags(const float *loc, const float *conf, const float *priors, int B, int64_t N, int C, int top_k, int nms_top_k,
                     float conf_thresh, float nms_thresh, float var0, float var1, float *out, int32_t *counts, int64_t *kept_prior,
                     void *ws, size_t ws_bytes, void *stream, unsigned flags);

constexpr int FDT_DETECT_MAX_DEPTH = 4;     // slots of the stateful Detect workspace (calls of one stream in flight)
int fdt_option_host_chunk();     // images per copy/compute chunk of the host-buffer pipeline (hostctx.cu); option "host_chunk"

// greedy NMS for any n and for float64 (nms_generic.cu): sort + pairwise bit mask + serial reduce
constexpr int FDT_MAX_NMS_GENERIC = 131072;
size_t fdt_nms_generic_workspace_bytes(int64_t n);
template <typename T>
int fdt_nms_generic(const T *boxes, const T *scores, int64_t n, T thresh, int64_t top_k, int variant,
                    int64_t *keep, int64_t *count, void *ws, size_t ws_bytes, cudaStream_t st);

#define FDT_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fdt_set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return FDT_E_CUDA;                                                                    \
        }                                                                                         \
    } while (0)

#define FDT_REQUIRE(cond, code, ...)                                                              \
    do {                                                                                          \
        if (!(cond)) { fdt_set_error(__VA_ARGS__); return (code); }                               \
    } while (0)

#define FDT_LAUNCH_CHECK() FDT_CUDA(cudaGetLastError())

static inline size_t fdt_align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline bool fdt_aligned(const void *p, size_t a) { return ((uintptr_t)p & (a - 1)) == 0; }

constexpr int FDT_NUM_SMS = 148;            // B200
constexpr int FDT_SMEM_MAX = 232448;        // 227 KB opt-in dynamic shared memory per CTA

// ---- device helpers -------------------------------------------------------------------------
// Order-preserving float -> uint32 map (larger float -> larger uint; +NaN above +inf like torch.sort).
__device__ __forceinline__ uint32_t fdt_float_key(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fdt_key_float(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// exp/log evaluated in fp64, rounded once to fp32 (see include/fdt_b200.h conventions)
__device__ __forceinline__ float fdt_expf_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float fdt_logf_cr(float x) { return (float)log((double)x); }

// decode (layers/box_utils.py:238-258); compiled with -fmad=false so no FMA is formed
__device__ __forceinline__ float4 fdt_decode1(float4 l, float4 p, float v0, float v1)
{
    float cx = p.x + (l.x * v0) * p.z;
    float cy = p.y + (l.y * v0) * p.w;
    float w = p.z * fdt_expf_cr(l.z * v1);
    float h = p.w * fdt_expf_cr(l.w * v1);
    float x1 = cx - w / 2.0f;
    float y1 = cy - h / 2.0f;
    return make_float4(x1, y1, w + x1, h + y1);
}
// encode (layers/box_utils.py:213-234)
__device__ __forceinline__ float4 fdt_encode1(float4 m, float4 p, float v0, float v1)
{
    float4 o;
    o.x = ((m.x + m.z) / 2.0f - p.x) / (v0 * p.z);
    o.y = ((m.y + m.w) / 2.0f - p.y) / (v0 * p.w);
    o.z = fdt_logf_cr((m.z - m.x) / p.z) / v1;
    o.w = fdt_logf_cr((m.w - m.y) / p.w) / v1;
    return o;
}
