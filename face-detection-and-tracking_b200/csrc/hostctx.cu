// Host-buffer entry points: the reference-facing call when the caller's tensors live in host memory
// (the reference's Detect builds its output with torch.zeros on the default device, detection.py:48).
// The context owns one stream and grow-only device buffers; copies are cudaMemcpyAsync on that stream
// (truly asynchronous when the host buffers are pinned) and the call returns after the results landed.
//
// Detect only decodes the candidates NMS actually looks at (<= 1024 rows of `loc` per image and round), so when
// `loc_h` is pinned (page-locked, hence mapped into the unified address space) it is NOT copied: k_sort_nms gathers
// those 16-byte rows straight from host memory over PCIe -- ~1 MB instead of 35 MB per batch of 64.  `conf` is streamed
// in full by K2, so it is copied (17.5 MB) and read from HBM.  A pinned `out_h` is likewise written in place: the output
// stage of k_sort_nms stores the rows over PCIe (posted writes) while other images are still in NMS, and the separate
// device-to-host copy of the [B,C,top_k,5] block disappears.
#include <cstdlib>
#include "fdt_common.cuh"

struct fdt_ctx {
    int device;
    cudaStream_t stream;
    void *buf;
    size_t cap;
};

static int ctx_reserve(fdt_ctx *c, size_t bytes)
{
    if (bytes <= c->cap) return FDT_OK;
    if (c->buf) { FDT_CUDA(cudaFree(c->buf)); c->buf = nullptr; c->cap = 0; }
    size_t want = bytes + bytes / 4;
    FDT_CUDA(cudaMalloc(&c->buf, want));
    c->cap = want;
    return FDT_OK;
}

FDT_API int fdt_ctx_create(int device, fdt_ctx **out)
{
    FDT_REQUIRE(out != nullptr, FDT_E_INVALID, "fdt_ctx_create: out is null");
    int rc = fdt_device_check(device);
    if (rc != FDT_OK) return rc;
    FDT_CUDA(cudaSetDevice(device));
    fdt_ctx *c = (fdt_ctx *)calloc(1, sizeof(fdt_ctx));
    FDT_REQUIRE(c != nullptr, FDT_E_INVALID, "fdt_ctx_create: out of host memory");
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { free(c); fdt_set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return FDT_E_CUDA; }
    *out = c;
    return FDT_OK;
}

FDT_API int fdt_ctx_destroy(fdt_ctx *c)
{
    if (!c) return FDT_OK;
    cudaSetDevice(c->device);
    if (c->buf) cudaFree(c->buf);
    cudaStreamDestroy(c->stream);
    free(c);
    return FDT_OK;
}

FDT_API int fdt_detect_host(fdt_ctx *c, const float *loc_h, const float *conf_h, const float *priors_h,
                            int B, int64_t N, int C, int top_k, int nms_top_k,
                            float conf_thresh, float nms_thresh, float var0, float var1,
                            float *out_h, int32_t *counts_h, int64_t *kept_prior_h)
{
    FDT_REQUIRE(c != nullptr, FDT_E_INVALID, "fdt_detect_host: null context");
    FDT_REQUIRE(B >= 0 && N >= 0 && C >= 1 && top_k >= 1, FDT_E_INVALID, "fdt_detect_host: bad sizes");
    if (B == 0) return FDT_OK;
    FDT_REQUIRE(loc_h && conf_h && priors_h && out_h, FDT_E_INVALID, "fdt_detect_host: null pointer argument");
    FDT_CUDA(cudaSetDevice(c->device));
    // pinned + 16-byte aligned loc: gather it in place (zero-copy), see the header comment
    const float *loc_dev_view = nullptr;
    {
        cudaPointerAttributes attr;
        cudaError_t e = cudaPointerGetAttributes(&attr, loc_h);
        if (e == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer && fdt_aligned(attr.devicePointer, 16))
            loc_dev_view = (const float *)attr.devicePointer;
        else if (e != cudaSuccess) (void)cudaGetLastError();          // pageable memory on old drivers reports an error: ignore
    }
    float *out_dev_view = nullptr;
    {
        cudaPointerAttributes attr;
        cudaError_t e = cudaPointerGetAttributes(&attr, out_h);
        if (e == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer && fdt_aligned(attr.devicePointer, 16))
            out_dev_view = (float *)attr.devicePointer;
        else if (e != cudaSuccess) (void)cudaGetLastError();
    }
    const size_t sz_loc = loc_dev_view ? 0 : fdt_align256((size_t)B * N * 16), sz_conf = fdt_align256((size_t)B * N * C * 4);
    const size_t sz_pri = fdt_align256((size_t)N * 16), sz_out = fdt_align256((size_t)B * C * top_k * 20);
    const size_t sz_cnt = fdt_align256((size_t)B * C * 4), sz_kept = fdt_align256((size_t)B * C * top_k * 8);
    const size_t sz_ws = fdt_detect_workspace_bytes(B, N, C);
    int rc = ctx_reserve(c, sz_loc + sz_conf + sz_pri + sz_out + sz_cnt + sz_kept + sz_ws);
    if (rc != FDT_OK) return rc;
    char *p = (char *)c->buf;
    const float *d_loc = loc_dev_view ? loc_dev_view : (const float *)p; p += sz_loc;
    float *d_conf = (float *)p; p += sz_conf;
    float *d_pri = (float *)p; p += sz_pri;
    float *d_out = out_dev_view ? out_dev_view : (float *)p; p += sz_out;
    int32_t *d_cnt = (int32_t *)p; p += sz_cnt;
    int64_t *d_kept = (int64_t *)p; p += sz_kept;
    void *d_ws = p;
    cudaStream_t st = c->stream;
    FDT_CUDA(cudaMemcpyAsync(d_conf, conf_h, (size_t)B * N * C * 4, cudaMemcpyHostToDevice, st));
    FDT_CUDA(cudaMemcpyAsync(d_pri, priors_h, (size_t)N * 16, cudaMemcpyHostToDevice, st));
    if (!loc_dev_view) FDT_CUDA(cudaMemcpyAsync((void *)d_loc, loc_h, (size_t)B * N * 16, cudaMemcpyHostToDevice, st));
    rc = fdt_detect_flags(d_loc, d_conf, d_pri, B, N, C, top_k, nms_top_k, conf_thresh, nms_thresh, var0, var1,
                          d_out, counts_h ? d_cnt : nullptr, kept_prior_h ? d_kept : nullptr, d_ws, sz_ws, st,
                          loc_dev_view ? FDT_FLAG_LOC_HOST_MAPPED : 0u);
    if (rc != FDT_OK) return rc;
    if (!out_dev_view) FDT_CUDA(cudaMemcpyAsync(out_h, d_out, (size_t)B * C * top_k * 20, cudaMemcpyDeviceToHost, st));
    if (counts_h) FDT_CUDA(cudaMemcpyAsync(counts_h, d_cnt, (size_t)B * C * 4, cudaMemcpyDeviceToHost, st));
    if (kept_prior_h) FDT_CUDA(cudaMemcpyAsync(kept_prior_h, d_kept, (size_t)B * C * top_k * 8, cudaMemcpyDeviceToHost, st));
    FDT_CUDA(cudaStreamSynchronize(st));
    return FDT_OK;
}
