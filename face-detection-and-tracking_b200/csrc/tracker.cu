// placeholder until the tracker kernels land (next commit): keeps the ABI complete and loud.
#include "fdt_common.cuh"
FDT_API size_t fdt_iou_track_workspace_bytes(int64_t, int64_t, int64_t) { return 256; }
FDT_API int fdt_iou_track(const double *, const int64_t *, int64_t, int64_t, int64_t, double, double, int64_t, int64_t *, int64_t *, int64_t *, int64_t *, double *, void *, size_t, fdt_stream_t) { fdt_set_error("fdt_iou_track: not built yet"); return FDT_E_UNSUPPORTED; }
