// placeholder until the MultiBoxLoss kernels land (next commit): keeps the ABI complete and loud.
#include "fdt_common.cuh"
FDT_API size_t fdt_match_workspace_bytes(int, int64_t, int64_t) { return 256; }
FDT_API int fdt_match_encode(const float *, const float *, const int64_t *, int, int64_t, float, float, float, int, float *, int64_t *, int32_t *, float *, void *, size_t, fdt_stream_t) { fdt_set_error("fdt_match_encode: not built yet"); return FDT_E_UNSUPPORTED; }
FDT_API size_t fdt_mine_workspace_bytes(int, int64_t) { return 256; }
FDT_API int fdt_hard_negative_mine(const float *, const uint8_t *, int, int64_t, int, uint8_t *, void *, size_t, fdt_stream_t) { fdt_set_error("fdt_hard_negative_mine: not built yet"); return FDT_E_UNSUPPORTED; }
FDT_API size_t fdt_multibox_workspace_bytes(int, int64_t, int, int64_t) { return 256; }
FDT_API int fdt_multibox_loss_forward(const float *, const float *, const float *, const float *, const int64_t *, int, int64_t, int, float, int, int, float, float, float *, float *, float *, int64_t *, uint8_t *, float *, void *, size_t, fdt_stream_t) { fdt_set_error("fdt_multibox_loss_forward: not built yet"); return FDT_E_UNSUPPORTED; }
FDT_API int fdt_multibox_loss_backward(const float *, const float *, const float *, const int64_t *, const uint8_t *, const float *, float, float, int, int64_t, int, float *, float *, fdt_stream_t) { fdt_set_error("fdt_multibox_loss_backward: not built yet"); return FDT_E_UNSUPPORTED; }
