// Library-level entry points: version, thread-local error string, device check.
#include <cstdarg>
#include <cstdio>
#include "fdt_common.cuh"

static thread_local char g_err[512] = "";

void fdt_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

FDT_API int fdt_version(void) { return 210; }
FDT_API const char *fdt_last_error(void) { return g_err; }

FDT_API int fdt_device_check(int device)
{
    int n = 0;
    FDT_CUDA(cudaGetDeviceCount(&n));
    FDT_REQUIRE(device >= 0 && device < n, FDT_E_DEVICE, "device %d not present (%d CUDA devices)", device, n);
    cudaDeviceProp p;
    FDT_CUDA(cudaGetDeviceProperties(&p, device));
    FDT_REQUIRE(p.major == 10 && p.minor == 0, FDT_E_DEVICE,
                "device %d is sm_%d%d (%s); libfdt_b200 is built for sm_100a (B200) only", device, p.major, p.minor, p.name);
    return FDT_OK;
}
