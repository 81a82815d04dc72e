// Probe: how far can a chain of programmatically-serialised kernels on ONE stream run ahead of a long-running kernel?
// A (64 CTAs x 1024 thr, 200 KB smem, spins `spin_us`, triggers at its top) -> B (1 block) -> C (1088 blocks x 256) -> A' -> B' -> C' ...
// none of them executes griddepcontrol.wait.  Prints globaltimer start/end of every grid relative to the first.
// nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_probe tools/pdl_probe.cu && ./pdl_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void kA(unsigned long long *rec, int id, long long spin_ns, int wait)
{
    extern __shared__ char sm[];
    if (wait) cudaGridDependencySynchronize();
    unsigned long long t0 = gtime();
    cudaTriggerProgrammaticLaunchCompletion();
    if (threadIdx.x == 0) { atomicMin(&rec[2 * id], t0); }
    while ((long long)(gtime() - t0) < spin_ns) { }
    if (threadIdx.x == 0) atomicMax(&rec[2 * id + 1], gtime());
    sm[threadIdx.x] = 0;
}
__global__ void kS(unsigned long long *rec, int id, long long spin_ns, int wait)
{
    if (wait) cudaGridDependencySynchronize();
    unsigned long long t0 = gtime();
    cudaTriggerProgrammaticLaunchCompletion();
    if (threadIdx.x == 0) atomicMin(&rec[2 * id], t0);
    while ((long long)(gtime() - t0) < spin_ns) { }
    if (threadIdx.x == 0) atomicMax(&rec[2 * id + 1], gtime());
}
template <typename K> void launch(K k, dim3 g, dim3 b, size_t sm, cudaStream_t st, bool pdl, unsigned long long *rec, int id, long long spin, int wait)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k, rec, id, spin, wait);
}
int main(int argc, char **argv)
{
    const int calls = 6;
    unsigned long long *rec, h[64];
    cudaMalloc(&rec, 64 * 8);
    cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaStream_t st; cudaStreamCreate(&st);
    for (int variant = 0; variant < 3; ++variant) {
        // variant 0: nobody waits; 1: C waits on B (hardware wait) and A' waits on C; 2: no PDL at all (plain stream order)
        for (int rep = 0; rep < 2; ++rep) {
            for (int i = 0; i < 64; ++i) h[i] = (i & 1) ? 0ull : ~0ull;
            cudaMemcpy(rec, h, sizeof(h), cudaMemcpyHostToDevice);
            kS<<<1, 32, 0, st>>>(rec, 31, 300000, 0);      // 0.3 ms spin: the host enqueues everything behind it
            for (int c = 0; c < calls; ++c) {
                const bool pdl = variant != 2;
                launch(kS, dim3(1), dim3(256), 0, st, pdl, rec, 3 * c + 0, 500, 0);
                launch(kS, dim3(1088), dim3(256), 0, st, pdl, rec, 3 * c + 1, 3000, variant == 1);
                launch(kA, dim3(64), dim3(1024), 200 * 1024, st, pdl, rec, 3 * c + 2, 40000, variant == 1);
            }
            cudaStreamSynchronize(st);
        }
        cudaMemcpy(h, rec, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long t0 = h[0];
        printf("variant %d (%s)\n", variant, variant == 0 ? "PDL, nobody waits" : variant == 1 ? "PDL, K2 and K3 wait on their predecessor" : "no PDL");
        for (int c = 0; c < calls; ++c)
            printf("  call %d: begin %7.1f..%7.1f  K2 %7.1f..%7.1f  K3 %7.1f..%7.1f us\n", c,
                   (h[6 * c] - t0) / 1e3, (h[6 * c + 1] - t0) / 1e3, (h[6 * c + 2] - t0) / 1e3, (h[6 * c + 3] - t0) / 1e3,
                   (h[6 * c + 4] - t0) / 1e3, (h[6 * c + 5] - t0) / 1e3);
        printf("  total %.1f us for %d calls\n", (h[6 * (calls - 1) + 5] - t0) / 1e3, calls);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
