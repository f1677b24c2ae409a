// Microbenchmark: TMEM -> register bandwidth of tcgen05.ld by shape / repeat count / number of warps (sm_100a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu ; run: ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD_ASM_32(shape, num, ...)                                                                                       \
    asm volatile("tcgen05.ld.sync.aligned." shape "." num ".b32 "                                                         \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),   \
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),  \
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                     \
                 : "r"(addr))

template <int MODE>
__global__ void __launch_bounds__(512) bench(unsigned long long* cycles, uint32_t* sink, int iters, int waits_every) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[32];
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t addr = base + (uint32_t)((it * 32) & 255) + (uint32_t)((warp >> 2) & 1) * 256;
        if (MODE == 0) LD_ASM_32("32x32b", "x32");          // 32 lanes x 32 cols  = 4 KB / warp
        if (MODE == 1) LD_ASM_32("16x256b", "x8");          // 16 lanes x 64 cols  = 4 KB / warp
        if (MODE == 2) LD_ASM_32("16x128b", "x16");         // 16 lanes x 64 cols  = 4 KB / warp
        if (MODE == 3) LD_ASM_32("16x64b", "x32");          // 16 lanes x 64 cols  = 4 KB / warp
        if ((it % waits_every) == waits_every - 1) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= v[i];
        }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= v[i];
    const unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

int main() {
    unsigned long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    const int iters = 4096;
    const char* names[4] = {"32x32b.x32", "16x256b.x8", "16x128b.x16", "16x64b.x32"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {4, 8, 16})
            for (int we : {1, 4}) {
                for (int rep = 0; rep < 2; ++rep) {
                    if (mode == 0) bench<0><<<148, warps * 32>>>(cyc, sink, iters, we);
                    if (mode == 1) bench<1><<<148, warps * 32>>>(cyc, sink, iters, we);
                    if (mode == 2) bench<2><<<148, warps * 32>>>(cyc, sink, iters, we);
                    if (mode == 3) bench<3><<<148, warps * 32>>>(cyc, sink, iters, we);
                }
                cudaError_t e = cudaDeviceSynchronize();
                unsigned long long h[148];
                cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                double c = (double)h[0];
                double bytes = (double)iters * warps * 4096.0;
                printf("%-12s warps=%2d wait_every=%d  cycles=%9.0f  B/cycle/SM=%7.1f  cycles/ld/warp=%6.1f  %s\n", names[mode], warps, we, c,
                       bytes / c, c / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
