// Micro-benchmark: throughput of shared-memory reductions (red.shared.add.u32 -> SASS ATOMS.POPC.INC) per SM,
// against LDS / STS of the same address patterns.  nvcc -arch=sm_100a -O3 -o atoms_bench atoms_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: red.shared, 1: lds+sts (non-atomic rmw), 2: sts only, 3: atom.shared with return
__global__ void k(int iters, int pattern, unsigned long long* cycles, unsigned* sink) {
    __shared__ unsigned hist[64 * 128];
    for (int i = threadIdx.x; i < 64 * 128; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t rng = threadIdx.x * 2654435761u + 12345u;
    const uint32_t base = smem_u32(hist);
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        rng = rng * 1664525u + 1013904223u;
        uint32_t idx;
        if (pattern == 0) idx = lane + 32 * (warp & 7);                       // conflict-free, distinct addresses
        else if (pattern == 1) idx = (lane * 128 + ((rng >> 10) & 127));      // own row per lane, random bin (S3-like)
        else if (pattern == 2) idx = ((lane % 22) * 128 + 60 + ((rng >> 10) % 12)); // 22 rows shared by lanes, clustered bins (membrane-like)
        else if (pattern == 3) idx = 7;                                       // one address
        else idx = (lane * 128 + 60 + ((rng >> 10) % 12));                    // own row, clustered bins
        const uint32_t addr = base + 4u * idx;
        if (MODE == 0) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
        else if (MODE == 1) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v + 1) : "memory"); }
        else if (MODE == 2) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(i) : "memory");
        else { unsigned v; asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(addr) : "memory"); acc += v; }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0xffffffffu) sink[0] = acc + hist[threadIdx.x];
}

int main() {
    unsigned long long* cyc; unsigned* sink;
    cudaMalloc(&cyc, 8 * 1024); cudaMalloc(&sink, 64);
    const int iters = 4096;
    const char* pn[] = {"conflict-free", "own row, random bin", "22 shared rows, clustered", "one address", "own row, clustered"};
    const char* mn[] = {"red.shared", "lds+sts", "sts", "atom.shared(ret)"};
    for (int warps = 4; warps <= 16; warps *= 2)
        for (int m = 0; m < 4; ++m)
            for (int p = 0; p < 5; ++p) {
                void (*fn)(int, int, unsigned long long*, unsigned*) = m == 0 ? k<0> : m == 1 ? k<1> : m == 2 ? k<2> : k<3>;
                fn<<<148, warps * 32>>>(iters, p, cyc, sink);
                cudaDeviceSynchronize();
                fn<<<148, warps * 32>>>(iters, p, cyc, sink);
                cudaDeviceSynchronize();
                unsigned long long h[148];
                cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
                double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
                printf("warps/SM %2d  %-18s %-28s  %.2f cycles per warp-instruction (SM-wide)\n", warps, mn[m], pn[p], avg / (double)(iters * warps));
            }
    return 0;
}
