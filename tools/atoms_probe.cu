// Microbenchmark: throughput of red.shared.xor (32- and 64-bit) against LDS/STS on one SM and on the whole chip.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/atoms_probe.cu -o tools/_abl/atoms_probe
// Prints cycles per warp instruction and per active lane for a block of 1024 threads (32 warps), addresses spread
// pseudo-randomly over 64 KB of shared memory (the event kernel's pattern), for several active-lane counts.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(uint32_t *out, long long *cyc, int active, int iters) {
    extern __shared__ __align__(16) uint32_t sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t x = threadIdx.x * 2654435761u + 12345u;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    const bool on = (int)lane < active;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x = x * 1664525u + 1013904223u;
            const uint32_t w = (x >> 10) & 16383u;
            const uint32_t m = 1u << (x & 31u);
            if (MODE == 0) {
                asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; @p red.shared.xor.b32 [%0], %1;}" ::"r"(base + w * 4), "r"(m), "r"((uint32_t)on) : "memory");
            } else if (MODE == 1) {
                unsigned long long m64 = (unsigned long long)m << (x >> 31 ? 32 : 0);
                asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; @p red.shared.xor.b64 [%0], %1;}" ::"r"(base + (w & ~1u) * 4), "l"(m64), "r"((uint32_t)on) : "memory");
            } else if (MODE == 2) {
                asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.b32 [%0], %1;}" ::"r"(base + w * 4), "r"(m), "r"((uint32_t)on) : "memory");
            } else if (MODE == 3) {
                uint32_t v;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(base + w * 4) : "memory");
                if (on) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + w * 4), "r"(v ^ m) : "memory");
            } else if (MODE == 4) {
                uint32_t v;
                asm volatile("{.reg .pred p; setp.ne.u32 p, %3, 0; @p atom.shared.xor.b32 %0, [%1], %2;}" : "=r"(v) : "r"(base + w * 4), "r"(m), "r"((uint32_t)on) : "memory");
                x ^= v & 0;
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) acc ^= sm[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + x;
}

template <int MODE>
void run(const char *name, int threads) {
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int active : {32, 16, 8, 4, 1}) {
        const int iters = 2000;
        probe<MODE><<<148, threads, 65536>>>(out, cyc, active, iters);
        probe<MODE><<<148, threads, 65536>>>(out, cyc, active, iters);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < 148; i++) c += (double)h[i];
        c /= 148;
        const double winst = (double)iters * 8 * (threads / 32);
        printf("%-28s threads %4d active %2d: %7.2f cyc/warp-inst/SM  %6.3f cyc/active-lane\n", name, threads, active, c / winst,
               c / winst / active);
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int threads : {1024, 256}) {
        run<0>("red.shared.xor.b32", threads);
        run<1>("red.shared.xor.b64", threads);
        run<4>("atom.shared.xor.b32", threads);
        run<2>("st.shared.b32", threads);
        run<3>("ld+xor+st (non-atomic)", threads);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return 0;
}
