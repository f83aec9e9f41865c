// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM (TMEM <-> registers), the unknown behind every tcgen05 softmax
// design: if reading fp32 scores out of TMEM is capped near 16 values / clk / SM it is as tight a bound as the MUFU pipe.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tmem_ld tools/microbench/tmem_ld.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(r, addr) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
    : "r"(addr))
#define ST32(addr, r) asm volatile( \
    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " \
    "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, " \
    "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" \
    :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
       "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), \
       "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), \
       "r"(r[30]), "r"(r[31]) : "memory")

// MODE 0: one x32 load per wait; 1: two x32 loads in flight per wait; 2: four; 3: x32 store + wait::st; 4: load + 32 FADD (consume)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cycles, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[32], q[32], acc = 0;
    for (int i = 0; i < 32; ++i) { r[i] = threadIdx.x + i; q[i] = i; }
    // initialise the columns this warp will read
    for (int c = 0; c < 512; c += 32) { ST32(base + c, r); }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    uint32_t col = (uint32_t)((warp >> 2) * 128) & 511u;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            LD32(r, base + col);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= r[0] ^ r[31];
            col = (col + 32) & 511u;
        } else if (MODE == 1) {
            LD32(r, base + col);
            LD32(q, base + ((col + 32) & 511u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= r[0] ^ q[31];
            col = (col + 64) & 511u;
        } else if (MODE == 2) {
            LD32(r, base + col);
            LD32(q, base + ((col + 32) & 511u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= r[0] ^ q[31];
            LD32(r, base + ((col + 64) & 511u));
            LD32(q, base + ((col + 96) & 511u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= r[1] ^ q[30];
            col = (col + 128) & 511u;
        } else if (MODE == 3) {
            ST32(base + col, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            col = (col + 32) & 511u;
        } else {
            LD32(r, base + col);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float f = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) f += __uint_as_float(r[i]);
            acc ^= __float_as_uint(f);
            col = (col + 32) & 511u;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) out[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
    }
}

template <int MODE> void run(const char* name, int warps, double bytes_per_iter_warp) {
    uint32_t* d; long long* c;
    cudaMalloc(&d, 4); cudaMalloc(&c, 148 * sizeof(long long));
    const int iters = 20000;
    k<MODE><<<148, warps * 32>>>(d, c, 100);
    k<MODE><<<148, warps * 32>>>(d, c, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < 148; ++i) mean += (double)h[i]; mean /= 148;
    printf("%-44s warps %2d: %8.1f cycles/iter/warp, %7.1f B/clk/SM (%5.1f fp32 values/clk/SM)\n", name, warps, mean / iters,
           bytes_per_iter_warp * warps * iters / mean, bytes_per_iter_warp * warps * iters / mean / 4);
    cudaFree(d); cudaFree(c);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("tcgen05.ld 32x32b.x32, 1 per wait", w, 32 * 32 * 4);
        run<1>("tcgen05.ld 32x32b.x32, 2 per wait", w, 2 * 32 * 32 * 4);
        run<2>("tcgen05.ld 32x32b.x32, 2 per wait, x2", w, 4 * 32 * 32 * 4);
        run<4>("tcgen05.ld x32 + 32 FADD", w, 32 * 32 * 4);
        run<3>("tcgen05.st 32x32b.x32, 1 per wait", w, 32 * 32 * 4);
    }
    return 0;
}
