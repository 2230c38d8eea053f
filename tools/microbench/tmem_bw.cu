// Microbenchmark (developer tool, not part of the product): tensor-memory read bandwidth per SM on B200.
// W warps of one CTA per SM read their TMEM lane quarter with tcgen05.ld in a loop; cycles are taken with clock64().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink);

template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t& sink) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
}
// packed 16-bit read: 32 registers carry 64 columns' low halves
__device__ __forceinline__ void ld_pack(uint32_t taddr, uint32_t& sink) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
}

// mode 0: 32x32b.x32 (32 columns / instruction), mode 1: x32.pack::16b (64 columns / instruction, 32 registers)
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, int mode, long long* cycles, uint32_t* out) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const uint32_t col = (uint32_t)((i * 64 + (warp >> 2) * 128) & 511);
        if (mode == 0) {
            ld<32>(base + col, sink);
            ld<32>(base + ((col + 32) & 511), sink);
        } else {
            ld_pack(base + (col & 448), sink);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s) : "memory");
}

int main() {
    long long* d_cycles;
    uint32_t* d_out;
    cudaMalloc(&d_cycles, 148 * 8);
    cudaMalloc(&d_out, 148 * 512 * 4);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int warps : {4, 8, 16}) {
            tmem_read_kernel<<<148, warps * 32>>>(iters, mode, d_cycles, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e));
                return 1;
            }
            long long h[148];
            cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; ++i) avg += (double)h[i];
            avg /= 148;
            // 32-bit TMEM cells read per iteration per warp: 32 lanes x 64 columns
            const double cells = (double)iters * warps * 32 * 64;
            printf("mode %s warps %2d: %.0f cycles, %.1f TMEM B/clk/SM (32-bit cells x4), %.1f register B/clk/SM\n",
                   mode == 0 ? "32x32b.x32      " : "x32.pack::16b   ", warps, avg, cells * 4 / avg, cells * (mode == 0 ? 4 : 2) / avg);
        }
    }
    return 0;
}
