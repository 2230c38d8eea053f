// proqa_b200 — common device helpers (sm_100a only).
//
// PTX wrappers for mbarrier / TMA / tcgen05 plus the 64-bit candidate key that every
// selection kernel in this engine sorts on.  Nothing in here is portable to other
// architectures on purpose: the engine is written for B200 (sm_100a) alone.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <float.h>

namespace pq {

constexpr int kDim = 128;  // ProQA embedding width (reference: retrieval/eval_retrieval.py:98, retriever.py:16-20)

// ---------------------------------------------------------------------------------------------
// Candidate key: (score, row) packed so that a plain unsigned 64-bit "greater" is exactly the
// engine's total order  "score descending, then row id ascending"  (FAISS keeps the lowest ids on
// a k-th place tie because its heap replaces on strict '>' while scanning ids upwards).
//   hi 32 bits: monotone map of the fp32 score     lo 32 bits: ~row  (smaller row => larger key)
// Key 0 is reserved as "empty slot" (it is below every real score, see kThrFloor).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
    return (uint64_t(f32_to_ordered(score)) << 32) | uint64_t(~row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) { return ordered_to_f32(uint32_t(key >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) { return ~uint32_t(key); }

// Smallest score that may enter a result list: FAISS initialises its heaps with -FLT_MAX and
// replaces on strict '>', so -FLT_MAX itself, -inf and NaN never enter.  A score passes iff
// score >= kThrFloor where kThrFloor is the fp32 successor of -FLT_MAX.
#define PQ_THR_FLOOR (-3.4028232635611926e38f)

// ---------------------------------------------------------------------------------------------
// The engine's DEFINED fp32 score (restated bit for bit in oracle/flat_oracle.c: engine_dot):
//   p_j = fmaf chain over dims 16j .. 16j+15 (ascending, starting from 0)            j = 0..7
//   <a,b> := ((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7))
// Eight independent chains instead of one 128-long one: the scan kernel gives each chain to one lane of an
// 8-lane group and folds them with three butterfly shuffles; any other kernel (rescoring, norms) computes the
// same bits with eight accumulators in one thread.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ float engine_dot(const float* __restrict__ a, const float* __restrict__ b) {
    float p[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 0; j < 8; ++j) {
        float acc = 0.f;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < 16; ++i) acc = fmaf(a[16 * j + i], b[16 * j + i], acc);
        p[j] = acc;
    }
    return ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// shared-memory address / mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---------------------------------------------------------------------------------------------
// TMA: 2-D tiled tensor-map load, global -> shared, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, int32_t x, int32_t y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// Same, with an L2 eviction-priority hint (createpolicy value).
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, int32_t x, int32_t y, uint64_t* bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(x), "r"(y), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 (5th-gen tensor cores, TMEM accumulators)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // whole warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp, same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 TMEM lanes (one per thread of the warp) x 32 consecutive fp32 columns -> 32 registers.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 B
// (64 bf16), 8-row core groups 1024 B apart.  Field layout as in the PTX ISA "matrix
// descriptor" table for tcgen05 (start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout=SWIZZLE_128B(2) [61,64)).
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
    d |= uint64_t(1) << 16;            // LBO (ignored for swizzled K-major)
    d |= uint64_t(1024 >> 4) << 32;    // SBO: 8 rows * 128 B
    d |= uint64_t(1) << 46;            // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;            // SWIZZLE_128B
    return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4)            // D format  = F32
         | (1u << 7)            // A format  = BF16
         | (1u << 10)           // B format  = BF16
         | ((N >> 3) << 17)     // N / 8
         | ((M >> 4) << 24);    // M / 16
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// 16-byte read-only load the compiler must issue where it stands (a plain __ldg may be sunk next to its use, which serialises
// the round trips of a latency-bound gather).
__device__ __forceinline__ float4 ldg_f4_now(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// 8-byte asynchronous copy global -> shared (LDGSTS): no register staging, any number in flight per thread.
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// engine_dot computed by one warp: lane l holds dims 4l..4l+3 of both vectors.  Lanes 4j..4j+3 pass chain p_j from lane
// to lane, then the eight p_j are tree-combined with three butterfly shuffles — the same bits as engine_dot above.
// Every lane returns the result.
__device__ __forceinline__ float warp_engine_dot(const float4 a, const float4 b, int lane) {
    float acc = 0.f;
#pragma unroll
    for (int tstep = 0; tstep < 4; ++tstep) {
        const float in = __shfl_up_sync(0xffffffffu, acc, 1);
        if ((lane & 3) == tstep) {
            float x = tstep == 0 ? 0.f : in;
            x = fmaf(a.x, b.x, x);
            x = fmaf(a.y, b.y, x);
            x = fmaf(a.z, b.z, x);
            x = fmaf(a.w, b.w, x);
            acc = x;
        }
    }
    acc = __shfl_sync(0xffffffffu, acc, lane | 3);  // p_j to all four lanes of group j
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    return acc;
}

// engine_dot of FOUR rows by one warp: lane = 8 r + j computes chain p_j (dims 16j .. 16j+15, ascending, from 0) of row r from its
// own 64 bytes of the row and the query in shared memory; three butterfly steps inside each group of eight lanes add the chains
// in the tree of engine_dot (xor 1: p0+p1 ..., xor 2, xor 4; fp addition commutes, so every lane of the group ends with the same
// bits).  35 instructions for four rows, against 4 x 35 with warp_engine_dot.  `row` may be null (lane group without a row).
__device__ __forceinline__ float quad_engine_dot(const float* __restrict__ row, const float* q_smem, int lane) {
    const int j = lane & 7;
    float a = 0.f;
    if (row != nullptr) {
        const float4* r4 = reinterpret_cast<const float4*>(row + 16 * j);
        const float4* q4 = reinterpret_cast<const float4*>(q_smem + 16 * j);
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldg(r4 + i);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 qq = q4[i];
            a = fmaf(v[i].x, qq.x, a);
            a = fmaf(v[i].y, qq.y, a);
            a = fmaf(v[i].z, qq.z, a);
            a = fmaf(v[i].w, qq.w, a);
        }
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    return a;
}

// Bitonic sort of n (power of two) keys in shared memory by the whole CTA; ascending or descending.
template <int kThreads>
__device__ __forceinline__ void block_sort(uint64_t* a, int n, bool ascending) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (n >> 1); i += kThreads) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0) != ascending;
                const uint64_t x = a[lo], y = a[hi];
                if ((x < y) == desc) {
                    a[lo] = y;
                    a[hi] = x;
                }
            }
            __syncthreads();
        }
    }
}
template <int kThreads>
__device__ __forceinline__ void block_sort_desc(uint64_t* a, int n) {
    block_sort<kThreads>(a, n, false);
}
// a[0..n) is bitonic (descending run followed by an ascending run): log2(n) steps leave it sorted descending.
template <int kThreads>
__device__ __forceinline__ void block_bitonic_merge_desc(uint64_t* a, int n) {
    for (int stride = n >> 1; stride > 0; stride >>= 1) {
        for (int i = threadIdx.x; i < (n >> 1); i += kThreads) {
            const int lo = 2 * i - (i & (stride - 1));
            const int hi = lo + stride;
            const uint64_t x = a[lo], y = a[hi];
            if (x < y) {
                a[lo] = y;
                a[hi] = x;
            }
        }
        __syncthreads();
    }
}
#endif  // __CUDACC__

}  // namespace pq
