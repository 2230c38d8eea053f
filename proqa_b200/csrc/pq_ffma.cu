// proqa_b200 — exact fp32 streaming scan with fused top-k (the bandwidth-bound small-batch path).
//
// Replaces, for one batch of <=16 queries, what FAISS IndexFlat::search does on the host
// (reference call site retrieval/eval_retrieval.py:104; FAISS 1.6.3 knn_inner_product ->
// per-query heap, see oracle/flat_oracle.c).  Design:
//   * corpus tiles of 128 rows x 512 B stream HBM -> shared memory through TMA (4 boxes of
//     32 floats x 128 rows, SWIZZLE_128B, so that "one thread = one row" reads are bank-conflict
//     free), 2-3 stages of 64 KB in flight per SM, one persistent CTA per SM;
//   * queries sit in constant memory, so the inner loop is LDS.128 + FFMA with a constant-bank
//     operand: score(q, row) is the sequential chain acc = fmaf(row[i], q[i], acc), i = 0..127 —
//     the engine's *defined* fp32 score, restated bit-exactly in oracle/flat_oracle.c;
//   * top-k is fused: scores below the running per-query threshold are dropped in registers,
//     survivors are appended to a shared-memory buffer that is bitonic-sorted back to k entries
//     when it fills; thresholds are exchanged between CTAs through a global atomicMax so every
//     CTA filters with (nearly) the global k-th best score.  The score matrix never exists.
#include "pq_common.cuh"
#include "pq_internal.h"

namespace pq {

__constant__ float c_queries[kFfmaMaxQ * kDim];

struct FfmaParams {
    const float* row_norms;
    uint64_t* out_keys;
    uint32_t* gthr;
    long long n_rows;
    int tiles_per_cta;
    int n_tiles;
    int nq;
    int k;
    int cap;
    int n_stages;
    int metric;
};

struct FfmaCtrl {                 // lives right after the tile stages in shared memory
    uint64_t full[4];             // TMA completion barriers, one per stage
    int cnt[kFfmaMaxQ];           // entries currently buffered per query
    float thr[kFfmaMaxQ];         // current admission threshold per query
    int need_compact;
    int pad[3];
};

// Sort one query's buffer, keep the best k, refresh and publish its threshold.
__device__ __forceinline__ void ffma_compact(uint64_t* buf, FfmaCtrl* ctrl, int q, const FfmaParams& p) {
    const int n = ctrl->cnt[q];
    for (int i = n + threadIdx.x; i < p.cap; i += kFfmaThreads) buf[i] = 0ull;
    __syncthreads();
    block_sort_desc<kFfmaThreads>(buf, p.cap);
    if (threadIdx.x == 0) {
        float thr = ctrl->thr[q];
        if (n >= p.k) {
            ctrl->cnt[q] = p.k;
            const float kth = key_score(buf[p.k - 1]);
            thr = fmaxf(thr, kth);
            atomicMax(p.gthr + q, f32_to_ordered(thr));
        }
        thr = fmaxf(thr, ordered_to_f32(ld_volatile_u32(p.gthr + q)));
        ctrl->thr[q] = thr;
    }
    __syncthreads();
}

template <int QOFF, int QH>
__device__ __forceinline__ void ffma_tile_dots(const uint8_t* stage, int r, float (&acc)[QH]) {
#pragma unroll
    for (int j = 0; j < QH; ++j) acc[j] = 0.f;
    const uint8_t* prow = stage + r * 128;
    const int sw = (r & 7) << 4;
#pragma unroll
    for (int pnl = 0; pnl < 4; ++pnl) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(prow + pnl * (kFfmaTileRows * 128) + ((c << 4) ^ sw));
            const int d0 = pnl * 32 + c * 4;
#pragma unroll
            for (int j = 0; j < QH; ++j) {
                acc[j] = fmaf(v.x, c_queries[(QOFF + j) * kDim + d0 + 0], acc[j]);
                acc[j] = fmaf(v.y, c_queries[(QOFF + j) * kDim + d0 + 1], acc[j]);
                acc[j] = fmaf(v.z, c_queries[(QOFF + j) * kDim + d0 + 2], acc[j]);
                acc[j] = fmaf(v.w, c_queries[(QOFF + j) * kDim + d0 + 3], acc[j]);
            }
        }
    }
}

template <int QOFF, int QH>
__device__ __forceinline__ void ffma_tile_half(const uint8_t* stage, int r, long long row, bool valid, uint64_t* bufs,
                                               FfmaCtrl* ctrl, const FfmaParams& p) {
    float acc[QH];
    ffma_tile_dots<QOFF, QH>(stage, r, acc);
    if (!valid) return;
    float bias = 0.f;
    if (p.metric == kMetricL2) bias = __ldg(p.row_norms + row);
#pragma unroll
    for (int j = 0; j < QH; ++j) {
        const int q = QOFF + j;
        if (q < p.nq) {
            float s = acc[j];
            if (p.metric == kMetricL2) s = fmaf(2.f, s, -bias);  // larger is closer: 2<q,x> - |x|^2
            if (s >= ctrl->thr[q]) {
                const int slot = atomicAdd(&ctrl->cnt[q], 1);
                bufs[(size_t)q * p.cap + slot] = make_key(s, (uint32_t)row);
                if (slot + 1 > p.cap - kFfmaTileRows) ctrl->need_compact = 1;
            }
        }
    }
}

template <int QH>
__global__ void __launch_bounds__(kFfmaThreads, 1)
pq_ffma_scan_kernel(const __grid_constant__ CUtensorMap tmap, const FfmaParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stages = smem;
    FfmaCtrl* ctrl = reinterpret_cast<FfmaCtrl*>(smem + (size_t)p.n_stages * kFfmaStageBytes);
    uint64_t* bufs = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * kFfmaStageBytes + 256);

    const int t = threadIdx.x;
    const int r = t & (kFfmaTileRows - 1);
    const int tile0 = blockIdx.x * p.tiles_per_cta;
    int ntiles = p.n_tiles - tile0;
    ntiles = ntiles < 0 ? 0 : (ntiles > p.tiles_per_cta ? p.tiles_per_cta : ntiles);

    if (t == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < p.n_stages; ++s) mbar_init(&ctrl->full[s], 1);
        fence_mbar_init();
        ctrl->need_compact = 0;
    }
    if (t < kFfmaMaxQ) {
        ctrl->cnt[t] = 0;
        ctrl->thr[t] = (t < p.nq) ? fmaxf(PQ_THR_FLOOR, ordered_to_f32(ld_volatile_u32(p.gthr + t))) : INFINITY;
    }
    __syncthreads();

    auto issue = [&](int it) {  // thread 0 only
        const int s = it % p.n_stages;
        uint8_t* dst = stages + (size_t)s * kFfmaStageBytes;
        mbar_arrive_expect_tx(&ctrl->full[s], kFfmaStageBytes);
        const int y = (tile0 + it) * kFfmaTileRows;
#pragma unroll
        for (int pnl = 0; pnl < 4; ++pnl) tma_load_2d(dst + pnl * (kFfmaTileRows * 128), &tmap, pnl * 32, y, &ctrl->full[s]);
    };
    if (t == 0) {
        for (int it = 0; it < p.n_stages - 1 && it < ntiles; ++it) issue(it);
    }

    for (int it = 0; it < ntiles; ++it) {
        // The stage being refilled was consumed in iteration it-1; the __syncthreads() that ended
        // that iteration is what makes the overwrite safe.
        if (t == 0 && it + p.n_stages - 1 < ntiles) issue(it + p.n_stages - 1);
        const int s = it % p.n_stages;
        mbar_wait(&ctrl->full[s], (uint32_t)((it / p.n_stages) & 1));
        const uint8_t* stage = stages + (size_t)s * kFfmaStageBytes;
        const long long row = (long long)(tile0 + it) * kFfmaTileRows + r;
        const bool valid = row < p.n_rows;
        if (t < kFfmaTileRows) {
            ffma_tile_half<0, QH>(stage, r, row, valid, bufs, ctrl, p);
        } else if (p.nq > QH) {
            ffma_tile_half<QH, QH>(stage, r, row, valid, bufs, ctrl, p);
        }
        __syncthreads();
        const bool refresh = ((it & 15) == 15);
        if (ctrl->need_compact) {  // block-uniform: written before the barrier above
            for (int q = 0; q < p.nq; ++q) {
                if (ctrl->cnt[q] > p.cap - kFfmaTileRows) ffma_compact(bufs + (size_t)q * p.cap, ctrl, q, p);
            }
            if (t == 0) ctrl->need_compact = 0;
            __syncthreads();
        } else if (refresh) {
            if (t < p.nq) ctrl->thr[t] = fmaxf(ctrl->thr[t], ordered_to_f32(ld_volatile_u32(p.gthr + t)));
            __syncthreads();
        }
    }

    // Final: every query's buffer sorted, best k written out (zero keys pad short lists).
    for (int q = 0; q < p.nq; ++q) {
        uint64_t* buf = bufs + (size_t)q * p.cap;
        ffma_compact(buf, ctrl, q, p);
        const int n = ctrl->cnt[q] < p.k ? ctrl->cnt[q] : p.k;
        uint64_t* out = p.out_keys + ((size_t)blockIdx.x * p.nq + q) * p.k;
        for (int i = t; i < p.k; i += kFfmaThreads) out[i] = (i < n) ? buf[i] : 0ull;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
static constexpr int kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA on sm_100

static int ffma_cap_for_k(int k) { return next_pow2(k + kFfmaTileRows) < 256 ? 256 : next_pow2(k + kFfmaTileRows); }

int ffma_max_queries_for_k(int k) {
    if (k < 1) return 0;
    const long long cap = ffma_cap_for_k(k);
    long long q = (kSmemLimit - 2LL * kFfmaStageBytes - 256 - 1024) / (cap * 8);  // prefer a double-buffered ring
    if (q < 1) q = (kSmemLimit - 1LL * kFfmaStageBytes - 256 - 1024) / (cap * 8); // huge k: single stage
    if (q > kFfmaMaxQ) q = kFfmaMaxQ;
    return (int)q;
}

cudaError_t ffma_scan_launch(const FfmaLaunch& a, cudaStream_t stream) {
    FfmaParams p;
    p.row_norms = a.row_norms;
    p.out_keys = a.out_keys;
    p.gthr = a.gthr;
    p.n_rows = a.n_rows;
    p.n_tiles = (int)((a.n_rows + kFfmaTileRows - 1) / kFfmaTileRows);
    p.tiles_per_cta = (p.n_tiles + a.n_ctas - 1) / a.n_ctas;
    p.nq = a.nq;
    p.k = a.k;
    p.cap = ffma_cap_for_k(a.k);
    p.metric = a.metric;
    const size_t buf_bytes = (size_t)a.nq * p.cap * 8 + 256;
    p.n_stages = 3;
    while (p.n_stages > 1 && (size_t)p.n_stages * kFfmaStageBytes + buf_bytes > (size_t)kSmemLimit) --p.n_stages;
    const size_t smem = (size_t)p.n_stages * kFfmaStageBytes + buf_bytes;
    if (a.nq < 1 || a.nq > kFfmaMaxQ || smem > (size_t)kSmemLimit) return cudaErrorInvalidValue;

    cudaError_t e = cudaMemcpyToSymbolAsync(c_queries, a.queries_dev, (size_t)a.nq * kDim * sizeof(float), 0,
                                            cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return e;

    const int qh = a.nq <= 1 ? 1 : (a.nq <= 2 ? 1 : (a.nq <= 4 ? 2 : (a.nq <= 8 ? 4 : 8)));
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kern<<<a.n_ctas, kFfmaThreads, smem, stream>>>(*a.tmap_rows_f32, p);
        return cudaGetLastError();
    };
    switch (qh) {
        case 1: return launch(pq_ffma_scan_kernel<1>);
        case 2: return launch(pq_ffma_scan_kernel<2>);
        case 4: return launch(pq_ffma_scan_kernel<4>);
        default: return launch(pq_ffma_scan_kernel<8>);
    }
}

}  // namespace pq
