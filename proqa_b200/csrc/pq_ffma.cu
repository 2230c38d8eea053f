// proqa_b200 — exact fp32 streaming scan with fused top-k (the bandwidth-bound small-batch path).
//
// Replaces, for one batch of <= 8 queries, what FAISS IndexFlat::search does on the host
// (reference call site retrieval/eval_retrieval.py:104; FAISS 1.6.3 knn_inner_product ->
// per-query heap, see oracle/flat_oracle.c).  Design:
//   * corpus tiles of 128 rows x 512 B stream HBM -> shared memory through TMA (4 boxes of
//     32 floats x 128 rows, SWIZZLE_128B), 3 stages of 64 KB in flight per SM, one persistent
//     CTA per SM;
//   * queries sit in shared memory too; a lane owns a 16-dimension slice of 4 rows (lane = rg + 4*kq),
//     so one LDS.128 of row data feeds 4*QT FFMAs and one LDS.128 of query data feeds 16: the inner
//     loop is ~90 % FFMA.  The eight 16-dim partial chains of a row are folded across lanes with three
//     butterfly shuffles — exactly the combination tree of the engine's *defined* fp32 score
//     (pq_common.cuh: engine_dot; restated bit-exactly in oracle/flat_oracle.c);
//   * top-k is fused: scores below the running per-query threshold are dropped in registers,
//     survivors are appended to a shared-memory buffer that is bitonic-sorted back to k entries
//     when it fills; thresholds are exchanged between CTAs through a global atomicMax so every
//     CTA filters with (nearly) the global k-th best score.  The score matrix never exists.
#include "pq_common.cuh"
#include "pq_internal.h"

namespace pq {

struct FfmaParams {
    const float* queries;  // [nq][128] fp32, device
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
    int ctas_per_batch;    // CTAs [b * ctas_per_batch, ...) serve query batch b
    int nq_total;          // queries over all batches (the last batch may hold fewer than nq)
};

struct FfmaCtrl {                 // lives right after the tile stages in shared memory
    uint64_t full[4];             // TMA completion barriers, one per stage
    int cnt[kFfmaMaxQ];           // entries currently buffered per query
    float thr[kFfmaMaxQ];         // current admission threshold per query
    int need_compact;
    int pad[3];
};
constexpr int kFfmaCtrlBytes = 256;
constexpr int kFfmaQsmBytes = kFfmaMaxQ * kDim * 4;  // queries re-laid out as [q][step][kq] float4

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

// One butterfly step of the cross-lane reduction: lanes whose `hi` flag is clear keep the lower half of v[],
// the others the upper half; each adds the partner's partial for the half it keeps.  fp32 addition is
// commutative, so both partners of a pair produce identical bits — the combination tree of the engine's
// defined score (pq_common.cuh: engine_dot) is reproduced exactly.
template <int N>
__device__ __forceinline__ void ffma_fold(const float (&v)[N], float (&out)[(N > 1 ? N / 2 : 1)], bool hi, int lane_mask) {
    if constexpr (N > 1) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float mine = hi ? v[i + N / 2] : v[i];
            const float theirs = hi ? v[i] : v[i + N / 2];
            out[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, lane_mask);
        }
    } else {
        out[0] = v[0] + __shfl_xor_sync(0xffffffffu, v[0], lane_mask);
    }
}

// One 128-row tile against QT queries.  Lane = rg + 4*kq: rg (0..3) picks the row inside a group of four,
// kq (0..7) the 16-dimension slice of the dot product this lane accumulates (partial chain p_kq).  A warp covers
// 16 rows, the CTA's 8 warps the 128 rows of the tile.  Row and query operands both come from shared memory as
// LDS.128: per 16*QT FFMAs a thread issues 4 + QT loads, all bank-conflict free (rows through TMA's 128-byte
// swizzle, queries through the [q][step][kq] layout).
// Returns true if one of this thread's appends pushed a query's buffer past its refill mark (the caller ORs the flags of
// all threads at the tile barrier: a shared flag read after that barrier raced with the next tile's writers).
template <int QT>
__device__ __forceinline__ bool ffma_tile(const uint8_t* stage, const float4* qsm, long long tile_row0, uint64_t* bufs, FfmaCtrl* ctrl,
                                          const FfmaParams& p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rg = lane & 3, kq = lane >> 2;
    float acc[4 * QT];
#pragma unroll
    for (int i = 0; i < 4 * QT; ++i) acc[i] = 0.f;
    const uint8_t* panel = stage + (kq >> 1) * (kFfmaTileRows * 128);
#pragma unroll
    for (int step = 0; step < 4; ++step) {
        const int c = (kq & 1) * 4 + step;
        float4 rv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = warp * 16 + r * 4 + rg;
            rv[r] = *reinterpret_cast<const float4*>(panel + row * 128 + ((c ^ (row & 7)) << 4));
        }
#pragma unroll
        for (int j = 0; j < QT; ++j) {
            const float4 qv = qsm[(j * 4 + step) * 8 + kq];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float a = acc[r * QT + j];
                a = fmaf(rv[r].x, qv.x, a);
                a = fmaf(rv[r].y, qv.y, a);
                a = fmaf(rv[r].z, qv.z, a);
                a = fmaf(rv[r].w, qv.w, a);
                acc[r * QT + j] = a;
            }
        }
    }
    // (p0+p1), then +(p2+p3), then +(p4..p7): three butterfly steps over the kq bits of the lane index
    constexpr int N = 4 * QT;
    constexpr int N1 = N / 2, N2 = (N1 > 1 ? N1 / 2 : 1), N3 = (N2 > 1 ? N2 / 2 : 1);
    const bool b0 = kq & 1, b1 = kq & 2, b2 = kq & 4;
    float w1[N1], w2[N2], w3[N3];
    ffma_fold<N>(acc, w1, b0, 4);
    ffma_fold<N1>(w1, w2, b1, 8);
    ffma_fold<N2>(w2, w3, b2, 16);
    // flat index a = r*QT + j of the first value this lane ends up owning
    int a0 = (b0 ? N / 2 : 0) + (b1 ? N / 4 : 0);
    bool owner = true;
    if constexpr (N2 > 1) a0 += b2 ? N / 8 : 0;
    else owner = !b2;  // N == 4: both partners of the last step hold the same value, one reports it
    bool want_compact = false;
    if (!owner) return false;
#pragma unroll
    for (int i = 0; i < N3; ++i) {
        const int a = a0 + i;
        const int r = a / QT, q = a % QT;
        const long long row = tile_row0 + warp * 16 + r * 4 + rg;
        if (row < p.n_rows && q < p.nq) {
            float s = w3[i];
            if (p.metric == kMetricL2) s = fmaf(2.f, s, -__ldg(p.row_norms + row));  // larger is closer: 2<q,x> - |x|^2
            if (s >= ctrl->thr[q]) {
                const int slot = atomicAdd(&ctrl->cnt[q], 1);
                bufs[(size_t)q * p.cap + slot] = make_key(s, (uint32_t)row);
                if (slot + 1 > p.cap - kFfmaTileRows) want_compact = true;
            }
        }
    }
    return want_compact;
}

template <int QT>
__global__ void __launch_bounds__(kFfmaThreads, 1)
pq_ffma_scan_kernel(const __grid_constant__ CUtensorMap tmap, const FfmaParams p_all) {
    // this CTA's query batch: its queries, thresholds and output lists
    FfmaParams p = p_all;
    const int batch = (int)blockIdx.x / p_all.ctas_per_batch;
    const int cta = (int)blockIdx.x - batch * p_all.ctas_per_batch;
    p.queries = p_all.queries + (size_t)batch * p_all.nq * kDim;
    p.gthr = p_all.gthr + (size_t)batch * kFfmaMaxQ;
    p.out_keys = p_all.out_keys + (size_t)batch * p_all.ctas_per_batch * p_all.nq * p_all.k;
    p.nq = min(p_all.nq, p_all.nq_total - batch * p_all.nq);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stages = smem;
    uint8_t* tail = smem + (size_t)p.n_stages * kFfmaStageBytes;
    FfmaCtrl* ctrl = reinterpret_cast<FfmaCtrl*>(tail);
    float4* qsm = reinterpret_cast<float4*>(tail + kFfmaCtrlBytes);
    uint64_t* bufs = reinterpret_cast<uint64_t*>(tail + kFfmaCtrlBytes + kFfmaQsmBytes);

    const int t = threadIdx.x;
    const int tile0 = cta * p.tiles_per_cta;
    int ntiles = p.n_tiles - tile0;
    ntiles = ntiles < 0 ? 0 : (ntiles > p.tiles_per_cta ? p.tiles_per_cta : ntiles);

    if (t == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < p.n_stages; ++s) mbar_init(&ctrl->full[s], 1);
        fence_mbar_init();
    }
    if (t < kFfmaMaxQ) {
        ctrl->cnt[t] = 0;
        ctrl->thr[t] = (t < p.nq) ? fmaxf(PQ_THR_FLOOR, ordered_to_f32(ld_volatile_u32(p.gthr + t))) : INFINITY;
    }
    {   // queries: global [q][128] -> shared [q][step][kq] float4 (dims 16*kq + 4*step .. +3); unused slots are zero
        const int q = t >> 5, c = t & 31;  // 256 threads = 8 queries x 32 float4
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < p.nq) v = __ldg(reinterpret_cast<const float4*>(p.queries + (size_t)q * kDim) + c);
        qsm[(q * 4 + (c & 3)) * 8 + (c >> 2)] = v;
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
        const bool want = ffma_tile<QT>(stages + (size_t)s * kFfmaStageBytes, qsm, (long long)(tile0 + it) * kFfmaTileRows, bufs, ctrl, p);
        const int need_compact = __syncthreads_or(want ? 1 : 0);  // also the barrier that frees the stage for the next TMA
        const bool refresh = ((it & 15) == 15);
        if (need_compact) {  // block-uniform
            for (int q = 0; q < p.nq; ++q) {
                if (ctrl->cnt[q] > p.cap - kFfmaTileRows) ffma_compact(bufs + (size_t)q * p.cap, ctrl, q, p);
            }
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
        uint64_t* out = p.out_keys + ((size_t)cta * p_all.nq + q) * p.k;   // (list stride: the full batch size, also in a short batch)
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
    const long long fixed = kFfmaCtrlBytes + kFfmaQsmBytes + 1024;
    long long q = (kSmemLimit - 2LL * kFfmaStageBytes - fixed) / (cap * 8);  // prefer a double-buffered ring
    if (q < 1) q = (kSmemLimit - 1LL * kFfmaStageBytes - fixed) / (cap * 8); // huge k: single stage
    if (q > kFfmaMaxQ) q = kFfmaMaxQ;
    return (int)q;
}

cudaError_t ffma_scan_launch(const FfmaLaunch& a, cudaStream_t stream) {
    FfmaParams p;
    p.queries = a.queries_dev;
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
    p.ctas_per_batch = a.n_ctas;
    p.nq_total = a.nq_total > 0 ? a.nq_total : a.nq;
    const int n_batches = a.n_batches > 0 ? a.n_batches : 1;
    const size_t buf_bytes = (size_t)a.nq * p.cap * 8 + kFfmaCtrlBytes + kFfmaQsmBytes;
    p.n_stages = 3;
    while (p.n_stages > 1 && (size_t)p.n_stages * kFfmaStageBytes + buf_bytes > (size_t)kSmemLimit) --p.n_stages;
    const size_t smem = (size_t)p.n_stages * kFfmaStageBytes + buf_bytes;
    if (a.nq < 1 || a.nq > kFfmaMaxQ || smem > (size_t)kSmemLimit) return cudaErrorInvalidValue;

    const int qh = a.nq <= 1 ? 1 : (a.nq <= 2 ? 2 : (a.nq <= 4 ? 4 : 8));
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kern<<<a.n_ctas * n_batches, kFfmaThreads, smem, stream>>>(*a.tmap_rows_f32, p);
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
