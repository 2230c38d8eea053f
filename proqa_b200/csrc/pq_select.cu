// proqa_b200 — selection / merge kernels and row preparation.
//
//  * pq_merge_lists_kernel : per query, fold any number of candidate key lists into the final
//    sorted top-k and write FAISS-shaped outputs (D fp32 [nq,k], I int64 [nq,k], padding
//    I=-1 / D=-FLT_MAX as IndexFlat::search does when k > ntotal; L2 reports squared distances
//    ascending, clamped at 0).  Used after the fp32 scan (one list per CTA).
//  * pq_merge_di_kernel    : fold G already-final (D, I) lists (one per GPU shard, after the
//    NCCL all-gather) into one; ties resolve to the lower global id.
//  * pq_prep_rows_kernel   : at add(): squared row norms, bf16 copy for the tensor-core filter,
//    running max norm and non-finite detection.
#include "pq_common.cuh"
#include "pq_internal.h"

namespace pq {

constexpr int kSelThreads = 256;

struct MergeParams {
    MergeLaunch a;
    int work;  // power-of-two size of the shared work array
};

__device__ __forceinline__ void emit_result(const MergeLaunch& a, int q, int i, uint64_t key) {
    float d;
    long long id;
    if (key == 0ull) {
        id = -1;
        d = (a.metric == kMetricL2) ? FLT_MAX : -FLT_MAX;
    } else {
        id = (long long)key_row(key) + a.id_base;
        const float s = key_score(key);
        d = (a.metric == kMetricL2) ? fmaxf(0.f, a.q_norms[q] - s) : s;
    }
    if (a.D) a.D[(size_t)q * a.k + i] = d;
    if (a.I) a.I[(size_t)q * a.k + i] = id;
    if (a.out_keys) a.out_keys[(size_t)q * a.k + i] = key;
}

// Candidates are examined 1024 at a time (4 per thread); accepted ones (non-empty, not below the running k-th key) are
// appended to the shared work array, which is sorted back to its best k only when it is about to overflow — so the number
// of sorts follows the number of ACCEPTED candidates, not the number examined (148 lists x k = 5000 used to cost a sort
// per round).  After every sort the k-th key becomes the new admission bar.
__global__ void __launch_bounds__(kSelThreads) pq_merge_lists_kernel(const MergeParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* work = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ int s_fill;
    __shared__ unsigned long long s_bar;
    const MergeLaunch& a = p.a;
    const int q = blockIdx.x;
    const int t = threadIdx.x;
    const int qb = a.batch_q > 0 ? q / a.batch_q : 0, qi = a.batch_q > 0 ? q - qb * a.batch_q : q;   // (batch, query in the batch)
    const uint64_t* base = a.keys + (size_t)qb * a.batch_stride + (size_t)qi * a.q_stride;
    const long long total = (long long)a.n_lists * a.list_len;
    constexpr int kPer = 4, kStep = kSelThreads * kPer;

    if (t == 0) {
        s_fill = 0;
        s_bar = a.gthr ? ((unsigned long long)a.gthr[a.batch_q > 0 ? qb * a.gthr_batch_stride + qi : q] << 32) : 0ull;
    }
    __syncthreads();
    for (long long r0 = 0; r0 < total; r0 += kStep) {
        const unsigned long long bar = s_bar;
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const long long idx = r0 + u * kSelThreads + t;
            if (idx < total) {
                const int list = (int)(idx / a.list_len);
                const int pos = (int)(idx - (long long)list * a.list_len);
                if (!a.counts || pos < (int)a.counts[(size_t)q * a.cnt_q_stride + list]) {
                    const uint64_t key = base[(size_t)list * a.list_stride + pos];
                    if (key != 0ull && key >= bar) work[atomicAdd(&s_fill, 1)] = key;
                }
            }
        }
        __syncthreads();
        const int fill = s_fill;
        __syncthreads();  // every thread has read the fill before the next batch's appends move it (the branch below must stay uniform)
        if (fill + kStep > p.work) {  // the next batch might not fit: keep the best k
            for (int i = fill + t; i < p.work; i += kSelThreads) work[i] = 0ull;
            __syncthreads();
            block_sort_desc<kSelThreads>(work, p.work);
            if (t == 0) {
                s_fill = fill < a.k ? fill : a.k;
                if (fill >= a.k && work[a.k - 1] > s_bar) s_bar = work[a.k - 1];
            }
            __syncthreads();
        }
    }
    {
        const int fill = s_fill;
        for (int i = fill + t; i < p.work; i += kSelThreads) work[i] = 0ull;
        __syncthreads();
        block_sort_desc<kSelThreads>(work, p.work);
    }
    for (int i = t; i < a.k; i += kSelThreads) emit_result(a, q, i, work[i]);
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

cudaError_t merge_lists_launch(const MergeLaunch& a, cudaStream_t stream) {
    if (a.nq <= 0) return cudaSuccess;
    MergeParams p;
    p.a = a;
    p.work = next_pow2(a.k + 1024);
    if (p.work < 2048) p.work = 2048;
    const size_t smem = (size_t)p.work * 8;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(pq_merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pq_merge_lists_kernel<<<a.nq, kSelThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// (D, I) list merge across shards.  Lists arrive ordered by shard rank; shards own ascending,
// disjoint id ranges and each list is already in (score desc, id asc) order, so the position
// g*k + i is a valid tie-break proxy for the global id.
// ------------------------------------------------------------------------------------------------
struct MergeDIParams {
    const float* D_in;
    const long long* I_in;
    float* D_out;
    long long* I_out;
    int n_lists, nq, k, metric, work;
};

__global__ void __launch_bounds__(kSelThreads) pq_merge_di_kernel(const MergeDIParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* work = reinterpret_cast<uint64_t*>(smem_raw);
    const int q = blockIdx.x;
    const int t = threadIdx.x;
    const int total = p.n_lists * p.k;
    // total <= work is guaranteed by the launcher
    for (int idx = t; idx < p.work; idx += kSelThreads) {
        uint64_t key = 0ull;
        if (idx < total) {
            const int g = idx / p.k, i = idx - g * p.k;
            const size_t off = ((size_t)g * p.nq + q) * p.k + i;
            const long long id = p.I_in[off];
            if (id >= 0) {
                const float d = p.D_in[off];
                key = make_key(p.metric == kMetricL2 ? -d : d, (uint32_t)idx);
            }
        }
        work[idx] = key;
    }
    __syncthreads();
    block_sort_desc<kSelThreads>(work, p.work);
    for (int i = t; i < p.k; i += kSelThreads) {
        const uint64_t key = work[i];
        float d;
        long long id;
        if (key == 0ull) {
            id = -1;
            d = (p.metric == kMetricL2) ? FLT_MAX : -FLT_MAX;
        } else {
            const int idx = (int)key_row(key);
            const int g = idx / p.k, j = idx - g * p.k;
            const size_t off = ((size_t)g * p.nq + q) * p.k + j;
            id = p.I_in[off];
            d = p.D_in[off];
        }
        p.D_out[(size_t)q * p.k + i] = d;
        p.I_out[(size_t)q * p.k + i] = id;
    }
}

// The same merge without shared memory, for n_lists * k beyond what one CTA can sort (two shards at k = 10000): every entry
// finds its output position directly — its index in its own list plus, by binary search, the number of entries of every
// other list that come before it.  "Before" is the order pq_merge_di_kernel sorts by: score, then (list, position) — a
// total order even when an L2 list holds equal distances whose ids are not ascending (the engine orders L2 lists by
// 2<q,x> - |x|^2 before clamping and rounding |q|^2 - that to D), so output positions never collide.  Within one list the
// scores are non-increasing (best first), which is all the binary search needs.
constexpr int kMergeRankMaxLists = 64;

__global__ void __launch_bounds__(kSelThreads) pq_merge_di_rank_kernel(const MergeDIParams p) {
    __shared__ int s_len[kMergeRankMaxLists];  // valid entries per list (the -1 padding is a suffix)
    const int q = blockIdx.x;
    const int t = threadIdx.x;
    const bool l2 = p.metric == kMetricL2;
    for (int g = t; g < p.n_lists; g += kSelThreads) {
        const long long* ids = p.I_in + ((size_t)g * p.nq + q) * p.k;
        int lo = 0, hi = p.k;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ids[mid] >= 0) lo = mid + 1;
            else hi = mid;
        }
        s_len[g] = lo;
    }
    __syncthreads();
    int total_valid = 0;
    for (int g = 0; g < p.n_lists; ++g) total_valid += s_len[g];
    for (int e = t; e < p.n_lists * p.k; e += kSelThreads) {
        const int g = e / p.k, i = e - g * p.k;
        if (i >= s_len[g]) continue;
        const size_t off = ((size_t)g * p.nq + q) * p.k + i;
        const float d = p.D_in[off];
        int pos = i;
        for (int h = 0; h < p.n_lists && pos < p.k; ++h) {
            if (h == g) continue;
            // entries of list h that precede (d, g, i): strictly better scores, and equal scores when h < g
            const float* Dh = p.D_in + ((size_t)h * p.nq + q) * p.k;
            int lo = 0, hi = s_len[h];
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const float dm = Dh[mid];
                const bool before = (dm != d) ? (l2 ? dm < d : dm > d) : (h < g);
                if (before) lo = mid + 1;
                else hi = mid;
            }
            pos += lo;
        }
        if (pos < p.k) {
            p.D_out[(size_t)q * p.k + pos] = d;
            p.I_out[(size_t)q * p.k + pos] = p.I_in[off];
        }
    }
    for (int i = total_valid + t; i < p.k; i += kSelThreads) {
        p.D_out[(size_t)q * p.k + i] = l2 ? FLT_MAX : -FLT_MAX;
        p.I_out[(size_t)q * p.k + i] = -1;
    }
}

cudaError_t merge_di_launch(const float* D_in, const long long* I_in, int n_lists, int nq, int k, int metric, float* D_out,
                            long long* I_out, cudaStream_t stream) {
    if (nq <= 0) return cudaSuccess;
    MergeDIParams p{D_in, I_in, D_out, I_out, n_lists, nq, k, metric, 0};
    p.work = next_pow2(n_lists * k);
    if (p.work < 2) p.work = 2;
    const size_t smem = (size_t)p.work * 8;
    if (smem > 200 * 1024) {  // too many keys for the in-CTA sort: position-by-ranking kernel
        if (n_lists > kMergeRankMaxLists) return cudaErrorInvalidValue;
        pq_merge_di_rank_kernel<<<nq, kSelThreads, 0, stream>>>(p);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(pq_merge_di_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pq_merge_di_kernel<<<nq, kSelThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// add(): per-row preparation.  One warp per row; lane l owns dims 4l..4l+3 for the bf16 copy,
// the squared norm is engine_dot(row,row) (part of the engine's defined L2 score).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pq_prep_rows_kernel(const float* __restrict__ rows, long long n, uint16_t* __restrict__ rows_bf16,
                                                           float* __restrict__ norms, uint32_t* max_norm_bits,
                                                           uint32_t* nonfinite_flag, uint8_t* __restrict__ row_bad,
                                                           float* __restrict__ resid2, uint32_t* max_resid_bits) {
    // Four rows per warp pass: lane = 8 r + j owns dims 16j .. 16j+15 of row r — chain p_j of engine_dot(row, row) runs in one
    // lane, three butterfly steps inside the group of eight add the chains in engine_dot's tree (as quad_engine_dot does).  A row per
    // warp pass with the chains handed from lane to lane cost ~35 instructions per row: 0.33 ms per million rows, 2.8x the time the
    // 776 bytes per row take to move.
    const int lane = threadIdx.x & 31, r = lane >> 3, j = lane & 7;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    float local_max = 0.f, local_max_resid = 0.f;
    bool bad = false;
    for (long long row0 = warp * 4; row0 < n; row0 += n_warps * 4) {
        const long long row = row0 + r;
        const bool in = row < n;
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = in ? *reinterpret_cast<const float4*>(rows + row * kDim + 16 * j + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float acc = 0.f, r2 = 0.f;
        bool lane_bad = false;
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc = fmaf(v[i].x, v[i].x, acc);
            acc = fmaf(v[i].y, v[i].y, acc);
            acc = fmaf(v[i].z, v[i].z, acc);
            acc = fmaf(v[i].w, v[i].w, acc);
            // bf16 copy (round to nearest even)
            __nv_bfloat162 lo = __floats2bfloat162_rn(v[i].x, v[i].y);
            __nv_bfloat162 hi = __floats2bfloat162_rn(v[i].z, v[i].w);
            packed[2 * i] = *reinterpret_cast<uint32_t*>(&lo);
            packed[2 * i + 1] = *reinterpret_cast<uint32_t*>(&hi);
            const float bx = __low2float(lo), by = __high2float(lo), bz = __low2float(hi), bw = __high2float(hi);
            // non-finite after rounding (covers inf/NaN inputs and fp32 values that overflow bf16)
            lane_bad |= !(isfinite(bx) && isfinite(by) && isfinite(bz) && isfinite(bw));
            // squared norm of what the bf16 rounding took away, |x - bf16(x)|^2 (drives the filter's error bound);
            // any summation order will do, the result is inflated to be an upper bound
            const float dx = v[i].x - bx, dy = v[i].y - by, dz = v[i].z - bz, dw = v[i].w - bw;
            r2 += fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
        }
        if (in) {
            uint4* dst = reinterpret_cast<uint4*>(rows_bf16 + row * kDim + 16 * j);
            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
        // the eight chains of a row: ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7))
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        r2 += __shfl_xor_sync(0xffffffffu, r2, 1);
        r2 += __shfl_xor_sync(0xffffffffu, r2, 2);
        r2 += __shfl_xor_sync(0xffffffffu, r2, 4);
        r2 *= 1.0001f;
        const unsigned bad_lanes = __ballot_sync(0xffffffffu, lane_bad);
        const bool row_is_bad = (((bad_lanes >> (8 * r)) & 0xffu) != 0u) || !(acc <= FLT_MAX);  // inf or NaN norm
        if (in) {
            if (j == 0) {
                norms[row] = acc;
                if (row_bad) row_bad[row] = row_is_bad ? 1 : 0;
                if (resid2) resid2[row] = r2;
            }
            bad |= row_is_bad;
            local_max = fmaxf(local_max, acc);
            if (!row_is_bad) local_max_resid = fmaxf(local_max_resid, r2);
        }
    }
    if (max_resid_bits && j == 0 && local_max_resid > 0.f) atomicMax(max_resid_bits, __float_as_uint(local_max_resid));
    if (nonfinite_flag && bad && j == 0) atomicOr(nonfinite_flag, 1u);
    if (max_norm_bits && j == 0 && local_max > 0.f) atomicMax(max_norm_bits, __float_as_uint(local_max));  // non-negative floats order as uints
}

cudaError_t prep_rows_launch(const float* rows, long long n, uint16_t* rows_bf16, float* norms, uint32_t* max_norm_bits,
                             uint32_t* nonfinite_flag, uint8_t* row_bad, float* resid2, uint32_t* max_resid_bits, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    long long blocks = (n + 31) / 32;   // 8 warps x 4 rows per CTA pass
    if (blocks > 148 * 16) blocks = 148 * 16;
    pq_prep_rows_kernel<<<(int)blocks, 256, 0, stream>>>(rows, n, rows_bf16, norms, max_norm_bits, nonfinite_flag, row_bad, resid2, max_resid_bits);
    return cudaGetLastError();
}

}  // namespace pq
