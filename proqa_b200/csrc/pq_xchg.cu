// proqa_b200 — the exchange step of the row-sharded search, fused with its merge, over NVLink peer memory.
//
// north_star (4): "merges the per-shard k x (score, id) lists with an all-gather over NVLink plus a merge kernel".  Done with a
// library all-gather, every GPU receives every shard's whole list (R x nq x k x 12 B) and then merges all nq queries itself:
// 8 GPUs x C3 (65,536 queries, k = 80) = 503 MB into every GPU and eight identical merges — measured 10.7 ms + 2.9 ms per
// search next to 36 ms of scoring.  Here the GPUs write straight into each other's HBM (buffers mapped through CUDA IPC, or peer
// access inside one process) and the merge is spread over them:
//
//   stage 1  scatter   rank (rr, rq) cuts its list [n_loc, k] into R query slices and stores slice p into the inbox of the rank
//                      that holds row shard p of its row group                     (1/R of the list to each peer)
//   stage 2  merge     every rank merges the R lists of ITS slice (n_loc / R queries; same order rule as pq_merge_shard_results:
//                      score, then shard, then position) and stores the merged slice into the result buffer of EVERY rank of
//                      the job — the final [nq, k] result assembles itself in place on all GPUs
//   stage 3  collect   each rank waits until every slice has arrived and copies the result to the caller's tensors
//
// Ordering: data stores are followed, in the next kernel on the same stream, by a __threadfence_system() and one flag store
// per destination (the search's sequence number); consumers spin on flags in their OWN memory.  R = 1 (queries split, corpus
// replicated) degenerates to "push my finished slice to everybody" — the same code replaces the all-gather of that layout.
// Per search and GPU this moves (n_loc/R + nq) x k x 12 B instead of R x n_loc x k x 12 B, and merges n_loc/R queries, not n_loc.
#include "pq_common.cuh"
#include "pq_host.h"

#include <string.h>

#include <algorithm>
#include <new>

namespace pq {

constexpr int kXchgMaxRanks = 16;
constexpr size_t kXchgHeaderBytes = 4096;   // flag1[16] @0, flag2[16] @128, error word @256
constexpr unsigned long long kXchgTimeoutNs = 10ull * 1000ull * 1000ull * 1000ull;

struct XchgParams {
    int world, rank, R, Q, rr, rq;
    int n_loc, per;                 // queries of this rank's query group; queries per merge slice = ceil(n_loc / R)
    int nq, k, metric;
    long long q_base;               // first global query of this rank's query group
    unsigned long long seq;
    size_t in_off, out_off;         // byte offsets of the inbox / result areas inside every rank's buffer
    size_t slot_bytes;              // inbox bytes per sender: per x k scores, then (16-byte aligned) per x k ids
    size_t slot_i_off;              // byte offset of the ids inside a slot
    size_t out_i_off;               // byte offset of the id part inside the result area
    uint8_t* peer[kXchgMaxRanks];   // base address of every rank's buffer as seen from this device (own one included)
};

__device__ __forceinline__ uint64_t xchg_ld_flag(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint64_t xchg_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Thread 0 waits until flags[0..n) have all reached seq, the block follows.  A peer that never arrives (it died) ends the wait
// after 10 s with the error word set: results are garbage then and pq_xchg_check reports it.
__device__ __forceinline__ void xchg_wait_flags(const uint64_t* flags, int n, unsigned long long seq, uint64_t* err) {
    if (threadIdx.x == 0) {
        const uint64_t t0 = xchg_timer_ns();
        for (int g = 0; g < n; ++g) {
            while (xchg_ld_flag(flags + g) < seq) {
                if (xchg_timer_ns() - t0 > kXchgTimeoutNs) {
                    *err = 1ull;
                    break;
                }
            }
        }
    }
    __syncthreads();
    __threadfence_system();
}

// stage 1: blockIdx.y = destination row shard p; grid-stride copy of my slice p (scores, then ids) into its inbox slot rr
__global__ void __launch_bounds__(256) pq_xchg_scatter_kernel(const XchgParams x, const float* __restrict__ D_local, const long long* __restrict__ I_local) {
    const int p = blockIdx.y;
    const long long q0 = (long long)p * x.per, q1 = min((long long)x.n_loc, q0 + x.per);
    if (q1 <= q0) return;
    const long long n = (q1 - q0) * x.k;
    uint8_t* slot = x.peer[x.rq * x.R + p] + x.in_off + (size_t)x.rr * x.slot_bytes;
    float* Dd = reinterpret_cast<float*>(slot);
    long long* Id = reinterpret_cast<long long*>(slot + x.slot_i_off);
    const float* Ds = D_local + q0 * x.k;
    const long long* Is = I_local + q0 * x.k;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Dd[i] = Ds[i];
        Id[i] = Is[i];
    }
}

// after a data kernel: make its stores visible system-wide, then raise my flag at every destination
__global__ void pq_xchg_signal_kernel(const XchgParams x, int stage) {
    __threadfence_system();
    const int t = threadIdx.x;
    if (stage == 1) {
        if (t < x.R) {
            uint64_t* f = reinterpret_cast<uint64_t*>(x.peer[x.rq * x.R + t]) + x.rr;
            asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(f), "l"(x.seq) : "memory");
        }
    } else if (t < x.world) {
        uint64_t* f = reinterpret_cast<uint64_t*>(x.peer[t]) + 16 + x.rank;
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(f), "l"(x.seq) : "memory");
    }
}

__device__ __forceinline__ void xchg_emit(const XchgParams& x, long long q_global, int i, float d, long long id) {
    const size_t e = (size_t)q_global * x.k + i;
    for (int t = 0; t < x.world; ++t) {
        uint8_t* out = x.peer[t] + x.out_off;
        reinterpret_cast<float*>(out)[e] = d;
        reinterpret_cast<long long*>(out + x.out_i_off)[e] = id;
    }
}

// stage 2, R x k <= 16384: one CTA per query of my slice, the R lists sorted together in shared memory.
// Key = (score, ~(shard * k + position)): best score first, then lower shard, then earlier position — the order of pq_merge_di_kernel.
__global__ void __launch_bounds__(256) pq_xchg_merge_sort_kernel(const XchgParams x, int work) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    uint8_t* mine = x.peer[x.rank];
    xchg_wait_flags(reinterpret_cast<const uint64_t*>(mine), x.R, x.seq, reinterpret_cast<uint64_t*>(mine) + 32);
    const int n_mine = (int)max(0LL, min((long long)x.n_loc, (long long)(x.rr + 1) * x.per) - (long long)x.rr * x.per);
    const uint8_t* in = mine + x.in_off;
    const bool l2 = x.metric == kMetricL2;
    for (int qi = blockIdx.x; qi < n_mine; qi += gridDim.x) {
        const int total = x.R * x.k;
        for (int idx = threadIdx.x; idx < work; idx += 256) {
            uint64_t key = 0ull;
            if (idx < total) {
                const int g = idx / x.k, i = idx - g * x.k;
                const uint8_t* slot = in + (size_t)g * x.slot_bytes;
                const long long id = reinterpret_cast<const long long*>(slot + x.slot_i_off)[(size_t)qi * x.k + i];
                if (id >= 0) {
                    const float d = reinterpret_cast<const float*>(slot)[(size_t)qi * x.k + i];
                    key = make_key(l2 ? -d : d, (uint32_t)idx);
                }
            }
            keys[idx] = key;
        }
        __syncthreads();
        block_sort_desc<256>(keys, work);
        const long long qg = x.q_base + (long long)x.rr * x.per + qi;
        for (int i = threadIdx.x; i < x.k; i += 256) {
            const uint64_t key = keys[i];
            float d = l2 ? FLT_MAX : -FLT_MAX;
            long long id = -1;
            if (key != 0ull) {
                const int idx = (int)key_row(key);
                const int g = idx / x.k, j = idx - g * x.k;
                const uint8_t* slot = in + (size_t)g * x.slot_bytes;
                d = reinterpret_cast<const float*>(slot)[(size_t)qi * x.k + j];
                id = reinterpret_cast<const long long*>(slot + x.slot_i_off)[(size_t)qi * x.k + j];
            }
            xchg_emit(x, qg, i, d, id);
        }
        __syncthreads();
    }
}

// stage 2 beyond the in-CTA sort (two shards at k = 10000): every entry finds its output position by ranking against the other
// lists (binary searches; the order rule of pq_merge_di_rank_kernel).  R <= 16 here.
__global__ void __launch_bounds__(256) pq_xchg_merge_rank_kernel(const XchgParams x) {
    __shared__ int s_len[kXchgMaxRanks];
    uint8_t* mine = x.peer[x.rank];
    xchg_wait_flags(reinterpret_cast<const uint64_t*>(mine), x.R, x.seq, reinterpret_cast<uint64_t*>(mine) + 32);
    const int n_mine = (int)max(0LL, min((long long)x.n_loc, (long long)(x.rr + 1) * x.per) - (long long)x.rr * x.per);
    const uint8_t* in = mine + x.in_off;
    const bool l2 = x.metric == kMetricL2;
    const int t = threadIdx.x;
    for (int qi = blockIdx.x; qi < n_mine; qi += gridDim.x) {
        if (t < x.R) {
            const long long* ids = reinterpret_cast<const long long*>(in + (size_t)t * x.slot_bytes + x.slot_i_off) + (size_t)qi * x.k;
            int lo = 0, hi = x.k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ids[mid] >= 0) lo = mid + 1;
                else hi = mid;
            }
            s_len[t] = lo;
        }
        __syncthreads();
        int total_valid = 0;
        for (int g = 0; g < x.R; ++g) total_valid += s_len[g];
        const long long qg = x.q_base + (long long)x.rr * x.per + qi;
        for (int e = t; e < x.R * x.k; e += 256) {
            const int g = e / x.k, i = e - g * x.k;
            if (i >= s_len[g]) continue;
            const uint8_t* slot = in + (size_t)g * x.slot_bytes;
            const float d = reinterpret_cast<const float*>(slot)[(size_t)qi * x.k + i];
            int pos = i;
            for (int h = 0; h < x.R && pos < x.k; ++h) {
                if (h == g) continue;
                const float* Dh = reinterpret_cast<const float*>(in + (size_t)h * x.slot_bytes) + (size_t)qi * x.k;
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
            if (pos < x.k) xchg_emit(x, qg, pos, d, reinterpret_cast<const long long*>(slot + x.slot_i_off)[(size_t)qi * x.k + i]);
        }
        for (int i = total_valid + t; i < x.k; i += 256) xchg_emit(x, qg, i, l2 ? FLT_MAX : -FLT_MAX, -1);
        __syncthreads();
    }
}

// stage 2 for R = 1 (nothing to merge): my finished list straight into everybody's result buffer
__global__ void __launch_bounds__(256) pq_xchg_push_kernel(const XchgParams x, const float* __restrict__ D_local, const long long* __restrict__ I_local) {
    const long long n = (long long)x.n_loc * x.k;
    const size_t base = (size_t)x.q_base * x.k;
    uint8_t* out = x.peer[blockIdx.y] + x.out_off;
    float* Dd = reinterpret_cast<float*>(out) + base;
    long long* Id = reinterpret_cast<long long*>(out + x.out_i_off) + base;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Dd[i] = D_local[i];
        Id[i] = I_local[i];
    }
}

// stage 3: wait for every rank's slice, then result area -> the caller's tensors
__global__ void __launch_bounds__(256) pq_xchg_collect_kernel(const XchgParams x, float* __restrict__ D_out, long long* __restrict__ I_out) {
    uint8_t* mine = x.peer[x.rank];
    xchg_wait_flags(reinterpret_cast<const uint64_t*>(mine) + 16, x.world, x.seq, reinterpret_cast<uint64_t*>(mine) + 32);
    const long long n = (long long)x.nq * x.k;
    const float* Ds = reinterpret_cast<const float*>(mine + x.out_off);
    const long long* Is = reinterpret_cast<const long long*>(mine + x.out_off + x.out_i_off);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        D_out[i] = Ds[i];
        I_out[i] = Is[i];
    }
}

}  // namespace pq

using namespace pq;

struct pq_xchg {
    int device = 0, world = 1, rank = 0;
    size_t capacity = 0;   // bytes after the header
    DevBuf buf;
    uint8_t* peer[kXchgMaxRanks];
    bool connected = false;
    unsigned long long seq = 0;
};

extern "C" {

int pq_xchg_create(int device, int world, int rank, int64_t payload_bytes, pq_xchg** out, void** base_dev_out, int64_t* bytes_out) {
    if (!out || world < 1 || world > kXchgMaxRanks || rank < 0 || rank >= world || payload_bytes < 0)
        return set_error(PQ_ERR_INVALID, "xchg_create: bad arguments (world=%d, rank=%d)", world, rank);
    *out = nullptr;
    int dev = -1;
    int rc = pick_device(device, &dev);
    if (rc) return rc;
    pq_xchg* x = new (std::nothrow) pq_xchg();
    if (!x) return set_error(PQ_ERR_OOM, "xchg_create: host allocation failed");
    x->device = dev;
    x->world = world;
    x->rank = rank;
    x->capacity = ((size_t)payload_bytes + 255) & ~size_t(255);
    PQ_CUDA(cudaSetDevice(dev));
    {
        // Load every kernel of the exchange NOW: with lazy module loading the first launch of a kernel can wait for the device
        // to go idle — which never happens while another stage of the same exchange spins on a flag that launch would raise.
        cudaFuncAttributes fa;
        PQ_CUDA(cudaFuncGetAttributes(&fa, pq_xchg_scatter_kernel));
        PQ_CUDA(cudaFuncGetAttributes(&fa, pq_xchg_signal_kernel));
        PQ_CUDA(cudaFuncGetAttributes(&fa, pq_xchg_merge_sort_kernel));
        PQ_CUDA(cudaFuncGetAttributes(&fa, pq_xchg_merge_rank_kernel));
        PQ_CUDA(cudaFuncGetAttributes(&fa, pq_xchg_push_kernel));
        PQ_CUDA(cudaFuncGetAttributes(&fa, pq_xchg_collect_kernel));
        PQ_CUDA(cudaFuncSetAttribute(pq_xchg_merge_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    }
    // (at least 2 MB, in 2 MB steps: small cudaMalloc blocks are carved out of shared driver allocations, and two IPC handles that
    // resolve to the same allocation cannot both be opened by a peer — cudaErrorAlreadyMapped)
    x->capacity = ((kXchgHeaderBytes + x->capacity + (2u << 20) - 1) / (2u << 20)) * (2u << 20) - kXchgHeaderBytes;
    rc = x->buf.ensure(kXchgHeaderBytes + x->capacity);
    if (rc) {
        delete x;
        return rc;
    }
    PQ_CUDA(cudaMemset(x->buf.p, 0, kXchgHeaderBytes));
    for (int i = 0; i < kXchgMaxRanks; ++i) x->peer[i] = nullptr;
    *out = x;
    if (base_dev_out) *base_dev_out = x->buf.p;
    if (bytes_out) *bytes_out = (int64_t)(kXchgHeaderBytes + x->capacity);
    return PQ_OK;
}

int pq_xchg_connect(pq_xchg* x, const void* const* peer_bases) {
    if (!x || !peer_bases) return set_error(PQ_ERR_INVALID, "xchg_connect: bad arguments");
    for (int i = 0; i < x->world; ++i) {
        if (!peer_bases[i]) return set_error(PQ_ERR_INVALID, "xchg_connect: buffer %d is null", i);
        x->peer[i] = (uint8_t*)peer_bases[i];
    }
    if (x->peer[x->rank] != (uint8_t*)x->buf.p) return set_error(PQ_ERR_INVALID, "xchg_connect: entry %d must be this rank's own buffer", x->rank);
    x->connected = true;
    return PQ_OK;
}

void pq_xchg_free(pq_xchg* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    x->buf.release();
    delete x;
}

// Bytes the buffers must hold (after the header) for searches of up to nq queries at k results in an R x Q layout.
int64_t pq_xchg_bytes_needed(int64_t nq, int64_t k, int row_shards, int query_groups) {
    if (nq < 1 || k < 1 || row_shards < 1 || query_groups < 1) return 0;
    const int64_t n_loc = (nq + query_groups - 1) / query_groups;
    const int64_t per = (n_loc + row_shards - 1) / row_shards;
    const int64_t slot = (((per * k * 4 + 15) / 16) * 16 + per * k * 8 + 15) / 16 * 16;
    const int64_t in_bytes = (int64_t)row_shards * slot;
    const int64_t out_bytes = (((nq * k * 4 + 15) / 16) * 16 + nq * k * 8 + 255) / 256 * 256;
    return ((in_bytes + 255) / 256) * 256 + 2 * out_bytes + 256;   // two result areas: consecutive searches alternate
}

// One exchange: this rank's list [n_loc, k] (n_loc = its query group's slice of the nq queries; device pointers) in, the full
// result [nq, k] out, everything enqueued on `cuda_stream`.  Every rank of the job calls it with the same nq, k, layout and seq
// (a strictly increasing sequence number per search).  rank = rq * row_shards + rr, as proqa_b200/sharded.py lays the ranks out.
int pq_xchg_run(pq_xchg* x, int metric, int row_shards, int64_t nq, int64_t k, const float* D_local_dev, const int64_t* I_local_dev, float* D_out_dev,
                int64_t* I_out_dev, uint64_t seq, void* cuda_stream) {
    if (!x || !x->connected) return set_error(PQ_ERR_INVALID, "xchg_run: not connected");
    const int R = row_shards, W = x->world;
    if (R < 1 || W % R != 0 || nq < 1 || k < 1 || !D_out_dev || !I_out_dev || seq == 0)
        return set_error(PQ_ERR_INVALID, "xchg_run: bad arguments");
    const int Q = W / R;
    XchgParams p;
    memset(&p, 0, sizeof(p));
    p.world = W;
    p.rank = x->rank;
    p.R = R;
    p.Q = Q;
    p.rr = x->rank % R;
    p.rq = x->rank / R;
    const int64_t perq = (nq + Q - 1) / Q;   // query slice of a group (shard_bounds of sharded.py)
    const int64_t qlo = std::min<int64_t>(nq, perq * p.rq), qhi = std::min<int64_t>(nq, qlo + perq);
    p.n_loc = (int)(qhi - qlo);
    p.q_base = qlo;
    // every rank must agree on the slice geometry: it derives from the largest group slice
    p.per = (int)((perq + R - 1) / R);
    p.nq = (int)nq;
    p.k = (int)k;
    p.metric = metric;
    p.seq = seq;
    p.slot_i_off = ((size_t)p.per * k * 4 + 15) / 16 * 16;
    p.slot_bytes = (p.slot_i_off + (size_t)p.per * k * 8 + 15) / 16 * 16;
    p.in_off = kXchgHeaderBytes;
    const size_t in_bytes = (((size_t)R * p.slot_bytes) + 255) / 256 * 256;
    p.out_i_off = (((size_t)nq * k * 4) + 15) / 16 * 16;
    const size_t out_bytes = (p.out_i_off + (size_t)nq * k * 8 + 255) / 256 * 256;
    // two result areas, alternating with the sequence number: a rank that is already pushing search s+1 (nothing makes it wait for
    // ranks outside its row group) must not write into the area a slower rank is still collecting search s from
    p.out_off = kXchgHeaderBytes + in_bytes + (size_t)(seq & 1ull) * out_bytes;
    if (in_bytes + 2 * out_bytes > x->capacity)
        return set_error(PQ_ERR_INVALID, "xchg_run: buffers too small for nq=%lld k=%lld (pq_xchg_bytes_needed)", (long long)nq, (long long)k);
    if (p.n_loc > 0 && (!D_local_dev || !I_local_dev)) return set_error(PQ_ERR_INVALID, "xchg_run: null list");
    for (int i = 0; i < W; ++i) p.peer[i] = x->peer[i];
    cudaStream_t st = (cudaStream_t)cuda_stream;
    PQ_CUDA(cudaSetDevice(x->device));
    // grids stay well below the GPU (spinning CTAs must never keep a peer's kernels from being scheduled when several ranks share
    // one device, as the single-GPU tests do)
    const long long n_elems = (long long)p.n_loc * k;
    if (R > 1) {
        if (p.n_loc > 0) {
            const int bx = (int)std::max<long long>(1, std::min<long long>(32, ((long long)p.per * k + 1023) / 1024));
            pq_xchg_scatter_kernel<<<dim3(bx, R), 256, 0, st>>>(p, D_local_dev, (const long long*)I_local_dev);
            PQ_CUDA(cudaGetLastError());
        }
        pq_xchg_signal_kernel<<<1, 32, 0, st>>>(p, 1);
        PQ_CUDA(cudaGetLastError());
        const int n_mine = (int)std::max<int64_t>(0, std::min<int64_t>(p.n_loc, (int64_t)(p.rr + 1) * p.per) - (int64_t)p.rr * p.per);
        const int grid = std::max(1, std::min(n_mine, 296));
        int work = 2;
        while (work < R * (int)k) work <<= 1;
        if ((size_t)work * 8 <= 128 * 1024) {
            const size_t smem = (size_t)work * 8;
            pq_xchg_merge_sort_kernel<<<grid, 256, smem, st>>>(p, work);
        } else {
            pq_xchg_merge_rank_kernel<<<grid, 256, 0, st>>>(p);
        }
        PQ_CUDA(cudaGetLastError());
    } else if (p.n_loc > 0) {
        const int bx = (int)std::max<long long>(1, std::min<long long>(16, (n_elems + 2047) / 2048));
        pq_xchg_push_kernel<<<dim3(bx, W), 256, 0, st>>>(p, D_local_dev, (const long long*)I_local_dev);
        PQ_CUDA(cudaGetLastError());
    }
    pq_xchg_signal_kernel<<<1, 32, 0, st>>>(p, 2);
    PQ_CUDA(cudaGetLastError());
    const int gc = (int)std::max<long long>(1, std::min<long long>(128, ((long long)nq * k + 4095) / 4096));
    pq_xchg_collect_kernel<<<gc, 256, 0, st>>>(p, D_out_dev, (long long*)I_out_dev);
    PQ_CUDA(cudaGetLastError());
    x->seq = seq;
    return PQ_OK;
}

// After a synchronisation point: did any wait of this rank time out (a peer never arrived)?
int pq_xchg_check(pq_xchg* x) {
    if (!x) return set_error(PQ_ERR_INVALID, "null exchange");
    PQ_CUDA(cudaSetDevice(x->device));
    unsigned long long err = 0;
    PQ_CUDA(cudaMemcpy(&err, (uint8_t*)x->buf.p + 256, 8, cudaMemcpyDeviceToHost));
    if (err) return set_error(PQ_ERR_CUDA, "list exchange: a peer GPU did not deliver its part within 10 s (rank %d)", x->rank);
    return PQ_OK;
}

}  // extern "C"
