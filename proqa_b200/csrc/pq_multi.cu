// proqa_b200 — one process, several GPUs: the multi-GPU layer behind the drop-in boundary.
//
// The reference calls faiss.IndexFlatIP(d).add(xb) / .search(xq, k) from ONE Python process (retrieval/eval_retrieval.py:102-104,
// retrieval/group_paras.py:36-51); with PROQA_B200_DEVICES=0,1,... the faiss shim hands it a pq_multi instead of a pq_index
// and the unmodified script uses every GPU of the box.  One host thread per device drives that device's shard (every
// pq_index has its own lock and stream); the GPUs talk through peer memory over NVLink — no NCCL, no second process:
//
//   large corpus (first add > kReplicateRows rows): rows sharded contiguously, queries replicated; every shard searches its
//       rows while exchanging thresholds with the others through peer mailboxes (pq_mma.cu: ShareParams), copies its
//       (D, I) list into GPU 0's gather buffer (cudaMemcpyPeerAsync), GPU 0 merges (pq_merge_di kernels)   — north_star (4)
//   small corpus (k-means centroids, group_paras.py:49-51): rows replicated on every GPU, the QUERIES are split — the
//       assignment search of millions of points needs no exchange at all
#include "pq_common.cuh"
#include "pq_host.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace pq {

constexpr int64_t kReplicateRows = 1 << 20;   // a first add() of at most this many rows is replicated (512 MB per GPU)
constexpr int kMultiMaxDevices = 16;

struct IdSegment {
    long long local_begin, global_begin, len;
};

__global__ void pq_multi_map_ids_kernel(long long* __restrict__ I, long long n, const IdSegment* __restrict__ segs, int n_segs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long id = I[i];
    if (id < 0) return;
    int lo = 0, hi = n_segs - 1;   // last segment whose local_begin <= id
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].local_begin <= id) lo = mid;
        else hi = mid - 1;
    }
    I[i] = id - segs[lo].local_begin + segs[lo].global_begin;
}

}  // namespace pq

using namespace pq;

struct pq_multi {
    std::mutex mu;
    int d = 128, metric = 0, n = 0;
    int devices[kMultiMaxDevices];
    pq_index* shard[kMultiMaxDevices];
    std::vector<IdSegment> segs[kMultiMaxDevices];
    DevBuf segs_dev[kMultiMaxDevices];
    bool segs_dirty[kMultiMaxDevices];
    int64_t ntotal = 0;
    int mode = 0;              // 0 undecided, 1 rows sharded, 2 rows replicated
    bool peers_enabled = false;
    int share_cap = 0;
    int64_t bounds_ntotal = -1;   // ntotal when the shards last agreed on the error-bound scalars
    bool share_off = false;
    uint32_t seq = 0;
    DevBuf gD, gI, oD, oI;     // on devices[0]: gathered shard lists [n][nq][k], merged result
    cudaStream_t stream0 = nullptr;
    int64_t stats[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

struct WorkerResult {
    int rc = PQ_OK;
    std::string msg;
};

// Runs fn(g) on one host thread per device; the first failure (lowest device) is reported through the caller's error slot.
template <typename F>
int for_each_device(pq_multi* m, F fn) {
    std::vector<WorkerResult> res((size_t)m->n);
    std::vector<std::thread> th;
    for (int g = 1; g < m->n; ++g)
        th.emplace_back([&, g] {
            res[g].rc = fn(g);
            if (res[g].rc) res[g].msg = pq_last_error();
        });
    res[0].rc = fn(0);
    if (res[0].rc) res[0].msg = pq_last_error();
    for (std::thread& t : th) t.join();
    for (int g = 0; g < m->n; ++g)
        if (res[g].rc) return set_error(res[g].rc, "device %d: %s", m->devices[g], res[g].msg.c_str());
    return PQ_OK;
}

int enable_all_peers(pq_multi* m) {
    if (m->peers_enabled) return PQ_OK;
    for (int a = 0; a < m->n; ++a)
        for (int b = 0; b < m->n; ++b)
            if (a != b) {
                const int rc = pq_enable_peer_access(m->devices[a], m->devices[b]);
                if (rc) return rc;
            }
    m->peers_enabled = true;
    return PQ_OK;
}

// Mailboxes for the threshold exchange (row-sharded mode): every shard's mailbox address is valid on every device of this
// process once peer access is on.  Also gives every shard the error-bound scalars of the whole corpus.
int ensure_share(pq_multi* m, int64_t nq) {
    if (m->share_off || m->n < 2) return PQ_OK;
    if (m->bounds_ntotal != m->ntotal) {
        float mx[2] = {0.f, 0.f};
        for (int g = 0; g < m->n; ++g) {
            float s[2];
            int rc = pq_index_get_bound_scalars(m->shard[g], s);
            if (rc) return rc;
            mx[0] = std::max(mx[0], s[0]);
            mx[1] = std::max(mx[1], s[1]);
        }
        for (int g = 0; g < m->n; ++g) {
            const int rc = pq_index_set_bound_scalars(m->shard[g], mx[0], mx[1]);
            if (rc) return rc;
        }
        m->bounds_ntotal = m->ntotal;
    }
    const int64_t want = std::min<int64_t>(nq, 1 << 18);
    if (want <= m->share_cap) return PQ_OK;
    int cap = 1024;
    while (cap < want) cap *= 2;
    void* boxes[kMultiMaxDevices];
    for (int g = 0; g < m->n; ++g) {
        const int rc = pq_index_share_alloc(m->shard[g], m->n, g, cap, &boxes[g], nullptr);
        if (rc) return rc;
    }
    for (int g = 0; g < m->n; ++g) {
        const int rc = pq_index_share_connect(m->shard[g], boxes);
        if (rc) return rc;
    }
    m->share_cap = cap;
    return PQ_OK;
}

}  // namespace

extern "C" {

int pq_multi_create(int d, int metric, int n_devices, const int* devices, pq_multi** out) {
    if (!out) return set_error(PQ_ERR_INVALID, "multi_create: null out pointer");
    *out = nullptr;
    if (n_devices < 1 || n_devices > kMultiMaxDevices || !devices) return set_error(PQ_ERR_INVALID, "multi_create: 1..%d devices", kMultiMaxDevices);
    pq_multi* m = new (std::nothrow) pq_multi();
    if (!m) return set_error(PQ_ERR_OOM, "multi_create: host allocation failed");
    m->d = d;
    m->metric = metric;
    m->n = n_devices;
    const char* so = getenv("PROQA_B200_SHARE");
    m->share_off = so && !strcmp(so, "0");
    for (int g = 0; g < n_devices; ++g) {
        m->devices[g] = devices[g];
        m->shard[g] = nullptr;
        m->segs_dirty[g] = false;
    }
    for (int g = 0; g < n_devices; ++g) {   // (no CUDA call yet: the reference forks after importing faiss)
        const int rc = pq_index_create(d, metric, devices[g], &m->shard[g]);
        if (rc) {
            for (int h = 0; h < g; ++h) pq_index_free(m->shard[h]);
            delete m;
            return rc;
        }
    }
    *out = m;
    return PQ_OK;
}

void pq_multi_free(pq_multi* m) {
    if (!m) return;
    for (int g = 0; g < m->n; ++g) {
        if (m->segs_dev[g].p) {
            cudaSetDevice(m->devices[g]);
            m->segs_dev[g].release();
        }
        pq_index_free(m->shard[g]);
    }
    if (m->gD.p || m->gI.p || m->oD.p || m->oI.p || m->stream0) {
        cudaSetDevice(m->devices[0]);
        m->gD.release();
        m->gI.release();
        m->oD.release();
        m->oI.release();
        if (m->stream0) cudaStreamDestroy(m->stream0);
    }
    delete m;
}

int64_t pq_multi_ntotal(const pq_multi* m) { return m ? m->ntotal : 0; }
int pq_multi_n_devices(const pq_multi* m) { return m ? m->n : 0; }
int pq_multi_mode(const pq_multi* m) { return m ? m->mode : 0; }

int pq_multi_reset(pq_multi* m) {
    if (!m) return set_error(PQ_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(m->mu);
    for (int g = 0; g < m->n; ++g) {
        const int rc = pq_index_reset(m->shard[g]);
        if (rc) return rc;
        m->segs[g].clear();
        m->segs_dirty[g] = false;
        pq_index_set_id_base(m->shard[g], 0);
    }
    m->ntotal = 0;
    m->mode = 0;
    m->bounds_ntotal = -1;
    return PQ_OK;
}

// index.add(xb): ids are insertion order.  Row-sharded mode cuts every add into n contiguous slices (shard g takes the g-th).
int pq_multi_add(pq_multi* m, int64_t n, const float* x_host) {
    if (!m) return set_error(PQ_ERR_INVALID, "null index");
    if (n < 0 || (n > 0 && !x_host)) return set_error(PQ_ERR_INVALID, "add: bad arguments");
    if (n == 0) return PQ_OK;
    std::lock_guard<std::mutex> lock(m->mu);
    if (m->mode == 0) m->mode = (m->n > 1 && n > kReplicateRows) ? 1 : 2;
    if (m->mode == 2 || m->n == 1) {
        const int rc = for_each_device(m, [&](int g) { return pq_index_add(m->shard[g], n, x_host); });
        if (rc) return rc;
        m->ntotal += n;
        return PQ_OK;
    }
    const int64_t per = (n + m->n - 1) / m->n;
    const int rc = for_each_device(m, [&](int g) {
        const int64_t lo = std::min(n, per * g), hi = std::min(n, lo + per);
        if (hi <= lo) return (int)PQ_OK;
        IdSegment s;
        s.local_begin = pq_index_ntotal(m->shard[g]);
        s.global_begin = m->ntotal + lo;
        s.len = hi - lo;
        const int r = pq_index_add(m->shard[g], hi - lo, x_host + (size_t)lo * kDim);
        if (r) return r;
        m->segs[g].push_back(s);
        m->segs_dirty[g] = true;
        // one segment: ids are local row + constant; more: the search maps them through the segment table
        pq_index_set_id_base(m->shard[g], m->segs[g].size() == 1 ? s.global_begin - s.local_begin : 0);
        return (int)PQ_OK;
    });
    if (rc) return rc;
    m->ntotal += n;
    return PQ_OK;
}

int pq_multi_search(pq_multi* m, int64_t nq, const float* xq, int64_t k, float* D, int64_t* I) {
    if (!m) return set_error(PQ_ERR_INVALID, "null index");
    if (nq < 0 || k < 1 || (nq > 0 && (!xq || !D || !I))) return set_error(PQ_ERR_INVALID, "search: bad arguments");
    if (nq == 0) return PQ_OK;
    std::lock_guard<std::mutex> lock(m->mu);
    memset(m->stats, 0, sizeof(m->stats));
    if (m->n == 1 || m->mode == 0) {
        const int rc = pq_index_search(m->shard[0], nq, xq, k, D, I);
        pq_index_last_stats(m->shard[0], m->stats, 10);
        return rc;
    }
    if (m->mode == 2) {   // replicated rows: split the queries, every device writes its slice of the caller's buffers
        const int64_t per = (nq + m->n - 1) / m->n;
        const int rc = for_each_device(m, [&](int g) {
            const int64_t lo = std::min(nq, per * g), hi = std::min(nq, lo + per);
            if (hi <= lo) return (int)PQ_OK;
            return pq_index_search(m->shard[g], hi - lo, xq + (size_t)lo * kDim, k, D + (size_t)lo * k, I + (size_t)lo * k);
        });
        for (int g = 0; g < m->n; ++g) {
            int64_t s[10];
            pq_index_last_stats(m->shard[g], s, 10);
            for (int i = 0; i < 10; ++i) m->stats[i] = (i == 6 || i == 7) ? std::max(m->stats[i], s[i]) : m->stats[i] + s[i];
        }
        return rc;
    }
    // ---- rows sharded ----
    if (k > PQ_MAX_K) return set_error(PQ_ERR_UNSUPPORTED, "search: k=%lld exceeds PQ_MAX_K=%d", (long long)k, PQ_MAX_K);
    int rc = enable_all_peers(m);
    if (!rc) rc = ensure_share(m, nq);
    if (rc) return rc;
    const int dev0 = m->devices[0];
    PQ_CUDA(cudaSetDevice(dev0));
    if (!m->stream0) PQ_CUDA(cudaStreamCreateWithFlags(&m->stream0, cudaStreamNonBlocking));
    const size_t list = (size_t)nq * k;
    rc = m->gD.ensure(list * m->n * 4);
    if (!rc) rc = m->gI.ensure(list * m->n * 8);
    if (!rc) rc = m->oD.ensure(list * 4);
    if (!rc) rc = m->oI.ensure(list * 8);
    if (rc) return rc;
    m->seq += 1;
    rc = for_each_device(m, [&](int g) {
        pq_index* ix = m->shard[g];
        std::lock_guard<std::mutex> il(ix->mu);
        int r = index_init_device(ix);
        if (r) return r;
        PQ_CUDA(cudaSetDevice(ix->device));
        ix->share.seq = (m->seq % 0x0ffffffeu) + 1u;
        r = ix->ws_q.ensure((size_t)nq * kDim * 4);
        if (!r) r = ix->ws_D.ensure(list * 4);
        if (!r) r = ix->ws_I.ensure(list * 8);
        if (r) return r;
        PQ_CUDA(cudaMemcpyAsync(ix->ws_q.p, xq, (size_t)nq * kDim * 4, cudaMemcpyHostToDevice, ix->stream));
        r = search_device_impl(ix, nq, (const float*)ix->ws_q.p, k, (float*)ix->ws_D.p, (long long*)ix->ws_I.p);
        if (r) return r;
        if (m->segs[g].size() > 1) {
            if (m->segs_dirty[g]) {
                r = m->segs_dev[g].ensure(m->segs[g].size() * sizeof(IdSegment));
                if (r) return r;
                PQ_CUDA(cudaMemcpyAsync(m->segs_dev[g].p, m->segs[g].data(), m->segs[g].size() * sizeof(IdSegment), cudaMemcpyHostToDevice, ix->stream));
                m->segs_dirty[g] = false;
            }
            pq_multi_map_ids_kernel<<<(unsigned)((list + 255) / 256), 256, 0, ix->stream>>>((long long*)ix->ws_I.p, (long long)list,
                                                                                            (const IdSegment*)m->segs_dev[g].p, (int)m->segs[g].size());
            PQ_CUDA(cudaGetLastError());
        }
        // this shard's list into GPU 0's gather buffer, over NVLink
        PQ_CUDA(cudaMemcpyPeerAsync((float*)m->gD.p + list * g, dev0, ix->ws_D.p, ix->device, list * 4, ix->stream));
        PQ_CUDA(cudaMemcpyPeerAsync((long long*)m->gI.p + list * g, dev0, ix->ws_I.p, ix->device, list * 8, ix->stream));
        PQ_CUDA(cudaStreamSynchronize(ix->stream));
        return (int)PQ_OK;
    });
    for (int g = 0; g < m->n; ++g)
        for (int i = 0; i < 10; ++i) m->stats[i] = (i == 6 || i == 7) ? std::max(m->stats[i], m->shard[g]->stats[i]) : m->stats[i] + m->shard[g]->stats[i];
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(dev0));
    rc = pq_merge_shard_results_async(dev0, m->metric, m->n, nq, k, (const float*)m->gD.p, (const int64_t*)m->gI.p, (float*)m->oD.p, (int64_t*)m->oI.p,
                                      m->stream0);
    if (rc) return rc;
    PQ_CUDA(cudaMemcpyAsync(D, m->oD.p, list * 4, cudaMemcpyDeviceToHost, m->stream0));
    PQ_CUDA(cudaMemcpyAsync(I, m->oI.p, list * 8, cudaMemcpyDeviceToHost, m->stream0));
    PQ_CUDA(cudaStreamSynchronize(m->stream0));
    m->stats[5] += 1;
    return PQ_OK;
}

int pq_multi_last_stats(const pq_multi* m, int64_t* out, int n) {
    if (!m || !out || n < 0) return set_error(PQ_ERR_INVALID, "last_stats: bad arguments");
    for (int i = 0; i < n; ++i) out[i] = i < 10 ? m->stats[i] : 0;
    return PQ_OK;
}

// The shard on the first device: faiss.Clustering.train(x, index) trains on one GPU and leaves the centroids in every replica
// through the reset()/add() the caller does next (group_paras.py:49-50).
pq_index* pq_multi_first_shard(pq_multi* m) { return m ? m->shard[0] : nullptr; }

}  // extern "C"
