// proqa_b200 — host-side state shared by pq_index.cu (C ABI, fp32 tier) and pq_mma.cu (tensor-core tier).
#pragma once
#include <stdlib.h>

#include <stdint.h>

#include <mutex>
#include <vector>

#include "../../include/proqa_b200.h"
#include "pq_internal.h"

namespace pq {

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* file, int line);

#define PQ_CUDA(expr)                                                    \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) return ::pq::cuda_fail(_e, __FILE__, __LINE__); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);  // grow-only; contents are NOT preserved across growth
    void release();
};

int make_row_tensor_map(CUtensorMap* out, const void* base, long long rows, int elem_bytes, int box_cols, int box_rows);

// Tensor-core tier limits
constexpr int kMmaMaxK = 1024;          // carry list K' <= 4096 entries
constexpr int kMmaMinQueries = 5;       // nq <= 4 on a small corpus: the fp32 scan streams it at the HBM roofline (measured 1.0 of peak) with no fixed cost ...
// ... but the tensor tier streams the bf16 copy, half the bytes, for ~0.3 ms of fixed epoch costs: from about 7M rows on it wins even
// for one query (measured at 21M rows, k = 80 / 1000 / 5000: nq = 1 1.66 / 3.45 / 4.26 ms on the scan, 1.10 / 2.15 / 2.01 ms on the
// tensor tier; nq = 4, k = 5000: 17.0 vs 2.05 ms — tools/gpu_runs/r02_zc_smallnq.sh)
constexpr long long kMmaSmallBatchMinRows = 1LL << 23;
// (the value in force: PROQA_B200_MMA_MIN_QUERIES overrides it, read once)
inline int mma_min_queries() {
    static const int v = [] {
        const char* e = getenv("PROQA_B200_MMA_MIN_QUERIES");
        return (e && atoi(e) >= 1) ? atoi(e) : kMmaMinQueries;
    }();
    return v;
}
constexpr int kMmaMinRows = 16384;      // below this the epoch machinery is pure overhead ...
constexpr long long kMmaMinPairs = 1LL << 24;  // ... unless the query side is large (k-means assignment: 10k centroids x millions of points)

}  // namespace pq

// Cross-shard threshold exchange (corpus row-sharded over several GPUs; pq_mma.cu: ShareParams): this shard's
// mailbox in its own HBM and the device addresses of every shard's mailbox (peer access within a process, CUDA IPC across).
struct pq_share_state {
    int n = 0, rank = 0;
    uint32_t seq = 0;            // search sequence number: every shard of a search passes the same one
    int cap_q = 0;               // queries a mailbox slot holds
    int wait_us = 200;           // longest wait for the peers' announcement of an epoch
    bool connected = false;
    pq::DevBuf mailbox;
    uint64_t* peer[16] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

struct pq_index {
    std::mutex mu;  // serialises the C-ABI calls on THIS index; different indexes (other GPUs, other threads) run concurrently
    int d = 128;
    int metric = 0;
    int tier = 0;
    int requested_device = -1;
    int device = -1;
    bool device_ready = false;
    int n_sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    int64_t ntotal = 0;
    int64_t capacity = 0;
    int64_t id_base = 0;
    float max_norm2 = 0.f;       // max squared row norm (drives the bf16 filter's error bound)
    float max_resid2 = 0.f;      // max |x - bf16(x)|^2 over the rows (same)
    bool has_nonfinite = false;  // some row is inf/NaN (or overflows bf16): tensor-core tier disabled

    // HBM layout of the shard: fp32 rows [cap,128] (512 B/row), bf16 copy [cap,128] (256 B/row), squared norms [cap]
    pq::DevBuf rows_f32, rows_bf16, norms, scalars;
    CUtensorMap tmap_f32, tmap_bf16;

    // large-k tier (1024 < k; PROQA_B200_LARGEK=0 turns it off): compact bf16 copy of every sample_step-th row (+ norms) the thresholds
    // are estimated on; rebuilt lazily after add()/reset() (sample_rows < 0 = stale)
    bool largek = true;
    pq::DevBuf sample_bf16, sample_norms;
    CUtensorMap tmap_sample;
    int64_t sample_rows = -1;
    int sample_step = 0;

    // per-search workspaces (grow-only)
    pq::DevBuf ws_q, ws_D, ws_I, ws_qnorm, ws_qbf16, ws_qbad, ws_qresid;
    pq::DevBuf ws_scan_keys, ws_gthr;
    pq::DevBuf ws_rr_idx, ws_rr_q, ws_rr_qn, ws_rr_D, ws_rr_I;
    pq::DevBuf ws_mma[12];
    pq::DevBuf ws_km[10];  // staged k-means (multi-GPU training): centroids, assignment, sort buffers

    int64_t stats[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    pq_share_state share;

    // optional per-kernel timing of the dominant kernel (fp32 scan / tensor-core filter): event pairs on the stream
    bool profile = false;
    bool stream_is_external = false;
    cudaStream_t own_stream = nullptr;
    std::vector<cudaEvent_t> prof_events;
    std::vector<int> prof_tags;   // what each event pair brackets: 0 dominant kernel, 1 epoch select, 2 threshold fold, 3 rescoring / finalize
    size_t prof_used = 0;
    int prof_begin(int tag = 0) {
        if (!profile) return 0;
        prof_tags.resize(prof_used / 2 + 1);
        prof_tags[prof_used / 2] = tag;
        if (prof_used + 2 > prof_events.size()) {
            for (int i = 0; i < 2; ++i) {
                cudaEvent_t e;
                if (cudaEventCreate(&e) != cudaSuccess) return -1;
                prof_events.push_back(e);
            }
        }
        return cudaEventRecord(prof_events[prof_used], stream) == cudaSuccess ? 0 : -1;
    }
    void prof_end() {
        if (!profile) return;
        cudaEventRecord(prof_events[prof_used + 1], stream);
        prof_used += 2;
    }
    // after the stream drained: total microseconds between the recorded pairs of the dominant kernel (returned; stats[7]);
    // the other tags land in stats[10 + tag]
    int64_t prof_collect() {
        double us[4] = {0.0, 0.0, 0.0, 0.0};
        for (size_t i = 0; i + 1 < prof_used; i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, prof_events[i], prof_events[i + 1]) == cudaSuccess) us[prof_tags[i / 2] & 3] += ms * 1000.0;
        }
        prof_used = 0;
        for (int t = 1; t < 4; ++t) stats[10 + t] = (int64_t)(us[t] + 0.5);
        return (int64_t)(us[0] + 0.5);
    }

    void release_all() {
        pq::DevBuf* all[] = {&rows_f32, &rows_bf16, &norms,    &scalars,  &ws_q,     &ws_D,      &ws_I,      &ws_qnorm, &ws_qbf16,
                             &ws_qbad,  &ws_qresid, &ws_scan_keys, &ws_gthr, &ws_rr_idx, &ws_rr_q, &ws_rr_qn, &ws_rr_D,   &ws_rr_I};
        for (pq::DevBuf* b : all) b->release();
        for (pq::DevBuf& b : ws_mma) b.release();
        for (pq::DevBuf& b : ws_km) b.release();
        sample_bf16.release();
        sample_norms.release();
        sample_rows = -1;
        share.mailbox.release();
        share.connected = false;
    }
};

namespace pq {
// (the *_locked helpers expect the index's own lock, pq_index::mu, held by the C-ABI entry point)
// requested < 0: PROQA_B200_DEVICE / LOCAL_RANK / the current device; validates that the device is an sm_100 part
int pick_device(int requested, int* out);
int index_init_device(pq_index* ix);
int index_add_locked(pq_index* ix, int64_t n, const float* x, bool on_device);
int index_reset_locked(pq_index* ix);
// Device-resident search: dq [nq,128] fp32 -> dD [nq,k], dI [nq,k]; runs on ix->stream and leaves it drained.
int search_device_impl(pq_index* ix, int64_t nq, const float* dq, int64_t k, float* dD, long long* dI);
// fp32 tier: exact scan of all local rows.  Queries, norms and outputs are device pointers.
int search_fp32_scan(pq_index* ix, int nq, const float* dq, const float* dq_norms, int k, float* dD, long long* dI);
// tensor-core tier (pq_mma.cu): bf16 tcgen05 filter + fp32 rescoring.  Requires ws_qbf16 / ws_qnorm / ws_qbad to be
// prepared.  On return dD/dI hold final results for every query whose exactness certificate passed; the indices of
// the others are appended to *rerun (the stream has been synchronised for the flag read-back).
int search_mma_filter(pq_index* ix, int nq, const float* dq, int k, float* dD, long long* dI, std::vector<int>* rerun);
// same contract for 1024 < k <= PQ_MAX_K (pq_mma.cu, "large k"): sample thresholds, one filter pass, finalize kernel
int search_mma_largek(pq_index* ix, int nq, const float* dq, int k, float* dD, long long* dI, std::vector<int>* rerun);
}  // namespace pq
