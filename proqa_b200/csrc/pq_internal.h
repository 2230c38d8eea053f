// proqa_b200 — internal declarations shared by the kernel translation units and the host driver.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pq {

enum Metric : int { kMetricIP = 0, kMetricL2 = 1 };

// ------------------------------------------------------------------------------------------------
// Exact fp32 streaming scan (pq_ffma.cu).  A CTA scans its share of local rows [0, n_rows) for up to 8
// queries held in shared memory and leaves one sorted k-list per (CTA, query).  One launch takes n_batches
// such query batches (CTAs [b * n_ctas, (b + 1) * n_ctas) serve batch b): thousands of queries re-run after a
// failed tensor-tier certificate against a small corpus (k-means: 10,000 centroids) are one launch, not thousands.
// ------------------------------------------------------------------------------------------------
constexpr int kFfmaTileRows = 128;
constexpr int kFfmaThreads = 256;
constexpr int kFfmaMaxQ = 8;
constexpr int kFfmaStageBytes = kFfmaTileRows * 512;

struct FfmaLaunch {
    const CUtensorMap* tmap_rows_f32;  // host copy of the [rows,128] fp32 tensor map (box 32 x 128, SWIZZLE_128B)
    const float* row_norms;            // squared row norms (used by L2 only)
    const float* queries_dev;          // nq_total x 128 fp32, device
    uint64_t* out_keys;                // [n_batches][n_ctas][nq][k]
    uint32_t* gthr;                    // [n_batches][kFfmaMaxQ], ordered-u32 thresholds, pre-initialised by the caller
    long long n_rows;
    int n_ctas;                        // CTAs per query batch
    int nq;                            // queries per batch (<= kFfmaMaxQ)
    int k;
    int metric;
    int n_batches = 1;
    int nq_total = 0;                  // 0: nq (a single batch); the last batch may be short
};
// Returns cudaSuccess or the launch error.  *cap_out receives the per-query buffer capacity used.
cudaError_t ffma_scan_launch(const FfmaLaunch& a, cudaStream_t stream);
// Largest number of queries one scan launch can take for this k (0 if k is unsupported).
int ffma_max_queries_for_k(int k);

// ------------------------------------------------------------------------------------------------
// Selection / merge kernels (pq_select.cu)
// ------------------------------------------------------------------------------------------------
struct MergeLaunch {
    const uint64_t* keys;     // candidate keys
    long long q_stride;       // element stride between consecutive queries
    long long list_stride;    // element stride between consecutive lists of one query
    int n_lists;
    int list_len;             // entries per list (zero keys are skipped)
    const uint32_t* counts;   // optional [q * cnt_q_stride + list]: valid entries per list (clamped to list_len)
    long long cnt_q_stride;
    const uint32_t* gthr;     // optional per-query ordered-u32 lower bound on the key's score half
    int batch_q;              // > 0: queries come in batches of batch_q (the fp32 scan's launch shape): query q = b * batch_q + i reads
    long long batch_stride;   //   keys + b * batch_stride + i * q_stride  and  gthr[b * gthr_batch_stride + i]
    int gthr_batch_stride;    //   (callers zero the struct first)
    int nq;
    int k;
    int metric;
    const float* q_norms;     // squared query norms (L2 only)
    long long id_base;        // added to local row ids
    float* D;                 // [nq][k]
    long long* I;             // [nq][k]
    uint64_t* out_keys;       // optional [nq][k] sorted keys
};
cudaError_t merge_lists_launch(const MergeLaunch& a, cudaStream_t stream);

// Merge G sorted (D, I) lists per query (multi-GPU all-gather result) into one.
cudaError_t merge_di_launch(const float* D_in, const long long* I_in, int n_lists, int nq, int k, int metric, float* D_out,
                            long long* I_out, cudaStream_t stream);

// Row preparation (corpus rows at add(), query rows at search()): squared norms (engine_dot(row,row)),
// bf16 copy, max squared norm, non-finite detection (global flag and/or per-row byte).  Optional outputs may be null
// except rows_bf16 and norms.
// resid2 (optional, per row) / max_resid_bits (optional, running maximum): |x - bf16(x)|^2, rounded up.
cudaError_t prep_rows_launch(const float* rows, long long n, uint16_t* rows_bf16, float* norms, uint32_t* max_norm_bits,
                             uint32_t* nonfinite_flag, uint8_t* row_bad, float* resid2, uint32_t* max_resid_bits, cudaStream_t stream);

}  // namespace pq
