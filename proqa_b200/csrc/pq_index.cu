// proqa_b200 — host driver and C ABI (include/proqa_b200.h).
//
// Mirrors the slice of the FAISS API that ProQA's retrieval path calls (IndexFlatIP/IndexFlatL2:
// add, search, reset, ntotal — retrieval/eval_retrieval.py:102-104, retrieval/group_paras.py:35-51,
// retrieval/trec_process.py:74-76) on top of the sm_100a kernels.  There is no CPU path: without an
// sm_100 device every call that needs one fails with PQ_ERR_NO_DEVICE.
#include "pq_common.cuh"
#include "pq_host.h"
#include "pq_plan.h"

#include <cuda_fp16.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace pq {

static thread_local char g_err[1024] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* file, int line) {
    const int code = (e == cudaErrorMemoryAllocation) ? PQ_ERR_OOM : PQ_ERR_CUDA;
    cudaGetLastError();  // clear the sticky-less error state
    return set_error(code, "CUDA error %s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
}

static int cuda_ok_or_fail(cudaError_t e) { return e == cudaSuccess ? PQ_OK : cuda_fail(e, __FILE__, __LINE__); }

// Locking: every index has its own lock (pq_index::mu) held across a C-ABI call, so one host thread per GPU can drive the
// shards of a row-sharded corpus concurrently.  Process-wide state is small and has its own locks: the pinned staging
// buffers (one pair per device), the per-device table of validated ordinals, the shared-memory attribute cache (pq_mma.cu).

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, 128] row-major matrix, box = box_cols x box_rows elements, 128-byte swizzle, zero fill out of bounds.
int make_row_tensor_map(CUtensorMap* out, const void* base, long long rows, int elem_bytes, int box_cols, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(PQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available from the driver");
    if (rows < 1) rows = 1;
    const cuuint64_t gdim[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)kDim * elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PQ_OK;
}

// ------------------------------------------------------------------------------------------------
// device buffers
// ------------------------------------------------------------------------------------------------
int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return PQ_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = (bytes + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        return set_error(PQ_ERR_OOM, "device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return PQ_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

int pick_device(int requested, int* out) {
    // cudaGetDeviceProperties costs milliseconds: validate each ordinal once per process
    static std::mutex mu;
    static int validated[64];
    std::lock_guard<std::mutex> lock(mu);
    if (requested >= 0 && requested < 64 && validated[requested]) {
        *out = requested;
        return PQ_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(PQ_ERR_NO_DEVICE, "no CUDA device visible (%s); proqa_b200 has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    int dev = requested;
    if (dev < 0) {
        const char* s = getenv("PROQA_B200_DEVICE");
        if (!s || !*s) s = getenv("LOCAL_RANK");
        if (s && *s) {
            dev = atoi(s) % n;
        } else if (cudaGetDevice(&dev) != cudaSuccess) {
            dev = 0;
        }
    }
    if (dev >= n) return set_error(PQ_ERR_INVALID, "device %d requested but only %d visible", dev, n);
    cudaDeviceProp prop;
    PQ_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return set_error(PQ_ERR_NO_DEVICE, "device %d is sm_%d%d; proqa_b200 only runs on sm_100 (B200) and has no fallback", dev,
                         prop.major, prop.minor);
    if (dev < 64) validated[dev] = 1;
    *out = dev;
    return PQ_OK;
}

// ------------------------------------------------------------------------------------------------
// index
// ------------------------------------------------------------------------------------------------
int index_init_device(pq_index* ix) {
    if (ix->device_ready) return PQ_OK;
    int dev = -1;
    int rc = pick_device(ix->requested_device, &dev);
    if (rc) return rc;
    ix->device = dev;
    PQ_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    PQ_CUDA(cudaGetDeviceProperties(&prop, dev));
    ix->n_sms = prop.multiProcessorCount;
    PQ_CUDA(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
    if (!ix->stream_is_external) ix->stream = ix->own_stream;
    PQ_CUDA(cudaEventCreate(&ix->ev0));
    PQ_CUDA(cudaEventCreate(&ix->ev1));
    rc = ix->scalars.ensure(256);
    if (rc) return rc;
    PQ_CUDA(cudaMemsetAsync(ix->scalars.p, 0, 256, ix->stream));
    ix->device_ready = true;
    return PQ_OK;
}

static int index_refresh_maps(pq_index* ix) {
    ix->sample_rows = -1;  // the rows changed: the large-k threshold sample is stale
    if (ix->ntotal == 0) return PQ_OK;
    int rc = make_row_tensor_map(&ix->tmap_f32, ix->rows_f32.p, ix->ntotal, 4, 32, kFfmaTileRows);
    if (rc) return rc;
    return make_row_tensor_map(&ix->tmap_bf16, ix->rows_bf16.p, ix->ntotal, 2, 64, 128);
}

static int index_grow(pq_index* ix, int64_t need_rows) {
    if (need_rows <= ix->capacity) return PQ_OK;
    int64_t newcap = need_rows;
    if (ix->capacity > 0) newcap = std::max<int64_t>(need_rows, ix->capacity + ix->capacity / 2);
    newcap = (newcap + 255) / 256 * 256;
    DevBuf nf, nb, nn;
    int rc = nf.ensure((size_t)newcap * kDim * 4);
    if (!rc) rc = nb.ensure((size_t)newcap * kDim * 2);
    if (!rc) rc = nn.ensure((size_t)newcap * 4);
    if (rc) {
        nf.release();
        nb.release();
        nn.release();
        return rc;
    }
    if (ix->ntotal > 0) {
        cudaError_t e = cudaMemcpyAsync(nf.p, ix->rows_f32.p, (size_t)ix->ntotal * kDim * 4, cudaMemcpyDeviceToDevice, ix->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nb.p, ix->rows_bf16.p, (size_t)ix->ntotal * kDim * 2, cudaMemcpyDeviceToDevice, ix->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nn.p, ix->norms.p, (size_t)ix->ntotal * 4, cudaMemcpyDeviceToDevice, ix->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
        if (e != cudaSuccess) {  // the new buffers (1.5x the index) must not outlive a failed copy
            nf.release();
            nb.release();
            nn.release();
            return cuda_fail(e, __FILE__, __LINE__);
        }
    }
    ix->rows_f32.release();
    ix->rows_bf16.release();
    ix->norms.release();
    ix->rows_f32 = nf;
    ix->rows_bf16 = nb;
    ix->norms = nn;
    ix->capacity = newcap;
    return PQ_OK;
}

// ------------------------------------------------------------------------------------------------
// Index load path: host rows -> device.  The reference hands add() a pageable numpy array
// (eval_retrieval.py:100,103: np.load(...).astype('float32'), 10.75 GB for the 21M-row index).  A plain cudaMemcpy from
// pageable memory is staged by the driver through one small bounce buffer; here the copy is pipelined explicitly:
// host threads fill one of two pinned 64 MB buffers while the other is in flight over PCIe and the row-preparation kernel
// (norms, bf16 copy) of the previous chunk runs on the same stream.
// ------------------------------------------------------------------------------------------------
struct StagingBuffers {
    static constexpr size_t kBytes = 64u << 20;
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool ok = false;
    bool init() {
        if (ok) return true;
        for (int i = 0; i < 2; ++i) {
            if (cudaMallocHost(&buf[i], kBytes) != cudaSuccess || cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
        }
        ok = true;
        return true;
    }
};
// One pair per device (the events belong to a device); held for the length of one staged add() through its lock.
static StagingBuffers g_staging_of[64];
static std::mutex g_staging_mu[64];

static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const int n_threads = 4;
    const size_t per = (bytes / n_threads + 4095) & ~size_t(4095);
    std::vector<std::thread> th;
    for (int i = 1; i < n_threads; ++i) {
        const size_t a = std::min(bytes, per * i), b = std::min(bytes, per * (i + 1));
        if (b > a) th.emplace_back([=] { memcpy((char*)dst + a, (const char*)src + a, b - a); });
    }
    memcpy(dst, src, std::min(bytes, per));
    for (std::thread& t : th) t.join();
}

// Copies n rows from pageable host memory to dst_dev and launches the row preparation chunk by chunk on `stream`.
static int staged_add_rows(pq_index* ix, float* dst_dev, const float* x_host, int64_t n, int64_t first_row) {
    uint32_t* sc = (uint32_t*)ix->scalars.p;
    const int64_t rows_per_chunk = (int64_t)(StagingBuffers::kBytes / (kDim * 4));
    std::unique_lock<std::mutex> staging_lock(g_staging_mu[ix->device & 63], std::defer_lock);
    StagingBuffers& g_staging = g_staging_of[ix->device & 63];
    const bool small = n * kDim * 4 < (int64_t)(8u << 20);
    if (!small) staging_lock.lock();
    if (small || !g_staging.init()) {  // small (k-means centroids, tests): one plain copy
        PQ_CUDA(cudaMemcpyAsync(dst_dev, x_host, (size_t)n * kDim * 4, cudaMemcpyHostToDevice, ix->stream));
        return cuda_ok_or_fail(prep_rows_launch(dst_dev, n, (uint16_t*)ix->rows_bf16.p + (size_t)first_row * kDim, (float*)ix->norms.p + first_row,
                                                sc + 0, sc + 1, nullptr, nullptr, sc + 2, ix->stream));
    }
    int which = 0;
    for (int64_t a = 0; a < n; a += rows_per_chunk, which ^= 1) {
        const int64_t rows = std::min(rows_per_chunk, n - a);
        PQ_CUDA(cudaEventSynchronize(g_staging.done[which]));  // the H2D that last used this buffer has finished
        parallel_memcpy(g_staging.buf[which], x_host + (size_t)a * kDim, (size_t)rows * kDim * 4);
        float* d = dst_dev + (size_t)a * kDim;
        PQ_CUDA(cudaMemcpyAsync(d, g_staging.buf[which], (size_t)rows * kDim * 4, cudaMemcpyHostToDevice, ix->stream));
        PQ_CUDA(cudaEventRecord(g_staging.done[which], ix->stream));
        PQ_CUDA(prep_rows_launch(d, rows, (uint16_t*)ix->rows_bf16.p + (size_t)(first_row + a) * kDim, (float*)ix->norms.p + first_row + a, sc + 0,
                                 sc + 1, nullptr, nullptr, sc + 2, ix->stream));
    }
    PQ_CUDA(cudaStreamSynchronize(ix->stream));  // the pinned buffers are free again before the lock is
    return PQ_OK;
}

// Large query matrices and result arrays of a host-buffer search (group_paras.py:51 hands index.search all 21M points: 10.75 GB of
// pageable memory; C5 returns 98 MB) go through the same pinned double buffer: a plain cudaMemcpy from / to pageable memory
// runs at 7-9 GB/s (measured: 2M points in 0.116 s, C5 results 14.5 ms), the staged path at what add() gets (29 GB/s).
constexpr size_t kStagedCopyMinBytes = 8u << 20;
static int staged_upload(pq_index* ix, void* dst_dev, const void* src_host, size_t bytes) {
    StagingBuffers& sb = g_staging_of[ix->device & 63];
    std::unique_lock<std::mutex> staging_lock(g_staging_mu[ix->device & 63], std::defer_lock);
    if (bytes >= kStagedCopyMinBytes) staging_lock.lock();
    if (bytes < kStagedCopyMinBytes || !sb.init()) {
        PQ_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ix->stream));
        return PQ_OK;
    }
    int which = 0;
    for (size_t a = 0; a < bytes; a += StagingBuffers::kBytes, which ^= 1) {
        const size_t n = std::min(StagingBuffers::kBytes, bytes - a);
        PQ_CUDA(cudaEventSynchronize(sb.done[which]));  // the H2D that last used this buffer has finished
        parallel_memcpy(sb.buf[which], (const char*)src_host + a, n);
        PQ_CUDA(cudaMemcpyAsync((char*)dst_dev + a, sb.buf[which], n, cudaMemcpyHostToDevice, ix->stream));
        PQ_CUDA(cudaEventRecord(sb.done[which], ix->stream));
    }
    PQ_CUDA(cudaStreamSynchronize(ix->stream));  // the pinned buffers are free again before the lock is
    return PQ_OK;
}
// Device -> pageable host memory; returns with the copy complete.
static int staged_download(pq_index* ix, void* dst_host, const void* src_dev, size_t bytes) {
    StagingBuffers& sb = g_staging_of[ix->device & 63];
    std::unique_lock<std::mutex> staging_lock(g_staging_mu[ix->device & 63], std::defer_lock);
    if (bytes >= kStagedCopyMinBytes) staging_lock.lock();
    if (bytes < kStagedCopyMinBytes || !sb.init()) {
        PQ_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ix->stream));
        PQ_CUDA(cudaStreamSynchronize(ix->stream));
        return PQ_OK;
    }
    const size_t chunk = StagingBuffers::kBytes, n_chunks = (bytes + chunk - 1) / chunk;
    for (size_t i = 0; i <= n_chunks; ++i) {  // chunk i flies into one buffer while the host copies chunk i - 1 out of the other
        if (i < n_chunks) {
            PQ_CUDA(cudaMemcpyAsync(sb.buf[i & 1], (const char*)src_dev + i * chunk, std::min(chunk, bytes - i * chunk), cudaMemcpyDeviceToHost,
                                    ix->stream));
            PQ_CUDA(cudaEventRecord(sb.done[i & 1], ix->stream));
        }
        if (i >= 1) {
            PQ_CUDA(cudaEventSynchronize(sb.done[(i - 1) & 1]));
            parallel_memcpy((char*)dst_host + (i - 1) * chunk, sb.buf[(i - 1) & 1], std::min(chunk, bytes - (i - 1) * chunk));
        }
    }
    return PQ_OK;
}

// fp16 rows (what get_embed.py --fp16 writes, get_embed.py:147-151) -> fp32 rows: exact, and half the bytes over PCIe.
__global__ void pq_half_to_float_kernel(const __half* __restrict__ in, float* __restrict__ out, long long n_elems) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i + 8 <= n_elems) {
        const uint4 raw = *reinterpret_cast<const uint4*>(in + i);
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
        float4 a, b;
        const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
        a = make_float4(f0.x, f0.y, f1.x, f1.y);
        b = make_float4(f2.x, f2.y, f3.x, f3.y);
        *reinterpret_cast<float4*>(out + i) = a;
        *reinterpret_cast<float4*>(out + i + 4) = b;
    }
}

// Caller holds the index's lock.  x_host: n rows of 128 IEEE half values.
int index_add_f16_locked(pq_index* ix, int64_t n, const void* x_host) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    if (n < 0 || (n > 0 && !x_host)) return set_error(PQ_ERR_INVALID, "add_f16: bad arguments");
    if (n == 0) return PQ_OK;
    if (ix->ntotal + n > (int64_t)0x7fffff00) return set_error(PQ_ERR_UNSUPPORTED, "add: more than 2^31 rows per shard");
    int rc = index_init_device(ix);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(ix->device));
    rc = index_grow(ix, ix->ntotal + n);
    if (rc) return rc;
    std::lock_guard<std::mutex> staging_lock(g_staging_mu[ix->device & 63]);  // (released after the final stream synchronise below)
    StagingBuffers& g_staging = g_staging_of[ix->device & 63];
    if (!g_staging.init()) return set_error(PQ_ERR_OOM, "add_f16: pinned staging buffers unavailable");
    DevBuf tmp[2];
    const int64_t rows_per_chunk = (int64_t)(StagingBuffers::kBytes / (kDim * 2));
    rc = tmp[0].ensure(StagingBuffers::kBytes);
    if (!rc) rc = tmp[1].ensure(StagingBuffers::kBytes);
    if (rc) {
        tmp[0].release();
        tmp[1].release();
        return rc;
    }
    uint32_t* sc = (uint32_t*)ix->scalars.p;
    const uint16_t* src = (const uint16_t*)x_host;
    int which = 0;
    cudaError_t e = cudaSuccess;
    for (int64_t a = 0; a < n && e == cudaSuccess; a += rows_per_chunk, which ^= 1) {
        const int64_t rows = std::min(rows_per_chunk, n - a);
        e = cudaEventSynchronize(g_staging.done[which]);
        if (e != cudaSuccess) break;
        parallel_memcpy(g_staging.buf[which], src + (size_t)a * kDim, (size_t)rows * kDim * 2);
        e = cudaMemcpyAsync(tmp[which].p, g_staging.buf[which], (size_t)rows * kDim * 2, cudaMemcpyHostToDevice, ix->stream);
        if (e != cudaSuccess) break;
        e = cudaEventRecord(g_staging.done[which], ix->stream);
        if (e != cudaSuccess) break;
        float* d = (float*)ix->rows_f32.p + (size_t)(ix->ntotal + a) * kDim;
        const long long elems = (long long)rows * kDim;
        pq_half_to_float_kernel<<<(unsigned)((elems / 8 + 255) / 256), 256, 0, ix->stream>>>((const __half*)tmp[which].p, d, elems);
        e = cudaGetLastError();
        if (e != cudaSuccess) break;
        e = prep_rows_launch(d, rows, (uint16_t*)ix->rows_bf16.p + (size_t)(ix->ntotal + a) * kDim, (float*)ix->norms.p + ix->ntotal + a, sc + 0, sc + 1,
                             nullptr, nullptr, sc + 2, ix->stream);
    }
    uint32_t host_sc[3] = {0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_sc, sc, 12, cudaMemcpyDeviceToHost, ix->stream);
    const cudaError_t e2 = cudaStreamSynchronize(ix->stream);
    tmp[0].release();
    tmp[1].release();
    if (e != cudaSuccess) return cuda_fail(e, __FILE__, __LINE__);
    if (e2 != cudaSuccess) return cuda_fail(e2, __FILE__, __LINE__);
    memcpy(&ix->max_norm2, &host_sc[0], 4);
    memcpy(&ix->max_resid2, &host_sc[2], 4);
    ix->has_nonfinite = host_sc[1] != 0;
    ix->ntotal += n;
    return index_refresh_maps(ix);
}

// Caller holds the index's lock.
int index_add_locked(pq_index* ix, int64_t n, const float* x, bool on_device) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    if (n < 0 || (n > 0 && !x)) return set_error(PQ_ERR_INVALID, "add: bad arguments (n=%lld, x=%p)", (long long)n, (const void*)x);
    if (n == 0) return PQ_OK;
    if (ix->ntotal + n > (int64_t)0x7fffff00) return set_error(PQ_ERR_UNSUPPORTED, "add: more than 2^31 rows per shard");
    int rc = index_init_device(ix);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(ix->device));
    rc = index_grow(ix, ix->ntotal + n);
    if (rc) return rc;
    float* dst = (float*)ix->rows_f32.p + (size_t)ix->ntotal * kDim;
    uint32_t* sc = (uint32_t*)ix->scalars.p;
    if (on_device) {
        PQ_CUDA(cudaMemcpyAsync(dst, x, (size_t)n * kDim * 4, cudaMemcpyDeviceToDevice, ix->stream));
        PQ_CUDA(prep_rows_launch(dst, n, (uint16_t*)ix->rows_bf16.p + (size_t)ix->ntotal * kDim, (float*)ix->norms.p + ix->ntotal, sc + 0,
                                 sc + 1, nullptr, nullptr, sc + 2, ix->stream));
    } else {
        rc = staged_add_rows(ix, dst, x, n, ix->ntotal);
        if (rc) return rc;
    }
    uint32_t host_sc[3];
    PQ_CUDA(cudaMemcpyAsync(host_sc, sc, 12, cudaMemcpyDeviceToHost, ix->stream));
    PQ_CUDA(cudaStreamSynchronize(ix->stream));
    memcpy(&ix->max_norm2, &host_sc[0], 4);
    memcpy(&ix->max_resid2, &host_sc[2], 4);
    ix->has_nonfinite = host_sc[1] != 0;
    ix->ntotal += n;
    return index_refresh_maps(ix);
}

// Exact fp32 scan over all local rows for queries [0, nq) already on the device.
int search_fp32_scan(pq_index* ix, int nq, const float* dq, const float* dq_norms, int k, float* dD, long long* dI) {
    const int qmax = ffma_max_queries_for_k(k);
    if (qmax < 1) return set_error(PQ_ERR_UNSUPPORTED, "k=%d is above what the fp32 scan supports", k);
    const int n_tiles = (int)((ix->ntotal + kFfmaTileRows - 1) / kFfmaTileRows);
    // Query batches of qmax (<= 8) share one launch: CTAs [b * n_ctas, (b + 1) * n_ctas) scan the rows for batch b.  With few
    // batches each takes every SM it can use; with many (k-means: thousands of points whose k = 1 certificate failed, against
    // 10,000 centroids) a batch gets fewer CTAs, each scanning more of the rows — about two waves of CTAs in all.  (Until round 2
    // every batch was three stream operations of its own: 170k re-run points cost 0.5 s of launches.)
    const int kMaxBatchesPerLaunch = 16384;
    const int n_batches_all = (nq + qmax - 1) / qmax;
    const int per_launch = std::min(n_batches_all, kMaxBatchesPerLaunch);
    const int n_ctas = std::max(1, std::min(std::min(ix->n_sms, n_tiles), 2 * ix->n_sms / per_launch));
    int rc = ix->ws_scan_keys.ensure((size_t)per_launch * n_ctas * qmax * k * 8);
    if (!rc) rc = ix->ws_gthr.ensure((size_t)per_launch * kFfmaMaxQ * 4);
    if (rc) return rc;
    for (int b0 = 0; b0 < n_batches_all; b0 += per_launch) {
        const int nb = std::min(per_launch, n_batches_all - b0);
        const int q0 = b0 * qmax;
        const int nq_l = std::min(nq - q0, nb * qmax);
        PQ_CUDA(cudaMemsetAsync(ix->ws_gthr.p, 0, (size_t)nb * kFfmaMaxQ * 4, ix->stream));
        FfmaLaunch f;
        f.tmap_rows_f32 = &ix->tmap_f32;
        f.row_norms = (const float*)ix->norms.p;
        f.queries_dev = dq + (size_t)q0 * kDim;
        f.out_keys = (uint64_t*)ix->ws_scan_keys.p;
        f.gthr = (uint32_t*)ix->ws_gthr.p;
        f.n_rows = ix->ntotal;
        f.n_ctas = n_ctas;
        f.nq = std::min(qmax, nq_l);
        f.k = k;
        f.metric = ix->metric;
        f.n_batches = nb;
        f.nq_total = nq_l;
        ix->prof_begin();
        PQ_CUDA(ffma_scan_launch(f, ix->stream));
        ix->prof_end();
        MergeLaunch m;
        memset(&m, 0, sizeof(m));
        m.keys = (const uint64_t*)ix->ws_scan_keys.p;
        m.q_stride = k;
        m.list_stride = (long long)f.nq * k;
        m.n_lists = n_ctas;
        m.list_len = k;
        m.gthr = (const uint32_t*)ix->ws_gthr.p;
        m.batch_q = f.nq;
        m.batch_stride = (long long)n_ctas * f.nq * k;
        m.gthr_batch_stride = kFfmaMaxQ;
        m.nq = nq_l;
        m.k = k;
        m.metric = ix->metric;
        m.q_norms = dq_norms + q0;
        m.id_base = ix->id_base;
        m.D = dD + (size_t)q0 * k;
        m.I = dI + (size_t)q0 * k;
        PQ_CUDA(merge_lists_launch(m, ix->stream));
        ix->stats[2] += 1;
        ix->stats[4] += 1;
        ix->stats[5] += 2;
    }
    return PQ_OK;
}

__global__ void pq_fill_empty_kernel(float* D, long long* I, long long n, float dval) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        D[i] = dval;
        I[i] = -1;
    }
}

// Gather / scatter helpers for the certificate-failure re-run.
__global__ void pq_gather_queries_kernel(const float* __restrict__ q, const float* __restrict__ qn, const int* __restrict__ idx, int n,
                                         float* __restrict__ out_q, float* __restrict__ out_qn) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const int src = idx[i];
    out_q[(size_t)i * kDim + threadIdx.x] = q[(size_t)src * kDim + threadIdx.x];
    if (threadIdx.x == 0) out_qn[i] = qn[src];
}
__global__ void pq_scatter_results_kernel(const float* __restrict__ Ds, const long long* __restrict__ Is, const int* __restrict__ idx,
                                          int n, int k, float* __restrict__ D, long long* __restrict__ I) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const int dst = idx[i];
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        D[(size_t)dst * k + j] = Ds[(size_t)i * k + j];
        I[(size_t)dst * k + j] = Is[(size_t)i * k + j];
    }
}

// 1024 < k: the tensor tier takes the sample-threshold path (pq_mma_largek.inl) when the corpus is large enough for a
// sample to mean something; PROQA_B200_LARGEK=0 sends such requests to the fp32 scan instead.
static bool tier_uses_largek(const pq_index* ix, int64_t nq, int64_t k) {
    if (!ix->largek || ix->has_nonfinite || ix->tier == PQ_TIER_FP32) return false;
    if (k > kMmaMaxK) return k <= PQ_MAX_K && (nq >= mma_min_queries() || ix->ntotal >= kMmaSmallBatchMinRows) && plan_large_k_applies(ix->ntotal, (int)k);
    // 512 <= k <= 1024 with a real batch (C5: 8192 queries, k = 1000): sample thresholds + one pass beat the epochs
    return k >= kPlanMidK && nq >= 256 && ix->ntotal >= (1 << 20) && plan_large_k_applies(ix->ntotal, (int)k);
}

static bool tier_uses_mma(const pq_index* ix, int64_t nq, int64_t k) {
    if (ix->has_nonfinite) return false;
    if (k > kMmaMaxK) return tier_uses_largek(ix, nq, k);
    if (ix->tier == PQ_TIER_FP32) return false;
    if (ix->tier == PQ_TIER_BF16) return ix->ntotal >= 1;
    // enough (query,row) pairs to pay for the epoch machinery: a big corpus, or the k-means shape (few centroids, millions of points)
    if (nq < mma_min_queries()) return ix->ntotal >= kMmaSmallBatchMinRows;
    return ix->ntotal >= kMmaMinRows || nq * ix->ntotal >= kMmaMinPairs;
}

// Device-resident search: dq [nq,128] fp32 -> dD [nq,k], dI [nq,k]; all on ix->stream; leaves the stream drained.
int search_device_impl(pq_index* ix, int64_t nq, const float* dq, int64_t k, float* dD, long long* dI) {
    memset(ix->stats, 0, sizeof(ix->stats));
    if (nq == 0) return PQ_OK;
    if (ix->ntotal == 0) {
        const long long n = nq * k;
        pq_fill_empty_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ix->stream>>>(dD, dI, n, ix->metric == kMetricL2 ? FLT_MAX : -FLT_MAX);
        PQ_CUDA(cudaGetLastError());
        PQ_CUDA(cudaStreamSynchronize(ix->stream));
        return PQ_OK;
    }
    PQ_CUDA(cudaEventRecord(ix->ev0, ix->stream));
    // Query preparation: squared norms (L2 output + error bound), bf16 copy, per-query non-finite flags.
    const int64_t nq_pad = (nq + 127) / 128 * 128;
    int rc = ix->ws_qnorm.ensure((size_t)nq_pad * 4);
    if (!rc) rc = ix->ws_qbf16.ensure((size_t)nq_pad * kDim * 2);
    if (!rc) rc = ix->ws_qbad.ensure((size_t)nq_pad);
    if (!rc) rc = ix->ws_qresid.ensure((size_t)nq_pad * 4);
    if (rc) return rc;
    PQ_CUDA(cudaMemsetAsync(ix->ws_qbf16.p, 0, (size_t)nq_pad * kDim * 2, ix->stream));
    PQ_CUDA(prep_rows_launch(dq, nq, (uint16_t*)ix->ws_qbf16.p, (float*)ix->ws_qnorm.p, nullptr, nullptr, (uint8_t*)ix->ws_qbad.p,
                             (float*)ix->ws_qresid.p, nullptr, ix->stream));
    ix->stats[5] += 1;

    if (tier_uses_mma(ix, nq, k)) {
        std::vector<int> rerun;
        rc = tier_uses_largek(ix, nq, k) ? search_mma_largek(ix, (int)nq, dq, (int)k, dD, dI, &rerun) : search_mma_filter(ix, (int)nq, dq, (int)k, dD, dI, &rerun);
        if (rc) return rc;
        ix->stats[0] = nq - (int64_t)rerun.size();
        ix->stats[1] = (int64_t)rerun.size();
        if (!rerun.empty()) {
            const int nr = (int)rerun.size();
            rc = ix->ws_rr_idx.ensure((size_t)nr * 4);
            if (!rc) rc = ix->ws_rr_q.ensure((size_t)nr * kDim * 4);
            if (!rc) rc = ix->ws_rr_qn.ensure((size_t)nr * 4);
            if (!rc) rc = ix->ws_rr_D.ensure((size_t)nr * k * 4);
            if (!rc) rc = ix->ws_rr_I.ensure((size_t)nr * k * 8);
            if (rc) return rc;
            PQ_CUDA(cudaMemcpyAsync(ix->ws_rr_idx.p, rerun.data(), (size_t)nr * 4, cudaMemcpyHostToDevice, ix->stream));
            pq_gather_queries_kernel<<<nr, kDim, 0, ix->stream>>>(dq, (const float*)ix->ws_qnorm.p, (const int*)ix->ws_rr_idx.p, nr,
                                                                  (float*)ix->ws_rr_q.p, (float*)ix->ws_rr_qn.p);
            PQ_CUDA(cudaGetLastError());
            rc = search_fp32_scan(ix, nr, (const float*)ix->ws_rr_q.p, (const float*)ix->ws_rr_qn.p, (int)k, (float*)ix->ws_rr_D.p,
                                  (long long*)ix->ws_rr_I.p);
            if (rc) return rc;
            pq_scatter_results_kernel<<<nr, 128, 0, ix->stream>>>((const float*)ix->ws_rr_D.p, (const long long*)ix->ws_rr_I.p,
                                                                  (const int*)ix->ws_rr_idx.p, nr, (int)k, dD, dI);
            PQ_CUDA(cudaGetLastError());
            ix->stats[5] += 2;
            PQ_CUDA(cudaStreamSynchronize(ix->stream));  // `rerun` (host vector) must outlive the H2D copy
        }
    } else {
        rc = search_fp32_scan(ix, (int)nq, dq, (const float*)ix->ws_qnorm.p, (int)k, dD, dI);
        if (rc) return rc;
    }
    PQ_CUDA(cudaEventRecord(ix->ev1, ix->stream));
    PQ_CUDA(cudaStreamSynchronize(ix->stream));
    float ms = 0.f;
    PQ_CUDA(cudaEventElapsedTime(&ms, ix->ev0, ix->ev1));
    ix->stats[6] = (int64_t)(ms * 1000.f);
    ix->stats[7] = ix->prof_collect();
    return PQ_OK;
}

static int check_search_args(pq_index* ix, int64_t nq, const void* xq, int64_t k, const void* D, const void* I) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    if (nq < 0 || k < 1) return set_error(PQ_ERR_INVALID, "search: bad arguments (nq=%lld, k=%lld)", (long long)nq, (long long)k);
    if (nq > 0 && (!xq || !D || !I)) return set_error(PQ_ERR_INVALID, "search: null buffer");
    if (k > PQ_MAX_K) return set_error(PQ_ERR_UNSUPPORTED, "search: k=%lld exceeds PQ_MAX_K=%d", (long long)k, PQ_MAX_K);
    if (nq > (int64_t)0x7fffff00) return set_error(PQ_ERR_UNSUPPORTED, "search: more than 2^31 queries in one call");
    return PQ_OK;
}

}  // namespace pq

using namespace pq;

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int pq_index_create(int d, int metric, int device, pq_index** out) {
    if (!out) return set_error(PQ_ERR_INVALID, "create: null out pointer");
    *out = nullptr;
    if (d != kDim) return set_error(PQ_ERR_INVALID, "create: d=%d, but this engine is built for d=128 (eval_retrieval.py:98)", d);
    if (metric != PQ_METRIC_IP && metric != PQ_METRIC_L2) return set_error(PQ_ERR_INVALID, "create: unknown metric %d", metric);
    pq_index* ix = new (std::nothrow) pq_index();
    if (!ix) return set_error(PQ_ERR_OOM, "create: host allocation failed");
    ix->d = d;
    ix->metric = metric;
    ix->requested_device = device;
    ix->tier = PQ_TIER_AUTO;
    const char* t = getenv("PROQA_B200_TIER");
    if (t) {
        if (!strcmp(t, "fp32")) ix->tier = PQ_TIER_FP32;
        else if (!strcmp(t, "bf16")) ix->tier = PQ_TIER_BF16;
    }
    const char* ks = getenv("PROQA_B200_K1_SETS");
    if (ks && (atoi(ks) == 2 || atoi(ks) == 4)) plan_k1_sets() = atoi(ks);
    const char* la = getenv("PROQA_B200_LOOSE_ABOVE");
    if (la && atof(la) > 0.0) plan_loose_above() = atof(la);
    const char* gr = getenv("PROQA_B200_GROWTH");
    if (gr && atoll(gr) > 1) plan_growth_override() = atoll(gr);
    const char* br = getenv("PROQA_B200_BOOT_ROWS");
    if (br && atoll(br) >= 1024) plan_boot_rows_override() = atoll(br) / 128 * 128;
    const char* lk = getenv("PROQA_B200_LARGEK");
    ix->largek = !(lk && !strcmp(lk, "0"));  // tensor tier for 1024 < k unless switched off
    *out = ix;  // CUDA is touched lazily (first add/search): the reference forks after importing faiss
    return PQ_OK;
}

void pq_index_free(pq_index* ix) {
    if (!ix) return;
    if (ix->device_ready) {
        std::lock_guard<std::mutex> lock(ix->mu);
        cudaSetDevice(ix->device);
        cudaStreamSynchronize(ix->stream);
        ix->release_all();
        cudaEventDestroy(ix->ev0);
        cudaEventDestroy(ix->ev1);
        for (cudaEvent_t e : ix->prof_events) cudaEventDestroy(e);
        cudaStreamDestroy(ix->own_stream);
    }
    delete ix;
}

int pq_index_add(pq_index* ix, int64_t n, const float* x_host) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(ix->mu);
    return index_add_locked(ix, n, x_host, false);
}
int pq_index_add_device(pq_index* ix, int64_t n, const float* x_dev) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(ix->mu);
    return index_add_locked(ix, n, x_dev, true);
}
int pq_index_add_f16(pq_index* ix, int64_t n, const void* x_host_f16) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(ix->mu);
    return index_add_f16_locked(ix, n, x_host_f16);
}

int pq_index_search(pq_index* ix, int64_t nq, const float* xq, int64_t k, float* D, int64_t* I) {
    int rc = check_search_args(ix, nq, xq, k, D, I);
    if (rc) return rc;
    if (nq == 0) return PQ_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    rc = index_init_device(ix);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(ix->device));
    rc = ix->ws_q.ensure((size_t)nq * kDim * 4);
    if (!rc) rc = ix->ws_D.ensure((size_t)nq * k * 4);
    if (!rc) rc = ix->ws_I.ensure((size_t)nq * k * 8);
    if (rc) return rc;
    rc = staged_upload(ix, ix->ws_q.p, xq, (size_t)nq * kDim * 4);
    if (rc) return rc;
    rc = search_device_impl(ix, nq, (const float*)ix->ws_q.p, k, (float*)ix->ws_D.p, (long long*)ix->ws_I.p);
    if (rc) return rc;
    if ((size_t)nq * k * 8 < kStagedCopyMinBytes) {  // small results (every eval_retrieval.py shape): two plain copies, one wait
        PQ_CUDA(cudaMemcpyAsync(D, ix->ws_D.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
        PQ_CUDA(cudaMemcpyAsync(I, ix->ws_I.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ix->stream));
        PQ_CUDA(cudaStreamSynchronize(ix->stream));
        return PQ_OK;
    }
    rc = staged_download(ix, D, ix->ws_D.p, (size_t)nq * k * 4);
    if (!rc) rc = staged_download(ix, I, ix->ws_I.p, (size_t)nq * k * 8);
    return rc;
}

int pq_index_search_device(pq_index* ix, int64_t nq, const float* xq_dev, int64_t k, float* D_dev, int64_t* I_dev) {
    int rc = check_search_args(ix, nq, xq_dev, k, D_dev, I_dev);
    if (rc) return rc;
    if (nq == 0) return PQ_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    rc = index_init_device(ix);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(ix->device));
    // The caller's tensors were produced on its own stream(s): unless we were given that very stream
    // (pq_index_set_stream), order after all prior device work.
    if (!ix->stream_is_external) PQ_CUDA(cudaDeviceSynchronize());
    return search_device_impl(ix, nq, xq_dev, k, D_dev, (long long*)I_dev);
}

int pq_index_reset(pq_index* ix) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(ix->mu);
    return pq::index_reset_locked(ix);
}
}  // extern "C"
namespace pq {
int index_reset_locked(pq_index* ix) {
    ix->ntotal = 0;
    ix->sample_rows = -1;
    ix->max_norm2 = 0.f;
    ix->max_resid2 = 0.f;
    ix->has_nonfinite = false;
    if (ix->device_ready) {
        PQ_CUDA(cudaSetDevice(ix->device));
        PQ_CUDA(cudaMemsetAsync(ix->scalars.p, 0, 256, ix->stream));
        PQ_CUDA(cudaStreamSynchronize(ix->stream));
    }
    return PQ_OK;
}
}  // namespace pq
extern "C" {

int64_t pq_index_ntotal(const pq_index* ix) { return ix ? ix->ntotal : 0; }
int pq_index_d(const pq_index* ix) { return ix ? ix->d : 0; }
int pq_index_metric(const pq_index* ix) { return ix ? ix->metric : -1; }

int pq_index_set_id_base(pq_index* ix, int64_t id_base) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    ix->id_base = id_base;
    return PQ_OK;
}
int pq_index_set_tier(pq_index* ix, int tier) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    if (tier < PQ_TIER_AUTO || tier > PQ_TIER_BF16) return set_error(PQ_ERR_INVALID, "unknown tier %d", tier);
    ix->tier = tier;
    return PQ_OK;
}
int pq_index_set_stream(pq_index* ix, void* cuda_stream, int is_external) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(ix->mu);
    if (ix->device_ready) {
        PQ_CUDA(cudaSetDevice(ix->device));
        PQ_CUDA(cudaStreamSynchronize(ix->stream));
    }
    ix->stream_is_external = is_external != 0;
    ix->stream = is_external ? (cudaStream_t)cuda_stream : ix->own_stream;
    return PQ_OK;
}
int pq_index_set_profile(pq_index* ix, int on) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    ix->profile = on != 0;
    return PQ_OK;
}
int pq_index_last_stats(const pq_index* ix, int64_t* out, int n) {
    if (!ix || !out || n < 0) return set_error(PQ_ERR_INVALID, "last_stats: bad arguments");
    for (int i = 0; i < n; ++i) out[i] = i < 16 ? ix->stats[i] : 0;
    return PQ_OK;
}

static int merge_shard_results_impl(int device, int metric, int n_lists, int64_t nq, int64_t k, const float* D_lists, const int64_t* I_lists,
                                    float* D_out, int64_t* I_out, cudaStream_t stream, bool sync) {
    if (n_lists < 1 || nq < 0 || k < 1 || !D_lists || !I_lists || !D_out || !I_out)
        return set_error(PQ_ERR_INVALID, "merge_shard_results: bad arguments");
    if (k > PQ_MAX_K) return set_error(PQ_ERR_UNSUPPORTED, "merge_shard_results: k=%lld exceeds PQ_MAX_K=%d", (long long)k, PQ_MAX_K);
    // up to 16384 keys per query are sorted inside one CTA; beyond that every entry is placed by ranking it against the
    // other lists (pq_merge_di_rank_kernel), which takes at most 64 lists
    if ((int64_t)n_lists * k > 16384 && n_lists > 64)
        return set_error(PQ_ERR_UNSUPPORTED, "merge_shard_results: %d lists of k=%lld (more than 64 lists beyond 16384 keys per query)", n_lists,
                         (long long)k);
    int dev = -1;
    int rc = pick_device(device, &dev);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(dev));
    if (sync) PQ_CUDA(cudaDeviceSynchronize());
    PQ_CUDA(merge_di_launch(D_lists, (const long long*)I_lists, n_lists, (int)nq, (int)k, metric, D_out, (long long*)I_out, stream));
    if (sync) PQ_CUDA(cudaDeviceSynchronize());
    return PQ_OK;
}

int pq_merge_shard_results(int device, int metric, int n_lists, int64_t nq, int64_t k, const float* D_lists, const int64_t* I_lists,
                           float* D_out, int64_t* I_out) {
    return merge_shard_results_impl(device, metric, n_lists, nq, k, D_lists, I_lists, D_out, I_out, 0, true);
}
int pq_merge_shard_results_async(int device, int metric, int n_lists, int64_t nq, int64_t k, const float* D_lists, const int64_t* I_lists,
                                 float* D_out, int64_t* I_out, void* cuda_stream) {
    return merge_shard_results_impl(device, metric, n_lists, nq, k, D_lists, I_lists, D_out, I_out, (cudaStream_t)cuda_stream, false);
}

// ---- cross-shard threshold exchange (pq_mma.cu: ShareParams; proqa_b200/sharded.py sets it up) ---------------------------
int pq_index_share_alloc(pq_index* ix, int n_ranks, int rank, int64_t cap_queries, void** mailbox_out, int64_t* bytes_out) {
    if (!ix || n_ranks < 1 || n_ranks > 16 || rank < 0 || rank >= n_ranks || cap_queries < 1 || cap_queries > (1 << 18))
        return set_error(PQ_ERR_INVALID, "share_alloc: bad arguments (n_ranks=%d, rank=%d, cap_queries=%lld)", n_ranks, rank, (long long)cap_queries);
    std::lock_guard<std::mutex> lock(ix->mu);
    int rc = index_init_device(ix);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(ix->device));
    pq_share_state& ss = ix->share;
    ss.connected = false;
    size_t bytes = ((size_t)n_ranks * (size_t)cap_queries * 2 + (size_t)n_ranks) * 8;
    // whole 2 MB blocks: the mailbox is exported through CUDA IPC, and small cudaMalloc blocks share a driver allocation with
    // whatever else is small (a peer cannot open two handles that resolve to the same allocation)
    const size_t alloc_bytes = (bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);
    ss.mailbox.release();
    rc = ss.mailbox.ensure(alloc_bytes);
    if (rc) return rc;
    PQ_CUDA(cudaMemsetAsync(ss.mailbox.p, 0, bytes, ix->stream));   // tag 0 is never used by a search
    PQ_CUDA(cudaStreamSynchronize(ix->stream));
    ss.n = n_ranks;
    ss.rank = rank;
    ss.cap_q = (int)cap_queries;
    ss.seq = 0;
    const char* w = getenv("PROQA_B200_SHARE_WAIT_US");
    if (w && *w) ss.wait_us = std::max(0, atoi(w));
    if (mailbox_out) *mailbox_out = ss.mailbox.p;
    if (bytes_out) *bytes_out = (int64_t)bytes;
    return PQ_OK;
}
int pq_index_share_connect(pq_index* ix, const void* const* peer_mailboxes) {
    if (!ix || !peer_mailboxes) return set_error(PQ_ERR_INVALID, "share_connect: bad arguments");
    std::lock_guard<std::mutex> lock(ix->mu);
    pq_share_state& ss = ix->share;
    if (ss.n < 1 || !ss.mailbox.p) return set_error(PQ_ERR_INVALID, "share_connect: call pq_index_share_alloc first");
    for (int i = 0; i < ss.n; ++i) {
        if (!peer_mailboxes[i]) return set_error(PQ_ERR_INVALID, "share_connect: mailbox %d is null", i);
        ss.peer[i] = (uint64_t*)peer_mailboxes[i];
    }
    if (ss.peer[ss.rank] != (uint64_t*)ss.mailbox.p) return set_error(PQ_ERR_INVALID, "share_connect: entry %d must be this index's own mailbox", ss.rank);
    ss.connected = true;
    return PQ_OK;
}
int pq_index_share_begin(pq_index* ix, uint32_t seq) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    ix->share.seq = (seq % 0x0ffffffeu) + 1u;   // 28 bits, never 0
    return PQ_OK;
}
int pq_index_share_close(pq_index* ix) {
    if (!ix) return set_error(PQ_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lock(ix->mu);
    ix->share.connected = false;
    ix->share.n = 0;
    return PQ_OK;
}
int pq_index_get_bound_scalars(const pq_index* ix, float* out2) {
    if (!ix || !out2) return set_error(PQ_ERR_INVALID, "get_bound_scalars: bad arguments");
    out2[0] = ix->max_norm2;
    out2[1] = ix->max_resid2;
    return PQ_OK;
}
int pq_index_set_bound_scalars(pq_index* ix, float max_norm2, float max_resid2) {
    if (!ix || !(max_norm2 >= 0.f) || !(max_resid2 >= 0.f)) return set_error(PQ_ERR_INVALID, "set_bound_scalars: bad arguments");
    std::lock_guard<std::mutex> lock(ix->mu);
    int rc = index_init_device(ix);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(ix->device));
    ix->max_norm2 = std::max(ix->max_norm2, max_norm2);
    ix->max_resid2 = std::max(ix->max_resid2, max_resid2);
    // the running maxima live on the device as well (later add() calls fold into them): non-negative floats order as uints
    uint32_t* sc = (uint32_t*)ix->scalars.p;
    PQ_CUDA(cudaMemcpyAsync(sc + 0, &ix->max_norm2, 4, cudaMemcpyHostToDevice, ix->stream));
    PQ_CUDA(cudaMemcpyAsync(sc + 2, &ix->max_resid2, 4, cudaMemcpyHostToDevice, ix->stream));
    PQ_CUDA(cudaStreamSynchronize(ix->stream));
    return PQ_OK;
}
int pq_ipc_export(const void* dev_ptr, void* handle_out64) {
    if (!dev_ptr || !handle_out64) return set_error(PQ_ERR_INVALID, "ipc_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    PQ_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle_out64, &h, 64);
    return PQ_OK;
}
int pq_ipc_open(const void* handle64, int device, void** dev_ptr_out) {
    if (!handle64 || !dev_ptr_out) return set_error(PQ_ERR_INVALID, "ipc_open: bad arguments");
    int dev = -1;
    int rc = pick_device(device, &dev);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(dev));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PQ_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return PQ_OK;
}
int pq_ipc_close(int device, void* dev_ptr) {
    int dev = -1;
    int rc = pick_device(device, &dev);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(dev));
    PQ_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return PQ_OK;
}
int pq_enable_peer_access(int device, int peer_device) {
    int dev = -1, peer = -1;
    int rc = pick_device(device, &dev);
    if (!rc) rc = pick_device(peer_device, &peer);
    if (rc) return rc;
    if (dev == peer) return PQ_OK;
    int can = 0;
    PQ_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer));
    if (!can) return set_error(PQ_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", dev, peer);
    PQ_CUDA(cudaSetDevice(dev));
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, __FILE__, __LINE__);
    cudaGetLastError();
    return PQ_OK;
}

const char* pq_last_error(void) { return g_err; }
const char* pq_version(void) { return "proqa_b200 0.1.0 sm_100a"; }

}  // extern "C"
