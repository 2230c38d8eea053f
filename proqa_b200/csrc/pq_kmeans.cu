// proqa_b200 — GPU k-means driver: the B200 counterpart of faiss.Clustering.train as ProQA calls it
// (retrieval/group_paras.py:40-45: Clustering(d, ncentroids); niter; max_points_per_centroid; train(x, index)).
//
// FAISS 1.6.3's Clustering.cpp is not in the reference tree (third-party wheel): what follows restates its published
// algorithm [upstream-memory, SURVEY.md §8c(6)] — subsample by rand_perm(seed) when n > k*max_points_per_centroid,
// centroids initialised from rand_perm(seed+1+redo*15486557), then niter rounds of
//     assign = index.search(x, 1)            -> the engine's k = 1 path (tensor-core filter + exact fp32 rescoring)
//     centroid c = mean of its points        -> here: stable sort of points by centroid, then one warp per centroid
//                                               adds its points IN INDEX ORDER in fp32 (the same sequential sum
//                                               FAISS's compute_centroids performs), times 1 / count
//     empty clusters split a big one         -> host, same RandomGenerator(1234) walk and +-1/1024 perturbation
//     index.reset(); index.add(centroids)
// The assignment search is the hot part (250 x 10M x 10k); everything else is < 10 % of an iteration.
#include "pq_common.cuh"
#include "pq_host.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <random>
#include <vector>


namespace pq {

// ---- FAISS RandomGenerator / rand_perm (utils/random.cpp) ---------------------------------------------------------
struct FaissRng {
    std::mt19937 mt;
    explicit FaissRng(int64_t seed) : mt((unsigned int)seed) {}
    int rand_int(int max) { return (int)(mt() % (unsigned)max); }
    float rand_float() { return mt() / float(mt.max()); }
};
static void rand_perm(std::vector<int>& perm, size_t n, int64_t seed) {
    perm.resize(n);
    for (size_t i = 0; i < n; ++i) perm[i] = (int)i;
    FaissRng rng(seed);
    for (size_t i = 0; i + 1 < n; ++i) {
        const int i2 = (int)i + rng.rand_int((int)(n - i));
        std::swap(perm[i], perm[i2]);
    }
}

// ---- kernels ------------------------------------------------------------------------------------------------------
__global__ void km_keys_kernel(const long long* __restrict__ assign, int n, int* __restrict__ keys, int* __restrict__ vals,
                               int* __restrict__ hist, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long a = assign[i];
    if (a < 0 || a >= k) a = 0;  // cannot happen with a non-empty index; keeps the sort keys in range
    keys[i] = (int)a;
    vals[i] = i;
    atomicAdd(hist + (int)a, 1);
}

// ---- stable sort of (centroid, point index) pairs by centroid: the engine's own LSD radix sort ---------------------------------
// The centroid sums must add each centroid's points in ascending point index (that is what makes them bit-identical to
// Clustering::train's sequential sums), so the sort has to be stable.  8-bit digits, ceil(key_bits / 8) passes, each:
//   km_sort_hist_kernel     block b counts the digits of its contiguous chunk of kSortChunk elements      -> H[digit][block]
//   km_sort_scan_kernel     exclusive prefix sum over H in (digit, block) order, one CTA                      -> base offsets
//   km_sort_scatter_kernel  block b walks its chunk in order, 256 elements at a time; an element's place is the base of
//                           (its digit, this block) + the elements of that digit seen earlier in the chunk (a shared-memory
//                           running count + the earlier warps of this tile + the earlier lanes of its warp: __match_any_sync)
// (Round 1 called cub::DeviceRadixSort here — the one library kernel on the f1 path.)
constexpr int kSortChunk = 4096;

__global__ void __launch_bounds__(256) km_sort_hist_kernel(const int* __restrict__ keys, int n, int shift, int n_blocks, int* __restrict__ H) {
    __shared__ int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const long long a = (long long)blockIdx.x * kSortChunk;
    const int end = (int)min((long long)n, a + kSortChunk);
    for (int i = (int)a + threadIdx.x; i < end; i += 256) atomicAdd(&s_h[(keys[i] >> shift) & 255], 1);
    __syncthreads();
    H[(size_t)threadIdx.x * n_blocks + blockIdx.x] = s_h[threadIdx.x];
}

__global__ void __launch_bounds__(1024) km_sort_scan_kernel(int* __restrict__ H, int total) {
    __shared__ int s_sum[1024];
    const int t = threadIdx.x;
    const int per = (total + 1023) / 1024;
    const int a = min(total, t * per), b = min(total, a + per);
    int run = 0;
    for (int i = a; i < b; ++i) run += H[i];
    s_sum[t] = run;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive scan of the per-thread sums
        const int v = t >= o ? s_sum[t - o] : 0;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    int base = s_sum[t] - run;
    for (int i = a; i < b; ++i) {
        const int c = H[i];
        H[i] = base;
        base += c;
    }
}

__global__ void __launch_bounds__(256) km_sort_scatter_kernel(const int* __restrict__ keys_in, const int* __restrict__ vals_in, int n, int shift,
                                                              int n_blocks, const int* __restrict__ H, int* __restrict__ keys_out,
                                                              int* __restrict__ vals_out) {
    __shared__ int s_run[256];       // elements of each digit placed so far by this block
    __shared__ int s_warp[8][256];   // per tile: digit counts of each warp
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    s_run[t] = H[(size_t)t * n_blocks + blockIdx.x];
    const long long a = (long long)blockIdx.x * kSortChunk;
    const int end = (int)min((long long)n, a + kSortChunk);
    for (int i0 = (int)a; i0 < end; i0 += 256) {
#pragma unroll
        for (int w = 0; w < 8; ++w) s_warp[w][t] = 0;
        __syncthreads();
        const int i = i0 + t;
        const bool in = i < end;
        const int key = in ? keys_in[i] : 0, val = in ? vals_in[i] : 0;
        const int digit = (key >> shift) & 255;
        const unsigned act = __ballot_sync(0xffffffffu, in);
        int rank_in_warp = 0;
        if (in) {
            const unsigned peers = __match_any_sync(act, digit);
            rank_in_warp = __popc(peers & ((1u << lane) - 1u));
            if (rank_in_warp == 0) s_warp[warp][digit] = __popc(peers);   // (the first lane of each digit group records the group's size)
        }
        __syncthreads();
        if (in) {
            int before = s_run[digit];
            for (int w = 0; w < warp; ++w) before += s_warp[w][digit];
            const int dst = before + rank_in_warp;
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += s_warp[w][t];
        s_run[t] += tot;
        __syncthreads();
    }
}

// Sorts n pairs by the low key_bits bits of the key, stably.  a and b are two (keys, vals) buffer pairs: the input is in a, the
// passes alternate between them, *vals_sorted tells where the sorted values ended up.  H: (256 * ceil(n / kSortChunk)) ints.
static cudaError_t km_sort_pairs(int* keys_a, int* vals_a, int* keys_b, int* vals_b, int n, int key_bits, int* H, cudaStream_t st,
                                 const int** vals_sorted) {
    const int n_blocks = (n + kSortChunk - 1) / kSortChunk;
    int *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
    for (int shift = 0; shift < key_bits; shift += 8) {
        km_sort_hist_kernel<<<n_blocks, 256, 0, st>>>(ki, n, shift, n_blocks, H);
        km_sort_scan_kernel<<<1, 1024, 0, st>>>(H, 256 * n_blocks);
        km_sort_scatter_kernel<<<n_blocks, 256, 0, st>>>(ki, vi, n, shift, n_blocks, H, ko, vo);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    *vals_sorted = vi;
    return cudaSuccess;
}
static size_t km_sort_workspace_bytes(long long n) { return (size_t)256 * (size_t)((n + kSortChunk - 1) / kSortChunk) * 4; }

// One warp per centroid: lane l owns dims 4l..4l+3; points are added in ascending point index (stable sort order).
// kMean = false leaves the plain sums (multi-GPU training: the shards' sums are all-reduced before the division).
template <bool kMean>
__global__ void __launch_bounds__(256) km_centroid_kernel(const float* __restrict__ x, const int* __restrict__ sorted_idx,
                                                          const int* __restrict__ offsets, const int* __restrict__ hist, int k,
                                                          float* __restrict__ centroids) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= k) return;
    const int n = hist[c];
    const int* idx = sorted_idx + offsets[c];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = 0;
    for (; j + 4 <= n; j += 4) {  // four row loads in flight, added strictly in order
        const int i0 = idx[j], i1 = idx[j + 1], i2 = idx[j + 2], i3 = idx[j + 3];
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(x + (size_t)i0 * kDim) + lane);
        const float4 r1 = __ldg(reinterpret_cast<const float4*>(x + (size_t)i1 * kDim) + lane);
        const float4 r2 = __ldg(reinterpret_cast<const float4*>(x + (size_t)i2 * kDim) + lane);
        const float4 r3 = __ldg(reinterpret_cast<const float4*>(x + (size_t)i3 * kDim) + lane);
        acc.x += r0.x; acc.y += r0.y; acc.z += r0.z; acc.w += r0.w;
        acc.x += r1.x; acc.y += r1.y; acc.z += r1.z; acc.w += r1.w;
        acc.x += r2.x; acc.y += r2.y; acc.z += r2.z; acc.w += r2.w;
        acc.x += r3.x; acc.y += r3.y; acc.z += r3.z; acc.w += r3.w;
    }
    for (; j < n; ++j) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(x + (size_t)idx[j] * kDim) + lane);
        acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
    }
    if (kMean && n > 0) {  // compute_centroids (Clustering.cpp, v1.6.3): float norm = 1 / hassign[ci]; c[j] *= norm;
        const float norm = 1.f / (float)n;
        acc.x *= norm; acc.y *= norm; acc.z *= norm; acc.w *= norm;
    }
    reinterpret_cast<float4*>(centroids + (size_t)c * kDim)[lane] = acc;
}

// centroid c = sums[c] / counts[c] (counts[c] == 0 leaves the zero sum; void clusters are handled on the host afterwards)
__global__ void __launch_bounds__(256) km_divide_kernel(const float* __restrict__ sums, const int* __restrict__ counts, int k,
                                                        float* __restrict__ centroids) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= k) return;
    float4 v = reinterpret_cast<const float4*>(sums + (size_t)c * kDim)[lane];
    const int n = counts[c];
    if (n > 0) {
        const float norm = 1.f / (float)n;
        v.x *= norm; v.y *= norm; v.z *= norm; v.w *= norm;
    }
    reinterpret_cast<float4*>(centroids + (size_t)c * kDim)[lane] = v;
}

// fvec_renorm_L2: every centroid scaled to unit norm (spherical k-means).
__global__ void __launch_bounds__(256) km_renorm_kernel(float* __restrict__ centroids, int k) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= k) return;
    float4 v = reinterpret_cast<float4*>(centroids + (size_t)c * kDim)[lane];
    const float nr = warp_engine_dot(v, v, lane);
    if (nr > 0.f) {
        const float inv = 1.0f / sqrtf(nr);
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        reinterpret_cast<float4*>(centroids + (size_t)c * kDim)[lane] = v;
    }
}

// Objective: sum of the k = 1 distances, in double, per-block partials summed on the host in block order (deterministic).
__global__ void __launch_bounds__(256) km_objective_kernel(const float* __restrict__ dis, int n, double* __restrict__ partial) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) s += (double)dis[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void km_gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ perm, int n, float* __restrict__ out) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    reinterpret_cast<float4*>(out + (size_t)i * kDim)[lane] = __ldg(reinterpret_cast<const float4*>(x + (size_t)perm[i] * kDim) + lane);
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// split_clusters (Clustering.cpp, v1.6.3) — the treatment of void clusters (host; touches only the few affected rows).
// FAISS keeps the cluster sizes as floats there: a split halves a size as a float, which feeds the probabilities of later splits.
// Host-side helpers of Clustering::train's input handling, spread over a few threads: the finite check reads the whole
// training matrix (10.75 GB for the 21M paragraph embeddings), the subsample gather copies 5 GB of scattered rows.
static int host_threads_for(int64_t work_items, int64_t per_thread_min) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    return (int)std::max<int64_t>(1, std::min<int64_t>(std::min(hw, 16u), work_items / std::max<int64_t>(1, per_thread_min)));
}
static bool all_finite_parallel(const float* x, int64_t n) {
    const int nt = host_threads_for(n, 1 << 22);
    std::atomic<bool> ok(true);
    auto scan = [&](int64_t a, int64_t b) {
        // a value is finite iff its exponent field is not all ones; OR-reduce a block at a time, stop early once something is found
        const uint32_t* u = reinterpret_cast<const uint32_t*>(x);
        for (int64_t i = a; i < b && ok.load(std::memory_order_relaxed); i += 4096) {
            const int64_t e = std::min(b, i + 4096);
            bool bad = false;
            for (int64_t j = i; j < e; ++j) bad |= (u[j] & 0x7f800000u) == 0x7f800000u;
            if (bad) ok.store(false, std::memory_order_relaxed);
        }
    };
    std::vector<std::thread> th;
    const int64_t per = (n + nt - 1) / nt;
    for (int t = 1; t < nt; ++t) th.emplace_back(scan, std::min(n, per * t), std::min(n, per * (t + 1)));
    scan(0, std::min(n, per));
    for (std::thread& t : th) t.join();
    return ok.load();
}
static void gather_rows_parallel(float* dst, const float* x, const int* rows, int64_t n) {
    const int nt = host_threads_for(n, 1 << 13);
    auto work = [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; ++i) memcpy(dst + (size_t)i * kDim, x + (size_t)rows[i] * kDim, sizeof(float) * kDim);
    };
    std::vector<std::thread> th;
    const int64_t per = (n + nt - 1) / nt;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, std::min(n, per * t), std::min(n, per * (t + 1)));
    work(0, std::min(n, per));
    for (std::thread& t : th) t.join();
}

static int split_void_clusters(std::vector<float>& cent, const std::vector<int>& counts, int64_t k, int64_t n) {
    const float EPS = 1.f / 1024.f;
    int nsplit = 0;
    std::vector<float> hassign(counts.begin(), counts.end());
    FaissRng rng(1234);
    for (int64_t ci = 0; ci < k; ++ci) {
        if (hassign[ci] != 0) continue;
        int64_t cj;
        for (cj = 0; true; cj = (cj + 1) % k) {
            const float p = (hassign[cj] - 1.0) / (float)(n - k);  // probability to pick this cluster for a split
            const float r = rng.rand_float();
            if (r < p) break;
        }
        memcpy(&cent[ci * kDim], &cent[cj * kDim], sizeof(float) * kDim);
        for (int j = 0; j < kDim; ++j) {  // small symmetric perturbation
            if (j % 2 == 0) {
                cent[ci * kDim + j] *= 1 + EPS;
                cent[cj * kDim + j] *= 1 - EPS;
            } else {
                cent[ci * kDim + j] *= 1 - EPS;
                cent[cj * kDim + j] *= 1 + EPS;
            }
        }
        hassign[ci] = hassign[cj] / 2;  // assume an even split of the cluster
        hassign[cj] -= hassign[ci];
        ++nsplit;
    }
    return nsplit;
}

static int kmeans_train_locked(pq_index* ix, int64_t k, const pq_kmeans_params& prm, int64_t n_in, const float* x_host, float* centroids_out,
                               float* obj_out, int64_t obj_cap, int64_t* n_obj) {
    if (k < 1 || n_in < k) return set_error(PQ_ERR_INVALID, "Number of training points (%lld) should be at least as large as number of clusters (%lld)",
                                            (long long)n_in, (long long)k);
    if (k > (1 << 24) || n_in > 0x7fffff00LL) return set_error(PQ_ERR_UNSUPPORTED, "kmeans: k or n too large");
    if (!all_finite_parallel(x_host, n_in * kDim)) return set_error(PQ_ERR_INVALID, "input contains NaN's or Inf's");
    int rc = index_init_device(ix);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const double t0 = now_s();

    // ---- subsample (Clustering::train: "Sampling a subset of %ld / %ld for training") ----
    int64_t nx = n_in;
    std::vector<int> perm;
    // every device buffer of this call is released on every way out (an error half-way must not strand gigabytes)
    DevBuf d_x, d_cent, d_D, d_I, d_keys, d_keys2, d_vals, d_vals2, d_hist, d_off, d_tmp, d_part, d_perm;
    void* pinned[2] = {nullptr, nullptr};
    auto release = [&]() {
        DevBuf* all[] = {&d_x, &d_cent, &d_D, &d_I, &d_keys, &d_keys2, &d_vals, &d_vals2, &d_hist, &d_off, &d_tmp, &d_part, &d_perm};
        for (DevBuf* b : all) b->release();
        for (void*& h : pinned) {
            if (h) cudaFreeHost(h);
            h = nullptr;
        }
    };
#define KM_TRY(expr)              \
    do {                          \
        int _rc = (expr);         \
        if (_rc) {                \
            release();            \
            return _rc;           \
        }                         \
    } while (0)
#define KM_CUDA(expr)                                     \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) {                          \
            release();                                    \
            return cuda_fail(_e, __FILE__, __LINE__);     \
        }                                                 \
    } while (0)
    const int64_t max_train = k * (int64_t)prm.max_points_per_centroid;
    if (n_in > max_train) {
        if (prm.verbose) printf("Sampling a subset of %ld / %ld for training\n", (long)max_train, (long)n_in);
        rand_perm(perm, (size_t)n_in, prm.seed);
        nx = max_train;
    } else if (n_in < k * (int64_t)prm.min_points_per_centroid) {
        fprintf(stderr, "WARNING clustering %ld points to %ld centroids: please provide at least %ld training points\n", (long)n_in, (long)k,
                (long)(k * (int64_t)prm.min_points_per_centroid));
    }
    KM_TRY(d_x.ensure((size_t)nx * kDim * 4));
    if (n_in > max_train) {
        // The subsample (group_paras.py defaults: 10M of 21M rows, 5 GB) is gathered by host threads into two pinned buffers
        // that alternate between being filled and being copied: the gather of chunk i+1 overlaps the H2D copy of chunk i.
        const int64_t chunk = 1 << 17;  // 64 MB
        cudaEvent_t copied[2] = {nullptr, nullptr};
        for (int i = 0; i < 2; ++i) {
            KM_CUDA(cudaMallocHost(&pinned[i], (size_t)chunk * kDim * 4));
            KM_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        }
        cudaError_t ge = cudaSuccess;
        int which = 0;
        for (int64_t a = 0; a < nx && ge == cudaSuccess; a += chunk, which ^= 1) {
            const int64_t b = std::min(nx, a + chunk);
            ge = cudaEventSynchronize(copied[which]);
            if (ge != cudaSuccess) break;
            gather_rows_parallel((float*)pinned[which], x_host, perm.data() + a, b - a);
            ge = cudaMemcpyAsync((float*)d_x.p + (size_t)a * kDim, pinned[which], (size_t)(b - a) * kDim * 4, cudaMemcpyHostToDevice, st);
            if (ge == cudaSuccess) ge = cudaEventRecord(copied[which], st);
        }
        if (ge == cudaSuccess) ge = cudaStreamSynchronize(st);
        for (int i = 0; i < 2; ++i) cudaEventDestroy(copied[i]);
        KM_CUDA(ge);
        for (void*& h : pinned) {
            cudaFreeHost(h);
            h = nullptr;
        }
    } else {
        KM_CUDA(cudaMemcpyAsync(d_x.p, x_host, (size_t)nx * kDim * 4, cudaMemcpyHostToDevice, st));
    }
    const float* dx = (const float*)d_x.p;

    if (nx == k) {  // "Number of training points same as number of clusters, just copying"
        KM_CUDA(cudaMemcpyAsync(centroids_out, dx, (size_t)k * kDim * 4, cudaMemcpyDeviceToHost, st));
        KM_CUDA(cudaStreamSynchronize(st));
        rc = index_reset_locked(ix);
        if (!rc) rc = index_add_locked(ix, k, dx, true);
        release();
        if (n_obj) *n_obj = 0;
        return rc;
    }
    if (prm.verbose)
        printf("Clustering %d points in %dD to %ld clusters, redo %d times, %d iterations\n", (int)nx, kDim, (long)k, prm.nredo, prm.niter);

    KM_TRY(d_cent.ensure((size_t)k * kDim * 4));
    KM_TRY(d_D.ensure((size_t)nx * 4));
    KM_TRY(d_I.ensure((size_t)nx * 8));
    KM_TRY(d_keys.ensure((size_t)nx * 4));
    KM_TRY(d_keys2.ensure((size_t)nx * 4));
    KM_TRY(d_vals.ensure((size_t)nx * 4));
    KM_TRY(d_vals2.ensure((size_t)nx * 4));
    KM_TRY(d_hist.ensure((size_t)k * 4));
    KM_TRY(d_off.ensure((size_t)k * 4));
    KM_TRY(d_part.ensure(1024 * 8));
    KM_TRY(d_perm.ensure((size_t)k * 4));
    int key_bits = 1;
    while ((1LL << key_bits) < k) ++key_bits;
    KM_TRY(d_tmp.ensure(km_sort_workspace_bytes(nx)));

    std::vector<float> cent((size_t)k * kDim), best_cent;
    std::vector<int> hassign((size_t)k), offsets((size_t)k);
    std::vector<float> obj, best_obj;
    std::vector<double> partial(1024);
    float best_err = HUGE_VALF;
    double t_search_tot = 0.0;

    for (int redo = 0; redo < prm.nredo; ++redo) {
        if (prm.verbose && prm.nredo > 1) printf("Outer iteration %d / %d\n", redo, prm.nredo);
        // initialise the centroids with random points of the (sub-sampled) training set
        std::vector<int> perm2;
        rand_perm(perm2, (size_t)nx, prm.seed + 1 + redo * 15486557L);
        KM_CUDA(cudaMemcpyAsync(d_perm.p, perm2.data(), (size_t)k * 4, cudaMemcpyHostToDevice, st));
        km_gather_rows_kernel<<<(unsigned)((k + 7) / 8), 256, 0, st>>>(dx, (const int*)d_perm.p, (int)k, (float*)d_cent.p);
        KM_CUDA(cudaGetLastError());
        if (prm.spherical) km_renorm_kernel<<<(unsigned)((k + 7) / 8), 256, 0, st>>>((float*)d_cent.p, (int)k);
        KM_CUDA(cudaStreamSynchronize(st));
        KM_TRY(index_reset_locked(ix));
        KM_TRY(index_add_locked(ix, k, (const float*)d_cent.p, true));
        obj.clear();
        float err = 0.f;
        for (int it = 0; it < prm.niter; ++it) {
            const double t0s = now_s();
            KM_TRY(search_device_impl(ix, nx, dx, 1, (float*)d_D.p, (long long*)d_I.p));  // leaves the stream drained
            t_search_tot += now_s() - t0s;
            // objective
            km_objective_kernel<<<1024, 256, 0, st>>>((const float*)d_D.p, (int)nx, (double*)d_part.p);
            KM_CUDA(cudaMemcpyAsync(partial.data(), d_part.p, 1024 * 8, cudaMemcpyDeviceToHost, st));
            // centroid update: histogram + stable sort by centroid, then one warp per centroid
            KM_CUDA(cudaMemsetAsync(d_hist.p, 0, (size_t)k * 4, st));
            km_keys_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, st>>>((const long long*)d_I.p, (int)nx, (int*)d_keys.p, (int*)d_vals.p, (int*)d_hist.p,
                                                                        (int)k);
            const int* sorted_idx = nullptr;
            KM_CUDA(km_sort_pairs((int*)d_keys.p, (int*)d_vals.p, (int*)d_keys2.p, (int*)d_vals2.p, (int)nx, key_bits, (int*)d_tmp.p, st, &sorted_idx));
            KM_CUDA(cudaMemcpyAsync(hassign.data(), d_hist.p, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
            KM_CUDA(cudaStreamSynchronize(st));
            double e64 = 0.0;
            for (int b = 0; b < 1024; ++b) e64 += partial[b];
            err = (float)e64;
            obj.push_back(err);
            int run = 0, n_void = 0;
            double uf = 0.0;
            for (int64_t c = 0; c < k; ++c) {
                offsets[c] = run;
                run += hassign[c];
                n_void += hassign[c] == 0;
                uf += (double)hassign[c] * (double)hassign[c];
            }
            const double imbalance = uf * (double)k / ((double)run * (double)run);
            KM_CUDA(cudaMemcpyAsync(d_off.p, offsets.data(), (size_t)k * 4, cudaMemcpyHostToDevice, st));
            km_centroid_kernel<true><<<(unsigned)((k + 7) / 8), 256, 0, st>>>(dx, sorted_idx, (const int*)d_off.p, (const int*)d_hist.p, (int)k,
                                                                       (float*)d_cent.p);
            KM_CUDA(cudaGetLastError());
            int nsplit = 0;
            if (n_void) {
                KM_CUDA(cudaMemcpyAsync(cent.data(), d_cent.p, (size_t)k * kDim * 4, cudaMemcpyDeviceToHost, st));
                KM_CUDA(cudaStreamSynchronize(st));
                nsplit = split_void_clusters(cent, hassign, k, nx);
                KM_CUDA(cudaMemcpyAsync(d_cent.p, cent.data(), (size_t)k * kDim * 4, cudaMemcpyHostToDevice, st));
            }
            if (prm.verbose) {
                printf("\r  Iteration %d (%.2f s, search %.2f s): objective=%g imbalance=%.3f nsplit=%d       ", it, now_s() - t0, t_search_tot, err, imbalance,
                       nsplit);
                fflush(stdout);
            }
            if (prm.spherical) km_renorm_kernel<<<(unsigned)((k + 7) / 8), 256, 0, st>>>((float*)d_cent.p, (int)k);
            KM_CUDA(cudaStreamSynchronize(st));
            KM_TRY(index_reset_locked(ix));
            KM_TRY(index_add_locked(ix, k, (const float*)d_cent.p, true));
        }
        if (prm.verbose) printf("\n");
        if (prm.nredo > 1) {
            if (err < best_err) {
                if (prm.verbose) printf("Objective improved: keep new clusters\n");
                best_cent.resize((size_t)k * kDim);
                KM_CUDA(cudaMemcpyAsync(best_cent.data(), d_cent.p, (size_t)k * kDim * 4, cudaMemcpyDeviceToHost, st));
                KM_CUDA(cudaStreamSynchronize(st));
                best_obj = obj;
                best_err = err;
            }
        }
    }
    if (prm.nredo > 1) {
        memcpy(centroids_out, best_cent.data(), (size_t)k * kDim * 4);
        obj = best_obj;
        KM_CUDA(cudaMemcpyAsync(d_cent.p, best_cent.data(), (size_t)k * kDim * 4, cudaMemcpyHostToDevice, st));
        KM_CUDA(cudaStreamSynchronize(st));
        KM_TRY(index_reset_locked(ix));
        KM_TRY(index_add_locked(ix, k, (const float*)d_cent.p, true));
    } else {
        KM_CUDA(cudaMemcpyAsync(centroids_out, d_cent.p, (size_t)k * kDim * 4, cudaMemcpyDeviceToHost, st));
        KM_CUDA(cudaStreamSynchronize(st));
    }
    if (obj_out)
        for (size_t i = 0; i < obj.size() && (int64_t)i < obj_cap; ++i) obj_out[i] = obj[i];
    if (n_obj) *n_obj = (int64_t)obj.size();
    release();
#undef KM_TRY
#undef KM_CUDA
    return PQ_OK;
}

// ---- the same iteration in three steps, for training with the points sharded over several GPUs ------------------------
// (proqa_b200/sharded_clustering.py: every rank calls partial on its points, the sums / counts / objective are all-reduced
//  over NCCL, every rank calls finish with the identical totals and ends up with identical centroids.)
static int kmeans_set_centroids_locked(pq_index* ix, int64_t k, const float* cent_host, int spherical) {
    int rc = index_init_device(ix);
    if (rc) return rc;
    PQ_CUDA(cudaSetDevice(ix->device));
    rc = ix->ws_km[0].ensure((size_t)k * kDim * 4);
    if (rc) return rc;
    PQ_CUDA(cudaMemcpyAsync(ix->ws_km[0].p, cent_host, (size_t)k * kDim * 4, cudaMemcpyHostToDevice, ix->stream));
    if (spherical) km_renorm_kernel<<<(unsigned)((k + 7) / 8), 256, 0, ix->stream>>>((float*)ix->ws_km[0].p, (int)k);
    PQ_CUDA(cudaStreamSynchronize(ix->stream));
    rc = index_reset_locked(ix);
    if (!rc) rc = index_add_locked(ix, k, (const float*)ix->ws_km[0].p, true);
    return rc;
}

static int kmeans_partial_locked(pq_index* ix, int64_t k, int64_t n, const float* dx, float* sums_dev, int* counts_dev, double* obj_out) {
    if (!ix->device_ready || ix->ntotal != k) return set_error(PQ_ERR_INVALID, "kmeans_partial: the index must hold the k current centroids");
    if (n < 0 || n > 0x7fffff00LL || k > (1 << 24)) return set_error(PQ_ERR_UNSUPPORTED, "kmeans_partial: k or n too large");
    PQ_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    PQ_CUDA(cudaMemsetAsync(sums_dev, 0, (size_t)k * kDim * 4, st));
    PQ_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)k * 4, st));
    *obj_out = 0.0;
    if (n == 0) {
        PQ_CUDA(cudaStreamSynchronize(st));
        return PQ_OK;
    }
    DevBuf* w = ix->ws_km;  // [1] D, [2] I, [3] keys, [4] keys2, [5] vals, [6] vals2, [7] offsets, [8] sort scratch, [9] partials
    int rc = w[1].ensure((size_t)n * 4);
    if (!rc) rc = w[2].ensure((size_t)n * 8);
    for (int i = 3; i <= 6 && !rc; ++i) rc = w[i].ensure((size_t)n * 4);
    if (!rc) rc = w[7].ensure((size_t)k * 4);
    if (!rc) rc = w[9].ensure(1024 * 8);
    if (rc) return rc;
    int key_bits = 1;
    while ((1LL << key_bits) < k) ++key_bits;
    rc = w[8].ensure(km_sort_workspace_bytes(n));
    if (rc) return rc;
    rc = search_device_impl(ix, n, dx, 1, (float*)w[1].p, (long long*)w[2].p);  // leaves the stream drained
    if (rc) return rc;
    std::vector<double> partial(1024);
    km_objective_kernel<<<1024, 256, 0, st>>>((const float*)w[1].p, (int)n, (double*)w[9].p);
    PQ_CUDA(cudaMemcpyAsync(partial.data(), w[9].p, 1024 * 8, cudaMemcpyDeviceToHost, st));
    km_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const long long*)w[2].p, (int)n, (int*)w[3].p, (int*)w[5].p, counts_dev, (int)k);
    const int* sorted_idx = nullptr;
    PQ_CUDA(km_sort_pairs((int*)w[3].p, (int*)w[5].p, (int*)w[4].p, (int*)w[6].p, (int)n, key_bits, (int*)w[8].p, st, &sorted_idx));
    std::vector<int> hassign((size_t)k), offsets((size_t)k);
    PQ_CUDA(cudaMemcpyAsync(hassign.data(), counts_dev, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
    PQ_CUDA(cudaStreamSynchronize(st));
    int run = 0;
    for (int64_t c = 0; c < k; ++c) {
        offsets[c] = run;
        run += hassign[c];
    }
    PQ_CUDA(cudaMemcpyAsync(w[7].p, offsets.data(), (size_t)k * 4, cudaMemcpyHostToDevice, st));
    km_centroid_kernel<false><<<(unsigned)((k + 7) / 8), 256, 0, st>>>(dx, sorted_idx, (const int*)w[7].p, counts_dev, (int)k, sums_dev);
    PQ_CUDA(cudaGetLastError());
    PQ_CUDA(cudaStreamSynchronize(st));  // `offsets` (host vector) must outlive its copy; the caller all-reduces next
    double e64 = 0.0;
    for (int b = 0; b < 1024; ++b) e64 += partial[b];
    *obj_out = e64;
    return PQ_OK;
}

static int kmeans_finish_locked(pq_index* ix, int64_t k, int64_t n_total, int spherical, const float* sums_dev, const int* counts_dev,
                                float* centroids_out, int* nsplit_out) {
    if (!ix->device_ready) return set_error(PQ_ERR_INVALID, "kmeans_finish: index has no device yet");
    if (n_total < k) return set_error(PQ_ERR_INVALID, "kmeans_finish: fewer points than clusters");
    PQ_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    int rc = ix->ws_km[0].ensure((size_t)k * kDim * 4);
    if (rc) return rc;
    float* d_cent = (float*)ix->ws_km[0].p;
    km_divide_kernel<<<(unsigned)((k + 7) / 8), 256, 0, st>>>(sums_dev, counts_dev, (int)k, d_cent);
    PQ_CUDA(cudaGetLastError());
    std::vector<int> hassign((size_t)k);
    PQ_CUDA(cudaMemcpyAsync(hassign.data(), counts_dev, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
    PQ_CUDA(cudaStreamSynchronize(st));
    int n_void = 0;
    for (int64_t c = 0; c < k; ++c) n_void += hassign[c] == 0;
    int nsplit = 0;
    if (n_void) {
        std::vector<float> cent((size_t)k * kDim);
        PQ_CUDA(cudaMemcpyAsync(cent.data(), d_cent, (size_t)k * kDim * 4, cudaMemcpyDeviceToHost, st));
        PQ_CUDA(cudaStreamSynchronize(st));
        nsplit = split_void_clusters(cent, hassign, k, n_total);
        PQ_CUDA(cudaMemcpyAsync(d_cent, cent.data(), (size_t)k * kDim * 4, cudaMemcpyHostToDevice, st));
        PQ_CUDA(cudaStreamSynchronize(st));
    }
    if (spherical) km_renorm_kernel<<<(unsigned)((k + 7) / 8), 256, 0, st>>>(d_cent, (int)k);
    if (centroids_out) PQ_CUDA(cudaMemcpyAsync(centroids_out, d_cent, (size_t)k * kDim * 4, cudaMemcpyDeviceToHost, st));
    PQ_CUDA(cudaStreamSynchronize(st));
    if (nsplit_out) *nsplit_out = nsplit;
    rc = index_reset_locked(ix);
    if (!rc) rc = index_add_locked(ix, k, d_cent, true);
    return rc;
}

}  // namespace pq

extern "C" {

void pq_rand_perm(int64_t n, int64_t seed, int32_t* out) {
    if (n <= 0 || !out) return;
    std::vector<int> perm;
    pq::rand_perm(perm, (size_t)n, seed);
    memcpy(out, perm.data(), (size_t)n * 4);
}

int pq_kmeans_set_centroids(pq_index* index, int64_t k, const float* centroids_host, int spherical) {
    if (!index || !centroids_host || k < 1) return pq::set_error(PQ_ERR_INVALID, "kmeans_set_centroids: bad argument");
    std::lock_guard<std::mutex> lock(index->mu);
    return pq::kmeans_set_centroids_locked(index, k, centroids_host, spherical);
}

int pq_kmeans_partial_device(pq_index* index, int64_t k, int64_t n_local, const float* x_dev, float* sums_dev, int32_t* counts_dev,
                             double* objective_out) {
    if (!index || !sums_dev || !counts_dev || !objective_out || (n_local > 0 && !x_dev) || k < 1)
        return pq::set_error(PQ_ERR_INVALID, "kmeans_partial_device: bad argument");
    std::lock_guard<std::mutex> lock(index->mu);
    return pq::kmeans_partial_locked(index, k, n_local, x_dev, sums_dev, (int*)counts_dev, objective_out);
}

int pq_kmeans_finish_device(pq_index* index, int64_t k, int64_t n_total, int spherical, const float* sums_dev, const int32_t* counts_dev,
                            float* centroids_out_host, int* nsplit_out) {
    if (!index || !sums_dev || !counts_dev || k < 1) return pq::set_error(PQ_ERR_INVALID, "kmeans_finish_device: bad argument");
    std::lock_guard<std::mutex> lock(index->mu);
    return pq::kmeans_finish_locked(index, k, n_total, spherical, sums_dev, (const int*)counts_dev, centroids_out_host, nsplit_out);
}

void pq_kmeans_default_params(pq_kmeans_params* p) {
    if (!p) return;
    p->niter = 25;
    p->nredo = 1;
    p->verbose = 0;
    p->spherical = 0;
    p->min_points_per_centroid = 39;
    p->max_points_per_centroid = 256;
    p->seed = 1234;
}

int pq_kmeans_train(pq_index* index, int64_t k, const pq_kmeans_params* params, int64_t n, const float* x_host, float* centroids_out, float* obj_out,
                    int64_t obj_cap, int64_t* n_obj) {
    if (!index || !params || !x_host || !centroids_out) return pq::set_error(PQ_ERR_INVALID, "kmeans_train: null argument");
    std::lock_guard<std::mutex> lock(index->mu);
    return pq::kmeans_train_locked(index, k, *params, n, x_host, centroids_out, obj_out, obj_cap, n_obj);
}

}  // extern "C"
