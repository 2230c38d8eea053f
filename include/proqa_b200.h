/* proqa_b200 — C ABI of the B200-native exact top-k MIPS engine.
 *
 * This is the drop-in boundary for ProQA's retrieval hot path.  The reference reaches the path
 * through the FAISS Python (SWIG) API of faiss-cpu==1.6.3 (/root/reference/requirements.txt:2);
 * each entry point below names the reference call it stands behind.  Plain pointers and sizes
 * only — no torch, numpy or C++ types cross this boundary.  All functions return 0 on success or
 * a negative pq_status; pq_last_error() gives the thread-local message.  Nothing aborts.
 *
 * Ownership: the caller owns every host buffer passed in; the engine copies what it keeps and
 * owns all device memory.  Output buffers are caller-allocated and fully written before return
 * (search is synchronous, like IndexFlat::search).  One index may be searched from one thread at a
 * time; distinct indexes may be used from distinct threads.
 *
 * CUDA is initialised lazily by the first call that needs the device (never at load time):
 * retrieval/eval_retrieval.py:92-96 forks its worker pool after `import faiss` (:4).
 */
#ifndef PROQA_B200_H_
#define PROQA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pq_index pq_index;

enum pq_metric { PQ_METRIC_IP = 0, PQ_METRIC_L2 = 1 };

enum pq_status {
    PQ_OK = 0,
    PQ_ERR_INVALID = -1,     /* bad argument (shape, k, null pointer, d != 128 ...)          */
    PQ_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed; message has the details  */
    PQ_ERR_NO_DEVICE = -3,   /* no sm_100 device visible: there is NO CPU fallback            */
    PQ_ERR_OOM = -4,         /* device or pinned-host allocation failed                       */
    PQ_ERR_UNSUPPORTED = -5  /* k above PQ_MAX_K, or a tier that cannot serve the request     */
};

/* Search tiers (pq_index_set_tier / env PROQA_B200_TIER):
 *   AUTO : fp32 scan for small corpora (and for fewer than 5 queries below 2^23 rows), tensor-core filter otherwise.
 *   FP32 : always the exact fp32 streaming scan (coalesced FFMA path).
 *   BF16 : always the tcgen05 bf16 filter + exact fp32 rescoring with an exactness certificate;
 *          queries whose certificate fails are re-run through the fp32 scan.
 * Every tier returns the same bits: the top-k under the engine's defined fp32 score. */
enum pq_tier { PQ_TIER_AUTO = 0, PQ_TIER_FP32 = 1, PQ_TIER_BF16 = 2 };

#define PQ_MAX_K 15360

/* faiss.IndexFlatIP(d) / faiss.IndexFlatL2(d)
 *   retrieval/eval_retrieval.py:102, retrieval/group_paras.py:36,38, retrieval/trec_process.py:74.
 * d must be 128 (eval_retrieval.py:98 hard-codes it).  device < 0 selects the current device
 * (or env PROQA_B200_DEVICE / LOCAL_RANK). */
int pq_index_create(int d, int metric, int device, pq_index** out);
void pq_index_free(pq_index* idx);

/* index.add(xb)  — eval_retrieval.py:103, group_paras.py:50, trec_process.py:75.
 * Appends n rows (C-contiguous float32 [n, d]); ids are sequential insertion order. */
int pq_index_add(pq_index* idx, int64_t n, const float* x_host);
/* Same, rows given as IEEE half (C-contiguous float16 [n, d]) — what get_embed.py --fp16 saves (retrieval/get_embed.py:147-151)
 * and eval_retrieval.py:100 widens on the host with .astype('float32'): here the widening happens on the device, exactly,
 * after half as many bytes crossed PCIe. */
int pq_index_add_f16(pq_index* idx, int64_t n, const void* x_host_f16);
/* Same, rows already in device memory (skips the .npy -> host -> device round trip). */
int pq_index_add_device(pq_index* idx, int64_t n, const float* x_dev);

/* index.search(xq, k) -> (D, I)  — eval_retrieval.py:104 (k=80), trec_process.py:76 (k=10000),
 * group_paras.py:51 and Clustering.train's per-iteration assignment (k=1).
 * D: float32 [nq, k] best-first (IP: descending score; L2: ascending squared distance),
 * I: int64  [nq, k]; slots beyond ntotal are I=-1, D=-FLT_MAX (IP) / +FLT_MAX (L2). */
int pq_index_search(pq_index* idx, int64_t nq, const float* xq_host, int64_t k, float* D_host, int64_t* I_host);
/* Same with queries and outputs in device memory (used by the multi-GPU layer and the k-means
 * driver; D_dev/I_dev are written on the index's stream and the call returns after it drained). */
int pq_index_search_device(pq_index* idx, int64_t nq, const float* xq_dev, int64_t k, float* D_dev, int64_t* I_dev);

/* index.reset()  — group_paras.py:49.  ntotal = 0; device storage is kept for reuse. */
int pq_index_reset(pq_index* idx);
/* index.ntotal / index.d / index.metric_type */
int64_t pq_index_ntotal(const pq_index* idx);
int pq_index_d(const pq_index* idx);
int pq_index_metric(const pq_index* idx);

/* Row ids reported by search are local row + id_base (multi-GPU row shards: base of the shard). */
int pq_index_set_id_base(pq_index* idx, int64_t id_base);
int pq_index_set_tier(pq_index* idx, int tier);
/* Run this index's device work on the caller's CUDA stream (cudaStream_t passed as void*; is_external=1), or
 * back on the index's own stream (is_external=0).  With an external stream, *_device calls do not
 * device-synchronise first: the caller's prior work on that stream is already ordered before ours. */
int pq_index_set_stream(pq_index* idx, void* cuda_stream, int is_external);
/* Time the dominant kernel of every search (fp32 scan / tensor-core filter) with CUDA events on the
 * launching stream; the total lands in stats[7] (microseconds). */
int pq_index_set_profile(pq_index* idx, int on);
/* Counters of the last search: [0]=queries served by the tensor-core tier, [1]=queries re-run by the
 * fp32 scan after a failed certificate, [2]=fp32-scan launches, [3]=tensor-core filter launches,
 * [4]=select/merge/rescore launches, [5]=total kernel launches, [6]=device microseconds (CUDA events),
 * [7]=microseconds inside the dominant kernel (only with pq_index_set_profile), [8]=(query, epoch) pairs repaired after the
 * last epoch (candidate slabs overflowed: rows in document order can bring a whole cluster above the threshold at once),
 * [9]=threshold exchanges between row shards in which every shard's values had arrived in time; with pq_index_set_profile also
 * [11]=microseconds inside the epoch-select kernels, [12]=inside the threshold-fold kernels, [13]=inside the rescoring kernel
 * (n up to 16). */
int pq_index_last_stats(const pq_index* idx, int64_t* out, int n);

/* Merge G per-shard result lists (each [nq,k], best-first, global ids) into one — the kernel run
 * after the NCCL all-gather in the multi-GPU layer.  All pointers are device pointers on `device`.  Up to 16384 keys per
 * query (G * k) are sorted inside one CTA; beyond that every entry finds its place by ranking against the other lists
 * (G <= 64). */
int pq_merge_shard_results(int device, int metric, int n_lists, int64_t nq, int64_t k, const float* D_lists_dev,
                           const int64_t* I_lists_dev, float* D_out_dev, int64_t* I_out_dev);

/* Same, enqueued on the caller's CUDA stream (cudaStream_t as void*) without any device synchronisation: the all-gather
 * that produced the lists and the consumer of the result are ordered by that stream. */
int pq_merge_shard_results_async(int device, int metric, int n_lists, int64_t nq, int64_t k, const float* D_lists_dev,
                                 const int64_t* I_lists_dev, float* D_out_dev, int64_t* I_out_dev, void* cuda_stream);

/* faiss.Clustering(d, k) ... clus.train(x, index)  — retrieval/group_paras.py:40-45 (FAISS 1.6.3 Clustering.cpp semantics).
 * Fields mirror the ClusteringParameters the script sets (verbose :41, niter :42, max_points_per_centroid :43) plus the
 * FAISS defaults it leaves alone.  `index` is the IndexFlatL2 / IndexFlatIP the script built (:35-38); on return it holds
 * the final centroids (as after FAISS's last index.add).  centroids_out: [k*d] floats; obj_out (optional): one objective
 * per iteration, at most obj_cap of them, *n_obj receives how many iterations ran. */
typedef struct pq_kmeans_params {
    int niter;                   /* 25   */
    int nredo;                   /* 1    */
    int verbose;                 /* 0    */
    int spherical;               /* 0: centroids L2-normalised after each iteration when set */
    int min_points_per_centroid; /* 39   */
    int max_points_per_centroid; /* 256: above k*this many points the training set is subsampled */
    int64_t seed;                /* 1234 */
} pq_kmeans_params;
void pq_kmeans_default_params(pq_kmeans_params* p);
int pq_kmeans_train(pq_index* index, int64_t k, const pq_kmeans_params* params, int64_t n, const float* x_host, float* centroids_out,
                    float* obj_out, int64_t obj_cap, int64_t* n_obj);

/* The same k-means iteration in steps, for training with the POINTS sharded over several GPUs (assignment
 * needs no exchange, the centroid update needs one all-reduce of [k,128] sums + [k] counts).  Every rank holds the k current
 * centroids in `index` and its own points on its device:
 *   pq_kmeans_set_centroids   load the initial centroids (host array; renormalised when spherical) into the index
 *   pq_kmeans_partial_device  assign the local points (index.search(x, 1)), leave per-centroid fp32 sums (added in index
 *                             order) and counts in caller-provided device buffers, return the local objective
 *   -- the caller all-reduces sums, counts and objective across ranks (NCCL) --
 *   pq_kmeans_finish_device   centroids = sums / counts, void clusters split as FAISS does, renormalise when spherical,
 *                             index.reset(); index.add(centroids).  Identical inputs give identical centroids on every rank.
 * pq_rand_perm is FAISS's rand_perm (mt19937 seeded with `seed`), which Clustering uses to subsample and to pick the
 * initial centroids.  Host mirror: proqa_b200/sharded_clustering.py. */
void pq_rand_perm(int64_t n, int64_t seed, int32_t* out);
int pq_kmeans_set_centroids(pq_index* index, int64_t k, const float* centroids_host, int spherical);
int pq_kmeans_partial_device(pq_index* index, int64_t k, int64_t n_local, const float* x_dev, float* sums_dev, int32_t* counts_dev,
                             double* objective_out);
int pq_kmeans_finish_device(pq_index* index, int64_t k, int64_t n_total, int spherical, const float* sums_dev, const int32_t* counts_dev,
                            float* centroids_out_host, int* nsplit_out);

/* Introspection (no device needed; used by the CPU tests of the host logic): the launch plan of one tensor-tier search of nq
 * queries, top-k, over ntotal local rows on a device with n_sms SMs.  out[0..7] = {epochs, CTA groups, query tiles per group
 * (base), groups owning base+1, max tiles per group, carry length K', padded queries, epilogue warp sets of the widest kernel
 * variant}; then per epoch 8 values {first row, end row, slices of the base+1 groups, slices of the base groups, slab capacity,
 * CTAs, slabs per query, epilogue warp sets of the variant the epoch runs (4 while the threshold is loose, 2 once it is tight)}. */
int pq_plan_describe(int64_t ntotal, int64_t nq, int64_t k, int n_sms, int64_t* out, int out_len);
/* The same for 512 <= k <= PQ_MAX_K (retrieval/trec_process.py:76 asks for k = 10000; BASELINE C5 for k = 1000).  That tensor-tier
 * path — thresholds from a row sample, one filter pass, a finalize kernel — is the default for such k on corpora of at least 64 k
 * rows (512 <= k <= 1024: with at least 256 queries and 2^20 rows); PROQA_B200_LARGEK=0 at pq_index_create switches it off.
 * out[0..15] = {applies, sample step,
 * k of the sample search, sample rows, finalize pool keys, sort length, sample epochs, pass slices (base+1 groups), pass slices
 * (base groups), slab capacity, pass CTAs, slabs per query, queries per batch, slab bytes of a batch, finalize shared-memory
 * bytes, carry length of the sample search}. */
int pq_plan_describe_large_k(int64_t ntotal, int64_t nq, int64_t k, int n_sms, int64_t* out, int out_len);

const char* pq_last_error(void);
/* ---- corpus row-sharded over several GPUs: threshold exchange over peer memory ----------------------------------------
 * north_star (4) shards the rows of IndexFlatIP (eval_retrieval.py:102-104) over the GPUs of a box.  Each shard searches its
 * rows; to keep the per-query work from being repeated blindly on every shard, the shards tell each other after every epoch
 * how good their k-th (and ceil(k/n)-th) best local score is, by plain stores into each other's HBM over NVLink, and each
 * admits only what can still reach the global top-k.  A shard's result list then holds the local rows that can be in the
 * global top-k (at most k, best first, padded with -1); pq_merge_shard_results of all lists is the exact global result.
 *   pq_index_share_alloc    allocate this shard's mailbox (device memory of the index's GPU); returns its address
 *   pq_index_share_connect  the n mailbox addresses as seen from THIS process/device, in shard order (own one included):
 *                           peer-enabled pointers within one process (pq_enable_peer_access), CUDA IPC mappings across
 *                           processes (pq_ipc_export / pq_ipc_open)
 *   pq_index_share_begin    before every search: a sequence number, the same on every shard of that search
 *   pq_index_share_close    back to independent shards
 * The exchange needs no barrier and cannot deadlock: every value carries the sequence number of its search, readers wait a
 * bounded time (PROQA_B200_SHARE_WAIT_US, default 200) and then use what has arrived. */
int pq_index_share_alloc(pq_index* idx, int n_shards, int shard, int64_t max_queries_per_search, void** mailbox_dev_out, int64_t* bytes_out);
int pq_index_share_connect(pq_index* idx, const void* const* mailboxes_dev);
int pq_index_share_begin(pq_index* idx, uint32_t seq);
int pq_index_share_close(pq_index* idx);
/* The filter's error bound uses max |x|^2 and max |x - bf16(x)|^2 over the rows: shards that exchange thresholds must all use
 * the maxima over the WHOLE corpus (all-reduce the two floats once after add()).  set folds by maximum. */
int pq_index_get_bound_scalars(const pq_index* idx, float* out2);
int pq_index_set_bound_scalars(pq_index* idx, float max_norm2, float max_resid2);
/* CUDA IPC plumbing for the one-process-per-GPU layout (64-byte handles), and peer access for one process driving several GPUs. */
int pq_ipc_export(const void* dev_ptr, void* handle_out64);
int pq_ipc_open(const void* handle64, int device, void** dev_ptr_out);
int pq_ipc_close(int device, void* dev_ptr);
int pq_enable_peer_access(int device, int peer_device);

/* ---- one process, several GPUs (proqa_b200/csrc/pq_multi.cu) ------------------------------------------------------------
 * The same four calls as above for an index spread over the GPUs of a box, driven from the single host process the reference
 * runs in (eval_retrieval.py:102-104): the faiss shim creates one of these instead of a pq_index when PROQA_B200_DEVICES lists
 * more than one device.  A first add() of more than 2^20 rows shards the rows contiguously over the devices (queries
 * replicated, per-shard search with threshold exchange over peer memory, lists gathered and merged on the first device);
 * a smaller first add() (k-means centroids, group_paras.py:49-51) replicates the rows and splits the QUERIES instead.
 * ids are insertion order either way; -1 / -+FLT_MAX padding as pq_index_search. */
typedef struct pq_multi pq_multi;
int pq_multi_create(int d, int metric, int n_devices, const int* devices, pq_multi** out);
void pq_multi_free(pq_multi* m);
int pq_multi_add(pq_multi* m, int64_t n, const float* x_host);
int pq_multi_search(pq_multi* m, int64_t nq, const float* xq_host, int64_t k, float* D_host, int64_t* I_host);
int pq_multi_reset(pq_multi* m);
int64_t pq_multi_ntotal(const pq_multi* m);
int pq_multi_n_devices(const pq_multi* m);
int pq_multi_mode(const pq_multi* m);            /* 0 empty, 1 rows sharded, 2 rows replicated */
int pq_multi_last_stats(const pq_multi* m, int64_t* out, int n);   /* sums over the shards ([6], [7]: maxima) */
pq_index* pq_multi_first_shard(pq_multi* m);     /* the shard on the first device (k-means training runs there) */

/* ---- the exchange step of the row-sharded search, fused with its merge, over peer memory (proqa_b200/csrc/pq_xchg.cu) ------
 * Replaces "NCCL all-gather of the per-shard lists, then pq_merge_shard_results on every GPU" (north_star (4)): every rank stores
 * the slice of its list that rank p is to merge straight into p's HBM, merges its own slice of the queries, and stores the merged
 * slice into every rank's result buffer.  One buffer per rank, mapped by all ranks (CUDA IPC across processes: pq_ipc_export /
 * pq_ipc_open; plain pointers + pq_enable_peer_access within one process).  rank = query_group * row_shards + row_shard.
 *   pq_xchg_bytes_needed  payload bytes for searches of up to nq queries at k results in that layout
 *   pq_xchg_create        allocate this rank's buffer (returns its device address and size for the IPC export)
 *   pq_xchg_connect       the world addresses, in rank order, as valid on THIS device
 *   pq_xchg_run           enqueue one exchange on a stream: local list [n_loc, k] -> the full result [nq, k] on every rank;
 *                         seq: the same strictly increasing number on every rank, one per search
 *   pq_xchg_check         after a synchronisation: PQ_ERR_CUDA if a peer never delivered (10 s timeout inside the kernels) */
typedef struct pq_xchg pq_xchg;
int64_t pq_xchg_bytes_needed(int64_t nq, int64_t k, int row_shards, int query_groups);
int pq_xchg_create(int device, int world, int rank, int64_t payload_bytes, pq_xchg** out, void** base_dev_out, int64_t* bytes_out);
int pq_xchg_connect(pq_xchg* x, const void* const* peer_bases_dev);
int pq_xchg_run(pq_xchg* x, int metric, int row_shards, int64_t nq, int64_t k, const float* D_local_dev, const int64_t* I_local_dev,
                float* D_out_dev, int64_t* I_out_dev, uint64_t seq, void* cuda_stream);
int pq_xchg_check(pq_xchg* x);
void pq_xchg_free(pq_xchg* x);

/* "proqa_b200 <version> sm_100a" */
const char* pq_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PROQA_B200_H_ */
