/* oracle/flat_oracle.c — CPU restatement of the reference's exact flat search.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's shared object; the product (proqa_b200/) never does and has no CPU path.
 *
 * PARITY UNPINNED: the arithmetic of ProQA's retrieval hot path lives in the third-party wheel
 * faiss-cpu==1.6.3 (/root/reference/requirements.txt:2), whose source is NOT under /root/reference and
 * which is not installable in this image (no network, no wheel).  The reference has no tests or golden
 * vectors for this path.  What follows restates FAISS 1.6.3's published algorithm
 * (IndexFlat::search -> knn_inner_product / knn_L2sqr in faiss/utils/distances.cpp, heaps in
 * faiss/utils/Heap.h) as called from
 *     /root/reference/retrieval/eval_retrieval.py:102-104   (IndexFlatIP, add, search k=80)
 *     /root/reference/retrieval/group_paras.py:35-51         (IndexFlatL2 / IndexFlatIP, search k=1)
 *     /root/reference/retrieval/trec_process.py:74-76        (IndexFlatIP, search k=10000)
 * and is pinned only against an fp64 brute force (tests/test_oracle.py) and hand-made fixtures
 * (tests/golden/).
 *
 * Two families of entry points:
 *   faiss_*  : FAISS semantics.  fp32 scores; per-query size-k binary heap, rows visited in ascending
 *              id, the heap root is replaced only on a strictly better score (so the lowest ids survive
 *              a k-th place tie); results best-first; unfilled slots id=-1, D=-FLT_MAX (IP) / +FLT_MAX
 *              (L2); L2 reported as squared distance, BLAS path  |x|^2+|y|^2-2<x,y>  clamped at 0
 *              (nq >= 20) or the direct sum of squared differences (nq < 20).
 *   engine_* : the GPU engine's *defined* score: eight fmaf chains of 16 dims each,
 *              tree-combined (engine_chain_dot below), ordered by (score desc, id asc).  The CUDA
 *              kernels must reproduce these bits exactly.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * binary heap keeping the k best; `worse(a,b)` = a is worse than b.  Root = worst of the kept set.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    float* val;
    int64_t* id;
    int64_t k;
    int is_l2; /* IP keeps largest (root = smallest); L2 keeps smallest (root = largest) */
} heap_t;

static inline int worse(const heap_t* h, float a, float b) { return h->is_l2 ? (a > b) : (a < b); }

static void heap_init(heap_t* h) {
    for (int64_t i = 0; i < h->k; ++i) {
        h->val[i] = h->is_l2 ? FLT_MAX : -FLT_MAX;
        h->id[i] = -1;
    }
}

/* Replace the root by (v, id) and restore the heap property by sifting down. */
static void heap_replace_root(heap_t* h, float v, int64_t id) {
    int64_t i = 0;
    const int64_t k = h->k;
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, c;
        if (l >= k) break;
        /* child that is the worse of the two becomes the candidate parent */
        c = (r < k && worse(h, h->val[r], h->val[l])) ? r : l;
        if (!worse(h, h->val[c], v)) break;
        h->val[i] = h->val[c];
        h->id[i] = h->id[c];
        i = c;
    }
    h->val[i] = v;
    h->id[i] = id;
}

/* FAISS: "if (C::cmp(simi[0], ip)) { heap_pop; heap_push; }"  — strict improvement over the root. */
static inline void heap_offer(heap_t* h, float v, int64_t id) {
    if (worse(h, h->val[0], v)) heap_replace_root(h, v, id);
}

/* Best-first output (FAISS heap_reorder): repeatedly extract the worst into the tail. */
static void heap_sort_best_first(heap_t* h) {
    int64_t n = h->k;
    heap_t tmp = *h;
    while (n > 1) {
        float v0 = h->val[0];
        int64_t i0 = h->id[0];
        float vl = h->val[n - 1];
        int64_t il = h->id[n - 1];
        tmp.k = n - 1;
        heap_replace_root(&tmp, vl, il);
        h->val[n - 1] = v0;
        h->id[n - 1] = i0;
        --n;
    }
    /* entries that were never filled keep id -1; FAISS moves them to the end (they are the worst) */
}

/* ------------------------------------------------------------------------------------------------
 * FAISS-semantics search
 * ------------------------------------------------------------------------------------------------ */
static float dot_f32(const float* a, const float* b, int d) {
    /* 8 partial sums, as a SIMD kernel would keep them (fvec_inner_product); order is unspecified upstream */
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    for (; i + 8 <= d; i += 8)
        for (int j = 0; j < 8; ++j) s[j] += a[i + j] * b[i + j];
    float t = ((s[0] + s[4]) + (s[2] + s[6])) + ((s[1] + s[5]) + (s[3] + s[7]));
    for (; i < d; ++i) t += a[i] * b[i];
    return t;
}
static float l2sqr_f32(const float* a, const float* b, int d) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    for (; i + 8 <= d; i += 8)
        for (int j = 0; j < 8; ++j) {
            const float t = a[i + j] - b[i + j];
            s[j] += t * t;
        }
    float t = ((s[0] + s[4]) + (s[2] + s[6])) + ((s[1] + s[5]) + (s[3] + s[7]));
    for (; i < d; ++i) {
        const float u = a[i] - b[i];
        t += u * u;
    }
    return t;
}

/* metric: 0 = inner product, 1 = squared L2.  Self-contained (no BLAS): scores by dot_f32. */
void faiss_flat_search(const float* xq, int64_t nq, const float* xb, int64_t nb, int d, int64_t k, int metric, float* D,
                       int64_t* I) {
    const int blas_path = nq >= 20; /* distance_compute_blas_threshold */
    float* nb2 = NULL;
    if (metric == 1 && blas_path) {
        nb2 = (float*)malloc(sizeof(float) * (size_t)(nb > 0 ? nb : 1));
        for (int64_t j = 0; j < nb; ++j) nb2[j] = dot_f32(xb + j * d, xb + j * d, d);
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t q = 0; q < nq; ++q) {
        heap_t h = {D + q * k, I + q * k, k, metric == 1};
        heap_init(&h);
        const float* x = xq + q * d;
        const float nq2 = (metric == 1 && blas_path) ? dot_f32(x, x, d) : 0.f;
        for (int64_t j = 0; j < nb; ++j) {
            float s;
            if (metric == 0) {
                s = dot_f32(x, xb + j * d, d);
            } else if (blas_path) {
                s = nq2 + nb2[j] - 2.f * dot_f32(x, xb + j * d, d);
                if (s < 0.f) s = 0.f;
            } else {
                s = l2sqr_f32(x, xb + j * d, d);
            }
            heap_offer(&h, s, j);
        }
        heap_sort_best_first(&h);
    }
    free(nb2);
}

/* BLAS-path building blocks (the caller runs sgemm on 4096 x 1024 blocks, e.g. numpy/OpenBLAS):
 * heaps live in (D, I), initialised by faiss_heaps_init, fed block by block, finished by reorder. */
void faiss_heaps_init(int64_t nq, int64_t k, int metric, float* D, int64_t* I) {
#pragma omp parallel for
    for (int64_t q = 0; q < nq; ++q) {
        heap_t h = {D + q * k, I + q * k, k, metric == 1};
        heap_init(&h);
    }
}
/* S is an [nq_blk, nb_blk] row-major block of scores for queries q0.. and rows j0..; for L2 the
 * caller passes inner products and the two norm vectors (as knn_L2sqr_blas does). */
void faiss_heaps_addn(int64_t nq_blk, int64_t q0, int64_t nb_blk, int64_t j0, const float* S, int64_t ldS, int64_t k, int metric,
                      const float* q_norms, const float* b_norms, float* D, int64_t* I) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t qi = 0; qi < nq_blk; ++qi) {
        heap_t h = {D + (q0 + qi) * k, I + (q0 + qi) * k, k, metric == 1};
        const float* s = S + qi * ldS;
        if (metric == 0) {
            for (int64_t j = 0; j < nb_blk; ++j) heap_offer(&h, s[j], j0 + j);
        } else {
            const float qn = q_norms[q0 + qi];
            for (int64_t j = 0; j < nb_blk; ++j) {
                float dis = qn + b_norms[j0 + j] - 2.f * s[j];
                if (dis < 0.f) dis = 0.f;
                heap_offer(&h, dis, j0 + j);
            }
        }
    }
}
void faiss_heaps_reorder(int64_t nq, int64_t k, int metric, float* D, int64_t* I) {
#pragma omp parallel for
    for (int64_t q = 0; q < nq; ++q) {
        heap_t h = {D + q * k, I + q * k, k, metric == 1};
        heap_sort_best_first(&h);
    }
}

/* ------------------------------------------------------------------------------------------------
 * engine-semantics search (bit-exact specification of the CUDA engine's output)
 * ------------------------------------------------------------------------------------------------ */
static inline uint32_t f32_ordered(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float ordered_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
/* The engine's defined score for d = 128 (proqa_b200/csrc/pq_common.cuh: engine_dot): eight fmaf chains of 16
 * consecutive dims each, combined as ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)).  For other d (tests only) the same rule
 * with chains of ceil(d/8) dims.  Compiled with -ffp-contract=off: the additions below are never fused. */
float engine_chain_dot(const float* row, const float* q, int d) {
    float p[8];
    const int len = (d + 7) / 8;
    for (int j = 0; j < 8; ++j) {
        float acc = 0.f;
        for (int i = j * len; i < (j + 1) * len && i < d; ++i) acc = fmaf(row[i], q[i], acc);
        p[j] = acc;
    }
    return ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
}
static int cmp_u64_desc(const void* a, const void* b) {
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? 1 : (x > y ? -1 : 0);
}

void engine_flat_search(const float* xq, int64_t nq, const float* xb, int64_t nb, int d, int64_t k, int metric, int64_t id_base,
                        float* D, int64_t* I) {
    float* nb2 = (float*)malloc(sizeof(float) * (size_t)(nb > 0 ? nb : 1));
    for (int64_t j = 0; j < nb; ++j) nb2[j] = engine_chain_dot(xb + j * d, xb + j * d, d);
    const float floor_thr = -3.4028232635611926e38f; /* fp32 successor of -FLT_MAX */
#pragma omp parallel
    {
        uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(nb > 0 ? nb : 1));
#pragma omp for schedule(dynamic, 2)
        for (int64_t q = 0; q < nq; ++q) {
            const float* x = xq + q * d;
            const float qn = engine_chain_dot(x, x, d);
            int64_t n = 0;
            for (int64_t j = 0; j < nb; ++j) {
                float s = engine_chain_dot(xb + j * d, x, d);
                if (metric == 1) s = fmaf(2.f, s, -nb2[j]);
                if (s >= floor_thr) keys[n++] = ((uint64_t)f32_ordered(s) << 32) | (uint64_t)(~(uint32_t)j);
            }
            qsort(keys, (size_t)n, sizeof(uint64_t), cmp_u64_desc);
            for (int64_t i = 0; i < k; ++i) {
                if (i < n) {
                    const float s = ordered_f32((uint32_t)(keys[i] >> 32));
                    const uint32_t row = ~(uint32_t)keys[i];
                    I[q * k + i] = (int64_t)row + id_base;
                    D[q * k + i] = metric == 1 ? fmaxf(0.f, qn - s) : s;
                } else {
                    I[q * k + i] = -1;
                    D[q * k + i] = metric == 1 ? FLT_MAX : -FLT_MAX;
                }
            }
        }
        free(keys);
    }
    free(nb2);
}

/* ------------------------------------------------------------------------------------------------
 * faiss.Clustering.train restated (FAISS 1.6.3 Clustering.cpp / utils/random.cpp) [upstream-memory: the source is the
 * un-vendored wheel; reference call site retrieval/group_paras.py:40-45].  Assignment = faiss_flat_search(x, 1) above.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt19937_t; /* std::mt19937 */
static void mt_seed(mt19937_t* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}
static uint32_t mt_next(mt19937_t* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            const uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
/* RandomGenerator::rand_int(max) = mt() % max ; rand_float() = mt() / float(mt.max()) */
static void faiss_rand_perm(int* perm, int64_t n, int64_t seed) {
    mt19937_t g;
    for (int64_t i = 0; i < n; ++i) perm[i] = (int)i;
    mt_seed(&g, (uint32_t)seed);
    for (int64_t i = 0; i + 1 < n; ++i) {
        const int64_t i2 = i + (int64_t)(mt_next(&g) % (uint32_t)(n - i));
        const int t = perm[i];
        perm[i] = perm[i2];
        perm[i2] = t;
    }
}
void faiss_rand_perm_export(int* perm, int64_t n, int64_t seed) { faiss_rand_perm(perm, n, seed); }

static void renorm_l2(int d, int64_t k, float* c) {
    for (int64_t i = 0; i < k; ++i) {
        float nr = 0.f;
        for (int j = 0; j < d; ++j) nr += c[i * d + j] * c[i * d + j];
        if (nr > 0.f) {
            const float inv = 1.0f / sqrtf(nr);
            for (int j = 0; j < d; ++j) c[i * d + j] *= inv;
        }
    }
}

/* Clustering.cpp at v1.6.3 (the release that added weighted / encoded k-means): compute_centroids + split_clusters.
 *   compute_centroids: per-centroid sums in point order (fp32); the counts are FLOATS (hassign[ci] += 1.0); the mean is a
 *                      multiplication by the reciprocal:  float norm = 1 / hassign[ci];  c[j] *= norm;
 *   split_clusters:    a void cluster takes a copy of a populated one, picked with probability (hassign[cj] - 1.0) / (n - k)
 *                      by RandomGenerator(1234), both perturbed by +-1/1024; the counts are split in half AS FLOATS
 *                      (hassign[ci] = hassign[cj] / 2), which matters for the probabilities of later splits. */
static int km_update_centroids(const float* x, float* centroids, const int64_t* assign, int d, int64_t k, int64_t n, float* hassign) {
    memset(centroids, 0, sizeof(float) * (size_t)(d * k));
    memset(hassign, 0, sizeof(float) * (size_t)k);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t ci = assign[i];
        float* c = centroids + ci * d;
        const float* xi = x + i * d;
        hassign[ci] += 1.0;
        for (int j = 0; j < d; ++j) c[j] += xi[j];
    }
    for (int64_t ci = 0; ci < k; ++ci) {
        if (hassign[ci] == 0) continue;
        const float norm = 1 / hassign[ci];
        float* c = centroids + ci * d;
        for (int j = 0; j < d; ++j) c[j] *= norm;
    }
    int nsplit = 0;
    const float EPS = 1.f / 1024.f;
    mt19937_t rng;
    mt_seed(&rng, 1234);
    for (int64_t ci = 0; ci < k; ++ci) {
        if (hassign[ci] != 0) continue;
        int64_t cj;
        for (cj = 0; 1; cj = (cj + 1) % k) {
            const float p = (hassign[cj] - 1.0) / (float)(n - k);
            const float r = mt_next(&rng) / 4294967296.0f; /* float(mt.max()) rounds to 2^32 */
            if (r < p) break;
        }
        memcpy(centroids + ci * d, centroids + cj * d, sizeof(float) * (size_t)d);
        for (int j = 0; j < d; ++j) {
            if (j % 2 == 0) {
                centroids[ci * d + j] *= 1 + EPS;
                centroids[cj * d + j] *= 1 - EPS;
            } else {
                centroids[ci * d + j] *= 1 - EPS;
                centroids[cj * d + j] *= 1 + EPS;
            }
        }
        hassign[ci] = hassign[cj] / 2;
        hassign[cj] -= hassign[ci];
        nsplit++;
    }
    return nsplit;
}

/* Returns the number of iterations run; obj_out[it] = sum of the k=1 distances, nsplit_out[it] = clusters split. */
int faiss_kmeans_train(const float* x_in, int64_t n_in, int d, int64_t k, int niter, int spherical, int max_points_per_centroid, int64_t seed,
                       int metric, float* centroids, float* obj_out, int* nsplit_out, int64_t* assign_last) {
    int64_t nx = n_in;
    const float* x = x_in;
    float* x_new = NULL;
    if (n_in > k * (int64_t)max_points_per_centroid) {
        int* perm = (int*)malloc(sizeof(int) * (size_t)n_in);
        faiss_rand_perm(perm, n_in, seed);
        nx = k * (int64_t)max_points_per_centroid;
        x_new = (float*)malloc(sizeof(float) * (size_t)(nx * d));
        for (int64_t i = 0; i < nx; ++i) memcpy(x_new + i * d, x_in + (int64_t)perm[i] * d, sizeof(float) * (size_t)d);
        x = x_new;
        free(perm);
    }
    if (nx == k) {
        memcpy(centroids, x, sizeof(float) * (size_t)(k * d));
        free(x_new);
        return 0;
    }
    int64_t* assign = (int64_t*)malloc(sizeof(int64_t) * (size_t)nx);
    float* dis = (float*)malloc(sizeof(float) * (size_t)nx);
    float* hassign = (float*)malloc(sizeof(float) * (size_t)k);
    int* perm = (int*)malloc(sizeof(int) * (size_t)nx);
    faiss_rand_perm(perm, nx, seed + 1);
    for (int64_t i = 0; i < k; ++i) memcpy(centroids + i * d, x + (int64_t)perm[i] * d, sizeof(float) * (size_t)d);
    if (spherical) renorm_l2(d, k, centroids);
    for (int it = 0; it < niter; ++it) {
        faiss_flat_search(x, nx, centroids, k, d, 1, metric, dis, assign);
        float err = 0;
        for (int64_t j = 0; j < nx; ++j) err += dis[j];
        if (obj_out) obj_out[it] = err;
        const int nsplit = km_update_centroids(x, centroids, assign, d, k, nx, hassign);
        if (nsplit_out) nsplit_out[it] = nsplit;
        if (spherical) renorm_l2(d, k, centroids);
    }
    if (assign_last) memcpy(assign_last, assign, sizeof(int64_t) * (size_t)(nx < n_in ? nx : n_in));
    free(assign);
    free(dis);
    free(hassign);
    free(perm);
    free(x_new);
    return niter;
}
