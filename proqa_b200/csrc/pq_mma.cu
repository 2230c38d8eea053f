// proqa_b200 — tensor-core tier: tcgen05 bf16 filter + exact fp32 rescoring with an exactness certificate.
//
// For large query batches the Q.C^T contraction of IndexFlat::search (reference call site
// retrieval/eval_retrieval.py:104; FAISS 1.6.3 runs it as sgemm_ on 4096x1024 blocks) is a dense
// GEMM with K = 128.  This file runs it on the 5th-generation tensor cores and never writes the
// score matrix:
//
//   pq_mma_filter_kernel  (one CTA per SM, warp-specialised, 576 threads; a 320-thread variant with two epilogue sets runs the
//                 tight-threshold epochs)
//     warp 0      TMA producer: corpus tiles (128 rows x 256 B bf16, SWIZZLE_128B; L2 also the tile's 128 row norms)
//                 through a 6-stage shared-memory ring, mbarrier completion; paced against the slowest CTA of its wave so that
//                 the CTA groups (which all stream the same rows for different queries) share every tile through L2 (pace_*)
//     warp 1      issues tcgen05.mma in the TS form: the stationary operand — up to 4 query tiles of 128 queries, written
//                 once per CTA into tensor memory with tcgen05.st — times the streamed corpus tile from shared memory;
//                 M=128 queries x N=128 rows x K=16, bf16 -> fp32, into two 128-column TMEM accumulators
//     warps 2-17  epilogue, four sets of four warps (one warp per TMEM lane quarter): set h drains columns 32h..32h+31 of
//                 every accumulator (tcgen05.ld 32 lanes x 32 columns -> registers); one thread owns one query (a TMEM
//                 lane), so the admission threshold is a register compare; survivors are appended to the query's private
//                 candidate slab in global memory (warp-uniform votes + predicated stores).  (Round 1 ran two sets of 64
//                 columns: per accumulator a warp then needed ~140 instructions behind a tcgen05.ld round trip — longer
//                 than the 540 cycles the tensor pipe takes for it, pipe 80 % busy; in the early epochs, where nearly every
//                 32-column chunk has a survivor, the two warps per scheduler were issue-bound at 2.4x the steady time.)
//   pq_epoch_select_kernel  folds the slabs of one epoch into a per-query carry list (top-K' by bf16 score, radix select)
//                 and raises the query's admission threshold to  A_k - 2E
//   pq_rescore_kernel       recomputes the carry list's scores with the engine's defined fp32 score, sorts, emits (D, I)
//                 and evaluates the exactness certificate
//   pq_k1_finalize_kernel   k = 1 (k-means assignment): one warp per query rescoring the few groups within 2E of the best
//
// Exactness: |bf16 score - fp32 score| <= E_q = eps * |q| * max|c|.  Every row of the
// true top-k has approximate score >= A_k - 2E (A_k = k-th best approximate score), so filtering at
// any threshold <= A_k(seen so far) - 2E loses nothing.  The only lossy step is truncating the carry
// list at K' entries; the largest score ever truncated is tracked ("dropmax") and the certificate is
// dropmax < A_k(final) - 2E.  Queries that fail it (or overflow a slab) are re-run by the fp32 scan.
#include "pq_common.cuh"
#include "pq_host.h"
#include "pq_plan.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

namespace pq {

constexpr int kBM = 128;            // queries per M tile (TMEM lanes)
constexpr int kBN = 128;            // corpus rows per B tile (one TMA stage)
constexpr int kMaxEpiSets = 4;      // epilogue warp sets of the widest kernel variant (pq_plan.h: kPlanSetsLoose)
constexpr int kStages = 6;          // B ring depth
constexpr int kAccBufs = 2;         // TMEM accumulators of kBN = 128 columns (UMMA N = 128): one per (B tile, M tile), double-buffered
constexpr int kTmemACol = kAccBufs * kBN;      // first TMEM column of the stationary query operand
constexpr int kPanelBytes = 128 * 128;         // 128 rows x 128 B (64 bf16): one swizzle-128B K panel
constexpr int kTileBytes = 2 * kPanelBytes;    // K = 128 -> two panels, 32 KB
constexpr int kStageBytes = kTileBytes + 1024; // + the tile's 128 squared row norms (L2 only), padded to keep 1024-B alignment
constexpr int kMaxMTiles = 4;       // 4 x 64 TMEM columns of bf16 queries + 2 x 128 columns of accumulators = 512
static_assert(kBM == kPlanQueryTile && kBN == kPlanTileRows && kMaxMTiles == kPlanMaxMTiles && kMaxEpiSets == kPlanSetsLoose && kPlanSetsTight == 2, "pq_plan.h must describe this kernel");


struct MmaCtrl {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[kAccBufs];
    uint64_t tmem_empty[kAccBufs];
    uint32_t tmem_base;
    uint32_t pad;
};

struct MmaParams {
    const uint16_t* q_bf16;  // [nq_pad][128] bf16 queries (rows beyond nq are zero)
    uint64_t* cand_keys;     // [nq_pad][n_sub][cap]
    uint32_t* cand_cnt;      // [nq_pad][n_sub], zeroed before the launch
    const float* row_norms;  // [ntotal] squared row norms (L2 only)
    const float* thr;        // [nq_pad]
    const float* two_e;      // [nq_pad]
    long long row_begin;     // multiple of 128
    long long row_end;       // exclusive, <= ntotal
    int n_mtiles;            // total query tiles
    // query tiles are dealt to `n_groups` CTA groups as evenly as possible: the first `rem` groups own base+1
    // tiles and are given s1 row slices each, the others own `base` tiles and s0 slices (s ~ proportional to
    // the tiles owned, so every CTA carries the same number of (query tile x row tile) products)
    int base, rem, s1, s0;
    int cap;
    int n_sub;               // candidate slabs per query = max(s1, s0) * sets (one per epilogue warp set)
    int sets;                // epilogue warp sets of the kernel variant launched (2 or 4)
    int k1_adapt;
    // second attempt at an epoch ("repair", launched after the last epoch and only when some slab overflowed): only the
    // queries whose slabs overflowed in that epoch (redo[q] & redo_bit) take part, with their final thresholds
    const uint32_t* redo;      // [nq_pad] or null
    uint32_t redo_bit;
    // pacing of the TMA producers (pace_* below): null, or one zeroed arrival counter per block of 2^pace_shift row tiles
    uint32_t* pace;          // [cohorts][pace_blocks]
    int pace_shift;
    int pace_blocks;
    int pace_cohort;         // CTAs per cohort: blockIdx / pace_cohort = the wave a CTA runs in (a single-wave grid is one cohort)
};

// D[tmem] (+)= A[tmem] * B[smem desc]^T: the stationary operand (queries) is read from tensor memory, so shared
// memory only has to feed the streamed corpus tile (64 B/clk instead of 128 B/clk for the SS form).
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 TMEM lanes (one per thread of the warp) x 32 consecutive 32-bit columns <- 32 registers.
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 1-D bulk copy global -> shared, completion counted on an mbarrier (same mechanism as the tensor-map loads).
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// L2: turn 32 inner products into the engine's ranking score 2<q,x> - |x|^2 (larger = closer).
__device__ __forceinline__ void mma_apply_l2_bias(float (&v)[32], const float4* norms) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 n = norms[c];
        v[4 * c + 0] = fmaf(2.f, v[4 * c + 0], -n.x);
        v[4 * c + 1] = fmaf(2.f, v[4 * c + 1], -n.y);
        v[4 * c + 2] = fmaf(2.f, v[4 * c + 2], -n.z);
        v[4 * c + 3] = fmaf(2.f, v[4 * c + 3], -n.w);
    }
}

// Threshold filter over 32 accumulator columns held in registers (one thread = one query).
//
// Control flow is kept WARP-UNIFORM: votes decide whether any lane of the warp has a survivor in the chunk, then in each
// group of four columns, and the appends themselves are predicated stores — no divergent branch anywhere.  (Earlier
// versions located survivors through per-lane nested branches, first out of line, then inline: in the early epochs,
// where the threshold is still loose and nearly every chunk has a survivor in some lane, a third of all issued
// instructions were BRA/BSSY/BSYNC and the tensor pipe sat at 5-20 %.)  A full slab keeps counting (overflow is detected
// from the count) and overwrites its last slot.
__device__ __forceinline__ void mma_append_if(bool hit, uint64_t* slab, uint32_t& cnt, uint32_t cap, float score, uint32_t row) {
    const uint64_t key = make_key(score, row);
    uint64_t* dst = slab + min(cnt, cap - 1);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %0, 0;\n\t"
        "@p st.global.u64 [%1], %2;\n\t}"
        ::"r"((uint32_t)hit), "l"(dst), "l"(key)
        : "memory");
    cnt += hit ? 1u : 0u;
}
// rows past the end of the epoch (TMA zero fill in the last tile, stale norms) must not look like scores
__device__ __forceinline__ void mma_mask_tail(float (&v)[32], uint32_t base_row, uint32_t row_end32) {
    const int left = (int)(row_end32 - base_row);  // rows of this chunk inside the epoch (negative: none); one uniform value,
    if (left < 32) {                               // compared with immediates — warp-uniform, last tile only
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (c >= left) v[c] = -INFINITY;
    }
}

__device__ __forceinline__ void mma_filter32(float (&v)[32], float th, uint32_t base_row, uint32_t row_end32, uint64_t* slab, uint32_t& cnt,
                                             uint32_t cap) {
    // (the caller has masked the rows past the end of the epoch: mma_mask_tail)
    float g4[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) g4[g] = fmaxf(fmaxf(v[4 * g], v[4 * g + 1]), fmaxf(v[4 * g + 2], v[4 * g + 3]));
    const float mx = fmaxf(fmaxf(fmaxf(g4[0], g4[1]), fmaxf(g4[2], g4[3])), fmaxf(fmaxf(g4[4], g4[5]), fmaxf(g4[6], g4[7])));
    if (__any_sync(0xffffffffu, mx >= th)) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            if (__any_sync(0xffffffffu, g4[g] >= th)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mma_append_if(v[4 * g + i] >= th, slab, cnt, cap, v[4 * g + i], base_row + 4 * g + i);
            }
        }
    }
}

// Maximum of all the accumulator values a thread holds for one accumulator (kChunks x 32 columns), as a tree of 3-input
// maxima (FMNMX3): 16 instructions per 32 values.  The steady state of a search — threshold tight, a survivor in one
// accumulator out of ten — is then ONE compare and ONE vote per accumulator and warp; only a warp that has a survivor
// goes on to locate it (mma_filter32).  The epilogue's instruction count is what this kernel's power budget is spent on
// besides the MMAs themselves (the kernel runs at the power cap: every instruction saved is clock gained).
__device__ __forceinline__ float mma_max3(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    float d;  // one FMNMX3; as opaque asm the compiler cannot re-associate the tree into the two-input maxima of the slow path
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}
template <int N>
__device__ __forceinline__ float mma_max_reduce(float (&a)[N]) {
    if constexpr (N == 1) {
        return a[0];
    } else if constexpr (N == 2) {
        return fmaxf(a[0], a[1]);
    } else {
        constexpr int T = N / 3, R = N % 3;
        float b[T + R];
#pragma unroll
        for (int i = 0; i < T; ++i) b[i] = mma_max3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
#pragma unroll
        for (int r = 0; r < R; ++r) b[T + r] = a[3 * T + r];
        return mma_max_reduce<T + R>(b);
    }
}
template <int kChunks>
__device__ __forceinline__ float mma_max_all(float (&v)[kChunks][32]) {
    float a[kChunks * 32];
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) a[c * 32 + i] = v[c][i];
    return mma_max_reduce<kChunks * 32>(a);
}

// k = 1 variant (k-means assignment: few rows per thread, so survivors are common).  The thread keeps a running maximum;
// a score more than 2E below any score already seen cannot be the best row.  Survivors are recorded per group of four
// rows (key = group maximum, first row of the group) — 8 predicated appends per chunk; the finalize kernel rescores the
// four rows of the few groups that end up within 2E of the overall maximum.
__device__ __forceinline__ void mma_filter32_k1(float (&v)[32], float& thr, float two_e, uint32_t base_row, uint32_t row_end32, uint64_t* slab,
                                                uint32_t& cnt, uint32_t cap) {
    mma_mask_tail(v, base_row, row_end32);
    float g4[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) g4[g] = fmaxf(fmaxf(v[4 * g], v[4 * g + 1]), fmaxf(v[4 * g + 2], v[4 * g + 3]));
    const float mx = fmaxf(fmaxf(fmaxf(g4[0], g4[1]), fmaxf(g4[2], g4[3])), fmaxf(fmaxf(g4[4], g4[5]), fmaxf(g4[6], g4[7])));
    thr = fmaxf(thr, mx - two_e);
    const float th = thr;
    if (__any_sync(0xffffffffu, mx >= th)) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            if (__any_sync(0xffffffffu, g4[g] >= th)) mma_append_if(g4[g] >= th, slab, cnt, cap, g4[g], base_row + 4 * g);
        }
    }
}

// Pacing of the TMA producers.  Every CTA group streams the whole epoch's rows (each group owns other queries); the groups
// advance at the same rate by construction, and as long as all of them are within an L2's worth of rows of each other a tile
// comes from DRAM once and from L2 for the other groups.  Left alone the CTAs drift apart (a 0.5 % difference in speed over
// an 11 ms launch is 25 MB of rows): ncu showed 14.1 GB of DRAM reads for the 4.3 GB of rows of the last C2 epoch.  So the row
// tiles of the epoch are cut into blocks of 2^pace_shift tiles; a producer counts itself out of every block it has left, and
// does not start block b before every CTA of the grid has left block b - kPaceWindow.  Rate control only — no data depends on
// it — and bounded: a producer that waits longer than ~1 ms (some CTA is not running: the grid is not co-resident) stops pacing
// for the rest of the launch.  The host enables it only for grids of at most one CTA per SM.
constexpr int kPaceWindow = 3;
constexpr int kPaceMaxPolls = 4096;
__device__ __forceinline__ void pace_leave(uint32_t* pace, int from, int to) {
    for (int b = from; b < to; ++b) atomicAdd(pace + b, 1u);
}
__device__ __forceinline__ bool pace_wait(const uint32_t* cnt, uint32_t n_ctas) {
    for (int i = 0; i < kPaceMaxPolls; ++i) {
        if (*reinterpret_cast<const volatile uint32_t*>(cnt) >= n_ctas) return true;
        __nanosleep(256);
    }
    return false;
}

// Accumulator schedule shared by the MMA issuer and the epilogue.  For row tile t and query tile mi (< m) one accumulator
// of 128 columns is produced, sequence number j = t*m + mi, in TMEM buffer j & 1 (use number j >> 1).  The epilogue
// warps form kEpiSets sets of four (one warp per TMEM lane quarter); set h drains columns h*kSubN .. (h+1)*kSubN-1 (those rows
// of the tile).  Every warp consumes every accumulator, in order — an mbarrier parity wait can only tell adjacent phases apart,
// so a consumer must never be able to run two uses ahead of a buffer (a round-robin over four buffers whose consumers
// changed from use to use aliased and hung).  N = 128 per instruction measured ~5 % faster end to end than two N = 64
// accumulators per tile (half the instructions, half the reads of the stationary operand from tensor memory).
template <int M_TILES, bool kL2, bool kK1, int kEpiSets>
__global__ void __launch_bounds__((2 + 4 * kEpiSets) * 32, 1)
pq_mma_filter_kernel(const __grid_constant__ CUtensorMap tmap_c, const MmaParams p) {
    constexpr int kEpiWarps = 4 * kEpiSets;        // 4 TMEM lane quarters x kEpiSets column sets
    constexpr int kSubN = kBN / kEpiSets;          // accumulator columns (corpus rows) one epilogue warp set drains
    constexpr int kSubChunks = kSubN / 32;         // 32-column register chunks per warp and accumulator
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_b = smem;                                        // kStages x (32 KB tile + norms)
    MmaCtrl* ctrl = reinterpret_cast<MmaCtrl*>(smem_b + (size_t)kStages * kStageBytes);

    // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role code below (barrier
    // addresses, descriptors, loop counters) in uniform registers — the tcgen05.mma issue path is a handful of uniform
    // instructions instead of a per-lane "waterfall" loop around every UTCHMMA
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    // which query tiles and which row slice this CTA owns
    int group, slice, n_slices;
    {
        const int x = blockIdx.x, big = p.rem * p.s1;
        if (x < big) {
            group = x / p.s1;
            slice = x - group * p.s1;
            n_slices = p.s1;
        } else {
            group = p.rem + (x - big) / p.s0;
            slice = (x - big) % p.s0;
            n_slices = p.s0;
        }
    }
    const int mt0 = group * p.base + min(group, p.rem);
    const int m = min(M_TILES, p.base + (group < p.rem ? 1 : 0));

    // Row tiles are dealt to the slices round-robin (slice s takes tiles s, s + n_slices, ...): a run of consecutive rows that
    // all beat the threshold — a topic cluster in a corpus stored in document order — spreads over every slab of the query
    // instead of flooding one.
    const long long total_tiles = (p.row_end - p.row_begin + kBN - 1) / kBN;
    const int ntiles = total_tiles > slice ? (int)((total_tiles - slice + n_slices - 1) / n_slices) : 0;
    const long long row0 = p.row_begin + (long long)slice * kBN;
    const long long tile_stride = (long long)n_slices * kBN;  // rows between two consecutive tiles of this CTA

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_c);
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(&ctrl->full[s], 1);
                // L2: the epilogue reads the stage's row norms, so the stage is only free once it has let go as well
                mbar_init(&ctrl->empty[s], kL2 ? 1 + kEpiWarps : 1);
            }
            for (int b = 0; b < kAccBufs; ++b) {
                mbar_init(&ctrl->tmem_full[b], 1);
                mbar_init(&ctrl->tmem_empty[b], kEpiWarps);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<512>(&ctrl->tmem_base);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = ctrl->tmem_base;

    const int e = warp - 2;
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int set = e >> 2;         // epilogue warp set (column range of every accumulator); meaningless for warps 0 and 1
    if (warp >= 2) {
        // ---- queries -> tensor memory: lane = query, 64 columns of packed bf16 pairs per query tile ----
        for (int mi = set; mi < m; mi += kEpiSets) {
            const uint4* src = reinterpret_cast<const uint4*>(p.q_bf16 + ((size_t)(mt0 + mi) * kBM + quarter * 32 + lane) * kDim);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(kTmemACol + mi * 64);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 w = __ldg(src + half * 8 + j);
                    r[4 * j + 0] = w.x;
                    r[4 * j + 1] = w.y;
                    r[4 * j + 2] = w.z;
                    r[4 * j + 3] = w.w;
                }
                tmem_st_32x32(taddr + half * 32, r);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    if (warp == 0) {
        // ===================== TMA producer =====================
        int pace_blk = 0;              // block of row tiles this producer is in (pacing, see pace_leave / pace_wait)
        bool pacing = p.pace != nullptr;
        uint32_t* pace = nullptr;      // this CTA's cohort: the CTAs of its wave, which start together and advance together
        uint32_t pace_n = 0;
        if (p.pace != nullptr) {
            const int cohort = (int)blockIdx.x / p.pace_cohort;
            pace = p.pace + (size_t)cohort * p.pace_blocks;
            pace_n = (uint32_t)min(p.pace_cohort, (int)gridDim.x - cohort * p.pace_cohort);
        }
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % kStages;
            const uint32_t ph = (uint32_t)(t / kStages) & 1u;
            if (p.pace != nullptr) {
                const int blk = (int)(((long long)slice + (long long)t * n_slices) >> p.pace_shift);
                if (blk != pace_blk) {  // warp-uniform
                    if (lane == 0) {
                        pace_leave(pace, pace_blk, blk);
                        if (pacing && blk >= kPaceWindow) pacing = pace_wait(pace + (blk - kPaceWindow), pace_n);
                    }
                    pace_blk = blk;
                    __syncwarp();
                }
            }
            mbar_wait(&ctrl->empty[s], ph ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(&ctrl->full[s], kTileBytes + (kL2 ? kBN * 4 : 0));
                const int y = (int)(row0 + (long long)t * tile_stride);
#pragma unroll
                for (int pnl = 0; pnl < 2; ++pnl)
                    tma_load_2d(smem_b + (size_t)s * kStageBytes + pnl * kPanelBytes, &tmap_c, pnl * 64, y, &ctrl->full[s]);
                if (kL2) bulk_load_1d(smem_b + (size_t)s * kStageBytes + kTileBytes, p.row_norms + y, kBN * 4, &ctrl->full[s]);
            }
            __syncwarp();
        }
        if (p.pace != nullptr && lane == 0) pace_leave(pace, pace_blk, p.pace_blocks);  // (a CTA without tiles leaves them all)
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
        constexpr uint32_t idesc = umma_idesc_bf16(kBM, kBN);
        const uint32_t b_addr = smem_u32(smem_b);
        // all 512 columns are ours (one CTA per SM): the allocation can only start at lane 0, column 0
        if (tmem_base != 0) __trap();
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % kStages;
            const uint32_t ph = (uint32_t)(t / kStages) & 1u;
            mbar_wait(&ctrl->full[s], ph);
            tc_fence_after_sync();
            for (int mi = 0; mi < m; ++mi) {
                const int j = t * m + mi;
                const int b = j & 1;
                const uint32_t aph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(&ctrl->tmem_empty[b], aph ^ 1u);
                tc_fence_after_sync();
                if (elect_one()) {
                    const uint32_t a_tmem = (uint32_t)(kTmemACol + mi * 64);
                    const uint32_t tile_addr = b_addr + (uint32_t)s * kStageBytes;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        // K step ks is 32 B into the 128-B row of panel ks >> 2
                        const uint32_t koff = (uint32_t)(ks >> 2) * kPanelBytes + (uint32_t)(ks & 3) * 32u;
                        umma_bf16_ts((uint32_t)(b * kBN), a_tmem + (uint32_t)ks * 8u, umma_desc_k128(tile_addr + koff), idesc, ks > 0 ? 1u : 0u);
                    }
                    umma_commit(&ctrl->tmem_full[b]);
                    if (mi == m - 1) umma_commit(&ctrl->empty[s]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: threshold filter =====================
        // per (query tile, thread) state in shared memory — threshold, 2E, slab fill — so the loop over query tiles
        // below needs no unrolling (each thread reads and writes only its own slots: no synchronisation)
        const int tid_e = (int)threadIdx.x - 64;
        constexpr int kEpiThreads = kEpiWarps * 32;
        float* s_thr = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctrl) + 256);
        float* s_2e = s_thr + kMaxMTiles * kEpiWarps * 32;
        uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_2e + kMaxMTiles * kEpiWarps * 32);
        const int lane_q = quarter * 32 + lane;
        const int sub = slice * kEpiSets + set;
        for (int mi = 0; mi < m; ++mi) {
            const size_t q = (size_t)(mt0 + mi) * kBM + lane_q;
            s_thr[mi * kEpiThreads + tid_e] = (p.redo == nullptr || (p.redo[q] & p.redo_bit) != 0u) ? p.thr[q] : INFINITY;
            s_2e[mi * kEpiThreads + tid_e] = p.two_e[q];
            s_cnt[mi * kEpiThreads + tid_e] = 0;
        }
        const uint32_t row_end32 = (uint32_t)p.row_end;
        const uint32_t cap = (uint32_t)p.cap;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * kSubN);
        uint32_t base_row = (uint32_t)(row0 + set * kSubN);  // first row of this warp set's columns in tile t
        uint32_t j = 0;                                      // accumulator sequence number t * m + mi
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t, base_row += (uint32_t)tile_stride) {
            const int s = t % kStages;
            const float4* norms = reinterpret_cast<const float4*>(smem_b + (size_t)s * kStageBytes + kTileBytes) + set * (kSubN / 4);
            if (kL2) mbar_wait(&ctrl->full[s], (uint32_t)(t / kStages) & 1u);  // already complete (the MMAs needed it): acquire only
            const bool tail = base_row + kSubN > row_end32;  // warp-uniform: rows past the end of the epoch (last tile only)
#pragma unroll 1
            for (int mi = 0; mi < m; ++mi, ++j) {
                const uint32_t b = j & 1u;
                mbar_wait(&ctrl->tmem_full[b], (j >> 1) & 1u);
                tc_fence_after_sync();
                float v[kSubChunks][32];
#pragma unroll
                for (int c = 0; c < kSubChunks; ++c) tmem_ld_32x32(taddr0 + b * kBN + 32 * c, v[c]);
                tmem_ld_wait();
                // The accumulator is in registers now: hand the TMEM buffer back before filtering.
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctrl->tmem_empty[b]);
                if (kL2) {
#pragma unroll
                    for (int c = 0; c < kSubChunks; ++c) mma_apply_l2_bias(v[c], norms + 8 * c);
                    if (mi == m - 1) {  // last read of this stage's norms by this warp
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&ctrl->empty[s]);
                    }
                }
                const int slot = mi * kEpiThreads + tid_e;
                if (kK1) {
                    uint64_t* slab = p.cand_keys + (((size_t)(mt0 + mi) * kBM + lane_q) * p.n_sub + sub) * (size_t)cap;
                    uint32_t cnt = s_cnt[slot];
                    const uint32_t cnt_in = cnt;
                    float th = s_thr[slot];
                    const float two_e = s_2e[slot];
#pragma unroll
                    for (int c = 0; c < kSubChunks; ++c) mma_filter32_k1(v[c], th, two_e, base_row + 32 * c, row_end32, slab, cnt, cap);
                    s_thr[slot] = th;
                    if (cnt != cnt_in) s_cnt[slot] = cnt;
                } else {
                    const float th = s_thr[slot];
                    if (tail) {
#pragma unroll
                        for (int c = 0; c < kSubChunks; ++c) mma_mask_tail(v[c], base_row + 32 * c, row_end32);
                    }
                    // fast path (tight-threshold variant): no thread of the warp has a survivor anywhere in this accumulator.
                    // The loose-threshold variant nearly always has one, so it goes straight to the per-chunk votes.
                    if (kEpiSets == kMaxEpiSets || __any_sync(0xffffffffu, mma_max_all<kSubChunks>(v) >= th)) {
                        uint64_t* slab = p.cand_keys + (((size_t)(mt0 + mi) * kBM + lane_q) * p.n_sub + sub) * (size_t)cap;
                        uint32_t cnt = s_cnt[slot];
                        const uint32_t cnt_in = cnt;
                        uint32_t row = base_row;
#if defined(__CUDA_ARCH__)
                        asm volatile("" : "+r"(row));  // keeps the 64 row numbers of the keys from being precomputed once per tile
#endif
#pragma unroll
                        for (int c = 0; c < kSubChunks; ++c) mma_filter32(v[c], th, row + 32 * c, row_end32, slab, cnt, cap);
                        if (cnt != cnt_in) s_cnt[slot] = cnt;
                    }
                }
            }
        }
        for (int mi = 0; mi < m; ++mi) {
            const size_t q = (size_t)(mt0 + mi) * kBM + lane_q;
            p.cand_cnt[q * p.n_sub + sub] = s_cnt[mi * kEpiThreads + tid_e];
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// per-query state
// ------------------------------------------------------------------------------------------------
struct QState {
    float* thr;          // [nq_pad] admission threshold (bf16-score domain)
    float* two_e;        // [nq_pad]
    float* dropmax;      // [nq_pad]
    uint32_t* overflow;  // [nq_pad]
    uint64_t* carry;     // [nq_pad][kp] candidates kept between epochs: a compacted prefix, zero keys after it
    uint32_t* redo;      // [nq_pad] bit e set: the slabs of epoch e overflowed for this query (repaired after the last epoch)
    uint32_t* counters;  // [0] OR of all redo masks, [1] (query, epoch) overflows, [2] failed queries, [3] threshold exchanges
                         //     in which every peer shard's values had arrived
};

// Error bound of the bf16 filter, per query:
//   |sum q^ c^ - sum q c| <= |q^ - q| |c^| + |q| |c^ - c|                      (Cauchy-Schwarz on the two rounding residuals)
//   + tensor-core accumulation (products of bf16 are exact in fp32; <= 128 truncating adds)   <= 2^-14 |q^| |c^|
//   + rounding of the engine's own fp32 score (19 roundings)                                    <= 2.4e-6 |q| |c|
// with |q^ - q| measured per query, max |c^ - c| and max |c| measured over the corpus at add().
__global__ void pq_mma_init_state_kernel(QState st, const float* __restrict__ q_norm2, const float* __restrict__ q_resid2,
                                         const uint8_t* __restrict__ q_bad, int nq, int nq_pad, int kp, float max_norm2, float max_resid2,
                                         int metric) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq_pad) return;
    const bool live = q < nq && !q_bad[q];
    float e2 = 0.f;
    if (live) {
        const float Q = sqrtf(q_norm2[q]) * 1.000001f, rq = sqrtf(q_resid2[q]) * 1.000001f;
        const float C = sqrtf(max_norm2) * 1.000001f, rc = sqrtf(max_resid2) * 1.000001f;
        const float E = 1.0002f * (rq * (C + rc) + Q * rc) + 6.2e-5f * (Q + rq) * (C + rc) + 2.4e-6f * Q * C;
        e2 = 2.f * E;
        // L2 ranks by 2<q,x> - |x|^2: twice the inner-product error, plus the fp32 rounding of that fmaf on both sides
        if (metric == kMetricL2) e2 = 2.f * e2 + 4.8e-7f * (2.f * Q * C + max_norm2);
    }
    st.two_e[q] = e2;
    st.thr[q] = live ? PQ_THR_FLOOR : INFINITY;
    st.dropmax[q] = -INFINITY;
    st.overflow[q] = (live && isfinite(e2)) ? 0u : (q < nq ? 1u : 0u);
    st.redo[q] = 0u;
}

// Cross-shard threshold exchange over peer memory (NVLink): corpus row-sharded over n GPUs (north_star (4)).  After every
// epoch each shard writes, for every query, two bf16-domain scores into the mailbox of every shard:
//     a = its k-th best local score         (k local rows score at least a)
//     b = its ceil(k/n)-th best local score (that many local rows score at least b)
// The k-th best score over ALL rows is at least max over shards of a, and at least min over shards of b (n x ceil(k/n) >= k rows
// reach it) — on exchangeable rows the latter is the k-th best of n times as many rows as one shard has seen.  Every word
// carries the tag of the search it belongs to, so a reader uses only values of its own search; values of any epoch of that
// search are valid lower bounds, hence no barrier: the fold kernel waits a bounded time for the peers' epoch headers and
// then takes what is there.  Results never depend on what arrives — only how much the filter admits.
constexpr int kShareMaxPeers = 16;
struct ShareParams {
    int n, rank;       // row shards (0 or 1: exchange off), this shard
    int kr;            // ceil(k / n)
    uint32_t tag;      // (search sequence << 4) | query batch
    int cap_q;         // queries per mailbox slot
    int wait_ns;       // longest wait for the peers' headers
    uint64_t* peer[kShareMaxPeers];   // peer[p]: shard p's mailbox — [n][cap_q][2] value words, then [n] header words
};
__device__ __forceinline__ uint64_t share_word(uint32_t tag, float v) { return ((uint64_t)tag << 32) | (uint64_t)__float_as_uint(v); }
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void share_publish(const ShareParams& sh, int q, float a, float b, int lane) {
    if (lane < sh.n) {
        uint64_t* dst = sh.peer[lane] + ((size_t)sh.rank * sh.cap_q + q) * 2;
        dst[0] = share_word(sh.tag, a);
        dst[1] = share_word(sh.tag, b);
    }
}

// Announces this shard's values of `epoch` (the select kernel that wrote them precedes this launch on the stream), waits —
// bounded — for the peers' announcements, then raises every query's threshold to what the shards know together.
__global__ void __launch_bounds__(256) pq_share_fold_kernel(const ShareParams sh, QState st, int nq, int epoch) {
    uint64_t* mine = sh.peer[sh.rank];
    uint64_t* hdr = mine + (size_t)sh.n * sh.cap_q * 2;
    const uint64_t want = ((uint64_t)sh.tag << 32) | (uint64_t)(uint32_t)(epoch + 1);
    if (blockIdx.x == 0 && (int)threadIdx.x < sh.n) {
        __threadfence_system();
        st_volatile_u64(sh.peer[threadIdx.x] + (size_t)sh.n * sh.cap_q * 2 + sh.rank, want);
    }
    if (threadIdx.x == 0) {
        const uint64_t t0 = global_timer_ns();
        const uint32_t full = sh.n >= 32 ? 0xffffffffu : ((1u << sh.n) - 1u);
        uint32_t fresh;
        do {
            fresh = 0u;
            for (int g = 0; g < sh.n; ++g) {
                const uint64_t v = ld_volatile_u64(hdr + g);
                if ((uint32_t)(v >> 32) == sh.tag && (uint32_t)v >= (uint32_t)(epoch + 1)) fresh |= 1u << g;
            }
        } while (fresh != full && global_timer_ns() - t0 < (uint64_t)sh.wait_ns);
        if (blockIdx.x == 0 && fresh == full) atomicAdd(st.counters + 3, 1u);
    }
    __syncthreads();
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        float amax = -INFINITY, bmin = INFINITY;
        bool all = true;
        for (int g = 0; g < sh.n; ++g) {
            const uint64_t wa = ld_volatile_u64(mine + ((size_t)g * sh.cap_q + q) * 2);
            const uint64_t wb = ld_volatile_u64(mine + ((size_t)g * sh.cap_q + q) * 2 + 1);
            if ((uint32_t)(wa >> 32) == sh.tag && (uint32_t)(wb >> 32) == sh.tag) {
                amax = fmaxf(amax, __uint_as_float((uint32_t)wa));
                bmin = fminf(bmin, __uint_as_float((uint32_t)wb));
            } else {
                all = false;
            }
        }
        const float best = all ? fmaxf(amax, bmin) : amax;
        if (best > -INFINITY) st.thr[q] = fmaxf(st.thr[q], best - st.two_e[q]);
    }
}

struct EpochSelParams {
    QState st;
    const uint64_t* cand_keys;
    const uint32_t* cand_cnt;
    int n_sub, cap, kp, k, lmax;  // n_sub: slab stride per query; lmax: keys the CTA kernel's shared pool holds
    int nq;
    int base, rem, s1, s0, sets;  // the filter's grid shape: a query of a group with s row slices has sets x s slabs, the rest is unwritten
    int is_redo;                  // repair of this epoch: only queries with (st.redo & epoch_bit) take part
    int allow_redo;               // first attempt: a slab overflow asks for a repair instead of failing the query
    uint32_t epoch_bit;
    long long row_begin, row_end; // rows of the epoch (a repair replaces the carry entries the first attempt took from them)
    ShareParams share;
};

// slabs the filter kernel wrote for query q
__device__ __forceinline__ int sel_slabs_of_query(const EpochSelParams& p, int q) {
    const int mt = q / kPlanQueryTile;
    return p.sets * (mt < p.rem * (p.base + 1) ? p.s1 : p.s0);
}

// One CTA per query (any K'): carry  <-  top-K' of (carry U this epoch's slabs); threshold <- A_k - 2E.
// Slab counts are staged and prefix-summed in shared memory, candidates gathered coalesced (one warp per slab) into a
// pool next to the old carry; when the pool exceeds K' an MSB-first radix select over the 64-bit keys (8-bit digits, bytes
// common to all keys skipped) finds the K'-th largest key and the survivors are compacted — no sort of the pool.  Only the
// K' survivors are sorted at the end (the threshold is read off the k-th entry).
template <int kT = 256>
__device__ __forceinline__ uint64_t block_radix_select(const uint64_t* pool, int n, int want, int* hist, uint64_t* s_u64, int* s_int) {
    constexpr int kW = kT / 32;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // bytes on which all keys agree need no pass
    uint64_t a = ~0ull, o = 0ull;
    for (int i = t; i < n; i += kT) {
        const uint64_t k = pool[i];
        a &= k;
        o |= k;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        a &= __shfl_xor_sync(0xffffffffu, a, s);
        o |= __shfl_xor_sync(0xffffffffu, o, s);
    }
    if (lane == 0) {
        s_u64[warp] = a;
        s_u64[kW + warp] = o;
    }
    __syncthreads();
    a = s_u64[0];
    o = s_u64[kW];
#pragma unroll
    for (int w = 1; w < kW; ++w) {
        a &= s_u64[w];
        o |= s_u64[kW + w];
    }
    const uint64_t differ = a ^ o;  // bit set = keys disagree there
    uint64_t prefix = a & ~differ, mask = ~differ;  // agreed bits are part of the prefix already
    __syncthreads();
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (((differ >> shift) & 0xffull) == 0) continue;  // block-uniform
        const uint64_t dmask = (differ >> shift) & 0xffull;  // only the disagreeing bits of this byte vary
        if (t < 256) hist[t] = 0;
        __syncthreads();
        // (one shared-memory atomic per key.  Aggregating equal digits within a warp first — __match_any_sync, one atomic per
        // distinct digit — was measured and is SLOWER: the match costs a round per distinct value, and in the low bytes all 32
        // lanes differ: trec shape 2.18 -> 3.07 ms, C5 shard 27.1 -> 29.6 ms, S0 nq = 16 1.04 -> 1.16 ms.)
        for (int i = t; i < n; i += kT) {
            const uint64_t k = pool[i];
            if (((k ^ prefix) & mask) == 0) atomicAdd(&hist[(int)((k >> shift) & dmask)], 1);
        }
        __syncthreads();
        if (warp == 0) {  // largest digit d with count(digit >= d) >= want
            int loc = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) loc += hist[lane * 8 + b];
            int suf = loc;  // suffix sum over lanes (lane 31 = highest digits)
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int v = __shfl_down_sync(0xffffffffu, suf, s);
                if (lane + s < 32) suf += v;
            }
            const unsigned ok = __ballot_sync(0xffffffffu, suf >= want);
            const int L = 31 - __clz((int)ok);
            if (lane == L) {
                int above = suf - loc;
                int d = lane * 8 + 7;
                for (; d > lane * 8; --d) {
                    if (above + hist[d] >= want) break;
                    above += hist[d];
                }
                s_int[0] = d;
                s_int[1] = want - above;
            }
        }
        __syncthreads();
        prefix |= (uint64_t)s_int[0] << shift;
        mask |= dmask << shift;
        want = s_int[1];
        __syncthreads();
    }
    return prefix;  // keys are unique: exactly `want_initial` keys are >= prefix
}

__global__ void __launch_bounds__(256) pq_epoch_select_kernel(const EpochSelParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* pool = reinterpret_cast<uint64_t*>(smem_raw);                      // lmax keys: old carry + gathered candidates
    uint64_t* out = pool + p.lmax;                                               // kp keys: survivors
    int* s_cnt = reinterpret_cast<int*>(out + p.kp);                             // n_sub
    int* s_off = s_cnt + p.n_sub;                                                // n_sub
    __shared__ int hist[256];
    __shared__ uint64_t s_u64[16];
    __shared__ int s_int[2];
    __shared__ int s_total, s_ovf, s_nc, s_slot;
    __shared__ int s_wbase[8];
    __shared__ unsigned long long s_dropkey;
    const int q = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t redo_mask = p.st.redo[q];
    if (p.is_redo && (redo_mask & p.epoch_bit) == 0u) return;  // block-uniform: this query's first attempt was fine
    const int n_sub = sel_slabs_of_query(p, q);
    uint64_t* carry = p.st.carry + (size_t)q * p.kp;
    const uint64_t* keys = p.cand_keys + (size_t)q * p.n_sub * p.cap;
    const uint32_t* cnts = p.cand_cnt + (size_t)q * p.n_sub;
    if (t == 0) {
        s_ovf = 0;
        s_nc = 0;
        s_slot = 0;
        s_dropkey = 0ull;
    }
    __syncthreads();
    for (int s = t; s < n_sub; s += 256) {
        const uint32_t c = cnts[s];
        if (c > (uint32_t)p.cap) s_ovf = 1;
        s_cnt[s] = (int)min(c, (uint32_t)p.cap);
    }
    if (!p.is_redo) {
        for (int i = t; i < p.kp; i += 256) {  // the carry's non-empty entries form a prefix
            const uint64_t k = carry[i];
            pool[i] = k;
            if (k != 0ull) atomicMax(&s_nc, i + 1);
        }
    } else {
        // repair: the entries the first attempt took from this epoch's rows are regenerated below — all of them, against the
        // final threshold; ordered compaction of the others, 256 entries at a time
        for (int i0 = 0; i0 < p.kp; i0 += 256) {
            const int i = i0 + t;
            const uint64_t key = i < p.kp ? carry[i] : 0ull;
            const long long row = (long long)key_row(key);
            const bool keep = key != 0ull && !(row >= p.row_begin && row < p.row_end);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_wbase[warp] = __popc(m);
            __syncthreads();
            int before = s_nc;
            for (int w = 0; w < warp; ++w) before += s_wbase[w];
            if (keep) pool[before + __popc(m & ((1u << lane) - 1u))] = key;
            __syncthreads();
            if (t == 0) {
                int tot = 0;
                for (int w = 0; w < 8; ++w) tot += s_wbase[w];
                s_nc += tot;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (warp == 0) {  // exclusive prefix sum of the slab counts: lane owns a contiguous run, warp scan across lanes
        const int per = (n_sub + 31) / 32;
        const int a = lane * per, b = min(n_sub, a + per);
        int run = 0;
        for (int s = a; s < b; ++s) run += s_cnt[s];
        int incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int base = incl - run;
        for (int s = a; s < b; ++s) {
            s_off[s] = base;
            base += s_cnt[s];
        }
        if (lane == 31) s_total = incl;
    }
    __syncthreads();
    const int total = s_total;
    int nc = s_nc;
    for (int c0 = 0; c0 < total;) {  // almost always a single chunk
        const int chunk = min(total - c0, p.lmax - nc);
        // Gather by flat candidate index: thread -> (slab, position) through a binary search over the offsets in shared memory, so
        // every global load is independent of the others and four are in flight per thread.  (A warp per slab walked its 74 slabs
        // of ~13 keys one L2 round trip after the other: 80-95 us per select at nq = 16, a fifth of the whole search.)
        for (int f0 = t; f0 < chunk; f0 += 4 * 256) {
            uint64_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int f = f0 + u * 256;
                v[u] = 0ull;
                if (f < chunk) {
                    const int g = c0 + f;
                    int lo = 0, hi = n_sub - 1;  // last slab whose offset is <= g: it holds candidate g (empty slabs share an offset
                    while (lo < hi) {            // with their successor and are never the last)
                        const int mid = (lo + hi + 1) >> 1;
                        if (s_off[mid] <= g) lo = mid;
                        else hi = mid - 1;
                    }
                    v[u] = keys[(size_t)lo * p.cap + (size_t)(g - s_off[lo])];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int f = f0 + u * 256;
                if (f < chunk) pool[nc + f] = v[u];
            }
        }
        __syncthreads();
        const int n = nc + chunk;
        if (n > p.kp) {
            const uint64_t pivot = block_radix_select(pool, n, p.kp, hist, s_u64, s_int);
            if (t == 0) s_slot = 0;
            __syncthreads();
            // truncation bookkeeping: when this epoch is going to be repaired, its own rows are regenerated, not lost
            const bool redo_coming = s_ovf && p.allow_redo;
            uint64_t dropped = 0ull;
            for (int i = t; i < n; i += 256) {
                const uint64_t k = pool[i];
                const long long row = (long long)key_row(k);
                if (k >= pivot) out[atomicAdd(&s_slot, 1)] = k;
                else if (!redo_coming || !(row >= p.row_begin && row < p.row_end)) dropped = max(dropped, k);
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) dropped = max(dropped, __shfl_xor_sync(0xffffffffu, dropped, s));
            if (lane == 0 && dropped) atomicMax(&s_dropkey, (unsigned long long)dropped);
            __syncthreads();
            for (int i = t; i < p.kp; i += 256) pool[i] = out[i];
            nc = p.kp;
            __syncthreads();
        } else {
            nc = n;
        }
        c0 += chunk;
    }
    // sort the (at most K') survivors, best first
    for (int i = nc + t; i < p.kp; i += 256) pool[i] = 0ull;
    __syncthreads();
    block_sort_desc<256>(pool, p.kp);
    if (t == 0) {
        // the k-th best of what fitted is the score of a real row: a valid (and usually much tighter) threshold either way
        const uint64_t kth = pool[p.k - 1];
        if (kth != 0ull) p.st.thr[q] = fmaxf(p.st.thr[q], key_score(kth) - p.st.two_e[q]);
        if (s_dropkey != 0ull) p.st.dropmax[q] = fmaxf(p.st.dropmax[q], key_score((uint64_t)s_dropkey));
        if (s_ovf && p.allow_redo) {
            p.st.redo[q] = redo_mask | p.epoch_bit;
            atomicOr(p.st.counters, p.epoch_bit);
            atomicAdd(p.st.counters + 1, 1u);  // running total of repairs, reported in the search statistics
        } else if (s_ovf) {
            p.st.overflow[q] = 1u;
        }
    }
    for (int i = t; i < p.kp; i += 256) carry[i] = pool[i];
    if (p.share.n > 1 && !p.is_redo && warp == 0) {
        const uint64_t ka = pool[p.k - 1], kb = pool[p.share.kr - 1];
        share_publish(p.share, q, ka != 0ull ? key_score(ka) : -INFINITY, kb != 0ull ? key_score(kb) : -INFINITY, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// K' <= 256 (k <= 102: every BASELINE search config): one WARP per query, eight queries per CTA, no block barrier.
// The warp keeps a pool of up to kSelWarpPool keys in its own shared memory: the carry, then the query's slabs, 32 slabs
// at a time (lane = slab: counts prefix-summed with shuffles, entries copied lane by lane).  Reducing the pool means
//   A_k = k-th largest key (warp radix select, 8-bit digits, histogram in the warp's shared memory, bytes all keys share
//         skipped, finished early when the digit bucket is taken whole)            -> threshold = A_k - 2E
//   keep the keys whose score reaches the new threshold (nothing below it can be in the final top-k): usually ~1.5 k keys;
//   only if more than K' reach it is the exact top-K' taken and the best dropped score remembered for the certificate.
// The carry is left unsorted (a compacted prefix): nothing downstream needs its order.
// ------------------------------------------------------------------------------------------------
constexpr int kSelWarpPool = 704;   // keys per warp: 8 x (704 x 8 + 1024) B = 53 KB per CTA, four CTAs = 32 query warps per SM
constexpr int kSelWarps = 8;

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// want-th largest of the n keys in pool (1 <= want <= n); hist: 256 ints private to the warp.  All lanes return it.
__device__ __forceinline__ uint64_t warp_radix_select(const uint64_t* pool, int n, int want, int* hist, int lane) {
    uint64_t a = ~0ull, o = 0ull;
    for (int i = lane; i < n; i += 32) {
        const uint64_t k = pool[i];
        a &= k;
        o |= k;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        a &= __shfl_xor_sync(0xffffffffu, a, s);
        o |= __shfl_xor_sync(0xffffffffu, o, s);
    }
    const uint64_t differ = a ^ o;
    uint64_t prefix = a & ~differ, mask = ~differ;
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (((differ >> shift) & 0xffull) == 0ull) continue;  // warp-uniform
#pragma unroll
        for (int b = 0; b < 8; ++b) hist[lane * 8 + b] = 0;
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const uint64_t k = pool[i];
            if (((k ^ prefix) & mask) == 0ull) atomicAdd(&hist[(int)((k >> shift) & 0xffull)], 1);
        }
        __syncwarp();
        int loc = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) loc += hist[lane * 8 + b];
        int suf = loc;  // suffix sum over lanes (lane 31 owns the highest digits)
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int v = __shfl_down_sync(0xffffffffu, suf, s);
            if (lane + s < 32) suf += v;
        }
        const unsigned ok = __ballot_sync(0xffffffffu, suf >= want);
        const int L = 31 - __clz((int)ok);
        int d = 0, rest = 0, bucket = 0;
        if (lane == L) {
            int above = suf - loc;
            d = lane * 8 + 7;
            for (; d > lane * 8; --d) {
                if (above + hist[d] >= want) break;
                above += hist[d];
            }
            rest = want - above;
            bucket = hist[d];
        }
        d = __shfl_sync(0xffffffffu, d, L);
        rest = __shfl_sync(0xffffffffu, rest, L);
        bucket = __shfl_sync(0xffffffffu, bucket, L);
        prefix = (prefix & ~(0xffull << shift)) | ((uint64_t)d << shift);
        mask |= 0xffull << shift;
        want = rest;
        __syncwarp();
        if (bucket == want) {  // the whole bucket is taken: the answer is its smallest key
            uint64_t m = ~0ull;
            for (int i = lane; i < n; i += 32) {
                const uint64_t k = pool[i];
                if (((k ^ prefix) & mask) == 0ull) m = min(m, k);
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, s));
            return m;
        }
    }
    return prefix;  // every byte decided: the key itself
}

struct WarpSelState {
    float thr, two_e, a_score;   // a_score: a score at least k keys seen by this warp reach (or -inf)
    float b_score;               // the same for ceil(k / n_shards) keys (threshold exchange)
    uint64_t dropkey;
    bool redo_coming;
};

// Largest digit d (0..255) with  count(digit >= d) >= want  in a 256-bin histogram owned by the warp (lane l holds bins 8l..8l+7);
// -1 when fewer than `want` entries were counted.  All lanes return it.
__device__ __forceinline__ int warp_hist_rank(const int* hist, int want, int lane) {
    int loc = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) loc += hist[lane * 8 + b];
    int suf = loc;  // suffix sum over lanes (lane 31 owns the highest bins)
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int v = __shfl_down_sync(0xffffffffu, suf, s);
        if (lane + s < 32) suf += v;
    }
    const unsigned ok = __ballot_sync(0xffffffffu, suf >= want);
    if (ok == 0u) return -1;
    const int L = 31 - __clz((int)ok);
    int d = 0;
    if (lane == L) {
        int above = suf - loc;
        d = lane * 8 + 7;
        for (; d > lane * 8; --d) {
            if (above + hist[d] >= want) break;
            above += hist[d];
        }
    }
    return __shfl_sync(0xffffffffu, d, L);
}

// A lower bound on the k-th best score of pool[0, n) from ONE pass: scores are counted in 256 bins of width 2E/4 above the current
// threshold; the lower edge of the bin in which the count from the top reaches k is reached by at least k keys.  It is at most
// 2E/4 below the true k-th best — the filter then admits as if the margin were 2.25 E instead of 2 E, a few per cent more
// survivors — and it costs a third of the exact radix select's passes, which is what this kernel's time is (ncu: 10k instructions
// per query, issue-bound).  Returns false (exact select needed) when the range does not reach: a threshold still at the floor
// (bootstrap epoch) or the k-th best more than 64 x 2E above it.
__device__ __forceinline__ bool warp_binned_bounds(const EpochSelParams& p, const uint64_t* pool, int n, WarpSelState& w, int* hist, int lane) {
    const float base = w.thr, width = 0.25f * w.two_e;
    if (!(width > 0.f) || !(base > -1e30f)) return false;
    const float inv = 1.f / width;
#pragma unroll
    for (int b = 0; b < 8; ++b) hist[lane * 8 + b] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const float sc = key_score(pool[i]);
        if (sc >= base) atomicAdd(&hist[(int)fminf((sc - base) * inv, 255.f)], 1);
    }
    __syncwarp();
    const int dk = warp_hist_rank(hist, p.k, lane);
    if (dk < 0 || dk >= 255) {  // fewer than k keys reach the threshold (nothing to raise), or the k-th best is beyond the bins
        __syncwarp();
        return dk < 0;
    }
    const float slack = 1e-4f * width + 1e-6f * fabsf(base);  // (the bin coordinate and the edge are rounded: stay below the true edge)
    const float edge = fmaf((float)dk, width, base) - slack;
    w.a_score = fmaxf(w.a_score, edge);
    w.thr = fmaxf(w.thr, edge - w.two_e);
    if (p.share.n > 1) {
        const int db = warp_hist_rank(hist, p.share.kr, lane);
        if (db >= 0 && db < 255) w.b_score = fmaxf(w.b_score, fmaf((float)db, width, base) - slack);
    }
    __syncwarp();
    return true;
}

// pool[0, n) -> pool[0, m): what can still matter (m <= kp); returns m.
__device__ __forceinline__ int warp_sel_reduce(const EpochSelParams& p, uint64_t* pool, int n, WarpSelState& w, int* hist, int lane) {
    if (n >= p.k && !warp_binned_bounds(p, pool, n, w, hist, lane)) {
        const float ak = key_score(warp_radix_select(pool, n, p.k, hist, lane));
        w.a_score = fmaxf(w.a_score, ak);
        w.thr = fmaxf(w.thr, ak - w.two_e);
    }
    int reach = 0;
    for (int i = lane; i < n; i += 32) reach += key_score(pool[i]) >= w.thr ? 1 : 0;
    reach = warp_sum(reach);
    uint64_t pivot = 0ull;
    if (reach > p.kp) pivot = warp_radix_select(pool, n, p.kp, hist, lane);  // the kp best keys all reach the threshold
    int m = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {  // ordered compaction: a chunk is read by the whole warp before it writes at or below it
        const int i = i0 + lane;
        const uint64_t key = i < n ? pool[i] : 0ull;
        const bool reaches = i < n && key_score(key) >= w.thr;
        const bool keep = reaches && key >= pivot;
        if (reaches && !keep) {
            const long long row = (long long)key_row(key);
            if (!w.redo_coming || !(row >= p.row_begin && row < p.row_end)) w.dropkey = max(w.dropkey, key);
        }
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) pool[m + __popc(km & ((1u << lane) - 1u))] = key;
        m += __popc(km);
        __syncwarp();
    }
    return m;
}

__global__ void __launch_bounds__(kSelWarps * 32) pq_epoch_select_warp_kernel(const EpochSelParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x + warp * gridDim.x;  // queries dealt round-robin to the CTAs (warp_query_grid): every SM gets the same number
    if (q >= p.nq) return;  // (whole warp; this kernel has no block-wide barrier)
    uint64_t* pool = reinterpret_cast<uint64_t*>(smem_raw) + (size_t)warp * kSelWarpPool;
    int* hist = reinterpret_cast<int*>(reinterpret_cast<uint64_t*>(smem_raw) + (size_t)kSelWarps * kSelWarpPool) + warp * 256;
    const uint32_t redo_mask = p.st.redo[q];
    if (p.is_redo && (redo_mask & p.epoch_bit) == 0u) return;
    const int n_sub = sel_slabs_of_query(p, q);
    uint64_t* carry = p.st.carry + (size_t)q * p.kp;
    const uint64_t* keys = p.cand_keys + (size_t)q * p.n_sub * p.cap;
    const uint32_t* cnts = p.cand_cnt + (size_t)q * p.n_sub;

    // Everything this warp reads from global memory is read with many loads in flight: a lone warp that waits out one L2 round
    // trip per load (counts, then carry, then slab after slab) spent 70 us on ~300 keys.
    bool ovf = false;
    for (int s = lane; s < n_sub; s += 128) {
        uint32_t c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = s + 32 * u < n_sub ? cnts[s + 32 * u] : 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) ovf |= c[u] > (uint32_t)p.cap;
    }
    ovf = __any_sync(0xffffffffu, ovf);

    WarpSelState w;
    w.thr = p.st.thr[q];
    w.two_e = p.st.two_e[q];
    w.a_score = -INFINITY;
    w.b_score = -INFINITY;
    w.dropkey = 0ull;
    w.redo_coming = ovf && p.allow_redo;

    // carry -> pool (a repair drops what the first attempt took from this epoch's rows: regenerated below)
    int fill = 0;
    {
        uint64_t ck[8];   // kp <= 256: at most eight keys per lane, all loads issued before the first is used
#pragma unroll
        for (int u = 0; u < 8; ++u) ck[u] = 32 * u < p.kp ? carry[32 * u + lane] : 0ull;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint64_t key = ck[u];
            const long long row = (long long)key_row(key);
            const bool keep = key != 0ull && !(p.is_redo && row >= p.row_begin && row < p.row_end);
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            if (keep) pool[fill + __popc(km & ((1u << lane) - 1u))] = key;
            fill += __popc(km);
        }
    }
    __syncwarp();

    for (int s0 = 0; s0 < n_sub; s0 += 32) {
        const int s = s0 + lane;
        const int c = s < n_sub ? (int)min(cnts[s], (uint32_t)p.cap) : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        if (fill + total > kSelWarpPool && fill > p.kp) {
            cp_async_wait_all();
            __syncwarp();
            fill = warp_sel_reduce(p, pool, fill, w, hist, lane);
        }
        if (fill + total <= kSelWarpPool) {
            // slab by slab, warp-wide, as asynchronous 8-byte copies straight into the pool: nothing waits until the pool is reduced
            const int off = fill + incl - c;
            unsigned nonempty = __ballot_sync(0xffffffffu, c > 0);
            while (nonempty) {  // (warp-uniform)
                const int j = __ffs(nonempty) - 1;
                nonempty &= nonempty - 1;
                const int cj = __shfl_sync(0xffffffffu, c, j);
                const int oj = __shfl_sync(0xffffffffu, off, j);
                const uint64_t* src = keys + (size_t)(s0 + j) * p.cap;
                for (int i = lane; i < cj; i += 32) cp_async_8(pool + oj + i, src + i);
            }
            fill += total;
        } else {  // 32 slabs hold more than the pool has room for (rows in document order): slab by slab, warp-wide copies
            cp_async_wait_all();
            __syncwarp();
            for (int j = 0; j < 32; ++j) {
                const int cj = __shfl_sync(0xffffffffu, c, j);
                const uint64_t* src = keys + (size_t)(s0 + j) * p.cap;
                for (int done = 0; done < cj;) {
                    if (fill == kSelWarpPool) fill = warp_sel_reduce(p, pool, fill, w, hist, lane);
                    const int take = min(cj - done, kSelWarpPool - fill);
                    for (int i = lane; i < take; i += 32) pool[fill + i] = src[done + i];
                    fill += take;
                    done += take;
                    __syncwarp();
                }
            }
        }
    }
    cp_async_wait_all();
    __syncwarp();
    const int m = warp_sel_reduce(p, pool, fill, w, hist, lane);
    for (int i = lane; i < p.kp; i += 32) carry[i] = i < m ? pool[i] : 0ull;
    if (lane == 0) {
        p.st.thr[q] = w.thr;
        if (w.dropkey != 0ull) p.st.dropmax[q] = fmaxf(p.st.dropmax[q], key_score(w.dropkey));
        if (ovf && p.allow_redo) {
            p.st.redo[q] = redo_mask | p.epoch_bit;
            atomicOr(p.st.counters, p.epoch_bit);
            atomicAdd(p.st.counters + 1, 1u);
        } else if (ovf) {
            p.st.overflow[q] = 1u;
        }
    }
    if (p.share.n > 1 && !p.is_redo) {
        float b = w.b_score;   // from the binned pass when it applied; otherwise (first epochs) the exact one
        if (!(b > -INFINITY) && m >= p.share.kr) b = key_score(warp_radix_select(pool, m, p.share.kr, hist, lane));
        share_publish(p.share, q, w.a_score, b, lane);
    }
}

struct RescoreParams {
    QState st;
    const float* queries;    // [nq][128] fp32
    const float* rows;       // [ntotal][128] fp32
    const float* row_norms;
    const float* q_norms;
    const uint8_t* q_bad;
    int kp, k, metric, work, nq;
    long long id_base;
    float* D;
    long long* I;
    uint8_t* fail;
    uint32_t* fail_count;
};

// Exactness certificate: the carry was only ever truncated below the final admission threshold.
__device__ __forceinline__ bool rescore_certificate_fails(const RescoreParams& p, int q) {
    bool fail = p.q_bad[q] != 0 || p.st.overflow[q] != 0;
    const float two_e = p.st.two_e[q];
    const float dropmax = p.st.dropmax[q];
    if (dropmax > -INFINITY && two_e > 0.f && dropmax >= p.st.thr[q]) fail = true;
    return fail;
}
__device__ __forceinline__ void rescore_emit(const RescoreParams& p, int q, int i, uint64_t key) {
    float d;
    long long id;
    if (key == 0ull) {
        id = -1;
        d = (p.metric == kMetricL2) ? FLT_MAX : -FLT_MAX;
    } else {
        id = (long long)key_row(key) + p.id_base;
        const float s = key_score(key);
        d = (p.metric == kMetricL2) ? fmaxf(0.f, p.q_norms[q] - s) : s;
    }
    p.D[(size_t)q * p.k + i] = d;
    p.I[(size_t)q * p.k + i] = id;
}

// One CTA per query (any K'): exact fp32 scores for the carry list, final order, certificate.
__global__ void __launch_bounds__(256) pq_rescore_kernel(const RescoreParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* work = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ float s_q[kDim];
    const int q = blockIdx.x;
    const int t = threadIdx.x;
    const uint64_t* carry = p.st.carry + (size_t)q * p.kp;
    if (t < kDim) s_q[t] = p.queries[(size_t)q * kDim + t];
    __syncthreads();
    // a carry entry whose bf16 score is below the final admission threshold (the best known k-th best bf16 score - 2E)
    // cannot be in the exact top-k: skip it
    const float bar = p.st.thr[q];
    for (int i = t; i < p.work; i += 256) {
        uint64_t out = 0ull;
        if (i < p.kp) {
            const uint64_t key = carry[i];
            if (key != 0ull && key_score(key) >= bar) {
                const uint32_t row = key_row(key);
                // the engine's defined score (pq_common.cuh: engine_dot): 8 chains of 16 dims, tree-combined
                const float4* r4 = reinterpret_cast<const float4*>(p.rows + (size_t)row * kDim);
                float pj[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float a = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 v = __ldg(r4 + 4 * j + i);
                        a = fmaf(v.x, s_q[16 * j + 4 * i + 0], a);
                        a = fmaf(v.y, s_q[16 * j + 4 * i + 1], a);
                        a = fmaf(v.z, s_q[16 * j + 4 * i + 2], a);
                        a = fmaf(v.w, s_q[16 * j + 4 * i + 3], a);
                    }
                    pj[j] = a;
                }
                float acc = ((pj[0] + pj[1]) + (pj[2] + pj[3])) + ((pj[4] + pj[5]) + (pj[6] + pj[7]));
                if (p.metric == kMetricL2) acc = fmaf(2.f, acc, -__ldg(p.row_norms + row));
                if (acc >= PQ_THR_FLOOR) out = make_key(acc, row);
            }
        }
        work[i] = out;
    }
    __syncthreads();
    block_sort_desc<256>(work, p.work);
    for (int i = t; i < p.k; i += 256) rescore_emit(p, q, i, work[i]);
    if (t == 0) {
        const bool fail = rescore_certificate_fails(p, q);
        p.fail[q] = fail ? 1 : 0;
        if (fail) atomicAdd(p.fail_count, 1u);
    }
}

// K' <= 256: one warp per query.  Rows are read by the whole warp (one coalesced 512-B read per row, four rows in flight),
// the score is the engine's defined one (warp_engine_dot), the K' exact keys are sorted by a warp-level bitonic network.
__global__ void __launch_bounds__(kSelWarps * 32) pq_rescore_warp_kernel(const RescoreParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x + warp * gridDim.x;
    if (q >= p.nq) return;
    uint64_t* work = reinterpret_cast<uint64_t*>(smem_raw) + (size_t)warp * p.kp;
    const uint64_t* carry = p.st.carry + (size_t)q * p.kp;
    const float bar = p.st.thr[q];
    __align__(16) __shared__ float s_q[kSelWarps][kDim];
    reinterpret_cast<float4*>(s_q[warp])[lane] = __ldg(reinterpret_cast<const float4*>(p.queries + (size_t)q * kDim) + lane);
    uint64_t ck[8];   // kp <= 256: the carry in eight loads, all in flight together
#pragma unroll
    for (int u = 0; u < 8; ++u) ck[u] = 32 * u < p.kp ? carry[32 * u + lane] : 0ull;
    __syncwarp();
    const int g = lane >> 3;  // lane group: scores the g-th of the (up to) four rows of a pass (quad_engine_dot)
#pragma unroll
    for (int u8 = 0; u8 < 8; ++u8) {
        const int i0 = 32 * u8;
        if (i0 >= p.kp) break;
        const uint64_t key = ck[u8];
        unsigned live = __ballot_sync(0xffffffffu, key != 0ull && key_score(key) >= bar);
        uint64_t mine = 0ull;
        while (live) {  // warp-uniform: four candidates per pass
            int src[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                src[u] = -1;
                if (live) {
                    src[u] = __ffs(live) - 1;
                    live &= live - 1;
                }
            }
            const int my_src = g == 0 ? src[0] : (g == 1 ? src[1] : (g == 2 ? src[2] : src[3]));
            const uint32_t row = key_row(__shfl_sync(0xffffffffu, key, my_src < 0 ? 0 : my_src));
            float sc = quad_engine_dot(my_src >= 0 ? p.rows + (size_t)row * kDim : nullptr, s_q[warp], lane);
            if (my_src >= 0 && p.metric == kMetricL2) sc = fmaf(2.f, sc, -__ldg(p.row_norms + row));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float su = __shfl_sync(0xffffffffu, sc, 8 * u);
                const uint32_t ru = __shfl_sync(0xffffffffu, row, 8 * u);
                if (lane == src[u] && su >= PQ_THR_FLOOR) mine = make_key(su, ru);
            }
        }
        work[i0 + lane] = mine;
    }
    __syncwarp();
    for (int size = 2; size <= p.kp; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < (p.kp >> 1); i += 32) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t x = work[lo], y = work[hi];
                if ((x < y) == desc) {
                    work[lo] = y;
                    work[hi] = x;
                }
            }
            __syncwarp();
        }
    }
    for (int i = lane; i < p.k; i += 32) rescore_emit(p, q, i, work[i]);
    if (lane == 0) {
        const bool fail = rescore_certificate_fails(p, q);
        p.fail[q] = fail ? 1 : 0;
        if (fail) atomicAdd(p.fail_count, 1u);
    }
}


// ------------------------------------------------------------------------------------------------
// k = 1 (k-means assignment, group_paras.py:45,51): one warp per query folds its slabs directly — no carry list,
// no sort.  The slabs hold one record per group of four rows whose best bf16 score was within 2E of the best score its
// thread had seen, so the exact best row sits in a group within 2E of the overall best bf16 score; the rows of those
// (typically 1-3) groups are rescored with the engine's defined fp32 score by the whole warp (coalesced 512-B reads).
// ------------------------------------------------------------------------------------------------
struct K1Params {
    const uint64_t* cand_keys;
    const uint32_t* cand_cnt;
    const float* two_e;
    const float* queries;    // [nq][128] fp32
    const float* rows;       // [ntotal][128] fp32
    const float* row_norms;
    const float* q_norms;
    const uint8_t* q_bad;
    int nq, n_sub, cap, metric;
    long long n_rows;
    long long id_base;
    float* D;
    long long* I;
    uint8_t* fail;
    uint32_t* fail_count;  // number of queries flagged in `fail` (the host fetches the flags only when it is not zero)
};

__global__ void __launch_bounds__(256) pq_k1_finalize_kernel(const K1Params p) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= p.nq) return;
    const uint64_t* keys = p.cand_keys + (size_t)q * p.n_sub * p.cap;
    const uint32_t* cnts = p.cand_cnt + (size_t)q * p.n_sub;
    // A lone warp per query: every global read below is issued together with its neighbours (query, bound and counts up front;
    // slab entries four slabs per round; the four rows of a group together) — the dependent chain is counts -> entries -> rows.
    const float4 qv = __ldg(reinterpret_cast<const float4*>(p.queries + (size_t)q * kDim) + lane);
    const float two_e = p.two_e[q];
    bool fail = p.q_bad[q] != 0;
    // pass 1: best bf16 score over all candidates
    uint32_t best_hi = 0;
    for (int s0 = 0; s0 < p.n_sub; s0 += 32) {
        const int s = s0 + lane;
        const uint32_t c = s < p.n_sub ? cnts[s] : 0u;
        fail |= c > (uint32_t)p.cap;
        const int n = (int)min(c, (uint32_t)p.cap);
        unsigned nonempty = __ballot_sync(0xffffffffu, n > 0);
        while (nonempty) {  // (warp-uniform) four slabs per round
            int sj[4], nj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                sj[u] = -1;
                nj[u] = 0;
                if (nonempty) {
                    sj[u] = __ffs(nonempty) - 1;
                    nonempty &= nonempty - 1;
                    nj[u] = __shfl_sync(0xffffffffu, n, sj[u]);
                }
            }
            const int longest = max(max(nj[0], nj[1]), max(nj[2], nj[3]));
            for (int i = lane; i < longest; i += 32) {
                uint64_t kv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) kv[u] = i < nj[u] ? keys[(size_t)(s0 + sj[u]) * p.cap + i] : 0ull;
#pragma unroll
                for (int u = 0; u < 4; ++u) best_hi = max(best_hi, (uint32_t)(kv[u] >> 32));
            }
        }
    }
    fail = __any_sync(0xffffffffu, fail);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best_hi = max(best_hi, __shfl_xor_sync(0xffffffffu, best_hi, o));
    uint64_t best = 0ull;
    __align__(16) __shared__ float s_q[8][kDim];
    if (best_hi != 0 && !fail) {
        const float bar = ordered_to_f32(best_hi) - two_e;
        const int w = threadIdx.x >> 5;
        reinterpret_cast<float4*>(s_q[w])[lane] = qv;
        __syncwarp();
        // pass 2: exact score of every candidate within 2E of the best (the slab entries come from L1 now).  A record is a group of
        // four consecutive rows: one quad_engine_dot scores all four (ncu: this kernel was issue-bound at 870 instructions per
        // query, most of them the row-by-row warp-wide dots).
        for (int s = 0; s < p.n_sub; ++s) {
            const int n = (int)min(cnts[s], (uint32_t)p.cap);
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                const uint64_t key = i < n ? keys[(size_t)s * p.cap + i] : 0ull;
                unsigned live = __ballot_sync(0xffffffffu, key != 0ull && key_score(key) >= bar);
                while (live) {
                    const int src = __ffs(live) - 1;
                    live &= live - 1;
                    const long long row = (long long)key_row(__shfl_sync(0xffffffffu, key, src)) + (lane >> 3);
                    const bool in = row < p.n_rows;
                    float sc = quad_engine_dot(in ? p.rows + (size_t)row * kDim : nullptr, s_q[w], lane);
                    if (in) {
                        if (p.metric == kMetricL2) sc = fmaf(2.f, sc, -__ldg(p.row_norms + row));
                        if (sc >= PQ_THR_FLOOR) best = max(best, make_key(sc, (uint32_t)row));
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    }
    if (lane == 0) {
        float d;
        long long id;
        if (best == 0ull) {
            id = -1;
            d = (p.metric == kMetricL2) ? FLT_MAX : -FLT_MAX;
        } else {
            id = (long long)key_row(best) + p.id_base;
            const float sc = key_score(best);
            d = (p.metric == kMetricL2) ? fmaxf(0.f, p.q_norms[q] - sc) : sc;
        }
        p.D[q] = d;
        p.I[q] = id;
        p.fail[q] = fail ? 1 : 0;
        if (fail) atomicAdd(p.fail_count, 1u);
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
static int next_pow2i(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// cudaFuncSetAttribute costs microseconds of host time per call: a kernel's dynamic shared-memory limit is raised only when
// a launch needs more than the kernel was last given on that device.
static cudaError_t ensure_dyn_smem_impl(const void* fn, size_t smem, int device, cudaError_t (*set)(const void*, int)) {
    struct Slot {
        const void* fn;
        int dev;
        size_t granted;
    };
    static Slot slots[256];
    static int n_slots = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    Slot* sl = nullptr;
    for (int i = 0; i < n_slots; ++i)
        if (slots[i].fn == fn && slots[i].dev == device) sl = &slots[i];
    if (sl && sl->granted >= smem) return cudaSuccess;
    const cudaError_t e = set(fn, (int)smem);
    if (e != cudaSuccess) return e;
    if (!sl && n_slots < 256) {
        sl = &slots[n_slots++];
        sl->fn = fn;
        sl->dev = device;
        sl->granted = 0;
    }
    if (sl) sl->granted = smem;
    return cudaSuccess;
}
template <typename K>
static cudaError_t ensure_dyn_smem(K* kernel, size_t smem, int device) {
    if (smem <= 48 * 1024) return cudaSuccess;
    return ensure_dyn_smem_impl((const void*)kernel, smem, device,
                                [](const void* f, int v) { return cudaFuncSetAttribute((K*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, v); });
}

template <int M, bool L2, bool K1, int SETS>
static cudaError_t launch_filter(const CUtensorMap& tc, const MmaParams& p, int n_ctas, int device, cudaStream_t stream) {
    const size_t smem = (size_t)kStages * kStageBytes + 256 + (size_t)kMaxMTiles * 4 * SETS * 32 * 12;  // ring + control + epilogue state
    cudaError_t e = ensure_dyn_smem(pq_mma_filter_kernel<M, L2, K1, SETS>, smem, device);
    if (e != cudaSuccess) return e;
    pq_mma_filter_kernel<M, L2, K1, SETS><<<n_ctas, (2 + 4 * SETS) * 32, smem, stream>>>(tc, p);
    return cudaGetLastError();
}
template <int M, bool L2, bool K1>
static cudaError_t launch_filter_s(const CUtensorMap& tc, const MmaParams& p, int n_ctas, int device, cudaStream_t stream) {
    if (p.sets == kPlanSetsTight) return launch_filter<M, L2, K1, kPlanSetsTight>(tc, p, n_ctas, device, stream);
    return launch_filter<M, L2, K1, kPlanSetsLoose>(tc, p, n_ctas, device, stream);
}
template <bool L2, bool K1>
static cudaError_t launch_filter_m(int m_max, const CUtensorMap& tc, const MmaParams& p, int n_ctas, int device, cudaStream_t stream) {
    if (m_max == 1) return launch_filter_s<1, L2, K1>(tc, p, n_ctas, device, stream);
    if (m_max == 2) return launch_filter_s<2, L2, K1>(tc, p, n_ctas, device, stream);
    return launch_filter_s<4, L2, K1>(tc, p, n_ctas, device, stream);
}
static cudaError_t launch_filter_any(int m_max, bool l2, bool k1, const CUtensorMap& tc, const MmaParams& p, int n_ctas, int device,
                                     cudaStream_t stream) {
    if (p.sets != kPlanSetsTight && p.sets != kPlanSetsLoose) return cudaErrorInvalidValue;
    if (l2) return k1 ? launch_filter_m<true, true>(m_max, tc, p, n_ctas, device, stream) : launch_filter_m<true, false>(m_max, tc, p, n_ctas, device, stream);
    return k1 ? launch_filter_m<false, true>(m_max, tc, p, n_ctas, device, stream) : launch_filter_m<false, false>(m_max, tc, p, n_ctas, device, stream);
}

// Epoch select and rescoring: one warp per query when the carry list fits a warp's pool (K' <= 256) and there are enough
// queries to fill the GPU with warps (1024; PROQA_B200_SELECT_WARP_MIN overrides — the CPU tests run both kernels on small
// batches through it); one CTA per query otherwise: eight warps on one list finish a lone list sooner.
static int sel_warp_min_queries() {
    const char* s = getenv("PROQA_B200_SELECT_WARP_MIN");
    return (s && *s) ? atoi(s) : 1024;
}
// Grid of the warp-per-query kernels: a whole number of CTAs per SM, the queries dealt to them round-robin (q = block + warp x grid).
// ceil(nq / 8) CTAs put 3.05 CTAs on an SM for C2's 3610 queries — eight SMs with a fourth CTA set the kernel's time (ncu: SMs busy
// 112k of 170k cycles).
static int warp_query_grid(int nq, int n_sms) {
    const int per_wave = std::max(1, n_sms) * kSelWarps;
    return std::max(1, n_sms) * ((nq + per_wave - 1) / per_wave);
}
static cudaError_t launch_epoch_select(const EpochSelParams& sp, int n_sms, int device, cudaStream_t stream) {
    if (sp.kp <= 256 && sp.nq >= sel_warp_min_queries()) {
        const size_t smem = (size_t)kSelWarps * (kSelWarpPool * 8 + 256 * 4);
        cudaError_t e = ensure_dyn_smem(pq_epoch_select_warp_kernel, smem, device);
        if (e != cudaSuccess) return e;
        pq_epoch_select_warp_kernel<<<warp_query_grid(sp.nq, n_sms), kSelWarps * 32, smem, stream>>>(sp);
    } else {
        const size_t smem = ((size_t)sp.lmax + sp.kp) * 8 + (size_t)sp.n_sub * 8;
        cudaError_t e = ensure_dyn_smem(pq_epoch_select_kernel, smem, device);
        if (e != cudaSuccess) return e;
        pq_epoch_select_kernel<<<sp.nq, 256, smem, stream>>>(sp);
    }
    return cudaGetLastError();
}
static cudaError_t launch_rescore(const RescoreParams& rp, int n_sms, int device, cudaStream_t stream) {
    if (rp.kp <= 256 && rp.nq >= sel_warp_min_queries()) {
        const size_t smem = (size_t)kSelWarps * rp.kp * 8;
        pq_rescore_warp_kernel<<<warp_query_grid(rp.nq, n_sms), kSelWarps * 32, smem, stream>>>(rp);
    } else {
        const size_t smem = (size_t)rp.work * 8;
        cudaError_t e = ensure_dyn_smem(pq_rescore_kernel, smem, device);
        if (e != cudaSuccess) return e;
        pq_rescore_kernel<<<rp.nq, 256, smem, stream>>>(rp);
    }
    return cudaGetLastError();
}

// Cross-shard threshold exchange for this batch, or n = 0 when it is off (pq_index_share_connect).
static ShareParams make_share_params(const pq_index* ix, int nq, int k, int batch) {
    ShareParams sh;
    memset(&sh, 0, sizeof(sh));
    const pq_share_state& ss = ix->share;
    // (the tag has four bits for the query batch: a search of more than 16 batches of 2^18 queries exchanges nothing after the 16th)
    if (!ss.connected || ss.n < 2 || ss.n > kShareMaxPeers || nq > ss.cap_q || k == 1 || batch >= 16) return sh;
    sh.n = ss.n;
    sh.rank = ss.rank;
    sh.kr = (k + ss.n - 1) / ss.n;
    sh.tag = (ss.seq << 4) | ((uint32_t)batch & 15u);
    sh.cap_q = ss.cap_q;
    sh.wait_ns = ss.wait_us * 1000;
    for (int i = 0; i < ss.n; ++i) sh.peer[i] = ss.peer[i];
    return sh;
}

// Producer pacing of a filter launch (pq_mma_filter_kernel: pace_leave / pace_wait): on when several CTA groups stream rows
// that do not fit L2 together.  A grid of at most one CTA per SM is resident at once and paces as one cohort.  A larger grid
// (65,536 queries: 1024 CTAs, seven waves) paces wave by wave — cohort c = CTAs [c n_sms, (c+1) n_sms): they start together as
// the previous, paced, wave ends together; a CTA whose cohort is not all there within the bounded wait simply runs unpaced.
// Returns the arrival counters one cohort needs (0: unpaced).  PROQA_B200_PACE = 0 off, 1 single-wave grids only, 2 (default)
// every grid; PROQA_B200_PACE_MIN_TILES / _SHIFT are test hooks.
static int pace_shift() {
    static const int shift = [] { const char* e = getenv("PROQA_B200_PACE_SHIFT"); return e ? atoi(e) : 8; }();
    return shift;
}
static int pace_cohorts(const pq_index* ix, int n_ctas) { return (n_ctas + ix->n_sms - 1) / ix->n_sms; }
static long long pace_blocks_for(const pq_index* ix, const GridShape& gs, int n_ctas, long long row_begin, long long row_end) {
    static const int mode = [] { const char* e = getenv("PROQA_B200_PACE"); return e ? atoi(e) : 2; }();
    static const long long min_tiles = [] { const char* e = getenv("PROQA_B200_PACE_MIN_TILES"); return e ? atoll(e) : 2048LL; }();
    const long long tiles = (row_end - row_begin + kBN - 1) / kBN;
    if (mode <= 0 || gs.n_groups < 2 || tiles < min_tiles) return 0;
    if (n_ctas > ix->n_sms && (mode < 2 || pace_cohorts(ix, n_ctas) > 64)) return 0;
    return ((tiles - 1) >> pace_shift()) + 1;
}
// One zeroed counter area for all paced launches of a search (a single memset node), handed out launch by launch.
struct PaceArea {
    uint32_t* base = nullptr;
    long long used = 0, total = 0;
    int reserve(pq_index* ix, long long counters) {
        used = 0;
        total = counters;
        base = nullptr;
        if (counters == 0) return PQ_OK;
        if (const int rc = ix->ws_mma[9].ensure((size_t)counters * 4)) return rc;
        base = (uint32_t*)ix->ws_mma[9].p;
        PQ_CUDA(cudaMemsetAsync(base, 0, (size_t)counters * 4, ix->stream));
        return PQ_OK;
    }
    void assign(const pq_index* ix, MmaParams& mp, long long blocks, int n_ctas) {
        const long long need = blocks * pace_cohorts(ix, n_ctas);
        mp.pace = nullptr;
        mp.pace_shift = pace_shift();
        mp.pace_blocks = 0;
        mp.pace_cohort = 1;
        if (blocks == 0 || base == nullptr || used + need > total) return;
        mp.pace = base + used;
        mp.pace_blocks = (int)blocks;
        mp.pace_cohort = n_ctas <= ix->n_sms ? n_ctas : ix->n_sms;
        used += need;
    }
};

int search_mma_filter(pq_index* ix, int nq_total, const float* dq_all, int k, float* dD_all, long long* dI_all,
                      std::vector<int>* rerun) {
    const long long N = ix->ntotal;
    const int kp = carry_size_for_k(k);
    const bool k1 = (k == 1);
    const bool l2 = ix->metric == kMetricL2;
    const int kMaxBatch = k1 ? (1 << 20) : (1 << 18);  // queries per pass (bounds the candidate slabs)

    for (int qb = 0, batch = 0; qb < nq_total; qb += kMaxBatch, ++batch) {
        const int nq = std::min(kMaxBatch, nq_total - qb);
        const int nq_pad = (nq + kBM - 1) / kBM * kBM;
        const int n_mtiles = nq_pad / kBM;
        const GridShape gs = make_grid_shape(n_mtiles);
        const float* dq = dq_all + (size_t)qb * kDim;
        const uint16_t* dq_bf16 = (const uint16_t*)ix->ws_qbf16.p + (size_t)qb * kDim;
        const float* dq_norm = (const float*)ix->ws_qnorm.p + qb;
        const uint8_t* dq_bad = (const uint8_t*)ix->ws_qbad.p + qb;
        const ShareParams share = make_share_params(ix, nq, k, batch);

        // ---- epoch plan (pq_plan.h) -----------------------------------------------------------
        const std::vector<EpochPlan> plan = plan_epochs(N, k, nq_pad, gs, ix->n_sms, share.n > 1 ? share.n : 1, l2);
        if (plan.size() > 31) return set_error(PQ_ERR_UNSUPPORTED, "search: %zu epochs", plan.size());
        size_t max_slab = 0, max_cnt = 0;
        for (const EpochPlan& ep : plan) {
            const size_t n_sub = (size_t)plan_n_sub(gs, ep);
            max_slab = std::max(max_slab, (size_t)nq_pad * n_sub * ep.cap * 8);
            max_cnt = std::max(max_cnt, (size_t)nq_pad * n_sub * 4);
        }

        // ---- workspaces -----------------------------------------------------------------------
        DevBuf* w = ix->ws_mma;
        int rc = w[0].ensure((size_t)nq_pad * 4);                 // thr
        if (!rc) rc = w[1].ensure((size_t)nq_pad * 4);            // two_e
        if (!rc) rc = w[2].ensure((size_t)nq_pad * 4);            // dropmax
        if (!rc) rc = w[3].ensure((size_t)nq_pad * 4);            // overflow
        if (!rc) rc = w[4].ensure((size_t)nq_pad * kp * 8);       // carry
        if (!rc) rc = w[5].ensure(max_slab);                      // candidate slabs
        if (!rc) rc = w[6].ensure(max_cnt);                       // slab counts
        if (!rc) rc = w[7].ensure((size_t)nq_pad);                // fail flags
        if (!rc) rc = w[8].ensure((size_t)nq_pad * 4 + 256);      // per-query repair masks, then the counters
        if (rc) return rc;
        QState st;
        st.thr = (float*)w[0].p;
        st.two_e = (float*)w[1].p;
        st.dropmax = (float*)w[2].p;
        st.overflow = (uint32_t*)w[3].p;
        st.carry = (uint64_t*)w[4].p;
        st.redo = (uint32_t*)w[8].p;
        st.counters = st.redo + nq_pad;

        if (!k1) PQ_CUDA(cudaMemsetAsync(st.carry, 0, (size_t)nq_pad * kp * 8, ix->stream));
        PQ_CUDA(cudaMemsetAsync(st.counters, 0, 16, ix->stream));
        PaceArea pace;
        {
            long long counters = 0;
            if (!k1)
                for (const EpochPlan& ep : plan)
                    counters += pace_blocks_for(ix, gs, plan_n_ctas(gs, ep), ep.begin, ep.end) * pace_cohorts(ix, plan_n_ctas(gs, ep));
            if (const int prc = pace.reserve(ix, counters)) return prc;
        }
        pq_mma_init_state_kernel<<<(nq_pad + 255) / 256, 256, 0, ix->stream>>>(st, dq_norm, (const float*)ix->ws_qresid.p + qb, dq_bad, nq, nq_pad, kp,
                                                                             ix->max_norm2, ix->max_resid2, ix->metric);
        PQ_CUDA(cudaGetLastError());
        ix->stats[5] += 1;

        MmaParams mp;
        mp.q_bf16 = dq_bf16;
        mp.cand_keys = (uint64_t*)w[5].p;
        mp.cand_cnt = (uint32_t*)w[6].p;
        mp.row_norms = (const float*)ix->norms.p;
        mp.thr = st.thr;
        mp.two_e = st.two_e;
        mp.n_mtiles = n_mtiles;
        mp.base = gs.base;
        mp.rem = gs.rem;
        mp.k1_adapt = k1 ? 1 : 0;
        EpochSelParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.st = st;
        sp.cand_keys = mp.cand_keys;
        sp.cand_cnt = mp.cand_cnt;
        sp.kp = kp;
        sp.k = k;
        sp.lmax = std::max(2 * kp, 4096);  // pool of the CTA kernel: old carry + one chunk of candidates
        sp.nq = nq;
        sp.base = gs.base;
        sp.rem = gs.rem;
        auto set_epoch = [&](const EpochPlan& ep, int e) {
            mp.row_begin = ep.begin;
            mp.row_end = ep.end;
            mp.s1 = ep.s1;
            mp.s0 = ep.s0;
            mp.cap = ep.cap;
            mp.n_sub = plan_n_sub(gs, ep);
            mp.sets = ep.sets;
            mp.pace = nullptr;   // (set per launch by PaceArea::assign; the repair launches run unpaced)
            mp.pace_shift = 0;
            mp.pace_blocks = 0;
            mp.pace_cohort = 1;
            sp.n_sub = mp.n_sub;
            sp.cap = ep.cap;
            sp.s1 = ep.s1;
            sp.s0 = ep.s0;
            sp.sets = ep.sets;
            sp.epoch_bit = 1u << e;
            sp.row_begin = ep.begin;
            sp.row_end = ep.end;
        };

        // ---- epochs: filter (every slab count a query's group owns is written by it), then select -------------------
        for (int e = 0; e < (int)plan.size(); ++e) {
            const EpochPlan& ep = plan[e];
            set_epoch(ep, e);
            mp.redo = nullptr;
            mp.redo_bit = 0u;
            const int n_ctas = plan_n_ctas(gs, ep);
            if (k1) PQ_CUDA(cudaMemsetAsync(mp.cand_cnt, 0, (size_t)nq_pad * mp.n_sub * 4, ix->stream));  // (its finalize reads every slab)
            pace.assign(ix, mp, pace_blocks_for(ix, gs, n_ctas, ep.begin, ep.end), n_ctas);
            ix->prof_begin();
            const cudaError_t fe = launch_filter_any(gs.m_max, l2, k1, ix->tmap_bf16, mp, n_ctas, ix->device, ix->stream);
            ix->prof_end();
            PQ_CUDA(fe);
            ix->stats[3] += 1;
            ix->stats[5] += 1;
            if (k1) {  // single pass: fold the slabs straight into (D, I)
                K1Params kp1;
                kp1.cand_keys = mp.cand_keys;
                kp1.cand_cnt = mp.cand_cnt;
                kp1.two_e = st.two_e;
                kp1.queries = dq;
                kp1.rows = (const float*)ix->rows_f32.p;
                kp1.row_norms = (const float*)ix->norms.p;
                kp1.q_norms = dq_norm;
                kp1.q_bad = dq_bad;
                kp1.nq = nq;
                kp1.n_sub = mp.n_sub;
                kp1.cap = ep.cap;
                kp1.metric = ix->metric;
                kp1.n_rows = N;
                kp1.id_base = ix->id_base;
                kp1.D = dD_all + (size_t)qb;
                kp1.I = dI_all + (size_t)qb;
                kp1.fail = (uint8_t*)w[7].p;
                kp1.fail_count = st.counters + 2;
                pq_k1_finalize_kernel<<<(nq + 7) / 8, 256, 0, ix->stream>>>(kp1);
                PQ_CUDA(cudaGetLastError());
                ix->stats[4] += 1;
                ix->stats[5] += 1;
                continue;
            }
            sp.is_redo = 0;
            sp.allow_redo = ep.begin > 0 ? 1 : 0;  // the bootstrap epoch's slabs hold every row they can see
            sp.share = share;
            ix->prof_begin(1);
            const cudaError_t se = launch_epoch_select(sp, ix->n_sms, ix->device, ix->stream);
            ix->prof_end();
            PQ_CUDA(se);
            ix->stats[4] += 1;
            ix->stats[5] += 1;
            // what the row shards know together, before the next epoch admits on it — and after the last epoch, before the rescoring:
            // a shard's own k-th best score is far below the global one, so without the last exchange every shard would rescore
            // (and ship) ~k rows per query instead of ~k/n
            if (share.n > 1) {
                ix->prof_begin(2);
                pq_share_fold_kernel<<<(nq + 255) / 256, 256, 0, ix->stream>>>(share, st, nq, e);
                ix->prof_end();
                PQ_CUDA(cudaGetLastError());
                ix->stats[5] += 1;
            }
        }
        if (k1) {
            uint32_t counts[4] = {0, 0, 0, 0};
            PQ_CUDA(cudaMemcpyAsync(counts, st.counters, 16, cudaMemcpyDeviceToHost, ix->stream));
            PQ_CUDA(cudaStreamSynchronize(ix->stream));
            if (counts[2] != 0) {
                std::vector<uint8_t> fail((size_t)nq);
                PQ_CUDA(cudaMemcpy(fail.data(), w[7].p, (size_t)nq, cudaMemcpyDeviceToHost));
                for (int q = 0; q < nq; ++q)
                    if (fail[q]) rerun->push_back(qb + q);
            }
            continue;
        }

        // ---- exact rescoring + certificate ------------------------------------------------------
        RescoreParams rp;
        rp.st = st;
        rp.queries = dq;
        rp.rows = (const float*)ix->rows_f32.p;
        rp.row_norms = (const float*)ix->norms.p;
        rp.q_norms = dq_norm;
        rp.q_bad = dq_bad;
        rp.kp = kp;
        rp.k = k;
        rp.metric = ix->metric;
        rp.work = std::max(kp, next_pow2i(k));
        rp.nq = nq;
        rp.id_base = ix->id_base;
        rp.D = dD_all + (size_t)qb * k;
        rp.I = dI_all + (size_t)qb * k;
        rp.fail = (uint8_t*)w[7].p;
        rp.fail_count = st.counters + 2;
        ix->prof_begin(3);
        const cudaError_t re = launch_rescore(rp, ix->n_sms, ix->device, ix->stream);
        ix->prof_end();
        PQ_CUDA(re);
        ix->stats[4] += 1;
        ix->stats[5] += 1;

        uint32_t counts[4] = {0, 0, 0, 0};  // OR of the repair masks, (query, epoch) overflows, failed queries, full exchanges
        PQ_CUDA(cudaMemcpyAsync(counts, st.counters, 16, cudaMemcpyDeviceToHost, ix->stream));
        PQ_CUDA(cudaStreamSynchronize(ix->stream));
        if (counts[0] != 0u) {
            // Some slabs overflowed (rows in document order: a whole cluster above the threshold at once).  What fitted was
            // used — real rows, valid thresholds — and every other epoch is complete, so only the overflowed epochs are run
            // again, for the queries concerned (the others filter at +inf), against their FINAL thresholds; then the
            // rescoring is redone.  The common case pays nothing for this: no launch, no flag check.
            for (int e = 0; e < (int)plan.size(); ++e) {
                if (!(counts[0] & (1u << e))) continue;
                set_epoch(plan[e], e);
                mp.redo = st.redo;
                mp.redo_bit = 1u << e;
                PQ_CUDA(launch_filter_any(gs.m_max, l2, false, ix->tmap_bf16, mp, plan_n_ctas(gs, plan[e]), ix->device, ix->stream));
                sp.is_redo = 1;
                sp.allow_redo = 0;
                sp.share.n = 0;
                PQ_CUDA(launch_epoch_select(sp, ix->n_sms, ix->device, ix->stream));
                ix->stats[3] += 1;
                ix->stats[4] += 1;
                ix->stats[5] += 2;
            }
            PQ_CUDA(cudaMemsetAsync(st.counters + 2, 0, 4, ix->stream));
            PQ_CUDA(launch_rescore(rp, ix->n_sms, ix->device, ix->stream));
            ix->stats[4] += 1;
            ix->stats[5] += 1;
            uint32_t again[4] = {0, 0, 0, 0};
            PQ_CUDA(cudaMemcpyAsync(again, st.counters, 16, cudaMemcpyDeviceToHost, ix->stream));
            PQ_CUDA(cudaStreamSynchronize(ix->stream));
            counts[2] = again[2];
        }
        ix->stats[8] += counts[1];
        ix->stats[9] += counts[3];
        if (counts[2] != 0) {  // rare: fetch the per-query flags
            std::vector<uint8_t> fail((size_t)nq);
            PQ_CUDA(cudaMemcpy(fail.data(), w[7].p, (size_t)nq, cudaMemcpyDeviceToHost));
            for (int q = 0; q < nq; ++q)
                if (fail[q]) rerun->push_back(qb + q);
        }
    }
    return PQ_OK;
}

#include "pq_mma_largek.inl"

}  // namespace pq

// Introspection for the CPU test-suite (no device needed): the launch plan of one tensor-tier search.
extern "C" int pq_plan_describe(int64_t ntotal, int64_t nq, int64_t k, int n_sms, int64_t* out, int out_len) {
    using namespace pq;
    if (ntotal < 1 || nq < 1 || k < 1 || k > kMmaMaxK || n_sms < 1 || !out || out_len < 8)
        return set_error(PQ_ERR_INVALID, "plan_describe: bad arguments");
    const int nq_pad = (int)((nq + kBM - 1) / kBM * kBM);
    const int n_mtiles = nq_pad / kBM;
    const GridShape gs = make_grid_shape(n_mtiles);
    const std::vector<EpochPlan> plan = plan_epochs(ntotal, (int)k, nq_pad, gs, n_sms);
    if ((int)(8 + 8 * plan.size()) > out_len) return set_error(PQ_ERR_INVALID, "plan_describe: output too small");
    out[0] = (int64_t)plan.size();
    out[1] = gs.n_groups;
    out[2] = gs.base;
    out[3] = gs.rem;
    out[4] = gs.m_max;
    out[5] = carry_size_for_k((int)k);
    out[6] = nq_pad;
    out[7] = kPlanSetsLoose;
    for (size_t e = 0; e < plan.size(); ++e) {
        int64_t* o = out + 8 + 8 * e;
        o[0] = plan[e].begin;
        o[1] = plan[e].end;
        o[2] = plan[e].s1;
        o[3] = plan[e].s0;
        o[4] = plan[e].cap;
        o[5] = plan_n_ctas(gs, plan[e]);
        o[6] = plan_n_sub(gs, plan[e]);
        o[7] = plan[e].sets;
    }
    return PQ_OK;
}


// Same for the large-k path (1024 < k <= PQ_MAX_K): out[0..15] =
//   applies, step, k_sample, sample_rows, pool, sort_n, sample epochs, pass s1, pass s0, pass cap, pass CTAs, pass slabs per
//   query, queries per batch, slab bytes of a full batch, finalize shared-memory bytes, carry length of the sample search
extern "C" int pq_plan_describe_large_k(int64_t ntotal, int64_t nq, int64_t k, int n_sms, int64_t* out, int out_len) {
    using namespace pq;
    if (ntotal < 1 || nq < 1 || k < kPlanMidK || k > PQ_MAX_K || n_sms < 1 || !out || out_len < 16)
        return set_error(PQ_ERR_INVALID, "plan_describe_large_k: bad arguments");
    const LargeKPlan lp = plan_large_k(ntotal, (int)k);
    const int nqb = (int)std::min<int64_t>(nq, kLargeKBatch);
    const int nq_pad = (nqb + kBM - 1) / kBM * kBM;
    const GridShape gs = make_grid_shape(nq_pad / kBM);
    const EpochPlan pass = plan_large_k_pass(ntotal, (int)k, nq_pad, gs, n_sms);
    out[0] = plan_large_k_applies(ntotal, (int)k) ? 1 : 0;
    out[1] = lp.step;
    out[2] = lp.k_sample;
    out[3] = lp.sample_rows;
    out[4] = lp.pool;
    out[5] = lp.sort_n;
    out[6] = lp.sample_rows >= 1 ? (int64_t)plan_epochs(lp.sample_rows, lp.k_sample, nq_pad, gs, n_sms).size() : 0;
    out[7] = pass.s1;
    out[8] = pass.s0;
    out[9] = pass.cap;
    out[10] = plan_n_ctas(gs, pass);
    out[11] = plan_n_sub(gs, pass);
    out[12] = kLargeKBatch;
    out[13] = (int64_t)nq_pad * plan_n_sub(gs, pass) * pass.cap * 8;
    out[14] = (int64_t)std::max(lp.pool, lp.sort_n) * 8 + (int64_t)plan_n_sub(gs, pass) * 4;
    out[15] = carry_size_for_k(lp.k_sample);
    return PQ_OK;
}
