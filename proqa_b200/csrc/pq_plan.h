// proqa_b200 — launch planning of the tensor-core tier (host only, no CUDA): how a query batch and a row shard are cut
// into CTA groups, row slices, epochs and candidate-slab capacities.  Pure functions of (ntotal, nq, k, #SMs) so that the
// CPU test-suite can check their invariants through pq_plan_describe (include/proqa_b200.h).
#pragma once

#include <math.h>

#include <algorithm>
#include <vector>

namespace pq {

constexpr int kPlanTileRows = 128;   // corpus rows per B tile (= kBN in pq_mma.cu)
constexpr int kPlanQueryTile = 128;  // queries per M tile (= kBM)
constexpr int kPlanMaxMTiles = 4;    // query tiles one CTA keeps in tensor memory (= kMaxMTiles)
// Epilogue warp sets of the filter kernel (4 warps each, one per TMEM lane quarter; a set drains 128 / sets columns of every
// accumulator and keeps its own candidate slab per (query, row slice)).  Two variants are compiled and chosen per epoch:
//   4 sets (16 epilogue warps) while the threshold is loose — most 32x32 chunks have a survivor and the append path is
//           issue-bound with two warps per scheduler (ncu: tensor pipe 25 %; 694 -> 556 us for rows 64k..512k of C2);
//   2 sets ( 8 epilogue warps) once it is tight — the kernel is then power-limited, and the four-set kernel executes 40 % more
//           instructions for the same products (ncu: 6.0e9 vs 4.3e9 in the last C2 epoch; sustained SM clock 1612 vs 1676 MHz).
constexpr int kPlanSetsLoose = 4, kPlanSetsTight = 2;
// expected share of 32x32 chunks with a survivor when rows [begin, ..) are filtered at the threshold known after `seen` rows
// (tuning hook, PROQA_B200_LOOSE_ABOVE: the share above which an epoch runs the four-set variant)
inline double& plan_loose_above() {
    static double v = 0.10;
    return v;
}
inline int plan_sets_for(double k, double seen_rows) {
    const double per_pair = 1.5 * k / std::max(1.0, seen_rows);
    return 1024.0 * per_pair > plan_loose_above() ? kPlanSetsLoose : kPlanSetsTight;
}

// (tuning hook, PROQA_B200_K1_SETS: epilogue warp sets of the k = 1 pass; 0 = the default)
inline int& plan_k1_sets() {
    static int v = 0;
    return v;
}

// (tuning hooks, PROQA_B200_GROWTH / PROQA_B200_BOOT_ROWS: epoch growth factor and rows of the bootstrap epoch; 0 = the defaults)
inline long long& plan_growth_override() {
    static long long v = 0;
    return v;
}
inline long long& plan_boot_rows_override() {
    static long long v = 0;
    return v;
}

struct EpochPlan {
    long long begin, end;
    int s1, s0, cap;   // row slices per CTA group (groups owning base+1 / base query tiles), slab capacity
    int sets;          // epilogue warp sets of the kernel variant this epoch runs (= slabs per row slice)
};

struct GridShape {
    int n_groups, base, rem, m_max;
};

inline int plan_next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Carry list length K' >= 2.5 k (power of two, at least 64).
inline int carry_size_for_k(int k) { return std::max(plan_next_pow2((k * 5 + 1) / 2), 64); }

// Query tiles are dealt to ceil(n_mtiles / 4) CTA groups as evenly as possible: `rem` groups own base+1 tiles, the others base.
inline GridShape make_grid_shape(int n_mtiles) {
    GridShape gs;
    gs.n_groups = (n_mtiles + kPlanMaxMTiles - 1) / kPlanMaxMTiles;
    gs.base = n_mtiles / gs.n_groups;
    gs.rem = n_mtiles % gs.n_groups;
    gs.m_max = gs.base + (gs.rem ? 1 : 0);
    return gs;
}

// Row slices per CTA group.
// Base split: with no more groups than SMs, slices in proportion to the query tiles a group owns (equal work per CTA);
// otherwise one slice each.  Then the multiplier c (1..8) that minimises
//     waves(c) x max over group kinds of  tiles_owned x (row tiles per CTA + 6)
// — the 6 stands for the fixed per-CTA cost (TMEM allocation, query staging, pipeline fill and drain).  c > 1 pays when
// the CTA count sits just above a multiple of the SM count or well below it (512 query tiles = 128 groups on 148 SMs:
// one slice each leaves 20 SMs idle, eight slices each fill 7 waves to 98.8 %).
inline void pick_slices(const GridShape& g, int n_mtiles, long long tiles, int n_sms, int* s1, int* s0) {
    long long a0 = 1, b0 = 1;
    if (g.n_groups <= n_sms) {
        a0 = std::max(1LL, (long long)n_sms * (g.base + 1) / n_mtiles);
        b0 = std::max(1LL, (long long)n_sms * g.base / n_mtiles);
        if (g.rem == 0) a0 = b0;
    }
    double best = 1e300;
    long long a = a0, b = b0;
    for (int c = 1; c <= 8; ++c) {
        const long long sa = std::max(1LL, std::min(a0 * c, tiles)), sb = std::max(1LL, std::min(b0 * c, tiles));
        const long long ctas = (long long)g.rem * sa + (long long)(g.n_groups - g.rem) * sb;
        const double waves = (double)((ctas + n_sms - 1) / n_sms);
        const double ta = g.rem ? (g.base + 1) * ((double)((tiles + sa - 1) / sa) + 6.0) : 0.0;
        const double tb = g.base * ((double)((tiles + sb - 1) / sb) + 6.0);
        const double cost = waves * std::max(ta, tb);
        if (cost < best * (1.0 - 1e-3)) {  // prefer the smaller c unless the gain is real
            best = cost;
            a = sa;
            b = sb;
        }
        if (sa >= tiles && sb >= tiles) break;
    }
    *s1 = (int)std::max(1LL, a);
    *s0 = (int)std::max(1LL, b);
}

// Epochs of one search over rows [0, N): contiguous, in order, covering every row once.
// share_n > 1: the corpus is row-sharded over share_n GPUs that exchange thresholds after every epoch (pq_mma.cu:
// ShareParams), so a threshold reflects share_n times the rows this shard has seen: slabs stay small.
inline std::vector<EpochPlan> plan_epochs(long long N, int k, int nq_pad, const GridShape& gs, int n_sms, int share_n = 1, bool l2 = true) {
    std::vector<EpochPlan> plan;
    const int n_mtiles = nq_pad / kPlanQueryTile;
    const int kp = carry_size_for_k(k);
    if (k == 1) {  // running-maximum filter: one pass
        EpochPlan ep;
        ep.begin = 0;
        ep.end = N;
        pick_slices(gs, n_mtiles, (N + kPlanTileRows - 1) / kPlanTileRows, n_sms, &ep.s1, &ep.s0);
        // the running-maximum filter records often.  L2 (bias epilogue: the most instructions per column) gains from four warp sets
        // (filter 61.8 -> 55.2 ms on C4, pass 87.5 -> 83.5 ms); IP does not (46.6 vs 45.8 ms) and pays for the two extra slabs per query
        // in the finalize (pass 71.3 vs 73.2 ms) — measured A/B, tools/gpu_runs/r02_r_k1sets.sh
        ep.sets = plan_k1_sets() ? plan_k1_sets() : (l2 ? kPlanSetsLoose : kPlanSetsTight);
        ep.cap = 128 / ep.sets;        // a thread keeps only rows within 2E of its running maximum: a few dozen at most
        plan.push_back(ep);
        return plan;
    }
    // Epoch growth: every epoch costs a fixed ~0.1-0.2 ms (launches, select) and about 1.5 k (growth - 1) survivors per
    // query; small batches are dominated by the fixed part, large ones by the survivors
    // (measured: nq=16 1.19 ms at 64x; nq=256 1.66 ms at 8x vs 1.81 at 64x).
    // Row shards that exchange thresholds: a threshold reflects share_n times the rows this shard has seen, so slabs can be
    // smaller — but a shard is share_n times shorter than the corpus, so the epochs must NOT grow faster: the loose-threshold
    // epochs would cover a larger part of the shard (8 shards of C3 with growth 32: 37 % of a shard's rows ran with 40 % of the
    // 32x32 chunks taking the append path — 51 ms against 36 ms for the same products with queries split instead).
    const int share_f = share_n >= 4 ? 4 : (share_n >= 2 ? 2 : 1);
    const long long growth = plan_growth_override() > 1 ? plan_growth_override() : (nq_pad <= 128 ? 64LL : (nq_pad <= 512 ? 16LL : 8LL));
    const long long n0 = std::min<long long>(N, std::max<long long>(std::max(1024, plan_next_pow2(2 * kp)), plan_boot_rows_override()));
    long long begin = 0, end = n0;
    while (begin < N) {
        EpochPlan ep;
        ep.begin = begin;
        ep.end = std::min(end, N);
        const long long tiles = (ep.end - ep.begin + kPlanTileRows - 1) / kPlanTileRows;
        if (begin == 0) {  // bootstrap: every score is a candidate, one row tile per slice
            ep.s1 = ep.s0 = (int)tiles;
            ep.sets = kPlanSetsLoose;
            ep.cap = kPlanTileRows / ep.sets;
        } else {
            pick_slices(gs, n_mtiles, tiles, n_sms, &ep.s1, &ep.s0);
            // survivors per query with the threshold frozen at the start of the epoch: about k * (end/begin - 1), times
            // ~1.5 for the 2E margin, on exchangeable rows; three times that is provisioned (rows in document order
            // bring whole clusters above the threshold at once — beyond the provision the epoch is run a second time)
            ep.sets = plan_sets_for((double)k, (double)ep.begin * share_n);
            const double slabs = (double)std::min(ep.s1, ep.s0) * ep.sets;
            const double expect = 1.5 * (double)k * ((double)(ep.end - ep.begin) / (double)ep.begin) / slabs / share_f;
            const int floor_cap = 128 / ep.sets;  // (the same slab memory per slice whatever the number of warp sets)
            ep.cap = std::min(4096, std::max(floor_cap, plan_next_pow2((int)(3.0 * expect) + floor_cap)));
        }
        plan.push_back(ep);
        begin = ep.end;
        end = (ep.end >= N / 2 || ep.end * growth >= N) ? N : ep.end * growth;
    }
    return plan;
}

// ---- large k (1024 < k <= PQ_MAX_K; trec_process.py:76 asks for k = 10000) --------------------------------------------------
// A carry list of 2.5 k keys no longer fits shared memory, so the epochs are replaced by
//   A. thresholds from a sample: the ordinary epoch search, k_sample <= 1024, over a compact copy of every step-th row;
//      its final threshold  A_k_sample(sample) - 2E  sits near full-corpus rank  k_sample * step = 1.35 k
//   B. ONE filter pass over all rows at those thresholds (about 2 k survivors per query, measured in tools/sim_largek.py)
//   C. a finalize kernel per query: k-th best bf16 score of the survivors (radix select over the slabs in global memory),
//      certificate, exact rescoring of the survivors within 2E of it (about 1.5 k rows, kept in shared memory: `pool`),
//      top-k by exact key, sort.
struct LargeKPlan {
    int step, k_sample, pool, sort_n;
    long long sample_rows;
};
constexpr int kLargeKPoolMax = 24576;   // keys (192 KB) the finalize kernel can hold next to its slab counters
constexpr int kLargeKBatch = 8192;      // queries per pass: bounds the candidate slabs (about 0.7 MB per query at k = 10000)

// k > 1024: the sample search runs at k_sample ~ 1024, its threshold aimed at full-corpus rank 1.35 k (rank noise ~ 1/sqrt(1024)
// = 3 %, so rank < k — a failed certificate, the query then costs a whole fp32 scan — is an 8-sigma event).
// kPlanMidK <= k <= 1024 (C5: k = 1000): the same scheme beats the epochs there too — K' = 4096 carry lists make every epoch's
// select expensive and the frozen thresholds let ~7 x 1.5 k rows per query through — but a 1024-row sample search would
// cost as much as it saves: k_sample ~ 160 on every ~10th row, aimed at rank 1.6 k (noise 8 %, 4.7 sigma above k).
constexpr int kPlanMidK = 512;
inline LargeKPlan plan_large_k(long long N, int k) {
    LargeKPlan lp;
    const bool mid = k <= 1024;
    const double rank_target = (mid ? 1.6 : 1.35) * (double)k;
    lp.step = std::max(2, (int)ceil(rank_target / (mid ? 160.0 : 1024.0)));
    lp.k_sample = (int)ceil(rank_target / (double)lp.step);
    lp.sample_rows = N / lp.step;
    lp.sort_n = plan_next_pow2(k);
    lp.pool = std::max(lp.sort_n, std::min(kLargeKPoolMax, std::max(2 * k, 4096)));
    return lp;
}
// The single pass of phase B over rows [0, N): slices as for any epoch, slabs provisioned for 3 x 2.2 k survivors per query.
inline EpochPlan plan_large_k_pass(long long N, int k, int nq_pad, const GridShape& gs, int n_sms) {
    EpochPlan ep;
    ep.begin = 0;
    ep.end = N;
    pick_slices(gs, nq_pad / kPlanQueryTile, (N + kPlanTileRows - 1) / kPlanTileRows, n_sms, &ep.s1, &ep.s0);
    // about 2.4 k survivors over N rows: with k in the thousands most chunks have one
    ep.sets = 1024.0 * 2.4 * (double)k / (double)std::max(1LL, N) > 0.10 ? kPlanSetsLoose : kPlanSetsTight;
    const double slabs = (double)std::min(ep.s1, ep.s0) * ep.sets;
    const double expect = (k <= 1024 ? 2.6 : 2.2) * (double)k / slabs;
    ep.cap = std::min(65536, std::max(128 / ep.sets, plan_next_pow2((int)(3.0 * expect) + 128 / ep.sets)));
    return ep;
}
// Large enough a corpus for the sample to mean something; otherwise the fp32 scan answers.
inline bool plan_large_k_applies(long long N, int k) { return N >= 64LL * k; }

inline int plan_n_ctas(const GridShape& gs, const EpochPlan& ep) { return gs.rem * ep.s1 + (gs.n_groups - gs.rem) * ep.s0; }
inline int plan_n_sub(const GridShape& gs, const EpochPlan& ep) { return std::max(ep.s1, ep.s0) * ep.sets; }

}  // namespace pq
