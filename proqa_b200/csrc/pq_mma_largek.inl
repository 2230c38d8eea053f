// proqa_b200 — tensor-core tier for 1024 < k <= PQ_MAX_K (included by pq_mma.cu inside namespace pq).
//
// Default for 1024 < k on corpora of at least 64 k rows (pq_index.cu: tier_uses_largek; PROQA_B200_LARGEK=0 sends such k back
// to the fp32 scan).  Validated on a B200 in round 2 (tests/test_gpu_largek.py, compute-sanitizer memcheck + racecheck clean;
// 256 queries x 8.8M rows x k = 10000: 3.8 ms on the device against ~6 ms per query on the scan).  The statistics it relies
// on (rank of the sample threshold, survivors per query, slab occupancy, size of the rescored set, certificate) were replayed
// on CPU for 8.8M rows / k = 10000 on exchangeable and on document-ordered rows: tools/sim_largek.py, DESIGN.md §5.6.
//
// Call site served: retrieval/trec_process.py:76 — index.search(xq, 10000) over the MS MARCO passages.
//
//   A. thresholds   ordinary epoch search (filter kernel + pq_epoch_select_kernel, k_sample <= 1024, no second attempts)
//                   over a compact copy of every step-th row: leaves  T = A_k_sample(sample) - 2E  in the query state,
//                   a bf16 score near full-corpus rank 1.35 k
//   B. one pass     pq_mma_filter_kernel over all rows admitting bf16 score >= T into the candidate slabs
//   C. finalize     pq_largek_finalize_kernel, one CTA per query:
//                     A_k  = k-th best bf16 score of the survivors   (radix select over the slabs where they lie)
//                     certificate  T <= A_k - 2E : every row the filter refused has bf16 score < T, hence exact score
//                       < T + E <= A_k - E <= the k-th best exact score (k rows have bf16 score >= A_k)
//                     survivors with bf16 score >= A_k - 2E  -> shared memory, exact fp32 rescoring (engine_dot order),
//                     radix select of the k best exact keys, sort, emit
//                   anything that does not fit (slab or pool overflow, fewer than k survivors, a failed certificate)
//                   marks the query for the fp32 scan, like the k <= 1024 path does.

// ------------------------------------------------------------------------------------------------
// sample copy: every step-th row (bf16) and its squared norm, compact
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pq_gather_sample_kernel(const uint16_t* __restrict__ rows_bf16, const float* __restrict__ norms,
                                                               long long n_sample, int step, uint16_t* __restrict__ out_bf16,
                                                               float* __restrict__ out_norms) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp; i < n_sample; i += n_warps) {
        const long long row = i * step + step / 2;
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(rows_bf16 + row * kDim) + lane);   // 32 lanes x 8 B = one 256-B row
        reinterpret_cast<uint2*>(out_bf16 + i * kDim)[lane] = v;
        if (lane == 0) out_norms[i] = norms[row];
    }
}

static int largek_ensure_sample(pq_index* ix, const LargeKPlan& lp) {
    if (ix->sample_rows == lp.sample_rows && ix->sample_step == lp.step) return PQ_OK;
    const long long ns = lp.sample_rows;
    const long long ns_pad = (ns + kBN - 1) / kBN * kBN;
    int rc = ix->sample_bf16.ensure((size_t)ns_pad * kDim * 2);
    if (!rc) rc = ix->sample_norms.ensure((size_t)(ns_pad + kBN) * 4);   // the L2 epilogue reads a whole tile of norms
    if (rc) return rc;
    PQ_CUDA(cudaMemsetAsync(ix->sample_norms.p, 0, (size_t)(ns_pad + kBN) * 4, ix->stream));
    const int blocks = (int)std::min<long long>((ns + 7) / 8, (long long)ix->n_sms * 16);
    pq_gather_sample_kernel<<<blocks, 256, 0, ix->stream>>>((const uint16_t*)ix->rows_bf16.p, (const float*)ix->norms.p, ns, lp.step,
                                                            (uint16_t*)ix->sample_bf16.p, (float*)ix->sample_norms.p);
    PQ_CUDA(cudaGetLastError());
    rc = make_row_tensor_map(&ix->tmap_sample, ix->sample_bf16.p, ns, 2, 64, 128);
    if (rc) return rc;
    ix->sample_rows = ns;
    ix->sample_step = lp.step;
    ix->stats[5] += 1;
    return PQ_OK;
}

// ------------------------------------------------------------------------------------------------
// finalize
// ------------------------------------------------------------------------------------------------
struct LargeKParams {
    const uint64_t* cand_keys;   // [nq_pad][n_sub][cap] survivors of the single pass
    const uint32_t* cand_cnt;    // [nq_pad][n_sub]
    const float* thr;            // [nq_pad] the threshold the pass admitted at
    const float* two_e;          // [nq_pad]
    const float* queries;        // [nq][128] fp32
    const float* rows;           // [ntotal][128] fp32
    const float* row_norms;
    const float* q_norms;
    const uint8_t* q_bad;
    int n_sub, cap, k, pool, sort_n, metric;
    long long id_base;
    float* D;
    long long* I;
    uint8_t* fail;
    uint32_t* fail_count;
};

// k-th largest 32-bit ordered score among the slab entries of one query, MSB-first radix select straight over the slabs
// (one warp per slab, coalesced).  Bytes on which every score agrees are skipped; inside a warp, lanes with the same
// digit are combined before the shared-memory atomic (the scores above a threshold share their leading bits).
template <int kT>
__device__ __forceinline__ uint32_t slab_radix_select_score(const uint64_t* keys, const int* s_cnt, int n_sub, int cap, int want, int* hist,
                                                            uint32_t* s_u32, int* s_int) {
    constexpr int kW = kT / 32;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint32_t a = ~0u, o = 0u;
    for (int s = warp; s < n_sub; s += kW) {
        const uint64_t* src = keys + (size_t)s * cap;
        const int n = s_cnt[s];
        for (int p0 = 0; p0 < n; p0 += 128) {  // four loads in flight per lane (the slabs are read from L2: one round trip each)
            uint64_t kv[4];
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) kv[u4] = p0 + 32 * u4 + lane < n ? src[p0 + 32 * u4 + lane] : 0ull;
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
                if (p0 + 32 * u4 + lane < n) {
                    const uint32_t u = (uint32_t)(kv[u4] >> 32);
                    a &= u;
                    o |= u;
                }
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        a &= __shfl_xor_sync(0xffffffffu, a, s);
        o |= __shfl_xor_sync(0xffffffffu, o, s);
    }
    if (lane == 0) {
        s_u32[warp] = a;
        s_u32[kW + warp] = o;
    }
    __syncthreads();
    a = s_u32[0];
    o = s_u32[kW];
#pragma unroll
    for (int w = 1; w < kW; ++w) {
        a &= s_u32[w];
        o |= s_u32[kW + w];
    }
    const uint32_t differ = a ^ o;
    uint32_t prefix = a & ~differ, mask = ~differ;
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
        const uint32_t dmask = (differ >> shift) & 0xffu;
        if (dmask == 0) continue;  // block-uniform
        if (t < 256) hist[t] = 0;
        __syncthreads();
        for (int s = warp; s < n_sub; s += kW) {
            const uint64_t* src = keys + (size_t)s * cap;
            const int n = s_cnt[s];
            for (int q0 = 0; q0 < n; q0 += 128) {  // warp-uniform trip counts; four loads in flight per lane
                uint64_t kv[4];
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) kv[u4] = q0 + 32 * u4 + lane < n ? src[q0 + 32 * u4 + lane] : 0ull;
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    const int p0 = q0 + 32 * u4;
                    if (p0 >= n) break;
                    const int pos = p0 + lane;
                    uint32_t u = 0;
                    bool in = false;
                    if (pos < n) {
                        u = (uint32_t)(kv[u4] >> 32);
                        in = ((u ^ prefix) & mask) == 0;
                    }
                    const unsigned act = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const int digit = (int)((u >> shift) & dmask);
                        const unsigned peers = __match_any_sync(act, digit);
                        if (lane == __ffs(peers) - 1) atomicAdd(&hist[digit], __popc(peers));
                    }
                }
            }
        }
        __syncthreads();
        if (warp == 0) {  // largest digit d with count(digit >= d) >= want
            int loc = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) loc += hist[lane * 8 + b];
            int suf = loc;
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
        prefix |= (uint32_t)s_int[0] << shift;
        mask |= dmask << shift;
        want = s_int[1];
        __syncthreads();
    }
    return prefix;
}

// kLargeKFinT threads: the kernel is a chain of dependent global reads (slab passes, then the row gather of the rescoring) with
// one CTA per SM because of its shared-memory pool — 8 warps left the SM waiting on memory 90 % of the time (ncu: issue slots
// 10 % busy, 65 % of the stall samples long_scoreboard), 32 warps keep four times the loads in flight.
constexpr int kLargeKFinT = 1024;
template <int kT>
__global__ void __launch_bounds__(kT) pq_largek_finalize_kernel(const LargeKParams p) {
    constexpr int kW = kT / 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* pool = reinterpret_cast<uint64_t*>(smem_raw);                       // max(pool, sort_n) keys
    int* s_cnt = reinterpret_cast<int*>(pool + max(p.pool, p.sort_n));            // n_sub
    __shared__ float s_q[kDim];
    __shared__ int hist[256];
    __shared__ uint64_t s_u64[2 * kW];
    __shared__ int s_int[2];
    __shared__ int s_wbase[kW];
    __shared__ int s_total, s_ovf, s_slot, s_valid;
    const int q = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint64_t* keys = p.cand_keys + (size_t)q * p.n_sub * p.cap;
    const uint32_t* cnts = p.cand_cnt + (size_t)q * p.n_sub;
    if (t == 0) {
        s_total = 0;
        s_ovf = 0;
        s_slot = 0;
        s_valid = 0;
    }
    if (t < kDim) s_q[t] = p.queries[(size_t)q * kDim + t];
    __syncthreads();
    {
        int mine = 0;
        for (int s = t; s < p.n_sub; s += kT) {
            const uint32_t c = cnts[s];
            if (c > (uint32_t)p.cap) s_ovf = 1;
            s_cnt[s] = (int)min(c, (uint32_t)p.cap);
            mine += s_cnt[s];
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, s);
        if (lane == 0 && mine) atomicAdd(&s_total, mine);
    }
    __syncthreads();
    const float two_e = p.two_e[q];
    // (all conditions below are block-uniform)
    bool fail = p.q_bad[q] != 0 || !isfinite(two_e) || s_ovf != 0 || s_total < p.k;
    float bar = 0.f;
    if (!fail) {
        // ---- A_k and the certificate ----
        const uint32_t kth = slab_radix_select_score<kT>(keys, s_cnt, p.n_sub, p.cap, p.k, hist, reinterpret_cast<uint32_t*>(s_u64), s_int);
        bar = ordered_to_f32(kth) - two_e;
        fail = !(p.thr[q] <= bar);
    }
    if (!fail) {
        // ---- survivors within 2E of A_k -> shared memory ----
        for (int s = warp; s < p.n_sub; s += kW) {
            const uint64_t* src = keys + (size_t)s * p.cap;
            const int n = s_cnt[s];
            for (int q0 = 0; q0 < n; q0 += 128) {
                uint64_t kv[4];
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) kv[u4] = q0 + 32 * u4 + lane < n ? src[q0 + 32 * u4 + lane] : 0ull;
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    const int p0 = q0 + 32 * u4;
                    if (p0 >= n) break;
                    const int pos = p0 + lane;
                    const uint64_t key = kv[u4];
                    const bool hit = pos < n && key_score(key) >= bar;
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    int base = 0;
                    if (lane == 0 && m) base = atomicAdd(&s_slot, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const int dst = base + __popc(m & ((1u << lane) - 1u));
                    if (hit && dst < p.pool) pool[dst] = key;
                }
            }
        }
        __syncthreads();
        fail = s_slot > p.pool;
    }
    int n_r = 0;
    if (!fail) {
        // ---- exact scores, in place ----
        // The engine's defined score (pq_common.cuh: engine_dot / quad_engine_dot): lane = 8 r + j computes chain p_j (dims
        // 16 j .. 16 j + 15 ascending) of row r of its group of four from its own 64 bytes of the row; three butterfly steps add
        // the chains in engine_dot's tree.  A warp takes EIGHT rows per pass (two groups): eight independent 16-byte loads per
        // lane, 4 KB per warp, in flight together; the lane's 16 query values stay in registers.
        n_r = s_slot;
        int valid = 0;
        const int j = lane & 7, r = lane >> 3;
        float4 qv[4];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) qv[i4] = reinterpret_cast<const float4*>(s_q + 16 * j)[i4];
        for (int i0 = warp * 8; i0 < n_r; i0 += kW * 8) {
            const int ia = i0 + r, ib = i0 + 4 + r;
            const bool has_a = ia < n_r, has_b = ib < n_r;
            const uint32_t row_a = has_a ? key_row(pool[ia]) : 0u, row_b = has_b ? key_row(pool[ib]) : 0u;
            const float4* ra = reinterpret_cast<const float4*>(p.rows + (size_t)row_a * kDim + 16 * j);
            const float4* rb = reinterpret_cast<const float4*>(p.rows + (size_t)row_b * kDim + 16 * j);
            float4 va[4], vb[4];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) va[i4] = ldg_f4_now(ra + i4);   // (row 0 stands in for a missing row: read, not used)
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) vb[i4] = ldg_f4_now(rb + i4);
            float na = 0.f, nb = 0.f;
            if (p.metric == kMetricL2 && j == 0) {
                na = __ldg(p.row_norms + row_a);
                nb = __ldg(p.row_norms + row_b);
            }
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                a = fmaf(va[i4].x, qv[i4].x, a);
                a = fmaf(va[i4].y, qv[i4].y, a);
                a = fmaf(va[i4].z, qv[i4].z, a);
                a = fmaf(va[i4].w, qv[i4].w, a);
                b = fmaf(vb[i4].x, qv[i4].x, b);
                b = fmaf(vb[i4].y, qv[i4].y, b);
                b = fmaf(vb[i4].z, qv[i4].z, b);
                b = fmaf(vb[i4].w, qv[i4].w, b);
            }
#pragma unroll
            for (int sft = 1; sft <= 4; sft <<= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, sft);
                b += __shfl_xor_sync(0xffffffffu, b, sft);
            }
            if (j == 0) {
                if (p.metric == kMetricL2) {
                    a = fmaf(2.f, a, -na);
                    b = fmaf(2.f, b, -nb);
                }
                if (has_a) {
                    const bool ok = a >= PQ_THR_FLOOR;
                    pool[ia] = ok ? make_key(a, row_a) : 0ull;
                    valid += ok ? 1 : 0;
                }
                if (has_b) {
                    const bool ok = b >= PQ_THR_FLOOR;
                    pool[ib] = ok ? make_key(b, row_b) : 0ull;
                    valid += ok ? 1 : 0;
                }
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, s);
        if (lane == 0 && valid) atomicAdd(&s_valid, valid);
        __syncthreads();
        fail = s_valid < p.k;  // a survivor's exact score was not a number: let the scan decide
    }
    if (fail) {
        if (t == 0) {
            p.fail[q] = 1;
            atomicAdd(p.fail_count, 1u);
        }
        return;
    }
    // ---- the k best exact keys to the front, in place ----
    if (n_r > p.k) {
        const uint64_t pivot = block_radix_select<kT>(pool, n_r, p.k, hist, s_u64, s_int);  // k-th largest; the keys >= it are unique
        if (t == 0) s_slot = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n_r; i0 += kT) {   // ordered compaction: a chunk is read by everyone before anyone writes at or below it
            const int i = i0 + t;
            const uint64_t key = i < n_r ? pool[i] : 0ull;
            const bool keep = i < n_r && key >= pivot;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_wbase[warp] = __popc(m);
            __syncthreads();
            int before = s_slot;
            for (int w = 0; w < warp; ++w) before += s_wbase[w];
            if (keep) pool[before + __popc(m & ((1u << lane) - 1u))] = key;
            __syncthreads();
            if (t == 0) {
                int tot = 0;
                for (int w = 0; w < kW; ++w) tot += s_wbase[w];
                s_slot += tot;
            }
            __syncthreads();
        }
    }
    for (int i = p.k + t; i < p.sort_n; i += kT) pool[i] = 0ull;
    __syncthreads();
    block_sort_desc<kT>(pool, p.sort_n);
    for (int i = t; i < p.k; i += kT) {
        const uint64_t key = pool[i];
        const float s = key_score(key);
        p.D[(size_t)q * p.k + i] = (p.metric == kMetricL2) ? fmaxf(0.f, p.q_norms[q] - s) : s;
        p.I[(size_t)q * p.k + i] = (long long)key_row(key) + p.id_base;
    }
    if (t == 0) p.fail[q] = 0;
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
int search_mma_largek(pq_index* ix, int nq_total, const float* dq_all, int k, float* dD_all, long long* dI_all, std::vector<int>* rerun) {
    const long long N = ix->ntotal;
    const LargeKPlan lp = plan_large_k(N, k);
    int rc = largek_ensure_sample(ix, lp);
    if (rc) return rc;
    const int kp_s = carry_size_for_k(lp.k_sample);
    const bool l2 = ix->metric == kMetricL2;

    for (int qb = 0; qb < nq_total; qb += kLargeKBatch) {
        const int nq = std::min(kLargeKBatch, nq_total - qb);
        const int nq_pad = (nq + kBM - 1) / kBM * kBM;
        const int n_mtiles = nq_pad / kBM;
        const GridShape gs = make_grid_shape(n_mtiles);
        const float* dq = dq_all + (size_t)qb * kDim;
        const uint16_t* dq_bf16 = (const uint16_t*)ix->ws_qbf16.p + (size_t)qb * kDim;
        const float* dq_norm = (const float*)ix->ws_qnorm.p + qb;
        const uint8_t* dq_bad = (const uint8_t*)ix->ws_qbad.p + qb;

        const std::vector<EpochPlan> plan_s = plan_epochs(lp.sample_rows, lp.k_sample, nq_pad, gs, ix->n_sms);
        const EpochPlan pass = plan_large_k_pass(N, k, nq_pad, gs, ix->n_sms);
        size_t max_slab = (size_t)nq_pad * plan_n_sub(gs, pass) * pass.cap * 8, max_cnt = (size_t)nq_pad * plan_n_sub(gs, pass) * 4;
        for (const EpochPlan& ep : plan_s) {
            const size_t n_sub = (size_t)plan_n_sub(gs, ep);
            max_slab = std::max(max_slab, (size_t)nq_pad * n_sub * ep.cap * 8);
            max_cnt = std::max(max_cnt, (size_t)nq_pad * n_sub * 4);
        }

        DevBuf* w = ix->ws_mma;
        rc = w[0].ensure((size_t)nq_pad * 4);                     // thr
        if (!rc) rc = w[1].ensure((size_t)nq_pad * 4);            // two_e
        if (!rc) rc = w[2].ensure((size_t)nq_pad * 4);            // dropmax (sample phase bookkeeping only)
        if (!rc) rc = w[3].ensure((size_t)nq_pad * 4);            // overflow (same)
        if (!rc) rc = w[4].ensure((size_t)nq_pad * kp_s * 8);     // carry of the sample search
        if (!rc) rc = w[5].ensure(max_slab);                      // candidate slabs
        if (!rc) rc = w[6].ensure(max_cnt);                       // slab counts
        if (!rc) rc = w[7].ensure((size_t)nq_pad);                // fail flags
        if (!rc) rc = w[8].ensure((size_t)nq_pad * 4 + 256);      // redo flags (unused here), then the counters
        if (rc) return rc;
        QState st;
        st.thr = (float*)w[0].p;
        st.two_e = (float*)w[1].p;
        st.dropmax = (float*)w[2].p;
        st.overflow = (uint32_t*)w[3].p;
        st.carry = (uint64_t*)w[4].p;
        st.redo = (uint32_t*)w[8].p;
        st.counters = st.redo + nq_pad;

        PQ_CUDA(cudaMemsetAsync(st.carry, 0, (size_t)nq_pad * kp_s * 8, ix->stream));
        PQ_CUDA(cudaMemsetAsync(st.counters, 0, 16, ix->stream));
        pq_mma_init_state_kernel<<<(nq_pad + 255) / 256, 256, 0, ix->stream>>>(st, dq_norm, (const float*)ix->ws_qresid.p + qb, dq_bad, nq, nq_pad, kp_s,
                                                                             ix->max_norm2, ix->max_resid2, ix->metric);
        PQ_CUDA(cudaGetLastError());
        ix->stats[5] += 1;

        MmaParams mp;
        mp.q_bf16 = dq_bf16;
        mp.cand_keys = (uint64_t*)w[5].p;
        mp.cand_cnt = (uint32_t*)w[6].p;
        mp.thr = st.thr;
        mp.two_e = st.two_e;
        mp.n_mtiles = n_mtiles;
        mp.base = gs.base;
        mp.rem = gs.rem;
        mp.k1_adapt = 0;
        mp.redo = nullptr;
        mp.redo_bit = 0u;
        mp.pace = nullptr;   // (the sample epochs are small; the full pass below sets it up)
        mp.pace_shift = 0;
        mp.pace_blocks = 0;
        mp.pace_cohort = 1;

        // ---- A. thresholds from the sample ----------------------------------------------------
        for (const EpochPlan& ep : plan_s) {
            mp.row_norms = (const float*)ix->sample_norms.p;
            mp.row_begin = ep.begin;
            mp.row_end = ep.end;
            mp.s1 = ep.s1;
            mp.s0 = ep.s0;
            mp.cap = ep.cap;
            mp.n_sub = plan_n_sub(gs, ep);
            mp.sets = ep.sets;
            PQ_CUDA(launch_filter_any(gs.m_max, l2, false, ix->tmap_sample, mp, plan_n_ctas(gs, ep), ix->device, ix->stream));
            EpochSelParams sp;
            memset(&sp, 0, sizeof(sp));
            sp.st = st;
            sp.cand_keys = mp.cand_keys;
            sp.cand_cnt = mp.cand_cnt;
            sp.n_sub = mp.n_sub;
            sp.cap = ep.cap;
            sp.kp = kp_s;
            sp.k = lp.k_sample;
            sp.lmax = std::max(2 * kp_s, 4096);
            sp.nq = nq;
            sp.base = gs.base;
            sp.rem = gs.rem;
            sp.s1 = ep.s1;
            sp.s0 = ep.s0;
            sp.sets = ep.sets;
            sp.is_redo = 0;
            sp.allow_redo = 0;   // a slab overflow only loosens the estimate (the k-th best of what fitted is still a real score)
            sp.epoch_bit = 0u;
            sp.row_begin = ep.begin;
            sp.row_end = ep.end;
            PQ_CUDA(launch_epoch_select(sp, ix->n_sms, ix->device, ix->stream));
            ix->stats[3] += 1;
            ix->stats[4] += 1;
            ix->stats[5] += 2;
        }

        // ---- B. one pass over all rows at those thresholds --------------------------------------
        mp.row_norms = (const float*)ix->norms.p;
        mp.row_begin = 0;
        mp.row_end = N;
        mp.s1 = pass.s1;
        mp.s0 = pass.s0;
        mp.cap = pass.cap;
        mp.n_sub = plan_n_sub(gs, pass);
        mp.sets = pass.sets;
        PQ_CUDA(cudaMemsetAsync(mp.cand_cnt, 0, (size_t)nq_pad * mp.n_sub * 4, ix->stream));
        PaceArea pace;
        const long long pace_blocks = pace_blocks_for(ix, gs, plan_n_ctas(gs, pass), 0, N);
        if (const int prc = pace.reserve(ix, pace_blocks * pace_cohorts(ix, plan_n_ctas(gs, pass)))) return prc;
        pace.assign(ix, mp, pace_blocks, plan_n_ctas(gs, pass));
        ix->prof_begin();
        const cudaError_t e = launch_filter_any(gs.m_max, l2, false, ix->tmap_bf16, mp, plan_n_ctas(gs, pass), ix->device, ix->stream);
        ix->prof_end();
        PQ_CUDA(e);
        ix->stats[3] += 1;
        ix->stats[5] += 1;

        // ---- C. finalize ------------------------------------------------------------------------
        LargeKParams fp;
        fp.cand_keys = mp.cand_keys;
        fp.cand_cnt = mp.cand_cnt;
        fp.thr = st.thr;
        fp.two_e = st.two_e;
        fp.queries = dq;
        fp.rows = (const float*)ix->rows_f32.p;
        fp.row_norms = (const float*)ix->norms.p;
        fp.q_norms = dq_norm;
        fp.q_bad = dq_bad;
        fp.n_sub = mp.n_sub;
        fp.cap = pass.cap;
        fp.k = k;
        fp.pool = lp.pool;
        fp.sort_n = lp.sort_n;
        fp.metric = ix->metric;
        fp.id_base = ix->id_base;
        fp.D = dD_all + (size_t)qb * k;
        fp.I = dI_all + (size_t)qb * k;
        fp.fail = (uint8_t*)w[7].p;
        fp.fail_count = st.counters + 2;
        const size_t fsmem = (size_t)std::max(lp.pool, lp.sort_n) * 8 + (size_t)fp.n_sub * 4;
        PQ_CUDA(ensure_dyn_smem(pq_largek_finalize_kernel<kLargeKFinT>, fsmem, ix->device));
        pq_largek_finalize_kernel<kLargeKFinT><<<nq, kLargeKFinT, fsmem, ix->stream>>>(fp);
        PQ_CUDA(cudaGetLastError());
        ix->stats[4] += 1;
        ix->stats[5] += 1;

        uint32_t counts[3] = {0, 0, 0};
        PQ_CUDA(cudaMemcpyAsync(counts, st.counters, 12, cudaMemcpyDeviceToHost, ix->stream));
        PQ_CUDA(cudaStreamSynchronize(ix->stream));
        if (counts[2] != 0) {
            std::vector<uint8_t> fail((size_t)nq);
            PQ_CUDA(cudaMemcpy(fail.data(), w[7].p, (size_t)nq, cudaMemcpyDeviceToHost));
            for (int q = 0; q < nq; ++q)
                if (fail[q]) rerun->push_back(qb + q);
        }
    }
    return PQ_OK;
}
