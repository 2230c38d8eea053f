"""Test infrastructure: build and call the CPU emulations of the tensor-core tier (see simt_emu.h, mma_host_emu.cpp.in).

The engine's own source text is extracted from proqa_b200/csrc (never copied into the repo a second time), `__shared__`
becomes static, `kernel<<<...>>>(...)` becomes EMU_LAUNCH, and the result is compiled with g++ into a throw-away .so.
Nothing here is importable by the product; it exists so that `-m "not gpu"` can execute kernel and driver LOGIC against
the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "proqa_b200", "csrc")
SIMT = os.path.join(ROOT, "tests", "simt")


def extract(text, signature, upto=None):
    """The top-level definition whose first line contains `signature` (plus a preceding template<> line), through the first
    line that is exactly '}' or '};' — or, when `upto` is given, through the first line containing it."""
    lines = text.split("\n")
    start = next(i for i, ln in enumerate(lines) if signature in ln)
    while start > 0 and (lines[start - 1].startswith("template") or lines[start - 1].startswith("__global__")):
        start -= 1
    end = start
    while not ((upto in lines[end]) if upto else lines[end] in ("}", "};")):
        end += 1
    return "\n".join(lines[start:end + 1]) + "\n"


def to_host(src):
    src = re.sub(r"extern __shared__ __align__\(16\)", "extern", src)
    src = src.replace("__shared__", "static")
    # (the kernel name goes in parentheses: template arguments contain commas the preprocessor would split on)
    return re.sub(r"(\w+(?:<[\w, ]+>)?)<<<(.*?)>>>\((.*?)\);", r"EMU_LAUNCH((\1), \2, \3);", src, flags=re.S)


def sources():
    return (open(os.path.join(CSRC, "pq_common.cuh")).read(), open(os.path.join(CSRC, "pq_mma.cu")).read(),
            open(os.path.join(CSRC, "pq_mma_largek.inl")).read())


def key_and_sort_helpers(common):
    return [extract(common, "uint32_t f32_to_ordered(float f)", upto="uint32_t key_row(uint64_t key)"),
            extract(common, "float warp_engine_dot(const float4 a, const float4 b, int lane)"),
            extract(common, "float quad_engine_dot(const float* __restrict__ row, const float* q_smem, int lane)"),
            extract(common, "void block_sort(uint64_t* a, int n, bool ascending)"),
            extract(common, "void block_sort_desc(uint64_t* a, int n)")]


def compile_so(cpp_text, workdir, name, opt="-O1", extra=()):
    cpp = os.path.join(str(workdir), name + ".cpp")
    open(cpp, "w").write(cpp_text)
    so = os.path.join(str(workdir), name + ".so")
    r = subprocess.run(["g++", opt, "-std=c++17", "-shared", "-fPIC", "-Wno-attributes", "-I", SIMT, *extra, cpp, "-o", so],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return ctypes.CDLL(so)


def build_host_emu(workdir, real_filter=False, mutate=None):
    """The tensor-tier host drivers (search_mma_filter, search_mma_largek) + their kernels, emulated.  The tcgen05 filter
    kernel is either replaced by a functional stand-in (filter_functional.inc: fast) or executed itself on models of
    mbarriers / TMA / TMEM / tcgen05.mma (filter_tcgen05.inc, real_filter=True)."""
    common, mma, inl = sources()
    if real_filter:
        kernel = "\n".join([
            extract(mma, "constexpr int kBM = 128;", upto="static_assert(kBM == kPlanQueryTile"),
            extract(mma, "struct MmaCtrl {"),
            extract(common, "uint64_t umma_desc_k128(uint32_t smem_addr)"),
            extract(common, "constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N)"),
            extract(mma, "void mma_apply_l2_bias(float (&v)[32], const float4* norms)"),
            extract(mma, "void mma_mask_tail(float (&v)[32], uint32_t base_row, uint32_t row_end32)"),
            extract(mma, "void mma_filter32(float (&v)[32], float th, uint32_t base_row"),
            extract(mma, "float mma_max3(float a, float b, float c)"),
            extract(mma, "float mma_max_reduce(float (&a)[N])"),
            extract(mma, "float mma_max_all(float (&v)[kChunks][32])"),
            extract(mma, "void mma_filter32_k1(float (&v)[32], float& thr, float two_e"),
            extract(mma, "constexpr int kPaceWindow = 3;", upto="constexpr int kPaceMaxPolls = 4096;"),
            extract(mma, "void pace_leave(uint32_t* pace, int from, int to)"),
            extract(mma, "pq_mma_filter_kernel(const __grid_constant__ CUtensorMap tmap_c, const MmaParams p)")
            .replace("extern __shared__ __align__(1024) uint8_t smem[];", "uint8_t* smem = smem_raw;"),
            extract(mma, "static cudaError_t launch_filter(const CUtensorMap& tc"),
            extract(mma, "static cudaError_t launch_filter_s(const CUtensorMap& tc"),
            extract(mma, "static cudaError_t launch_filter_m(int m_max"),
            extract(mma, "static cudaError_t launch_filter_any(int m_max"),
        ])
        if mutate:   # negative tests: (old, new) applied to the kernel text, must hit exactly once
            assert kernel.count(mutate[0]) == 1, f"{mutate[0]!r} occurs {kernel.count(mutate[0])} times"
            kernel = kernel.replace(mutate[0], mutate[1])
        filter_impl = open(os.path.join(SIMT, "filter_tcgen05.inc")).read().replace("@FILTER_KERNEL@", to_host(kernel))
    else:
        filter_impl = open(os.path.join(SIMT, "filter_functional.inc")).read()
    device = "\n".join(key_and_sort_helpers(common) + [
        extract(mma, "struct MmaParams {"),
        extract(mma, "struct QState {"),
        extract(mma, "__global__ void pq_mma_init_state_kernel(QState st"),
        extract(mma, "constexpr int kShareMaxPeers = 16;", upto="constexpr int kShareMaxPeers = 16;"),
        extract(mma, "struct ShareParams {"),
        extract(mma, "uint64_t share_word(uint32_t tag, float v)", upto="uint64_t share_word(uint32_t tag, float v)"),
        extract(mma, "void share_publish(const ShareParams& sh, int q, float a, float b, int lane)"),
        extract(mma, "pq_share_fold_kernel(const ShareParams sh, QState st, int nq, int epoch)"),
        extract(mma, "struct EpochSelParams {"),
        extract(mma, "int sel_slabs_of_query(const EpochSelParams& p, int q)"),
        extract(mma, "uint64_t block_radix_select(const uint64_t* pool"),
        extract(mma, "pq_epoch_select_kernel(const EpochSelParams p)"),
        extract(mma, "constexpr int kSelWarpPool = 704;", upto="constexpr int kSelWarps = 8;"),
        extract(mma, "int warp_sum(int v)"),
        extract(mma, "uint64_t warp_radix_select(const uint64_t* pool, int n, int want, int* hist, int lane)"),
        extract(mma, "struct WarpSelState {"),
        extract(mma, "int warp_hist_rank(const int* hist, int want, int lane)"),
        extract(mma, "bool warp_binned_bounds(const EpochSelParams& p"),
        extract(mma, "int warp_sel_reduce(const EpochSelParams& p"),
        extract(mma, "pq_epoch_select_warp_kernel(const EpochSelParams p)"),
        extract(mma, "struct RescoreParams {"),
        extract(mma, "bool rescore_certificate_fails(const RescoreParams& p, int q)"),
        extract(mma, "void rescore_emit(const RescoreParams& p, int q, int i, uint64_t key)"),
        extract(mma, "pq_rescore_kernel(const RescoreParams p)"),
        extract(mma, "pq_rescore_warp_kernel(const RescoreParams p)"),
        extract(mma, "struct K1Params {"),
        extract(mma, "pq_k1_finalize_kernel(const K1Params p)"),
    ])
    host = (extract(mma, "static int next_pow2i(int v)") + extract(mma, "static cudaError_t ensure_dyn_smem_impl(const void* fn")
            + extract(mma, "static cudaError_t ensure_dyn_smem(K* kernel, size_t smem, int device)")
            + extract(mma, "static int sel_warp_min_queries()")
            + extract(mma, "static int warp_query_grid(int nq, int n_sms)")
            + extract(mma, "static cudaError_t launch_epoch_select(const EpochSelParams& sp")
            + extract(mma, "static cudaError_t launch_rescore(const RescoreParams& rp")
            + extract(mma, "static ShareParams make_share_params(const pq_index* ix")
            + extract(mma, "static int pace_shift() {")
            + extract(mma, "static int pace_cohorts(const pq_index* ix, int n_ctas)", upto="static int pace_cohorts(const pq_index* ix, int n_ctas)")
            + extract(mma, "static long long pace_blocks_for(const pq_index* ix")
            + extract(mma, "struct PaceArea {")
            + extract(mma, "int search_mma_filter(pq_index* ix"))
    tmpl = open(os.path.join(SIMT, "mma_host_emu.cpp.in")).read()
    text = (tmpl.replace("@EXTRACTED_DEVICE@", to_host(device)).replace("@FILTER_IMPL@", filter_impl)
            .replace("@EXTRACTED_HOST@", to_host(host)).replace("@EXTRACTED_LARGEK@", to_host(inl)))
    # -Bsymbolic: our fake CUDA runtime, not a libcudart some other module of the test process has loaded
    lib = compile_so(text, workdir, "mma_host_emu_real" if real_filter else "mma_host_emu", opt="-O2", extra=("-I", CSRC, "-I", "/usr/local/cuda/include", "-Wl,-Bsymbolic"))
    vp = ctypes.c_void_p
    lib.emu_set_share.restype = None
    lib.emu_set_share.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_longlong, vp]
    lib.emu_search_mma.restype = ctypes.c_char_p
    lib.emu_search_mma.argtypes = [vp, vp, vp, ctypes.c_longlong, ctypes.c_float, ctypes.c_float, vp, vp, vp, vp, vp,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]
    return lib


# ---- what add() and the query preparation leave on the device, computed on the host ----------------------------------------
def bf16_round(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def bf16_bits(x):
    return (bf16_round(x).view(np.uint32) >> 16).astype(np.uint16)


def f32_ordered(f):
    u = np.asarray(f, np.float32).view(np.uint32)
    return np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)


def engine_norms(x):
    f32p = ctypes.POINTER(ctypes.c_float)
    L = oracle._lib()
    return np.array([L.engine_chain_dot(r.ctypes.data_as(f32p), r.ctypes.data_as(f32p), 128) for r in x], np.float32)


def resid2(x):
    return (((x - bf16_round(x)).astype(np.float64) ** 2).sum(1) * 1.0001).astype(np.float32)


def run_host_emu(lib, xb, xq, k, metric, n_sms=8, schedule=0, share=None, bound=None, force_largek=False):
    """-> D, I, sorted list of queries the driver wants re-run by the fp32 scan, the driver's statistics.
    schedule != 0: the emulator resumes the threads of a block in a different pseudo-random order at every pass."""
    lib.emu_set_schedule(schedule)
    nq = len(xq)
    nq_pad = (nq + 127) // 128 * 128
    xb = np.ascontiguousarray(xb, np.float32)
    xq = np.ascontiguousarray(xq, np.float32)
    xb_b = bf16_bits(xb)
    xq_b = np.zeros((nq_pad, 128), np.uint16)
    xq_b[:nq] = bf16_bits(xq)
    norms = np.zeros(len(xb) + 256, np.float32)         # (add() keeps the norm buffer padded to whole tiles)
    norms[:len(xb)] = engine_norms(xb)
    q_norm = np.zeros(nq_pad, np.float32)
    q_norm[:nq] = engine_norms(xq)
    q_resid = np.zeros(nq_pad, np.float32)
    q_resid[:nq] = resid2(xq)
    q_bad = np.zeros(nq_pad, np.uint8)
    D = np.full((nq, k), np.nan, np.float32)
    I = np.full((nq, k), -7, np.int64)
    rerun = np.zeros(nq, np.int32)
    n_rerun = ctypes.c_int(0)
    stats = np.zeros(10, np.int64)
    if share is not None:   # (n, rank, cap_q, wait_us, seq, id_base, [mailbox arrays]): this call is one row shard of n
        n_sh, rank, cap_q, wait_us, seq, id_base, boxes = share
        arr = (ctypes.c_void_p * 16)(*[b.ctypes.data for b in boxes])
        lib.emu_set_share(n_sh, rank, cap_q, wait_us, seq, id_base, arr)
    else:
        lib.emu_set_share(0, 0, 0, 0, 0, 0, None)
    lib.emu_force_largek(1 if force_largek else 0)
    max_norm2, max_resid2 = bound if bound is not None else (float(norms.max()), float(resid2(xb).max()))
    msg = lib.emu_search_mma(xb.ctypes.data, xb_b.ctypes.data, norms.ctypes.data, len(xb), max_norm2, max_resid2,
                             xq.ctypes.data, xq_b.ctypes.data, q_norm.ctypes.data, q_resid.ctypes.data, q_bad.ctypes.data, nq, k, metric,
                             n_sms, D.ctypes.data, I.ctypes.data, rerun.ctypes.data, ctypes.byref(n_rerun), stats.ctypes.data)
    assert msg is None, msg.decode()
    return D, I, sorted(rerun[:n_rerun.value].tolist()), stats


def build_select_emu(workdir, drop=None):
    """pq_select.cu: pq_merge_lists_kernel (fp32 scan: per-CTA lists -> result) and pq_merge_di_kernel (multi-GPU: shard
    results -> result) with their launchers.  `drop`: the beginning of one source line to leave out (negative tests)."""
    common = open(os.path.join(CSRC, "pq_common.cuh")).read()
    sel = open(os.path.join(CSRC, "pq_select.cu")).read()
    if drop:
        lines = sel.split("\n")
        hit = [i for i, ln in enumerate(lines) if ln.strip().startswith(drop)]
        assert len(hit) == 1, f"expected exactly one line starting with {drop!r}"
        del lines[hit[0]]
        sel = "\n".join(lines)
    parts = key_and_sort_helpers(common) + [
        extract(common, "void block_bitonic_merge_desc(uint64_t* a, int n)"),
        extract(sel, "constexpr int kSelThreads = 256;", upto="constexpr int kSelThreads = 256;"),
        extract(sel, "struct MergeParams {"),
        extract(sel, "void emit_result(const MergeLaunch& a, int q, int i, uint64_t key)"),
        extract(sel, "pq_merge_lists_kernel(const MergeParams p)"),
        extract(sel, "static int next_pow2(int v)"),
        extract(sel, "cudaError_t merge_lists_launch(const MergeLaunch& a, cudaStream_t stream)"),
        extract(sel, "struct MergeDIParams {"),
        extract(sel, "pq_merge_di_kernel(const MergeDIParams p)"),
        extract(sel, "constexpr int kMergeRankMaxLists = 64;", upto="constexpr int kMergeRankMaxLists = 64;"),
        extract(sel, "pq_merge_di_rank_kernel(const MergeDIParams p)"),
        extract(sel, "pq_prep_rows_kernel(const float* __restrict__ rows"),
        extract(sel, "cudaError_t prep_rows_launch(const float* rows"),
        extract(sel, "cudaError_t merge_di_launch(const float* D_in"),
    ]
    tmpl = open(os.path.join(SIMT, "select_emu.cpp.in")).read()
    compile_so(tmpl.replace("@EXTRACTED@", to_host("\n".join(parts))), workdir, "select_emu", opt="-O2",
               extra=("-I", CSRC, "-I", "/usr/local/cuda/include", "-Wl,-Bsymbolic"))
    return load_select_emu(os.path.join(str(workdir), "select_emu.so"))


def load_select_emu(path):
    lib = ctypes.CDLL(path)
    lib.path = path
    vp, ll, i32 = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    lib.emu_merge_di.restype = ctypes.c_char_p
    lib.emu_merge_di.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    lib.emu_prep_rows.restype = ctypes.c_char_p
    lib.emu_prep_rows.argtypes = [vp, ll, vp, vp, vp, vp, vp, vp, vp]
    lib.emu_merge_lists.restype = ctypes.c_char_p
    lib.emu_merge_lists.argtypes = [vp, ll, ll, i32, i32, vp, ll, vp, i32, i32, i32, vp, ll, vp, vp, vp]
    return lib


def merge_di(lib, D_all, I_all, k, metric):
    """pq_merge_shard_results on numpy arrays [G, nq, k] -> (D [nq,k], I [nq,k])."""
    D_all = np.ascontiguousarray(D_all, np.float32)
    I_all = np.ascontiguousarray(I_all, np.int64)
    G, nq, _ = D_all.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    msg = lib.emu_merge_di(D_all.ctypes.data, I_all.ctypes.data, G, nq, k, metric, D.ctypes.data, I.ctypes.data)
    assert msg is None, msg.decode()
    return D, I


def build_kmeans_emu(workdir):
    """pq_kmeans.cu: the single-GPU driver and the staged (multi-GPU) steps with their kernels (incl. the engine's own stable
    radix sort); the index replaced by a stand-in (kmeans_emu.cpp.in)."""
    common = open(os.path.join(CSRC, "pq_common.cuh")).read()
    km = open(os.path.join(CSRC, "pq_kmeans.cu")).read()
    helpers = "\n".join(key_and_sort_helpers(common)[:2] + [extract(common, "float engine_dot(const float* __restrict__ a")])
    a = km.index("namespace pq {") + len("namespace pq {")
    b = km.index("}  // namespace pq")
    tmpl = open(os.path.join(SIMT, "kmeans_emu.cpp.in")).read()
    text = tmpl.replace("@HELPERS@", to_host(helpers)).replace("@EXTRACTED@", to_host(km[a:b]))
    compile_so(text, workdir, "kmeans_emu", opt="-O2", extra=("-I", CSRC, "-I", "/usr/local/cuda/include", "-Wl,-Bsymbolic"))
    return load_kmeans_emu(os.path.join(str(workdir), "kmeans_emu.so"))


def load_kmeans_emu(path):
    lib = ctypes.CDLL(path)
    lib.path = path
    vp, ll, i32 = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    lib.emu_km_new.argtypes = [i32]
    lib.emu_km_train.restype = ctypes.c_char_p
    lib.emu_km_train.argtypes = [vp, ll, ll, i32, i32, i32, ll, vp, vp, ll, ctypes.POINTER(ll)]
    lib.emu_rand_perm.argtypes = [ll, ll, vp]
    for name, args in (("emu_km_set_centroids", [ll, vp, i32]), ("emu_km_partial", [ll, ll, vp, vp, vp, ctypes.POINTER(ctypes.c_double)]),
                       ("emu_km_finish", [ll, ll, i32, vp, vp, vp, ctypes.POINTER(i32)])):
        getattr(lib, name).restype = ctypes.c_char_p
        getattr(lib, name).argtypes = args
    lib.emu_km_assign.restype = ll
    lib.emu_km_assign.argtypes = [vp, ll, vp, vp]
    lib.emu_km_sort.restype = ctypes.c_char_p
    lib.emu_km_sort.argtypes = [vp, vp, i32, i32, vp]
    return lib


def build_fp32_scan_emu(workdir):
    """The exact fp32 tier: pq_ffma_scan_kernel + launcher, search_fp32_scan, pq_merge_lists_kernel; TMA and mbarriers replaced
    by stand-ins (ffma_emu.cpp.in)."""
    common = open(os.path.join(CSRC, "pq_common.cuh")).read()
    sel = open(os.path.join(CSRC, "pq_select.cu")).read()
    ffma = open(os.path.join(CSRC, "pq_ffma.cu")).read()
    index = open(os.path.join(CSRC, "pq_index.cu")).read()
    helpers = "\n".join(key_and_sort_helpers(common))
    select = "\n".join([
        extract(sel, "constexpr int kSelThreads = 256;", upto="constexpr int kSelThreads = 256;"),
        extract(sel, "struct MergeParams {"),
        extract(sel, "void emit_result(const MergeLaunch& a, int q, int i, uint64_t key)"),
        extract(sel, "pq_merge_lists_kernel(const MergeParams p)"),
        extract(sel, "static int next_pow2(int v)").replace("next_pow2", "sel_next_pow2"),
        extract(sel, "cudaError_t merge_lists_launch(const MergeLaunch& a, cudaStream_t stream)").replace("next_pow2", "sel_next_pow2"),
    ])
    a = ffma.index("namespace pq {") + len("namespace pq {")
    b = ffma.index("}  // namespace pq")
    body = ffma[a:b].replace("extern __shared__ __align__(1024) uint8_t smem[];", "uint8_t* smem = smem_raw;")
    tmpl = open(os.path.join(SIMT, "ffma_emu.cpp.in")).read()
    text = (tmpl.replace("@HELPERS@", to_host(helpers)).replace("@SELECT@", to_host(select)).replace("@FFMA@", to_host(body))
            .replace("@SCAN_DRIVER@", to_host(extract(index, "int search_fp32_scan(pq_index* ix"))))
    lib = compile_so(text, workdir, "ffma_emu", opt="-O2", extra=("-I", CSRC, "-I", "/usr/local/cuda/include", "-Wl,-Bsymbolic"))
    vp, ll, i32 = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    lib.emu_fp32_scan.restype = ctypes.c_char_p
    lib.emu_fp32_scan.argtypes = [vp, vp, ll, vp, vp, i32, i32, i32, i32, ll, vp, vp, vp]
    return lib


def run_fp32_scan_emu(lib, xb, xq, k, metric, n_sms=6, id_base=0, schedule=0):
    lib.emu_set_schedule(schedule)
    xb = np.ascontiguousarray(xb, np.float32)
    xq = np.ascontiguousarray(xq, np.float32)
    norms = np.zeros(len(xb) + 256, np.float32)
    norms[:len(xb)] = engine_norms(xb) if len(xb) else 0
    q_norm = engine_norms(xq)
    D = np.full((len(xq), k), np.nan, np.float32)
    I = np.full((len(xq), k), -7, np.int64)
    stats = np.zeros(10, np.int64)
    msg = lib.emu_fp32_scan(xb.ctypes.data, norms.ctypes.data, len(xb), xq.ctypes.data, q_norm.ctypes.data, len(xq), k, metric, n_sms, id_base,
                            D.ctypes.data, I.ctypes.data, stats.ctypes.data)
    assert msg is None, msg.decode()
    return D, I, stats
