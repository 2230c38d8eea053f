// Test infrastructure: a fake CUDA runtime for host drivers compiled with g++ (tests/simt/harness.py) — host memory
// instead of device memory, kernel launches through the SIMT emulator with the shared-memory rules checked:
// dynamic shared memory above 48 KB must have been allowed by cudaFuncSetAttribute, and a kernel must not write past it.
// Include AFTER the CUDA host headers and simt_emu.h.
#pragma once

static const char* g_emu_error = nullptr;
static std::map<const void*, int> g_max_dyn_smem;

extern "C" void emu_set_schedule(unsigned seed) { simt::g_schedule_seed = seed; }

extern "C" {
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, enum cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, enum cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void* f, enum cudaFuncAttribute a, int v) {
    if (a == cudaFuncAttributeMaxDynamicSharedMemorySize) g_max_dyn_smem[f] = v;
    return v <= 232448 ? cudaSuccess : cudaErrorInvalidValue;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
}


namespace pq {
alignas(16) uint8_t smem_raw[232448 + 4096];

// (cuda_runtime.h offers this typed overload to nvcc only)
template <class T>
static cudaError_t cudaFuncSetAttribute(T* entry, enum cudaFuncAttribute a, int v) { return ::cudaFuncSetAttribute((const void*)entry, a, v); }

template <class F>
static void emu_launch(const void* fn, long long grid, int block, size_t smem, F body) {
    if (g_emu_error) return;
    const auto it = g_max_dyn_smem.find(fn);
    const size_t allowed = std::max<size_t>(48 * 1024, it == g_max_dyn_smem.end() ? 0 : (size_t)it->second);
    if (smem > allowed) { g_emu_error = "launch asks for more dynamic shared memory than cudaFuncSetAttribute allowed"; return; }
    if (grid < 1 || block < 1 || block > 1024) { g_emu_error = "bad launch configuration"; return; }
    memset(smem_raw + smem, 0xA5, sizeof(smem_raw) - smem);
    const char* m = simt::launch((unsigned)grid, (unsigned)block, body);
    if (m) { g_emu_error = m; return; }
    for (size_t i = smem; i < sizeof(smem_raw); ++i)
        if (smem_raw[i] != 0xA5) { g_emu_error = "a kernel wrote past the dynamic shared memory it was given"; return; }
}
}  // namespace pq
#define EMU_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu_launch((const void*)(kernel), (grid), (block), (smem), [&] { (kernel)(__VA_ARGS__); })
