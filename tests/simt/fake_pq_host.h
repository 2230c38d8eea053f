// Test infrastructure: the engine's host-side plumbing on host memory — error reporting and DevBuf on malloc (filled with
// garbage, so that anything a driver forgets to initialise shows).  Include after pq_host.h, simt_emu.h and fake_cuda.h.
#pragma once

namespace pq {
int set_error(int code, const char*, ...) { return code; }
int cuda_fail(cudaError_t e, const char*, int line) {
    static char msg[128];
    snprintf(msg, sizeof(msg), "CUDA call failed (error %d) at line %d of the generated source", (int)e, line);
    g_emu_error = msg;
    return PQ_ERR_CUDA;
}
int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return PQ_OK;
    free(p);
    p = malloc(bytes + 64);
    memset(p, 0xCD, bytes + 64);
    cap = bytes;
    return PQ_OK;
}
void DevBuf::release() { free(p); p = nullptr; cap = 0; }
}  // namespace pq
