// Error plumbing and small stateless entry points of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace rapid {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();   // clear the sticky-less error state
    return e == cudaErrorMemoryAllocation ? RAPID_ENOMEM : RAPID_ECUDA;
}

}  // namespace rapid

using namespace rapid;

extern "C" {

const char* rapid_version(void) { return "rapid_b200 0.1.0 (sm_100a)"; }

int32_t rapid_last_error(char* buf, size_t cap) {
    if (!buf || cap == 0) return RAPID_EINVAL;
    strncpy(buf, g_err, cap - 1);
    buf[cap - 1] = 0;
    return RAPID_OK;
}

int32_t rapid_device_count(int32_t* out) {
    if (!out) return RAPID_EINVAL;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); *out = 0; set_error("no CUDA device (%s)", cudaGetErrorString(e)); return RAPID_ECUDA; }
    *out = n;
    return RAPID_OK;
}

int32_t rapid_proposal_fingerprint(const int32_t* ids, int64_t n, uint64_t* h1, uint64_t* h2) {
    if (n < 0 || (n && !ids)) { set_error("bad arguments"); return RAPID_EINVAL; }
    uint64_t a = 0, b = 0;
    for (int64_t i = 0; i < n; ++i) { a += fp_mix1(ids[i]); b += fp_mix2(ids[i]); }
    if (h1) *h1 = a;
    if (h2) *h2 = b;
    return RAPID_OK;
}

int32_t rapid_fp_quorum(int64_t membership_size, int64_t* out) {
    if (!out || membership_size < 1) { set_error("bad arguments"); return RAPID_EINVAL; }
    *out = membership_size - (membership_size - 1) / 4;    // N - floor((N-1)/4.0), FastPaxos.java:145
    return RAPID_OK;
}

}  // extern "C"
