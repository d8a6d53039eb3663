// Alert generation on the device: the K ping-pong edge failure detectors of every virtual node
// (PingPongFailureDetector.java:38-121, one per entry of getSubjectsOf(myAddr), MembershipService.java:697-707) and the
// AlertMessages their notifiers raise (edgeFailureNotification, MembershipService.java:472-495), as cells ready for
// rapid_cd_apply_batch_dev.  One 32-bit word per detector; one tick = one failure-detector interval of the whole cluster:
// a single elementwise pass (HBM-bound: 8 B of state + 4 B of subject table per detector), a prefix sum, a scatter.
#include <limits.h>

#include <cub/cub.cuh>

#include "common.cuh"

namespace rapid {

// detector word: bits 0..22 failureCount, bit 23 notified, bits 24..31 bootstrapResponseCount (saturating; only "> 30" is read)
#define FD_CNT_MASK 0x7fffffu
#define FD_NOTIFIED (1u << 23)

struct FdScal {
    unsigned long long totals;          // (alerts << 32) | cells of the last tick
};

// One interval: run() of every detector (:75-85).  out[idx] = (1 << 32) | #cells if the notifier fired, else 0.
__global__ void k_fd_tick(int64_t n, int K, const int32_t* __restrict__ subj, const uint8_t* __restrict__ flags,
                          const uint8_t* __restrict__ edge_fail, int32_t thr, int32_t boot_thr, uint32_t* __restrict__ st,
                          unsigned long long* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * K) return;
    const int64_t o = idx / K;
    unsigned long long res = 0;
    const uint8_t fo = flags[o];
    if (!(fo & RAPID_FD_CRASHED)) {                                              // a crashed process runs nothing
        uint32_t w = st[idx];
        const int32_t s = subj[idx];
        if ((int32_t)(w & FD_CNT_MASK) >= thr && !(w & FD_NOTIFIED)) {           // hasFailed() && !notified (:76-79)
            w |= FD_NOTIFIED;
            int cells = 0;                                                       // getRingNumbers(myAddr, subject) (MembershipView.java:397-418)
            for (int r = 0; r < K; ++r) cells += subj[o * K + r] == s ? 1 : 0;
            res = (1ull << 32) | (unsigned long long)cells;
        } else {                                                                 // probe (:80-84) and its callback
            const uint8_t fs = flags[s];
            const bool fail = (edge_fail && edge_fail[idx]) || (fo & RAPID_FD_EGRESS_BLOCKED) || (fs & (RAPID_FD_CRASHED | RAPID_FD_INGRESS_BLOCKED));
            bool count = fail;
            if (!fail && (fs & RAPID_FD_BOOTSTRAPPING)) {                        // :97-104
                uint32_t b = w >> 24;
                if (b < 255) ++b;
                w = (w & 0x00ffffffu) | (b << 24);
                count = (int32_t)b > boot_thr;
            }
            if (count && (w & FD_CNT_MASK) < FD_CNT_MASK) w = (w & ~FD_CNT_MASK) | ((w & FD_CNT_MASK) + 1);   // :120-123
        }
        st[idx] = w;
    }
    out[idx] = res;
}

// alerts and cells of the detectors that fired, in detector order (node, then ring of the detector), rings ascending
__global__ void k_fd_emit(int64_t n, int K, const int32_t* __restrict__ subj, const unsigned long long* __restrict__ cnt,
                          const unsigned long long* __restrict__ pos, int64_t cfg, int32_t* __restrict__ a_obs,
                          int32_t* __restrict__ a_subj, uint16_t* __restrict__ a_mask, int32_t* __restrict__ c_src,
                          int32_t* __restrict__ c_dst, uint8_t* __restrict__ c_ring, uint8_t* __restrict__ c_status,
                          int64_t* __restrict__ c_cfg) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * K) return;
    if (cnt[idx] == 0) return;
    const int64_t o = idx / K;
    const int32_t s = subj[idx];
    const uint32_t a = (uint32_t)(pos[idx] >> 32);
    uint32_t c = (uint32_t)pos[idx];
    uint32_t mask = 0;
    for (int r = 0; r < K; ++r) {
        if (subj[o * K + r] != s) continue;
        mask |= 1u << r;
        c_src[c] = (int32_t)o; c_dst[c] = s; c_ring[c] = (uint8_t)r; c_status[c] = RAPID_EDGE_DOWN; c_cfg[c] = cfg;
        ++c;
    }
    a_obs[a] = (int32_t)o; a_subj[a] = s; a_mask[a] = (uint16_t)mask;
}

__global__ void k_fd_totals(int64_t D, const unsigned long long* __restrict__ cnt, const unsigned long long* __restrict__ pos,
                            FdScal* __restrict__ sc) {
    sc->totals = pos[D - 1] + cnt[D - 1];
}

struct FD {
    const View* view = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    int64_t n = 0;
    int K = 0;
    int32_t thr = 10, boot_thr = 30;
    uint64_t view_epoch = 0;
    DevBuf<uint32_t> st;
    DevBuf<unsigned long long> cnt, pos;
    DevBuf<uint8_t> flags, edge, cub_tmp;
    DevBuf<FdScal> sc;
    PinnedBuf<FdScal> h_sc;
    int64_t n_alerts = 0, n_cells = 0;
    DevBuf<int32_t> a_obs, a_subj, c_src, c_dst;
    DevBuf<uint16_t> a_mask;
    DevBuf<uint8_t> c_ring, c_status;
    DevBuf<int64_t> c_cfg;
};

static const int TB = 256;
static inline unsigned grid_for(int64_t n) { return (unsigned)ceil_div<int64_t>(n > 0 ? n : 1, TB); }

static int32_t fd_alloc(FD* fd) {
    const View* v = fd->view;
    fd->n = v->n; fd->K = v->K; fd->view_epoch = v->epoch;
    const size_t D = (size_t)std::max<int64_t>(fd->n * fd->K, 1);
    RAPID_CHECK(fd->st.reserve(D)); RAPID_CHECK(fd->cnt.reserve(D)); RAPID_CHECK(fd->pos.reserve(D));
    RAPID_CHECK(fd->flags.reserve((size_t)std::max<int64_t>(fd->n, 1))); RAPID_CHECK(fd->edge.reserve(D));
    // at most every detector fires in one tick; a fired detector yields at most K cells
    RAPID_CHECK(fd->a_obs.reserve(D)); RAPID_CHECK(fd->a_subj.reserve(D)); RAPID_CHECK(fd->a_mask.reserve(D));
    RAPID_CUDA(cudaMemsetAsync(fd->st.p, 0, D * sizeof(uint32_t), fd->stream));
    fd->n_alerts = fd->n_cells = 0;
    return RAPID_OK;
}

static int32_t fd_tick_device(FD* fd, const uint8_t* d_flags, const uint8_t* d_edge, int64_t cfg, int64_t* n_alerts, int64_t* n_cells) {
    cudaStream_t s = fd->stream;
    const int64_t D = fd->n * fd->K;
    fd->n_alerts = fd->n_cells = 0;
    if (fd->n >= 2 && D > 0) {                               // getSubjectsOf is empty in a one-node view (MembershipView.java:270-272)
        k_fd_tick<<<grid_for(D), TB, 0, s>>>(fd->n, fd->K, fd->view->subj.p, d_flags, d_edge, fd->thr, fd->boot_thr, fd->st.p, fd->cnt.p);
        RAPID_KERNEL_CHECK();
        size_t bytes = 0;
        RAPID_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, fd->cnt.p, fd->pos.p, (int)D, s));
        RAPID_CHECK(fd->cub_tmp.reserve(bytes));
        RAPID_CUDA(cub::DeviceScan::ExclusiveSum(fd->cub_tmp.p, bytes, fd->cnt.p, fd->pos.p, (int)D, s));
        // the cell buffers are sized exactly (a bound of K cells per detector would be 180 MB per million nodes): one 8-byte readback
        k_fd_totals<<<1, 1, 0, s>>>(D, fd->cnt.p, fd->pos.p, fd->sc.p);
        RAPID_KERNEL_CHECK();
        RAPID_CUDA(cudaMemcpyAsync(fd->h_sc.p, fd->sc.p, sizeof(FdScal), cudaMemcpyDeviceToHost, s));
        RAPID_CUDA(cudaStreamSynchronize(s));
        const unsigned long long tot = fd->h_sc.p->totals;
        fd->n_alerts = (int64_t)(tot >> 32); fd->n_cells = (int64_t)(tot & 0xffffffffull);
        if (fd->n_alerts > 0) {
            const size_t C = (size_t)std::max<int64_t>(fd->n_cells, 1);
            RAPID_CHECK(fd->c_src.reserve(C)); RAPID_CHECK(fd->c_dst.reserve(C)); RAPID_CHECK(fd->c_ring.reserve(C));
            RAPID_CHECK(fd->c_status.reserve(C)); RAPID_CHECK(fd->c_cfg.reserve(C));
            k_fd_emit<<<grid_for(D), TB, 0, s>>>(fd->n, fd->K, fd->view->subj.p, fd->cnt.p, fd->pos.p, cfg, fd->a_obs.p, fd->a_subj.p, fd->a_mask.p,
                                                fd->c_src.p, fd->c_dst.p, fd->c_ring.p, fd->c_status.p, fd->c_cfg.p);
            RAPID_KERNEL_CHECK();
        }
    }
    if (n_alerts) *n_alerts = fd->n_alerts;
    if (n_cells) *n_cells = fd->n_cells;
    return RAPID_OK;
}

}  // namespace rapid

using namespace rapid;

struct rapid_fdet : rapid::FD {};

extern "C" {

int32_t rapid_fdet_create(rapid_fdet** out, const rapid_view* v, int32_t failure_threshold, int32_t bootstrap_threshold) {
    if (!out || !v || failure_threshold < 1 || failure_threshold > 1000000 || bootstrap_threshold < 0 || bootstrap_threshold > 254) { set_error("bad arguments"); return RAPID_EINVAL; }
    *out = nullptr;
    DeviceGuard g(v->device);
    rapid_fdet* fd = new rapid_fdet();
    fd->view = v; fd->device = v->device; fd->thr = failure_threshold; fd->boot_thr = bootstrap_threshold;
    int32_t rc = RAPID_OK;
    do {
        if (cudaStreamCreateWithFlags(&fd->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&fd->ev0) != cudaSuccess ||
            cudaEventCreate(&fd->ev1) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "stream", __FILE__, __LINE__); break; }
        if ((rc = fd->sc.reserve(1)) || (rc = fd->h_sc.reserve(1))) break;
        if ((rc = fd_alloc(fd))) break;
        if (cudaStreamSynchronize(fd->stream) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "init", __FILE__, __LINE__); break; }
    } while (0);
    if (rc) { rapid_fdet_destroy(fd); return rc; }
    *out = fd;
    return RAPID_OK;
}

int32_t rapid_fdet_destroy(rapid_fdet* h) {
    rapid_fdet* fd = h;
    if (!fd) return RAPID_OK;
    DeviceGuard g(fd->device);
    if (fd->stream) { cudaStreamSynchronize(fd->stream); cudaStreamDestroy(fd->stream); }
    if (fd->ev0) cudaEventDestroy(fd->ev0);
    if (fd->ev1) cudaEventDestroy(fd->ev1);
    delete fd;
    return RAPID_OK;
}

int32_t rapid_fdet_reset(rapid_fdet* h) {
    rapid_fdet* fd = h;
    if (!fd) { set_error("NULL handle"); return RAPID_EINVAL; }
    DeviceGuard g(fd->device);
    return fd_alloc(fd);
}

int32_t rapid_fdet_tick(rapid_fdet* h, const uint8_t* node_flags, const uint8_t* edge_fail, int64_t cfg_id, int64_t* n_alerts, int64_t* n_cells) {
    rapid_fdet* fd = h;
    if (!fd || (fd->n && !node_flags)) { set_error("bad arguments"); return RAPID_EINVAL; }
    if (fd->view_epoch != fd->view->epoch || fd->n != fd->view->n) { set_error("the view changed: call rapid_fdet_reset (detectors are re-created per configuration)"); return RAPID_EINVAL; }
    DeviceGuard g(fd->device);
    cudaStream_t s = fd->stream;
    RAPID_CUDA(cudaEventRecord(fd->ev0, s));
    if (fd->n) RAPID_CUDA(cudaMemcpyAsync(fd->flags.p, node_flags, (size_t)fd->n, cudaMemcpyHostToDevice, s));
    if (edge_fail && fd->n) RAPID_CUDA(cudaMemcpyAsync(fd->edge.p, edge_fail, (size_t)(fd->n * fd->K), cudaMemcpyHostToDevice, s));
    const int32_t rc = fd_tick_device(fd, fd->flags.p, edge_fail ? fd->edge.p : nullptr, cfg_id, n_alerts, n_cells);
    if (rc == RAPID_OK) { cudaEventRecord(fd->ev1, s); cudaEventSynchronize(fd->ev1); cudaEventElapsedTime(&fd->last_ms, fd->ev0, fd->ev1); }
    return rc;
}

int32_t rapid_fdet_tick_dev(rapid_fdet* h, const uint8_t* node_flags_dev, const uint8_t* edge_fail_dev, int64_t cfg_id, int64_t* n_alerts,
                            int64_t* n_cells) {
    rapid_fdet* fd = h;
    if (!fd || (fd->n && !node_flags_dev)) { set_error("bad arguments"); return RAPID_EINVAL; }
    if (fd->view_epoch != fd->view->epoch || fd->n != fd->view->n) { set_error("the view changed: call rapid_fdet_reset"); return RAPID_EINVAL; }
    DeviceGuard g(fd->device);
    cudaStream_t s = fd->stream;
    RAPID_CUDA(cudaEventRecord(fd->ev0, s));
    const int32_t rc = fd_tick_device(fd, node_flags_dev, edge_fail_dev, cfg_id, n_alerts, n_cells);
    if (rc == RAPID_OK) { cudaEventRecord(fd->ev1, s); cudaEventSynchronize(fd->ev1); cudaEventElapsedTime(&fd->last_ms, fd->ev0, fd->ev1); }
    return rc;
}

int32_t rapid_fdet_cells_dev(const rapid_fdet* h, const int32_t** src, const int32_t** dst, const uint8_t** ring, const uint8_t** status,
                             const int64_t** cfg) {
    const rapid_fdet* fd = h;
    if (!fd) { set_error("NULL handle"); return RAPID_EINVAL; }
    if (src) *src = fd->c_src.p;
    if (dst) *dst = fd->c_dst.p;
    if (ring) *ring = fd->c_ring.p;
    if (status) *status = fd->c_status.p;
    if (cfg) *cfg = fd->c_cfg.p;
    return RAPID_OK;
}

int32_t rapid_fdet_read_cells(const rapid_fdet* h, int32_t* src, int32_t* dst, uint8_t* ring, uint8_t* status, int64_t* cfg) {
    const rapid_fdet* fd = h;
    if (!fd) { set_error("NULL handle"); return RAPID_EINVAL; }
    DeviceGuard g(fd->device);
    const size_t n = (size_t)fd->n_cells;
    if (n == 0) return RAPID_OK;
    cudaStream_t s = fd->stream;
    if (src) RAPID_CUDA(cudaMemcpyAsync(src, fd->c_src.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (dst) RAPID_CUDA(cudaMemcpyAsync(dst, fd->c_dst.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (ring) RAPID_CUDA(cudaMemcpyAsync(ring, fd->c_ring.p, n, cudaMemcpyDeviceToHost, s));
    if (status) RAPID_CUDA(cudaMemcpyAsync(status, fd->c_status.p, n, cudaMemcpyDeviceToHost, s));
    if (cfg) RAPID_CUDA(cudaMemcpyAsync(cfg, fd->c_cfg.p, n * 8, cudaMemcpyDeviceToHost, s));
    RAPID_CUDA(cudaStreamSynchronize(s));
    return RAPID_OK;
}

int32_t rapid_fdet_read_alerts(const rapid_fdet* h, int32_t* observer, int32_t* subject, uint16_t* ring_mask) {
    const rapid_fdet* fd = h;
    if (!fd) { set_error("NULL handle"); return RAPID_EINVAL; }
    DeviceGuard g(fd->device);
    const size_t n = (size_t)fd->n_alerts;
    if (n == 0) return RAPID_OK;
    cudaStream_t s = fd->stream;
    if (observer) RAPID_CUDA(cudaMemcpyAsync(observer, fd->a_obs.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (subject) RAPID_CUDA(cudaMemcpyAsync(subject, fd->a_subj.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (ring_mask) RAPID_CUDA(cudaMemcpyAsync(ring_mask, fd->a_mask.p, n * 2, cudaMemcpyDeviceToHost, s));
    RAPID_CUDA(cudaStreamSynchronize(s));
    return RAPID_OK;
}

int32_t rapid_fdet_state(const rapid_fdet* h, int64_t node, int32_t k, int32_t* failure_count, int32_t* notified) {
    const rapid_fdet* fd = h;
    if (!fd || node < 0 || node >= fd->n || k < 0 || k >= fd->K) { set_error("bad arguments"); return RAPID_EINVAL; }
    DeviceGuard g(fd->device);
    uint32_t w = 0;
    RAPID_CUDA(cudaStreamSynchronize(fd->stream));
    RAPID_CUDA(cudaMemcpy(&w, fd->st.p + node * fd->K + k, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (failure_count) *failure_count = (int32_t)(w & FD_CNT_MASK);
    if (notified) *notified = (w & FD_NOTIFIED) ? 1 : 0;
    return RAPID_OK;
}

int32_t rapid_fdet_last_device_ms(const rapid_fdet* h, float* total_ms) {
    const rapid_fdet* fd = h;
    if (!fd || !total_ms) return RAPID_EINVAL;
    *total_ms = fd->last_ms;
    return RAPID_OK;
}

}  // extern "C"
