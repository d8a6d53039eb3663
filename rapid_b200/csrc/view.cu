// MembershipView on the device: K rings as structure-of-arrays in HBM.
//
// Follows rapid/src/main/java/com/vrg/rapid/MembershipView.java:
//   ring k  = members sorted by signed key  xx_k(hostname)*31 + xx_k.hashInt(port)        (:562-587)
//   observers = ring successors (:234-257), subjects / expected observers = predecessors (:308-322)
//   configuration id = 37-ary polynomial hash over identifiersSeen then ring-0 order       (:544-556)
// Not a port: the Java keeps K red-black trees of Endpoint objects with a memoised comparator; here every
// (ring, node) key is hashed in one kernel, each ring is one radix sort, and the observer/subject relations
// become two dense int32 tables that the cut-detection kernels index directly.
#include <cub/cub.cuh>

#include <algorithm>
#include <string>
#include <unordered_map>

#include "common.cuh"

namespace rapid {

// ------------------------------------------------------------------ kernels
__global__ void k_ring_keys(const uint8_t* __restrict__ hb, const int32_t* __restrict__ off,
                            const int32_t* __restrict__ port, int64_t first, int64_t count, int K,
                            int64_t* __restrict__ key, size_t stride) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * K) return;
    const int k = (int)(t / count);
    const int64_t i = first + t % count;
    const int32_t o = off[i];
    key[(size_t)k * stride + i] = ring_key(hb + o, off[i + 1] - o, port[i], k);
}

// sortable unsigned image of a signed key
__global__ void k_flip_keys(const int64_t* __restrict__ key, size_t stride, int k, int64_t n,
                            uint64_t* __restrict__ ukey, int32_t* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ukey[i] = (uint64_t)key[(size_t)k * stride + i] ^ 0x8000000000000000ULL;
    ids[i] = (int32_t)i;
}

__global__ void k_unflip_and_check(const uint64_t* __restrict__ ukey_sorted, int64_t n, int64_t* __restrict__ sorted_key,
                                   int32_t* __restrict__ collision /* [2]: position, ring */, int k) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sorted_key[i] = (int64_t)(ukey_sorted[i] ^ 0x8000000000000000ULL);
    if (i + 1 < n && ukey_sorted[i] == ukey_sorted[i + 1]) {
        if (atomicCAS(&collision[0], -1, (int32_t)i) == -1) collision[1] = k;
    }
}

__global__ void k_tables(const int32_t* __restrict__ ring, int64_t n, int K, int32_t* __restrict__ obs,
                         int32_t* __restrict__ subj, int32_t* __restrict__ pos0) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * K) return;
    const int k = (int)(t / n);
    const int64_t p = t % n;
    const int32_t* r = ring + (size_t)k * n;
    const int32_t node = r[p];
    if (n <= 1) {
        obs[(size_t)node * K + k] = -1;
        subj[(size_t)node * K + k] = -1;
    } else {
        obs[(size_t)node * K + k] = r[p + 1 == n ? 0 : p + 1];        // TreeSet.higher, wrap to first()
        subj[(size_t)node * K + k] = r[p == 0 ? n - 1 : p - 1];       // TreeSet.lower, wrap to last()
    }
    if (k == 0) pos0[node] = (int32_t)p;
}

// Joiners: expected observers = predecessor of the joiner's key on every ring (getExpectedObserversOf :292-303).
// flag[j] = 1 if the identical endpoint is already a member.
__global__ void k_joiner_rows(const uint8_t* __restrict__ hb, const int32_t* __restrict__ off,
                              const int32_t* __restrict__ port, const int64_t* __restrict__ key, size_t stride,
                              const int64_t* __restrict__ sorted_key, const int32_t* __restrict__ ring, int64_t n,
                              int K, int64_t first, int64_t count, int32_t* __restrict__ obs, int32_t* __restrict__ flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * K) return;
    const int k = (int)(t / count);
    const int64_t id = first + t % count;
    const int64_t kj = key[(size_t)k * stride + id];
    const int64_t* sk = sorted_key + (size_t)k * n;
    // lower_bound: first position with sk[pos] >= kj
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sk[mid] < kj) lo = mid + 1; else hi = mid;
    }
    int32_t pred = -1;
    if (n > 0) pred = ring[(size_t)k * n + (lo == 0 ? n - 1 : lo - 1)];
    obs[(size_t)id * K + k] = pred;
    if (k == 0 && lo < n && sk[lo] == kj) {
        const int32_t m = ring[lo];
        const int32_t oa = off[id], ob = off[m];
        const int32_t la = off[id + 1] - oa, lb = off[m + 1] - ob;
        bool same = (la == lb) && (port[id] == port[m]);
        for (int32_t c = 0; same && c < la; ++c) same = hb[oa + c] == hb[ob + c];
        if (same) flag[t % count] = 1;
    }
}

// Configuration id: hash = 1; for x in X: hash = hash*37 + x  (wrapping).  X = hashLong(high),hashLong(low) of the
// sorted identifiers, then xx0(hostname), xx0.hashInt(port) in ring-0 order.
__device__ __forceinline__ uint64_t pow37(uint64_t e) {
    uint64_t r = 1, b = 37;
    while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
    return r;
}

__global__ void k_config_id(const int64_t* __restrict__ id_high, const int64_t* __restrict__ id_low,
                            const int32_t* __restrict__ id_order, int64_t n_ids, const uint8_t* __restrict__ hb,
                            const int32_t* __restrict__ off, const int32_t* __restrict__ port,
                            const int32_t* __restrict__ ring0, int64_t n, unsigned long long* __restrict__ out) {
    const int64_t M = 2 * n_ids + 2 * n;
    const int SEG = 32;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = t * SEG;
    uint64_t h = 0;
    if (s < M) {
        const int64_t e = (s + SEG < M) ? s + SEG : M;
        for (int64_t i = s; i < e; ++i) {
            uint64_t x;
            if (i < 2 * n_ids) {
                const int32_t j = id_order[i >> 1];
                x = xxh64_long((i & 1) ? id_low[j] : id_high[j], 0);
            } else {
                const int64_t q = i - 2 * n_ids;
                const int32_t node = ring0[q >> 1];
                if (q & 1) x = xxh64_int(port[node], 0);
                else { const int32_t o = off[node]; x = xxh64_bytes(hb + o, off[node + 1] - o, 0); }
            }
            h = h * 37 + x;
        }
        h *= pow37((uint64_t)(M - e));
        if (t == 0) h += pow37((uint64_t)M);     // the leading "hash = 1"
    } else if (t == 0) {
        h = 1;                                    // M == 0
    }
    // block reduce (wrapping add)
    __shared__ unsigned long long sm[32];
    unsigned long long v = h;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (blockDim.x + 31) / 32 ? sm[threadIdx.x] : 0ULL;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0 && v) atomicAdd(out, v);
    }
}

__global__ void k_id_sort_keys(const int64_t* __restrict__ v, int64_t n, uint64_t* __restrict__ out, int32_t* idx,
                               const int32_t* __restrict__ order_in) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t j = order_in ? order_in[i] : (int32_t)i;
    out[i] = (uint64_t)v[j] ^ 0x8000000000000000ULL;
    idx[i] = j;
}

// ------------------------------------------------------------------ host helpers
static int32_t sort_pairs(DevBuf<uint8_t>& tmp, const uint64_t* kin, uint64_t* kout, const int32_t* vin, int32_t* vout,
                          int64_t n, cudaStream_t s) {
    size_t bytes = 0;
    RAPID_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, 64, s));
    RAPID_CHECK(tmp.reserve(bytes ? bytes : 1));
    RAPID_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kin, kout, vin, vout, (int)n, 0, 64, s));
    return RAPID_OK;
}

static int32_t ensure_total_capacity(View* v, int64_t ntot) {
    if ((size_t)ntot <= v->key_stride) return RAPID_OK;
    size_t ns = v->key_stride ? v->key_stride : 1;
    while (ns < (size_t)ntot) ns *= 2;
    DevBuf<int64_t> nk;
    RAPID_CHECK(nk.reserve(ns * (size_t)v->K));
    if (v->key.p && v->key_stride) {
        RAPID_CUDA(cudaMemcpy2DAsync(nk.p, ns * sizeof(int64_t), v->key.p, v->key_stride * sizeof(int64_t),
                                     (size_t)(v->n + v->nj) * sizeof(int64_t), (size_t)v->K, cudaMemcpyDeviceToDevice,
                                     v->stream));
        RAPID_CUDA(cudaStreamSynchronize(v->stream));
    }
    std::swap(v->key.p, nk.p);
    std::swap(v->key.cap, nk.cap);
    v->key_stride = ns;
    RAPID_CHECK(v->obs.reserve(ns * (size_t)v->K, true, v->stream));
    return RAPID_OK;
}

static int32_t upload_endpoints(View* v, int64_t count, const uint8_t* hb, const int32_t* off, const int32_t* port) {
    // append to host mirrors, then to device
    const int64_t base = v->n + v->nj;      // v->n must already be set for members (0 during create)
    const size_t old_bytes = v->h_host_bytes.size();
    const size_t add_bytes = count ? (size_t)(off[count] - off[0]) : 0;
    if (v->h_host_off.empty()) v->h_host_off.push_back(0);
    v->h_host_bytes.insert(v->h_host_bytes.end(), hb + (count ? off[0] : 0), hb + (count ? off[0] : 0) + add_bytes);
    for (int64_t i = 0; i < count; ++i) {
        const int32_t len = off[i + 1] - off[i];
        if (len < 0) { set_error("host_off must be non-decreasing"); return RAPID_EINVAL; }
        v->h_host_off.push_back(v->h_host_off.back() + len);
        v->h_port.push_back(port[i]);
    }
    if (v->h_host_bytes.size() > 0x7fffffffULL) { set_error("hostname bytes exceed 2 GiB"); return RAPID_EINVAL; }
    RAPID_CHECK(v->host_bytes.reserve(std::max<size_t>(1, v->h_host_bytes.size()), true, v->stream));
    RAPID_CHECK(v->host_off.reserve((size_t)(base + count + 1), true, v->stream));
    RAPID_CHECK(v->port.reserve(std::max<size_t>(1, (size_t)(base + count)), true, v->stream));
    if (add_bytes)
        RAPID_CUDA(cudaMemcpyAsync(v->host_bytes.p + old_bytes, v->h_host_bytes.data() + old_bytes, add_bytes,
                                   cudaMemcpyHostToDevice, v->stream));
    RAPID_CUDA(cudaMemcpyAsync(v->host_off.p + base, v->h_host_off.data() + base, (size_t)(count + 1) * sizeof(int32_t),
                               cudaMemcpyHostToDevice, v->stream));
    if (count)
        RAPID_CUDA(cudaMemcpyAsync(v->port.p + base, v->h_port.data() + base, (size_t)count * sizeof(int32_t),
                                   cudaMemcpyHostToDevice, v->stream));
    RAPID_CUDA(cudaStreamSynchronize(v->stream));
    return RAPID_OK;
}

static bool same_endpoint(const View* v, int64_t a, int64_t b) {
    const int32_t la = v->h_host_off[a + 1] - v->h_host_off[a], lb = v->h_host_off[b + 1] - v->h_host_off[b];
    return la == lb && v->h_port[a] == v->h_port[b] &&
           memcmp(v->h_host_bytes.data() + v->h_host_off[a], v->h_host_bytes.data() + v->h_host_off[b], (size_t)la) == 0;
}

static int32_t build_rings(View* v) {
    const int64_t n = v->n;
    const int K = v->K;
    cudaStream_t s = v->stream;
    RAPID_CHECK(ensure_total_capacity(v, std::max<int64_t>(n, 1)));
    RAPID_CHECK(v->sorted_key.reserve(std::max<size_t>(1, (size_t)n * K)));
    RAPID_CHECK(v->ring.reserve(std::max<size_t>(1, (size_t)n * K)));
    RAPID_CHECK(v->pos0.reserve(std::max<size_t>(1, (size_t)n)));
    RAPID_CHECK(v->subj.reserve(std::max<size_t>(1, (size_t)n * K)));
    if (n == 0) return RAPID_OK;
    const int TB = 256;
    k_ring_keys<<<(unsigned)ceil_div<int64_t>(n * K, TB), TB, 0, s>>>(v->host_bytes.p, v->host_off.p, v->port.p, 0, n, K,
                                                                      v->key.p, v->key_stride);
    RAPID_KERNEL_CHECK();
    DevBuf<uint64_t> uk_in, uk_out;
    DevBuf<int32_t> id_in, collision;
    DevBuf<uint8_t> tmp;
    RAPID_CHECK(uk_in.reserve((size_t)n));
    RAPID_CHECK(uk_out.reserve((size_t)n));
    RAPID_CHECK(id_in.reserve((size_t)n));
    RAPID_CHECK(collision.reserve(2));
    RAPID_CUDA(cudaMemsetAsync(collision.p, 0xff, 2 * sizeof(int32_t), s));
    for (int k = 0; k < K; ++k) {
        k_flip_keys<<<(unsigned)ceil_div<int64_t>(n, TB), TB, 0, s>>>(v->key.p, v->key_stride, k, n, uk_in.p, id_in.p);
        RAPID_KERNEL_CHECK();
        RAPID_CHECK(sort_pairs(tmp, uk_in.p, uk_out.p, id_in.p, v->ring.p + (size_t)k * n, n, s));
        k_unflip_and_check<<<(unsigned)ceil_div<int64_t>(n, TB), TB, 0, s>>>(uk_out.p, n, v->sorted_key.p + (size_t)k * n,
                                                                             collision.p, k);
        RAPID_KERNEL_CHECK();
    }
    int32_t coll[2];
    RAPID_CUDA(cudaMemcpyAsync(coll, collision.p, sizeof(coll), cudaMemcpyDeviceToHost, s));
    RAPID_CUDA(cudaStreamSynchronize(s));
    if (coll[0] >= 0) {
        int32_t ids[2];
        RAPID_CUDA(cudaMemcpy(ids, v->ring.p + (size_t)coll[1] * n + coll[0], sizeof(ids), cudaMemcpyDeviceToHost));
        if (same_endpoint(v, ids[0], ids[1])) {
            set_error("endpoint given twice (node ids %d and %d): NodeAlreadyInRingException", ids[0], ids[1]);
            return RAPID_EALREADY_IN_RING;
        }
        set_error("ring-%d key collision between node ids %d and %d (TreeSet would silently drop one)", coll[1], ids[0], ids[1]);
        return RAPID_EHASH_COLLISION;
    }
    k_tables<<<(unsigned)ceil_div<int64_t>(n * K, TB), TB, 0, s>>>(v->ring.p, n, K, v->obs.p, v->subj.p, v->pos0.p);
    RAPID_KERNEL_CHECK();
    RAPID_CUDA(cudaStreamSynchronize(s));
    return RAPID_OK;
}

}  // namespace rapid

using namespace rapid;

extern "C" {

int32_t rapid_view_create(rapid_view** out, int32_t K, int64_t n, const uint8_t* host_bytes, const int32_t* host_off,
                          const int32_t* port, int32_t device) {
    if (!out) { set_error("out is NULL"); return RAPID_EINVAL; }
    *out = nullptr;
    if (K < 1 || K > RAPID_MAX_K) { set_error("K must be in [1, %d], got %d", RAPID_MAX_K, K); return RAPID_EINVAL; }
    if (n < 0 || n > 0x7ffffff0LL) { set_error("bad n"); return RAPID_EINVAL; }
    if (n > 0 && (!host_bytes || !host_off || !port)) { set_error("NULL endpoint arrays"); return RAPID_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: librapid_b200 has no CPU fallback");
        return RAPID_ECUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return RAPID_EINVAL; }
    DeviceGuard g(device);
    rapid_view* v = new rapid_view();
    v->device = device;
    v->K = K;
    int32_t rc = RAPID_OK;
    do {
        if (cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "stream", __FILE__, __LINE__); break; }
        static const int32_t zero_off[1] = {0};
        rc = upload_endpoints(v, n, host_bytes, n ? host_off : zero_off, port);
        if (rc) break;
        v->n = n;
        rc = build_rings(v);
    } while (0);
    if (rc) { rapid_view_destroy(v); return rc; }
    *out = v;
    return RAPID_OK;
}

int32_t rapid_view_destroy(rapid_view* v) {
    if (!v) return RAPID_OK;
    DeviceGuard g(v->device);
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
    return RAPID_OK;
}

int32_t rapid_view_size(const rapid_view* v, int64_t* out_n) {
    if (!v || !out_n) { set_error("NULL argument"); return RAPID_EINVAL; }
    *out_n = v->n;
    return RAPID_OK;
}

int32_t rapid_view_num_joiners(const rapid_view* v, int64_t* out) {
    if (!v || !out) { set_error("NULL argument"); return RAPID_EINVAL; }
    *out = v->nj;
    return RAPID_OK;
}

int32_t rapid_view_ring(const rapid_view* v, int32_t k, int32_t* out_ids) {
    if (!v || k < 0 || k >= v->K || (!out_ids && v->n)) { set_error("bad ring index %d", k); return RAPID_EINVAL; }
    DeviceGuard g(v->device);
    if (v->n) RAPID_CUDA(cudaMemcpy(out_ids, v->ring.p + (size_t)k * v->n, (size_t)v->n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return RAPID_OK;
}

int32_t rapid_view_keys(const rapid_view* v, int32_t k, int64_t* out_keys) {
    if (!v || k < 0 || k >= v->K) { set_error("bad ring index %d", k); return RAPID_EINVAL; }
    DeviceGuard g(v->device);
    const int64_t tot = v->n + v->nj;
    if (tot) RAPID_CUDA(cudaMemcpy(out_keys, v->key.p + (size_t)k * v->key_stride, (size_t)tot * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return RAPID_OK;
}

static int32_t member_row(const rapid_view* v, const int32_t* table, int32_t node, int32_t* out, int32_t* out_count) {
    if (!v || !out || !out_count) { set_error("NULL argument"); return RAPID_EINVAL; }
    if (node < 0 || node >= v->n) { set_error("node %d is not in the ring", node); return RAPID_ENOT_IN_RING; }
    if (v->n <= 1) { *out_count = 0; return RAPID_OK; }
    DeviceGuard g(v->device);
    RAPID_CUDA(cudaMemcpy(out, table + (size_t)node * v->K, (size_t)v->K * sizeof(int32_t), cudaMemcpyDeviceToHost));
    *out_count = v->K;
    return RAPID_OK;
}

int32_t rapid_view_observers(const rapid_view* v, int32_t node, int32_t* out, int32_t* out_count) {
    return member_row(v, v ? v->obs.p : nullptr, node, out, out_count);
}

int32_t rapid_view_subjects(const rapid_view* v, int32_t node, int32_t* out, int32_t* out_count) {
    return member_row(v, v ? v->subj.p : nullptr, node, out, out_count);
}

int32_t rapid_view_tables(const rapid_view* v, int32_t* out_obs, int32_t* out_subj) {
    if (!v) { set_error("NULL view"); return RAPID_EINVAL; }
    DeviceGuard g(v->device);
    const size_t bytes = (size_t)v->n * v->K * sizeof(int32_t);
    if (bytes && out_obs) RAPID_CUDA(cudaMemcpy(out_obs, v->obs.p, bytes, cudaMemcpyDeviceToHost));
    if (bytes && out_subj) RAPID_CUDA(cudaMemcpy(out_subj, v->subj.p, bytes, cudaMemcpyDeviceToHost));
    return RAPID_OK;
}

int32_t rapid_view_joiner_tables(const rapid_view* v, int32_t* out) {
    if (!v || (!out && v->nj)) { set_error("NULL argument"); return RAPID_EINVAL; }
    DeviceGuard g(v->device);
    if (v->nj) RAPID_CUDA(cudaMemcpy(out, v->obs.p + (size_t)v->n * v->K, (size_t)v->nj * v->K * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return RAPID_OK;
}

int32_t rapid_view_ring_numbers(const rapid_view* v, int32_t observer, int32_t subject, uint16_t* out_mask) {
    if (!v || !out_mask) { set_error("NULL argument"); return RAPID_EINVAL; }
    int32_t row[RAPID_MAX_K], cnt = 0;
    RAPID_CHECK(rapid_view_subjects(v, observer, row, &cnt));
    uint16_t m = 0;
    for (int k = 0; k < cnt; ++k) if (row[k] == subject) m |= (uint16_t)(1u << k);
    *out_mask = m;
    return RAPID_OK;
}

// Computes joiner rows for endpoints [first, first+count) already uploaded; flags identical members.
static int32_t joiner_rows(rapid_view* v, int64_t first, int64_t count, std::vector<int32_t>& flags) {
    cudaStream_t s = v->stream;
    const int TB = 256;
    k_ring_keys<<<(unsigned)ceil_div<int64_t>(count * v->K, TB), TB, 0, s>>>(v->host_bytes.p, v->host_off.p, v->port.p, first,
                                                                             count, v->K, v->key.p, v->key_stride);
    RAPID_KERNEL_CHECK();
    DevBuf<int32_t> flag;
    RAPID_CHECK(flag.reserve((size_t)count));
    RAPID_CUDA(cudaMemsetAsync(flag.p, 0, (size_t)count * sizeof(int32_t), s));
    k_joiner_rows<<<(unsigned)ceil_div<int64_t>(count * v->K, TB), TB, 0, s>>>(
        v->host_bytes.p, v->host_off.p, v->port.p, v->key.p, v->key_stride, v->sorted_key.p, v->ring.p, v->n, v->K, first,
        count, v->obs.p, flag.p);
    RAPID_KERNEL_CHECK();
    flags.resize((size_t)count);
    RAPID_CUDA(cudaMemcpyAsync(flags.data(), flag.p, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    RAPID_CUDA(cudaStreamSynchronize(s));
    return RAPID_OK;
}

static void pop_endpoints(rapid_view* v, int64_t count) {
    for (int64_t i = 0; i < count; ++i) { v->h_host_off.pop_back(); v->h_port.pop_back(); }
    v->h_host_bytes.resize((size_t)v->h_host_off.back());
}

int32_t rapid_view_register_joiners(rapid_view* v, int64_t n_add, const uint8_t* host_bytes, const int32_t* host_off,
                                    const int32_t* port, int32_t* out_first_id) {
    if (!v || n_add < 0 || (n_add && (!host_bytes || !host_off || !port))) { set_error("bad arguments"); return RAPID_EINVAL; }
    DeviceGuard g(v->device);
    const int64_t first = v->n + v->nj;
    if (out_first_id) *out_first_id = (int32_t)first;
    if (n_add == 0) return RAPID_OK;
    if (first + n_add > 0x7ffffff0LL) { set_error("too many endpoints"); return RAPID_EINVAL; }
    RAPID_CHECK(ensure_total_capacity(v, first + n_add));
    RAPID_CHECK(upload_endpoints(v, n_add, host_bytes, host_off, port));
    std::vector<int32_t> flags;
    int32_t rc = joiner_rows(v, first, n_add, flags);
    if (rc == RAPID_OK) {
        for (int64_t j = 0; j < n_add; ++j)
            if (flags[(size_t)j]) { set_error("joiner %lld is already a member", (long long)j); rc = RAPID_EALREADY_IN_RING; break; }
    }
    if (rc != RAPID_OK) { pop_endpoints(v, n_add); return rc; }
    v->nj += n_add;
    ++v->epoch;
    return RAPID_OK;
}

// decideViewChange (MembershipService.java:385-444): every node of the decided cut that is a member leaves (ringDelete,
// MembershipView.java:167-201), every other one — a registered joiner — is added (ringAdd, :123-160).  The K rings are
// rebuilt on the device from the surviving endpoints (hash + radix sort + tables); ids are renumbered densely: surviving
// members keep their relative order, the admitted joiners follow in id order, joiners not in the cut are dropped.
int32_t rapid_view_apply_cut(rapid_view* v, const int32_t* cut_ids, int64_t n_cut, int32_t* out_old_to_new) {
    if (!v || n_cut < 0 || (n_cut && !cut_ids)) { set_error("bad arguments"); return RAPID_EINVAL; }
    DeviceGuard g(v->device);
    const int64_t tot = v->n + v->nj;
    std::vector<uint8_t> in_cut((size_t)tot, 0);
    for (int64_t i = 0; i < n_cut; ++i) {
        const int32_t id = cut_ids[i];
        if (id < 0 || id >= tot) { set_error("cut id %d outside [0, members + joiners)", id); return RAPID_EINVAL; }
        if (in_cut[(size_t)id]) {
            set_error("cut names node %d twice", id);
            return id < v->n ? RAPID_ENOT_IN_RING : RAPID_EALREADY_IN_RING;      // second ringDelete / ringAdd would throw
        }
        in_cut[(size_t)id] = 1;
    }
    std::vector<uint8_t> hb;
    std::vector<int32_t> off(1, 0), port;
    std::vector<int32_t> map((size_t)tot, -1);
    hb.reserve(v->h_host_bytes.size());
    int32_t next = 0;
    for (int64_t id = 0; id < tot; ++id) {
        const bool keep = id < v->n ? !in_cut[(size_t)id] : in_cut[(size_t)id];
        if (!keep) continue;
        const int32_t o = v->h_host_off[(size_t)id], len = v->h_host_off[(size_t)id + 1] - o;
        hb.insert(hb.end(), v->h_host_bytes.begin() + o, v->h_host_bytes.begin() + o + len);
        off.push_back(off.back() + len);
        port.push_back(v->h_port[(size_t)id]);
        map[(size_t)id] = next++;
    }
    // swap in the new endpoint list and rebuild
    v->h_host_bytes.clear(); v->h_host_off.clear(); v->h_port.clear();
    v->n = 0; v->nj = 0;
    static const uint8_t dummy = 0;
    static const int32_t zero_off[1] = {0};
    RAPID_CHECK(upload_endpoints(v, next, next ? hb.data() : &dummy, next ? off.data() : zero_off, port.data()));
    v->n = next;
    ++v->epoch;
    RAPID_CHECK(build_rings(v));
    if (out_old_to_new) memcpy(out_old_to_new, map.data(), (size_t)tot * sizeof(int32_t));
    return RAPID_OK;
}

int32_t rapid_view_expected_observers(const rapid_view* cv, const uint8_t* host, int32_t len, int32_t port, int32_t* out,
                                      int32_t* out_count) {
    rapid_view* v = const_cast<rapid_view*>(cv);   // uses scratch space past the registered endpoints; logically const
    if (!v || !out || !out_count || len < 0 || (len && !host)) { set_error("bad arguments"); return RAPID_EINVAL; }
    if (v->n == 0) { *out_count = 0; return RAPID_OK; }
    DeviceGuard g(v->device);
    const int64_t first = v->n + v->nj;
    RAPID_CHECK(ensure_total_capacity(v, first + 1));
    const int32_t off[2] = {0, len};
    static const uint8_t dummy = 0;
    RAPID_CHECK(upload_endpoints(v, 1, len ? host : &dummy, off, &port));
    std::vector<int32_t> flags;
    int32_t rc = joiner_rows(v, first, 1, flags);
    if (rc == RAPID_OK) {
        cudaError_t e = cudaMemcpy(out, v->obs.p + (size_t)first * v->K, (size_t)v->K * sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "copy", __FILE__, __LINE__);
    }
    pop_endpoints(v, 1);
    if (rc == RAPID_OK) *out_count = v->K;
    return rc;
}

int32_t rapid_view_config_id(const rapid_view* v, const int64_t* id_high, const int64_t* id_low, int64_t n_ids, int64_t* out) {
    if (!v || !out || n_ids < 0 || (n_ids && (!id_high || !id_low))) { set_error("bad arguments"); return RAPID_EINVAL; }
    DeviceGuard g(v->device);
    cudaStream_t s = v->stream;
    const int TB = 256;
    DevBuf<int64_t> dh, dl;
    DevBuf<uint64_t> k_in, k_out;
    DevBuf<int32_t> i_in, i_mid, order;
    DevBuf<uint8_t> tmp;
    DevBuf<unsigned long long> acc;
    const size_t m = std::max<int64_t>(1, n_ids);
    RAPID_CHECK(dh.reserve(m)); RAPID_CHECK(dl.reserve(m));
    RAPID_CHECK(k_in.reserve(m)); RAPID_CHECK(k_out.reserve(m));
    RAPID_CHECK(i_in.reserve(m)); RAPID_CHECK(i_mid.reserve(m)); RAPID_CHECK(order.reserve(m));
    RAPID_CHECK(acc.reserve(1));
    if (n_ids) {
        RAPID_CUDA(cudaMemcpyAsync(dh.p, id_high, (size_t)n_ids * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        RAPID_CUDA(cudaMemcpyAsync(dl.p, id_low, (size_t)n_ids * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        // NodeIdComparator (:474-500): signed (high, low).  LSD: stable sort by low, then by high.
        const unsigned gb = (unsigned)ceil_div<int64_t>(n_ids, TB);
        k_id_sort_keys<<<gb, TB, 0, s>>>(dl.p, n_ids, k_in.p, i_in.p, nullptr);
        RAPID_KERNEL_CHECK();
        RAPID_CHECK(sort_pairs(tmp, k_in.p, k_out.p, i_in.p, i_mid.p, n_ids, s));
        k_id_sort_keys<<<gb, TB, 0, s>>>(dh.p, n_ids, k_in.p, i_in.p, i_mid.p);
        RAPID_KERNEL_CHECK();
        RAPID_CHECK(sort_pairs(tmp, k_in.p, k_out.p, i_in.p, order.p, n_ids, s));
    }
    RAPID_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(unsigned long long), s));
    const int64_t M = 2 * n_ids + 2 * v->n;
    const int64_t threads = std::max<int64_t>(1, ceil_div<int64_t>(M, 32));
    k_config_id<<<(unsigned)ceil_div<int64_t>(threads, TB), TB, 0, s>>>(dh.p, dl.p, order.p, n_ids, v->host_bytes.p, v->host_off.p,
                                                                        v->port.p, v->ring.p, v->n, acc.p);
    RAPID_KERNEL_CHECK();
    unsigned long long h = 0;
    RAPID_CUDA(cudaMemcpyAsync(&h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    RAPID_CUDA(cudaStreamSynchronize(s));
    *out = (int64_t)h;
    return RAPID_OK;
}

}  // extern "C"
