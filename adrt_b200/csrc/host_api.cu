// Host-pointer (NumPy path) entry points: adrt_b200_host_*.
//
// The reference's API is NumPy-in / NumPy-out (_wrappers.py:96-112).  These
// functions keep that contract: they take host buffers, move the batch through
// the GPU in chunks and return when the host output is complete.  Two slots
// (device in/out/workspace buffers + a stream each) are used round-robin so
// that the H2D copy of chunk i+1 overlaps the kernels and the D2H copy of
// chunk i (B200 has independent copy engines per direction).
//
// Pageable host memory is staged through pinned bounce buffers by a small pool
// of host threads (a single-threaded memcpy cannot feed PCIe Gen5); buffers
// that are already pinned (cudaHostAlloc / cudaHostRegister / torch
// pin_memory) are detected with cudaPointerGetAttributes and copied directly.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace adrt_b200 {
namespace {

constexpr int kSlots = 2;
constexpr int kMaxDevices = 16;

struct Slot {
    void *d_in = nullptr, *d_out = nullptr, *d_ws = nullptr;
    size_t cap_in = 0, cap_out = 0, cap_ws = 0;
    void *p_in = nullptr, *p_out = nullptr;  // pinned bounce buffers
    size_t pcap_in = 0, pcap_out = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    // pending D2H unstaging for this slot
    void *pending_dst = nullptr;
    size_t pending_bytes = 0;
};

struct Arena {
    std::mutex mu;
    bool init = false;
    Slot slot[kSlots];
};

Arena g_arena[kMaxDevices];

int grow(void **p, size_t *cap, size_t need, bool pinned)
{
    if (need <= *cap) return ADRT_B200_OK;
    if (*p) {
        if (pinned) cudaFreeHost(*p); else cudaFree(*p);
        *p = nullptr; *cap = 0;
    }
    // round up to 1 MiB so slightly different shapes reuse the allocation
    size_t sz = (need + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    cudaError_t e = pinned ? cudaHostAlloc(p, sz, cudaHostAllocDefault) : cudaMalloc(p, sz);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s of %zu bytes failed: %s", pinned ? "cudaHostAlloc" : "cudaMalloc", sz, cudaGetErrorString(e));
        *p = nullptr;
        return ADRT_B200_ENOMEM;
    }
    *cap = sz;
    return ADRT_B200_OK;
}

}  // namespace
bool is_pinned(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}
namespace {

int host_threads()
{
    static int n = [] {
        if (const char *e = getenv("ADRT_B200_HOST_THREADS")) return std::max(1, atoi(e));
        unsigned hc = std::thread::hardware_concurrency();
        return (int)std::min<unsigned>(std::max<unsigned>(hc, 1), 16);
    }();
    return n;
}

// memcpy split over several threads (pageable <-> pinned staging)
void parallel_memcpy(void *dst, const void *src, size_t bytes)
{
    const int nt = host_threads();
    if (bytes < (size_t(8) << 20) || nt == 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes / nt) + 4095) & ~size_t(4095);
    for (int i = 0; i < nt; ++i) {
        const size_t off = per * i;
        if (off >= bytes) break;
        const size_t len = std::min(per, bytes - off);
        th.emplace_back([=] { memcpy((char *)dst + off, (const char *)src + off, len); });
    }
    for (auto &t : th) t.join();
}

size_t chunk_budget_bytes()
{
    static size_t b = [] {
        if (const char *e = getenv("ADRT_B200_HOST_CHUNK_MB")) return (size_t)std::max(1, atoi(e)) << 20;
        // measured (64 x 2048^2 fp32 adrt + bdrt, pinned): 2048 MB 349 ms, 1024 MB 339, 512 MB 334, 300 MB 333 --
        // the first upload and the last download of a call are not overlapped, so short chunks win
        return size_t(512) << 20;
    }();
    return b;
}

using WsFn = std::function<size_t(int64_t)>;
using OpFn = std::function<int(const void *, void *, int64_t, void *, size_t, cudaStream_t)>;

int flush_pending(Slot &sl)
{
    if (sl.pending_dst) {
        ADRT_CUDA_CHECK(cudaEventSynchronize(sl.done));
        parallel_memcpy(sl.pending_dst, sl.p_out, sl.pending_bytes);
        sl.pending_dst = nullptr;
        sl.pending_bytes = 0;
    }
    return ADRT_B200_OK;
}

int run_host(const void *h_in, void *h_out, int64_t B, size_t in_item, size_t out_item, const WsFn &ws_fn,
             const OpFn &op, int device)
{
    ADRT_REQUIRE(h_in && h_out, "null pointer argument");
    ADRT_REQUIRE(device >= 0 && device < kMaxDevices, "bad device ordinal %d", device);
    int prev_dev = 0;
    ADRT_CUDA_CHECK(cudaGetDevice(&prev_dev));
    ADRT_CUDA_CHECK(cudaSetDevice(device));
    Arena &ar = g_arena[device];
    std::lock_guard<std::mutex> lock(ar.mu);
    int rc = ADRT_B200_OK;
    auto body = [&]() -> int {
        if (!ar.init) {
            for (auto &sl : ar.slot) {
                ADRT_CUDA_CHECK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
                ADRT_CUDA_CHECK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
            }
            ar.init = true;
        }
        const bool pin_in = is_pinned(h_in), pin_out = is_pinned(h_out);
        // chunk: bound (in + out) per slot by the budget, at least one item
        int64_t cb = (int64_t)(chunk_budget_bytes() / (in_item + out_item));
        cb = std::max<int64_t>(1, std::min<int64_t>(cb, B));
        // balance the chunks
        const int64_t nchunks = (B + cb - 1) / cb;
        cb = (B + nchunks - 1) / nchunks;
        for (auto &sl : ar.slot) { sl.pending_dst = nullptr; sl.pending_bytes = 0; }
        int64_t done_items = 0;
        for (int64_t ci = 0; done_items < B; ++ci) {
            Slot &sl = ar.slot[ci % kSlots];
            const int64_t b = std::min(cb, B - done_items);
            const size_t in_bytes = in_item * b, out_bytes = out_item * b, ws_bytes = ws_fn(b);
            // the slot's previous chunk must be fully drained before reuse
            int r = flush_pending(sl);
            if (r) return r;
            ADRT_CUDA_CHECK(cudaStreamSynchronize(sl.stream));
            if ((r = grow(&sl.d_in, &sl.cap_in, in_bytes, false))) return r;
            if ((r = grow(&sl.d_out, &sl.cap_out, out_bytes, false))) return r;
            if ((r = grow(&sl.d_ws, &sl.cap_ws, ws_bytes, false))) return r;
            const char *src = (const char *)h_in + in_item * done_items;
            char *dst = (char *)h_out + out_item * done_items;
            if (pin_in) {
                ADRT_CUDA_CHECK(cudaMemcpyAsync(sl.d_in, src, in_bytes, cudaMemcpyHostToDevice, sl.stream));
            } else {
                if ((r = grow(&sl.p_in, &sl.pcap_in, in_bytes, true))) return r;
                parallel_memcpy(sl.p_in, src, in_bytes);
                ADRT_CUDA_CHECK(cudaMemcpyAsync(sl.d_in, sl.p_in, in_bytes, cudaMemcpyHostToDevice, sl.stream));
            }
            if ((r = op(sl.d_in, sl.d_out, b, sl.d_ws, ws_bytes, sl.stream))) return r;
            if (pin_out) {
                ADRT_CUDA_CHECK(cudaMemcpyAsync(dst, sl.d_out, out_bytes, cudaMemcpyDeviceToHost, sl.stream));
            } else {
                if ((r = grow(&sl.p_out, &sl.pcap_out, out_bytes, true))) return r;
                ADRT_CUDA_CHECK(cudaMemcpyAsync(sl.p_out, sl.d_out, out_bytes, cudaMemcpyDeviceToHost, sl.stream));
                ADRT_CUDA_CHECK(cudaEventRecord(sl.done, sl.stream));
                sl.pending_dst = dst;
                sl.pending_bytes = out_bytes;
            }
            done_items += b;
        }
        for (auto &sl : ar.slot) {
            int r = flush_pending(sl);
            if (r) return r;
            ADRT_CUDA_CHECK(cudaStreamSynchronize(sl.stream));
        }
        return ADRT_B200_OK;
    };
    rc = body();
    if (rc != ADRT_B200_OK) {
        // leave the device in a clean state for the next call
        cudaDeviceSynchronize();
        cudaGetLastError();
        for (auto &sl : ar.slot) { sl.pending_dst = nullptr; sl.pending_bytes = 0; }
    }
    cudaSetDevice(prev_dev);
    return rc;
}

inline size_t sino_bytes(int64_t n, int dtype) { return (size_t)4 * (2 * n - 1) * n * dtype_size(dtype); }
inline size_t img_bytes(int64_t n, int dtype) { return (size_t)n * n * dtype_size(dtype); }

int check_host(int64_t B, int64_t n, int dtype)
{
    ADRT_REQUIRE(dtype_ok(dtype), "unsupported dtype %d", dtype);
    ADRT_REQUIRE(B > 0, "batch must be positive");
    ADRT_REQUIRE(is_pow2(n) && n <= kMaxN, "n must be a power of two <= %lld", (long long)kMaxN);
    return ADRT_B200_OK;
}

const WsFn kNoWs = [](int64_t) { return size_t(0); };

}  // namespace
}  // namespace adrt_b200

using namespace adrt_b200;

extern "C" {

int adrt_b200_host_adrt(const void *in, void *out, int64_t B, int64_t n, int dtype, int device)
{
    int rc = check_host(B, n, dtype);
    if (rc) return rc;
    return run_host(in, out, B, img_bytes(n, dtype), sino_bytes(n, dtype),
                    [=](int64_t b) { return adrt_b200_adrt_workspace_bytes(b, n, dtype); },
                    [=](const void *di, void *dout, int64_t b, void *ws, size_t wsb, cudaStream_t s) {
                        return adrt_b200_adrt(di, dout, b, n, dtype, ws, wsb, s);
                    },
                    device);
}

int adrt_b200_host_bdrt(const void *in, void *out, int64_t B, int64_t n, int dtype, int device)
{
    int rc = check_host(B, n, dtype);
    if (rc) return rc;
    return run_host(in, out, B, sino_bytes(n, dtype), sino_bytes(n, dtype),
                    [=](int64_t b) { return adrt_b200_bdrt_workspace_bytes(b, n, dtype); },
                    [=](const void *di, void *dout, int64_t b, void *ws, size_t wsb, cudaStream_t s) {
                        return adrt_b200_bdrt(di, dout, b, n, dtype, ws, wsb, s);
                    },
                    device);
}

int adrt_b200_host_iadrt(const void *in, void *out, int64_t B, int64_t n, int dtype, int device)
{
    int rc = check_host(B, n, dtype);
    if (rc) return rc;
    return run_host(in, out, B, sino_bytes(n, dtype), sino_bytes(n, dtype),
                    [=](int64_t b) { return adrt_b200_iadrt_workspace_bytes(b, n, dtype); },
                    [=](const void *di, void *dout, int64_t b, void *ws, size_t wsb, cudaStream_t s) {
                        return adrt_b200_iadrt(di, dout, b, n, dtype, ws, wsb, s);
                    },
                    device);
}

int adrt_b200_host_adrt_step(const void *in, void *out, int64_t B, int64_t n, int step, int dtype, int device)
{
    int rc = check_host(B, n, dtype);
    if (rc) return rc;
    ADRT_REQUIRE(step >= 0 && step < num_iters(n), "step %d is out of range for n=%lld", step, (long long)n);
    return run_host(in, out, B, sino_bytes(n, dtype), sino_bytes(n, dtype), kNoWs,
                    [=](const void *di, void *dout, int64_t b, void *, size_t, cudaStream_t s) {
                        return adrt_b200_adrt_step(di, dout, b, n, step, dtype, s);
                    },
                    device);
}

int adrt_b200_host_bdrt_step(const void *in, void *out, int64_t B, int64_t n, int step, int dtype, int device)
{
    int rc = check_host(B, n, dtype);
    if (rc) return rc;
    ADRT_REQUIRE(step >= 0 && step < num_iters(n), "step %d is out of range for n=%lld", step, (long long)n);
    return run_host(in, out, B, sino_bytes(n, dtype), sino_bytes(n, dtype), kNoWs,
                    [=](const void *di, void *dout, int64_t b, void *, size_t, cudaStream_t s) {
                        return adrt_b200_bdrt_step(di, dout, b, n, step, dtype, s);
                    },
                    device);
}

int adrt_b200_host_adrt_init(const void *in, void *out, int64_t B, int64_t n, int dtype, int device)
{
    int rc = check_host(B, n, dtype);
    if (rc) return rc;
    return run_host(in, out, B, img_bytes(n, dtype), sino_bytes(n, dtype), kNoWs,
                    [=](const void *di, void *dout, int64_t b, void *, size_t, cudaStream_t s) {
                        return adrt_b200_adrt_init(di, dout, b, n, dtype, s);
                    },
                    device);
}

int adrt_b200_host_fmg_restriction(const void *in, void *out, int64_t B, int64_t n, int dtype, int device)
{
    ADRT_REQUIRE(dtype_ok(dtype) && B > 0 && n >= 2 && n % 2 == 0 && n <= kMaxN, "bad argument");
    return run_host(in, out, B, sino_bytes(n, dtype), (size_t)4 * (n - 1) * (n / 2) * dtype_size(dtype), kNoWs,
                    [=](const void *di, void *dout, int64_t b, void *, size_t, cudaStream_t s) {
                        return adrt_b200_fmg_restriction(di, dout, b, n, dtype, s);
                    },
                    device);
}

int adrt_b200_host_fmg_prolongation(const void *in, void *out, int64_t B, int64_t h, int64_t w, int dtype, int device)
{
    ADRT_REQUIRE(dtype_ok(dtype) && B > 0 && h > 0 && w > 0, "bad argument");
    return run_host(in, out, B, (size_t)h * w * dtype_size(dtype), (size_t)4 * h * w * dtype_size(dtype), kNoWs,
                    [=](const void *di, void *dout, int64_t b, void *, size_t, cudaStream_t s) {
                        return adrt_b200_fmg_prolongation(di, dout, b, h, w, dtype, s);
                    },
                    device);
}

int adrt_b200_host_fmg_highpass(const void *in, void *out, int64_t B, int64_t h, int64_t w, int dtype, int device)
{
    ADRT_REQUIRE(dtype_ok(dtype) && B > 0 && h >= 2 && w >= 2, "bad argument");
    return run_host(in, out, B, (size_t)h * w * dtype_size(dtype), (size_t)h * w * dtype_size(dtype), kNoWs,
                    [=](const void *di, void *dout, int64_t b, void *, size_t, cudaStream_t s) {
                        return adrt_b200_fmg_highpass(di, dout, b, h, w, dtype, s);
                    },
                    device);
}

int adrt_b200_host_interp_to_cart(const void *in, void *out, int64_t B, int64_t n, int dtype, int device)
{
    int rc = check_host(B, n, dtype);
    if (rc) return rc;
    ADRT_REQUIRE(n >= 2, "interp_to_cart needs n >= 2");
    return run_host(in, out, B, sino_bytes(n, dtype), (size_t)n * 4 * n * dtype_size(dtype), kNoWs,
                    [=](const void *di, void *dout, int64_t b, void *, size_t, cudaStream_t s) {
                        return adrt_b200_interp_to_cart(di, dout, b, n, dtype, s);
                    },
                    device);
}

void *adrt_b200_host_alloc_pinned(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaHostAlloc of %zu bytes failed", bytes);
        return nullptr;
    }
    return p;
}

void adrt_b200_host_free_pinned(void *p)
{
    if (p) cudaFreeHost(p);
}

int adrt_b200_host_is_pinned(const void *p) { return p && adrt_b200::is_pinned(p) ? 1 : 0; }

}  // extern "C"
