// C ABI of libadrt_b200.so (declared in include/adrt_b200.h): argument
// validation, dtype dispatch, the per-stage fallback drivers, the cached
// interp_to_cart tables and the host-pointer (NumPy) pipeline.
#include "common.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

namespace adrt_b200 {

std::atomic<int64_t> g_launch_count{0};
std::atomic<int> g_mode{0};

static thread_local char t_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

namespace {

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// entry points that run the fused passes (16/32-byte vector accesses on the caller's arrays)
int check_vec_aligned(const void *in, const void *out, const void *ws)
{
    auto ok = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 32 == 0; };
    ADRT_REQUIRE(ok(in) && ok(out) && ok(ws), "device pointers must be 32-byte aligned");
    return ADRT_B200_OK;
}

int check_image(const void *in, const void *out, int64_t B, int64_t n, int dtype)
{
    ADRT_REQUIRE(in && out, "null pointer argument");
    // every transform here is out of place: CTAs read tiles of `in` while others write `out`
    ADRT_REQUIRE(in != out, "input and output must not alias (no transform runs in place)");
    ADRT_REQUIRE(dtype_ok(dtype), "unsupported dtype %d", dtype);
    ADRT_REQUIRE(B > 0, "batch must be positive, got %lld", (long long)B);
    ADRT_REQUIRE(is_pow2(n) && n <= kMaxN, "n must be a power of two <= %lld, got %lld", (long long)kMaxN, (long long)n);
    return ADRT_B200_OK;
}

// ---- per-stage drivers (also the oracle-like definition of the full ops) -----
template <typename T>
int adrt_by_steps(const T *in, T *out, int64_t B, int64_t n, T *ws, cudaStream_t s)
{
    const int K = num_iters(n);
    T *a = (K % 2 == 0) ? out : ws, *b = (K % 2 == 0) ? ws : out;
    int rc = launch_adrt_init<T>(in, a, B, n, s);
    for (int i = 0; i < K && rc == ADRT_B200_OK; ++i) {
        rc = launch_adrt_step<T>(a, b, B, n, i, s);
        T *t = a; a = b; b = t;
    }
    return rc;
}

template <typename T>
int bdrt_by_steps(const T *in, T *out, int64_t B, int64_t n, T *ws, cudaStream_t s)
{
    const int K = num_iters(n);
    if (K == 0) {
        ADRT_CUDA_CHECK(cudaMemcpyAsync(out, in, sizeof(T) * sino_elems(B, n), cudaMemcpyDeviceToDevice, s));
        return ADRT_B200_OK;
    }
    const T *src = in;
    T *a = (K % 2 == 1) ? out : ws, *b = (K % 2 == 1) ? ws : out;
    int rc = ADRT_B200_OK;
    for (int i = 0; i < K && rc == ADRT_B200_OK; ++i) {
        rc = launch_bdrt_step<T>(src, a, B, n, i, /*core_semantics=*/true, s);
        src = a;
        T *t = a; a = b; b = t;
    }
    return rc;
}

template <typename T>
int iadrt_by_stages(const T *in, T *out, int64_t B, int64_t n, T *ws, cudaStream_t s)
{
    const int K = num_iters(n);
    if (K == 0) {
        ADRT_CUDA_CHECK(cudaMemcpyAsync(out, in, sizeof(T) * sino_elems(B, n), cudaMemcpyDeviceToDevice, s));
        return ADRT_B200_OK;
    }
    const T *src = in;
    T *a = (K % 2 == 1) ? out : ws, *b = (K % 2 == 1) ? ws : out;
    int rc = ADRT_B200_OK;
    for (int i = 0; i < K && rc == ADRT_B200_OK; ++i) {
        rc = launch_iadrt_stage<T>(src, a, B, n, i, s);
        src = a;
        T *t = a; a = b; b = t;
    }
    return rc;
}

template <typename T>
size_t adrt_ws_elems(int64_t B, int64_t n)
{
    if (g_mode.load() == 0) {
        size_t f = fused_adrt_workspace_elems<T>(B, n, 4);
        if (f != (size_t)-1) return f;
    }
    return num_iters(n) == 0 ? 0 : (size_t)sino_elems(B, n);
}

template <typename T>
size_t bdrt_ws_elems(int64_t B, int64_t n)
{
    if (g_mode.load() == 0) {
        size_t f = fused_bdrt_workspace_elems<T>(B, n, 4);
        if (f != (size_t)-1) return f;
    }
    return num_iters(n) <= 1 ? 0 : (size_t)sino_elems(B, n);
}

template <typename T>
int adrt_impl(const T *in, T *out, int64_t B, int64_t n, T *ws, size_t ws_bytes, cudaStream_t s)
{
    const size_t need = adrt_ws_elems<T>(B, n) * sizeof(T);
    if (need > 0 && (!ws || ws_bytes < need)) {
        set_error("adrt workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        return ADRT_B200_EWORKSPACE;
    }
    if (g_mode.load() == 0) {
        bool handled = false;
        int rc = fused_adrt<T>(in, out, B, n, 0, 4, ws, ws_bytes / sizeof(T), s, &handled);
        if (rc != ADRT_B200_OK || handled) return rc;
    }
    return adrt_by_steps<T>(in, out, B, n, ws, s);
}

template <typename T>
int bdrt_impl(const T *in, T *out, int64_t B, int64_t n, T *ws, size_t ws_bytes, cudaStream_t s)
{
    const size_t need = bdrt_ws_elems<T>(B, n) * sizeof(T);
    if (need > 0 && (!ws || ws_bytes < need)) {
        set_error("bdrt workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        return ADRT_B200_EWORKSPACE;
    }
    if (g_mode.load() == 0) {
        bool handled = false;
        int rc = fused_bdrt<T>(in, out, B, n, 4, -1, ws, ws_bytes / sizeof(T), s, &handled);
        if (rc != ADRT_B200_OK || handled) return rc;
    }
    return bdrt_by_steps<T>(in, out, B, n, ws, s);
}

// ---- interp_to_cart tables ------------------------------------------------------
// Host half of adrt_cdefs_interp_adrtcart.hpp:61-114 with float_index = float
// (adrt_cdefs_py.cpp:638): everything that involves libm transcendentals
// (tanf, cosf) depends on the output column only, so it is evaluated here on
// the host with the very expressions the reference uses and shipped as O(n)
// tables; the device finishes with correctly-rounded IEEE div/mul/add and
// exact roundf, which are bit-identical to the host's.
struct InterpTable {
    float *t = nullptr;         // [n]   t_left * lerp(-1,1,offset/(n-1))
    int32_t *base = nullptr;    // [4n]  q*D*n + si
    float *h_base = nullptr;    // [4n]  0.5f + tan/2
    float *cosv = nullptr;      // [4n]  cosf(th0)
    int32_t *sgn = nullptr;     // [4n]  1 if even quadrant else 0
    void *factor = nullptr;     // [4n]  T
};

std::mutex g_interp_mu;
std::map<std::tuple<int, int64_t, int>, InterpTable> g_interp_cache;

int get_interp_table(int device, int64_t n, int dtype, cudaStream_t s, InterpTable *out)
{
    std::lock_guard<std::mutex> lock(g_interp_mu);
    auto key = std::make_tuple(device, n, dtype);
    auto it = g_interp_cache.find(key);
    if (it != g_interp_cache.end()) { *out = it->second; return ADRT_B200_OK; }

    const int64_t D = 2 * n - 1, W = 4 * n;
    std::vector<float> t(n), hb(W), cv(W);
    std::vector<int32_t> base(W), sg(W);
    std::vector<float> f32(W);
    std::vector<double> f64(W);
    const float sqrt2_2 = (float)1.41421356237309504880168872420969808L / 2.0f;
    const float pi = (float)3.14159265358979323846264338327950288L;
    const float pi_2 = pi / 2.0f, pi_4 = pi / 4.0f, pi_8 = pi / 8.0f;
    const float t_left = sqrt2_2 - (sqrt2_2 / (float)n);
    const float th_left = pi_2 - (pi_8 / (float)n);
    for (int64_t off = 0; off < n; ++off) {
        const volatile float of = (float)off / (float)(n - 1);
        const volatile float p1 = of * 1.0f;
        const volatile float p2 = (1.0f - of) * -1.0f;
        const volatile float l = p1 + p2;   // std::lerp(-1, 1, of), opposite-sign branch
        t[off] = t_left * l;
    }
    for (int64_t ang = 0; ang < W; ++ang) {
        const volatile float af = (float)ang / (float)(W - 1);
        const volatile float p1 = af * -1.0f;
        const volatile float p2 = (1.0f - af) * 1.0f;
        const volatile float l = p1 + p2;   // std::lerp(1, -1, af)
        const float th = th_left * l;
        float qf = -th / pi_4;
        qf = qf < -2.0f ? -2.0f : (qf > 1.0f ? 1.0f : qf);
        const int q = (int)(std::floor(qf) + 2);
        const float th0 = pi_4 - std::fabs(std::fabs(th) - pi_4);
        float tant = std::tan(th0);
        tant = tant < 0.0f ? 0.0f : (tant > 1.0f ? 1.0f : tant);
        const float si = std::round(tant * (float)(n - 1));
        {
            const volatile float sa = si / (float)(n - 1);
            const volatile float sq = sa * sa;
            f32[ang] = std::sqrt(sq + 1.0f);
            const volatile double sad = (double)si / (double)(n - 1);
            const volatile double sqd = sad * sad;
            f64[ang] = std::sqrt(sqd + 1.0);
        }
        const volatile float half_tan = tant / 2.0f;
        hb[ang] = 0.5f + half_tan;
        cv[ang] = std::cos(th0);
        sg[ang] = (q % 2 == 0) ? 1 : 0;
        base[ang] = (int32_t)((int64_t)q * D * n + (int64_t)si);
    }
    InterpTable tab;
    const size_t fsz = dtype == ADRT_B200_F64 ? sizeof(double) : sizeof(float);
    // One allocation for the six tables (256-byte aligned pieces): nothing to unwind piecewise when
    // an allocation or a copy fails.  Synchronous copies from pageable memory: the vectors die at
    // scope exit.  (First use of an (n, dtype) on a device therefore must not happen inside a stream
    // capture; later calls only launch the gather kernel.)
    auto pad = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t o_t = 0, o_base = o_t + pad(n * sizeof(float)), o_hb = o_base + pad(W * sizeof(int32_t)),
                 o_cos = o_hb + pad(W * sizeof(float)), o_sgn = o_cos + pad(W * sizeof(float)),
                 o_fac = o_sgn + pad(W * sizeof(int32_t)), total = o_fac + pad(W * fsz);
    char *blob = nullptr;
    ADRT_CUDA_CHECK(cudaMalloc(&blob, total));
    auto up = [&](size_t off, const void *src, size_t bytes) { return cudaMemcpy(blob + off, src, bytes, cudaMemcpyHostToDevice); };
    cudaError_t ce = up(o_t, t.data(), n * sizeof(float));
    if (ce == cudaSuccess) ce = up(o_base, base.data(), W * sizeof(int32_t));
    if (ce == cudaSuccess) ce = up(o_hb, hb.data(), W * sizeof(float));
    if (ce == cudaSuccess) ce = up(o_cos, cv.data(), W * sizeof(float));
    if (ce == cudaSuccess) ce = up(o_sgn, sg.data(), W * sizeof(int32_t));
    if (ce == cudaSuccess) ce = up(o_fac, dtype == ADRT_B200_F64 ? (const void *)f64.data() : (const void *)f32.data(), W * fsz);
    if (ce != cudaSuccess) {
        cudaFree(blob);
        set_error("interp_to_cart table upload failed: %s", cudaGetErrorString(ce));
        return ADRT_B200_ECUDA;
    }
    tab.t = reinterpret_cast<float *>(blob + o_t);
    tab.base = reinterpret_cast<int32_t *>(blob + o_base);
    tab.h_base = reinterpret_cast<float *>(blob + o_hb);
    tab.cosv = reinterpret_cast<float *>(blob + o_cos);
    tab.sgn = reinterpret_cast<int32_t *>(blob + o_sgn);
    tab.factor = blob + o_fac;
    (void)s;
    g_interp_cache[key] = tab;
    *out = tab;
    return ADRT_B200_OK;
}

}  // namespace

}  // namespace adrt_b200

using namespace adrt_b200;

template <typename T>
int adrt_quadrants_impl(const T *in, T *out, int64_t B, int64_t n, int q_first, int q_count, T *ws, size_t ws_bytes,
                        cudaStream_t s)
{
    if (n == 1) {  // K = 0: every quadrant is the image itself
        for (int64_t b = 0; b < B; ++b)
            for (int q = 0; q < q_count; ++q)
                ADRT_CUDA_CHECK(cudaMemcpyAsync(out + b * q_count + q, in + b, sizeof(T), cudaMemcpyDeviceToDevice, s));
        return ADRT_B200_OK;
    }
    const size_t need = fused_adrt_workspace_elems<T>(B, n, q_count) * sizeof(T);
    if (need > 0 && (!ws || ws_bytes < need)) {
        set_error("adrt_quadrants workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        return ADRT_B200_EWORKSPACE;
    }
    bool handled = false;
    return fused_adrt<T>(in, out, B, n, q_first, q_count, ws, ws_bytes / sizeof(T), s, &handled);
}

template <typename T>
int bdrt_planes_impl(const T *in, T *out, int64_t planes, int64_t n, int64_t rows, T *ws, size_t ws_bytes, cudaStream_t s,
                     const T *sub = nullptr)
{
    if (n == 1) {
        ADRT_CUDA_CHECK(cudaMemcpyAsync(out, in, sizeof(T) * planes, cudaMemcpyDeviceToDevice, s));
        return ADRT_B200_OK;
    }
    const size_t need = fused_bdrt_workspace_elems<T>(planes, n, 1) * sizeof(T);
    if (need > 0 && (!ws || ws_bytes < need)) {
        set_error("bdrt_planes workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        return ADRT_B200_EWORKSPACE;
    }
    bool handled = false;
    return fused_bdrt<T>(in, out, planes, n, 1, rows, ws, ws_bytes / sizeof(T), s, &handled, false, sub);
}

// ---- fused normal operator ---------------------------------------
// bdrt(adrt(x)) for quadrants q_first .. q_first + q_count - 1, offsets d < n only (all that truncate
// keeps).  The sinogram never exists in the public (d, column) layout: the last forward pass stores
// R-layout rows and the first transposed pass loads them (fused_plan.h rows_out / rows_in), so the
// transposing public-layout store of adrt and load of bdrt are gone.  Same adds in the same order as
// the two separate transforms => bit-identical.
template <typename T>
size_t normal_rows_elems(int64_t B, int64_t n, int q_count) { return (size_t)(B * q_count) * (size_t)n * (size_t)((2 * n - 1 + 3) & ~int64_t(3)); }

template <typename T>
size_t normal_transform_ws_elems(int64_t B, int64_t n, int q_count)
{
    const size_t a = fused_adrt_workspace_elems<T>(B, n, q_count), b = fused_bdrt_workspace_elems<T>(B * q_count, n, 1);
    if (a == (size_t)-1 || b == (size_t)-1) return (size_t)-1;
    return a > b ? a : b;
}

template <typename T>
int adrt_bdrt_rows_impl(const T *in, T *out, int64_t B, int64_t n, int q_first, int q_count, T *ws, size_t ws_elems,
                        cudaStream_t s)
{
    const size_t mid = (normal_rows_elems<T>(B, n, q_count) + 63) & ~size_t(63);
    const size_t tws = normal_transform_ws_elems<T>(B, n, q_count);
    if (tws == (size_t)-1 || !ws || ws_elems < mid + tws) {
        set_error("normal operator workspace too small: need %zu elements, got %zu", mid + tws, ws_elems);
        return ADRT_B200_EWORKSPACE;
    }
    bool handled = false;
    int rc = fused_adrt<T>(in, ws, B, n, q_first, q_count, ws + mid, ws_elems - mid, s, &handled, /*rows_out=*/true);
    if (rc != ADRT_B200_OK) return rc;
    if (!handled) { set_error("internal: no fused forward plan for n=%lld", (long long)n); return ADRT_B200_EINVAL; }
    rc = fused_bdrt<T>(ws, out, B * q_count, n, 1, n, ws + mid, ws_elems - mid, s, &handled, /*rows_in=*/true);
    if (rc == ADRT_B200_OK && !handled) { set_error("internal: no fused transposed plan for n=%lld", (long long)n); return ADRT_B200_EINVAL; }
    return rc;
}

template <typename T>
size_t normal_operator_ws_elems(int64_t B, int64_t n)
{
    if (n < 2) return 0;
    const size_t tws = normal_transform_ws_elems<T>(B, n, 4);
    if (tws == (size_t)-1) return (size_t)-1;
    return ((normal_rows_elems<T>(B, n, 4) + 63) & ~size_t(63)) + tws + (((size_t)sino_elems(B, n) + 63) & ~size_t(63));
}

template <typename T>
int normal_operator_impl(const T *in, T *out, int64_t B, int64_t n, T divisor, T *ws, size_t ws_bytes, cudaStream_t s)
{
    if (n == 1) {
        // adrt and bdrt are identities on the four copies of the pixel: ((x/div + x/div) + x/div + x/div) / 4
        const size_t need = (size_t)B * 4 * sizeof(T);
        if (!ws || ws_bytes < need) { set_error("normal operator workspace too small"); return ADRT_B200_EWORKSPACE; }
        int rc = launch_adrt_init<T>(in, ws, B, 1, s);
        if (rc) return rc;
        return launch_truncate_mean<T>(ws, out, B, 1, divisor, s);
    }
    const size_t need = normal_operator_ws_elems<T>(B, n);
    if (need == (size_t)-1 || !ws || ws_bytes / sizeof(T) < need) {
        set_error("normal operator workspace too small: need %zu bytes, got %zu", need * sizeof(T), ws_bytes);
        return ADRT_B200_EWORKSPACE;
    }
    const size_t back = ((size_t)sino_elems(B, n) + 63) & ~size_t(63);
    T *back_buf = ws, *rest = ws + back;
    int rc = adrt_bdrt_rows_impl<T>(in, back_buf, B, n, 0, 4, rest, ws_bytes / sizeof(T) - back, s);
    if (rc) return rc;
    return launch_truncate_mean<T>(back_buf, out, B, n, divisor, s);
}

// ---- one full-multigrid pass, all levels, no host round trips --------------------
// core.py:318-331 (iadrt_fmg_step): restrict the sinogram down to 1 x 1, then per level
//   ret = prolongation(ret); ret -= highpass(mean_q(truncate(bdrt(adrt(ret) - a_level)) / (m - 1)))
// with the reference's operator order, so the result is bit-identical to composing the
// public operators.  Workspace segments are padded to 64 elements.
inline size_t pad64(size_t v) { return (v + 63) & ~size_t(63); }

template <typename T>
size_t fmg_transform_ws_elems(int64_t B, int64_t n)
{
    size_t a = adrt_ws_elems<T>(B, n), b = bdrt_ws_elems<T>(B, n);
    if (g_mode.load() == 0 && n >= 2) {
        const size_t c = fused_bdrt_workspace_elems<T>(B * 4, n, 1);
        if (c != (size_t)-1 && c > b) b = c;
    }
    return a > b ? a : b;
}

template <typename T>
size_t fmg_step_ws_elems(int64_t B, int64_t n)
{
    size_t total = 0;
    for (int64_t m = n / 2; m >= 1; m /= 2) total += pad64((size_t)sino_elems(B, m));
    total += 2 * pad64((size_t)sino_elems(B, n)) + 4 * pad64((size_t)(B * n * n));
    return total + pad64(fmg_transform_ws_elems<T>(B, n));
}

template <typename T>
int fmg_step_impl(const T *in, T *out, int64_t B, int64_t n, T *ws, size_t ws_bytes, cudaStream_t s)
{
    const size_t need = fmg_step_ws_elems<T>(B, n) * sizeof(T);
    if (!ws || ws_bytes < need) {
        set_error("fmg_step workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        return ADRT_B200_EWORKSPACE;
    }
    const int K = num_iters(n);
    // restricted sinograms: level[0] = caller's input, level[k] has side n >> k
    std::vector<const T *> level(K + 1);
    level[0] = in;
    T *p = ws;
    int rc = ADRT_B200_OK;
    for (int k = 1; k <= K && rc == ADRT_B200_OK; ++k) {
        rc = launch_fmg_restriction<T>(level[k - 1], p, B, n >> (k - 1), s);
        level[k] = p;
        p += pad64((size_t)sino_elems(B, n >> k));
    }
    if (rc != ADRT_B200_OK) return rc;
    T *sino_a = p; p += pad64((size_t)sino_elems(B, n));
    T *sino_b = p; p += pad64((size_t)sino_elems(B, n));
    T *img[2] = {p, p + pad64((size_t)(B * n * n))};
    p += 2 * pad64((size_t)(B * n * n));
    T *grad = p; p += pad64((size_t)(B * n * n));
    T *hp = p; p += pad64((size_t)(B * n * n));
    T *tws = p;
    const size_t tws_bytes = ws_bytes - (size_t)(tws - ws) * sizeof(T);

    // ret = a_K[..., 0, :, :]: element 0 of every image's (4, 1, 1) block
    T *cur = K == 0 ? out : img[0];
    ADRT_CUDA_CHECK(cudaMemcpy2DAsync(cur, sizeof(T), level[K], 4 * sizeof(T), sizeof(T), (size_t)B,
                                      cudaMemcpyDeviceToDevice, s));
    int which = 0;
    for (int k = K - 1; k >= 0; --k) {
        const int64_t m = n >> k;
        T *pro = img[which ^ 1];
        if ((rc = launch_fmg_prolongation<T>(cur, pro, B, m / 2, m / 2, s))) return rc;
        if ((rc = adrt_impl<T>(pro, sino_a, B, m, tws, tws_bytes, s))) return rc;
        // adrt(ret) - a_level: subtracted by the loader of bdrt's first pass where that pass is one of the
        // public-layout kernels with a fused first step (large levels: three sinogram sweeps less), by a
        // kernel of its own otherwise; element-wise either way, so the result does not depend on which
        const bool sub_on_load = g_mode.load() == 0 && m >= 2 && getenv("ADRT_B200_FMG_SUB_SEPARATE") == nullptr &&
                                 fused_bdrt_sub_ok<T>(m, m);
        if (!sub_on_load && (rc = launch_binary<T>(sino_a, level[k], sino_a, sino_elems(B, m), 0, s))) return rc;
        if (g_mode.load() == 0) rc = bdrt_planes_impl<T>(sino_a, sino_b, B * 4, m, m, tws, tws_bytes, s, sub_on_load ? level[k] : (const T *)nullptr);
        else rc = bdrt_impl<T>(sino_a, sino_b, B, m, tws, tws_bytes, s);
        if (rc) return rc;
        if ((rc = launch_truncate_mean<T>(sino_b, grad, B, m, (T)(m - 1), s))) return rc;
        if ((rc = launch_fmg_highpass<T>(grad, hp, B, m, m, s))) return rc;
        if ((rc = launch_binary<T>(pro, hp, k == 0 ? out : pro, B * m * m, 0, s))) return rc;
        cur = pro;
        which ^= 1;
    }
    return ADRT_B200_OK;
}

#define DISPATCH(dtype, CALL_F32, CALL_F64) ((dtype) == ADRT_B200_F64 ? (CALL_F64) : (CALL_F32))

extern "C" {

int adrt_b200_version(void) { return 100; }
const char *adrt_b200_last_error(void) { return t_error; }
int adrt_b200_device_count(void)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return -1; }
    return c;
}
void adrt_b200_set_mode(int mode) { g_mode.store(mode); }
int adrt_b200_get_mode(void) { return g_mode.load(); }
int64_t adrt_b200_launch_count(void) { return g_launch_count.load(); }
int adrt_b200_num_iters(int64_t n) { return num_iters(n); }

size_t adrt_b200_adrt_workspace_bytes(int64_t B, int64_t n, int dtype)
{
    if (B <= 0 || !is_pow2(n) || !dtype_ok(dtype)) return 0;
    return DISPATCH(dtype, adrt_ws_elems<float>(B, n) * 4, adrt_ws_elems<double>(B, n) * 8);
}

int adrt_b200_adrt(const void *in, void *out, int64_t B, int64_t n, int dtype, void *ws, size_t ws_bytes, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    if ((rc = check_vec_aligned(in, out, ws))) return rc;
    return DISPATCH(dtype,
                    adrt_impl<float>((const float *)in, (float *)out, B, n, (float *)ws, ws_bytes, as_stream(stream)),
                    adrt_impl<double>((const double *)in, (double *)out, B, n, (double *)ws, ws_bytes, as_stream(stream)));
}

size_t adrt_b200_bdrt_workspace_bytes(int64_t B, int64_t n, int dtype)
{
    if (B <= 0 || !is_pow2(n) || !dtype_ok(dtype)) return 0;
    return DISPATCH(dtype, bdrt_ws_elems<float>(B, n) * 4, bdrt_ws_elems<double>(B, n) * 8);
}

int adrt_b200_bdrt(const void *in, void *out, int64_t B, int64_t n, int dtype, void *ws, size_t ws_bytes, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    if ((rc = check_vec_aligned(in, out, ws))) return rc;
    return DISPATCH(dtype,
                    bdrt_impl<float>((const float *)in, (float *)out, B, n, (float *)ws, ws_bytes, as_stream(stream)),
                    bdrt_impl<double>((const double *)in, (double *)out, B, n, (double *)ws, ws_bytes, as_stream(stream)));
}

// ---- quadrant / plane subsets (single-image multi-GPU sharding) -------------------
size_t adrt_b200_adrt_quadrants_workspace_bytes(int64_t B, int64_t n, int dtype, int q_count)
{
    if (B <= 0 || !is_pow2(n) || n > kMaxN || n < 2 || !dtype_ok(dtype) || q_count < 1 || q_count > 4) return 0;
    return DISPATCH(dtype, fused_adrt_workspace_elems<float>(B, n, q_count) * 4, fused_adrt_workspace_elems<double>(B, n, q_count) * 8);
}

int adrt_b200_adrt_quadrants(const void *in, void *out, int64_t B, int64_t n, int dtype, int q_first, int q_count,
                             void *ws, size_t ws_bytes, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    if ((rc = check_vec_aligned(in, out, ws))) return rc;
    ADRT_REQUIRE(q_first >= 0 && q_count >= 1 && q_first + q_count <= 4, "bad quadrant range %d+%d", q_first, q_count);
    return DISPATCH(dtype,
                    adrt_quadrants_impl<float>((const float *)in, (float *)out, B, n, q_first, q_count, (float *)ws, ws_bytes, as_stream(stream)),
                    adrt_quadrants_impl<double>((const double *)in, (double *)out, B, n, q_first, q_count, (double *)ws, ws_bytes, as_stream(stream)));
}

size_t adrt_b200_bdrt_planes_workspace_bytes(int64_t planes, int64_t n, int dtype)
{
    if (planes <= 0 || !is_pow2(n) || n > kMaxN || n < 2 || !dtype_ok(dtype)) return 0;
    return DISPATCH(dtype, fused_bdrt_workspace_elems<float>(planes, n, 1) * 4, fused_bdrt_workspace_elems<double>(planes, n, 1) * 8);
}

int adrt_b200_bdrt_planes(const void *in, void *out, int64_t planes, int64_t n, int dtype, void *ws, size_t ws_bytes,
                          void *stream)
{
    int rc = check_image(in, out, planes, n, dtype);
    if (rc) return rc;
    if ((rc = check_vec_aligned(in, out, ws))) return rc;
    return DISPATCH(dtype,
                    bdrt_planes_impl<float>((const float *)in, (float *)out, planes, n, -1, (float *)ws, ws_bytes, as_stream(stream)),
                    bdrt_planes_impl<double>((const double *)in, (double *)out, planes, n, -1, (double *)ws, ws_bytes, as_stream(stream)));
}

int adrt_b200_bdrt_rows(const void *in, void *out, int64_t planes, int64_t n, int64_t rows, int dtype, void *ws,
                        size_t ws_bytes, void *stream)
{
    int rc = check_image(in, out, planes, n, dtype);
    if (rc) return rc;
    if ((rc = check_vec_aligned(in, out, ws))) return rc;
    ADRT_REQUIRE(rows >= 1 && rows <= 2 * n - 1, "rows %lld out of range", (long long)rows);
    return DISPATCH(dtype,
                    bdrt_planes_impl<float>((const float *)in, (float *)out, planes, n, rows, (float *)ws, ws_bytes, as_stream(stream)),
                    bdrt_planes_impl<double>((const double *)in, (double *)out, planes, n, rows, (double *)ws, ws_bytes, as_stream(stream)));
}

// ---- fused normal operator -----------------------------------------------------------
size_t adrt_b200_normal_operator_workspace_bytes(int64_t B, int64_t n, int dtype)
{
    if (B <= 0 || !is_pow2(n) || n > kMaxN || !dtype_ok(dtype)) return 0;
    if (n == 1) return (size_t)B * 4 * dtype_size(dtype);
    const size_t e = DISPATCH(dtype, normal_operator_ws_elems<float>(B, n), normal_operator_ws_elems<double>(B, n));
    return e == (size_t)-1 ? 0 : e * dtype_size(dtype);
}

int adrt_b200_normal_operator(const void *in, void *out, int64_t B, int64_t n, double divisor, int dtype, void *ws,
                              size_t ws_bytes, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    if ((rc = check_vec_aligned(in, out, ws))) return rc;
    return DISPATCH(dtype,
                    normal_operator_impl<float>((const float *)in, (float *)out, B, n, (float)divisor, (float *)ws, ws_bytes, as_stream(stream)),
                    normal_operator_impl<double>((const double *)in, (double *)out, B, n, divisor, (double *)ws, ws_bytes, as_stream(stream)));
}

size_t adrt_b200_adrt_bdrt_rows_workspace_bytes(int64_t B, int64_t n, int dtype, int q_count)
{
    if (B <= 0 || !is_pow2(n) || n > kMaxN || n < 2 || !dtype_ok(dtype) || q_count < 1 || q_count > 4) return 0;
    const size_t t = DISPATCH(dtype, normal_transform_ws_elems<float>(B, n, q_count), normal_transform_ws_elems<double>(B, n, q_count));
    if (t == (size_t)-1) return 0;
    const size_t mid = DISPATCH(dtype, normal_rows_elems<float>(B, n, q_count), normal_rows_elems<double>(B, n, q_count));
    return (((mid + 63) & ~size_t(63)) + t) * dtype_size(dtype);
}

int adrt_b200_adrt_bdrt_rows(const void *in, void *out, int64_t B, int64_t n, int dtype, int q_first, int q_count, void *ws,
                             size_t ws_bytes, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    if ((rc = check_vec_aligned(in, out, ws))) return rc;
    ADRT_REQUIRE(n >= 2, "adrt_bdrt_rows needs n >= 2");
    ADRT_REQUIRE(q_first >= 0 && q_count >= 1 && q_first + q_count <= 4, "bad quadrant range %d+%d", q_first, q_count);
    return DISPATCH(dtype,
                    adrt_bdrt_rows_impl<float>((const float *)in, (float *)out, B, n, q_first, q_count, (float *)ws, ws_bytes / 4, as_stream(stream)),
                    adrt_bdrt_rows_impl<double>((const double *)in, (double *)out, B, n, q_first, q_count, (double *)ws, ws_bytes / 8, as_stream(stream)));
}

// ---- angle-block sharding of single large images --------------------------
static int check_part(int64_t n, int part, int parts, int m_last, int phase)
{
    ADRT_REQUIRE(parts >= 2 && (parts & (parts - 1)) == 0 && part >= 0 && part < parts, "bad part %d of %d", part, parts);
    ADRT_REQUIRE(m_last >= 1 && (1 << m_last) >= parts && m_last < num_iters(n), "bad m_last %d for n=%lld, parts=%d", m_last,
                 (long long)n, parts);
    ADRT_REQUIRE(phase == 0 || phase == 1, "phase must be 0 or 1");
    return ADRT_B200_OK;
}

size_t adrt_b200_part_exchange_pitch(int64_t n, int dtype, int m_last, int forward)
{
    if (!is_pow2(n) || n > kMaxN || !dtype_ok(dtype) || m_last < 1 || m_last >= num_iters(n)) return 0;
    return DISPATCH(dtype, part_exchange_pitch<float>(n, m_last, forward != 0), part_exchange_pitch<double>(n, m_last, forward != 0));
}

size_t adrt_b200_part_exchange_cols(int64_t n, int dtype, int m_last, int64_t rows)
{
    if (!is_pow2(n) || n > kMaxN || !dtype_ok(dtype) || m_last < 1 || m_last >= num_iters(n)) return 0;
    return DISPATCH(dtype, part_exchange_cols<float>(n, m_last, rows), part_exchange_cols<double>(n, m_last, rows));
}

size_t adrt_b200_part_workspace_bytes(int64_t planes, int64_t n, int dtype, int m_last)
{
    if (planes <= 0 || !is_pow2(n) || n > kMaxN || !dtype_ok(dtype) || m_last < 1 || m_last >= num_iters(n)) return 0;
    const size_t e = DISPATCH(dtype, part_workspace_elems<float>(planes, n, m_last), part_workspace_elems<double>(planes, n, m_last));
    return e == (size_t)-1 ? 0 : e * dtype_size(dtype);
}

int adrt_b200_adrt_part(const void *img, void *xbuf, void *sino, int64_t B, int64_t n, int dtype, int q_first, int q_count,
                        int part, int parts, int m_last, int phase, void *ws, size_t ws_bytes, void *stream)
{
    ADRT_REQUIRE(xbuf && (phase == 0 ? img != nullptr : sino != nullptr), "null pointer argument");
    ADRT_REQUIRE(dtype_ok(dtype) && B > 0 && is_pow2(n) && n <= kMaxN, "bad argument");
    ADRT_REQUIRE(q_first >= 0 && q_count >= 1 && q_first + q_count <= 4, "bad quadrant range %d+%d", q_first, q_count);
    int rc = check_part(n, part, parts, m_last, phase);
    if (rc) return rc;
    if ((rc = check_vec_aligned(phase == 0 ? img : sino, xbuf, ws))) return rc;
    return DISPATCH(dtype,
                    fused_adrt_part<float>((const float *)img, (float *)xbuf, (float *)sino, B, n, q_first, q_count, part, parts, m_last, phase, (float *)ws, ws_bytes / 4, as_stream(stream)),
                    fused_adrt_part<double>((const double *)img, (double *)xbuf, (double *)sino, B, n, q_first, q_count, part, parts, m_last, phase, (double *)ws, ws_bytes / 8, as_stream(stream)));
}

int adrt_b200_bdrt_part(const void *sino, void *xbuf, void *out, int64_t planes, int64_t n, int64_t rows, int dtype, int part,
                        int parts, int m_last, int phase, void *ws, size_t ws_bytes, void *stream)
{
    ADRT_REQUIRE(xbuf && (phase == 0 ? sino != nullptr : out != nullptr), "null pointer argument");
    ADRT_REQUIRE(dtype_ok(dtype) && planes > 0 && is_pow2(n) && n <= kMaxN, "bad argument");
    ADRT_REQUIRE(rows == -1 || (rows >= 1 && rows <= 2 * n - 1), "rows %lld out of range", (long long)rows);
    int rc = check_part(n, part, parts, m_last, phase);
    if (rc) return rc;
    if ((rc = check_vec_aligned(phase == 0 ? sino : out, xbuf, ws))) return rc;
    return DISPATCH(dtype,
                    fused_bdrt_part<float>((const float *)sino, (float *)xbuf, (float *)out, planes, n, rows, part, parts, m_last, phase, (float *)ws, ws_bytes / 4, as_stream(stream)),
                    fused_bdrt_part<double>((const double *)sino, (double *)xbuf, (double *)out, planes, n, rows, part, parts, m_last, phase, (double *)ws, ws_bytes / 8, as_stream(stream)));
}

int adrt_b200_adrt_step(const void *in, void *out, int64_t B, int64_t n, int step, int dtype, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    ADRT_REQUIRE(step >= 0 && step < num_iters(n), "step %d is out of range for n=%lld", step, (long long)n);
    return DISPATCH(dtype,
                    launch_adrt_step<float>((const float *)in, (float *)out, B, n, step, as_stream(stream)),
                    launch_adrt_step<double>((const double *)in, (double *)out, B, n, step, as_stream(stream)));
}

int adrt_b200_bdrt_step(const void *in, void *out, int64_t B, int64_t n, int step, int dtype, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    ADRT_REQUIRE(step >= 0 && step < num_iters(n), "step %d is out of range for n=%lld", step, (long long)n);
    return DISPATCH(dtype,
                    launch_bdrt_step<float>((const float *)in, (float *)out, B, n, step, false, as_stream(stream)),
                    launch_bdrt_step<double>((const double *)in, (double *)out, B, n, step, false, as_stream(stream)));
}

int adrt_b200_adrt_init(const void *in, void *out, int64_t B, int64_t n, int dtype, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    return DISPATCH(dtype,
                    launch_adrt_init<float>((const float *)in, (float *)out, B, n, as_stream(stream)),
                    launch_adrt_init<double>((const double *)in, (double *)out, B, n, as_stream(stream)));
}

size_t adrt_b200_iadrt_workspace_bytes(int64_t B, int64_t n, int dtype)
{
    if (B <= 0 || !is_pow2(n) || !dtype_ok(dtype)) return 0;
    // enough for either path: per-stage ping-pong (one sinogram) or the fused passes' column-major buffers
    const size_t steps = num_iters(n) <= 1 ? 0 : (size_t)sino_elems(B, n);
    const size_t fused = n > kMaxN ? 0 : DISPATCH(dtype, fused_iadrt_workspace_elems<float>(B, n), fused_iadrt_workspace_elems<double>(B, n));
    return (steps > fused ? steps : fused) * dtype_size(dtype);
}

int adrt_b200_iadrt(const void *in, void *out, int64_t B, int64_t n, int dtype, void *ws, size_t ws_bytes, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    const size_t need = adrt_b200_iadrt_workspace_bytes(B, n, dtype);
    if (need > 0 && (!ws || ws_bytes < need)) {
        set_error("iadrt workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        return ADRT_B200_EWORKSPACE;
    }
    // fused multi-stage passes (iadrt_fused.cu) need 32-byte aligned workspace rows (256-bit accesses); mode 1 and
    // misaligned workspaces take the one-kernel-per-stage path
    if (g_mode.load() == 0 && num_iters(n) >= 1 && reinterpret_cast<uintptr_t>(ws) % 32 == 0)
        return DISPATCH(dtype,
                        fused_iadrt<float>((const float *)in, (float *)out, B, n, (float *)ws, ws_bytes / 4, as_stream(stream)),
                        fused_iadrt<double>((const double *)in, (double *)out, B, n, (double *)ws, ws_bytes / 8, as_stream(stream)));
    return DISPATCH(dtype,
                    iadrt_by_stages<float>((const float *)in, (float *)out, B, n, (float *)ws, as_stream(stream)),
                    iadrt_by_stages<double>((const double *)in, (double *)out, B, n, (double *)ws, as_stream(stream)));
}

int adrt_b200_fmg_restriction(const void *in, void *out, int64_t B, int64_t n, int dtype, void *stream)
{
    ADRT_REQUIRE(in && out && in != out && dtype_ok(dtype) && B > 0, "bad argument");
    ADRT_REQUIRE(n >= 2 && n % 2 == 0 && n <= kMaxN, "restriction needs even n >= 2, got %lld", (long long)n);
    return DISPATCH(dtype,
                    launch_fmg_restriction<float>((const float *)in, (float *)out, B, n, as_stream(stream)),
                    launch_fmg_restriction<double>((const double *)in, (double *)out, B, n, as_stream(stream)));
}

int adrt_b200_fmg_prolongation(const void *in, void *out, int64_t B, int64_t h, int64_t w, int dtype, void *stream)
{
    ADRT_REQUIRE(in && out && in != out && dtype_ok(dtype) && B > 0 && h > 0 && w > 0, "bad argument");
    return DISPATCH(dtype,
                    launch_fmg_prolongation<float>((const float *)in, (float *)out, B, h, w, as_stream(stream)),
                    launch_fmg_prolongation<double>((const double *)in, (double *)out, B, h, w, as_stream(stream)));
}

int adrt_b200_fmg_highpass(const void *in, void *out, int64_t B, int64_t h, int64_t w, int dtype, void *stream)
{
    ADRT_REQUIRE(in && out && in != out && dtype_ok(dtype) && B > 0, "bad argument");
    ADRT_REQUIRE(h >= 2 && w >= 2, "array is too small to high-pass filter");
    return DISPATCH(dtype,
                    launch_fmg_highpass<float>((const float *)in, (float *)out, B, h, w, as_stream(stream)),
                    launch_fmg_highpass<double>((const double *)in, (double *)out, B, h, w, as_stream(stream)));
}

size_t adrt_b200_fmg_step_workspace_bytes(int64_t B, int64_t n, int dtype)
{
    if (B <= 0 || !is_pow2(n) || n > kMaxN || !dtype_ok(dtype)) return 0;
    return DISPATCH(dtype, fmg_step_ws_elems<float>(B, n) * 4, fmg_step_ws_elems<double>(B, n) * 8);
}

int adrt_b200_fmg_step(const void *in, void *out, int64_t B, int64_t n, int dtype, void *ws, size_t ws_bytes,
                       void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    if ((rc = check_vec_aligned(in, out, ws))) return rc;
    return DISPATCH(dtype,
                    fmg_step_impl<float>((const float *)in, (float *)out, B, n, (float *)ws, ws_bytes, as_stream(stream)),
                    fmg_step_impl<double>((const double *)in, (double *)out, B, n, (double *)ws, ws_bytes, as_stream(stream)));
}

int adrt_b200_interp_to_cart(const void *in, void *out, int64_t B, int64_t n, int dtype, void *stream)
{
    int rc = check_image(in, out, B, n, dtype);
    if (rc) return rc;
    ADRT_REQUIRE(n >= 2, "interp_to_cart needs n >= 2");
    int device = 0;
    ADRT_CUDA_CHECK(cudaGetDevice(&device));
    InterpTable tab;
    rc = get_interp_table(device, n, dtype, as_stream(stream), &tab);
    if (rc) return rc;
    return DISPATCH(dtype,
                    launch_interp_to_cart<float>((const float *)in, (float *)out, tab.t, tab.base, tab.h_base, tab.cosv,
                                                 tab.sgn, (const float *)tab.factor, B, n, as_stream(stream)),
                    launch_interp_to_cart<double>((const double *)in, (double *)out, tab.t, tab.base, tab.h_base, tab.cosv,
                                                  tab.sgn, (const double *)tab.factor, B, n, as_stream(stream)));
}

int adrt_b200_truncate(const void *in, void *out, int64_t B, int64_t n, int dtype, void *stream)
{
    ADRT_REQUIRE(in && out && in != out && dtype_ok(dtype) && B > 0 && is_pow2(n) && n <= kMaxN, "bad argument");
    return DISPATCH(dtype,
                    launch_truncate<float>((const float *)in, (float *)out, B, n, as_stream(stream)),
                    launch_truncate<double>((const double *)in, (double *)out, B, n, as_stream(stream)));
}

int adrt_b200_stitch(const void *in, void *out, int64_t B, int64_t n, int remove_repeated, int dtype, void *stream)
{
    ADRT_REQUIRE(in && out && in != out && dtype_ok(dtype) && B > 0 && n >= 1 && n <= kMaxN, "bad argument");
    return DISPATCH(dtype,
                    launch_stitch<float>((const float *)in, (float *)out, B, n, remove_repeated != 0, as_stream(stream)),
                    launch_stitch<double>((const double *)in, (double *)out, B, n, remove_repeated != 0, as_stream(stream)));
}

int adrt_b200_unstitch(const void *in, void *out, int64_t B, int64_t n, int trimmed, int dtype, void *stream)
{
    ADRT_REQUIRE(in && out && in != out && dtype_ok(dtype) && B > 0 && n >= 1 && n <= kMaxN, "bad argument");
    ADRT_REQUIRE(!(trimmed && n < 2), "the narrow stitched form needs n >= 2");
    return DISPATCH(dtype,
                    launch_unstitch<float>((const float *)in, (float *)out, B, n, trimmed != 0, as_stream(stream)),
                    launch_unstitch<double>((const double *)in, (double *)out, B, n, trimmed != 0, as_stream(stream)));
}

int adrt_b200_truncate_mean(const void *in, void *out, int64_t B, int64_t n, double divisor, int dtype, void *stream)
{
    ADRT_REQUIRE(in && out && in != out && dtype_ok(dtype) && B > 0 && is_pow2(n) && n <= kMaxN, "bad argument");
    return DISPATCH(dtype,
                    launch_truncate_mean<float>((const float *)in, (float *)out, B, n, (float)divisor, as_stream(stream)),
                    launch_truncate_mean<double>((const double *)in, (double *)out, B, n, divisor, as_stream(stream)));
}

int adrt_b200_truncate_mean_shares(const void *in, void *out, int64_t B, int64_t n, int per, int parts, double divisor, int dtype,
                                   void *stream)
{
    ADRT_REQUIRE(in && out && in != out && dtype_ok(dtype) && B > 0 && is_pow2(n) && n <= kMaxN, "bad argument");
    ADRT_REQUIRE((per == 1 || per == 2 || per == 4) && parts >= 1 && (parts & (parts - 1)) == 0 && n / parts >= 32,
                 "bad share layout: per=%d parts=%d n=%lld (needs n / parts >= 32)", per, parts, (long long)n);
    return DISPATCH(dtype,
                    launch_truncate_mean_shares<float>((const float *)in, (float *)out, B, n, per, parts, (float)divisor, as_stream(stream)),
                    launch_truncate_mean_shares<double>((const double *)in, (double *)out, B, n, per, parts, divisor, as_stream(stream)));
}

int adrt_b200_sub(const void *a, const void *b, void *out, int64_t count, int dtype, void *stream)
{
    ADRT_REQUIRE(a && b && out && dtype_ok(dtype) && count > 0, "bad argument");
    return DISPATCH(dtype,
                    launch_binary<float>((const float *)a, (const float *)b, (float *)out, count, 0, as_stream(stream)),
                    launch_binary<double>((const double *)a, (const double *)b, (double *)out, count, 0, as_stream(stream)));
}

int adrt_b200_add(const void *a, const void *b, void *out, int64_t count, int dtype, void *stream)
{
    ADRT_REQUIRE(a && b && out && dtype_ok(dtype) && count > 0, "bad argument");
    return DISPATCH(dtype,
                    launch_binary<float>((const float *)a, (const float *)b, (float *)out, count, 1, as_stream(stream)),
                    launch_binary<double>((const double *)a, (const double *)b, (double *)out, count, 1, as_stream(stream)));
}

size_t adrt_b200_cg_workspace_bytes(void) { return cg_workspace_bytes(); }

int adrt_b200_cg_dot(const void *a, const void *b, void *state, int slot, int64_t count, int dtype, void *ws, size_t ws_bytes,
                     void *stream)
{
    ADRT_REQUIRE(a && b && state && ws && dtype_ok(dtype) && count > 0 && slot >= 0 && slot < 4, "bad argument");
    ADRT_REQUIRE(ws_bytes >= cg_workspace_bytes(), "cg workspace too small: need %zu bytes, got %zu", cg_workspace_bytes(), ws_bytes);
    return DISPATCH(dtype,
                    launch_cg_dot<float>((const float *)a, (const float *)b, (double *)state, slot, count, (double *)ws, as_stream(stream)),
                    launch_cg_dot<double>((const double *)a, (const double *)b, (double *)state, slot, count, (double *)ws, as_stream(stream)));
}

int adrt_b200_cg_update(void *x, void *r, const void *p, const void *ap, void *state, int64_t count, int dtype, void *ws,
                        size_t ws_bytes, void *stream)
{
    ADRT_REQUIRE(x && r && p && ap && state && ws && dtype_ok(dtype) && count > 0, "bad argument");
    ADRT_REQUIRE(x != r && x != p && x != ap && r != p && r != ap, "cg_update: the vectors must be distinct");
    ADRT_REQUIRE(ws_bytes >= cg_workspace_bytes(), "cg workspace too small: need %zu bytes, got %zu", cg_workspace_bytes(), ws_bytes);
    return DISPATCH(dtype,
                    launch_cg_update<float>((float *)x, (float *)r, (const float *)p, (const float *)ap, (double *)state, count, (double *)ws, as_stream(stream)),
                    launch_cg_update<double>((double *)x, (double *)r, (const double *)p, (const double *)ap, (double *)state, count, (double *)ws, as_stream(stream)));
}

int adrt_b200_cg_direction(void *p, const void *r, void *state, int64_t count, int dtype, void *stream)
{
    ADRT_REQUIRE(p && r && state && p != r && dtype_ok(dtype) && count > 0, "bad argument");
    return DISPATCH(dtype,
                    launch_cg_direction<float>((float *)p, (const float *)r, (double *)state, count, as_stream(stream)),
                    launch_cg_direction<double>((double *)p, (const double *)r, (double *)state, count, as_stream(stream)));
}

}  // extern "C"
