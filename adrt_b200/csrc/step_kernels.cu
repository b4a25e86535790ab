// One-kernel-per-stage operators in the public Q-layout (B,4,D,n), D = 2n-1.
//
// These implement the reference's single-step API (adrt_step / bdrt_step),
// adrt_init, the exact inverse iadrt, the Press FMG operators, the
// interp_to_cart gather and the elementwise glue of the multigrid driver.
// They are simple streaming kernels (every element read/written once per
// stage, c is the contiguous, coalesced axis).  The fast full transforms
// live in fused_adrt.cu.
//
// Bit-exactness rules (SURVEY.md section 8a "exactness"): exactly one IEEE add
// of the two named operands per output, a copy where the reference copies.
// A copy is expressed as `x + (-0.0)`, which is the identity for every x
// including -0.0 (x + (+0.0) would turn -0.0 into +0.0).  nvcc never contracts
// a lone add, and -fmad=false is set for the whole library anyway.
#include "common.cuh"

#include <cstdlib>

namespace adrt_b200 {

namespace {

constexpr int kThreads = 256;
constexpr int kRowThreadsS = 128;   // stitch / unstitch: threads along a row

__host__ __device__ inline int ilog2(int64_t n)
{
    int k = 0;
    while ((int64_t(1) << k) < n) ++k;
    return k;
}

// grid: x over in-plane chunks, y over planes (looped when > 65535)
inline dim3 plane_grid(int64_t plane_elems, int64_t planes)
{
    int64_t gx = (plane_elems + kThreads - 1) / kThreads;
    int64_t gy = planes < 65535 ? planes : 65535;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

// ---- adrt_init: core.py:169-176 / adrt_cdefs_adrt.hpp:124-186 -----------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
adrt_init_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t planes, int n, int logn)
{
    const int64_t D = 2 * (int64_t)n - 1;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= D * n) return;
    const int d = (int)(idx >> logn), c = (int)(idx & (n - 1));
    for (int64_t p = blockIdx.y; p < planes; p += gridDim.y) {
        const int q = (int)(p & 3);
        const T *img = in + (p >> 2) * (int64_t)n * n;
        T v = T(0);
        if (d < n) {
            int r, k;
            switch (q) {
            case 0: r = c; k = n - 1 - d; break;
            case 1: r = n - 1 - d; k = c; break;
            case 2: r = d; k = c; break;
            default: r = n - 1 - c; k = n - 1 - d; break;
            }
            v = img[(int64_t)r * n + k];
        }
        out[p * D * n + idx] = v;
    }
}

// kVec consecutive outputs leave a thread as one aligned vector store
template <typename T, int kVec> struct alignas(sizeof(T) * kVec) OutVec { T v[kVec]; };

template <typename T, int kVec>
__device__ __forceinline__ void store_outvec(T *o, const T (&v)[kVec])
{
    OutVec<T, kVec> t;
#pragma unroll
    for (int i = 0; i < kVec; ++i) t.v[i] = v[i];
    *reinterpret_cast<OutVec<T, kVec> *>(o) = t;
}

// ---- adrt_step: adrt_cdefs_adrt.hpp:215-258 -----------------------------------
template <typename T>
__device__ __forceinline__ T adrt_step_one(const T *__restrict__ I, int64_t d, int c, int n, int iter)
{
    const int e = 1 << iter;
    const int a = c & (2 * e - 1);
    const int cA = (c - a) + (a >> 1);
    const int sh = (a + 1) >> 1;
    const T av = I[d * n + cA];
    const T bv = (d >= sh) ? I[(d - sh) * n + cA + e] : T(-0.0);
    return av + bv;
}

// kVec = 4 consecutive columns per thread (one 16/32-byte store), n % 4 == 0
template <typename T, int kVec>
__global__ void __launch_bounds__(kThreads)
adrt_step_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t planes, int n, int logn, int iter)
{
    const int64_t D = 2 * (int64_t)n - 1;
    const int64_t idx = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kVec;
    if (idx >= D * n) return;
    const int64_t d = idx >> logn;
    const int c = (int)(idx & (n - 1));
    for (int64_t p = blockIdx.y; p < planes; p += gridDim.y) {
        const T *I = in + p * D * n;
        T v[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) v[i] = adrt_step_one<T>(I, d, c + i, n, iter);
        store_outvec<T, kVec>(out + p * D * n + idx, v);
    }
}

// ---- bdrt_step: adrt_cdefs_bdrt.hpp:190-244 (step semantics) and --------------
// ---- bdrt_core: adrt_cdefs_bdrt.hpp:55-116 (core semantics) -------------------
// step semantics: missing operands are +0 and are still added (aval=0; bval=0).
// core semantics: last valid row is a copy of la_val, rows past it are +0.
template <typename T, bool kCore>
__device__ __forceinline__ T bdrt_step_one(const T *__restrict__ I, int64_t d, int c, int n, int64_t D, int adrt_iter)
{
    const int e = 1 << adrt_iter;
    const int cb = c >> adrt_iter, ci = c & (e - 1);
    const int beta = 2 * (ci + e * (cb >> 1));
    if (!(cb & 1)) return I[d * n + beta] + I[d * n + beta + 1];
    const int64_t r = d + ci;
    const T av = (r < D) ? I[r * n + beta] : T(0.0);
    const T bv = (r + 1 < D) ? I[(r + 1) * n + beta + 1] : (kCore ? T(-0.0) : T(0.0));
    return av + bv;
}

template <typename T, bool kCore, int kVec>
__global__ void __launch_bounds__(kThreads)
bdrt_step_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t planes, int n, int logn, int adrt_iter)
{
    const int64_t D = 2 * (int64_t)n - 1;
    const int64_t idx = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kVec;
    if (idx >= D * n) return;
    const int64_t d = idx >> logn;
    const int c = (int)(idx & (n - 1));
    for (int64_t p = blockIdx.y; p < planes; p += gridDim.y) {
        const T *I = in + p * D * n;
        T v[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) v[i] = bdrt_step_one<T, kCore>(I, d, c + i, n, D, adrt_iter);
        store_outvec<T, kVec>(out + p * D * n + idx, v);
    }
}

// Same outputs, CTA = 32 offsets x 32 columns (8 threads of 4 columns per offset): the operands of
// an odd block sit on a diagonal of the input, in[d + ci][2 ci ..], so the 32-byte sector a thread
// touches is shared with the 3 offsets below it -- a 2-D CTA finds them in L1 instead of L2
// (a CTA spanning one offset row re-reads every sector 4 times).  n % 32 == 0.
template <typename T, bool kCore>
__global__ void __launch_bounds__(256)
bdrt_step_tiled_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t planes, int n, int logn, int adrt_iter)
{
    const int64_t D = 2 * (int64_t)n - 1;
    const int ctiles = n >> 5;
    const int ct = blockIdx.x & (ctiles - 1), dt = blockIdx.x >> (logn - 5);
    const int c = ct * 32 + (threadIdx.x & 7) * 4;
    const int64_t d = (int64_t)dt * 32 + (threadIdx.x >> 3);
    if (d >= D) return;
    for (int64_t p = blockIdx.y; p < planes; p += gridDim.y) {
        const T *I = in + p * D * n;
        T v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = bdrt_step_one<T, kCore>(I, d, c + i, n, D, adrt_iter);
        store_outvec<T, 4>(out + p * D * n + d * n + c, v);
    }
}

// ---- iadrt stage: adrt_cdefs_iadrt.hpp:52-105 in the Q-layout -----------------
// One thread per (plane, output column); the offset axis is walked serially
// from D-1 down (the reference's "must be serial" loop, iadrt.hpp:73-98) with
// the running value kept in a register.  Adjacent threads own adjacent
// columns, so every load and store is coalesced.  Only the running sum is
// serial: the operands of kIadrtBatch rows are fetched together before the
// chain of adds consumes them, which keeps enough loads in flight to hide the
// DRAM latency (one row at a time is latency bound at a quarter of the bandwidth).

template <typename T, int kIadrtBatch>
__global__ void __launch_bounds__(128)
iadrt_stage_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t planes, int n, int stage)
{
    // in-plane offsets fit 32 bits (n <= 16384: D * n < 2^30)
    const int D = 2 * n - 1;
    const int co = blockIdx.x * 128 + threadIdx.x;
    if (co >= n) return;
    const int Cin = n >> stage, C = Cin >> 1;
    const int l = co / C, col = co - l * C;
    const int A = (l >> 1) * Cin + 2 * col;
    const bool even = (l & 1) == 0;
    // even l: a = in[d][A], b = in[d+1][A+1]; odd l: a = in[d+1+col][A+1], b = in[d+1+col][A]
    const int ra_off = even ? 0 : 1 + col, rb_off = 1 + (even ? 0 : col);
    const int ca = even ? A : A + 1, cb = even ? A + 1 : A;
    for (int64_t p = blockIdx.y; p < planes; p += gridDim.y) {
        const T *Ia = in + p * (int64_t)D * n + ca;
        const T *Ib = in + p * (int64_t)D * n + cb;
        T *O = out + p * (int64_t)D * n + co;
        T prev = T(0);
        for (int dtop = D - 1; dtop >= 0; dtop -= kIadrtBatch) {
            T a[kIadrtBatch], b[kIadrtBatch];
#pragma unroll
            for (int u = 0; u < kIadrtBatch; ++u) {
                const int d = dtop - u;
                a[u] = T(0);
                b[u] = T(0);
                if (d >= 0) {
                    // even: a always exists, b needs d+1 < D; odd: both need d+1+col < D
                    if (d + ra_off < D) a[u] = Ia[(d + ra_off) * n];
                    if (d + rb_off < D) b[u] = Ib[(d + rb_off) * n];
                }
            }
#pragma unroll
            for (int u = 0; u < kIadrtBatch; ++u) {
                const int d = dtop - u;
                if (d >= 0) {
                    T val = T(0);
                    if (even) {
                        val += a[u];
                        if (d + 1 < D) val -= b[u];
                    } else if (d + 1 + col < D) {
                        val += a[u];
                        val -= b[u];
                    }
                    if (d + 1 < D) val += prev;
                    O[d * n] = val;
                    prev = val;
                }
            }
        }
    }
}

// ---- Press FMG operators: adrt_cdefs_fmg.hpp ----------------------------------
// Grids: x over (vectors of) columns, y over rows, z over planes (y and z looped when
// they exceed the grid limits), so no thread ever divides.
constexpr int kRowThreads = 128;

template <typename T, int kVec> struct alignas(sizeof(T) * kVec) InVec { T v[kVec]; };

inline dim3 row_grid(int64_t col_items, int64_t rows, int64_t planes)
{
    return dim3((unsigned)((col_items + kRowThreads - 1) / kRowThreads), (unsigned)(rows < 65535 ? rows : 65535),
                (unsigned)(planes < 65535 ? planes : 65535));
}

// out[r,c] = (in[2r,2c] + in[2r+1,2c]) / 4  (fmg.hpp:53-73).  kVec = 2: two output
// columns per thread from two 4-element row vectors (n % 4 == 0).
template <typename T, int kVec>
__global__ void __launch_bounds__(kRowThreads)
fmg_restriction_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t planes, int n)
{
    const int64_t D = 2 * (int64_t)n - 1, R = n - 1, C = n / 2;
    const int c = (blockIdx.x * kRowThreads + threadIdx.x) * kVec;
    if (c >= C) return;
    for (int64_t p = blockIdx.z; p < planes; p += gridDim.z)
        for (int64_t r = blockIdx.y; r < R; r += gridDim.y) {
            const T *I = in + p * D * n + (2 * r) * n + 2 * c;
            T *O = out + p * R * C + r * C + c;
            if (kVec == 2) {
                const InVec<T, 4> va = *reinterpret_cast<const InVec<T, 4> *>(I);
                const InVec<T, 4> vb = *reinterpret_cast<const InVec<T, 4> *>(I + n);
                const T v[2] = {(va.v[0] + vb.v[0]) / T(4), (va.v[2] + vb.v[2]) / T(4)};
                store_outvec<T, 2>(O, v);
            } else {
                O[0] = (I[0] + I[n]) / T(4);
            }
        }
}

// out[2r+{0,1}, 2c+{0,1}] = in[r,c]  (fmg.hpp:75-95).  kVec = 2: two input columns per
// thread, written as one 4-element vector to each of the two output rows (w even).
template <typename T, int kVec>
__global__ void __launch_bounds__(kRowThreads)
fmg_prolongation_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t B, int64_t h, int64_t w)
{
    const int64_t W2 = 2 * w;
    const int64_t c = ((int64_t)blockIdx.x * kRowThreads + threadIdx.x) * kVec;
    if (c >= w) return;
    for (int64_t b = blockIdx.z; b < B; b += gridDim.z)
        for (int64_t r = blockIdx.y; r < h; r += gridDim.y) {
            const T *I = in + (b * h + r) * w + c;
            T *O = out + (b * 2 * h + 2 * r) * W2 + 2 * c;
            if (kVec == 2) {
                const InVec<T, 2> x = *reinterpret_cast<const InVec<T, 2> *>(I);
                const T v[4] = {x.v[0], x.v[0], x.v[1], x.v[1]};
                store_outvec<T, 4>(O, v);
                store_outvec<T, 4>(O + W2, v);
            } else {
                const T x = I[0];
                O[0] = x; O[1] = x; O[W2] = x; O[W2 + 1] = x;
            }
        }
}

// Four consecutive columns per thread, walking down a strip of kStrip rows with the three
// live rows in registers: one vector + two scalar loads per row instead of nine per output.
// w % 4 == 0 and 16-byte aligned planes.
constexpr int kStrip = 8;

template <typename T> struct HpRow { T l, v[4], r; };

template <typename T>
__device__ __forceinline__ HpRow<T> hp_load(const T *__restrict__ row, int c, int w)
{
    HpRow<T> o;
    const InVec<T, 4> x = *reinterpret_cast<const InVec<T, 4> *>(row + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) o.v[i] = x.v[i];
    o.l = c == 0 ? x.v[1] : row[c - 1];             // reflect-101 (fmg.hpp:108-118)
    o.r = c + 4 == w ? x.v[2] : row[c + 4];
    return o;
}

template <typename T>
__global__ void __launch_bounds__(kRowThreads)
fmg_highpass_vec_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t B, int h, int w)
{
    const T ca = T(-0.0625), cb = T(-0.125), cc = T(0.75);
    const int c = (blockIdx.x * kRowThreads + threadIdx.x) * 4;
    if (c >= w) return;
    for (int64_t b = blockIdx.z; b < B; b += gridDim.z)
        for (int r0 = blockIdx.y * kStrip; r0 < h; r0 += gridDim.y * kStrip) {
            const T *I = in + b * h * w;
            T *O = out + b * h * w;
            HpRow<T> P = hp_load(I + (int64_t)(r0 == 0 ? 1 : r0 - 1) * w, c, w);
            HpRow<T> C = hp_load(I + (int64_t)r0 * w, c, w);
#pragma unroll
            for (int i = 0; i < kStrip; ++i) {
                const int r = r0 + i;
                if (r >= h) break;
                const HpRow<T> N = (r == h - 1) ? P : hp_load(I + (int64_t)(r + 1) * w, c, w);
                const T p[6] = {P.l, P.v[0], P.v[1], P.v[2], P.v[3], P.r};
                const T m[6] = {C.l, C.v[0], C.v[1], C.v[2], C.v[3], C.r};
                const T q[6] = {N.l, N.v[0], N.v[1], N.v[2], N.v[3], N.r};
                T v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // products rounded one by one, summed column by column (fmg.hpp:129,150,167)
                    const T v11 = ca * p[k], v12 = cb * p[k + 1], v13 = ca * p[k + 2];
                    const T v21 = cb * m[k], v22 = cc * m[k + 1], v23 = cb * m[k + 2];
                    const T v31 = ca * q[k], v32 = cb * q[k + 1], v33 = ca * q[k + 2];
                    v[k] = ((v11 + v21) + v31) + ((v12 + v22) + v32) + ((v13 + v23) + v33);
                }
                store_outvec<T, 4>(O + (int64_t)r * w + c, v);
                P = C;
                C = N;
            }
        }
}

template <typename T>
__global__ void __launch_bounds__(kRowThreads)
fmg_highpass_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t B, int64_t h, int64_t w)
{
    const T ca = T(-0.0625), cb = T(-0.125), cc = T(0.75);
    const int64_t c = (int64_t)blockIdx.x * kRowThreads + threadIdx.x;
    if (c >= w) return;
    const int64_t pc = (c == 0 ? 1 : c - 1), nc = (c == w - 1 ? c - 1 : c + 1);
    for (int64_t b = blockIdx.z; b < B; b += gridDim.z)
        for (int64_t r = blockIdx.y; r < h; r += gridDim.y) {
            const int64_t pr = (r == 0 ? 1 : r - 1), nr = (r == h - 1 ? r - 1 : r + 1);
            const T *I = in + b * h * w;
            // every product is rounded on its own (no FMA), then summed column by
            // column exactly like fmg.hpp:129,150,167
            const T v11 = ca * I[pr * w + pc], v12 = cb * I[pr * w + c], v13 = ca * I[pr * w + nc];
            const T v21 = cb * I[r * w + pc], v22 = cc * I[r * w + c], v23 = cb * I[r * w + nc];
            const T v31 = ca * I[nr * w + pc], v32 = cb * I[nr * w + c], v33 = ca * I[nr * w + nc];
            out[(b * h + r) * w + c] = ((v11 + v21) + v31) + ((v12 + v22) + v32) + ((v13 + v23) + v33);
        }
}

// ---- interp_to_cart: adrt_cdefs_interp_adrtcart.hpp:61-114 --------------------
// The per-column transcendental pieces come from the host tables built in
// api.cu (get_interp_table); this kernel finishes the index computation with
// IEEE-exact float ops in the reference's order (interp.hpp:97-100):
//   h0 = (0.5 + tan/2) + ((sgn >= 0 ? t : -t) / cos(th0))
//   hi = (round(h0 * 2n) - 1) / 2
// and gathers factor * in[q, floor(hi), si] (0 when hi is out of range).
template <typename T>
__global__ void __launch_bounds__(kThreads)
interp_kernel(const T *__restrict__ in, T *__restrict__ out, const float *__restrict__ tt,
              const int32_t *__restrict__ base, const float *__restrict__ h_base,
              const float *__restrict__ cosv, const int32_t *__restrict__ sgn,
              const T *__restrict__ factor, int64_t B, int n, int logn)
{
    const int64_t D = 2 * (int64_t)n - 1, N = (int64_t)n * 4 * n;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= N) return;
    const int off = (int)(idx >> (logn + 2)), ang = (int)(idx & (4 * n - 1));
    const float t = tt[off];
    const float ts = sgn[ang] ? t : -t;
    const float h0 = __fadd_rn(h_base[ang], __fdiv_rn(ts, cosv[ang]));
    const float hi = __fdiv_rn(__fadd_rn(roundf(__fmul_rn(h0, (float)(2 * n))), -1.0f), 2.0f);
    const bool ok = hi >= 0.0f && hi < (float)D;
    const int64_t src = ok ? (int64_t)base[ang] + (int64_t)hi * n : 0;
    const T f = factor[ang];
    // the index computation above is shared by all images: each thread serves B / gridDim.y of them
#pragma unroll 4
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        T v = T(0);
        if (ok) v = f * in[b * 4 * D * n + src];
        out[b * N + idx] = v;
    }
}

// ---- truncate (utils.py:231-242) and truncate+divide+mean (core.py:329) -------
template <typename T>
__device__ inline T truncate_load(const T *plane0, int q, int r, int c, int n, int64_t D)
{
    // T0 = flip_rows(a0[:n,:n])^T  -> T0[r,c] = a0[n-1-c, r]
    // T1 = flip_rows(a1[:n,:n])    -> T1[r,c] = a1[n-1-r, c]
    // T2 = a2[:n,:n]
    // T3 = flip_both(a3[:n,:n])^T  -> T3[r,c] = a3[n-1-c, n-1-r]
    int d, k;
    switch (q) {
    case 0: d = n - 1 - c; k = r; break;
    case 1: d = n - 1 - r; k = c; break;
    case 2: d = r; k = c; break;
    default: d = n - 1 - c; k = n - 1 - r; break;
    }
    return plane0[((int64_t)q * D + d) * n + k];
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
truncate_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t planes, int n, int logn)
{
    const int64_t D = 2 * (int64_t)n - 1;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (int64_t)n * n) return;
    const int r = (int)(idx >> logn), c = (int)(idx & (n - 1));
    for (int64_t p = blockIdx.y; p < planes; p += gridDim.y) {
        const int q = (int)(p & 3);
        out[p * n * n + idx] = truncate_load(in + (p >> 2) * 4 * D * n, q, r, c, n, D);
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
truncate_mean_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t B, int n, int logn, T divisor)
{
    const int64_t D = 2 * (int64_t)n - 1;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (int64_t)n * n) return;
    const int r = (int)(idx >> logn), c = (int)(idx & (n - 1));
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const T *I = in + b * 4 * D * n;
        const T t0 = truncate_load(I, 0, r, c, n, D) / divisor;
        const T t1 = truncate_load(I, 1, r, c, n, D) / divisor;
        const T t2 = truncate_load(I, 2, r, c, n, D) / divisor;
        const T t3 = truncate_load(I, 3, r, c, n, D) / divisor;
        out[b * n * n + idx] = (((t0 + t1) + t2) + t3) / T(4);
    }
}

// n >= 32: 32 x 32 output tiles; the two transposed quadrants (0 and 3) go through
// shared memory so that every global access runs along the contiguous axis.
// kShares: the input is not a (B, 4, 2n-1, n) sinogram but the all-gathered shares of a sharded
// back-projection (adrt_b200/_shard.py): rank r = group * parts + p contributes (B, per, n, w) --
// offsets d < n, columns [p*w, (p+1)*w), w = n / parts, of the group's `per` quadrants -- and the
// shares lie rank-major: element (b, q, d, k) sits at ((r*B + b)*per + q % per)*n*w + d*w + k % w with
// r = (q / per)*parts + k / w.
struct ShareLayout {
    int per, parts, logw;
};

template <typename T, bool kShares>
__device__ __forceinline__ int64_t tm_addr(int64_t b, int64_t B, int q, int64_t d, int k, int n, int64_t D, const ShareLayout &L)
{
    if (!kShares) return ((b * 4 + q) * D + d) * n + k;
    const int w = 1 << L.logw;
    const int64_t r = (int64_t)(q / L.per) * L.parts + (k >> L.logw);
    return (((r * B + b) * L.per + q % L.per) * n + d) * w + (k & (w - 1));
}

template <typename T, bool kShares>
__global__ void __launch_bounds__(256)
truncate_mean_tiled_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t B, int n, T divisor, ShareLayout L)
{
    __shared__ T s0[32][33], s3[32][33];
    const int64_t D = 2 * (int64_t)n - 1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int64_t b = blockIdx.z; b < B; b += gridDim.z) {
        // s0[i][j] = a0[n-1-(c0+i), r0+j] = T0[r0+j, c0+i];  s3[i][j] = a3[n-1-(c0+i), n-1-(r0+j)] = T3[r0+j, c0+i]
        // all sixteen loads of a thread are issued before anything waits on them
        T a0[4], a3[4], a1[4], a2[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int i = ty + 8 * m;
            const int64_t d = n - 1 - (c0 + i);
            a0[m] = in[tm_addr<T, kShares>(b, B, 0, d, r0 + tx, n, D, L)];
            a3[m] = in[tm_addr<T, kShares>(b, B, 3, d, n - 1 - (r0 + tx), n, D, L)];
            const int r = r0 + i;
            a1[m] = in[tm_addr<T, kShares>(b, B, 1, n - 1 - r, c0 + tx, n, D, L)];
            a2[m] = in[tm_addr<T, kShares>(b, B, 2, r, c0 + tx, n, D, L)];
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            s0[ty + 8 * m][tx] = a0[m];
            s3[ty + 8 * m][tx] = a3[m];
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int j = ty + 8 * m;
            const T t0 = s0[tx][j] / divisor;
            const T t1 = a1[m] / divisor;
            const T t2 = a2[m] / divisor;
            const T t3 = s3[tx][j] / divisor;
            out[b * n * n + (int64_t)(r0 + j) * n + c0 + tx] = (((t0 + t1) + t2) + t3) / T(4);
        }
        __syncthreads();
    }
}

template <typename T, int kOp>
__global__ void __launch_bounds__(kThreads)
binary_kernel(const T *a, const T *b, T *out, int64_t count)  // out may alias a or b
{
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kThreads)
        out[i] = kOp == 0 ? a[i] - b[i] : a[i] + b[i];
}

// 16-byte vectors (count and the three pointers are multiples of the vector size)
template <typename T, int kOp>
__global__ void __launch_bounds__(kThreads)
binary_vec_kernel(const T *a, const T *b, T *out, int64_t vectors)
{
    constexpr int L = 16 / sizeof(T);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < vectors; i += (int64_t)gridDim.x * kThreads) {
        const InVec<T, L> x = reinterpret_cast<const InVec<T, L> *>(a)[i];
        const InVec<T, L> y = reinterpret_cast<const InVec<T, L> *>(b)[i];
        T v[L];
#pragma unroll
        for (int k = 0; k < L; ++k) v[k] = kOp == 0 ? x.v[k] - y.v[k] : x.v[k] + y.v[k];
        store_outvec<T, L>(out + i * L, v);
    }
}

}  // namespace

// ---- launchers ------------------------------------------------------------------
// vector paths need the base pointers aligned to the vector size
inline bool aligned_to(const void *p, size_t bytes) { return reinterpret_cast<uintptr_t>(p) % bytes == 0; }

template <typename T>
int launch_adrt_init(const T *in, T *out, int64_t B, int64_t n, cudaStream_t s)
{
    const int64_t D = 2 * n - 1;
    adrt_init_kernel<T><<<plane_grid(D * n, B * 4), kThreads, 0, s>>>(in, out, B * 4, (int)n, ilog2(n));
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_adrt_step(const T *in, T *out, int64_t B, int64_t n, int step, cudaStream_t s)
{
    const int64_t D = 2 * n - 1;
    if (n % 4 == 0 && aligned_to(out, 4 * sizeof(T)))
        adrt_step_kernel<T, 4><<<plane_grid(D * n / 4, B * 4), kThreads, 0, s>>>(in, out, B * 4, (int)n, ilog2(n), step);
    else
        adrt_step_kernel<T, 1><<<plane_grid(D * n, B * 4), kThreads, 0, s>>>(in, out, B * 4, (int)n, ilog2(n), step);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_bdrt_step(const T *in, T *out, int64_t B, int64_t n, int step, bool core_semantics, cudaStream_t s)
{
    const int64_t D = 2 * n - 1;
    const int adrt_iter = num_iters(n) - 1 - step;
    const bool vec = n % 4 == 0 && aligned_to(out, 4 * sizeof(T));
    if (vec && n % 32 == 0) {
        const dim3 tg = plane_grid(((D + 31) / 32) * (n / 32) * 256, B * 4);
        if (core_semantics) bdrt_step_tiled_kernel<T, true><<<tg, 256, 0, s>>>(in, out, B * 4, (int)n, ilog2(n), adrt_iter);
        else bdrt_step_tiled_kernel<T, false><<<tg, 256, 0, s>>>(in, out, B * 4, (int)n, ilog2(n), adrt_iter);
        ADRT_LAUNCH_CHECK();
        return ADRT_B200_OK;
    }
    const dim3 grid = plane_grid(vec ? D * n / 4 : D * n, B * 4);
    if (core_semantics && vec) bdrt_step_kernel<T, true, 4><<<grid, kThreads, 0, s>>>(in, out, B * 4, (int)n, ilog2(n), adrt_iter);
    else if (core_semantics) bdrt_step_kernel<T, true, 1><<<grid, kThreads, 0, s>>>(in, out, B * 4, (int)n, ilog2(n), adrt_iter);
    else if (vec) bdrt_step_kernel<T, false, 4><<<grid, kThreads, 0, s>>>(in, out, B * 4, (int)n, ilog2(n), adrt_iter);
    else bdrt_step_kernel<T, false, 1><<<grid, kThreads, 0, s>>>(in, out, B * 4, (int)n, ilog2(n), adrt_iter);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_iadrt_stage(const T *in, T *out, int64_t B, int64_t n, int stage, cudaStream_t s)
{
    const int64_t planes = B * 4;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)(planes < 65535 ? planes : 65535), 1);
    // rows fetched together per thread: measured best 16 (fp32) / 12 (fp64) at 16 x 2048^2 (profiles/r01_ops.jsonl)
    static const int batch_env = [] { const char *e = getenv("ADRT_B200_IADRT_BATCH"); return e ? atoi(e) : 0; }();
    const int batch = batch_env > 0 ? batch_env : (sizeof(T) == 4 ? 16 : 12);
    if (batch >= 16) iadrt_stage_kernel<T, 16><<<grid, 128, 0, s>>>(in, out, planes, (int)n, stage);
    else if (batch >= 12) iadrt_stage_kernel<T, 12><<<grid, 128, 0, s>>>(in, out, planes, (int)n, stage);
    else iadrt_stage_kernel<T, 8><<<grid, 128, 0, s>>>(in, out, planes, (int)n, stage);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_fmg_restriction(const T *in, T *out, int64_t B, int64_t n, cudaStream_t s)
{
    if (n % 4 == 0 && aligned_to(in, 4 * sizeof(T)) && aligned_to(out, 2 * sizeof(T))) fmg_restriction_kernel<T, 2><<<row_grid(n / 4, n - 1, B * 4), kRowThreads, 0, s>>>(in, out, B * 4, (int)n);
    else fmg_restriction_kernel<T, 1><<<row_grid(n / 2, n - 1, B * 4), kRowThreads, 0, s>>>(in, out, B * 4, (int)n);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_fmg_prolongation(const T *in, T *out, int64_t B, int64_t h, int64_t w, cudaStream_t s)
{
    if (w % 2 == 0 && aligned_to(in, 2 * sizeof(T)) && aligned_to(out, 4 * sizeof(T))) fmg_prolongation_kernel<T, 2><<<row_grid(w / 2, h, B), kRowThreads, 0, s>>>(in, out, B, h, w);
    else fmg_prolongation_kernel<T, 1><<<row_grid(w, h, B), kRowThreads, 0, s>>>(in, out, B, h, w);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_fmg_highpass(const T *in, T *out, int64_t B, int64_t h, int64_t w, cudaStream_t s)
{
    if (w % 4 == 0 && w >= 4 && h < (1 << 30) && w < (1 << 30) && aligned_to(in, 4 * sizeof(T)) && aligned_to(out, 4 * sizeof(T)))
        fmg_highpass_vec_kernel<T><<<row_grid(w / 4, (h + kStrip - 1) / kStrip, B), kRowThreads, 0, s>>>(in, out, B, (int)h, (int)w);
    else
        fmg_highpass_kernel<T><<<row_grid(w, h, B), kRowThreads, 0, s>>>(in, out, B, h, w);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_interp_to_cart(const T *in, T *out, const float *t, const int32_t *base, const float *h_base,
                          const float *cosv, const int32_t *sgn, const T *factor, int64_t B, int64_t n,
                          cudaStream_t s)
{
    dim3 grid = plane_grid(4 * n * n, B);
    // enough threads to fill the GPU with one or two images: share the index math over the batch
    if (4 * n * n >= (int64_t)148 * 2048 * 4) grid.y = (unsigned)((B + 7) / 8);
    interp_kernel<T><<<grid, kThreads, 0, s>>>(in, out, t, base, h_base, cosv, sgn, factor, B, (int)n, ilog2(n));
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

// ---- stitch_adrt / unstitch_adrt (utils.py:111-134, 162-188) ------------------------
// Stitched image: (3n-2) rows x 4 bands of w columns (w = n, or n-1 with remove_repeated); band i is
// quadrant i, flipped on both axes when i is odd; bands 0,1 sit at rows 0 .. 2n-2, bands 2,3 at rows
// n-1 .. 3n-3; everything else is zero.  Grid: x over the 4w columns, y over rows, z over images.
template <typename T>
__global__ void __launch_bounds__(kRowThreadsS)
stitch_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t B, int n, int w)
{
    const int D = 2 * n - 1, R = 3 * n - 2, W = 4 * w;
    const int x = blockIdx.x * kRowThreadsS + threadIdx.x;
    if (x >= W) return;
    const int band = x / w, c = x - band * w;
    const int off = band < 2 ? 0 : n - 1;
    for (int64_t b = blockIdx.z; b < B; b += gridDim.z)
        for (int r = blockIdx.y; r < R; r += gridDim.y) {
            const int rr = r - off;
            T v = T(0);
            if (rr >= 0 && rr < D) {
                const T *q = in + (b * 4 + band) * (int64_t)D * n;
                v = (band & 1) ? q[(int64_t)(D - 1 - rr) * n + (n - 1 - c)] : q[(int64_t)rr * n + c];
            }
            out[(b * R + r) * (int64_t)W + x] = v;
        }
}

// Inverse: out[b, q, d, c].  With the narrow (4n-4) form the dropped last column of a band is the
// first column of the next band (read upside down for the wrap from band 3 to band 0).
template <typename T>
__global__ void __launch_bounds__(kRowThreadsS)
unstitch_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t B, int n, int w)
{
    const int D = 2 * n - 1, R = 3 * n - 2, W = 4 * w;
    const int c = blockIdx.x * kRowThreadsS + threadIdx.x;
    if (c >= n) return;
    for (int64_t p = blockIdx.z; p < B * 4; p += gridDim.z) {
        const int q = (int)(p & 3);
        const T *I = in + (p >> 2) * (int64_t)R * W;
        for (int d = blockIdx.y; d < D; d += gridDim.y) {
            // position in the band before the odd quadrants' flip
            const int dd = (q & 1) ? D - 1 - d : d, cc = (q & 1) ? n - 1 - c : c;
            const int row = dd + (q < 2 ? 0 : n - 1);
            T v;
            if (cc < w) v = I[(int64_t)row * W + q * w + cc];
            else if (q == 3) v = I[(int64_t)(R - 1 - row) * W];
            else v = I[(int64_t)row * W + (q + 1) * w];
            out[(p * D + d) * (int64_t)n + c] = v;
        }
    }
}

template <typename T>
int launch_stitch(const T *in, T *out, int64_t B, int64_t n, bool remove_repeated, cudaStream_t s)
{
    const int w = (int)n - (remove_repeated ? 1 : 0);
    if (w <= 0) return ADRT_B200_OK;   // n = 1 with remove_repeated: zero-width result
    const int64_t R = 3 * n - 2;
    dim3 grid((unsigned)((4 * w + kRowThreadsS - 1) / kRowThreadsS), (unsigned)(R < 65535 ? R : 65535), (unsigned)(B < 65535 ? B : 65535));
    stitch_kernel<T><<<grid, kRowThreadsS, 0, s>>>(in, out, B, (int)n, w);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_unstitch(const T *in, T *out, int64_t B, int64_t n, bool trimmed, cudaStream_t s)
{
    const int w = (int)n - (trimmed ? 1 : 0);
    const int64_t D = 2 * n - 1;
    dim3 grid((unsigned)((n + kRowThreadsS - 1) / kRowThreadsS), (unsigned)(D < 65535 ? D : 65535), (unsigned)(B * 4 < 65535 ? B * 4 : 65535));
    unstitch_kernel<T><<<grid, kRowThreadsS, 0, s>>>(in, out, B, (int)n, w);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_truncate(const T *in, T *out, int64_t B, int64_t n, cudaStream_t s)
{
    truncate_kernel<T><<<plane_grid(n * n, B * 4), kThreads, 0, s>>>(in, out, B * 4, (int)n, ilog2(n));
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_truncate_mean(const T *in, T *out, int64_t B, int64_t n, T divisor, cudaStream_t s)
{
    if (n >= 32) {
        dim3 grid((unsigned)(n / 32), (unsigned)(n / 32), (unsigned)(B < 65535 ? B : 65535));
        truncate_mean_tiled_kernel<T, false><<<grid, 256, 0, s>>>(in, out, B, (int)n, divisor, ShareLayout{4, 1, 0});
    } else {
        truncate_mean_kernel<T><<<plane_grid(n * n, B), kThreads, 0, s>>>(in, out, B, (int)n, ilog2(n), divisor);
    }
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

// quadrant mean straight from the all-gathered shares of a sharded back-projection (see ShareLayout)
template <typename T>
int launch_truncate_mean_shares(const T *in, T *out, int64_t B, int64_t n, int per, int parts, T divisor, cudaStream_t s)
{
    dim3 grid((unsigned)(n / 32), (unsigned)(n / 32), (unsigned)(B < 65535 ? B : 65535));
    truncate_mean_tiled_kernel<T, true><<<grid, 256, 0, s>>>(in, out, B, (int)n, divisor, ShareLayout{per, parts, ilog2(n / parts)});
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_binary(const T *a, const T *b, T *out, int64_t count, int op, cudaStream_t s)
{
    constexpr int L = 16 / sizeof(T);
    const bool vec = count % L == 0 && aligned_to(a, 16) && aligned_to(b, 16) && aligned_to(out, 16);
    const int64_t items = vec ? count / L : count;
    int64_t blocks = (items + kThreads - 1) / kThreads;
    if (blocks > 148 * 64) blocks = 148 * 64;
    if (blocks < 1) blocks = 1;
    if (vec && op == 0)
        binary_vec_kernel<T, 0><<<(unsigned)blocks, kThreads, 0, s>>>(a, b, out, items);
    else if (vec)
        binary_vec_kernel<T, 1><<<(unsigned)blocks, kThreads, 0, s>>>(a, b, out, items);
    else if (op == 0)
        binary_kernel<T, 0><<<(unsigned)blocks, kThreads, 0, s>>>(a, b, out, count);
    else
        binary_kernel<T, 1><<<(unsigned)blocks, kThreads, 0, s>>>(a, b, out, count);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

// ---------------------------------------------------------------------------
// CG vector updates (recipes.iadrt_cg; the reference's recipe hands the normal operator to
// scipy.sparse.linalg.cg, docs/examples.cginverse.md:40-67).  One iteration = the operator plus
//   dot:       state[1] = p . Ap
//   update:    alpha = state[0] / state[1];  x += alpha p;  r -= alpha Ap;  state[2] = r . r
//   direction: beta = state[2] / state[0];   p = r + beta p;  state[0] = state[2]
// i.e. three passes over the vectors instead of a dozen elementwise ones.  Dot products: per-thread
// and per-block partial sums in double in a fixed order, a one-block pass adds the block partials --
// deterministic, so ranks that hold replicas of the vectors stay bit-identical.
constexpr int kCgBlocks = 148 * 8;

template <typename T>
__device__ __forceinline__ double cg_block_sum(double v)
{
    __shared__ double warp_sums[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += warp_sums[w];
    return t;   // valid in thread 0
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cg_dot_kernel(const T *__restrict__ a, const T *__restrict__ b, double *__restrict__ partials, int64_t count)
{
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kThreads)
        acc += (double)a[i] * (double)b[i];
    const double t = cg_block_sum<T>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cg_update_kernel(T *__restrict__ x, T *__restrict__ r, const T *__restrict__ p, const T *__restrict__ ap,
                 const double *__restrict__ state, double *__restrict__ partials, int64_t count)
{
    const T alpha = (T)(state[0] / state[1]);
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kThreads) {
        x[i] = x[i] + alpha * p[i];
        const T rn = r[i] - alpha * ap[i];
        r[i] = rn;
        acc += (double)rn * (double)rn;
    }
    const double t = cg_block_sum<T>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kThreads)
cg_finish_kernel(const double *__restrict__ partials, int n, double *__restrict__ out)
{
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) acc += partials[i];
    const double t = cg_block_sum<double>(acc);
    if (threadIdx.x == 0) *out = t;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cg_direction_kernel(T *__restrict__ p, const T *__restrict__ r, const double *__restrict__ state, int64_t count)
{
    const T beta = (T)(state[2] / state[0]);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kThreads)
        p[i] = r[i] + beta * p[i];
}

__global__ void cg_shift_kernel(double *state) { state[0] = state[2]; }

size_t cg_workspace_bytes() { return (size_t)kCgBlocks * sizeof(double); }

template <typename T>
int launch_cg_dot(const T *a, const T *b, double *state, int slot, int64_t count, double *partials, cudaStream_t s)
{
    cg_dot_kernel<T><<<kCgBlocks, kThreads, 0, s>>>(a, b, partials, count);
    ADRT_LAUNCH_CHECK();
    cg_finish_kernel<<<1, kThreads, 0, s>>>(partials, kCgBlocks, state + slot);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_cg_update(T *x, T *r, const T *p, const T *ap, double *state, int64_t count, double *partials, cudaStream_t s)
{
    cg_update_kernel<T><<<kCgBlocks, kThreads, 0, s>>>(x, r, p, ap, state, partials, count);
    ADRT_LAUNCH_CHECK();
    cg_finish_kernel<<<1, kThreads, 0, s>>>(partials, kCgBlocks, state + 2);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T>
int launch_cg_direction(T *p, const T *r, double *state, int64_t count, cudaStream_t s)
{
    cg_direction_kernel<T><<<kCgBlocks, kThreads, 0, s>>>(p, r, state, count);
    ADRT_LAUNCH_CHECK();
    cg_shift_kernel<<<1, 1, 0, s>>>(state);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

#define INSTANTIATE(T)                                                                                  \
    template int launch_adrt_init<T>(const T *, T *, int64_t, int64_t, cudaStream_t);                  \
    template int launch_adrt_step<T>(const T *, T *, int64_t, int64_t, int, cudaStream_t);             \
    template int launch_bdrt_step<T>(const T *, T *, int64_t, int64_t, int, bool, cudaStream_t);       \
    template int launch_iadrt_stage<T>(const T *, T *, int64_t, int64_t, int, cudaStream_t);           \
    template int launch_fmg_restriction<T>(const T *, T *, int64_t, int64_t, cudaStream_t);            \
    template int launch_fmg_prolongation<T>(const T *, T *, int64_t, int64_t, int64_t, cudaStream_t);  \
    template int launch_fmg_highpass<T>(const T *, T *, int64_t, int64_t, int64_t, cudaStream_t);      \
    template int launch_interp_to_cart<T>(const T *, T *, const float *, const int32_t *, const float *, const float *, const int32_t *, const T *, int64_t, int64_t, cudaStream_t); \
    template int launch_truncate<T>(const T *, T *, int64_t, int64_t, cudaStream_t);                   \
    template int launch_truncate_mean<T>(const T *, T *, int64_t, int64_t, T, cudaStream_t);           \
    template int launch_truncate_mean_shares<T>(const T *, T *, int64_t, int64_t, int, int, T, cudaStream_t); \
    template int launch_stitch<T>(const T *, T *, int64_t, int64_t, bool, cudaStream_t);               \
    template int launch_unstitch<T>(const T *, T *, int64_t, int64_t, bool, cudaStream_t);             \
    template int launch_binary<T>(const T *, const T *, T *, int64_t, int, cudaStream_t);             \
    template int launch_cg_dot<T>(const T *, const T *, double *, int, int64_t, double *, cudaStream_t); \
    template int launch_cg_update<T>(T *, T *, const T *, const T *, double *, int64_t, double *, cudaStream_t); \
    template int launch_cg_direction<T>(T *, const T *, double *, int64_t, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace adrt_b200
