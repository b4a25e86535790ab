// Tile-level algorithms of the fused multi-stage ADRT / bdrt passes.
//
// This header is plain C++ that compiles both as CUDA device code (included by
// fused_adrt.cu) and as host code (tests/emu/emu_fused.cpp runs every phase
// for tid = 0..NT-1 with the barriers in between, so the index algebra is
// checked bit-for-bit against the oracle on a machine without a GPU).
//
// ---------------------------------------------------------------------------
// The maths (SURVEY.md section 8a rows a2', a5'; reference adrt_cdefs_adrt.hpp:55-96,
// adrt_cdefs_bdrt.hpp:55-116)
//
// After s butterfly stages a quadrant is a set of n/e row-blocks (e = 2^s), each
// holding e "angles"; row (blk, a) is a function of the offset d with support
// d < n + a.  A *pass* fuses the next M stages (G = 2^M).  For the group
// g = k0*e + a_g the G input rows  j -> (blk = k0*G + j, a_g)  produce the G
// output rows  p -> (blk' = k0, a' = a_g*G + p):
//
//     out_p[d] = sum_j in_j[d - a_g*j - small(j, p)]          (forward)
//     out_j[d - a_g*j] = sum_p in_p[d + small(j, p)]          (transposed, bdrt)
//
// combined in the reference's radix-2 tree order.  Absorbing the uniform
// per-row shift a_g*j into the global<->shared copy makes the shared-memory
// computation the same for every group: a *local* M-stage transform that
// starts from angle 0.  Only rows of the same group interact, exactly like a
// decimation-in-time FFT.
//
// Shared-memory tile: G rows (one per input/output row of the group) x XT
// offsets, offset axis contiguous ("R-layout").  Offsets are padded by one
// word per 32 (phys()), and the row pitch is odd, so that
//   * a warp whose lanes own V = 8 consecutive offsets each reads/writes any
//     uniformly shifted position without bank conflicts, and
//   * a warp whose lanes walk the rows at a fixed offset (the transposing
//     copies to/from the public (d, c) layout) is conflict free as well.
// A thread computes a radix-4 butterfly (two stages) for 8 consecutive offsets
// in registers: 38 shared loads and 32 stores for 66 adds, instead of the 128
// loads / 64 stores of two separate stages.
//
// Signed zeros / missing operands: the reference copies instead of adding when
// the shifted operand does not exist.  Forward: positions below offset 0 hold
// -0.0 (x + -0.0 == x bit for bit, also for x = -0.0) and real zero padding is
// +0.0.  Transposed: a missing first operand is +0.0 and a missing second
// operand -0.0; because "missing" means "row >= D of that intermediate", tiles
// that touch the end of the offset axis take a masked path (kMask).
#pragma once

#ifdef __CUDACC__
#define ADRT_HD __host__ __device__ __forceinline__
#else
#define ADRT_HD inline
#endif

namespace adrt_b200 {
namespace tile {

constexpr int V = 8;                 // consecutive offsets per thread
constexpr int XW = 256;              // computed offsets per tile (32 lanes x V)
constexpr int MARGIN = 8;            // slack for reads just outside the computed window
constexpr int XT = XW + MARGIN;      // offsets held per row
constexpr int NT = 256;              // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int NCHUNK = XW / V;       // = 32: one warp covers one row-group

ADRT_HD constexpr int phys(int xt) { return xt + (xt >> 5); }
constexpr int PITCH = 273;           // >= phys(XT - 1) + 1 = 272, odd

enum LoadKind { LOAD_IMAGE = 0, LOAD_WROWS = 1, LOAD_QCOLS = 2 };
enum StoreKind { STORE_WROWS = 0, STORE_QCOLS = 1 };

template <int M> struct Geo {
    static constexpr int G = 1 << M;
    static constexpr int HALO = G - 1;        // total shift consumed by M stages
    static constexpr int TD = XW - HALO;      // valid output offsets per tile
};

// Everything a CTA needs to know about its tile.
struct TileCtx {
    int n, D;          // image side, 2n-1
    int q;             // quadrant (image loader only)
    int e;             // block height before the pass, 2^s
    int k0, a_g;       // group = k0*e + a_g
    int g;             // group index
    int d0;            // forward: first valid output offset; transposed: first input offset
    long long in_pitch, out_pitch;   // elements per row of the R-layout workspaces
};

// ===========================================================================
// forward
// ===========================================================================

// ---- loaders: fill rows j of buf with in_j[d0 - HALO - MARGIN - a_g*j + xt] ----
template <typename T, int M>
ADRT_HD void fwd_load_wrows(T *buf, const T *src_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, HALO = Geo<M>::HALO;
    const int warp = tid >> 5, lane = tid & 31;
    const int sup = c.n + c.a_g;  // support of every input row of this group
    for (int j = warp; j < G; j += NWARP) {
        const T *row = src_plane + ((long long)(c.k0 * G + j) * c.e + c.a_g) * c.in_pitch;
        const int dbase = c.d0 - HALO - MARGIN - c.a_g * j;
        for (int xt = lane; xt < XT; xt += 32) {
            const int d = dbase + xt;
            T v;
            if (d < 0) v = T(-0.0);
            else if (d < sup) v = row[d];
            else v = T(0.0);
            buf[j * PITCH + phys(xt)] = v;
        }
    }
}

// Image loader (first pass, e = 1, a_g = 0): row j is oriented image row
// r = g*G + j of quadrant q (core.py:169-176):
//   q0: I[r][d] = x[r, n-1-d]      q1: I[r][d] = x[n-1-d, r]
//   q2: I[r][d] = x[d, r]          q3: I[r][d] = x[n-1-r, n-1-d]
template <typename T, int M>
ADRT_HD void fwd_load_image(T *buf, const T *img, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, HALO = Geo<M>::HALO;
    const int warp = tid >> 5, lane = tid & 31;
    const int n = c.n;
    const int dbase = c.d0 - HALO - MARGIN;
    const int rows = G < n ? G : n;
    if (c.q == 0 || c.q == 3) {
        // image rows are contiguous along d (reversed): lanes walk d
        for (int j = warp; j < rows; j += NWARP) {
            const int r = c.g * G + j;
            const T *row = img + (long long)(c.q == 0 ? r : n - 1 - r) * n;
            for (int xt = lane; xt < XT; xt += 32) {
                const int d = dbase + xt;
                T v;
                if (d < 0) v = T(-0.0);
                else if (d < n) v = row[n - 1 - d];
                else v = T(0.0);
                buf[j * PITCH + phys(xt)] = v;
            }
        }
    } else {
        // image rows are contiguous along r: lanes walk the tile rows j
        for (int xt = warp; xt < XT; xt += NWARP) {
            const int d = dbase + xt;
            const int px = phys(xt);
            for (int j = lane; j < rows; j += 32) {
                const int r = c.g * G + j;
                T v;
                if (d < 0) v = T(-0.0);
                else if (d < n) v = (c.q == 1) ? img[(long long)(n - 1 - d) * n + r] : img[(long long)d * n + r];
                else v = T(0.0);
                buf[j * PITCH + px] = v;
            }
        }
    }
}

// ---- radix-4 step: local stages t and t+1 (e = 2^t) ---------------------------
//   u[k][al][d]  = in_{2k}[d] + in_{2k+1}[d - a - al]                 (stage t)
//   out[p][d]    = u[0][p>>1][d] + u[1][p>>1][d - 2a - ceil(p/2)]     (stage t+1)
// with input rows r_j = (k0*4 + j)*e + a, output rows k0*4e + 4a + p.
template <typename T, int M>
ADRT_HD void fwd_radix4(const T *in, T *out, int t, int tid)
{
    constexpr int G = Geo<M>::G;
    const int e = 1 << t;
    const int lo_out = 4 * e - 1;  // offsets below this are not valid after the step
    for (int item = tid; item < (G / 4) * NCHUNK; item += NT) {
        const int gi = item / NCHUNK, ch = item % NCHUNK;
        if (V * ch + V <= lo_out) continue;
        const int k0 = gi >> t, a = gi & (e - 1);
        const int x = MARGIN + V * ch;
        const T *r0 = in + ((k0 * 4 + 0) * e + a) * PITCH;
        const T *r1 = in + ((k0 * 4 + 1) * e + a) * PITCH;
        const T *r2 = in + ((k0 * 4 + 2) * e + a) * PITCH;
        const T *r3 = in + ((k0 * 4 + 3) * e + a) * PITCH;
        T y0[V], y1[V + 1], y2[V + 2], y3[V + 3];
#pragma unroll
        for (int i = 0; i < V; ++i) y0[i] = r0[phys(x + i)];
#pragma unroll
        for (int i = 0; i < V + 1; ++i) y1[i] = r1[phys(x - a - 1 + i)];
#pragma unroll
        for (int i = 0; i < V + 2; ++i) y2[i] = r2[phys(x - 2 * a - 2 + i)];
#pragma unroll
        for (int i = 0; i < V + 3; ++i) y3[i] = r3[phys(x - 3 * a - 3 + i)];
        T u00[V], u01[V], u10[V + 2], u11[V + 2];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            u00[i] = y0[i] + y1[i + 1];
            u01[i] = y0[i] + y1[i];
        }
#pragma unroll
        for (int i = 0; i < V + 2; ++i) {
            u10[i] = y2[i] + y3[i + 1];
            u11[i] = y2[i] + y3[i];
        }
        T *o = out + (k0 * 4 * e + 4 * a) * PITCH;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int px = phys(x + i);
            o[0 * PITCH + px] = u00[i] + u10[i + 2];
            o[1 * PITCH + px] = u00[i] + u10[i + 1];
            o[2 * PITCH + px] = u01[i] + u11[i + 1];
            o[3 * PITCH + px] = u01[i] + u11[i];
        }
    }
}

// ---- radix-2 step: local stage t ------------------------------------------------
template <typename T, int M>
ADRT_HD void fwd_radix2(const T *in, T *out, int t, int tid)
{
    constexpr int G = Geo<M>::G;
    const int e = 1 << t;
    const int lo_out = 2 * e - 1;
    for (int item = tid; item < (G / 2) * NCHUNK; item += NT) {
        const int gi = item / NCHUNK, ch = item % NCHUNK;
        if (V * ch + V <= lo_out) continue;
        const int k = gi >> t, b = gi & (e - 1);
        const int x = MARGIN + V * ch;
        const T *rA = in + ((2 * k) * e + b) * PITCH;
        const T *rB = in + ((2 * k + 1) * e + b) * PITCH;
        T yA[V], yB[V + 1];
#pragma unroll
        for (int i = 0; i < V; ++i) yA[i] = rA[phys(x + i)];
#pragma unroll
        for (int i = 0; i < V + 1; ++i) yB[i] = rB[phys(x - b - 1 + i)];
        T *o = out + (k * 2 * e + 2 * b) * PITCH;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int px = phys(x + i);
            o[px] = yA[i] + yB[i + 1];
            o[PITCH + px] = yA[i] + yB[i];
        }
    }
}

// Number of barrier-separated compute steps for M stages (radix-4 first, one
// radix-2 at the end when M is odd) and which buffer holds the result.
ADRT_HD constexpr int num_steps(int M) { return (M + 1) / 2; }

// step i of the forward local transform; returns nothing, caller syncs.
template <typename T, int M>
ADRT_HD void fwd_step(T *bufA, T *bufB, int step, int tid)
{
    const T *in = (step & 1) ? bufB : bufA;
    T *out = (step & 1) ? bufA : bufB;
    const int t = 2 * step;
    if (t + 2 <= M) fwd_radix4<T, M>(in, out, t, tid);
    else fwd_radix2<T, M>(in, out, t, tid);
}

// ---- stores -----------------------------------------------------------------------
// R-layout workspace: output row p -> row (g*G + p), offsets [d0, d0+TD) below the
// support bound of the row (n + a'), a' = a_g*G + p; nothing else is ever read back.
template <typename T, int M>
ADRT_HD void fwd_store_wrows(const T *buf, T *dst_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, HALO = Geo<M>::HALO;
    const int warp = tid >> 5, lane = tid & 31;
    for (int p = warp; p < G; p += NWARP) {
        T *row = dst_plane + ((long long)c.g * G + p) * c.out_pitch;
        int lim = c.n + c.a_g * G + p;
        if (lim > c.D) lim = c.D;
        for (int xc = HALO + lane; xc < XW; xc += 32) {
            const int d = c.d0 + xc - HALO;
            if (d < lim) row[d] = buf[p * PITCH + phys(MARGIN + xc)];
        }
    }
}

// Public layout (D, n) of the plane: column g*G + p, all offsets < D.
template <typename T, int M>
ADRT_HD void fwd_store_qcols(const T *buf, T *dst_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, HALO = Geo<M>::HALO;
    const int warp = tid >> 5, lane = tid & 31;
    for (int xc = HALO + warp; xc < XW; xc += NWARP) {
        const int d = c.d0 + xc - HALO;
        if (d >= c.D) break;
        const int px = phys(MARGIN + xc);
        T *orow = dst_plane + (long long)d * c.n + c.g * G;
        for (int p = lane; p < G; p += 32) orow[p] = buf[p * PITCH + px];
    }
}

// Tile entirely above the support of all its output columns: plain zeros.
template <typename T, int M>
ADRT_HD void fwd_store_qcols_zero(T *dst_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, TD = Geo<M>::TD;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < TD; i += NWARP) {
        const int d = c.d0 + i;
        if (d >= c.D) break;
        T *orow = dst_plane + (long long)d * c.n + c.g * G;
        for (int p = lane; p < G; p += 32) orow[p] = T(0.0);
    }
}

// ===========================================================================
// transposed (bdrt)
// ===========================================================================
// Tile coordinate xc = xt (margin on the right); input row p holds
// in_p[d0 + xt].  Output row j is stored at offset d0 + xc - a_g*j.

template <typename T, int M>
ADRT_HD void bwd_load_wrows(T *buf, const T *src_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G;
    const int warp = tid >> 5, lane = tid & 31;
    for (int p = warp; p < G; p += NWARP) {
        const T *row = src_plane + ((long long)c.g * G + p) * c.in_pitch;
        for (int xt = lane; xt < XT; xt += 32) {
            const int d = c.d0 + xt;
            buf[p * PITCH + phys(xt)] = d < c.D ? row[d] : T(0.0);
        }
    }
}

template <typename T, int M>
ADRT_HD void bwd_load_qcols(T *buf, const T *src_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G;
    const int warp = tid >> 5, lane = tid & 31;
    const int cols = G < c.n ? G : c.n;
    for (int xt = warp; xt < XT; xt += NWARP) {
        const int d = c.d0 + xt;
        const int px = phys(xt);
        const T *irow = src_plane + (long long)d * c.n + c.g * G;
        for (int p = lane; p < cols; p += 32) buf[p * PITCH + px] = d < c.D ? irow[p] : T(0.0);
    }
}

// Missing-operand rule of bdrt_core (adrt_cdefs_bdrt.hpp:96-109) for a value read
// at tile position `pos` from a row whose intermediate ends at `lim`.
template <typename T, bool kMask>
ADRT_HD T bmask(T v, int pos, int lim, bool odd)
{
    if (kMask) return pos >= lim ? (odd ? T(-0.0) : T(0.0)) : v;
    return v;
}

// ---- transposed radix-4 step: from local stage t+2 rows back to stage t rows ----
//   gu[0][b][d]  = gin[2b][d] + gin[2b+1][d]
//   gu[1][b][d'] = gin[2b][d' + 2a + b] + gin[2b+1][d' + 2a + b + 1]
//   out_{2k}[d]    = gu[k][0][d] + gu[k][1][d]
//   out_{2k+1}[d'] = gu[k][0][d' + a] + gu[k][1][d' + a + 1]
// `dt` = D - d0 and `ag` = global base angle give each row's end:
//   parent rows (block k0 at stage t+2): dt + ag*k0*4e;  node k=1: + ag*2e.
template <typename T, int M, bool kMask>
ADRT_HD void bwd_radix4(const T *in, T *out, int t, int dt, int ag, int tid)
{
    constexpr int G = Geo<M>::G;
    const int e = 1 << t;
    for (int item = tid; item < (G / 4) * NCHUNK; item += NT) {
        const int gi = item / NCHUNK, ch = item % NCHUNK;
        const int k0 = gi >> t, a = gi & (e - 1);
        const int x = V * ch;
        const T *ip = in + (k0 * 4 * e + 4 * a) * PITCH;
        const int lim_p = dt + ag * (k0 * 4 * e);
        const int lim_1 = lim_p + ag * 2 * e;
        T g0[V], g1[V + 1], g2[V + 2], g3[V + 3];
#pragma unroll
        for (int i = 0; i < V; ++i) g0[i] = bmask<T, kMask>(ip[0 * PITCH + phys(x + i)], x + i, lim_p, false);
#pragma unroll
        for (int i = 0; i < V + 1; ++i) g1[i] = bmask<T, kMask>(ip[1 * PITCH + phys(x + i)], x + i, lim_p, true);
#pragma unroll
        for (int i = 0; i < V + 2; ++i) g2[i] = bmask<T, kMask>(ip[2 * PITCH + phys(x + i)], x + i, lim_p, false);
#pragma unroll
        for (int i = 0; i < V + 3; ++i) g3[i] = bmask<T, kMask>(ip[3 * PITCH + phys(x + i)], x + i, lim_p, true);
        T u00[V], u01[V + 1], u10[V], u11[V + 1];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            u00[i] = g0[i] + g1[i];
            u10[i] = g0[i] + g1[i + 1];
        }
#pragma unroll
        for (int i = 0; i < V + 1; ++i) {
            // odd-angle children: a missing entry must act as -0.0 when it is the
            // second operand below
            u01[i] = bmask<T, kMask>(g2[i] + g3[i], x + i, lim_p, true);
            u11[i] = bmask<T, kMask>(g2[i + 1] + g3[i + 2], x - 2 * a + i, lim_1, true);
        }
        T *o0 = out + ((k0 * 4 + 0) * e + a) * PITCH;
        T *o1 = out + ((k0 * 4 + 1) * e + a) * PITCH;
        T *o2 = out + ((k0 * 4 + 2) * e + a) * PITCH;
        T *o3 = out + ((k0 * 4 + 3) * e + a) * PITCH;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int p0 = x + i, p1 = x - a + i, p2 = x - 2 * a + i, p3 = x - 3 * a + i;
            o0[phys(p0)] = u00[i] + u01[i];
            if (p1 >= 0) o1[phys(p1)] = u00[i] + u01[i + 1];
            if (p2 >= 0) o2[phys(p2)] = u10[i] + u11[i];
            if (p3 >= 0) o3[phys(p3)] = u10[i] + u11[i + 1];
        }
    }
}

// ---- transposed radix-2 step: from local stage t+1 rows back to stage t rows ----
template <typename T, int M, bool kMask>
ADRT_HD void bwd_radix2(const T *in, T *out, int t, int dt, int ag, int tid)
{
    constexpr int G = Geo<M>::G;
    const int e = 1 << t;
    for (int item = tid; item < (G / 2) * NCHUNK; item += NT) {
        const int gi = item / NCHUNK, ch = item % NCHUNK;
        const int k = gi >> t, b = gi & (e - 1);
        const int x = V * ch;
        const T *ip = in + (k * 2 * e + 2 * b) * PITCH;
        const int lim_p = dt + ag * (k * 2 * e);
        T g0[V], g1[V + 1];
#pragma unroll
        for (int i = 0; i < V; ++i) g0[i] = bmask<T, kMask>(ip[phys(x + i)], x + i, lim_p, false);
#pragma unroll
        for (int i = 0; i < V + 1; ++i) g1[i] = bmask<T, kMask>(ip[PITCH + phys(x + i)], x + i, lim_p, true);
        T *oA = out + ((2 * k) * e + b) * PITCH;
        T *oB = out + ((2 * k + 1) * e + b) * PITCH;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int pB = x - b + i;
            oA[phys(x + i)] = g0[i] + g1[i];
            if (pB >= 0) oB[phys(pB)] = g0[i] + g1[i + 1];
        }
    }
}

// Transposed steps run the forward schedule backwards: forward step i covers
// local stages 2i (and 2i+1); transposed step i undoes forward step nsteps-1-i.
template <typename T, int M, bool kMask>
ADRT_HD void bwd_step(T *bufA, T *bufB, int step, int dt, int ag, int tid)
{
    const T *in = (step & 1) ? bufB : bufA;
    T *out = (step & 1) ? bufA : bufB;
    const int t = 2 * (num_steps(M) - 1 - step);
    if (t + 2 <= M) bwd_radix4<T, M, kMask>(in, out, t, dt, ag, tid);
    else bwd_radix2<T, M, kMask>(in, out, t, dt, ag, tid);
}

// Output row j -> workspace row (k0*G + j)*e + a_g at offset d0 + xc - a_g*j.
template <typename T, int M>
ADRT_HD void bwd_store_wrows(const T *buf, T *dst_plane, const TileCtx &c, bool zero, int tid)
{
    constexpr int G = Geo<M>::G, TD = Geo<M>::TD;
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < G; j += NWARP) {
        T *row = dst_plane + ((long long)(c.k0 * G + j) * c.e + c.a_g) * c.out_pitch;
        const int shift = c.a_g * j;
        for (int xc = lane; xc < TD; xc += 32) {
            const int d = c.d0 + xc - shift;
            if (d >= 0 && d < c.D) row[d] = zero ? T(0.0) : buf[j * PITCH + phys(xc)];
        }
    }
}

// Public layout: column g*G + j (last transposed pass: e = 1, a_g = 0).
template <typename T, int M>
ADRT_HD void bwd_store_qcols(const T *buf, T *dst_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, TD = Geo<M>::TD;
    const int warp = tid >> 5, lane = tid & 31;
    const int cols = G < c.n ? G : c.n;
    for (int xc = warp; xc < TD; xc += NWARP) {
        const int d = c.d0 + xc;
        if (d >= c.D) break;
        const int px = phys(xc);
        T *orow = dst_plane + (long long)d * c.n + c.g * G;
        for (int j = lane; j < cols; j += 32) orow[j] = buf[j * PITCH + px];
    }
}

}  // namespace tile
}  // namespace adrt_b200

// ===========================================================================
// Tile programs: the barrier-separated phase sequence of one CTA.  The CUDA
// kernels run phase(ph, ..., threadIdx.x) with __syncthreads() in between; the
// host emulator runs every phase for tid = 0..NT-1.
// ===========================================================================
namespace adrt_b200 {
namespace tile {

enum TileMode { TILE_SKIP = 0, TILE_ZERO = 1, TILE_FULL = 2, TILE_FULL_MASKED = 3 };

template <typename T, int M, int LOADK, int STOREK>
struct FwdProgram {
    static constexpr int G = Geo<M>::G;
    static constexpr int kPhases = 2 + num_steps(M);

    // c.d0 must already be set.
    ADRT_HD static int classify(const TileCtx &c)
    {
        int sup = c.n + c.a_g * G + G - 1;  // first offset at which every output row is zero
        if (sup > c.D) sup = c.D;
        if (c.d0 >= c.D) return TILE_SKIP;
        if (c.d0 >= sup) return STOREK == STORE_QCOLS ? TILE_ZERO : TILE_SKIP;
        return TILE_FULL;
    }

    ADRT_HD static void phase(int ph, int mode, T *bufA, T *bufB, const T *src, T *dst, const TileCtx &c, int tid)
    {
        if (mode == TILE_ZERO) {
            if (ph == 0) fwd_store_qcols_zero<T, M>(dst, c, tid);
            return;
        }
        if (ph == 0) {
            if (LOADK == LOAD_IMAGE) fwd_load_image<T, M>(bufA, src, c, tid);
            else fwd_load_wrows<T, M>(bufA, src, c, tid);
        } else if (ph <= num_steps(M)) {
            fwd_step<T, M>(bufA, bufB, ph - 1, tid);
        } else {
            const T *res = (num_steps(M) & 1) ? bufB : bufA;
            if (STOREK == STORE_QCOLS) fwd_store_qcols<T, M>(res, dst, c, tid);
            else fwd_store_wrows<T, M>(res, dst, c, tid);
        }
    }
};

template <typename T, int M, int LOADK, int STOREK>
struct BwdProgram {
    static constexpr int G = Geo<M>::G;
    static constexpr int kPhases = 2 + num_steps(M);

    ADRT_HD static int classify(const TileCtx &c)
    {
        if (c.d0 >= c.D + c.a_g * (G - 1)) return TILE_SKIP;   // no output row reaches this far
        if (c.d0 >= c.D) return STOREK == STORE_QCOLS ? TILE_SKIP : TILE_ZERO;
        if (c.d0 + XT > c.D) return TILE_FULL_MASKED;
        return TILE_FULL;
    }

    ADRT_HD static void phase(int ph, int mode, T *bufA, T *bufB, const T *src, T *dst, const TileCtx &c, int tid)
    {
        if (mode == TILE_ZERO) {
            if (ph == 0) bwd_store_wrows<T, M>(bufA, dst, c, true, tid);
            return;
        }
        if (ph == 0) {
            if (LOADK == LOAD_QCOLS) bwd_load_qcols<T, M>(bufA, src, c, tid);
            else bwd_load_wrows<T, M>(bufA, src, c, tid);
        } else if (ph <= num_steps(M)) {
            if (mode == TILE_FULL_MASKED) bwd_step<T, M, true>(bufA, bufB, ph - 1, c.D - c.d0, c.a_g, tid);
            else bwd_step<T, M, false>(bufA, bufB, ph - 1, c.D - c.d0, c.a_g, tid);
        } else {
            const T *res = (num_steps(M) & 1) ? bufB : bufA;
            if (STOREK == STORE_QCOLS) bwd_store_qcols<T, M>(res, dst, c, tid);
            else bwd_store_wrows<T, M>(res, dst, c, false, tid);
        }
    }
};

}  // namespace tile
}  // namespace adrt_b200
