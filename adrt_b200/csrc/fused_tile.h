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
// Shared-memory tile: G rows x 256 offsets, offset axis contiguous, linear
// (no padding inside a row), row pitch = 4 words mod 32.  A thread owns V = 4
// consecutive offsets of one radix-4 butterfly (two stages): consecutive lanes
// own consecutive 16-byte slots, so every LDS.128/STS.128 is conflict free,
// and so are the transposing copies (lanes walk the rows, 16 bytes each).
// Shifted operand windows are fetched as the enclosing aligned 16-byte slots
// and the wanted elements are picked by *compile-time* register indices: the
// alignment residue depends only on (angle mod 4), which is warp-uniform, so
// each step is instantiated for the 4 residues and dispatched by a switch.
// Outputs are always stored aligned.  In the transposed direction this means a
// child row is stored with a skew (j * angle) relative to its logical offsets;
// the next step folds the skew into its operand windows (again only the skew
// mod 4 matters for code selection).
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

#ifndef ADRT_QCOLS_PAIR_F32
#define ADRT_QCOLS_PAIR_F32 0
#endif

constexpr int V = 4;                 // consecutive offsets per thread
constexpr int XW = 256;              // offsets per tile row
constexpr int NCHUNK = XW / V;       // 64 chunks per row

// 16-byte vector access -----------------------------------------------------------
template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int L = 4; };
template <> struct VecOf<double> { static constexpr int L = 2; };
template <typename T> struct alignas(16) Pack { T v[VecOf<T>::L]; };

// Row pitch (elements): >= XW + 8 slack for over-fetch, 16-byte multiple, and
// = 4 words (mod 32) so that 8 lanes reading 16 bytes from 8 consecutive rows
// hit 32 distinct banks.
template <typename T> struct Pitch;
template <> struct Pitch<float> { static constexpr int value = 260; };
template <> struct Pitch<double> { static constexpr int value = 274; };

enum LoadKind { LOAD_IMAGE = 0, LOAD_WROWS = 1, LOAD_QCOLS = 2 };
enum StoreKind { STORE_WROWS = 0, STORE_QCOLS = 1 };

template <int M> struct Geo {
    static constexpr int G = 1 << M;
    // offsets consumed by M stages, rounded so that chunk boundaries stay aligned
    static constexpr int HALO = G < 4 ? 4 : G;
    static constexpr int NT = G >= 64 ? 512 : 256;   // threads per CTA: one warp per radix-4 group
    static constexpr int MIN_CTAS = G >= 64 ? 2 : (G == 32 ? 5 : 3); // launch-bounds target (caps registers at 64)
    static constexpr int NWARP = NT / 32;
};

// Everything a CTA needs to know about its tile.
struct TileCtx {
    int n, D;          // image side, 2n-1
    int q;             // quadrant (image loader only)
    int e;             // block height before (forward) / after (transposed) the pass, 2^s
    int k0, a_g;       // group = k0*e + a_g
    int g;             // group index
    int d0;            // forward: first valid output offset; transposed: first input offset
    long long in_pitch, out_pitch;   // elements per row of the R-layout workspaces (multiples of 4)
    int next_g;        // forward: group size (rows) of the pass that will read the workspace, 0 if none
    int d_need;        // transposed: only output offsets < d_need are wanted (D: all of them)
    // transposed passes that read workspace rows of a producer which skips its all-zero tiles
    // (plan::Pass::skip_zero): row r of the workspace is leaf j = (r >> sup_loge) & sup_gmask of the
    // producer's group with angle a = r & ((1 << sup_loge) - 1), and holds data only at offsets below
    // D - a*j (everything above is the structural +0.0 that the producer no longer writes);
    // sup_gmask < 0: rows are complete up to D
    int sup_loge, sup_gmask;
    // transposed passes that read the public layout, kSub instantiations only: byte distance from the
    // sinogram to a second one of the same shape that is subtracted on load (the residual
    // adrt(x) - b of iadrt_fmg_step, core.py:329, without a pass of its own); 0: none
    long long sub_delta = 0;
};

// first offset at which workspace row r of a transposed plan is structurally zero
ADRT_HD int bwd_row_support(const TileCtx &c, long long r)
{
    if (c.sup_gmask < 0) return c.D;
    const int a = (int)(r & ((1LL << c.sup_loge) - 1)), j = (int)((r >> c.sup_loge) & c.sup_gmask);
    const int sup = c.D - a * j;
    return sup;
}

// Valid offsets a tile produces.  Passes that store workspace rows give up 4
// more offsets: a row may be shifted by up to 3 elements against the 16-byte
// grid of global memory (forward: storage skew, transposed: a_g*j), and each
// tile then owns whole aligned 16-byte chunks of every row.
template <int M, int STOREK> struct TileTD {
    static constexpr int value = XW - Geo<M>::HALO - (STOREK == STORE_WROWS ? 4 : 0);
};

// Forward workspace rows are stored with a skew of s = (a * j) & 3 elements,
// (a = angle of the row, j = its index inside the group of the pass that reads
// it) so that the reader's shifted row start d - a*j lands on a 16-byte boundary.
ADRT_HD int fwd_row_skew(int angle, int j) { return (angle * j) & 3; }

// Load N consecutive elements that start Q elements after the 16-byte aligned
// position `p` (Q compile time): ceil((Q+N)/L) vector loads + static selection.
template <typename T, int N, int Q>
ADRT_HD void load_window(const T *p, T (&dst)[N])
{
    constexpr int L = VecOf<T>::L;
    constexpr int NV = (Q + N + L - 1) / L;
    Pack<T> tmp[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) tmp[v] = *reinterpret_cast<const Pack<T> *>(p + v * L);
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = tmp[(Q + i) / L].v[(Q + i) % L];
}

// one 16-byte vector
template <typename T>
ADRT_HD void store_cv(T *p, const T *src)
{
    constexpr int L = VecOf<T>::L;
    Pack<T> t;
#pragma unroll
    for (int i = 0; i < L; ++i) t.v[i] = src[i];
    *reinterpret_cast<Pack<T> *>(p) = t;
}

template <typename T>
ADRT_HD void store_chunk(T *p, const T *src)
{
    constexpr int L = VecOf<T>::L;
#pragma unroll
    for (int v = 0; v < V / L; ++v) {
        Pack<T> t;
#pragma unroll
        for (int i = 0; i < L; ++i) t.v[i] = src[v * L + i];
        *reinterpret_cast<Pack<T> *>(p + v * L) = t;
    }
}

// residue of a non-negative start position (x multiple of 4) minus `shift`
ADRT_HD constexpr int neg_mod(int shift, int L) { return ((-shift) % L + L) % L; }

// ===========================================================================
// forward
// ===========================================================================

// ---- loaders: row j of buf <- in_j[d0 - LH - a_g*j + x], x in [0, XW) ----------
// (LH = tile position of offset d0).  The stored row is skewed by (a_g*j)&3, which
// makes the global start 16-byte aligned.
template <typename T, int M, int LH>
ADRT_HD void fwd_load_wrows(T *buf, const T *src_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    constexpr int L = VecOf<T>::L;
    const int warp = tid >> 5, lane = tid & 31;
    const int sup = c.n + c.a_g;  // support of every input row of this group
    for (int j = warp; j < G; j += NWARP) {
        const T *row = src_plane + ((long long)(c.k0 * G + j) * c.e + c.a_g) * c.in_pitch;
        const int dbase = c.d0 - LH - c.a_g * j;          // logical offset of x = 0
        const int gbase = dbase + fwd_row_skew(c.a_g, j);  // its position in the stored row (multiple of 4)
        T *dst = buf + j * P;
        if (dbase >= 0 && dbase + XW <= sup) {
            Pack<T> v[XW / (32 * L)];
#pragma unroll
            for (int k = 0; k < XW / (32 * L); ++k) v[k] = *reinterpret_cast<const Pack<T> *>(row + gbase + (k * 32 + lane) * L);
#pragma unroll
            for (int k = 0; k < XW / (32 * L); ++k) *reinterpret_cast<Pack<T> *>(dst + (k * 32 + lane) * L) = v[k];
        } else if (dbase >= sup || dbase + XW <= 0) {
            // the whole window lies above the row's support (+0.0) or below offset 0 (the -0.0 sentinels): rows
            // are shifted by up to a_g * (G - 1) against the tile, so this is a quarter of the rows of a pass with
            // large block height (26 % of the row pairs of the second forward pass at 2048^2)
            Pack<T> z;
#pragma unroll
            for (int i = 0; i < L; ++i) z.v[i] = dbase >= sup ? T(0.0) : T(-0.0);
#pragma unroll
            for (int k = 0; k < XW / (32 * L); ++k) *reinterpret_cast<Pack<T> *>(dst + (k * 32 + lane) * L) = z;
        } else {
#pragma unroll 4
            for (int k = 0; k < XW / 32; ++k) {
                const int x = k * 32 + lane, d = dbase + x;
                T v;
                if (d < 0) v = T(-0.0);
                else if (d < sup) v = row[gbase + x];
                else v = T(0.0);
                dst[x] = v;
            }
        }
    }
}

// Image loader (first pass, e = 1, a_g = 0): row j is oriented image row
// r = g*G + j of quadrant q (core.py:169-176):
//   q0: I[r][d] = x[r, n-1-d]      q1: I[r][d] = x[n-1-d, r]
//   q2: I[r][d] = x[d, r]          q3: I[r][d] = x[n-1-r, n-1-d]
template <typename T, int M, int LH>
ADRT_HD void fwd_load_image(T *buf, const T *img, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    const int warp = tid >> 5, lane = tid & 31;
    const int n = c.n;
    const int dbase = c.d0 - LH;
    const int rows = G < n ? G : n;
    if (c.q == 0 || c.q == 3) {
        // image rows are contiguous along d (reversed): lanes walk d
        for (int j = warp; j < rows; j += NWARP) {
            const int r = c.g * G + j;
            const T *row = img + (long long)(c.q == 0 ? r : n - 1 - r) * n;
            T *dst = buf + j * P;
            if (dbase >= 0 && dbase + XW <= n) {
                // offsets x..x+3 are image columns n-1-d-3 .. n-1-d: one aligned vector, reversed
                // (n and d are multiples of 4 here)
                constexpr int NV = XW / (32 * V);
                T v[NV][V];
#pragma unroll
                for (int k = 0; k < NV; ++k) load_window<T, V, 0>(row + (n - V - (dbase + (k * 32 + lane) * V)), v[k]);
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    T r[V];
#pragma unroll
                    for (int i = 0; i < V; ++i) r[i] = v[k][V - 1 - i];
                    store_chunk<T>(dst + (k * 32 + lane) * V, r);
                }
            } else {
#pragma unroll 4
                for (int k = 0; k < XW / 32; ++k) {
                    const int x = k * 32 + lane, d = dbase + x;
                    T v;
                    if (d < 0) v = T(-0.0);
                    else if (d < n) v = row[n - 1 - d];
                    else v = T(0.0);
                    dst[x] = v;
                }
            }
        }
    } else {
        // image rows are contiguous along r: lanes walk the tile rows j, each
        // thread gathers V consecutive offsets and stores them as one vector
        constexpr int NIT = XW / (NWARP * V);
        const bool fast = dbase >= 0 && dbase + XW <= n;
        for (int j = lane; j < rows; j += 32) {
            const int r = c.g * G + j;
            T *b = buf + j * P;
            if (fast) {
                // q1 walks the image rows downwards, q2 upwards
                const long long step = (c.q == 1) ? -(long long)n : (long long)n;
                const T *ib = (c.q == 1) ? img + (long long)(n - 1 - dbase) * n + r : img + (long long)dbase * n + r;
#pragma unroll 1
                for (int it0 = 0; it0 < NIT; it0 += 2) {
                    T v[2][V];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int x = ((it0 + u) * NWARP + warp) * V;
#pragma unroll
                        for (int i = 0; i < V; ++i) v[u][i] = ib[(long long)(x + i) * step];
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) store_chunk<T>(b + ((it0 + u) * NWARP + warp) * V, v[u]);
                }
            } else {
#pragma unroll 2
                for (int it = 0; it < NIT; ++it) {
                    const int x = (it * NWARP + warp) * V;
                    T v[V];
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const int d = dbase + x + i;
                        if (d < 0) v[i] = T(-0.0);
                        else if (d < n) v[i] = (c.q == 1) ? img[(long long)(n - 1 - d) * n + r] : img[(long long)d * n + r];
                        else v[i] = T(0.0);
                    }
                    store_chunk<T>(b + x, v);
                }
            }
        }
    }
}

// Steps run IN PLACE on the single tile buffer: every thread first reads its
// operand windows and computes its outputs into registers (xxx_compute), the
// CTA synchronises, then the outputs are written over the tile (xxx_store) and
// the CTA synchronises again.  A thread owns one butterfly group (one warp per
// group) and the two chunks x = 4*lane and 4*lane + 128 of it, so row pointers
// and the alignment variant are set up once for 2 x 4 offsets.
constexpr int NREG = 32;   // outputs a thread holds across the barrier of a step
// In the butterfly steps a chunk is ONE 16-byte vector (4 floats / 2 doubles), so that
// consecutive lanes always touch consecutive 16-byte slots (conflict free for both
// types); a thread owns XW / (32 * chunk) chunks of its group: 2 (float) or 4 (double).

// ---- radix-4 step: local stages t and t+1 (e = 2^t) ---------------------------
//   u[k][al][d]  = in_{2k}[d] + in_{2k+1}[d - a - al]                 (stage t)
//   out[p][d]    = u[0][p>>1][d] + u[1][p>>1][d - 2a - ceil(p/2)]     (stage t+1)
// with input rows r_j = (k0*4 + j)*e + a, output rows k0*4e + 4a + p.
// AM = a & 3 fixes the alignment residues of the three shifted windows.
// the arithmetic of one chunk on operand windows that are already in registers
template <typename T>
ADRT_HD void fwd_radix4_math(const T (&y0)[VecOf<T>::L], const T (&y1)[VecOf<T>::L + 1], const T (&y2)[VecOf<T>::L + 2],
                             const T (&y3)[VecOf<T>::L + 3], T *o)
{
    constexpr int V = VecOf<T>::L;
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
#pragma unroll
    for (int i = 0; i < V; ++i) {
        o[0 * V + i] = u00[i] + u10[i + 2];
        o[1 * V + i] = u00[i] + u10[i + 1];
        o[2 * V + i] = u01[i] + u11[i + 1];
        o[3 * V + i] = u01[i] + u11[i];
    }
}

template <typename T, int AM>
ADRT_HD void fwd_radix4_chunk(const T *r0, const T *r1, const T *r2, const T *r3, int a, int x, T *o)
{
    constexpr int V = VecOf<T>::L;
    constexpr int L = VecOf<T>::L;
    constexpr int Q1 = neg_mod(AM + 1, L), Q2 = neg_mod(2 * AM + 2, L), Q3 = neg_mod(3 * AM + 3, L);
    T y0[V], y1[V + 1], y2[V + 2], y3[V + 3];
    load_window<T, V, 0>(r0 + x, y0);
    load_window<T, V + 1, Q1>(r1 + (x - a - 1 - Q1), y1);
    load_window<T, V + 2, Q2>(r2 + (x - 2 * a - 2 - Q2), y2);
    load_window<T, V + 3, Q3>(r3 + (x - 3 * a - 3 - Q3), y3);
    fwd_radix4_math<T>(y0, y1, y2, y3, o);
}

template <typename T, int AM>
ADRT_HD void fwd_radix4_group(const T *buf, int e, int k0, int a, int lane, int lo_out, T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int P = Pitch<T>::value;
    const T *r0 = buf + ((k0 * 4) * e + a) * P;
    const T *r1 = r0 + e * P, *r2 = r1 + e * P, *r3 = r2 + e * P;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int x = (lane + 32 * c) * V;
        if (x >= lo_out) fwd_radix4_chunk<T, AM>(r0, r1, r2, r3, a, x, &o[c * 4 * V]);
    }
}

template <typename T, int M, int t>
ADRT_HD void fwd_radix4_compute(const T *buf, int tid, T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int G = Geo<M>::G;
    constexpr int e = 1 << t, lo_out = 4 * e;
    const int gi = tid >> 5, lane = tid & 31;
    if (gi >= G / 4) return;
    const int k0 = gi >> t, a = gi & (e - 1);
    switch (a & 3) {
    case 0: fwd_radix4_group<T, 0>(buf, e, k0, a, lane, lo_out, o); break;
    case 1: fwd_radix4_group<T, 1>(buf, e, k0, a, lane, lo_out, o); break;
    case 2: fwd_radix4_group<T, 2>(buf, e, k0, a, lane, lo_out, o); break;
    default: fwd_radix4_group<T, 3>(buf, e, k0, a, lane, lo_out, o); break;
    }
}

template <typename T, int M, int t>
ADRT_HD void fwd_radix4_store(T *buf, int tid, const T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int G = Geo<M>::G, P = Pitch<T>::value;
    constexpr int e = 1 << t, lo_out = 4 * e;
    const int gi = tid >> 5, lane = tid & 31;
    if (gi >= G / 4) return;
    const int k0 = gi >> t, a = gi & (e - 1);
    T *orow = buf + (k0 * 4 * e + 4 * a) * P;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int x = (lane + 32 * c) * V;
        if (x >= lo_out) {
#pragma unroll
            for (int p = 0; p < 4; ++p) store_cv<T>(orow + p * P + x, &o[(c * 4 + p) * V]);
        }
    }
}

// ---- radix-2 step: local stage t; a thread serves two groups ---------------------
template <typename T, int BM>
ADRT_HD void fwd_radix2_group(const T *buf, int e, int k, int b, int lane, int lo_out, T *o)
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int P = Pitch<T>::value, L = VecOf<T>::L;
    constexpr int Q = neg_mod(BM + 1, L);
    const T *rA = buf + ((2 * k) * e + b) * P;
    const T *rB = rA + e * P;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int x = (lane + 32 * c) * V;
        if (x < lo_out) continue;
        T yA[V], yB[V + 1];
        load_window<T, V, 0>(rA + x, yA);
        load_window<T, V + 1, Q>(rB + (x - b - 1 - Q), yB);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            o[(c * 2 + 0) * V + i] = yA[i] + yB[i + 1];
            o[(c * 2 + 1) * V + i] = yA[i] + yB[i];
        }
    }
}

template <typename T, int M, int t>
ADRT_HD void fwd_radix2_compute(const T *buf, int tid, T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP;
    constexpr int e = 1 << t, lo_out = 2 * e < 4 ? 4 : 2 * e;
    const int lane = tid & 31;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int gi = (tid >> 5) + u * NWARP;
        if (gi >= G / 2) continue;
        const int k = gi >> t, b = gi & (e - 1);
        T *ou = &o[u * CHUNKS * 2 * V];
        switch (b & 3) {
        case 0: fwd_radix2_group<T, 0>(buf, e, k, b, lane, lo_out, ou); break;
        case 1: fwd_radix2_group<T, 1>(buf, e, k, b, lane, lo_out, ou); break;
        case 2: fwd_radix2_group<T, 2>(buf, e, k, b, lane, lo_out, ou); break;
        default: fwd_radix2_group<T, 3>(buf, e, k, b, lane, lo_out, ou); break;
        }
    }
}

template <typename T, int M, int t>
ADRT_HD void fwd_radix2_store(T *buf, int tid, const T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    constexpr int e = 1 << t, lo_out = 2 * e < 4 ? 4 : 2 * e;
    const int lane = tid & 31;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int gi = (tid >> 5) + u * NWARP;
        if (gi >= G / 2) continue;
        const int k = gi >> t, b = gi & (e - 1);
        T *orow = buf + (k * 2 * e + 2 * b) * P;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int x = (lane + 32 * c) * V;
            if (x >= lo_out) {
                store_cv<T>(orow + x, &o[((u * CHUNKS + c) * 2 + 0) * V]);
                store_cv<T>(orow + P + x, &o[((u * CHUNKS + c) * 2 + 1) * V]);
            }
        }
    }
}

// Image loader that also performs local stages 0 and 1 (first pass: e = 1, a_g = 0).  The four
// parents of butterfly k0 are the oriented image rows 4*k0 .. 4*k0+3 of the group:
//   q1, q2: four ADJACENT image columns -- one 16/32-byte piece of every image row the tile
//           needs; lanes = (butterfly pair, 4 chunks) so that a warp reads whole 128-byte row
//           pieces and its stores hit 8 distinct 16-byte bank groups;
//   q0, q3: four image rows read along the row (reversed): lanes walk the chunks.
// The thread runs the radix-4 step on what it fetched and stores the four children where
// fwd_radix4_store would put them: the tile is never staged as loaded (one tile write and the
// step's 1.9 tile reads of shared-memory traffic less).  Offsets below 0 are -0.0, offsets
// >= n are +0.0 (same rule as fwd_load_image).  Needs ~64 registers for its 4 x 7 operands: used
// by the fp64 passes (128 registers anyway) and the fp32 six-stage pass (64); fp32 five-stage passes
// are faster with the plain loader at 48 registers / 5 CTAs per SM (measured 292 vs 315 us).
// Interior tiles of the fp64 instantiations (every window of every job inside the image -- a CTA-uniform test):
// the jobs of TWO rounds are loaded before either is used.  A six-stage fp64 tile is alone on its SM (140 KB,
// 512 threads at 128 registers), and four dependent rounds of 8-10 16-byte loads per thread left only 64 KB per SM
// in flight: 38 % of the kernel's stall samples sat on the first use of each round
// (profiles/r06_pass_phases_f64.txt).  Same loads, same adds in the same order.
template <typename T, int M, bool kCols>
ADRT_HD void fwd_radix4_from_image_interior(T *buf, const T *img, const TileCtx &c, int tid, int dbase)
{
    constexpr int W = VecOf<T>::L, G = Geo<M>::G, NT = Geo<M>::NT, P = Pitch<T>::value;
    constexpr int NB = G / 4, NCH = XW / W, JOBS = NB * NCH, ROUNDS = JOBS / NT;
    constexpr int NP = (2 * W + 2) / W;
    constexpr int NRAW = kCols ? (W + 3) * (4 / W) : 4 * NP;
    static_assert(JOBS % NT == 0 && ROUNDS % 2 == 0, "rounds are paired");
    const int n = c.n;
#pragma unroll
    for (int rd = 0; rd < ROUNDS; rd += 2) {
        Pack<T> raw[2][NRAW];
        int k0s[2], xs[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int job = tid + (rd + u) * NT;
            int k0, chunk;
            if (kCols) {
                const int rest = job >> 3;
                k0 = 2 * (rest % (NB / 2)) + (job & 1);
                chunk = 4 * (rest / (NB / 2)) + ((job >> 1) & 3);
            } else {
                k0 = job / NCH;
                chunk = job % NCH;
            }
            k0s[u] = k0;
            xs[u] = chunk * W;
            // jobs below the step's lo_out store nothing; they fetch the window of x = 4 so that no address leaves the image
            const int x = xs[u] < 4 ? 4 : xs[u];
            const int r0 = c.g * G + 4 * k0;
            if (kCols) {
                const int dlo = dbase + x - 3;
#pragma unroll
                for (int i = 0; i < W + 3; ++i) {
                    const int d = dlo + i;
                    const T *rp = img + (long long)(c.q == 1 ? n - 1 - d : d) * n + r0;
#pragma unroll
                    for (int k = 0; k < 4 / W; ++k) raw[u][i * (4 / W) + k] = *reinterpret_cast<const Pack<T> *>(rp + k * W);
                }
            } else {
                const int col_lo = n - dbase - x - W;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const T *rp = img + (long long)(c.q == 0 ? r0 + j : n - 1 - (r0 + j)) * n + col_lo;
#pragma unroll
                    for (int k = 0; k < NP; ++k) raw[u][j * NP + k] = *reinterpret_cast<const Pack<T> *>(rp + k * W);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (xs[u] < 4) continue;
            T w[4][W + 3];                       // w[j][i] = parent j at offset dlo + i
            if (kCols) {
#pragma unroll
                for (int i = 0; i < W + 3; ++i)
#pragma unroll
                    for (int k = 0; k < 4 / W; ++k)
#pragma unroll
                        for (int q = 0; q < W; ++q) w[k * W + q][i] = raw[u][i * (4 / W) + k].v[q];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < W + 3; ++i) w[j][i] = raw[u][j * NP + (W + 2 - i) / W].v[(W + 2 - i) % W];
            }
            T y0[W], y1[W + 1], y2[W + 2], y3[W + 3];
#pragma unroll
            for (int i = 0; i < W; ++i) y0[i] = w[0][i + 3];
#pragma unroll
            for (int i = 0; i < W + 1; ++i) y1[i] = w[1][i + 2];
#pragma unroll
            for (int i = 0; i < W + 2; ++i) y2[i] = w[2][i + 1];
#pragma unroll
            for (int i = 0; i < W + 3; ++i) y3[i] = w[3][i];
            T o[4 * W];
            fwd_radix4_math<T>(y0, y1, y2, y3, o);
            T *orow = buf + (4 * k0s[u]) * P + xs[u];
#pragma unroll
            for (int p4 = 0; p4 < 4; ++p4) store_cv<T>(orow + p4 * P, &o[p4 * W]);
        }
    }
}

template <typename T, int M, int LH>
ADRT_HD void fwd_radix4_from_image(T *buf, const T *img, const TileCtx &c, int tid)
{
    constexpr int W = VecOf<T>::L, G = Geo<M>::G, NT = Geo<M>::NT, P = Pitch<T>::value;
    constexpr int NB = G / 4, NCH = XW / W, JOBS = NB * NCH, ROUNDS = (JOBS + NT - 1) / NT;
    static_assert(NB >= 2 && NB % 2 == 0, "lanes pair the butterflies");
    const int n = c.n;
    const int dbase = c.d0 - LH;
    const bool cols = (c.q == 1 || c.q == 2);
    if constexpr (sizeof(T) == 8 && JOBS % NT == 0 && ROUNDS % 2 == 0) {
        if (dbase >= 0 && dbase + XW <= n) {
            if (cols) fwd_radix4_from_image_interior<T, M, true>(buf, img, c, tid, dbase);
            else fwd_radix4_from_image_interior<T, M, false>(buf, img, c, tid, dbase);
            return;
        }
    }
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
        const int job = tid + rd * NT;
        if (job >= JOBS) continue;
        int k0, chunk;
        if (cols) {
            const int rest = job >> 3;
            k0 = 2 * (rest % (NB / 2)) + (job & 1);
            chunk = 4 * (rest / (NB / 2)) + ((job >> 1) & 3);
        } else {
            k0 = job / NCH;
            chunk = job % NCH;
        }
        const int x = chunk * W;
        if (x < 4) continue;                 // lo_out of the step: windows would start below the tile
        const int r0 = c.g * G + 4 * k0;     // oriented image row of parent 0
        const int dlo = dbase + x - 3;       // offset of window element 0
        T w[4][W + 3];                       // w[j][i] = parent j at offset dlo + i
        if (cols && dlo >= 0 && dlo + W + 3 <= n) {
#pragma unroll
            for (int i = 0; i < W + 3; ++i) {
                const int d = dlo + i;
                const T *rp = img + (long long)(c.q == 1 ? n - 1 - d : d) * n + r0;
#pragma unroll
                for (int k = 0; k < 4 / W; ++k) {
                    const Pack<T> v = *reinterpret_cast<const Pack<T> *>(rp + k * W);
#pragma unroll
                    for (int q = 0; q < W; ++q) w[k * W + q][i] = v.v[q];
                }
            }
        } else if (!cols && dbase + x >= 4 && dbase + x + W <= n) {
            // offsets dlo .. dlo+W+2 are the columns col_lo+W+2 .. col_lo of the image row
            constexpr int NP = (2 * W + 2) / W;
            const int col_lo = n - dbase - x - W;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const T *rp = img + (long long)(c.q == 0 ? r0 + j : n - 1 - (r0 + j)) * n + col_lo;
                Pack<T> v[NP];
#pragma unroll
                for (int k = 0; k < NP; ++k) v[k] = *reinterpret_cast<const Pack<T> *>(rp + k * W);
#pragma unroll
                for (int i = 0; i < W + 3; ++i) w[j][i] = v[(W + 2 - i) / W].v[(W + 2 - i) % W];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + j;
#pragma unroll
                for (int i = 0; i < W + 3; ++i) {
                    const int d = dlo + i;
                    T v;
                    if (d < 0) v = T(-0.0);
                    else if (d >= n) v = T(0.0);
                    else if (c.q == 0) v = img[(long long)r * n + (n - 1 - d)];
                    else if (c.q == 1) v = img[(long long)(n - 1 - d) * n + r];
                    else if (c.q == 2) v = img[(long long)d * n + r];
                    else v = img[(long long)(n - 1 - r) * n + (n - 1 - d)];
                    w[j][i] = v;
                }
            }
        }
        T y0[W], y1[W + 1], y2[W + 2], y3[W + 3];
#pragma unroll
        for (int i = 0; i < W; ++i) y0[i] = w[0][i + 3];
#pragma unroll
        for (int i = 0; i < W + 1; ++i) y1[i] = w[1][i + 2];
#pragma unroll
        for (int i = 0; i < W + 2; ++i) y2[i] = w[2][i + 1];
#pragma unroll
        for (int i = 0; i < W + 3; ++i) y3[i] = w[3][i];
        T o[4 * W];
        fwd_radix4_math<T>(y0, y1, y2, y3, o);
        T *orow = buf + (4 * k0) * P + x;
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) store_cv<T>(orow + p4 * P, &o[p4 * W]);
    }
}

// Number of compute steps for M stages (radix-4 first, one radix-2 at the end
// when M is odd).  Each step is two barrier-separated phases (compute, store).
ADRT_HD constexpr int num_steps(int M) { return (M + 1) / 2; }

// step starting at local stage t: radix-4 if two stages are left, else radix-2
template <typename T, int M, int t>
ADRT_HD void fwd_step_compute(const T *buf, int tid, T (&o)[NREG])
{
    if constexpr (t + 2 <= M) fwd_radix4_compute<T, M, t>(buf, tid, o);
    else fwd_radix2_compute<T, M, t>(buf, tid, o);
}

template <typename T, int M, int t>
ADRT_HD void fwd_step_store(T *buf, int tid, const T (&o)[NREG])
{
    if constexpr (t + 2 <= M) fwd_radix4_store<T, M, t>(buf, tid, o);
    else fwd_radix2_store<T, M, t>(buf, tid, o);
}

// Loader that also performs local stage 0 (odd M on the workspace side): the two
// input rows 2k, 2k+1 of a pair are read straight from global memory and the
// tile receives   row 2k   = in_2k[x] + in_2k+1[x]      (even angle, shift 0)
//                 row 2k+1 = in_2k[x] + in_2k+1[x - 1]   (odd angle, shift 1)
// so the radix-2 step and its shared-memory round trip disappear.  Scalar,
// coalesced accesses: the rows' different alignments do not matter here.
template <typename T, int M, int LH>
ADRT_HD void fwd_load_wrows_stage0(T *buf, const T *src_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    const int warp = tid >> 5, lane = tid & 31;
    const int sup = c.n + c.a_g;
    for (int k = warp; k < G / 2; k += NWARP) {
        const int jA = 2 * k, jB = 2 * k + 1;
        const T *rowA = src_plane + ((long long)(c.k0 * G + jA) * c.e + c.a_g) * c.in_pitch + fwd_row_skew(c.a_g, jA);
        const T *rowB = src_plane + ((long long)(c.k0 * G + jB) * c.e + c.a_g) * c.in_pitch + fwd_row_skew(c.a_g, jB);
        const int dA = c.d0 - LH - c.a_g * jA, dB = c.d0 - LH - c.a_g * jB;   // offsets of tile position 0
        T *oe = buf + jA * P, *oo = buf + jB * P;
        if (dB >= 4 && dA + XW <= sup) {   // dB < dA: the whole pair (and x - 4) is inside [0, sup)
            // rowX + dX is 16-byte aligned (storage skew): vector loads, 4 offsets per lane
#pragma unroll
            for (int cc = 0; cc < XW / (32 * V); ++cc) {
                const int x = (cc * 32 + lane) * V;
                T a0[V], b[V + 1];
                load_window<T, V, 0>(rowA + dA + x, a0);
                load_window<T, V + 1, V - 1>(rowB + dB + x - V, b);   // offsets x-1 .. x+3
                T ve[V], vo[V];
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    ve[i] = a0[i] + b[i + 1];
                    vo[i] = a0[i] + b[i];
                }
                store_chunk<T>(oe + x, ve);
                store_chunk<T>(oo + x, vo);
            }
        } else if (dB - 1 >= sup || dA + XW <= 0) {
            // every operand of the pair is a structural +0.0 (above the support; 0 + 0 = +0.0) or a -0.0 sentinel
            // (below offset 0; -0 + -0 = -0.0): constant rows, no loads (see fwd_load_wrows)
            constexpr int L = VecOf<T>::L;
            Pack<T> z;
#pragma unroll
            for (int i = 0; i < L; ++i) z.v[i] = dB - 1 >= sup ? T(0.0) : T(-0.0);
#pragma unroll
            for (int cc = 0; cc < XW / (32 * L); ++cc) {
                *reinterpret_cast<Pack<T> *>(oe + (cc * 32 + lane) * L) = z;
                *reinterpret_cast<Pack<T> *>(oo + (cc * 32 + lane) * L) = z;
            }
        } else {
#pragma unroll 2
            for (int i = 0; i < XW / 32; ++i) {
                const int x = i * 32 + lane;
                const int da = dA + x, db = dB + x;
                const T a0 = da < 0 ? T(-0.0) : (da < sup ? rowA[da] : T(0.0));
                const T b0 = db < 0 ? T(-0.0) : (db < sup ? rowB[db] : T(0.0));
                const T b1 = db - 1 < 0 ? T(-0.0) : (db - 1 < sup ? rowB[db - 1] : T(0.0));
                oe[x] = a0 + b0;
                oo[x] = a0 + b1;
            }
        }
    }
}

// ---- stores -----------------------------------------------------------------------
// R-layout workspace: output row p -> row (g*G + p), stored with the skew the next
// pass wants.  The tile owns the aligned global chunks [d0, d0 + TD) of every row;
// chunk position gp holds offsets gp - s .., read from the tile at LH + (gp - d0) - s.
// Rows are zero above their support bound n + a' and are written up to the pitch.
template <typename T, int M, int LH, int TD, int S>
ADRT_HD void fwd_store_row(const T *b, T *row, const TileCtx &c, int lim, bool zero, int lane)
{
    // a lane moves ONE 16-byte vector per trip (4 floats / 2 doubles), so consecutive lanes touch consecutive
    // 16-byte slots of the tile row and of the global row: with 4 doubles per lane (32-byte lane stride) the
    // window loads paid 2.2 x their ideal shared-memory wavefronts (profiles/r06_pass_phases_f64.txt, F1)
    constexpr int CW = VecOf<T>::L;
    constexpr int Q = ((CW - S) % CW + CW) % CW;
#pragma unroll
    for (int k = 0; k < (TD / CW + 31) / 32; ++k) {
        const int xc = (k * 32 + lane) * CW;     // gp - d0
        const int gp = c.d0 + xc;
        if (xc < TD && gp < c.out_pitch) {
            T v[CW];
            if (zero) {
#pragma unroll
                for (int i = 0; i < CW; ++i) v[i] = T(0.0);
            } else {
                load_window<T, CW, Q>(b + (LH + xc - S - Q), v);
#pragma unroll
                for (int i = 0; i < CW; ++i)
                    if (gp - S + i >= lim) v[i] = T(0.0);
            }
            store_cv<T>(row + gp, v);
        }
    }
}

template <typename T, int M, int LH, int TD>
ADRT_HD void fwd_store_wrows(const T *buf, T *dst_plane, const TileCtx &c, bool zero, int tid)
{
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    const int warp = tid >> 5, lane = tid & 31;
    for (int p = warp; p < G; p += NWARP) {
        T *row = dst_plane + ((long long)c.g * G + p) * c.out_pitch;
        const int ang = c.a_g * G + p;
        int lim = c.n + ang;
        if (lim > c.D) lim = c.D;
        const int s = c.next_g ? fwd_row_skew(ang, c.k0 & (c.next_g - 1)) : 0;
        const T *b = buf + p * P;
        switch (s) {
        case 0: fwd_store_row<T, M, LH, TD, 0>(b, row, c, lim, zero, lane); break;
        case 1: fwd_store_row<T, M, LH, TD, 1>(b, row, c, lim, zero, lane); break;
        case 2: fwd_store_row<T, M, LH, TD, 2>(b, row, c, lim, zero, lane); break;
        default: fwd_store_row<T, M, LH, TD, 3>(b, row, c, lim, zero, lane); break;
        }
    }
}

// Public layout (D, n) of the plane: column g*G + p, all offsets < D.  Lanes walk
// the rows p (consecutive columns), each thread moves V consecutive offsets.
template <typename T, int M, int TD>
ADRT_HD void store_qcols(const T *buf, T *dst_plane, const TileCtx &c, int xoff, bool zero, int tid)
{
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    constexpr int NIT = (TD + NWARP * V - 1) / (NWARP * V);
    const int warp = tid >> 5, lane = tid & 31;
    const int cols = G < c.n ? G : c.n;
    const bool fast = c.d0 + TD <= c.D;
    const long long n1 = c.n;
    for (int p = lane; p < cols; p += 32) {
        T *o = dst_plane + (long long)(c.d0 + warp * V) * n1 + c.g * G + p;   // row d0 + warp*V, column p
        const T *b = buf + p * P + xoff + warp * V;
        const long long ostep = (long long)NWARP * V * n1;
#pragma unroll 3
        for (int it = 0; it < NIT; ++it, o += ostep, b += NWARP * V) {
            const int xc = (it * NWARP + warp) * V;
            if (xc >= TD) break;
            T v[V];
            if (zero) {
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = T(0.0);
            } else {
                load_window<T, V, 0>(b, v);
            }
            if (fast) {
                o[0] = v[0];
                o[n1] = v[1];
                o[2 * n1] = v[2];
                o[3 * n1] = v[3];
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (c.d0 + xc + i < c.D) o[i * n1] = v[i];
            }
        }
    }
}

// ===========================================================================
// transposed (bdrt)
// ===========================================================================
// Input row p holds in_p[d0 + x].  A row written by a step is stored with a
// skew: logical position = stored position - skew, skew = j * angle of the row
// in the step that produced it (0 for loaded rows and after the last step).

template <typename T, int M>
ADRT_HD void bwd_load_wrows(T *buf, const T *src_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value, L = VecOf<T>::L;
    constexpr int RPW = (G + NWARP - 1) / NWARP;       // rows per warp
    constexpr int NV = XW / (32 * L);                  // vectors per lane and row
    const int warp = tid >> 5, lane = tid & 31;
    if (c.d0 + XW <= c.D) {
        // d0 and the pitch are multiples of 4: aligned vector copy, pointers advanced by constants
        const T *rp = src_plane + ((long long)c.g * G + warp) * c.in_pitch + c.d0 + lane * L;
        T *dp = buf + warp * P + lane * L;
        const long long rstep = (long long)NWARP * c.in_pitch;
#pragma unroll
        for (int r0 = 0; r0 < RPW; r0 += 2) {
            Pack<T> v[2][NV];
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (r0 + u < RPW && warp + (r0 + u) * NWARP < G) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) v[u][k] = *reinterpret_cast<const Pack<T> *>(rp + (r0 + u) * rstep + k * 32 * L);
                }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (r0 + u < RPW && warp + (r0 + u) * NWARP < G) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) *reinterpret_cast<Pack<T> *>(dp + (r0 + u) * NWARP * P + k * 32 * L) = v[u][k];
                }
        }
    } else {
        for (int p = warp; p < G; p += NWARP) {
            const T *row = src_plane + ((long long)c.g * G + p) * c.in_pitch;
            T *dst = buf + p * P;
#pragma unroll 4
            for (int k = 0; k < XW / 32; ++k) {
                const int x = k * 32 + lane, d = c.d0 + x;
                dst[x] = d < c.D ? row[d] : T(0.0);
            }
        }
    }
}

template <typename T, int M>
ADRT_HD void bwd_load_qcols(T *buf, const T *src_plane, const TileCtx &c, int tid)
{
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    constexpr int NIT = XW / (NWARP * V);
    const int warp = tid >> 5, lane = tid & 31;
    const int cols = G < c.n ? G : c.n;
    const bool fast = c.d0 + XW <= c.D;
    const long long n1 = c.n;
    for (int p = lane; p < cols; p += 32) {
        const T *ir = src_plane + (long long)(c.d0 + warp * V) * n1 + c.g * G + p;
        T *b = buf + p * P + warp * V;
        const long long istep = (long long)NWARP * V * n1;
        if (fast) {
#pragma unroll 1
            for (int it0 = 0; it0 < NIT; it0 += 2, ir += 2 * istep, b += 2 * NWARP * V) {
                // two chunks (8 independent loads) in flight per thread
                T v0[V], v1[V];
                v0[0] = ir[0]; v0[1] = ir[n1]; v0[2] = ir[2 * n1]; v0[3] = ir[3 * n1];
                const T *ir1 = ir + istep;
                v1[0] = ir1[0]; v1[1] = ir1[n1]; v1[2] = ir1[2 * n1]; v1[3] = ir1[3 * n1];
                store_chunk<T>(b, v0);
                store_chunk<T>(b + NWARP * V, v1);
            }
        } else {
#pragma unroll 2
            for (int it = 0; it < NIT; ++it) {
                const int x = (it * NWARP + warp) * V;
                T v[V];
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = (c.d0 + x + i < c.D) ? ir[it * istep + i * n1] : T(0.0);
                store_chunk<T>(b + it * NWARP * V, v);
            }
        }
    }
}

// Missing-operand rule of bdrt_core (adrt_cdefs_bdrt.hpp:96-109) for a value read
// at logical tile position `pos` from a row whose intermediate ends at `lim`.
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
// The thread reads the four parent windows at logical positions [x, x+4+p) and
// produces child j at logical positions x - j*a + [0,4), stored at x (skew j*a).
// Parent row p carries skew jp*(4a+p) (jp = index of the parent block in the
// previous step, 0 if the rows were loaded); JP = jp & 3 fixes the residues.
// `dt` = D - d0 and `ag` = global base angle give each row's logical end:
//   parent rows (block k0 at stage t+2): dt + ag*k0*4e;  node k=1: + ag*2e.
// the arithmetic of one chunk on parent windows that are already in registers
template <typename T, bool kMask>
ADRT_HD void bwd_radix4_math(T (&g0)[VecOf<T>::L], T (&g1)[VecOf<T>::L + 1], T (&g2)[VecOf<T>::L + 2], T (&g3)[VecOf<T>::L + 3],
                             int a, int x, int lim_p, int lim_1, T *o)
{
    constexpr int V = VecOf<T>::L;
    if (kMask) {
#pragma unroll
        for (int i = 0; i < V; ++i) g0[i] = bmask<T, kMask>(g0[i], x + i, lim_p, false);
#pragma unroll
        for (int i = 0; i < V + 1; ++i) g1[i] = bmask<T, kMask>(g1[i], x + i, lim_p, true);
#pragma unroll
        for (int i = 0; i < V + 2; ++i) g2[i] = bmask<T, kMask>(g2[i], x + i, lim_p, false);
#pragma unroll
        for (int i = 0; i < V + 3; ++i) g3[i] = bmask<T, kMask>(g3[i], x + i, lim_p, true);
    }
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
#pragma unroll
    for (int i = 0; i < V; ++i) {
        o[0 * V + i] = u00[i] + u01[i];
        o[1 * V + i] = u00[i] + u01[i + 1];
        o[2 * V + i] = u10[i] + u11[i];
        o[3 * V + i] = u10[i] + u11[i + 1];
    }
}

template <typename T, bool kMask, int JP>
ADRT_HD void bwd_radix4_chunk(const T *ip, long long P, int a, int x, int jp, int lim_p, int lim_1, T *o)
{
    // P = distance between the four parent rows (tile pitch, or the workspace pitch
    // when the parents are read straight from global memory)
    constexpr int V = VecOf<T>::L;
    constexpr int L = VecOf<T>::L;
    constexpr int Q1 = (JP * 1) % L, Q2 = (JP * 2) % L, Q3 = (JP * 3) % L;
    const int s0 = jp * 4 * a;  // skew of parent 0; parent p adds jp*p
    T g0[V], g1[V + 1], g2[V + 2], g3[V + 3];
    load_window<T, V, 0>(ip + x + s0, g0);
    load_window<T, V + 1, Q1>(ip + P + (x + s0 + jp - Q1), g1);
    load_window<T, V + 2, Q2>(ip + 2 * P + (x + s0 + 2 * jp - Q2), g2);
    load_window<T, V + 3, Q3>(ip + 3 * P + (x + s0 + 3 * jp - Q3), g3);
    bwd_radix4_math<T, kMask>(g0, g1, g2, g3, a, x, lim_p, lim_1, o);
}

// every window (and its over-fetch) must stay inside the row; chunks that fail
// this only produce positions beyond the valid region
template <typename T>
ADRT_HD bool bwd_chunk_ok(int x, int jp, int a)
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    return x + jp * (4 * a + 3) + V + 3 + 3 <= Pitch<T>::value;
}

template <typename T, bool kMask, int JP>
ADRT_HD void bwd_radix4_group(const T *buf, int e, int k0, int a, int lane, int jp, int dt, int ag, T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int P = Pitch<T>::value;
    const T *ip = buf + (k0 * 4 * e + 4 * a) * P;
    const int lim_p = dt + ag * (k0 * 4 * e), lim_1 = lim_p + ag * 2 * e;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int x = (lane + 32 * c) * V;
        if (bwd_chunk_ok<T>(x, jp, a)) bwd_radix4_chunk<T, kMask, JP>(ip, P, a, x, jp, lim_p, lim_1, &o[c * 4 * V]);
    }
}

// rprev = radix of the step that produced the parent rows (0: they were loaded)
template <typename T, int M, bool kMask, int t, int rprev>
ADRT_HD void bwd_radix4_compute(const T *buf, int dt, int ag, int tid, T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int G = Geo<M>::G;
    constexpr int e = 1 << t;
    const int gi = tid >> 5, lane = tid & 31;
    if (gi >= G / 4) return;
    const int k0 = gi >> t, a = gi & (e - 1);
    const int jp = rprev ? (k0 & (rprev - 1)) : 0;
    switch (jp) {
    case 0: bwd_radix4_group<T, kMask, 0>(buf, e, k0, a, lane, jp, dt, ag, o); break;
    case 1: bwd_radix4_group<T, kMask, 1>(buf, e, k0, a, lane, jp, dt, ag, o); break;
    case 2: bwd_radix4_group<T, kMask, 2>(buf, e, k0, a, lane, jp, dt, ag, o); break;
    default: bwd_radix4_group<T, kMask, 3>(buf, e, k0, a, lane, jp, dt, ag, o); break;
    }
}

// First transposed step of an interior tile whose parents are workspace rows: the
// operand windows are read straight from global memory (aligned: d0 and the pitch
// are multiples of 4), so the tile is never staged in shared memory before step 0.
template <typename T, int M, int t>
ADRT_HD void bwd_radix4_compute_global(const T *src_plane, const TileCtx &c, int tid, T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);
    constexpr int G = Geo<M>::G;
    constexpr int e = 1 << t;
    const int gi = tid >> 5, lane = tid & 31;
    if (gi >= G / 4) return;
    const int k0 = gi >> t, a = gi & (e - 1);
    const T *ip = src_plane + ((long long)c.g * G + k0 * 4 * e + 4 * a) * c.in_pitch + c.d0;
#pragma unroll
    for (int cc = 0; cc < CHUNKS; ++cc) {
        const int x = (lane + 32 * cc) * V;
        // doubles: x + V + 6 <= pitch for every chunk, and without the test the chunks form one basic block whose
        // global loads the compiler may issue ahead of the previous chunk's adds
        if (XW + 6 <= Pitch<T>::value || bwd_chunk_ok<T>(x, 0, a))
            bwd_radix4_chunk<T, false, 0>(ip, c.in_pitch, a, x, 0, 0, 0, &o[cc * 4 * V]);
    }
}

// First transposed step of a pass that reads the public (offset, column) layout: the four
// parents of butterfly `a` are the four ADJACENT sinogram columns g*G + 4a .. 4a+3, i.e. one
// 16-byte (fp32) / 32-byte (fp64) piece of every offset row, so the thread that fetches them
// can run the radix-4 step on the spot and store the four children -- the tile is never staged
// as loaded, which saves one tile write and the step's over-wide window reads (1.9 tile reads)
// of shared-memory traffic.  Lanes walk the butterflies first (NB * 16 contiguous bytes per
// offset row), then the chunks.  Children are stored exactly where bwd_radix4_store puts them.
template <typename T, int M, bool kMask, int t, bool kSub = false>
ADRT_HD void bwd_radix4_from_qcols(T *buf, const T *src_plane, const TileCtx &c, int tid)
{
    constexpr int W = VecOf<T>::L, G = Geo<M>::G, NT = Geo<M>::NT, P = Pitch<T>::value;
    constexpr int e = 1 << t;
    static_assert(4 * e == G, "the first step's butterflies span the whole group");
    constexpr int NB = G / 4, SLOTS = NT / NB, NCH = XW / W, CPT = (NCH + SLOTS - 1) / SLOTS;
    const int a = tid % NB, slot = tid / NB;
    const long long n1 = c.n;
    const int dt = c.D - c.d0;
    const int lim_p = dt, lim_1 = dt + c.a_g * 2 * e;
    const T *col = src_plane + (long long)c.d0 * n1 + c.g * G + 4 * a;
    T *orow = buf + a * P;
    // fp64 interior tiles: the pieces of TWO chunks are fetched before either is used (2 x 10 16-byte loads in
    // flight per thread instead of 10, two dependent rounds per tile instead of four: 53 % of the five-stage fp64
    // kernel's stall samples sat in this loader with 2 x 256 threads per SM, profiles/r06_pass_phases_f64.txt);
    // every chunk of a double tile is live (NCH = SLOTS * CPT, XW + 6 <= pitch), so the pairs need no tests.
    // -DADRT_QCOLS_PAIR_F32=1 (A/B builds, tools/build_variant.sh) pairs the two chunks of the fp32 pass as well.
    if constexpr ((sizeof(T) == 8 || ADRT_QCOLS_PAIR_F32) && !kMask && !kSub && NCH % SLOTS == 0 && CPT % 2 == 0) {
#pragma unroll
        for (int cc = 0; cc < CPT; cc += 2) {
            Pack<T> raw[2][W + 3][4 / W];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int x = (slot + (cc + u) * SLOTS) * W;
#pragma unroll
                for (int i = 0; i < W + 3; ++i)
#pragma unroll
                    for (int k = 0; k < 4 / W; ++k)
                        raw[u][i][k] = *reinterpret_cast<const Pack<T> *>(col + (long long)(x + i) * n1 + k * W);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int x = (slot + (cc + u) * SLOTS) * W;
                // floats: the last chunk of a row fails bwd_chunk_ok (its outputs lie beyond the valid region); its
                // loads above are harmless (an unmasked tile ends 8 offsets before D)
                if (XW + 6 > P && !bwd_chunk_ok<T>(x, 0, a)) continue;
                T g0[W], g1[W + 1], g2[W + 2], g3[W + 3];
#pragma unroll
                for (int i = 0; i < W + 3; ++i) {
                    if (i < W) g0[i] = raw[u][i][0].v[0];
                    if (i < W + 1) g1[i] = raw[u][i][1 / W].v[1 % W];
                    if (i < W + 2) g2[i] = raw[u][i][2 / W].v[2 % W];
                    g3[i] = raw[u][i][3 / W].v[3 % W];
                }
                T o[4 * W];
                bwd_radix4_math<T, false>(g0, g1, g2, g3, a, x, lim_p, lim_1, o);
#pragma unroll
                for (int j = 0; j < 4; ++j) store_cv<T>(orow + j * e * P + x, &o[j * W]);   // skew j*a
            }
        }
        return;
    }
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
        const int x = (slot + cc * SLOTS) * W;
        if (x >= XW || !bwd_chunk_ok<T>(x, 0, a)) continue;
        T g0[W], g1[W + 1], g2[W + 2], g3[W + 3];
#pragma unroll
        for (int i = 0; i < W + 3; ++i) {
            Pack<T> v[4 / W];
            if (!kMask || x + i < dt) {
#pragma unroll
                for (int k = 0; k < 4 / W; ++k) {
                    const T *pv = col + (long long)(x + i) * n1 + k * W;
                    v[k] = *reinterpret_cast<const Pack<T> *>(pv);
                    if constexpr (kSub) {
                        // element-wise "a - b" exactly as the separate subtraction kernel computes it
                        const Pack<T> w = *reinterpret_cast<const Pack<T> *>(reinterpret_cast<const char *>(pv) + c.sub_delta);
#pragma unroll
                        for (int q = 0; q < W; ++q) v[k].v[q] = v[k].v[q] - w.v[q];
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4 / W; ++k)
#pragma unroll
                    for (int q = 0; q < W; ++q) v[k].v[q] = T(0.0);
            }
            if (i < W) g0[i] = v[0].v[0];
            if (i < W + 1) g1[i] = v[1 / W].v[1 % W];
            if (i < W + 2) g2[i] = v[2 / W].v[2 % W];
            g3[i] = v[3 / W].v[3 % W];
        }
        T o[4 * W];
        bwd_radix4_math<T, kMask>(g0, g1, g2, g3, a, x, lim_p, lim_1, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) store_cv<T>(orow + j * e * P + x, &o[j * W]);   // skew j*a
    }
}

template <typename T, int M, int t, int rprev>
ADRT_HD void bwd_radix4_store(T *buf, int tid, const T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int G = Geo<M>::G, P = Pitch<T>::value;
    constexpr int e = 1 << t;
    const int gi = tid >> 5, lane = tid & 31;
    if (gi >= G / 4) return;
    const int k0 = gi >> t, a = gi & (e - 1);
    const int jp = rprev ? (k0 & (rprev - 1)) : 0;
    T *orow = buf + ((k0 * 4) * e + a) * P;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int x = (lane + 32 * c) * V;
        if (bwd_chunk_ok<T>(x, jp, a)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) store_cv<T>(orow + j * e * P + x, &o[(c * 4 + j) * V]);   // skew j*a
        }
    }
}

// ---- transposed radix-2 step (always the first transposed step: parents loaded) --
template <typename T, int M, bool kMask, int t>
ADRT_HD void bwd_radix2_compute(const T *buf, int dt, int ag, int tid, T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    constexpr int e = 1 << t;
    const int lane = tid & 31;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int gi = (tid >> 5) + u * NWARP;
        if (gi >= G / 2) continue;
        const int k = gi >> t, b = gi & (e - 1);
        const T *ip = buf + (k * 2 * e + 2 * b) * P;
        const int lim_p = dt + ag * (k * 2 * e);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int x = (lane + 32 * c) * V;
            T g0[V], g1[V + 1];
            load_window<T, V, 0>(ip + x, g0);
            load_window<T, V + 1, 0>(ip + P + x, g1);
            if (kMask) {
#pragma unroll
                for (int i = 0; i < V; ++i) g0[i] = bmask<T, kMask>(g0[i], x + i, lim_p, false);
#pragma unroll
                for (int i = 0; i < V + 1; ++i) g1[i] = bmask<T, kMask>(g1[i], x + i, lim_p, true);
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {
                o[((u * CHUNKS + c) * 2 + 0) * V + i] = g0[i] + g1[i];
                o[((u * CHUNKS + c) * 2 + 1) * V + i] = g0[i] + g1[i + 1];
            }
        }
    }
}

template <typename T, int M, int t>
ADRT_HD void bwd_radix2_store(T *buf, int tid, const T (&o)[NREG])
{
    constexpr int V = VecOf<T>::L, CHUNKS = XW / (32 * V);  // chunk = one 16-byte vector
    (void)CHUNKS;
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    constexpr int e = 1 << t;
    const int lane = tid & 31;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int gi = (tid >> 5) + u * NWARP;
        if (gi >= G / 2) continue;
        const int k = gi >> t, b = gi & (e - 1);
        T *oA = buf + ((2 * k) * e + b) * P;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int x = (lane + 32 * c) * V;
            store_cv<T>(oA + x, &o[((u * CHUNKS + c) * 2 + 0) * V]);
            store_cv<T>(oA + e * P + x, &o[((u * CHUNKS + c) * 2 + 1) * V]);   // skew b
        }
    }
}

// Transposed step that produces the rows of local stage t (radix-4 if stage t+2
// exists, else radix-2).  RPREV = radix of the step that produced its parent rows
// (0: they were loaded).
template <typename T, int M, bool kMask, int t, int RPREV>
ADRT_HD void bwd_step_compute(const T *buf, int dt, int ag, int tid, T (&o)[NREG])
{
    if constexpr (t + 2 <= M) bwd_radix4_compute<T, M, kMask, t, RPREV>(buf, dt, ag, tid, o);
    else bwd_radix2_compute<T, M, kMask, t>(buf, dt, ag, tid, o);
}

template <typename T, int M, int t, int RPREV>
ADRT_HD void bwd_step_store(T *buf, int tid, const T (&o)[NREG])
{
    if constexpr (t + 2 <= M) bwd_radix4_store<T, M, t, RPREV>(buf, tid, o);
    else bwd_radix2_store<T, M, t>(buf, tid, o);
}

// Store that also performs the transposed local stage 0 (odd M on the workspace
// side): output row j = 2k (+1) is  g0[x] + g1[x (+1)]  of the parent pair (rows 2k,
// 2k+1 of the tile, angles 0 / 1), written to workspace row (k0*G + j)*e + a_g at
// offset d0 + x - a_g*j.  The odd parent carries the skew (k & 3) of the radix-4
// step that produced it (kPrevR4).  Scalar, coalesced accesses.
template <typename T, int M, int TD, bool kMask, bool kPrevR4>
ADRT_HD void bwd_store_wrows_stage0(const T *buf, T *dst_plane, const TileCtx &c, bool zero, int tid)
{
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    constexpr int NS = (TD + 31) / 32;
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < G; j += NWARP) {
        const int k = j >> 1, odd = j & 1;
        const T *g0 = buf + (2 * k) * P;
        const T *g1 = buf + (2 * k + 1) * P + (kPrevR4 ? (k & 3) : 0) + odd;
        const int lim_p = (c.D - c.d0) + c.a_g * (2 * k);
        T *row = dst_plane + ((long long)(c.k0 * G + j) * c.e + c.a_g) * c.out_pitch;
        const int dbase = c.d0 - c.a_g * j;
        if (!kMask && !zero && dbase >= 0 && dbase + TD <= c.D) {
            // whole row segment in range: immediate offsets, 2 LDS + 1 add + 1 STG per element
            const T *p0 = g0 + lane, *p1 = g1 + lane;
            T *r = row + dbase + lane;
#pragma unroll
            for (int i = 0; i < TD / 32; ++i) r[i * 32] = p0[i * 32] + p1[i * 32];
            if ((TD % 32) != 0 && lane < TD % 32) r[(TD / 32) * 32] = p0[(TD / 32) * 32] + p1[(TD / 32) * 32];
            continue;
        }
#pragma unroll 2
        for (int i = 0; i < NS; ++i) {
            const int xc = i * 32 + lane, d = dbase + xc;
            if (xc < TD && d >= 0 && d < c.D) {
                if (zero) {
                    row[d] = T(0.0);   // tile beyond row D of the input: every surviving output is +0
                } else {
                    const T a = bmask<T, kMask>(g0[xc], xc, lim_p, false);
                    const T b = bmask<T, kMask>(g1[xc], xc + odd, lim_p, true);
                    row[d] = a + b;
                }
            }
        }
    }
}

// Output row j -> workspace row (k0*G + j)*e + a_g; tile position xc is offset
// d0 - a_g*j + xc.  The tile owns the aligned chunks [ceil4(dbase), ceil4(dbase) + TD)
// of the row (valid tile positions reach TD + 3); Q = (a_g*j) & 3 is the residue of
// the tile-side window.
template <typename T, int M, int TD, int Q>
ADRT_HD void bwd_store_row(const T *b, T *row, const TileCtx &c, int dbase, bool zero, int lane)
{
#pragma unroll
    for (int k = 0; k < (TD / V + 31) / 32; ++k) {
        const int xa = (k * 32 + lane) * V;       // aligned tile position below the window
        const int gp = dbase + Q + xa;            // aligned global offset of the chunk
        if (xa < TD && gp + V > 0 && gp < c.D) {
            T v[V];
            if (zero) {
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = T(0.0);
            } else {
                load_window<T, V, Q>(b + xa, v);
            }
            if (gp >= 0 && gp + V <= c.D) {
                store_chunk<T>(row + gp, v);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (gp + i >= 0 && gp + i < c.D) row[gp + i] = v[i];
            }
        }
    }
}

template <typename T, int M, int TD>
ADRT_HD void bwd_store_wrows(const T *buf, T *dst_plane, const TileCtx &c, bool zero, int tid)
{
    constexpr int G = Geo<M>::G, NWARP = Geo<M>::NWARP, P = Pitch<T>::value;
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < G; j += NWARP) {
        T *row = dst_plane + ((long long)(c.k0 * G + j) * c.e + c.a_g) * c.out_pitch;
        const int dbase = c.d0 - c.a_g * j;
        const T *b = buf + j * P;
        switch ((c.a_g * j) & 3) {
        case 0: bwd_store_row<T, M, TD, 0>(b, row, c, dbase, zero, lane); break;
        case 1: bwd_store_row<T, M, TD, 1>(b, row, c, dbase, zero, lane); break;
        case 2: bwd_store_row<T, M, TD, 2>(b, row, c, dbase, zero, lane); break;
        default: bwd_store_row<T, M, TD, 3>(b, row, c, dbase, zero, lane); break;
        }
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

// Phases: 0 = load, then (compute, store) per step, last = store to global.
// `regs` is the thread's private NREG-element array (registers on the GPU).
template <typename T, int M, int LOADK, int STOREK>
struct FwdProgram {
    static constexpr int G = Geo<M>::G;
    // odd M reading workspace rows: stage 0 is done by the loader, the rest is radix-4
    static constexpr bool kFused0 = (M & 1) && LOADK == LOAD_WROWS;
    static constexpr int NS = kFused0 ? (M - 1) / 2 : num_steps(M);
    static constexpr int kPhases = 2 + 2 * NS;
    // fp64 image passes of 4+ stages and the fp32 six-stage one (64 registers anyway) fuse the first
    // radix-4 step into the loader (fwd_radix4_from_image); phases 1 and 2 are then empty
    static constexpr bool kFusedLoad = LOADK == LOAD_IMAGE && M >= 4 && (sizeof(T) == 8 || M == 6);
    static constexpr int TD = TileTD<M, STOREK>::value;   // offsets produced per tile
    static constexpr int LH = XW - TD;                    // tile position of offset d0

    // c.d0 must already be set.
    ADRT_HD static int classify(const TileCtx &c)
    {
        int sup = c.n + c.a_g * G + G - 1;  // first offset at which every output row is zero
        if (sup > c.D) sup = c.D;
        if (STOREK == STORE_QCOLS) {
            if (c.d0 >= c.D) return TILE_SKIP;
            if (c.d0 >= sup) return TILE_ZERO;
        } else {
            if (c.d0 >= c.out_pitch) return TILE_SKIP;
            // stored rows are skewed by up to 3 elements: position d0 holds offset d0 - s
            if (c.d0 - 3 >= sup) return TILE_ZERO;
        }
        return TILE_FULL;
    }

    // whole tile is zeros: just write them
    ADRT_HD static void zero_tile(T *buf, T *dst, const TileCtx &c, int tid)
    {
        if (STOREK == STORE_QCOLS) store_qcols<T, M, TD>(buf, dst, c, 0, true, tid);
        else fwd_store_wrows<T, M, LH, TD>(buf, dst, c, true, tid);
    }

    // phase PH of a full tile; every per-step quantity is a compile-time constant
    template <int PH>
    ADRT_HD static void phase_ct(int mode, T *buf, T (&regs)[NREG], const T *src, T *dst, const TileCtx &c, int tid)
    {
        (void)mode;
        if constexpr (PH == 0 && kFusedLoad) {
            fwd_radix4_from_image<T, M, LH>(buf, src, c, tid);
        } else if constexpr (kFusedLoad && (PH == 1 || PH == 2)) {
            // done in phase 0
        } else if constexpr (PH == 0) {
            if (LOADK == LOAD_IMAGE) fwd_load_image<T, M, LH>(buf, src, c, tid);
            else if (kFused0) fwd_load_wrows_stage0<T, M, LH>(buf, src, c, tid);
            else fwd_load_wrows<T, M, LH>(buf, src, c, tid);
        } else if constexpr (PH <= 2 * NS) {
            constexpr int step = (PH - 1) >> 1;
            constexpr int t = kFused0 ? 1 + 2 * step : 2 * step;
            if constexpr ((PH - 1) & 1) fwd_step_store<T, M, t>(buf, tid, regs);
            else fwd_step_compute<T, M, t>(buf, tid, regs);
        } else {
            if (STOREK == STORE_QCOLS) store_qcols<T, M, TD>(buf, dst, c, LH, false, tid);
            else fwd_store_wrows<T, M, LH, TD>(buf, dst, c, false, tid);
        }
    }
};

template <typename T, int M, int LOADK, int STOREK, bool kSub = false>
struct BwdProgram {
    static constexpr int G = Geo<M>::G;
    // odd M writing workspace rows: the transposed stage 0 is done by the store
    static constexpr bool kFused0 = (M & 1) && STOREK == STORE_WROWS;
    static constexpr int NS = kFused0 ? (M - 1) / 2 : num_steps(M);
    static constexpr int kPhases = 2 + 2 * NS;
    static constexpr int TD = TileTD<M, STOREK>::value;
    // stage produced by step i, and the radix of the step before it
    static constexpr int step_t(int i) { return kFused0 ? (M - 2) - 2 * i : 2 * (NS - 1 - i); }
    static constexpr int step_rprev(int i) { return i == 0 ? 0 : ((step_t(i - 1) + 2 <= M) ? 4 : 2); }
    // interior tiles of passes that read workspace rows and start with a radix-4 step
    // skip the staging copy: step 0 reads its windows from global memory
    static constexpr bool kDirect0 = LOADK == LOAD_WROWS && NS > 0 && (step_t(0) + 2 <= M);
    // passes that read the public layout and start with a radix-4 step fuse that step into the
    // loader (bwd_radix4_from_qcols); phases 1 and 2 are then empty
    static constexpr bool kFusedLoad = LOADK == LOAD_QCOLS && NS > 0 && (step_t(0) + 2 <= M) && M >= 4 &&
                                       (4 << step_t(0)) == G;
    static_assert(!kSub || kFusedLoad, "subtract-on-load exists for the fused public-layout loader only");

    ADRT_HD static int classify(const TileCtx &c)
    {
        // no wanted output row reaches this far (output row j sits at tile offset d + a_g*j)
        if (c.d0 >= c.d_need + c.a_g * (G - 1)) return TILE_SKIP;
        if (c.d0 >= c.D) return STOREK == STORE_QCOLS ? TILE_SKIP : TILE_ZERO;
        if (c.d0 + XW + 8 > c.D) return TILE_FULL_MASKED;
        return TILE_FULL;
    }

    ADRT_HD static void zero_tile(T *buf, T *dst, const TileCtx &c, int tid)
    {
        // same row coverage as the store of a full tile
        if (kFused0) bwd_store_wrows_stage0<T, M, TD, false, (NS > 0)>(buf, dst, c, true, tid);
        else bwd_store_wrows<T, M, TD>(buf, dst, c, true, tid);
    }

    template <int PH>
    ADRT_HD static void phase_ct(int mode, T *buf, T (&regs)[NREG], const T *src, T *dst, const TileCtx &c, int tid)
    {
        if constexpr (PH == 0 && kFusedLoad) {
            if (mode == TILE_FULL_MASKED) bwd_radix4_from_qcols<T, M, true, step_t(0), kSub>(buf, src, c, tid);
            else bwd_radix4_from_qcols<T, M, false, step_t(0), kSub>(buf, src, c, tid);
        } else if constexpr (kFusedLoad && (PH == 1 || PH == 2)) {
            // done in phase 0
        } else if constexpr (PH == 0) {
            if (LOADK == LOAD_QCOLS) bwd_load_qcols<T, M>(buf, src, c, tid);
            else if (!(kDirect0 && mode == TILE_FULL)) bwd_load_wrows<T, M>(buf, src, c, tid);
        } else if constexpr (PH <= 2 * NS) {
            constexpr int step = (PH - 1) >> 1;
            constexpr int t = step_t(step), rp = step_rprev(step);
            if constexpr ((PH - 1) & 1) bwd_step_store<T, M, t, rp>(buf, tid, regs);
            else if (mode == TILE_FULL_MASKED) bwd_step_compute<T, M, true, t, rp>(buf, c.D - c.d0, c.a_g, tid, regs);
            else if constexpr (kDirect0 && step == 0) bwd_radix4_compute_global<T, M, t>(src, c, tid, regs);
            else bwd_step_compute<T, M, false, t, rp>(buf, c.D - c.d0, c.a_g, tid, regs);
        } else {
            if (STOREK == STORE_QCOLS) store_qcols<T, M, TD>(buf, dst, c, 0, false, tid);
            else if (!kFused0) bwd_store_wrows<T, M, TD>(buf, dst, c, false, tid);
            else if (mode == TILE_FULL_MASKED) bwd_store_wrows_stage0<T, M, TD, true, (NS > 0)>(buf, dst, c, false, tid);
            else bwd_store_wrows_stage0<T, M, TD, false, (NS > 0)>(buf, dst, c, false, tid);
        }
    }
};

// Run-time phase index -> compile-time phase (host emulator; the CUDA kernel
// unrolls the sequence instead).
template <typename Prog, typename T, int PH = 0>
ADRT_HD void run_phase(int ph, int mode, T *buf, T (&regs)[NREG], const T *src, T *dst, const TileCtx &c, int tid)
{
    if constexpr (PH < Prog::kPhases) {
        if (ph == PH) Prog::template phase_ct<PH>(mode, buf, regs, src, dst, c, tid);
        else run_phase<Prog, T, PH + 1>(ph, mode, buf, regs, src, dst, c, tid);
    }
}

}  // namespace tile
}  // namespace adrt_b200
