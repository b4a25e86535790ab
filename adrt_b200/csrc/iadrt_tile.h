// Fused multi-stage passes of the exact inverse (adrt.iadrt) -- warp-level algorithm.
//
// Plain C++ that compiles as CUDA device code (iadrt_fused.cu) and as host code
// (tests/emu/emu_fused.cpp plays the 32 lanes of a warp phase by phase), like fused_tile.h.
//
// ---------------------------------------------------------------------------
// The maths (reference adrt_cdefs_iadrt.hpp:52-105).  Stage s (C_in = n >> s
// columns per block, C = C_in / 2, L = 2^(s+1) blocks) in the public (d, column) layout: output
// column (l, col), l < L, col < C, reads the input columns A = (l >> 1, 2 col) and A1 = (l >> 1, 2 col + 1):
//
//   l even:  delta[d] = (0 + A[d]) - A1[d + 1]                      (second term if d + 1 < D)
//   l odd:   delta[d] = (0 + A1[d + 1 + col]) - A[d + 1 + col]      (0 if d + 1 + col >= D)
//   out[d]   = delta[d] + out[d + 1]                                 (no add at d = D - 1)
//
// i.e. a serial suffix scan along d (the reference's "must be serial" loop) of a difference of two
// input columns.  One kernel per stage moves the sinogram through HBM 2 K times.
//
// Fusing m stages s0 .. s0+m-1 (G = 2^m): the G adjacent input columns (l', c0 G + j), j < G, produce
// the G output columns (l' G + lambda, c0), lambda < G.  At inner level t (0 = inputs, m = outputs) the
// nodes are (lambda_t < 2^t, j_t < 2^(m-t)) = column (l' 2^t + lambda_t, c0 2^(m-t) + j_t); G nodes per
// level -> one lane each, k = lambda_t 2^(m-t) + j_t.  The odd-branch shift 1 + col has the large part
// c0 2^(m-t) in common for the whole group.  Pre-shifted frames remove it: node (t, lambda_t) is
// handled in the coordinate
//
//       x = d + psi,   psi = c0 2^(m-t) lambda_t
//
// and then a level-t node at x reads its parents (level t-1, lanes kA = (lambda_t >> 1) 2^(m-t+1) + 2 j_t
// and kA + 1) at x and x + 1 (even lambda_t) or at x + 1 + j_t (odd lambda_t): look-ahead <= 2^(m-t).
// The scans run downwards in d, hence downwards in x, so everything a node needs was produced a few
// rows earlier: a warp sweeps x from the top, every lane scanning its node of every level, the
// levels one row apart (level t works on row X + t while level t-1 works on X + t - 1), the last
// rows of every level in a small shared-memory ring.  Each element costs exactly the reference's
// operations in the reference's order; only the iteration space is re-indexed, so results are
// bit-identical (signed zeros included: "0 + a" is kept).
//
// Layouts: the first pass reads the public layout (rows coalesced across the lanes' adjacent columns),
// the last pass writes it (its psi is 0: C_out = 1 means c0 = 0).  Between passes the data lives in a
// column-major workspace W[column][2n] (2n - 1 offsets + 1 pad: rows 16-byte aligned): a lane streams
// its own column with 16-byte vectors, four offsets per access, in phase with all other lanes.
#pragma once

#include "fused_tile.h"

namespace adrt_b200 {
namespace itile {

constexpr int kLanes = 32;

ADRT_HD constexpr int pow2_ceil(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <int M> struct Geo {
    static_assert(M >= 1 && M <= 5, "a team is at most one warp");
    static constexpr int G = 1 << M;              // lanes per team
    static constexpr int TEAMS = kLanes / G;      // teams per warp
    // rows of level t (t < M) that are live at once: the reader (level t+1) looks ahead 2^(M-1-t) + 1 rows;
    // level 0 is committed eight rows at a time (two trips), seven rows early
    ADRT_HD static constexpr int depth(int t) { return pow2_ceil((1 << (M - 1 - t)) + 2 + (t == 0 ? 7 : 0)); }
    // Every level ring t < M has three MIRROR rows in front of it: physical row p = r + 3 holds ring row r,
    // and ring rows depth-3 .. depth-1 are stored a second time at p = 0 .. 2.  The four rows r0, r0 - 1,
    // r0 - 2, r0 - 3 (mod depth) a reader needs in one trip are then the contiguous physical rows
    // r0 + 3 - u, whatever the lane's own look-ahead: one address per operand and trip, the rows as
    // immediate offsets (the address arithmetic was 10 of the 16 instructions per node and row).
    static constexpr int MIRROR = 3;
    // The LAST level reads its two parents from registers: node (M, lambda) has j = 0, so its parents are the
    // level-(M-1) nodes of its own lane pair (lanes lambda & ~1 and lambda | 1) one row and two rows back --
    // a lane keeps its last two level-(M-1) values (LaneState::h1 / h2) and one warp shuffle per row step
    // brings the partner's; level M-1 has no ring (M >= 2; level 0 is the input ring).
    static constexpr bool kRegLast = M >= 2;
    ADRT_HD static constexpr int rows_of(int t) { return (kRegLast && t == M - 1) ? 0 : depth(t) + MIRROR; }
    ADRT_HD static constexpr int base(int t) { return t == 0 ? 0 : base(t - 1) + rows_of(t - 1); }
    static constexpr int OUT_DEPTH = 16;          // level-M ring (workspace stores): one aligned group of 8 + skew
    static constexpr int OUT_BASE = base(M);
    static constexpr int ROWS = OUT_BASE + OUT_DEPTH;   // ring rows per warp (x 32 lanes)
};

// Everything a lane needs to know about its team's group.
struct Team {
    int n, D;          // image side, 2n - 1
    int c0;            // group index inside its block
    long long in_col;  // first input column of the group (public layout: column index; workspace: row index)
    long long out_col; // output column of lambda = 0 (public layout) / workspace row of lambda = 0
    long long out_stride;  // distance between the output columns / rows of consecutive lambda
    bool active;       // false: padding team (more teams in the warp than groups left)
};

// per-lane registers that live across iterations
template <typename T, int M> struct LaneState {
    T prev[M + 1];     // running scan value of the lane's node at level t (index t, 1..M)
    T v[8];            // eight input rows fetched for the next pair of trips
    T h1, h2;          // Geo::kRegLast: the lane's level-(M-1) values at the last level's row x and at x + 1
};

// Per-lane constants of the sweep, one set per level t = 1..M (index t): where the two operands of the
// lane's node live in the ring of level t-1 and for which base rows X (the warp-uniform loop variable;
// the node's row is x = X + t, its offset d = x - psi) each term of the reference's expression exists.
//   even lambda:  acc = (0 + A[x]) - A1[x + 1]            second term iff d + 1 < D
//   odd lambda:   acc = (0 + A1[x + 1 + j]) - A[x + 1 + j]  both terms iff d + 1 + col < D
//   acc += prev iff d + 1 < D;  the node exists iff 0 <= d < D
// Both parities are "first minus second" with per-lane operand positions, so the warp does not diverge.
template <int M> struct LaneConst {
    int colF[M + 1], colS[M + 1];   // ring element offset (row 0) of the first / second operand's column
    int dF[M + 1], dS[M + 1];       // their row look-ahead relative to x
    int lo[M + 1], hi[M + 1];       // node exists for lo <= X <= hi
    int thr1[M + 1], thr2[M + 1];   // first / second term exists for X <= thr
    int own[M + 1];                 // ring element offset (row 0) of the lane's own column at level t
    bool odd_last;                  // parity of the lane's last-level node (Geo::kRegLast)
};

template <int M, int t = 1>
ADRT_HD void setup_levels(const Team &tm, int team_lane0, int k, int lane, LaneConst<M> &lc)
{
    if constexpr (t <= M) {
        constexpr int sh = M - t;
        const int lam = k >> sh, j = k & ((1 << sh) - 1);
        const int psi = (tm.c0 << sh) * lam;
        const int kA = ((lam >> 1) << (sh + 1)) + 2 * j + team_lane0;
        constexpr int rb = Geo<M>::base(t - 1);
        const bool odd = lam & 1;
        lc.colF[t] = rb * kLanes + (odd ? kA + 1 : kA);
        lc.colS[t] = rb * kLanes + (odd ? kA : kA + 1);
        lc.dF[t] = odd ? 1 + j : 0;
        lc.dS[t] = odd ? 1 + j : 1;
        // d = X + t - psi
        lc.lo[t] = psi - t;
        lc.hi[t] = tm.D - 1 + psi - t;
        const int col = (tm.c0 << sh) + j;
        const int t_odd = tm.D - 2 - col + psi - t;     // d + 1 + col < D
        lc.thr1[t] = odd ? t_odd : 0x7fffffff;
        lc.thr2[t] = odd ? t_odd : tm.D - 2 + psi - t;  // even: d + 1 < D
        if (!tm.active) { lc.lo[t] = 1; lc.hi[t] = 0; } // padding team: no node ever exists
        lc.own[t] = (t < M ? Geo<M>::base(t) : Geo<M>::OUT_BASE) * kLanes + lane;
        if (t == M) lc.odd_last = odd;
        setup_levels<M, t + 1>(tm, team_lane0, k, lane, lc);
    }
}

// Ring positions of one trip (base rows X0 .. X0 - 3, X0 = 3 mod 4), element offsets, one set per level t:
//   F, S   the lane's two operands at u = 0 in the level-(t-1) ring; row u is at - 32 u (mirror rows)
//   W0, W1 the lane's own cell of row 0 of the two aligned row groups the trip's rows x = X0 + t - u fall
//          into: x = c (mod 4) at u = 0 with c = (3 + t) & 3 known at compile time, so u <= c is row c - u
//          of group W0 and u > c row 4 + c - u of group W1 (the group below, modulo the ring depth)
//   top    bit 2t / 2t+1: group W0 / W1 is the ring's last one -- its rows 1 .. 3 are mirrored
template <int M> struct TripAddr {
    int F[M + 1], S[M + 1], W0[M + 1], W1[M + 1];
    unsigned top;
};

template <int M, bool kOutQ, int t = 1>
ADRT_HD void trip_setup(const LaneConst<M> &lc, int X0, TripAddr<M> &ta)
{
    if constexpr (t == 1) ta.top = 0;
    if constexpr (t <= M) {
        constexpr int dp = Geo<M>::depth(t - 1), MR = Geo<M>::MIRROR;
        const int x0 = X0 + t;
        ta.F[t] = lc.colF[t] + ((((x0 + lc.dF[t]) & (dp - 1)) + MR) << 5);
        ta.S[t] = lc.colS[t] + ((((x0 + lc.dS[t]) & (dp - 1)) + MR) << 5);
        constexpr int c = (3 + t) & 3;
        constexpr int dw = t < M ? Geo<M>::depth(t) : Geo<M>::OUT_DEPTH;
        const int g0 = (x0 - c) & (dw - 1), g1 = (x0 - c - 4) & (dw - 1);
        ta.W0[t] = lc.own[t] + ((g0 + (t < M ? MR : 0)) << 5);
        ta.W1[t] = lc.own[t] + ((g1 + (t < M ? MR : 0)) << 5);
        if (t < M && dw > 4) ta.top |= (g0 == dw - 4 ? 1u << (2 * t) : 0u) | (g1 == dw - 4 ? 2u << (2 * t) : 0u);
        trip_setup<M, kOutQ, t + 1>(lc, X0, ta);
    }
}

// ---- input side ---------------------------------------------------------------------------------
// One 32-byte sector of a lane's workspace row (8 floats / 4 doubles), 32-byte aligned: the lanes of a warp
// stream 32 different rows, so every global access of a pass costs 32 L1 wavefronts whatever its width --
// the widest access per lane halves that cost per element against 16-byte vectors.
ADRT_HD void load_sector(const float *p, float *dst)
{
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=f"(dst[0]), "=f"(dst[1]), "=f"(dst[2]), "=f"(dst[3]), "=f"(dst[4]), "=f"(dst[5]), "=f"(dst[6]), "=f"(dst[7])
                 : "l"(p));
#else
    for (int i = 0; i < 8; ++i) dst[i] = p[i];
#endif
}
ADRT_HD void load_sector(const double *p, double *dst)
{
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];\n" : "=d"(dst[0]), "=d"(dst[1]), "=d"(dst[2]), "=d"(dst[3]) : "l"(p));
#else
    for (int i = 0; i < 4; ++i) dst[i] = p[i];
#endif
}
ADRT_HD void store_sector(float *p, const float *v)
{
#ifdef __CUDA_ARCH__
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
#else
    for (int i = 0; i < 8; ++i) p[i] = v[i];
#endif
}
ADRT_HD void store_sector(double *p, const double *v)
{
#ifdef __CUDA_ARCH__
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
#else
    for (int i = 0; i < 4; ++i) p[i] = v[i];
#endif
}
template <typename T> struct Sector { static constexpr int L = 32 / (int)sizeof(T); };

// Rows X0-7 .. X0 (X0 = 7 mod 8) of the lane's input column, to be committed at the start of the pair of
// trips with base rows X0 .. X0-3 and X0-4 .. X0-7.  kInQ: public layout in[d][col] (pitch n), else
// workspace W[col][2n].  `col_ptr`: the lane's column (public layout: &in[0][col]; workspace: &W[col][0]).
template <typename T, bool kInQ>
ADRT_HD void fetch_inputs(const T *col_ptr, const Team &tm, int X0, T (&v)[8])
{
    if (!tm.active || X0 < 0 || X0 - 7 >= tm.D) return;
    if (kInQ) {
        const T *p = col_ptr + (long long)(X0 - 7) * tm.n;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (X0 - 7 + e < tm.D) v[e] = p[(long long)e * tm.n];
    } else {
        // aligned 32-byte sector(s) of the padded row; the pad cell (offset 2n - 1) is loaded and never used
        const T *p = col_ptr + (X0 - 7);
        constexpr int L = Sector<T>::L;
#pragma unroll
        for (int g = 0; g < 8 / L; ++g) load_sector(p + g * L, &v[g * L]);
    }
}

// rows X0-7 .. X0 into the level-0 ring: two aligned groups of four
template <typename T, int M>
ADRT_HD void commit_inputs(T *ring, const Team &tm, int lane, int X0, const T (&v)[8])
{
    if (!tm.active || X0 < 0 || X0 - 7 >= tm.D) return;
    constexpr int dp = Geo<M>::depth(0), MR = Geo<M>::MIRROR;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int x4 = X0 - 7 + 4 * h;       // a multiple of 4: one aligned row group
        const int g = x4 & (dp - 1);
        const bool top = g == dp - 4;        // the ring's last group: rows 1 .. 3 also go to the mirror rows 0 .. 2
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (x4 + e < tm.D) {
                ring[((g + e + MR) << 5) + lane] = v[4 * h + e];
                if (e >= 1 && top) ring[((e - 1) << 5) + lane] = v[4 * h + e];
            }
        }
    }
}

// ---- all levels of one base row X = X0 - U for one lane ------------------------------------------------
// A level reads rows of the level below that were written in EARLIER row steps only (the levels run one
// row apart), so a row step is: all operands of all levels -> registers, all sums, all stores.  Written
// that way the compiler sees the independence (it cannot prove that the ring cells do not alias) and the
// 2 M shared-memory loads of a row step are in flight together instead of one level after the other.
// kOutQ: the last level is stored straight to the public layout (`out_ptr` = &out[0][lane's column]; the
// last pass has psi = 0, so every lane is at the same offset d = X + M: one coalesced row piece).
template <typename T, int M, int U, int t = 1>
ADRT_HD void load_levels(const T *ring, const TripAddr<M> &ta, T (&first)[M + 1], T (&second)[M + 1])
{
    if constexpr (t <= (Geo<M>::kRegLast ? M - 1 : M)) {
        first[t] = ring[ta.F[t] - (U << 5)];
        second[t] = ring[ta.S[t] - (U << 5)];
        load_levels<T, M, U, t + 1>(ring, ta, first, second);
    }
}

// operands of the last level from the lane pair's registers (`partner_h2`: the other lane's h2, by shuffle)
template <typename T, int M>
ADRT_HD void last_level_operands(const LaneConst<M> &lc, const LaneState<T, M> &st, T partner_h2, T (&first)[M + 1],
                                 T (&second)[M + 1])
{
    if constexpr (Geo<M>::kRegLast) {
        first[M] = lc.odd_last ? st.h2 : st.h1;    // even: A[x] (own, one row back); odd: A1[x + 1] (own, two back)
        second[M] = partner_h2;                     // even: A1[x + 1]; odd: A[x + 1]
    }
}

template <typename T, int M, bool kOutQ, int U, int t>
ADRT_HD void put_node(T *ring, const TripAddr<M> &ta, int n, int x, T acc, T *out_ptr, T &hnew)
{
    constexpr int c = (3 + t) & 3;
    constexpr int i = U <= c ? c - U : 4 + c - U;          // row of its aligned group
    if constexpr (Geo<M>::kRegLast && t == M - 1) {
        hnew = acc;
    } else if constexpr (t < M) {
        const int w = (U <= c ? ta.W0[t] : ta.W1[t]) + (i << 5);
        constexpr int dw = Geo<M>::depth(t);
        ring[w] = acc;
        if constexpr (i >= 1) {
            const int mirror = Geo<M>::base(t) * kLanes + ((i - 1) << 5);
            const int lane_col = (U <= c ? ta.W0[t] : ta.W1[t]) & (kLanes - 1);
            if (dw == 4 || (ta.top >> (2 * t + (U <= c ? 0 : 1))) & 1u) ring[mirror + lane_col] = acc;
        }
    } else if constexpr (kOutQ) {
        out_ptr[(long long)x * n] = acc;
    } else {
        ring[(U <= c ? ta.W0[t] : ta.W1[t]) + (i << 5)] = acc;
    }
}

template <typename T, int M, bool kOutQ, int U, int t = 1>
ADRT_HD void guarded_levels(T *ring, const LaneConst<M> &lc, const TripAddr<M> &ta, int n, int X0, LaneState<T, M> &st,
                            T *out_ptr, const T (&first)[M + 1], const T (&second)[M + 1], T &hnew)
{
    if constexpr (t <= M) {
        const int X = X0 - U;
        T acc = X <= lc.thr1[t] ? T(0) + first[t] : T(0);
        acc = X <= lc.thr2[t] ? acc - second[t] : acc;
        acc = X < lc.hi[t] ? acc + st.prev[t] : acc;
        if (X >= lc.lo[t] && X <= lc.hi[t]) {
            st.prev[t] = acc;
            put_node<T, M, kOutQ, U, t>(ring, ta, n, X + t, acc, out_ptr, hnew);
        }
        guarded_levels<T, M, kOutQ, U, t + 1>(ring, lc, ta, n, X0, st, out_ptr, first, second, hnew);
    }
}

template <typename T, int M, bool kOutQ, int U>
ADRT_HD void all_levels(T *ring, const LaneConst<M> &lc, const TripAddr<M> &ta, int n, int X0, LaneState<T, M> &st,
                        T *out_ptr, T partner_h2)
{
    T first[M + 1], second[M + 1];
    load_levels<T, M, U>(ring, ta, first, second);
    last_level_operands<T, M>(lc, st, partner_h2, first, second);
    T hnew = st.h1;     // where the level-(M-1) node does not exist the value is never used
    guarded_levels<T, M, kOutQ, U>(ring, lc, ta, n, X0, st, out_ptr, first, second, hnew);
    st.h2 = st.h1;
    st.h1 = hnew;
}

// The same for base rows where every term of every level exists for every lane of the warp
// (interior_range): no comparisons, no selects.
template <typename T, int M, bool kOutQ, int U, int t = 1>
ADRT_HD void interior_levels(T *ring, const TripAddr<M> &ta, int n, int X0, LaneState<T, M> &st, T *out_ptr,
                             const T (&first)[M + 1], const T (&second)[M + 1], T &hnew)
{
    if constexpr (t <= M) {
        const T acc = ((T(0) + first[t]) - second[t]) + st.prev[t];
        st.prev[t] = acc;
        put_node<T, M, kOutQ, U, t>(ring, ta, n, X0 - U + t, acc, out_ptr, hnew);
        interior_levels<T, M, kOutQ, U, t + 1>(ring, ta, n, X0, st, out_ptr, first, second, hnew);
    }
}

template <typename T, int M, bool kOutQ, int U>
ADRT_HD void all_levels_interior(T *ring, const LaneConst<M> &lc, const TripAddr<M> &ta, int n, int X0, LaneState<T, M> &st,
                                 T *out_ptr, T partner_h2)
{
    T first[M + 1], second[M + 1];
    load_levels<T, M, U>(ring, ta, first, second);
    last_level_operands<T, M>(lc, st, partner_h2, first, second);
    T hnew = st.h1;
    interior_levels<T, M, kOutQ, U>(ring, ta, n, X0, st, out_ptr, first, second, hnew);
    st.h2 = st.h1;
    st.h1 = hnew;
}

// Base rows X for which this lane needs no guard at any level: [lo, hi] (empty for padding teams).
// The warp takes the unguarded path for a trip X0 .. X0-3 when it lies inside every lane's range.
template <int M>
ADRT_HD void interior_range(const LaneConst<M> &lc, bool active, int &lo, int &hi)
{
    lo = -0x40000000;
    hi = 0x40000000;
#pragma unroll
    for (int t = 1; t <= M; ++t) {
        if (lc.lo[t] > lo) lo = lc.lo[t];
        int h = lc.hi[t] - 1;                       // "+ prev" needs X < hi
        if (lc.thr1[t] < h) h = lc.thr1[t];
        if (lc.thr2[t] < h) h = lc.thr2[t];
        if (h < hi) hi = h;
    }
    if (!active) { lo = 1; hi = 0; }
}

// ---- workspace stores -------------------------------------------------------------------------------
// After the pair of trips with base rows X0 .. X0-7 the lane's output column is complete down to offset
// X0 - 7 + M - psi: flush the one aligned group of eight offsets that became complete in this pair.
// `row_ptr`: the lane's workspace row (&W[output column][0]); psi = c0 * lambda.
template <typename T, int M>
ADRT_HD void flush_outputs(const T *ring, const Team &tm, int psi, int lane, int X0, T *row_ptr)
{
    if (!tm.active) return;
    const int lo = X0 - 7 + M - psi;           // lowest offset computed so far
    const int d0 = (lo + 7) & ~7;              // lowest complete aligned group
    if (d0 < 0 || d0 >= 2 * tm.n) return;
    T w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = ring[Geo<M>::OUT_BASE * kLanes + lane + (((d0 + e + psi) & (Geo<M>::OUT_DEPTH - 1)) << 5)];
    constexpr int L = Sector<T>::L;
#pragma unroll
    for (int g = 0; g < 8 / L; ++g) store_sector(row_ptr + d0 + g * L, &w[g * L]);
}

// Sweep bounds of a warp: base rows from x_top (= 7 mod 8) downwards, two trips of four rows per iteration.
ADRT_HD int sweep_top(int D, int psi_max)
{
    return ((D - 1 + psi_max) | 7);
}

}  // namespace itile
}  // namespace adrt_b200

// ---- pass / team bookkeeping shared by the CUDA driver and the host emulator -----------------------
namespace adrt_b200 {
namespace itile {

// stages per pass: at most 5 per pass (a team is at most one warp), the FIRST pass the shortest -- its
// groups span the whole block (c0 up to n / G), so its lanes' frames are up to n rows apart and every lane
// idles through the other lanes' ramps; later passes have short ramps and amortise the per-row overhead
// over more levels.  Measured at 2048^2 (profiles/r05_iadrt.jsonl): fp32 2,4,5 4.61 ms / 3,4,4 4.93 /
// 4,4,3 5.13 / 1,5,5 5.09; fp64 (five-level rings are 2 x the shared memory) 3,4,4 4.82 / 2,4,5 4.96 /
// 4,4,3 5.26.  ADRT_B200_IADRT_SPLIT="4,4,3" overrides.
inline int iadrt_split(int K, int *ms /* [8] */, int elem_size = 4)
{
    if (const char *e = getenv("ADRT_B200_IADRT_SPLIT")) {
        int cnt = 0, sum = 0;
        bool ok = true;
        for (const char *p = e; *p && cnt < 8;) {
            const int v = (int)strtol(p, const_cast<char **>(&p), 10);
            if (v < 1 || v > 5) ok = false;
            ms[cnt++] = v;
            sum += v;
            if (*p == ',') ++p;
        }
        if (ok && sum == K && cnt <= 3) return cnt;
    }
    const int np = (K + 4) / 5;
    if (np == 3 && elem_size == 4 && K >= 10) {
        ms[0] = K - 9;
        ms[1] = 4;
        ms[2] = 5;
        return np;
    }
    // as even as possible, ascending
    int left = K;
    for (int i = 0; i < np; ++i) {
        const int m = left / (np - i);
        ms[i] = m;
        left -= m;
    }
    return np;
}

// teams of one plane are numbered tp = 0 .. n/G - 1: heaviest groups (largest c0: longest sweep) first,
// the 2^s0 blocks of one c0 next to each other (a warp's teams then sweep the same rows)
template <int M>
ADRT_HD Team make_team(int n, int s0, int tp)
{
    constexpr int G = Geo<M>::G;
    Team tm;
    tm.n = n;
    tm.D = 2 * n - 1;
    const int Cin0 = n >> s0, gpb = Cin0 / G, L0 = 1 << s0;
    tm.active = tp < gpb * L0;
    const int lp = tp & (L0 - 1);
    tm.c0 = gpb - 1 - (tp >> s0);
    if (!tm.active) tm.c0 = 0;
    tm.in_col = (long long)lp * Cin0 + (long long)tm.c0 * G;
    tm.out_col = (long long)lp * Cin0 + tm.c0;
    tm.out_stride = Cin0 / G;
    return tm;
}

}  // namespace itile
}  // namespace adrt_b200
