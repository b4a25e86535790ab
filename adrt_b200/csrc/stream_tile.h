// "Streaming" tile algorithms of the fused fp32 passes with M = 5 or 6 stages.
//
// Same maths, tiling and workspace formats as fused_tile.h (read its header
// first); what changes is how the butterfly steps walk the shared-memory tile.
// fused_tile.h gives a thread 4 consecutive offsets of a radix-4 butterfly and
// fetches the shifted operands as over-wide aligned windows (1.9 tile reads +
// 1 tile write per TWO stages).  Here a thread owns a radix-8 (or radix-4)
// butterfly for a whole SEGMENT of the offset axis and walks it serially,
// 4 offsets per iteration: the shifted operands of iteration i are the values
// the thread itself loaded or computed in iteration i-1, kept in registers, so
// every tile element is read once and written once per THREE stages.
//
//   pass of M stages = step 1 (radix 8, local stages 0-2)
//                    + step 2 (radix 2^(M-3), local stages 3..M-1)
//
// Local transform in pre-shifted frames (SURVEY.md 8a row a2'): node (level L,
// block m, local angle a) lives in the frame of its first leaf, and
//     node(L+1, k, 2a+b)[x] = node(L, 2k, a)[x] + node(L, 2k+1, a)[x - a - b]
// with shifts a + b <= 4 for L <= 2, i.e. inside the previous 4-vector.  The
// uniform part p*jj of the shift of leaf jj of a step-2 group with base angle p
// is an address offset of the leaf's row.
//
// In place, position preserving: a butterfly writes output q over the row of
// its input q (like an in-place FFT), so the rows of a tile end up digit
// reversed.  Passes that store workspace rows (any row order is as good as any
// other there) load naturally and store reversed; passes that store the public
// (d, column) layout load reversed and store naturally, so the transposing
// copies always walk consecutive tile rows.
//
// Hazards of the in-place update: segment S+1 needs the 8 offsets below its
// start (and, because of the alignment rule below, its first vector) before
// segment S overwrites them: every step is  prologue (warm-up + first loads)
// -> barrier -> main loop, and the main loop loads iteration i+1 before it
// stores iteration i.
//
// Alignment: all vector accesses are 16-byte aligned.  Step 2 reads leaf jj at
// offset -p*jj: rows with jj % 4 == 0 use LDS.128, jj % 4 == 2 LDS.64, odd jj
// LDS.32 (same bytes, conflict free: the 4 groups of a warp have 4 consecutive
// base angles p, hence 4 distinct residues of p*jj mod 4 for odd jj).  Output q
// of step 2 is stored with the skew floor4(p*q) (logical offset x sits at
// x - floor4(p*q)): never below the position its row was read from, at most 3
// above it, which the one-iteration look-ahead covers.
//
// Segments: 8 per group, lengths 36 / 28 (= 4 mod 8 vectors-of-4... 9 or 7
// vectors), so the 8 lanes of a quarter warp hit 8 distinct 16-byte bank groups.
#pragma once

#include "fused_tile.h"

namespace adrt_b200 {
namespace stile {

using tile::TileCtx;
using tile::Pack;
using tile::LOAD_IMAGE;
using tile::LOAD_WROWS;
using tile::LOAD_QCOLS;
using tile::STORE_WROWS;
using tile::STORE_QCOLS;
using tile::fwd_row_skew;

// iterations of a step's main loop unrolled together: the per-thread history lives in registers, so
// the loop must be unrolled enough for the compiler to rename it instead of moving it, but the whole
// kernel should stay inside the instruction cache (several CTAs in different phases share an SM)
#ifndef ADRT_STREAM_UNROLL
#define ADRT_STREAM_UNROLL 3
#endif
constexpr int kStreamUnroll = ADRT_STREAM_UNROLL;
// 4 x 4 blocks a thread of a transposing loader keeps in flight (UB * 4 vector loads); the 5-stage
// kernels run 6 CTAs per SM and have 168 registers
template <int M> struct LoadBatch { static constexpr int value = (M == 6) ? 6 : 3; };

constexpr int V = 4;
constexpr int XW = 288;          // offsets per tile row
constexpr int P = 292;           // row pitch (floats): 4 mod 32
constexpr int NVEC = XW / V;     // 72 vectors per row

typedef Pack<float> F4;
struct alignas(8) F2 { float v[2]; };

// 16-byte global -> shared copy that does not pass through registers (LDGSTS): a thread
// queues all its copies of a tile and waits once, so the whole tile is in flight at once.
ADRT_HD void copy16_async(float *dst_smem, const float *src)
{
#ifdef __CUDA_ARCH__
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
#else
    *reinterpret_cast<F4 *>(dst_smem) = *reinterpret_cast<const F4 *>(src);
#endif
}
// ---- bulk (TMA) row copies signalled on an mbarrier ----------------------------------
// One thread moves a whole row segment with one instruction; the tile's threads then wait on
// the barrier's phase.  Host emulator: synchronous copies.
struct BulkBar {
    unsigned long long *bar;   // 8 bytes of shared memory
    unsigned phase;            // parity of the next completion
};

ADRT_HD void bulk_init(BulkBar &b, unsigned long long *bar, int nthreads, int tid)
{
    b.bar = bar;
    b.phase = 0;
#ifdef __CUDA_ARCH__
    if (tid == 0) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(nthreads) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
#else
    (void)nthreads; (void)tid;
#endif
}

// order earlier generic-proxy accesses of shared memory before the async-proxy writes that follow
ADRT_HD void bulk_fence()
{
#ifdef __CUDA_ARCH__
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#endif
}

// global -> shared, `bytes` (multiple of 16, both addresses 16-byte aligned)
ADRT_HD void bulk_load(const BulkBar &b, float *dst_smem, const float *src, int bytes)
{
#ifdef __CUDA_ARCH__
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const unsigned a = (unsigned)__cvta_generic_to_shared(b.bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(a) : "memory");
#else
    (void)b;
    for (int i = 0; i < bytes / 4; ++i) dst_smem[i] = src[i];
#endif
}

// every thread of the tile arrives once, announcing the bytes of the copies it issued, then waits
ADRT_HD void bulk_arrive_wait(BulkBar &b, int my_bytes)
{
#ifdef __CUDA_ARCH__
    const unsigned a = (unsigned)__cvta_generic_to_shared(b.bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(my_bytes) : "memory");
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(b.phase) : "memory");
    b.phase ^= 1;
#else
    (void)b; (void)my_bytes;
#endif
}

// 16-byte shared-memory load the compiler may not narrow (narrowed window loads turn into 4-byte
// loads with a 16-byte lane stride: 4-way bank conflicts)
ADRT_HD F4 lds128(const float *p)
{
#ifdef __CUDA_ARCH__
    F4 v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.v[0]), "=f"(v.v[1]), "=f"(v.v[2]), "=f"(v.v[3]) : "r"(a));
    return v;
#else
    return *reinterpret_cast<const F4 *>(p);
#endif
}

// N consecutive elements starting Q after the 16-byte aligned p, from whole 16-byte loads
template <int N, int Q>
ADRT_HD void load_window_wide(const float *p, float (&dst)[N])
{
    constexpr int NV = (Q + N + 3) / 4;
    F4 tmp[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) tmp[v] = lds128(p + 4 * v);
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = tmp[(Q + i) / 4].v[(Q + i) % 4];
}

// Sources of the constant fills of a tile row (+0.0 above the data, -0.0 "copy" sentinels below
// offset 0), so that fills are bulk copies too instead of shared-memory stores.
#define ADRT_R8(x) x, x, x, x, x, x, x, x
#define ADRT_R32(x) ADRT_R8(x), ADRT_R8(x), ADRT_R8(x), ADRT_R8(x)
#define ADRT_R288(x) ADRT_R32(x), ADRT_R32(x), ADRT_R32(x), ADRT_R32(x), ADRT_R32(x), ADRT_R32(x), ADRT_R32(x), ADRT_R32(x), ADRT_R32(x)
#ifdef __CUDACC__
__device__ __align__(16) const float g_fill_pos[288] = {ADRT_R288(0.0f)};
__device__ __align__(16) const float g_fill_neg[288] = {ADRT_R288(-0.0f)};
#endif
static const float h_fill_pos[288] = {ADRT_R288(0.0f)};
static const float h_fill_neg[288] = {ADRT_R288(-0.0f)};
ADRT_HD const float *fill_src(bool neg)
{
#ifdef __CUDA_ARCH__
    return neg ? g_fill_neg : g_fill_pos;
#else
    return neg ? h_fill_neg : h_fill_pos;
#endif
}

// one 32-byte sector (8 consecutive columns of the public layout)
ADRT_HD void store8(float *p, const float (&v)[8])
{
#if defined(__CUDA_ARCH__) && defined(ADRT_STORE8_SPLIT)
    // A/B: two 16-byte stores (the L1 data pipe spends one wavefront per LANE on a 256-bit store)
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
#elif defined(__CUDA_ARCH__)
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
#else
    for (int i = 0; i < 8; ++i) p[i] = v[i];
#endif
}

ADRT_HD void copy4_async(float *dst_smem, const float *src)
{
#ifdef __CUDA_ARCH__
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(src) : "memory");
#else
    *dst_smem = *src;
#endif
}
ADRT_HD void copy_async_wait()
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
#endif
}

// The radix-8 step is the one next to the public layout (last forward step, last transposed step):
// its 8 outputs are 8 consecutive columns = one 32-byte sector per offset, written straight from
// registers.  Forward: step 1 radix 2^(M-3), step 2 radix 8.  Transposed: step A (the transpose of
// step 2) radix 2^(M-3), step B (the transpose of step 1) radix 8.
template <int M, bool kFwd> struct SGeo {
    static_assert(M == 5 || M == 6, "streaming passes fuse 5 or 6 stages");
    static constexpr int G = 1 << M;
    static constexpr int LOGR1 = kFwd ? M - 3 : 3, R1 = 1 << LOGR1;
    static constexpr int LOGR2 = kFwd ? 3 : M - 3, R2 = 1 << LOGR2;
    // step 1 covers the whole row, step 2 the offsets [X2, XW)
    static constexpr int SEG1 = 36, NSEG1 = 8;
    static constexpr int SEG2 = (M == 6) ? 28 : 36, NSEG2 = (M == 6) ? 8 : 7;
    static constexpr int X2 = XW - SEG2 * NSEG2;     // 64 (M = 6), 36 (M = 5): >= G, the offsets M stages consume
};

template <int M, int STOREK> struct STileTD {
    static constexpr int value = XW - SGeo<M, true>::X2 - (STOREK == STORE_WROWS ? 4 : 0);
};

// Row order.  kRev = false: leaf j sits in tile row j and output angle A ends in
// row (A % R2) * R1 + A / R2.  kRev = true: leaf j = k0*R1 + jj sits in row
// jj*R2 + k0 and output angle A ends in row A.
template <int M, bool kRev, bool kFwd> struct RowMap {
    static constexpr int R1 = SGeo<M, kFwd>::R1, R2 = SGeo<M, kFwd>::R2;
    ADRT_HD static int leaf_row(int j) { return kRev ? (j % R1) * R2 + j / R1 : j; }
    ADRT_HD static int out_row(int A) { return kRev ? A : (A % R2) * R1 + A / R2; }
    // storage skew of output angle A (multiple of 4)
    ADRT_HD static int out_skew(int A) { return ((A / R2) * (A % R2)) & ~3; }
    // rows of a butterfly: base + i * stride
    ADRT_HD static int s1_base(int k0) { return kRev ? k0 : k0 * R1; }
    static constexpr int s1_stride = kRev ? R2 : 1;
    ADRT_HD static int s2_base(int p) { return kRev ? p * R2 : p; }
    static constexpr int s2_stride = kRev ? 1 : R1;
};

// ===========================================================================
// forward butterfly on 4-vectors with register history
// ===========================================================================
template <int LOGR> struct FwdBfly {
    static constexpr int R = 1 << LOGR;
    // previous 4-vector of every right-hand child: level L has (R >> (L+1)) of them x 2^L angles = R/2
    float hist[LOGR][R / 2][4];

    ADRT_HD void clear()
    {
#pragma unroll
        for (int l = 0; l < LOGR; ++l)
#pragma unroll
            for (int k = 0; k < R / 2; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) hist[l][k][i] = 0.0f;
    }

    template <int L>
    ADRT_HD void level(const float (&cur)[R][4], float (&nxt)[R][4])
    {
        constexpr int nodes = R >> L, angles = 1 << L;
#pragma unroll
        for (int k = 0; k < nodes / 2; ++k)
#pragma unroll
            for (int a = 0; a < angles; ++a) {
                const int il = (2 * k) * angles + a, ir = (2 * k + 1) * angles + a, ih = k * angles + a;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int s = a + b;   // <= 4
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = 4 + i - s;
                        const float rv = idx >= 4 ? cur[ir][idx - 4] : hist[L][ih][idx];
                        nxt[k * 2 * angles + 2 * a + b][i] = cur[il][i] + rv;
                    }
                }
            }
#pragma unroll
        for (int k = 0; k < nodes / 2; ++k)
#pragma unroll
            for (int a = 0; a < angles; ++a)
#pragma unroll
                for (int i = 0; i < 4; ++i) hist[L][k * angles + a][i] = cur[(2 * k + 1) * angles + a][i];
    }

    // leaf[jj] = the 4 new offsets of leaf jj (pre-shifted frame); out[q] = the same 4 offsets of output q
    ADRT_HD void iterate(const float (&leaf)[R][4], float (&out)[R][4])
    {
        if constexpr (LOGR == 1) {
            level<0>(leaf, out);
        } else if constexpr (LOGR == 2) {
            float t[R][4];
            level<0>(leaf, t);
            level<1>(t, out);
        } else {
            float t[R][4], u[R][4];
            level<0>(leaf, t);
            level<1>(t, u);
            level<2>(u, out);
        }
    }
};

// Per-thread state that lives across the barrier between prologue and main loop.
template <int LOGR> struct FwdStepState {
    FwdBfly<LOGR> b;
    float nxt[1 << LOGR][4];
};

// state of the two steps of a pass (only one of them is live at any time)
template <int M> struct FwdState {
    FwdStepState<SGeo<M, true>::LOGR1> s1;
    FwdStepState<SGeo<M, true>::LOGR2> s2;
    BulkBar bar;
};

// Load the 4 offsets [c, c+4) (pre-shifted frame) of the R leaves of a butterfly whose
// row i starts at buf + (base + i*stride)*P; leaf jj is read p*jj positions lower.
template <int R, bool kShift>
ADRT_HD void fwd_load_leaves(const float *buf, int base, int stride, int p, int c, float (&leaf)[R][4])
{
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
        const float *rp = buf + (base + jj * stride) * P + c - (kShift ? p * jj : 0);
        if (!kShift || (jj & 3) == 0) {
            const F4 v = *reinterpret_cast<const F4 *>(rp);
#pragma unroll
            for (int i = 0; i < 4; ++i) leaf[jj][i] = v.v[i];
        } else if ((jj & 1) == 0) {
            const F2 v0 = *reinterpret_cast<const F2 *>(rp), v1 = *reinterpret_cast<const F2 *>(rp + 2);
            leaf[jj][0] = v0.v[0]; leaf[jj][1] = v0.v[1]; leaf[jj][2] = v1.v[0]; leaf[jj][3] = v1.v[1];
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) leaf[jj][i] = rp[i];
        }
    }
}

// A step of radix R = 2^LOGR for the segment [c0, c0 + 4*NIT) of one butterfly.
//   kShift: leaf jj is read at -p*jj and output q stored with skew floor4(p*q) (step 2)
//   warm:   the 8 offsets below c0 exist (false only for the segment that starts the row)
template <int LOGR, bool kShift>
ADRT_HD void fwd_step_prologue(const float *buf, int base, int stride, int p, int c0, bool warm, FwdStepState<LOGR> &st)
{
    constexpr int R = 1 << LOGR;
    st.b.clear();
    if (warm) {
        float leaf[R][4], out[R][4];
#pragma unroll
        for (int w = 2; w >= 1; --w) {
            fwd_load_leaves<R, kShift>(buf, base, stride, p, c0 - 4 * w, leaf);
            st.b.iterate(leaf, out);
        }
    }
    fwd_load_leaves<R, kShift>(buf, base, stride, p, c0, st.nxt);
}

template <int LOGR, bool kShift, int NIT>
ADRT_HD void fwd_step_main(float *buf, int base, int stride, int p, int c0, FwdStepState<LOGR> &st)
{
    constexpr int R = 1 << LOGR;
#pragma unroll kStreamUnroll
    for (int it = 0; it < NIT; ++it) {
        float cur[R][4], out[R][4];
#pragma unroll
        for (int jj = 0; jj < R; ++jj)
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[jj][i] = st.nxt[jj][i];
        if (it + 1 < NIT) fwd_load_leaves<R, kShift>(buf, base, stride, p, c0 + 4 * (it + 1), st.nxt);
        st.b.iterate(cur, out);
#pragma unroll
        for (int q = 0; q < R; ++q) {
            F4 v;
#pragma unroll
            for (int i = 0; i < 4; ++i) v.v[i] = out[q][i];
            const int f = kShift ? ((p * q) & ~3) : 0;
            *reinterpret_cast<F4 *>(buf + (base + q * stride) * P + c0 + 4 * it - f) = v;
        }
    }
}

// Last step of a pass that stores the public (d, column) layout: the R outputs of a
// butterfly are R consecutive columns, so the thread writes them straight to global memory
// (whole 32-byte sectors for R = 8, 16 bytes next to the neighbour group's 16 for R = 4)
// instead of going back through the tile.  `o` points at (offset of tile position c0, first
// column of the butterfly), dmax = number of offsets from there to the end of the plane.
template <int LOGR, int NIT>
ADRT_HD void fwd_step_main_direct(const float *buf, int base, int stride, int p, int c0, FwdStepState<LOGR> &st,
                                  float *o, long long n1, int dmax)
{
    constexpr int R = 1 << LOGR;
#pragma unroll kStreamUnroll
    for (int it = 0; it < NIT; ++it) {
        float cur[R][4], out[R][4];
#pragma unroll
        for (int jj = 0; jj < R; ++jj)
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[jj][i] = st.nxt[jj][i];
        if (it + 1 < NIT) fwd_load_leaves<R, true>(buf, base, stride, p, c0 + 4 * (it + 1), st.nxt);
        st.b.iterate(cur, out);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (4 * it + i < dmax) {
                static_assert(R == 8, "direct stores write 8 columns");
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = out[k][i];
                store8(o + (long long)(4 * it + i) * n1, v);
            }
        }
    }
}

// thread -> (butterfly, segment) of the two steps; returns false for idle threads
template <int M, bool kRev>
ADRT_HD bool fwd_s1_map(int tid, int &base, int &c0, bool &warm)
{
    typedef SGeo<M, true> Geo;
    const int grp = tid >> 3, seg = tid & 7;
    if (grp >= Geo::G / Geo::R1) return false;
    base = RowMap<M, kRev, true>::s1_base(grp);
    c0 = seg * Geo::SEG1;
    warm = seg > 0;
    return true;
}

template <int M, bool kRev>
ADRT_HD bool fwd_s2_map(int tid, int &base, int &p, int &c0)
{
    typedef SGeo<M, true> Geo;
    const int grp = tid >> 3, seg = tid & 7;
    if (grp >= Geo::G / Geo::R2 || seg >= Geo::NSEG2) return false;
    p = grp;    // one block group per tile: the butterfly's base angle is its index
    base = RowMap<M, kRev, true>::s2_base(p);
    c0 = Geo::X2 + seg * Geo::SEG2;
    return true;
}

// ===========================================================================
// forward loaders (row j of the group -> tile row leaf_row(j))
// ===========================================================================
// in_j[d0 - LH - a_g*j + x], x in [0, XW); see fused_tile.h fwd_load_wrows for the skew rule.
// Every stored position of a workspace row is meaningful: positions below the skew hold the
// -0.0 "copy" sentinels of negative offsets (they are computed like any output), positions above
// the support hold +0.0 up to the pitch.  So a row's segment is ONE bulk copy of the positions
// that exist, [max(0, gbase), min(pitch, gbase + XW)); tile cells that map below position 0 get
// -0.0 and cells beyond the pitch +0.0 (whole 16-byte chunks: gbase and the pitch are multiples of 4).
template <int M, bool kRev, int LH, int NT>
ADRT_HD void fwd_load_wrows(float *buf, const float *src_plane, const TileCtx &c, int tid, BulkBar &bar)
{
    constexpr int G = SGeo<M, true>::G;
    static_assert(G <= NT, "one row per thread");
    const int pitch = (int)c.in_pitch;
    int my_bytes = 0;
    bulk_fence();
    if (tid < G) {
        // thread j moves row j: one bulk copy, plus the fill of the cells that have no stored position
        const int j = tid;
        const float *row = src_plane + ((long long)(c.k0 * G + j) * c.e + c.a_g) * c.in_pitch;
        const int dbase = c.d0 - LH - c.a_g * j;
        const int gbase = dbase + fwd_row_skew(c.a_g, j);
        float *dst = buf + RowMap<M, kRev, true>::leaf_row(j) * P;
        int xlo = gbase < 0 ? -gbase : 0;            // first tile column whose position exists
        int xhi = pitch - gbase;                      // first tile column beyond the pitch
        if (xlo > XW) xlo = XW;
        if (xhi > XW) xhi = XW;
        if (xhi < xlo) xhi = xlo;
        if (xlo > 0) bulk_load(bar, dst, fill_src(true), xlo * 4);
        if (xhi > xlo) bulk_load(bar, dst + xlo, row + gbase + xlo, (xhi - xlo) * 4);
        if (xhi < XW) bulk_load(bar, dst + xhi, fill_src(false), (XW - xhi) * 4);
        my_bytes = XW * 4;
    }
    bulk_arrive_wait(bar, my_bytes);
}

// Image loader (first pass, e = 1, a_g = 0); orientations as in fused_tile.h fwd_load_image.
template <int M, bool kRev, int LH, int NWARP>
ADRT_HD void fwd_load_image(float *buf, const float *img, const TileCtx &c, int tid)
{
    constexpr int G = SGeo<M, true>::G;
    const int warp = tid >> 5, lane = tid & 31;
    const int n = c.n;
    const int dbase = c.d0 - LH;        // multiple of 4, like n: a 4-chunk lies inside [0, n) or outside
    const int rows = G < n ? G : n;     // multiple of 32 (n >= 32 for these passes)
    if (c.q == 0 || c.q == 3) {
        // offsets x..x+3 are image columns n-1-d-3 .. n-1-d: one aligned vector, reversed;
        // 4 rows (12 vectors) in flight per thread
        for (int j0 = warp * 4; j0 < rows; j0 += NWARP * 4) {
            F4 v[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = c.g * G + j0 + u;
                const float *row = img + (long long)(c.q == 0 ? r : n - 1 - r) * n;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int d = dbase + (k * 32 + lane) * V;
                    if (k * 32 + lane < NVEC && d >= 0 && d < n) v[u][k] = *reinterpret_cast<const F4 *>(row + (n - V - d));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float *dst = buf + RowMap<M, kRev, true>::leaf_row(j0 + u) * P;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int d = dbase + (k * 32 + lane) * V;
                    if (k * 32 + lane < NVEC) {
                        F4 r4;
                        const float fill = d < 0 ? -0.0f : 0.0f;
                        const bool in = d >= 0 && d < n;
#pragma unroll
                        for (int i = 0; i < V; ++i) r4.v[i] = in ? v[u][k].v[V - 1 - i] : fill;
                        *reinterpret_cast<F4 *>(dst + (k * 32 + lane) * V) = r4;
                    }
                }
            }
        }
    } else {
        // image rows are contiguous along r (the tile rows): I[r][d] = x[d, r] (q2) or x[n-1-d, r] (q1).
        // 4 x 4 blocks: a thread loads 4 image rows d..d+3 x 4 columns r..r+3 as four vectors,
        // transposes them in registers and stores four tile-row vectors.  Lane bits: b0, b3, b4 ->
        // row group (8 groups = 128 contiguous bytes per image row), b1, b2 -> offset chunk; the 8
        // lanes of a quarter warp then write 8 distinct 16-byte bank groups.
        constexpr int UB = LoadBatch<M>::value;
        const int rg_lo = (lane & 1) | (((lane >> 3) & 3) << 1), ch_lo = (lane >> 1) & 3;
        const int nblk = (rows / 32) * (NVEC / 4);
        const long long step = (c.q == 1) ? -(long long)n : (long long)n;
        const float *ib = (c.q == 1) ? img + (long long)(n - 1 - dbase) * n + c.g * G : img + (long long)dbase * n + c.g * G;
        for (int blk0 = warp; blk0 < nblk; blk0 += UB * NWARP) {
            F4 v[UB][4];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int blk = blk0 + u * NWARP;
                const int rg = (blk / (NVEC / 4)) * 8 + rg_lo, ch = (blk % (NVEC / 4)) * 4 + ch_lo;
                const int d = dbase + 4 * ch;
                if (blk < nblk && d >= 0 && d < n) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        v[u][i] = *reinterpret_cast<const F4 *>(ib + (long long)(4 * ch + i) * step + 4 * rg);
                }
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int blk = blk0 + u * NWARP;
                if (blk < nblk) {
                    const int rg = (blk / (NVEC / 4)) * 8 + rg_lo, ch = (blk % (NVEC / 4)) * 4 + ch_lo;
                    const int d = dbase + 4 * ch;
                    const float fill = d < 0 ? -0.0f : 0.0f;
                    const bool in = d >= 0 && d < n;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        F4 w;
#pragma unroll
                        for (int i = 0; i < 4; ++i) w.v[i] = in ? v[u][i].v[t] : fill;
                        *reinterpret_cast<F4 *>(buf + RowMap<M, kRev, true>::leaf_row(4 * rg + t) * P + 4 * ch) = w;
                    }
                }
            }
        }
    }
}

// ===========================================================================
// forward stores
// ===========================================================================
// R-layout workspace row of output angle p (see fused_tile.h fwd_store_wrows): the tile
// owns the aligned chunks [d0, d0 + TD); chunk position gp holds offsets gp - S .., which
// sit in the tile row at LH + (gp - d0) - S - skew.
template <int LH, int TD>
ADRT_HD void fwd_store_row_aligned(const float *b, float *row, const TileCtx &c, int lim, bool zero, int lane)
{
#pragma unroll
    for (int k = 0; k < (TD / V + 31) / 32; ++k) {
        const int xc = (k * 32 + lane) * V;
        const int gp = c.d0 + xc;
        if (xc < TD && gp < c.out_pitch) {
            F4 v;
            if (zero) {
#pragma unroll
                for (int i = 0; i < V; ++i) v.v[i] = 0.0f;
            } else {
                v = *reinterpret_cast<const F4 *>(b + LH + xc);
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (gp + i >= lim) v.v[i] = 0.0f;
            }
            *reinterpret_cast<F4 *>(row + gp) = v;
        }
    }
}

// skewed row (S = 1..3): chunk position gp holds offsets gp - S .., which sit in the tile row
// at LH + (gp - d0) - S: two aligned vectors and a compile-time selection
template <int LH, int TD, int S>
ADRT_HD void fwd_store_row_skewed(const float *b, float *row, const TileCtx &c, int lim, bool zero, int lane)
{
    constexpr int Q = (4 - S) & 3;
#pragma unroll
    for (int k = 0; k < (TD / V + 31) / 32; ++k) {
        const int xc = (k * 32 + lane) * V;
        const int gp = c.d0 + xc;
        if (xc < TD && gp < c.out_pitch) {
            float v[V];
            if (zero) {
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = 0.0f;
            } else {
                load_window_wide<V, Q>(b + (LH + xc - S - Q), v);
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (gp - S + i >= lim) v[i] = 0.0f;
            }
            tile::store_chunk<float>(row + gp, v);
        }
    }
}

template <int M, bool kRev, int LH, int TD, int NWARP>
ADRT_HD void fwd_store_wrows(const float *buf, float *dst_plane, const TileCtx &c, bool zero, int tid)
{
    constexpr int G = SGeo<M, true>::G;
    const int warp = tid >> 5, lane = tid & 31;
    for (int p = warp; p < G; p += NWARP) {
        float *row = dst_plane + ((long long)c.g * G + p) * c.out_pitch;
        const int ang = c.a_g * G + p;
        int lim = c.n + ang;
        if (lim > c.D) lim = c.D;
        const int s = c.next_g ? fwd_row_skew(ang, c.k0 & (c.next_g - 1)) : 0;
        const float *b = buf + RowMap<M, kRev, true>::out_row(p) * P - RowMap<M, kRev, true>::out_skew(p);
        switch (s) {
        case 0: fwd_store_row_aligned<LH, TD>(b, row, c, lim, zero, lane); break;
        case 1: fwd_store_row_skewed<LH, TD, 1>(b, row, c, lim, zero, lane); break;
        case 2: fwd_store_row_skewed<LH, TD, 2>(b, row, c, lim, zero, lane); break;
        default: fwd_store_row_skewed<LH, TD, 3>(b, row, c, lim, zero, lane); break;
        }
    }
}

// Public layout (D, n): column g*G + p, offsets d0 .. d0 + TD; lanes walk the rows p.
template <int M, bool kRev, int TD, int NWARP>
ADRT_HD void store_qcols(const float *buf, float *dst_plane, const TileCtx &c, int xoff, bool zero, int tid)
{
    constexpr int G = SGeo<M, true>::G;
    constexpr int NIT = (TD + NWARP * V - 1) / (NWARP * V);
    const int warp = tid >> 5, lane = tid & 31;
    const int cols = G < c.n ? G : c.n;
    const bool fast = c.d0 + TD <= c.D;
    const long long n1 = c.n;
    for (int p = lane; p < cols; p += 32) {
        float *o = dst_plane + (long long)(c.d0 + warp * V) * n1 + c.g * G + p;
        const float *b = buf + RowMap<M, kRev, true>::out_row(p) * P - RowMap<M, kRev, true>::out_skew(p) + xoff + warp * V;
        const long long ostep = (long long)NWARP * V * n1;
#pragma unroll 3
        for (int it = 0; it < NIT; ++it, o += ostep, b += NWARP * V) {
            const int xc = (it * NWARP + warp) * V;
            if (xc >= TD) break;
            float v[V];
            if (zero) {
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = 0.0f;
            } else {
                tile::load_window<float, V, 0>(b, v);
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
// tile program (forward)
// ===========================================================================
// Phases: 0 load | 1 step-1 prologue | 2 step-1 main | 3 step-2 prologue | 4 step-2 main | 5 store,
// with a CTA barrier after each.
// kSplit = 2: every (butterfly, segment) of a step is shared by TWO threads (lower / upper part of the
// segment, 5 + 4 or 4 + 3 vectors), 128 threads per tile: a tile occupies its shared memory for the copy
// latency plus the butterflies, and the butterflies are serial per thread
template <int M, int LOADK, int STOREK, int kSplit = 1>
struct FwdStream {
    static_assert(kSplit == 1 || kSplit == 2, "one or two threads per segment");
    typedef SGeo<M, true> Geo;
    static constexpr int G = Geo::G;
    static constexpr bool kRev = (STOREK == STORE_QCOLS);
    static constexpr bool kImage = (LOADK == LOAD_IMAGE);
    ADRT_HD static constexpr bool runs(int mode) { return mode != tile::TILE_SKIP; }   // forward tiles are never masked
    // passes that store the public layout write it from the registers of step 2 (no store phase)
    static constexpr bool kDirect = (STOREK == STORE_QCOLS);
    static constexpr int kPhases = kDirect ? 5 : 6;
    static constexpr int TD = STileTD<M, STOREK>::value;
    static constexpr int LH = XW - TD;
    // image loads go through registers (more threads = more loads in flight); workspace rows are
    // copied asynchronously, so those passes only need the 64 threads of the butterfly steps
    // 64 threads: the butterfly steps want ~200 registers, and a scheduler's 16K registers are shared by
    // the warps resident on it (3 CTAs x 2 warps = at most 2 warps per scheduler)
    static constexpr int NT = 64 * kSplit;
    static constexpr int NWARP = NT / 32;
    static constexpr int MIN_CTAS = (M == 6) ? 3 : 6;
    typedef FwdState<M> State;
    // a direct step 2 writes nothing to the tile: its prologue and main loop need no barrier between them
    ADRT_HD static constexpr bool barrier_after(int ph) { return !(kDirect && ph == 3); }
    // vectors of a segment that its first thread walks (the second one takes the rest)
    static constexpr int N1 = Geo::SEG1 / V, N1A = kSplit == 2 ? (N1 + 1) / 2 : N1, N1B = N1 - N1A;
    static constexpr int N2 = Geo::SEG2 / V, N2A = kSplit == 2 ? (N2 + 1) / 2 : N2, N2B = N2 - N2A;

    ADRT_HD static int classify(const TileCtx &c)
    {
        int sup = c.n + c.a_g * G + G - 1;
        if (sup > c.D) sup = c.D;
        if (STOREK == STORE_QCOLS) {
            if (c.d0 >= c.D) return tile::TILE_SKIP;
            if (c.d0 >= sup) return tile::TILE_ZERO;
        } else {
            if (c.d0 >= c.out_pitch) return tile::TILE_SKIP;
            if (c.d0 - 3 >= sup) return tile::TILE_ZERO;
        }
        return tile::TILE_FULL;
    }

    ADRT_HD static void zero_tile(float *buf, float *dst, const TileCtx &c, int tid)
    {
        if (STOREK == STORE_QCOLS) store_qcols<M, kRev, TD, NWARP>(buf, dst, c, 0, true, tid);
        else fwd_store_wrows<M, kRev, LH, TD, NWARP>(buf, dst, c, true, tid);
    }

    template <int PH>
    ADRT_HD static void phase_ct(int mode, float *buf, State &st, const float *src, float *dst, const TileCtx &c, int tid)
    {
        (void)mode;
        typedef RowMap<M, kRev, true> RM;
        if constexpr (PH == 0) {
            if (LOADK == LOAD_IMAGE) fwd_load_image<M, kRev, LH, NWARP>(buf, src, c, tid);
            else fwd_load_wrows<M, kRev, LH, NT>(buf, src, c, tid, st.bar);
        } else if constexpr (PH == 1 || PH == 2) {
            int base, c0;
            bool warm;
            const int half = kSplit == 2 ? tid >> 6 : 0;    // warp uniform
            if (!fwd_s1_map<M, kRev>(tid & 63, base, c0, warm)) return;
            if (half) { c0 += V * N1A; warm = true; }
            if constexpr (PH == 1) fwd_step_prologue<Geo::LOGR1, false>(buf, base, RM::s1_stride, 0, c0, warm, st.s1);
            else if (kSplit == 1 || !half) fwd_step_main<Geo::LOGR1, false, N1A>(buf, base, RM::s1_stride, 0, c0, st.s1);
            else if constexpr (kSplit == 2) fwd_step_main<Geo::LOGR1, false, N1B>(buf, base, RM::s1_stride, 0, c0, st.s1);
        } else if constexpr (PH == 3 || PH == 4) {
            int base, p, c0;
            const int half = kSplit == 2 ? tid >> 6 : 0;
            if (!fwd_s2_map<M, kRev>(tid & 63, base, p, c0)) return;
            if (half) c0 += V * N2A;
            if constexpr (PH == 3) {
                fwd_step_prologue<Geo::LOGR2, true>(buf, base, RM::s2_stride, p, c0, true, st.s2);
            } else if constexpr (kDirect) {
                // tile position c0 is offset d0 + (c0 - LH); the butterfly's columns start at g*G + p*R2
                const int d = c.d0 + (c0 - LH);
                float *o = dst + (long long)d * c.n + c.g * G + p * Geo::R2;
                if (kSplit == 1 || !half) fwd_step_main_direct<Geo::LOGR2, N2A>(buf, base, RM::s2_stride, p, c0, st.s2, o, c.n, c.D - d);
                else if constexpr (kSplit == 2) fwd_step_main_direct<Geo::LOGR2, N2B>(buf, base, RM::s2_stride, p, c0, st.s2, o, c.n, c.D - d);
            } else {
                if (kSplit == 1 || !half) fwd_step_main<Geo::LOGR2, true, N2A>(buf, base, RM::s2_stride, p, c0, st.s2);
                else if constexpr (kSplit == 2) fwd_step_main<Geo::LOGR2, true, N2B>(buf, base, RM::s2_stride, p, c0, st.s2);
            }
        } else {
            if (STOREK == STORE_QCOLS) store_qcols<M, kRev, TD, NWARP>(buf, dst, c, LH, false, tid);
            else fwd_store_wrows<M, kRev, LH, TD, NWARP>(buf, dst, c, false, tid);
        }
    }
};

// ===========================================================================
// transposed (bdrt) butterfly: exact transpose of FwdBfly
// ===========================================================================
//   g(L, 2k,   a)[x] = g(L+1, k, 2a)[x]     + g(L+1, k, 2a+1)[x]
//   g(L, 2k+1, a)[x] = g(L+1, k, 2a)[x + a] + g(L+1, k, 2a+1)[x + a + 1]
// (first operand = even parent angle, adrt_cdefs_bdrt.hpp:216-237).  The shifts look AHEAD,
// so the segment is walked downwards and the history is the previous, higher 4-vector.
// Masked instantiation (tiles that reach offset D): an operand read at a position at or beyond
// the end `thr` of its node is absent and acts as +0.0 (first operand) / -0.0 (second operand),
// the copy-vs-add rule of adrt_cdefs_bdrt.hpp:96-109.
template <int LOGR> struct BwdBfly {
    static constexpr int R = 1 << LOGR;
    float hist[LOGR][R][4];   // hist[L]: higher 4-vector of the level-(L+1) nodes; only the first a+b entries are live

    ADRT_HD void clear()
    {
#pragma unroll
        for (int l = 0; l < LOGR; ++l)
#pragma unroll
            for (int k = 0; k < R; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) hist[l][k][i] = 0.0f;
    }

    // thr0 + k * thrk = end (in the coordinates of x0) of node k of level L+1
    template <int L, bool kMask>
    ADRT_HD void level(const float (&cur)[R][4], float (&nxt)[R][4], int x0, int thr0, int thrk)
    {
        constexpr int nodes1 = R >> (L + 1), angles = 1 << L, angles1 = 2 << L;
#pragma unroll
        for (int k = 0; k < nodes1; ++k) {
            const int thr = kMask ? thr0 + k * thrk : 0;
#pragma unroll
            for (int a = 0; a < angles; ++a) {
                const int iA = k * angles1 + 2 * a, iB = iA + 1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float vA = cur[iA][i], vB = cur[iB][i];
                    if (kMask && x0 + i >= thr) { vA = 0.0f; vB = -0.0f; }
                    nxt[(2 * k) * angles + a][i] = vA + vB;
                    const int jA = i + a, jB = i + a + 1;
                    float wA = jA < 4 ? cur[iA][jA] : hist[L][iA][jA - 4];
                    float wB = jB < 4 ? cur[iB][jB] : hist[L][iB][jB - 4];
                    if (kMask && x0 + jA >= thr) wA = 0.0f;
                    if (kMask && x0 + jB >= thr) wB = -0.0f;
                    nxt[(2 * k + 1) * angles + a][i] = wA + wB;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) hist[L][k][i] = cur[k][i];
    }

    // par[q] = the 4 offsets [x0, x0+4) of parent angle q; out[jj] = the same offsets of child jj
    // (each in its own frame).  thr_top = end of the parents; a node k of level L ends at
    // thr_top + k * 2^L * thr_leaf.
    template <bool kMask>
    ADRT_HD void iterate(const float (&par)[R][4], float (&out)[R][4], int x0, int thr_top, int thr_leaf)
    {
        if constexpr (LOGR == 1) {
            level<0, kMask>(par, out, x0, thr_top, 2 * thr_leaf);
        } else if constexpr (LOGR == 2) {
            float t[R][4];
            level<1, kMask>(par, t, x0, thr_top, 4 * thr_leaf);
            level<0, kMask>(t, out, x0, thr_top, 2 * thr_leaf);
        } else {
            float t[R][4], u[R][4];
            level<2, kMask>(par, t, x0, thr_top, 8 * thr_leaf);
            level<1, kMask>(t, u, x0, thr_top, 4 * thr_leaf);
            level<0, kMask>(u, out, x0, thr_top, 2 * thr_leaf);
        }
    }
};

template <int LOGR> struct BwdStepState {
    BwdBfly<LOGR> b;
    float nxt[1 << LOGR][4];
};

template <int M> struct BwdState {
    BwdStepState<SGeo<M, false>::LOGR2> sa;   // first transposed step: radix 2^(M-3), local stages M-1 .. 3
    BwdStepState<SGeo<M, false>::LOGR1> sb;   // second: radix 8, local stages 2 .. 0
    BulkBar bar;
};

// Storage conventions of the transposed tile (mirror image of the forward ones):
//   parents of step A      : logical column x at x (as loaded)
//   children of step A     : child (block jj, angle p), logical column c = x - p*jj, at c + floor4(p*jj)
//                            = x - ((p*jj) & 3): mixed-width stores (128 / 64 / 32 bit by jj), never
//                            above the position its row was read from and at most 3 below it
//   step B reads parent q of block k0 at c + floor4(q*k0) (aligned) and writes leaf jj of block k0 at
//   c + floor4(jj*k0) (aligned, the very positions it read from row jj)
template <int R, bool kB>
ADRT_HD void bwd_load_parents(const float *buf, int base, int stride, int k0, int x, float (&par)[R][4])
{
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const float *rp = buf + (base + q * stride) * P + x + (kB ? ((q * k0) & ~3) : 0);
        const F4 v = *reinterpret_cast<const F4 *>(rp);
#pragma unroll
        for (int i = 0; i < 4; ++i) par[q][i] = v.v[i];
    }
}

template <int R, bool kB>
ADRT_HD void bwd_store_children(float *buf, int base, int stride, int pk, int x, const float (&out)[R][4])
{
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
        float *rp = buf + (base + jj * stride) * P;
        if (kB) {
            F4 v;
#pragma unroll
            for (int i = 0; i < 4; ++i) v.v[i] = out[jj][i];
            *reinterpret_cast<F4 *>(rp + x + ((jj * pk) & ~3)) = v;
        } else {
            const int col = x - ((pk * jj) & 3);     // may be negative only for x = 0: those columns are nobody's
            if ((jj & 3) == 0) {
                F4 v;
#pragma unroll
                for (int i = 0; i < 4; ++i) v.v[i] = out[jj][i];
                *reinterpret_cast<F4 *>(rp + col) = v;
            } else if ((jj & 1) == 0) {
                F2 v0, v1;
                v0.v[0] = out[jj][0]; v0.v[1] = out[jj][1]; v1.v[0] = out[jj][2]; v1.v[1] = out[jj][3];
                if (col >= 0) *reinterpret_cast<F2 *>(rp + col) = v0;
                *reinterpret_cast<F2 *>(rp + col + 2) = v1;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i == 3 || col + i >= 0) rp[col + i] = out[jj][i];
            }
        }
    }
}

// Transposed step for the segment [c0, c0 + 4*NIT) of one butterfly, walked downwards.
//   kB = false: step A (group base angle pk = p);  kB = true: step B (pk = k0, block index)
//   warm: the 8 columns above the segment exist (false for the segment that ends the row)
template <int LOGR, bool kB, int NIT>
ADRT_HD void bwd_step_prologue(const float *buf, int base, int stride, int pk, int c0, bool warm, BwdStepState<LOGR> &st)
{
    constexpr int R = 1 << LOGR;
    st.b.clear();
    const int top = c0 + 4 * NIT;
    if (warm) {
        float par[R][4], out[R][4];
#pragma unroll
        for (int w = 1; w >= 0; --w) {
            bwd_load_parents<R, kB>(buf, base, stride, pk, top + 4 * w, par);
            st.b.template iterate<false>(par, out, 0, 0, 0);
        }
    }
    bwd_load_parents<R, kB>(buf, base, stride, pk, top - 4, st.nxt);
}

template <int LOGR, bool kB, bool kMask, int NIT>
ADRT_HD void bwd_step_main(float *buf, int base, int stride, int pk, int c0, BwdStepState<LOGR> &st, int thr_top, int thr_leaf)
{
    constexpr int R = 1 << LOGR;
#pragma unroll kStreamUnroll
    for (int it = NIT - 1; it >= 0; --it) {
        float cur[R][4], out[R][4];
#pragma unroll
        for (int q = 0; q < R; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[q][i] = st.nxt[q][i];
        if (it > 0) bwd_load_parents<R, kB>(buf, base, stride, pk, c0 + 4 * (it - 1), st.nxt);
        st.b.template iterate<kMask>(cur, out, c0 + 4 * it, thr_top, thr_leaf);
        bwd_store_children<R, kB>(buf, base, stride, pk, c0 + 4 * it, out);
    }
}

// Step B of a pass that stores the public layout: the 8 leaves of a butterfly are 8 consecutive
// columns; `o` points at (offset of tile column c0, first column), dmax = offsets from there to D.
template <int LOGR, bool kMask, int NIT>
ADRT_HD void bwd_step_main_direct(const float *buf, int base, int stride, int pk, int c0, BwdStepState<LOGR> &st,
                                  int thr_top, int thr_leaf, float *o, long long n1, int dmax)
{
    constexpr int R = 1 << LOGR;
#pragma unroll kStreamUnroll
    for (int it = NIT - 1; it >= 0; --it) {
        float cur[R][4], out[R][4];
#pragma unroll
        for (int q = 0; q < R; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[q][i] = st.nxt[q][i];
        if (it > 0) bwd_load_parents<R, true>(buf, base, stride, pk, c0 + 4 * (it - 1), st.nxt);
        st.b.template iterate<kMask>(cur, out, c0 + 4 * it, thr_top, thr_leaf);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (4 * it + i < dmax) {
                static_assert(R == 8, "direct stores write 8 columns");
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = out[k][i];
                store8(o + (long long)(4 * it + i) * n1, v);
            }
        }
    }
}

// ---- transposed loaders ---------------------------------------------------------------
// Public layout: tile row of parent angle A <- column g*G + A, tile column x <- offset d0 + x
// (zero from offset D on).  4 x 4 register transposes as in fwd_load_image.
template <int M, bool kRev, int NWARP, bool kEdge>
ADRT_HD void bwd_load_qcols(float *buf, const float *src_plane, const TileCtx &c, int tid)
{
    constexpr int G = SGeo<M, false>::G;
    const int warp = tid >> 5, lane = tid & 31;
    const int n = c.n;
    const int cols = G < n ? G : n;
    constexpr int UB = LoadBatch<M>::value;
    const int rg_lo = (lane & 1) | (((lane >> 3) & 3) << 1), ch_lo = (lane >> 1) & 3;
    const int nblk = (cols / 32) * (NVEC / 4);
    const float *ib = src_plane + (long long)c.d0 * n + c.g * G + 4 * rg_lo + 4 * ch_lo * n;   // in-plane offsets fit 32 bits
    const int dt = c.D - c.d0;
    for (int blk0 = warp; blk0 < nblk; blk0 += UB * NWARP) {
        F4 v[UB][4];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int blk = blk0 + u * NWARP;
            const int rb = blk / (NVEC / 4), cb = blk % (NVEC / 4);   // 8 row groups x 4 chunks per block
            if (blk < nblk) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!kEdge || 16 * cb + 4 * ch_lo + i < dt) {
                        v[u][i] = *reinterpret_cast<const F4 *>(ib + (16 * cb + i) * n + 32 * rb);
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t) v[u][i].v[t] = 0.0f;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int blk = blk0 + u * NWARP;
            if (blk < nblk) {
                const int rg = (blk / (NVEC / 4)) * 8 + rg_lo, ch = (blk % (NVEC / 4)) * 4 + ch_lo;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    F4 w;
#pragma unroll
                    for (int i = 0; i < 4; ++i) w.v[i] = v[u][i].v[t];
                    *reinterpret_cast<F4 *>(buf + RowMap<M, kRev, false>::out_row(4 * rg + t) * P + 4 * ch) = w;
                }
            }
        }
    }
}

// Workspace rows (dense, pitch round4(D), offsets >= D never written): parent A <- row g*G + A,
// columns [d0, d0 + XW): one bulk copy of the whole chunks below D, scalar tail, zeros above.
template <int M, bool kRev, int NT>
ADRT_HD void bwd_load_wrows(float *buf, const float *src_plane, const TileCtx &c, int tid, BulkBar &bar)
{
    constexpr int G = SGeo<M, false>::G;
    static_assert(G <= NT, "one row per thread");
    int my_bytes = 0;
    bulk_fence();
    if (tid < G) {
        const int A = tid;
        const long long r = (long long)c.g * G + A;
        const float *row = src_plane + r * c.in_pitch + c.d0;
        float *dst = buf + RowMap<M, kRev, false>::out_row(A) * P;
        int xv = tile::bwd_row_support(c, r) - c.d0;   // columns that exist (the producer may have skipped the row's zero tail)
        if (xv > XW) xv = XW;
        if (xv < 0) xv = 0;
        const int xb = xv & ~3;          // whole chunks
        if (xb > 0) {
            bulk_load(bar, dst, row, xb * 4);
            my_bytes = xb * 4;
        }
        if (xb < XW) {
            // the chunk that straddles D by hand, whole chunks above it from the zero source
            const int xz = xb + 4 < XW ? xb + 4 : XW;
            for (int x = xb; x < xz; ++x) dst[x] = x < xv ? row[x] : 0.0f;
            if (xz < XW) {
                bulk_load(bar, dst + xz, fill_src(false), (XW - xz) * 4);
                my_bytes += (XW - xz) * 4;
            }
        }
    }
    bulk_arrive_wait(bar, my_bytes);
}

// ---- transposed stores ----------------------------------------------------------------
// Workspace row of leaf j: (k0*G + j)*e + a_g; tile column xc is offset d0 - a_g*j + xc.  The tile owns
// the aligned chunks [ceil4(dbase), ceil4(dbase) + TD) of the row; Q = (a_g*j) & 3 is the residue of the
// tile-side window (see fused_tile.h bwd_store_row).  `b` already includes the leaf's storage skew.
template <int TD, int Q>
ADRT_HD void bwd_store_row(const float *b, float *row, const TileCtx &c, int dbase, bool zero, int lane)
{
    if (!zero && dbase + Q >= 0 && dbase + Q + TD <= c.D) {
        // the whole segment lies inside the row
        float *o = row + dbase + Q;
#pragma unroll
        for (int k = 0; k < (TD / V + 31) / 32; ++k) {
            const int xa = (k * 32 + lane) * V;
            if (xa < TD) {
                float v[V];
                load_window_wide<V, Q>(b + xa, v);
                tile::store_chunk<float>(o + xa, v);
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < (TD / V + 31) / 32; ++k) {
        const int xa = (k * 32 + lane) * V;
        const int gp = dbase + Q + xa;
        if (xa < TD && gp + V > 0 && gp < c.D) {
            float v[V];
            if (zero) {
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = 0.0f;
            } else {
                load_window_wide<V, Q>(b + xa, v);
            }
            if (gp >= 0 && gp + V <= c.D) {
                tile::store_chunk<float>(row + gp, v);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (gp + i >= 0 && gp + i < c.D) row[gp + i] = v[i];
            }
        }
    }
}

template <int M, bool kRev, int TD, int NWARP>
ADRT_HD void bwd_store_wrows(const float *buf, float *dst_plane, const TileCtx &c, bool zero, int tid)
{
    constexpr int G = SGeo<M, false>::G, R1 = SGeo<M, false>::R1;
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < G; j += NWARP) {
        float *row = dst_plane + ((long long)(c.k0 * G + j) * c.e + c.a_g) * c.out_pitch;
        const int dbase = c.d0 - c.a_g * j;
        const int k0 = j / R1, jj = j % R1;
        const float *b = buf + RowMap<M, kRev, false>::leaf_row(j) * P + ((jj * k0) & ~3);
        switch ((c.a_g * j) & 3) {
        case 0: bwd_store_row<TD, 0>(b, row, c, dbase, zero, lane); break;
        case 1: bwd_store_row<TD, 1>(b, row, c, dbase, zero, lane); break;
        case 2: bwd_store_row<TD, 2>(b, row, c, dbase, zero, lane); break;
        default: bwd_store_row<TD, 3>(b, row, c, dbase, zero, lane); break;
        }
    }
}

// thread -> (butterfly, segment)
template <int M, bool kRev>
ADRT_HD bool bwd_sa_map(int tid, int &base, int &p, int &c0, bool &warm)
{
    typedef SGeo<M, false> Geo;
    const int grp = tid >> 3, seg = tid & 7;
    if (grp >= Geo::G / Geo::R2) return false;
    p = grp;
    base = RowMap<M, kRev, false>::s2_base(p);
    c0 = seg * Geo::SEG1;        // step A walks the whole row: 8 segments of 36
    warm = seg < 7;
    return true;
}

template <int M, bool kRev>
ADRT_HD bool bwd_sb_map(int tid, int &base, int &k0, int &c0)
{
    typedef SGeo<M, false> Geo;
    const int grp = tid >> 3, seg = tid & 7;
    if (grp >= Geo::G / Geo::R1 || seg >= Geo::NSEG2) return false;
    k0 = grp;
    base = RowMap<M, kRev, false>::s1_base(k0);
    c0 = seg * Geo::SEG2;        // step B produces the columns [0, TD): NSEG2 segments of SEG2
    return true;
}

// ===========================================================================
// tile program (transposed)
// ===========================================================================
// Phases: 0 load | 1 step-A prologue | 2 step-A main | 3 step-B prologue | 4 step-B main | 5 store
template <int M, int LOADK, int STOREK, bool kMaskTiles>
struct BwdStream {
    typedef SGeo<M, false> Geo;
    static constexpr int G = Geo::G;
    // parents arrive in natural row order from the public layout (transposing load); workspace rows may
    // be placed anywhere, and then the leaves come out in natural order for the public-layout store
    static constexpr bool kRev = (LOADK == LOAD_QCOLS);
    static constexpr bool kImage = false;
    // the masked program also writes the all-zero tiles beyond offset D
    ADRT_HD static constexpr bool runs(int mode) { return kMaskTiles ? (mode == tile::TILE_FULL_MASKED || mode == tile::TILE_ZERO) : mode == tile::TILE_FULL; }
    static constexpr bool kDirect = (STOREK == STORE_QCOLS);
    static constexpr int kPhases = kDirect ? 5 : 6;
    static constexpr int TD = STileTD<M, STOREK>::value;
    static constexpr int TDB = XW - Geo::X2;            // columns step B produces (TD, or TD + 4 for workspace stores)
    static constexpr int NT = 64;
    static constexpr int NWARP = NT / 32;
    static constexpr int MIN_CTAS = (M == 6) ? 3 : 6;
    typedef BwdState<M> State;
    ADRT_HD static constexpr bool barrier_after(int ph) { return !(kDirect && ph == 3); }

    ADRT_HD static int classify(const TileCtx &c)
    {
        if (c.d0 >= c.d_need + c.a_g * (G - 1)) return tile::TILE_SKIP;
        if (c.d0 >= c.D) return STOREK == STORE_QCOLS ? tile::TILE_SKIP : tile::TILE_ZERO;
        if (c.d0 + XW + 8 > c.D) return tile::TILE_FULL_MASKED;
        return tile::TILE_FULL;
    }

    ADRT_HD static void zero_tile(float *buf, float *dst, const TileCtx &c, int tid)
    {
        bwd_store_wrows<M, kRev, TD, NWARP>(buf, dst, c, true, tid);
    }

    template <int PH, bool kMask>
    ADRT_HD static void phase_m(float *buf, State &st, const float *src, float *dst, const TileCtx &c, int tid)
    {
        typedef RowMap<M, kRev, false> RM;
        const int dt = c.D - c.d0;
        if constexpr (PH == 0) {
            if (LOADK == LOAD_QCOLS) bwd_load_qcols<M, kRev, NWARP, kMask>(buf, src, c, tid);
            else bwd_load_wrows<M, kRev, NT>(buf, src, c, tid, st.bar);
        } else if constexpr (PH == 1 || PH == 2) {
            int base, p, c0;
            bool warm;
            if (!bwd_sa_map<M, kRev>(tid, base, p, c0, warm)) return;
            // node k of a level whose nodes span 2^L step leaves ends at dt + a_g*(8*2^L*k) + p*(2^L*k):
            // per step leaf (= a block of 8 tile leaves) a_g*8 + p
            if constexpr (PH == 1) bwd_step_prologue<Geo::LOGR2, false, Geo::SEG1 / V>(buf, base, RM::s2_stride, p, c0, warm, st.sa);
            else bwd_step_main<Geo::LOGR2, false, kMask, Geo::SEG1 / V>(buf, base, RM::s2_stride, p, c0, st.sa, dt, c.a_g * Geo::R1 + p);
        } else if constexpr (PH == 3 || PH == 4) {
            int base, k0, c0;
            if (!bwd_sb_map<M, kRev>(tid, base, k0, c0)) return;
            // block k0 lives in the frame of its first leaf 8*k0: its nodes end at dt + a_g*(8*k0 + 2^L*k)
            const int thr_top = dt + c.a_g * (Geo::R1 * k0);
            if constexpr (PH == 3) {
                bwd_step_prologue<Geo::LOGR1, true, Geo::SEG2 / V>(buf, base, RM::s1_stride, k0, c0, true, st.sb);
            } else if constexpr (kDirect) {
                float *o = dst + (long long)(c.d0 + c0) * c.n + c.g * G + k0 * Geo::R1;
                bwd_step_main_direct<Geo::LOGR1, kMask, Geo::SEG2 / V>(buf, base, RM::s1_stride, k0, c0, st.sb, thr_top, c.a_g,
                                                                       o, c.n, dt - c0);
            } else {
                bwd_step_main<Geo::LOGR1, true, kMask, Geo::SEG2 / V>(buf, base, RM::s1_stride, k0, c0, st.sb, thr_top, c.a_g);
            }
        } else {
            bwd_store_wrows<M, kRev, TD, NWARP>(buf, dst, c, false, tid);
        }
    }

    // kMaskTiles programs run the tiles that reach offset D (TILE_FULL_MASKED), the others the interior
    // tiles: two launches, so that neither kernel carries both copies of the unrolled steps
    template <int PH>
    ADRT_HD static void phase_ct(int mode, float *buf, State &st, const float *src, float *dst, const TileCtx &c, int tid)
    {
        (void)mode;
        phase_m<PH, kMaskTiles>(buf, st, src, dst, c, tid);
    }
    // first d-tile that must run masked
    ADRT_HD static int first_masked_tile(int D)
    {
        const int lim = D - XW - 8;
        return lim < 0 ? 0 : lim / TD + 1;
    }
};

template <typename Prog, int PH = 0>
ADRT_HD void run_phase(int ph, int mode, float *buf, typename Prog::State &st, const float *src, float *dst, const TileCtx &c, int tid)
{
    if constexpr (PH < Prog::kPhases) {
        if (ph == PH) Prog::template phase_ct<PH>(mode, buf, st, src, dst, c, tid);
        else run_phase<Prog, PH + 1>(ph, mode, buf, st, src, dst, c, tid);
    }
}

}  // namespace stile
}  // namespace adrt_b200
