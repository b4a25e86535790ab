// "Staged" tile programs: the two five-stage fp32 passes that talk to the PUBLIC layout on their
// input side (forward: images; transposed: the (d, column) sinogram) as streaming passes whose
// input tile is brought into shared memory by the TMA engine instead of by the threads.
//
// Why (profiles/r01_pass_phases.txt, r02_ncu_full_summary.txt): the fused_tile.h kernels that ran
// these two pass kinds are bound by the SM's L1/shared-memory pipe (l1tex 80-86 %, 4-7 issued
// instructions per useful add), and half of their stall samples sit in the load phase; the
// streaming kernels of stream_tile.h need half the instructions and shared-memory wavefronts per
// stage, but their register-transposing public-layout loaders (64 threads, three rounds of DRAM
// latency per tile) made them slower for exactly these two kinds.  Here
//
//   * the input tile arrives as it lies in global memory -- [offset][column] boxes of 32 columns
//     (128-byte rows) by 2-D tensor-map copies (cp.async.bulk.tensor, one elected thread, no
//     registers, no LSU wavefronts), or, for the row-oriented image quadrants, as 32 one-row bulk
//     copies -- into a STAGING buffer;
//   * the first butterfly step reads its leaves from the staging buffer (the four leaves /
//     parents of a radix-4 butterfly are four adjacent columns = ONE 16-byte vector per offset,
//     so the transposition costs nothing) and writes the [row][offset] work tile: out of place,
//     so the step needs no prologue barrier;
//   * the second step is out of place too, back into the buffer the input came in, and it leaves
//     every output row at the alignment its workspace row has in global memory (the realignment is
//     register renaming between the thread's previous and current 4-vector, selected at compile
//     time by the tile's residue class); the store "phase" is then ONE bulk copy per row
//     (cp.async.bulk shared -> global) instead of a 16-byte-window walk by the threads, which was
//     half of either kernel's time (profiles/s5_ncu_staged_v1.txt);
//   * the CTA is persistent and the two buffers swap roles from tile to tile: the copies of the
//     CTA's NEXT tile are issued into the buffer the second step has just finished reading while
//     the bulk stores of the current tile drain from the other one.
//
// The butterflies, the tile geometry (XW = 288, TD = 248), the workspace formats and the tile
// classification are those of stream_tile.h (FwdStream<5, IMAGE, WROWS> /
// BwdStream<5, QCOLS, WROWS>), which these programs inherit from: same adds in the same order.
#pragma once

#include <utility>

#include "stream_tile.h"

namespace adrt_b200 {
namespace sgtile {

using stile::BulkBar;
using stile::F4;
using stile::P;
using stile::XW;
using stile::SGeo;
using stile::RowMap;
using tile::TileCtx;

constexpr int SW = 32;            // columns of a staging row (128 bytes)
constexpr int BOX_ROWS = 144;     // a tile is two boxes of 144 rows (box dimensions are limited to 256)
constexpr int STG_FLOATS = 32 * P;   // staging buffer: [XW][SW] boxes (36864 B) or 32 rows of pitch P (37376 B)
constexpr int kUnroll = stile::kStreamUnroll;

// A rank-3 tensor (columns contiguous, rows, planes) as the TMA engine and as the host emulator see it.
struct alignas(64) TmaMap {
    unsigned long long opaque[16];   // CUtensorMap, encoded by the host (stage_adrt.cu); box = SW x BOX_ROWS x 1
    const float *base;               // the same tensor for the host emulator
    int dim0, dim1, dim2;            // columns, rows, planes
    long long stride1, stride2;      // elements between rows / planes
};

// box (SW columns from c0) x (BOX_ROWS rows from c1) of plane c2 -> dst[BOX_ROWS][SW]; elements outside
// the tensor arrive as +0.0
ADRT_HD void tma_load_box(const BulkBar &b, float *dst_smem, const TmaMap *m, int c0, int c1, int c2)
{
#ifdef __CUDA_ARCH__
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const unsigned a = (unsigned)__cvta_generic_to_shared(b.bar);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 ::"r"(d), "l"(reinterpret_cast<const void *>(m->opaque)), "r"(c0), "r"(c1), "r"(c2), "r"(a) : "memory");
#else
    (void)b;
    for (int r = 0; r < BOX_ROWS; ++r)
        for (int x = 0; x < SW; ++x) {
            const int row = c1 + r, col = c0 + x;
            const bool in = row >= 0 && row < m->dim1 && col >= 0 && col < m->dim0 && c2 >= 0 && c2 < m->dim2;
            dst_smem[r * SW + x] = in ? m->base[(long long)c2 * m->stride2 + (long long)row * m->stride1 + col] : 0.0f;
        }
#endif
}

// every thread of the tile announces the bytes of the copies it issued ...
ADRT_HD void bulk_arrive(const BulkBar &b, int my_bytes)
{
#ifdef __CUDA_ARCH__
    const unsigned a = (unsigned)__cvta_generic_to_shared(b.bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(my_bytes) : "memory");
#else
    (void)b; (void)my_bytes;
#endif
}

// ... and waits for all of them to land, possibly much later
ADRT_HD void bulk_wait(BulkBar &b)
{
#ifdef __CUDA_ARCH__
    const unsigned a = (unsigned)__cvta_generic_to_shared(b.bar);
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
    (void)b;
#endif
}

// shared -> global bulk copy (`bytes` a multiple of 16, both addresses 16-byte aligned), in the thread's
// current bulk group
ADRT_HD void bulk_store(float *dst_global, const float *src_smem, int bytes)
{
#ifdef __CUDA_ARCH__
    const unsigned s = (unsigned)__cvta_generic_to_shared(src_smem);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst_global), "r"(s), "r"(bytes) : "memory");
#else
    for (int i = 0; i < bytes / 4; ++i) dst_global[i] = src_smem[i];
#endif
}
ADRT_HD void bulk_store_commit()
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
#endif
}
// until the thread's bulk stores have READ their shared-memory source (which may then be overwritten)
ADRT_HD void bulk_store_wait_read()
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
#endif
}

// thread -> (butterfly, segment) of the first step: the 8 lanes of a quarter warp take the 8 butterflies
// (8 adjacent 16-byte vectors of a staging row) and segments chosen so that their stores into the work
// tile (rows 4 apart = 16 banks, segments 36 apart = 4 banks) hit 8 distinct bank groups
ADRT_HD void first_step_map(int tid, int &bf, int &seg)
{
    const int l = tid & 7, k = tid >> 3;
    bf = l;
    seg = ((l >> 1) + k) & 7;
}

// Out-of-place second step: row r of the output buffer holds tile coordinate x at position x + DL
// (DL = 0..3, compile time): the aligned chunk [p0, p0 + 4) is the last DL elements of the vector
// below it and the first 4 - DL of the vector at p0.  A thread walks its segment upwards (kUp) or
// downwards and owns exactly the cells of its own coordinates: the chunk it shares with the
// neighbouring segment is written as single elements by the two threads.
//   lo / hi: the vector at the lower / higher coordinates of the two that meet in the chunk
//   at_start: first iteration of the segment (the other vector belongs to the neighbour)
template <int DL, bool kUp>
ADRT_HD void oop_store_vec(float *row, int x, const float (&cur)[4], const float (&prev)[4], bool at_start)
{
    if constexpr (DL == 0) {
        F4 v;
#pragma unroll
        for (int i = 0; i < 4; ++i) v.v[i] = cur[i];
        *reinterpret_cast<F4 *>(row + x) = v;
    } else if constexpr (kUp) {
        // chunk at x: coordinates x - DL .. x - DL + 3 = prev[4 - DL ..], cur[.. 3 - DL]
        if (at_start) {
#pragma unroll
            for (int e = DL; e < 4; ++e) row[x + e] = cur[e - DL];
        } else {
            F4 v;
#pragma unroll
            for (int e = 0; e < 4; ++e) v.v[e] = e < DL ? prev[4 - DL + e] : cur[e - DL];
            *reinterpret_cast<F4 *>(row + x) = v;
        }
    } else {
        // chunk at x + 4: coordinates x + 4 - DL .. x + 7 - DL = cur[4 - DL ..], prev[.. 3 - DL]
        if (at_start) {
#pragma unroll
            for (int e = 0; e < DL; ++e) row[x + 4 + e] = cur[4 - DL + e];
        } else {
            F4 v;
#pragma unroll
            for (int e = 0; e < 4; ++e) v.v[e] = e < DL ? cur[4 - DL + e] : prev[e - DL];
            *reinterpret_cast<F4 *>(row + x + 4) = v;
        }
    }
}

// the elements of the segment's last vector that no chunk store has written yet
template <int DL, bool kUp>
ADRT_HD void oop_store_rest(float *row, int x, const float (&last)[4])
{
    if constexpr (DL != 0) {
        if constexpr (kUp) {
#pragma unroll
            for (int e = 0; e < DL; ++e) row[x + 4 + e] = last[4 - DL + e];   // x = the last vector's coordinate
        } else {
#pragma unroll
            for (int e = DL; e < 4; ++e) row[x + e] = last[e - DL];
        }
    }
}

// the 8 output rows row0 + 4*J of a radix-8 butterfly, row J with displacement (MUL * J) & 3
template <int MUL, bool kUp, int... J>
ADRT_HD void oop_store_all(float *xo, int row0, int x, const float (&out)[8][4], const float (&prev)[8][4], bool at_start,
                           std::integer_sequence<int, J...>)
{
    (oop_store_vec<(MUL * J) & 3, kUp>(xo + (row0 + 4 * J) * P, x, out[J], prev[J], at_start), ...);
}
template <int MUL, bool kUp, int... J>
ADRT_HD void oop_store_rest_all(float *xo, int row0, int x, const float (&last)[8][4], std::integer_sequence<int, J...>)
{
    (oop_store_rest<(MUL * J) & 3, kUp>(xo + (row0 + 4 * J) * P, x, last[J]), ...);
}

// ===========================================================================
// transposed: sinogram (d, column) -> workspace rows, 5 stages
// ===========================================================================
// parents q = 0..3 of butterfly p at the offsets x .. x+3: staging[x + i][4p + q]
ADRT_HD void stg_load_parents4(const float *stg, int p, int x, float (&par)[4][4])
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const F4 v = stile::lds128(stg + (x + i) * SW + 4 * p);
#pragma unroll
        for (int q = 0; q < 4; ++q) par[q][i] = v.v[q];
    }
}

// Step A (radix 4) of one butterfly for the segment [c0, c0 + 4*NIT), walked downwards: parents from
// the staging buffer, children into the work tile (conventions of stile::bwd_store_children)
template <bool kMask, int NIT>
ADRT_HD void bwd_stepA_staged(const float *stg, float *buf, int p, int c0, bool warm, stile::BwdStepState<2> &st,
                              int thr_top, int thr_leaf)
{
    constexpr int R = 4;
    st.b.clear();
    const int top = c0 + 4 * NIT;
    if (warm) {
        float par[R][4], out[R][4];
#pragma unroll
        for (int w = 1; w >= 0; --w) {
            stg_load_parents4(stg, p, top + 4 * w, par);
            st.b.template iterate<false>(par, out, 0, 0, 0);
        }
    }
    stg_load_parents4(stg, p, top - 4, st.nxt);
#pragma unroll kUnroll
    for (int it = NIT - 1; it >= 0; --it) {
        float cur[R][4], out[R][4];
#pragma unroll
        for (int q = 0; q < R; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[q][i] = st.nxt[q][i];
        if (it > 0) stg_load_parents4(stg, p, c0 + 4 * (it - 1), st.nxt);
        st.b.template iterate<kMask>(cur, out, c0 + 4 * it, thr_top, thr_leaf);
        stile::bwd_store_children<R, false>(buf, 4 * p, 1, p, c0 + 4 * it, out);
    }
}

// Step B (radix 8) of block k0 for the segment [c0, c0 + 4*NIT), walked downwards, out of place:
// parents from the work tile `buf` (row k0 + 4q, column c at c + floor4(q*k0), as step A left them),
// leaf jj into row k0 + 4jj of `xo` with column c at c + ((-AG*jj) & 3), AG = a_g & 3: the alignment
// of the leaf's workspace row, whose position of column c is d0 - a_g*(8*k0 + jj) + c.
template <int AG, bool kMask, int NIT>
ADRT_HD void bwd_stepB_oop(const float *buf, float *xo, int k0, int c0, stile::BwdStepState<3> &st, int thr_top, int thr_leaf)
{
    constexpr int R = 8;
    st.b.clear();
    const int top = c0 + 4 * NIT;
    {
        float par[R][4], out[R][4];
#pragma unroll
        for (int w = 1; w >= 0; --w) {
            stile::bwd_load_parents<R, true>(buf, k0, 4, k0, top + 4 * w, par);
            st.b.template iterate<false>(par, out, 0, 0, 0);
        }
    }
    stile::bwd_load_parents<R, true>(buf, k0, 4, k0, top - 4, st.nxt);
    float prev[R][4];
#pragma unroll
    for (int jj = 0; jj < R; ++jj)
#pragma unroll
        for (int i = 0; i < 4; ++i) prev[jj][i] = 0.0f;
    auto body = [&](int it, bool at_start) {
        float cur[R][4], out[R][4];
#pragma unroll
        for (int q = 0; q < R; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[q][i] = st.nxt[q][i];
        if (it > 0) stile::bwd_load_parents<R, true>(buf, k0, 4, k0, c0 + 4 * (it - 1), st.nxt);
        st.b.template iterate<kMask>(cur, out, c0 + 4 * it, thr_top, thr_leaf);
        // (-AG * jj) & 3 = ((4 - AG) * jj) & 3
        oop_store_all<(4 - AG) & 3, false>(xo, k0, c0 + 4 * it, out, prev, at_start, std::make_integer_sequence<int, R>());
#pragma unroll
        for (int jj = 0; jj < R; ++jj)
#pragma unroll
            for (int i = 0; i < 4; ++i) prev[jj][i] = out[jj][i];
    };
    body(NIT - 1, true);
#pragma unroll 4
    for (int it = NIT - 2; it >= 0; --it) body(it, false);
    oop_store_rest_all<(4 - AG) & 3, false>(xo, k0, c0, prev, std::make_integer_sequence<int, R>());
}

// Phases: 0 issue the copies into `in` | 1 wait for them (and for the bulk stores that still read
// `mid`) | 2 step A: in -> mid | 3 step B: mid -> in | 4 bulk stores from `in`.  Barriers after 1, 2
// and 3; the kernel (stage_adrt.cu) swaps the two buffers after every tile and issues phase 0 of
// the CTA's next tile right after phase 4.
template <int M, bool kMaskTiles>
struct BwdStaged : stile::BwdStream<M, tile::LOAD_QCOLS, tile::STORE_WROWS, kMaskTiles> {
    static_assert(M == 5, "staged passes fuse 5 stages (two 6-stage buffers leave one CTA per SM)");
    typedef stile::BwdStream<M, tile::LOAD_QCOLS, tile::STORE_WROWS, kMaskTiles> Base;
    typedef SGeo<M, false> Geo;
    static constexpr int G = Geo::G;
    static constexpr bool kRev = true;
    static constexpr bool kImage = false;
    static constexpr int kPhases = 5;
    static constexpr int NT = 64;
    static constexpr int NWARP = NT / 32;
    static constexpr int MIN_CTAS = 3;
    static constexpr int TD = Base::TD;
    typedef typename Base::State State;

    template <int AG>
    ADRT_HD static void step_b(const float *mid, float *in, State &st, const TileCtx &c, int tid)
    {
        int base, k0, c0;
        if (!stile::bwd_sb_map<M, kRev>(tid, base, k0, c0)) return;
        const int thr_top = c.D - c.d0 + c.a_g * (Geo::R1 * k0);
        bwd_stepB_oop<AG, kMaskTiles, Geo::SEG2 / stile::V>(mid, in, k0, c0, st.sb, thr_top, c.a_g);
    }

    template <int PH>
    ADRT_HD static void phase_ct(int mode, float *in, float *mid, State &st, const TmaMap &tm, const float *src, float *dst,
                                 const TileCtx &c, int plane, int tid)
    {
        (void)mode; (void)src;
        const int dt = c.D - c.d0;
        if constexpr (PH == 0) {
            // rows d0 .. d0 + XW of columns g*G .. g*G + 31; rows from D on arrive as +0.0
            stile::bulk_fence();
            if (tid == 0) {
                tma_load_box(st.bar, in, &tm, c.g * G, c.d0, plane);
                tma_load_box(st.bar, in + BOX_ROWS * SW, &tm, c.g * G, c.d0 + BOX_ROWS, plane);
            }
            bulk_arrive(st.bar, tid == 0 ? XW * SW * 4 : 0);
        } else if constexpr (PH == 1) {
            bulk_wait(st.bar);
            bulk_store_wait_read();
        } else if constexpr (PH == 2) {
            int p, seg;
            first_step_map(tid, p, seg);
            // per step leaf (= a block of 8 tile leaves) a node ends a_g*8 + p later (see BwdStream)
            bwd_stepA_staged<kMaskTiles, Geo::SEG1 / stile::V>(in, mid, p, seg * Geo::SEG1, seg < 7, st.sa, dt, c.a_g * Geo::R1 + p);
        } else if constexpr (PH == 3) {
            switch (c.a_g & 3) {
            case 0: step_b<0>(mid, in, st, c, tid); break;
            case 1: step_b<1>(mid, in, st, c, tid); break;
            case 2: step_b<2>(mid, in, st, c, tid); break;
            default: step_b<3>(mid, in, st, c, tid); break;
            }
            stile::bulk_fence();   // this thread's writes of `in` before the bulk copies that read it
        } else {
            // leaf j: workspace row (k0*G + j)*e + a_g, whose position of tile column xc is d0 - a_g*j + xc;
            // the tile owns the aligned chunks at dbase + Q + xa, xa in [0, TD), which hold the columns from
            // Q + xa on = cells xa + (Q ? 4 : 0) on of the leaf's row of `in`
            if (tid < G) {
                const int j = tid;
                float *row = dst + ((long long)(c.k0 * G + j) * c.e + c.a_g) * c.out_pitch;
                const int Q = (c.a_g * j) & 3, gbase = c.d0 - c.a_g * j + Q;   // multiple of 4
                const float *b = in + stile::RowMap<M, kRev, false>::leaf_row(j) * P + (Q ? 4 : 0);
                int lo = gbase < 0 ? -gbase : 0;
                int hi = (c.D - gbase) & ~3;          // whole chunks below D
                if (hi > TD) hi = TD;
                if (hi > lo) bulk_store(row + gbase + lo, b + lo, (hi - lo) * 4);
                bulk_store_commit();
                if (hi >= lo && hi < TD) {
                    // the chunk that straddles D (positions from D on are not the tile's to write)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gbase + hi + i < c.D) row[gbase + hi + i] = b[hi + i];
                }
            }
        }
    }
};

// ===========================================================================
// forward: image -> workspace rows, 5 stages
// ===========================================================================
// Staging layouts.  Quadrants 1 / 2 (tile row j = image column g*G + j, offsets run along the image
// rows): [s][j] boxes, s = x (q2) or XW-1-x (q1: I[r][d] = x[n-1-d, r]).  Quadrants 0 / 3 (tile row
// j = image row, offsets run backwards along it): 32 rows of pitch P, cell XW-1-x of row j.
template <bool kCols>
ADRT_HD void stg_load_leaves4(const float *stg, int k0, int c, bool flip, float (&leaf)[4][4])
{
    if (kCols) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int s = flip ? XW - 1 - (c + i) : c + i;
            const F4 v = stile::lds128(stg + s * SW + 4 * k0);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) leaf[jj][i] = v.v[jj];
        }
    } else {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const F4 v = stile::lds128(stg + (4 * k0 + jj) * P + (XW - 4 - c));
#pragma unroll
            for (int i = 0; i < 4; ++i) leaf[jj][i] = v.v[3 - i];
        }
    }
}

// Step 1 (radix 4, no shifts between the leaves' frames) of butterfly k0 for the segment
// [c0, c0 + 4*NIT): leaves from the staging buffer, outputs into the work tile rows 4*k0 + q
template <bool kCols, int NIT>
ADRT_HD void fwd_step1_staged(const float *stg, float *buf, int k0, int c0, bool warm, bool flip, stile::FwdStepState<2> &st)
{
    constexpr int R = 4;
    st.b.clear();
    if (warm) {
        float leaf[R][4], out[R][4];
#pragma unroll
        for (int w = 2; w >= 1; --w) {
            stg_load_leaves4<kCols>(stg, k0, c0 - 4 * w, flip, leaf);
            st.b.iterate(leaf, out);
        }
    }
    stg_load_leaves4<kCols>(stg, k0, c0, flip, st.nxt);
#pragma unroll kUnroll
    for (int it = 0; it < NIT; ++it) {
        float cur[R][4], out[R][4];
#pragma unroll
        for (int jj = 0; jj < R; ++jj)
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[jj][i] = st.nxt[jj][i];
        if (it + 1 < NIT) stg_load_leaves4<kCols>(stg, k0, c0 + 4 * (it + 1), flip, st.nxt);
        st.b.iterate(cur, out);
#pragma unroll
        for (int q = 0; q < R; ++q) {
            F4 v;
#pragma unroll
            for (int i = 0; i < 4; ++i) v.v[i] = out[q][i];
            *reinterpret_cast<F4 *>(buf + (4 * k0 + q) * P + c0 + 4 * it) = v;
        }
    }
}

// Step 2 (radix 8) of the butterfly with base angle p for the segment [c0, c0 + 4*NIT), out of
// place: leaves from the work tile `buf` (row p + 4jj, read p*jj lower), output q (angle 8p + q) into
// row p + 4q of `xo` with tile coordinate x at x + ((q*KK) & 3), KK = kk & 3: the storage skew
// fwd_row_skew(angle, kk) of the output's workspace row, see stile::fwd_store_wrows.
template <int KK, int NIT>
ADRT_HD void fwd_step2_oop(const float *buf, float *xo, int p, int c0, stile::FwdStepState<3> &st)
{
    constexpr int R = 8;
    st.b.clear();
    {
        float leaf[R][4], out[R][4];
#pragma unroll
        for (int w = 2; w >= 1; --w) {
            stile::fwd_load_leaves<R, true>(buf, p, 4, p, c0 - 4 * w, leaf);
            st.b.iterate(leaf, out);
        }
    }
    stile::fwd_load_leaves<R, true>(buf, p, 4, p, c0, st.nxt);
    float prev[R][4];
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) prev[q][i] = 0.0f;
    auto body = [&](int it, bool at_start) {
        float cur[R][4], out[R][4];
#pragma unroll
        for (int jj = 0; jj < R; ++jj)
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[jj][i] = st.nxt[jj][i];
        if (it + 1 < NIT) stile::fwd_load_leaves<R, true>(buf, p, 4, p, c0 + 4 * (it + 1), st.nxt);
        st.b.iterate(cur, out);
        oop_store_all<KK, true>(xo, p, c0 + 4 * it, out, prev, at_start, std::make_integer_sequence<int, R>());
#pragma unroll
        for (int q = 0; q < R; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) prev[q][i] = out[q][i];
    };
    body(0, true);
#pragma unroll 4
    for (int it = 1; it < NIT; ++it) body(it, false);
    oop_store_rest_all<KK, true>(xo, p, c0 + 4 * (NIT - 1), prev, std::make_integer_sequence<int, R>());
}

// Phases as for BwdStaged (1 also patches the cells below offset 0, which must hold the -0.0 "copy"
// sentinels where the tensor copy delivered +0.0): 2 = step 1, 3 = step 2, 4 = bulk stores.
template <int M>
struct FwdStaged : stile::FwdStream<M, tile::LOAD_IMAGE, tile::STORE_WROWS> {
    static_assert(M == 5, "staged passes fuse 5 stages");
    typedef stile::FwdStream<M, tile::LOAD_IMAGE, tile::STORE_WROWS> Base;
    typedef SGeo<M, true> Geo;
    static constexpr int G = Geo::G;
    static constexpr bool kRev = false;
    static constexpr bool kImage = true;
    static constexpr int kPhases = 5;
    static constexpr int NT = 64;
    static constexpr int NWARP = NT / 32;
    static constexpr int MIN_CTAS = 3;
    static constexpr int TD = Base::TD, LH = Base::LH;
    typedef typename Base::State State;

    template <int KK>
    ADRT_HD static void step_2(const float *mid, float *in, State &st, int tid)
    {
        int base, p, c0;
        if (!stile::fwd_s2_map<M, kRev>(tid, base, p, c0)) return;
        fwd_step2_oop<KK, Geo::SEG2 / stile::V>(mid, in, p, c0, st.s2);
    }

    // `plane` = image index of the tile; c.q its quadrant
    template <int PH>
    ADRT_HD static void phase_ct(int mode, float *in, float *mid, State &st, const TmaMap &tm, const float *src, float *dst,
                                 const TileCtx &c, int plane, int tid)
    {
        (void)mode;
        const int n = c.n;
        const int dbase = c.d0 - LH;          // offset of tile position 0 (multiple of 4)
        const bool cols = c.q == 1 || c.q == 2;
        if constexpr (PH == 0) {
            stile::bulk_fence();
            int my_bytes = 0;
            if (cols) {
                if (tid == 0) {
                    const int r0 = c.q == 2 ? dbase : n - XW - dbase;   // image row of staging row 0
                    tma_load_box(st.bar, in, &tm, c.g * G, r0, plane);
                    tma_load_box(st.bar, in + BOX_ROWS * SW, &tm, c.g * G, r0 + BOX_ROWS, plane);
                    my_bytes = XW * SW * 4;
                }
            } else if (tid < G) {
                // image columns [cs, cs + XW) of the row, clipped: columns < 0 are offsets >= n (+0.0),
                // columns >= n offsets < 0 (-0.0)
                const int r = c.g * G + tid;
                const float *row = src + (long long)plane * n * n + (long long)(c.q == 0 ? r : n - 1 - r) * n;
                const int cs = n - XW - dbase;
                float *d = in + tid * P;
                int lo = cs < 0 ? -cs : 0, hi = n - cs;
                if (lo > XW) lo = XW;
                if (hi > XW) hi = XW;
                if (hi < lo) hi = lo;
                if (lo > 0) stile::bulk_load(st.bar, d, stile::fill_src(false), lo * 4);
                if (hi > lo) stile::bulk_load(st.bar, d + lo, row + cs + lo, (hi - lo) * 4);
                if (hi < XW) stile::bulk_load(st.bar, d + hi, stile::fill_src(true), (XW - hi) * 4);
                my_bytes = XW * 4;
            }
            bulk_arrive(st.bar, my_bytes);
        } else if constexpr (PH == 1) {
            bulk_wait(st.bar);
            bulk_store_wait_read();
            if (cols && dbase < 0) {
                // offsets x < -dbase (a multiple of 4) are below 0: whole staging rows
                const int nx = -dbase;
                for (int k = tid; k < nx * (SW / 4); k += NT) {
                    const int x = k / (SW / 4), s = c.q == 2 ? x : XW - 1 - x;
                    F4 v;
#pragma unroll
                    for (int i = 0; i < 4; ++i) v.v[i] = -0.0f;
                    *reinterpret_cast<F4 *>(in + s * SW + 4 * (k % (SW / 4))) = v;
                }
            }
        } else if constexpr (PH == 2) {
            int k0, seg;
            first_step_map(tid, k0, seg);
            if (cols) fwd_step1_staged<true, Geo::SEG1 / stile::V>(in, mid, k0, seg * Geo::SEG1, seg > 0, c.q == 1, st.s1);
            else fwd_step1_staged<false, Geo::SEG1 / stile::V>(in, mid, k0, seg * Geo::SEG1, seg > 0, false, st.s1);
        } else if constexpr (PH == 3) {
            const int kk = c.next_g ? (c.k0 & (c.next_g - 1)) : 0;
            switch (kk & 3) {
            case 0: step_2<0>(mid, in, st, tid); break;
            case 1: step_2<1>(mid, in, st, tid); break;
            case 2: step_2<2>(mid, in, st, tid); break;
            default: step_2<3>(mid, in, st, tid); break;
            }
            stile::bulk_fence();   // this thread's writes of `in` before the bulk copies that read it
        } else {
            // output angle A: workspace row g*G + A, whose chunk at position d0 + xc holds the offsets from
            // d0 + xc - S on (S = the row's storage skew) = tile coordinates LH + xc - S on = cells LH + xc on
            // of the angle's row of `in`; everything above a row's support is exact +0.0 already
            if (tid < G) {
                const int A = tid;
                long long len = c.out_pitch - c.d0;
                if (len > TD) len = TD;
                if (len > 0)
                    bulk_store(dst + ((long long)c.g * G + A) * c.out_pitch + c.d0,
                               in + stile::RowMap<M, kRev, true>::out_row(A) * P + LH, (int)len * 4);
                bulk_store_commit();
            }
        }
    }
};

template <typename Prog, int PH = 0>
ADRT_HD void run_phase(int ph, int mode, float *in, float *mid, typename Prog::State &st, const TmaMap &tm, const float *src,
                       float *dst, const TileCtx &c, int plane, int tid)
{
    if constexpr (PH < Prog::kPhases) {
        if (ph == PH) Prog::template phase_ct<PH>(mode, in, mid, st, tm, src, dst, c, plane, tid);
        else run_phase<Prog, PH + 1>(ph, mode, in, mid, st, tm, src, dst, c, plane, tid);
    }
}

}  // namespace sgtile
}  // namespace adrt_b200
