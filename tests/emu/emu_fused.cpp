// Host emulator of the fused CUDA passes (TEST INFRASTRUCTURE).
//
// Compiles adrt_b200/csrc/fused_tile.h + fused_plan.h as plain C++ and runs
// every CTA of every pass sequentially: for each phase, for tid = 0..NT-1.
// The arithmetic, index algebra, tiling, masking and workspace layouts are the
// very code the GPU executes, so tests/test_emu_fused.py can compare the fused
// algorithm bit-for-bit with the oracle without a GPU.  It is NOT a product
// path: nothing in adrt_b200/ links or loads it.
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <type_traits>
#include <vector>

#include "../../adrt_b200/csrc/fused_plan.h"
#include "../../adrt_b200/csrc/iadrt_tile.h"
#include "../../adrt_b200/csrc/stage_tile.h"

using namespace adrt_b200;

namespace {

// groups (blockIdx.y range) the next pass runs: angle-block sharding runs a rank's share only
int g_y_off = 0, g_y_cnt = -1;

// subtract-on-load (BwdProgram<..., kSub>): byte distance from the sinogram to the one subtracted, 0 = off
long long g_sub_delta = 0;

template <typename T, int M, int LOADK, int STOREK, bool kForward, bool kSub = false>
void run_pass(const plan::Pass &p, const T *src, T *dst, int n, int D, int planes, long long sps, long long dps)
{
    using Prog = typename std::conditional<kForward, tile::FwdProgram<T, M, LOADK, STOREK>,
                                           tile::BwdProgram<T, M, LOADK, STOREK, kSub>>::type;
    // Pack<T> accesses need 16-byte alignment: allocate as Pack vectors
    const size_t cells = (size_t)tile::Geo<M>::G * tile::Pitch<T>::value;
    std::vector<tile::Pack<T>> storeA(cells / tile::VecOf<T>::L + 1);
    T *buf = reinterpret_cast<T *>(storeA.data());
    // per-thread "registers" that live across the barriers of a step
    std::vector<T> regfile((size_t)tile::Geo<M>::NT * tile::NREG);
    const int e = 1 << p.s;
    const int y_lo = g_y_cnt < 0 ? 0 : g_y_off, y_hi = g_y_cnt < 0 ? p.grid_y : g_y_off + g_y_cnt;
    for (int plane = 0; plane < planes; ++plane)
        for (int by = y_lo; by < y_hi; ++by)
            for (int bx = 0; bx < p.grid_x; ++bx) {
                tile::TileCtx c;
                c.n = n; c.D = D; c.e = e; c.g = by; c.k0 = by / e; c.a_g = by % e;
                c.d0 = bx * Prog::TD;
                c.next_g = p.next_g;
                c.d_need = p.d_need;
                c.sup_loge = p.sup_loge; c.sup_gmask = p.sup_gmask;
                c.in_pitch = p.in_pitch; c.out_pitch = p.out_pitch;
                c.q = 0;
                c.sub_delta = kSub ? g_sub_delta : 0;
                const int mode = Prog::classify(c);
                if (mode == tile::TILE_SKIP || (mode == tile::TILE_ZERO && p.skip_zero)) continue;
                const T *sp;
                if (LOADK == tile::LOAD_IMAGE) { c.q = plane & 3; sp = src + (long long)(plane >> 2) * sps; }
                else sp = src + (long long)plane * sps;
                T *dp = dst + (long long)plane * dps;
                // poison shared memory so that reads of never-written cells are visible
                for (size_t i = 0; i < cells; ++i) buf[i] = T(1e30);
                for (auto &v : regfile) v = T(-1e30);
                if (mode == tile::TILE_ZERO) {
                    for (int tid = 0; tid < tile::Geo<M>::NT; ++tid) Prog::zero_tile(buf, dp, c, tid);
                    continue;
                }
                for (int ph = 0; ph < Prog::kPhases; ++ph)
                    for (int tid = 0; tid < tile::Geo<M>::NT; ++tid)
                        tile::run_phase<Prog, T>(ph, mode, buf, *reinterpret_cast<T(*)[tile::NREG]>(&regfile[(size_t)tid * tile::NREG]), sp, dp, c, tid);
            }
}

// thread order inside a phase: 0 = ascending tid, 1 = descending (a result that depends on
// the order means two threads race on a tile cell within one phase)
int g_order = 0;
long long g_stream_tiles = 0;   // full tiles run by the streaming programs

// Streaming passes (stream_tile.h): fp32, M = 5 or 6.
template <typename Prog>
void run_stream_pass(const plan::Pass &p, const float *src, float *dst, int n, int D, int planes, long long sps, long long dps,
                     bool image_loader)
{
    const size_t cells = (size_t)Prog::G * stile::P + 64;   // + slack: the top segment of a transposed step reads a few cells past the last row
    std::vector<tile::Pack<float>> storeA(cells / 4 + 1);
    float *buf = reinterpret_cast<float *>(storeA.data());
    std::vector<typename Prog::State> states(Prog::NT);
    const int e = 1 << p.s;
    const int y_lo = g_y_cnt < 0 ? 0 : g_y_off, y_hi = g_y_cnt < 0 ? p.grid_y : g_y_off + g_y_cnt;
    for (int plane = 0; plane < planes; ++plane)
        for (int by = y_lo; by < y_hi; ++by)
            for (int bx = 0; bx < p.grid_x; ++bx) {
                tile::TileCtx c;
                c.n = n; c.D = D; c.e = e; c.g = by; c.k0 = by / e; c.a_g = by % e;
                c.d0 = bx * Prog::TD;
                c.next_g = p.next_g;
                c.d_need = p.d_need;
                c.sup_loge = p.sup_loge; c.sup_gmask = p.sup_gmask;
                c.in_pitch = p.in_pitch; c.out_pitch = p.out_pitch;
                c.q = 0;
                const int mode = Prog::classify(c);
                if (mode == tile::TILE_SKIP || (mode == tile::TILE_ZERO && p.skip_zero)) continue;
                if (!Prog::runs(mode)) continue;
                const float *sp;
                if (image_loader) { c.q = plane & 3; sp = src + (long long)(plane >> 2) * sps; }
                else sp = src + (long long)plane * sps;
                float *dp = dst + (long long)plane * dps;
                for (size_t i = 0; i < cells; ++i) buf[i] = 1e30f;
                if (mode == tile::TILE_ZERO) {
                    for (int tid = 0; tid < Prog::NT; ++tid) Prog::zero_tile(buf, dp, c, tid);
                    continue;
                }
                ++g_stream_tiles;
                unsigned long long fake_bar = 0;
                for (int tid = 0; tid < Prog::NT; ++tid) stile::bulk_init(states[tid].bar, &fake_bar, Prog::NT, tid);
                for (int ph = 0; ph < Prog::kPhases; ++ph)
                    for (int i = 0; i < Prog::NT; ++i) {
                        const int tid = g_order ? Prog::NT - 1 - i : i;
                        stile::run_phase<Prog>(ph, mode, buf, states[tid], sp, dp, c, tid);
                    }
            }
}

// Staged passes (stage_tile.h): the tensor-map / bulk copies are played by the host branch of
// sgtile::tma_load_box / stile::bulk_load / sgtile::bulk_store; every tile runs its five phases in order
// (the GPU kernel swaps the two buffers from tile to tile and issues phase 0 of a CTA's next tile early
// -- an ordering the barriers protect, not the arithmetic).
long long g_staged_tiles = 0;
template <typename Prog>
void run_staged_pass(const plan::Pass &p, const float *src, float *dst, int n, int D, int planes, long long sps, long long dps,
                     bool image_loader)
{
    const size_t cells = 2 * (size_t)sgtile::STG_FLOATS + 64;
    std::vector<tile::Pack<float>> storeA(cells / 4 + 1);
    float *stg = reinterpret_cast<float *>(storeA.data());
    float *buf = stg + sgtile::STG_FLOATS;
    std::vector<typename Prog::State> states(Prog::NT);
    const int e = 1 << p.s;
    sgtile::TmaMap tm;
    tm.base = src;
    tm.dim0 = n;
    tm.dim1 = image_loader ? n : D;
    tm.dim2 = image_loader ? (planes + 3) / 4 : planes;
    tm.stride1 = n;
    tm.stride2 = sps;
    const int y_lo = g_y_cnt < 0 ? 0 : g_y_off, y_hi = g_y_cnt < 0 ? p.grid_y : g_y_off + g_y_cnt;
    for (int plane = 0; plane < planes; ++plane)
        for (int by = y_lo; by < y_hi; ++by)
            for (int bx = 0; bx < p.grid_x; ++bx) {
                tile::TileCtx c;
                c.n = n; c.D = D; c.e = e; c.g = by; c.k0 = by / e; c.a_g = by % e;
                c.d0 = bx * Prog::TD;
                c.next_g = p.next_g;
                c.d_need = p.d_need;
                c.sup_loge = p.sup_loge; c.sup_gmask = p.sup_gmask;
                c.in_pitch = p.in_pitch; c.out_pitch = p.out_pitch;
                c.q = 0;
                const int mode = Prog::classify(c);
                if (mode == tile::TILE_SKIP || (mode == tile::TILE_ZERO && p.skip_zero)) continue;
                if (!Prog::runs(mode)) continue;
                int lplane = plane;
                if (image_loader) { c.q = plane & 3; lplane = plane >> 2; }
                float *dp = dst + (long long)plane * dps;
                for (size_t i = 0; i < cells; ++i) stg[i] = 1e30f;
                if (mode == tile::TILE_ZERO) {
                    for (int tid = 0; tid < Prog::NT; ++tid) Prog::zero_tile(buf, dp, c, tid);
                    continue;
                }
                ++g_staged_tiles;
                unsigned long long fake_bar = 0;
                for (int tid = 0; tid < Prog::NT; ++tid) stile::bulk_init(states[tid].bar, &fake_bar, Prog::NT, tid);
                for (int ph = 0; ph < Prog::kPhases; ++ph)
                    for (int i = 0; i < Prog::NT; ++i) {
                        const int tid = g_order ? Prog::NT - 1 - i : i;
                        sgtile::run_phase<Prog>(ph, mode, stg, buf, states[tid], tm, src, dp, c, lplane, tid);
                    }
            }
}

template <int M, bool kForward>
void run_stream_kinds(const plan::Pass &p, const float *src, float *dst, int n, int D, int planes, long long sps, long long dps)
{
    using namespace tile;
    if (kForward) {
        if (p.load == LOAD_IMAGE && p.store == STORE_WROWS) run_stream_pass<stile::FwdStream<M, LOAD_IMAGE, STORE_WROWS>>(p, src, dst, n, D, planes, sps, dps, true);
        else if (p.load == LOAD_IMAGE && p.store == STORE_QCOLS) run_stream_pass<stile::FwdStream<M, LOAD_IMAGE, STORE_QCOLS>>(p, src, dst, n, D, planes, sps, dps, true);
        else if (p.load == LOAD_WROWS && p.store == STORE_WROWS) run_stream_pass<stile::FwdStream<M, LOAD_WROWS, STORE_WROWS>>(p, src, dst, n, D, planes, sps, dps, false);
        else if (M == 6 && plan::stream_split2()) run_stream_pass<stile::FwdStream<6, LOAD_WROWS, STORE_QCOLS, 2>>(p, src, dst, n, D, planes, sps, dps, false);
        else run_stream_pass<stile::FwdStream<M, LOAD_WROWS, STORE_QCOLS>>(p, src, dst, n, D, planes, sps, dps, false);
    } else {
        if (p.load == LOAD_QCOLS && p.store == STORE_WROWS) { run_stream_pass<stile::BwdStream<M, LOAD_QCOLS, STORE_WROWS, false>>(p, src, dst, n, D, planes, sps, dps, false); run_stream_pass<stile::BwdStream<M, LOAD_QCOLS, STORE_WROWS, true>>(p, src, dst, n, D, planes, sps, dps, false); }
        else if (p.load == LOAD_QCOLS && p.store == STORE_QCOLS) { run_stream_pass<stile::BwdStream<M, LOAD_QCOLS, STORE_QCOLS, false>>(p, src, dst, n, D, planes, sps, dps, false); run_stream_pass<stile::BwdStream<M, LOAD_QCOLS, STORE_QCOLS, true>>(p, src, dst, n, D, planes, sps, dps, false); }
        else if (p.load == LOAD_WROWS && p.store == STORE_WROWS) { run_stream_pass<stile::BwdStream<M, LOAD_WROWS, STORE_WROWS, false>>(p, src, dst, n, D, planes, sps, dps, false); run_stream_pass<stile::BwdStream<M, LOAD_WROWS, STORE_WROWS, true>>(p, src, dst, n, D, planes, sps, dps, false); }
        else { run_stream_pass<stile::BwdStream<M, LOAD_WROWS, STORE_QCOLS, false>>(p, src, dst, n, D, planes, sps, dps, false); run_stream_pass<stile::BwdStream<M, LOAD_WROWS, STORE_QCOLS, true>>(p, src, dst, n, D, planes, sps, dps, false); }
    }
}

template <typename T, bool kForward>
void run_stream(const plan::Pass &p, const T *src, T *dst, int n, int D, int planes, long long sps, long long dps)
{
    if constexpr (std::is_same<T, float>::value) {
        if (p.staged) {
            if (kForward) run_staged_pass<sgtile::FwdStaged<5>>(p, src, dst, n, D, planes, sps, dps, true);
            else {
                run_staged_pass<sgtile::BwdStaged<5, false>>(p, src, dst, n, D, planes, sps, dps, false);
                run_staged_pass<sgtile::BwdStaged<5, true>>(p, src, dst, n, D, planes, sps, dps, false);
            }
            return;
        }
        if (p.M == 6) run_stream_kinds<6, kForward>(p, src, dst, n, D, planes, sps, dps);
        else run_stream_kinds<5, kForward>(p, src, dst, n, D, planes, sps, dps);
    }
}

template <typename T, int M, bool kForward>
void run_kinds(const plan::Pass &p, const T *src, T *dst, int n, int D, int planes, long long sps, long long dps)
{
    using namespace tile;
    if (kForward) {
        if (p.load == LOAD_IMAGE && p.store == STORE_WROWS) run_pass<T, M, LOAD_IMAGE, STORE_WROWS, kForward>(p, src, dst, n, D, planes, sps, dps);
        else if (p.load == LOAD_IMAGE && p.store == STORE_QCOLS) run_pass<T, M, LOAD_IMAGE, STORE_QCOLS, kForward>(p, src, dst, n, D, planes, sps, dps);
        else if (p.load == LOAD_WROWS && p.store == STORE_WROWS) run_pass<T, M, LOAD_WROWS, STORE_WROWS, kForward>(p, src, dst, n, D, planes, sps, dps);
        else run_pass<T, M, LOAD_WROWS, STORE_QCOLS, kForward>(p, src, dst, n, D, planes, sps, dps);
    } else {
        if constexpr (M >= 4) {
            if (g_sub_delta != 0 && p.src_buf < 0 && p.load == LOAD_QCOLS && p.store == STORE_WROWS) {
                run_pass<T, M, LOAD_QCOLS, STORE_WROWS, kForward, true>(p, src, dst, n, D, planes, sps, dps);
                return;
            }
        }
        if (p.load == LOAD_QCOLS && p.store == STORE_WROWS) run_pass<T, M, LOAD_QCOLS, STORE_WROWS, kForward>(p, src, dst, n, D, planes, sps, dps);
        else if (p.load == LOAD_QCOLS && p.store == STORE_QCOLS) run_pass<T, M, LOAD_QCOLS, STORE_QCOLS, kForward>(p, src, dst, n, D, planes, sps, dps);
        else if (p.load == LOAD_WROWS && p.store == STORE_WROWS) run_pass<T, M, LOAD_WROWS, STORE_WROWS, kForward>(p, src, dst, n, D, planes, sps, dps);
        else run_pass<T, M, LOAD_WROWS, STORE_QCOLS, kForward>(p, src, dst, n, D, planes, sps, dps);
    }
}

// rows_layout: forward stores / transposed loads the caller's array as R-layout rows (fused normal operator)
template <typename T, bool kForward>
int run_one_pass(const plan::Pass &p, const T *src, T *dst, int n, int D, int planes, long long sps, long long dps)
{
    if (p.stream) {
        run_stream<T, kForward>(p, src, dst, n, D, planes, sps, dps);
        return 0;
    }
    switch (p.M) {
    case 1: run_kinds<T, 1, kForward>(p, src, dst, n, D, planes, sps, dps); break;
    case 2: run_kinds<T, 2, kForward>(p, src, dst, n, D, planes, sps, dps); break;
    case 3: run_kinds<T, 3, kForward>(p, src, dst, n, D, planes, sps, dps); break;
    case 4: run_kinds<T, 4, kForward>(p, src, dst, n, D, planes, sps, dps); break;
    case 5: run_kinds<T, 5, kForward>(p, src, dst, n, D, planes, sps, dps); break;
    case 6: run_kinds<T, 6, kForward>(p, src, dst, n, D, planes, sps, dps); break;
    default: return 2;
    }
    return 0;
}

template <typename T, bool kForward>
int run(const T *in, T *out, int64_t B, int64_t n64, int64_t rows = -1, bool rows_layout = false)
{
    plan::Plan pl;
    const bool ok = kForward ? plan::make_forward_plan(n64, sizeof(T), &pl, rows_layout)
                             : plan::make_transposed_plan(n64, sizeof(T), &pl, rows, rows_layout);
    if (!ok) return 1;
    const int n = pl.n, D = pl.D, planes = (int)B * 4;
    // workspaces start as NaN so that any read of a never-written element shows up
    std::vector<T> ws0(pl.ws_slot_elems[0] * planes, T(NAN)), ws1(pl.ws_slot_elems[1] * planes, T(NAN));
    T *slot[2] = {ws0.data(), ws1.data()};
    const long long img = (long long)n * n, sino = (long long)D * n;
    for (int i = 0; i < pl.npass; ++i) {
        const plan::Pass &p = pl.pass[i];
        const T *src;
        T *dst;
        long long sps, dps;
        if (p.src_buf < 0) { src = in; sps = kForward ? img : (p.in_pitch ? (long long)n * p.in_pitch : sino); }
        else { src = slot[p.src_buf]; sps = (long long)n * p.in_pitch; }
        if (p.dst_buf < 0) { dst = out; dps = p.out_pitch ? (long long)n * p.out_pitch : sino; }
        else { dst = slot[p.dst_buf]; dps = (long long)n * p.out_pitch; }
        if (run_one_pass<T, kForward>(p, src, dst, n, D, planes, sps, dps)) return 2;
    }
    return 0;
}

// ---- angle-block sharding (fused_plan.h part_*): `parts` ranks emulated one after the other ----------
// Forward: every rank runs the head passes on its own blocks into its own copy of the exchange buffer
// (NaN elsewhere), the exchange copies exactly the rows the protocol sends, the tail pass runs on the
// rank's angle range and fills the rank's columns of `out`.  Transposed: the mirror image.
template <typename T, bool kForward>
int run_parts(const T *in, T *out, int64_t B, int64_t n64, int parts, int m_last, int64_t rows)
{
    plan::Plan pl;
    const int K = plan::ilog2(n64);
    const std::vector<int> ms = plan::part_split(K, sizeof(T), m_last);
    if (ms.empty() || (1 << m_last) < parts) return 1;
    const bool ok = kForward ? plan::make_forward_plan_split(n64, sizeof(T), ms, &pl)
                             : plan::make_transposed_plan_split(n64, sizeof(T), ms, &pl, rows);
    if (!ok) return 1;
    const int n = pl.n, D = pl.D, planes = (int)B * 4, np = pl.npass;
    const long long img = (long long)n * n, sino = (long long)D * n;
    // the exchange buffer is the workspace between the head passes and the tail pass (forward) /
    // between the first pass and the rest (transposed)
    const int xi = kForward ? np - 2 : 0;                       // pass that WRITES the exchange buffer
    const long long xpitch = pl.pass[xi].out_pitch;
    const int e = 1 << (K - m_last), nblk = 1 << m_last;        // rows (blk, a): blk < nblk, a < e
    const size_t xelems = (size_t)planes * n * xpitch;
    std::vector<std::vector<T>> xbuf(parts, std::vector<T>(xelems, T(NAN)));
    std::vector<T> ws0(pl.ws_slot_elems[0] * planes), ws1(pl.ws_slot_elems[1] * planes);
    auto run_range = [&](int first, int last, int part, const T *src_in, T *dst_out) -> int {
        // passes [first, last]; reads src_in at `first`, writes dst_out at `last`; slots in between
        for (int i = first; i <= last; ++i) {
            const plan::Pass &p = pl.pass[i];
            const bool angle_pass = kForward ? (i == np - 1) : (i == 0);
            const plan::YRange yr = angle_pass ? plan::part_angle_range(p, part, parts) : plan::part_block_range(p, n, part, parts);
            g_y_off = yr.y_off; g_y_cnt = yr.y_cnt;
            T *slot[2] = {ws0.data(), ws1.data()};
            const T *src = i == first ? src_in : slot[p.src_buf];
            T *dst = i == last ? dst_out : slot[p.dst_buf];
            const long long sps = p.src_buf < 0 ? (kForward ? img : sino) : (long long)n * p.in_pitch;
            const long long dps = p.dst_buf < 0 ? sino : (long long)n * p.out_pitch;
            const int rc = run_one_pass<T, kForward>(p, src, dst, n, D, planes, sps, dps);
            g_y_cnt = -1;
            if (rc) return rc;
        }
        return 0;
    };
    // phase 1
    for (int part = 0; part < parts; ++part) {
        std::fill(ws0.begin(), ws0.end(), T(NAN));
        std::fill(ws1.begin(), ws1.end(), T(NAN));
        if (run_range(0, xi, part, in, xbuf[part].data())) return 2;
    }
    // exchange: row (blk, a) of every plane goes from the rank that computed it to the rank that needs it
    for (int dstp = 0; dstp < parts; ++dstp)
        for (int srcp = 0; srcp < parts; ++srcp) {
            if (srcp == dstp) continue;
            for (int plane = 0; plane < planes; ++plane)
                for (int blk = 0; blk < nblk; ++blk)
                    for (int a = 0; a < e; ++a) {
                        const int blk_owner = blk / (nblk / parts), ang_owner = a / (e / parts);
                        // forward: computed by the block owner, needed by the angle owner; transposed: the reverse
                        const int from = kForward ? blk_owner : ang_owner, to = kForward ? ang_owner : blk_owner;
                        if (from != srcp || to != dstp) continue;
                        const size_t off = ((size_t)plane * n + (size_t)blk * e + a) * xpitch;
                        std::copy(xbuf[srcp].begin() + off, xbuf[srcp].begin() + off + xpitch, xbuf[dstp].begin() + off);
                    }
        }
    // phase 2
    for (int part = 0; part < parts; ++part) {
        std::fill(ws0.begin(), ws0.end(), T(NAN));
        std::fill(ws1.begin(), ws1.end(), T(NAN));
        if (run_range(xi + 1, np - 1, part, xbuf[part].data(), out)) return 3;
    }
    return 0;
}

// bdrt(a - b) with the subtraction done by the loader of the first pass (iadrt_fmg_step's residual);
// returns 4 when the plan of this size has no such first pass (the engine then subtracts separately)
template <typename T>
int run_bdrt_sub(const T *a, const T *b, T *out, int64_t B, int64_t n, int64_t rows)
{
    plan::Plan pl;
    if (!plan::make_transposed_plan(n, sizeof(T), &pl, rows, false)) return 1;
    const plan::Pass &p = pl.pass[0];
    if (!(pl.npass >= 2 && !p.stream && !p.staged && p.src_buf < 0 && p.in_pitch == 0 && p.load == tile::LOAD_QCOLS &&
          p.store == tile::STORE_WROWS && p.M >= 4 && p.M <= 6))
        return 4;
    g_sub_delta = (long long)(reinterpret_cast<const char *>(b) - reinterpret_cast<const char *>(a));
    const int rc = run<T, false>(a, out, B, n, rows, false);
    g_sub_delta = 0;
    return rc;
}

// fused normal operator data flow: adrt with R-layout rows out (`mid`: B * 4 * n * round4(2n-1) elements),
// then bdrt (offsets d < rows) with R-layout rows in
template <typename T>
int run_normal(const T *in, T *mid, T *out, int64_t B, int64_t n, int64_t rows)
{
    int rc = run<T, true>(in, mid, B, n, -1, true);
    if (rc) return rc;
    return run<T, false>(mid, out, B, n, rows, true);
}
// ---- fused iadrt passes (iadrt_tile.h): the 32 lanes of every warp played phase by phase ------------
// one base row X0 - U for the 32 lanes of a warp, in the configured lane order; the shuffle that brings the
// partner lane's h2 reads all lanes' registers before any lane computes
template <typename T, int M, bool kOutQ, int U>
void iadrt_row(bool interior, T *ring, const itile::LaneConst<M> *lc, const itile::TripAddr<M> *ta, int n, int X0,
               itile::LaneState<T, M> *st, T **op)
{
    T ph2[32];
    for (int lane = 0; lane < 32; ++lane) ph2[lane] = st[lane ^ 1].h2;
    for (int i = 0; i < 32; ++i) {
        const int lane = g_order ? 31 - i : i;
        if (interior) itile::all_levels_interior<T, M, kOutQ, U>(ring, lc[lane], ta[lane], n, X0, st[lane], op[lane], ph2[lane]);
        else itile::all_levels<T, M, kOutQ, U>(ring, lc[lane], ta[lane], n, X0, st[lane], op[lane], ph2[lane]);
    }
}

template <typename T, int M, bool kInQ, bool kOutQ>
void run_iadrt_pass(const T *in, T *out, int64_t planes, int n, int s0)
{
    using G = itile::Geo<M>;
    const int teams = n >> M, warps = (teams + G::TEAMS - 1) / G::TEAMS;
    const int D = 2 * n - 1;
    const long long in_plane = kInQ ? (long long)D * n : (long long)n * 2 * n;
    const long long out_plane = kOutQ ? (long long)D * n : (long long)n * 2 * n;
    std::vector<T> ring((size_t)G::ROWS * itile::kLanes);
    for (int64_t plane = 0; plane < planes; ++plane)
        for (int w = 0; w < warps; ++w) {
            const int tp0 = w * G::TEAMS;
            const itile::Team t0 = itile::make_team<M>(n, s0, tp0);
            if (!t0.active) continue;
            const int top = itile::sweep_top(D, t0.c0 * (G::G - 1));
            std::fill(ring.begin(), ring.end(), T(NAN));   // reads of never-written cells show up
            itile::Team tm[32];
            itile::LaneState<T, M> st[32];
            itile::LaneConst<M> lc[32];
            const T *ip[32];
            T *op[32];
            for (int lane = 0; lane < 32; ++lane) {
                const int k = lane % G::G;
                tm[lane] = itile::make_team<M>(n, s0, tp0 + lane / G::G);
                itile::setup_levels<M>(tm[lane], (lane / G::G) * G::G, k, lane, lc[lane]);
                ip[lane] = in + plane * in_plane + (kInQ ? tm[lane].in_col + k : (tm[lane].in_col + k) * (long long)(2 * n));
                op[lane] = out + plane * out_plane + (kOutQ ? tm[lane].out_col + k : (tm[lane].out_col + (long long)k * tm[lane].out_stride) * (long long)(2 * n));
                for (int t = 0; t <= M; ++t) st[lane].prev[t] = T(0);
                st[lane].h1 = st[lane].h2 = T(NAN);
                for (int e = 0; e < 8; ++e) st[lane].v[e] = T(NAN);
                itile::fetch_inputs<T, kInQ>(ip[lane], tm[lane], top, st[lane].v);
            }
            int ilo = -0x40000000, ihi = 0x40000000;   // the warp-wide reduction of the lanes' interior ranges
            for (int lane = 0; lane < 32; ++lane) {
                int lo, hi;
                itile::interior_range<M>(lc[lane], tm[lane].active, lo, hi);
                ilo = std::max(ilo, lo);
                ihi = std::min(ihi, hi);
            }
            for (int X0 = top; X0 >= -M; X0 -= 8) {
                for (int i = 0; i < 32; ++i) {
                    const int lane = g_order ? 31 - i : i;
                    itile::commit_inputs<T, M>(ring.data(), tm[lane], lane, X0, st[lane].v);
                    itile::fetch_inputs<T, kInQ>(ip[lane], tm[lane], X0 - 8, st[lane].v);
                }
                for (int h = 0; h < 2; ++h) {
                    const int X4 = X0 - 4 * h;
                    itile::TripAddr<M> ta[32];
                    for (int lane = 0; lane < 32; ++lane) itile::trip_setup<M, kOutQ>(lc[lane], X4, ta[lane]);
                    const bool interior = X4 - 3 >= ilo && X4 <= ihi;
                    iadrt_row<T, M, kOutQ, 0>(interior, ring.data(), lc, ta, n, X4, st, op);
                    iadrt_row<T, M, kOutQ, 1>(interior, ring.data(), lc, ta, n, X4, st, op);
                    iadrt_row<T, M, kOutQ, 2>(interior, ring.data(), lc, ta, n, X4, st, op);
                    iadrt_row<T, M, kOutQ, 3>(interior, ring.data(), lc, ta, n, X4, st, op);
                }
                if (!kOutQ)
                    for (int lane = 0; lane < 32; ++lane)
                        itile::flush_outputs<T, M>(ring.data(), tm[lane], tm[lane].c0 * (lane % G::G), lane, X0, op[lane]);
            }
        }
}

template <typename T, bool kInQ, bool kOutQ>
void run_iadrt_m(int M, const T *in, T *out, int64_t planes, int n, int s0)
{
    switch (M) {
    case 1: run_iadrt_pass<T, 1, kInQ, kOutQ>(in, out, planes, n, s0); break;
    case 2: run_iadrt_pass<T, 2, kInQ, kOutQ>(in, out, planes, n, s0); break;
    case 3: run_iadrt_pass<T, 3, kInQ, kOutQ>(in, out, planes, n, s0); break;
    case 4: run_iadrt_pass<T, 4, kInQ, kOutQ>(in, out, planes, n, s0); break;
    default: run_iadrt_pass<T, 5, kInQ, kOutQ>(in, out, planes, n, s0); break;
    }
}

template <typename T>
int run_iadrt(const T *in, T *out, int64_t B, int64_t n64)
{
    const int n = (int)n64, K = plan::ilog2(n64);
    if (K < 1) return 1;
    int ms[8];
    const int np = itile::iadrt_split(K, ms, (int)sizeof(T));
    const int64_t planes = B * 4;
    const size_t w = (size_t)planes * n * 2 * n;
    std::vector<T> w0(np > 1 ? w : 0, T(NAN)), w1(np > 2 ? w : 0, T(NAN));
    T *wbuf[2] = {w0.data(), w1.data()};
    const T *src = in;
    int s0 = 0;
    for (int i = 0; i < np; ++i) {
        const bool first = i == 0, last = i == np - 1;
        T *dst = last ? out : wbuf[i & 1];
        if (first && last) run_iadrt_m<T, true, true>(ms[i], src, dst, planes, n, s0);
        else if (first) run_iadrt_m<T, true, false>(ms[i], src, dst, planes, n, s0);
        else if (last) run_iadrt_m<T, false, true>(ms[i], src, dst, planes, n, s0);
        else run_iadrt_m<T, false, false>(ms[i], src, dst, planes, n, s0);
        src = dst;
        s0 += ms[i];
    }
    return 0;
}

}  // namespace

extern "C" {
int emu_iadrt_f32(const float *in, float *out, int64_t B, int64_t n) { return run_iadrt<float>(in, out, B, n); }
int emu_iadrt_f64(const double *in, double *out, int64_t B, int64_t n) { return run_iadrt<double>(in, out, B, n); }
void emu_set_order(int order) { g_order = order; }
int emu_bdrt_sub_f32(const float *a, const float *b, float *out, int64_t B, int64_t n, int64_t rows) { return run_bdrt_sub<float>(a, b, out, B, n, rows); }
int emu_bdrt_sub_f64(const double *a, const double *b, double *out, int64_t B, int64_t n, int64_t rows) { return run_bdrt_sub<double>(a, b, out, B, n, rows); }
long long emu_stream_tiles(void) { return g_stream_tiles; }
long long emu_staged_tiles(void) { return g_staged_tiles; }
int emu_adrt_f32(const float *in, float *out, int64_t B, int64_t n) { return run<float, true>(in, out, B, n); }
int emu_adrt_f64(const double *in, double *out, int64_t B, int64_t n) { return run<double, true>(in, out, B, n); }
int emu_bdrt_f32(const float *in, float *out, int64_t B, int64_t n) { return run<float, false>(in, out, B, n); }
int emu_bdrt_f64(const double *in, double *out, int64_t B, int64_t n) { return run<double, false>(in, out, B, n); }
// only offsets d < rows of every output plane are produced
int emu_bdrt_rows_f32(const float *in, float *out, int64_t B, int64_t n, int64_t rows) { return run<float, false>(in, out, B, n, rows); }
int emu_bdrt_rows_f64(const double *in, double *out, int64_t B, int64_t n, int64_t rows) { return run<double, false>(in, out, B, n, rows); }
// angle-block sharding over `parts` emulated ranks, last pass of m_last stages (rows < 0: all offsets)
int emu_adrt_parts_f32(const float *in, float *out, int64_t B, int64_t n, int parts, int m_last) { return run_parts<float, true>(in, out, B, n, parts, m_last, -1); }
int emu_adrt_parts_f64(const double *in, double *out, int64_t B, int64_t n, int parts, int m_last) { return run_parts<double, true>(in, out, B, n, parts, m_last, -1); }
int emu_bdrt_parts_f32(const float *in, float *out, int64_t B, int64_t n, int parts, int m_last, int64_t rows) { return run_parts<float, false>(in, out, B, n, parts, m_last, rows); }
int emu_bdrt_parts_f64(const double *in, double *out, int64_t B, int64_t n, int parts, int m_last, int64_t rows) { return run_parts<double, false>(in, out, B, n, parts, m_last, rows); }
int emu_normal_f32(const float *in, float *mid, float *out, int64_t B, int64_t n, int64_t rows) { return run_normal<float>(in, mid, out, B, n, rows); }
int emu_normal_f64(const double *in, double *mid, double *out, int64_t B, int64_t n, int64_t rows) { return run_normal<double>(in, mid, out, B, n, rows); }
}
