// Host-side planning of the fused passes: how the K = log2(n) butterfly stages
// are split into passes, the R-layout workspace each pass writes and the grid
// each pass needs.  Shared by the CUDA driver (fused_adrt.cu) and the host
// emulator used by the CPU tests (tests/emu/emu_fused.cpp).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "fused_tile.h"
#include "stream_tile.h"

namespace adrt_b200 {
namespace plan {

constexpr int kMaxPasses = 4;

struct Pass {
    int M;            // stages fused in this pass
    int s;            // stages done before it (block height e = 2^s)
    int load;         // tile::LoadKind
    int store;        // tile::StoreKind
    long long in_pitch, out_pitch;  // row length (elements) of the R-layout workspaces
    int src_buf, dst_buf;           // -1 = caller's input / output, else workspace slot 0/1
    int grid_x, grid_y;             // d-tiles, groups
    int next_g;                     // forward: group size of the pass reading this pass's workspace (0: none)
    int d_need;                     // transposed: output offsets >= d_need are not needed (tiles skipped)
    bool stream;                    // run by the streaming kernels (stream_tile.h) instead of fused_tile.h
    bool staged;                    // ... by their staged variants (stage_tile.h: TMA-fed, persistent); implies stream
    // transposed plans: this pass does not write its all-zero tiles (d0 >= D: only structural zeros of the
    // sheared output rows) because the next pass synthesises them (TileCtx::sup_*); and the geometry of
    // the pass that wrote this pass's input rows when that one skipped (sup_gmask < 0: it did not)
    bool skip_zero;
    int sup_loge, sup_gmask;
};

struct Plan {
    int n, K, D;
    int npass;
    Pass pass[kMaxPasses];
    size_t ws_slot_elems[2];  // per plane
};

inline int ilog2(int64_t n)
{
    int k = 0;
    while ((int64_t(1) << k) < n) ++k;
    return k;
}

// Largest M whose (single, in-place) tile buffer fits in shared memory (227 KB).
inline int max_stages_per_pass(size_t elem_size)
{
    int m = 1;
    const size_t pitch = elem_size == 8 ? tile::Pitch<double>::value : tile::Pitch<float>::value;
    while (m < 6 && (2ull << m) * pitch * elem_size <= 227ull * 1024) ++m;
    return m;
}

// Split K stages into passes.  ADRT_B200_SPLIT="6,5" overrides (testing/tuning).
inline std::vector<int> split_stages(int K, size_t elem_size, const char *env_name, bool small_first = false)
{
    std::vector<int> out;
    if (const char *e = getenv(env_name)) {
        int sum = 0;
        const int cap = max_stages_per_pass(elem_size);
        bool ok = true;
        for (const char *p = e; *p;) {
            int v = (int)strtol(p, const_cast<char **>(&p), 10);
            if (v < 1 || v > cap) ok = false;
            out.push_back(v);
            sum += v;
            if (*p == ',') ++p;
        }
        if (ok && sum == K && (int)out.size() <= kMaxPasses) return out;
        out.clear();
    }
    const int cap = max_stages_per_pass(elem_size);
    const int np = (K + cap - 1) / cap;
    int left = K;
    for (int i = 0; i < np; ++i) {
        const int m = (left + (np - i) - 1) / (np - i);  // as even as possible, larger first
        out.push_back(m);
        left -= m;
    }
    if (small_first) std::reverse(out.begin(), out.end());
    return out;
}

inline long long round4(long long v) { return (v + 3) & ~3LL; }

// Row length of the forward workspace after s stages: the support n + 2^s
// (capped at D), rounded up so that rows stay 16-byte aligned.
// (+3: rows are stored with a skew of up to 3 elements, see tile::fwd_row_skew)
inline long long fwd_pitch(int n, int s)
{
    const long long D = 2LL * n - 1, p = (long long)n + (1LL << s);
    return round4((p < D ? p : D) + 3);
}

inline int tile_td(int M, int store, bool stream = false)
{
    if (stream) return M == 6 ? stile::STileTD<6, 0>::value + (store == tile::STORE_WROWS ? 0 : 4)
                              : stile::STileTD<5, 0>::value + (store == tile::STORE_WROWS ? 0 : 4);
    const int G = 1 << M;
    return tile::XW - (G < 4 ? 4 : G) - (store == tile::STORE_WROWS ? 4 : 0);
}

// Which passes the streaming kernels (stream_tile.h) take: fp32, 5 or 6 stages, and only the pass
// kinds where they beat the fused_tile.h kernels (profiles/): a kind is named <f|b><M><p|w> --
// forward / transposed, stages, loading from the public side (image / sinogram) or from workspace
// rows.  ADRT_B200_STREAM_SET="f6p,f5w,..." overrides the default set ("" = none, "all" = every kind).
inline bool use_stream(int M, size_t elem_size, bool forward, int load)
{
    if (elem_size != 4 || (M != 5 && M != 6)) return false;
    char key[4] = {forward ? 'f' : 'b', (char)('0' + M), load == tile::LOAD_WROWS ? 'w' : 'p', 0};
    const char *set = getenv("ADRT_B200_STREAM_SET");
    // b5p / b6p / f5p: the fused_tile.h kernels win -- their loader runs the first radix-4 step on the
    // public-layout columns it fetches (16 x 4096^2 bdrt: 8.9 ms with a streaming b6p, 7.3 ms without)
    // f6p: 16 x 4096^2 adrt 5.53 ms streaming, 5.31 ms with fused_tile.h's fused image loader
    if (!set) set = "f6w,f5w,b6w,b5w";
    if (!strcmp(set, "all")) return true;
    return strstr(set, key) != nullptr;
}

// ADRT_B200_STREAM_SPLIT2=1: the six-stage forward streaming pass that stores the public layout runs two
// threads per (butterfly, segment), 128 per tile (stile::FwdStream kSplit = 2)
inline bool stream_split2()
{
    const char *e = getenv("ADRT_B200_STREAM_SPLIT2");
    return e && atoi(e) != 0;
}

// Which passes the staged kernels (stage_tile.h) take: the fp32 five-stage passes that read the public
// layout and store workspace rows -- "f5p" (images) and "b5p" (sinograms).  ADRT_B200_STAGE_SET overrides
// the default set ("" = none).  Default "f5p": 64 x 2048^2 adrt 4.75 -> 4.50 ms, 64 x 1024^2 1.27 -> 1.19 ms;
// "b5p" is bit-identical too but slower than the fused_tile.h kernel with its fused loader (bdrt 6.37 ->
// 9.1 ms: its masked boundary tiles cost 7x an interior tile), profiles/s5_*.
inline bool use_staged(int M, size_t elem_size, bool forward, int load, int store)
{
    if (elem_size != 4 || M != 5 || store != tile::STORE_WROWS) return false;
    if (load != (forward ? tile::LOAD_IMAGE : tile::LOAD_QCOLS)) return false;
    const char *set = getenv("ADRT_B200_STAGE_SET");
    if (!set) set = "f5p";
    return strstr(set, forward ? "f5p" : "b5p") != nullptr;
}

// `rows_out`: the last pass stores its result as R-layout rows (row = angle, pitch round4(D), logical
// positions, zeros above each row's support) into the caller's buffer instead of the public (d, column)
// layout -- the format the first pass of a transposed plan built with `rows_in` loads.  The fused
// normal operator hands adrt's result to bdrt this way: neither the public-layout
// store nor the public-layout load happens.
inline bool make_forward_plan_split(int64_t n64, size_t elem_size, const std::vector<int> &ms, Plan *pl, bool rows_out = false);

inline bool make_forward_plan(int64_t n64, size_t elem_size, Plan *pl, bool rows_out = false)
{
    const int K = ilog2(n64);
    if (K < 1) return false;
    // fp32: the pass next to the public layout is the long one in both directions (it is a streaming
    // pass, stream_tile.h, and the fastest kernel per stage) -- measured 4.7 vs 5.05 ms at 64 x 2048^2
    std::vector<int> ms = split_stages(K, elem_size, "ADRT_B200_SPLIT", elem_size == 4);
    // 256^2 fp32: (3, 5) ends with a streaming five-stage pass -- 4096 images: 12.5 ms vs 16.3 ms for (4, 4)
    if (elem_size == 4 && K == 8 && !getenv("ADRT_B200_SPLIT")) ms = {3, 5};
    return make_forward_plan_split(n64, elem_size, ms, pl, rows_out);
}

// Same with the stages-per-pass given explicitly (they must add up to log2 n).
inline bool make_forward_plan_split(int64_t n64, size_t elem_size, const std::vector<int> &ms, Plan *pl, bool rows_out)
{
    const int n = (int)n64;
    const int K = ilog2(n64);
    if (K < 1 || ms.empty() || (int)ms.size() > kMaxPasses) return false;
    {
        int sum = 0;
        for (int m : ms) {
            if (m < 1 || m > max_stages_per_pass(elem_size)) return false;
            sum += m;
        }
        if (sum != K) return false;
    }
    pl->n = n; pl->K = K; pl->D = 2 * n - 1;
    pl->npass = (int)ms.size();
    pl->ws_slot_elems[0] = pl->ws_slot_elems[1] = 0;
    int s = 0;
    for (int i = 0; i < pl->npass; ++i) {
        Pass &p = pl->pass[i];
        const bool first = i == 0, last = i == pl->npass - 1;
        p.M = ms[i];
        p.s = s;
        p.load = first ? tile::LOAD_IMAGE : tile::LOAD_WROWS;
        p.store = (last && !rows_out) ? tile::STORE_QCOLS : tile::STORE_WROWS;
        p.in_pitch = first ? 0 : fwd_pitch(n, s);
        p.out_pitch = last ? (rows_out ? round4(pl->D) : 0) : fwd_pitch(n, s + p.M);
        p.src_buf = first ? -1 : (i - 1) & 1;
        p.dst_buf = last ? -1 : i & 1;
        const int G = 1 << p.M;
        p.staged = use_staged(p.M, elem_size, true, p.load, p.store);
        p.stream = p.staged || use_stream(p.M, elem_size, true, p.load);
        const int TD = tile_td(p.M, p.store, p.stream);
        p.next_g = last ? 0 : (1 << ms[i + 1]);
        p.d_need = pl->D;
        p.skip_zero = false; p.sup_loge = 0; p.sup_gmask = -1;
        const long long extent = (last && !rows_out) ? pl->D : p.out_pitch;  // offsets that must be written
        p.grid_x = (int)((extent + TD - 1) / TD);
        p.grid_y = n / G;
        if (!last) {
            const size_t need = (size_t)n * (size_t)p.out_pitch;
            if (need > pl->ws_slot_elems[p.dst_buf]) pl->ws_slot_elems[p.dst_buf] = need;
        }
        s += p.M;
    }
    return true;
}

// Transposed plan: forward pass i is undone by transposed pass npass-1-i.
// `rows` < D asks only for output offsets d < rows of the final result (what
// utils.truncate keeps): every pass then skips the tiles that cannot reach them.
// `rows_in`: the first pass loads R-layout rows (see make_forward_plan) from the caller's buffer.
inline bool make_transposed_plan_split(int64_t n64, size_t elem_size, const std::vector<int> &ms, Plan *pl, int64_t rows = -1,
                                       bool rows_in = false);

inline bool make_transposed_plan(int64_t n64, size_t elem_size, Plan *pl, int64_t rows = -1, bool rows_in = false)
{
    const int K = ilog2(n64);
    if (K < 1) return false;
    return make_transposed_plan_split(n64, elem_size, split_stages(K, elem_size, "ADRT_B200_SPLIT_BDRT"), pl, rows, rows_in);
}

// `ms`: stages per pass in FORWARD order (transposed pass i undoes forward pass npass-1-i).
inline bool make_transposed_plan_split(int64_t n64, size_t elem_size, const std::vector<int> &ms, Plan *pl, int64_t rows,
                                       bool rows_in)
{
    const int n = (int)n64;
    const int K = ilog2(n64);
    if (K < 1 || ms.empty() || (int)ms.size() > kMaxPasses) return false;
    {
        int sum = 0;
        for (int m : ms) {
            if (m < 1 || m > max_stages_per_pass(elem_size)) return false;
            sum += m;
        }
        if (sum != K) return false;
    }
    pl->n = n; pl->K = K; pl->D = 2 * n - 1;
    pl->npass = (int)ms.size();
    pl->ws_slot_elems[0] = pl->ws_slot_elems[1] = 0;
    int s = K;
    for (int i = 0; i < pl->npass; ++i) {
        Pass &p = pl->pass[i];
        const bool first = i == 0, last = i == pl->npass - 1;
        p.M = ms[pl->npass - 1 - i];
        s -= p.M;
        p.s = s;  // block height of the rows this pass PRODUCES is 2^s
        p.load = (first && !rows_in) ? tile::LOAD_QCOLS : tile::LOAD_WROWS;
        p.store = last ? tile::STORE_QCOLS : tile::STORE_WROWS;
        p.in_pitch = (first && !rows_in) ? 0 : round4(pl->D);
        p.out_pitch = last ? 0 : round4(pl->D);
        p.src_buf = first ? -1 : (i - 1) & 1;
        p.dst_buf = last ? -1 : i & 1;
        const int G = 1 << p.M;
        p.staged = use_staged(p.M, elem_size, false, p.load, p.store);
        p.stream = p.staged || use_stream(p.M, elem_size, false, p.load);
        p.next_g = 0;
        p.grid_y = n / G;
        p.skip_zero = false; p.sup_loge = 0; p.sup_gmask = -1;
        if (!last) {
            const size_t need = (size_t)n * (size_t)round4(pl->D);
            if (need > pl->ws_slot_elems[p.dst_buf]) pl->ws_slot_elems[p.dst_buf] = need;
        }
    }
    // A streaming pass reads its workspace rows with per-row bulk copies and fills what a row does not have
    // from a constant source, so its producer need not write the rows' structural-zero tails: the tiles at
    // d0 >= D (12 % of the workspace at 2048^2).  Opt-in (ADRT_B200_SKIP_ZERO=1): bit-identical, but measured
    // 2-3 % SLOWER (64 x 2048^2 bdrt 6.36 -> 6.55 ms, profiles/r02_skipzero.jsonl) -- a clipped row costs a
    // second bulk copy from the constant source, which is worth more than the DRAM bytes it saves.
    {
        const char *env = getenv("ADRT_B200_SKIP_ZERO");
        const bool on = env && atoi(env) != 0;
        for (int i = 1; on && i < pl->npass; ++i)
            if (pl->pass[i].stream) {
                pl->pass[i - 1].skip_zero = true;
                pl->pass[i].sup_loge = pl->pass[i - 1].s;
                pl->pass[i].sup_gmask = (1 << pl->pass[i - 1].M) - 1;
            }
    }
    // Offsets each pass must produce: an output at offset d of a pass with block height e
    // and G rows per group reads inputs at offsets < d + (e-1)*(G-1) + G (SURVEY 8a row a5');
    // +8 covers the 16-byte chunking of the workspace rows.
    long long need = (rows > 0 && rows < pl->D) ? rows : pl->D;
    for (int i = pl->npass - 1; i >= 0; --i) {
        Pass &p = pl->pass[i];
        p.d_need = (int)(need < pl->D ? need : pl->D);
        const long long G = 1LL << p.M, e = 1LL << p.s;
        need = p.d_need + (e - 1) * (G - 1) + G + 8;
        const int TD = tile_td(p.M, p.store, p.stream);
        const long long extent = p.d_need + (e - 1) * (G - 1);  // tile coordinates that hold wanted outputs
        p.grid_x = (int)((extent + TD - 1) / TD);
    }
    return true;
}

// ---- angle-block sharding of one plane over `parts` ranks (single large image) --------
// The last forward pass fuses m_last stages (2^m_last >= parts).  Every earlier pass only combines
// image rows of the same 2^(K - m_last)-row block, so rank `part` runs them on its own blocks alone;
// the last pass only combines rows with the same incoming angle a_g, so after ONE exchange of
// workspace rows (row (blk, a): the owner of block blk sends it to the owner of angle a) rank `part`
// runs it for its own angle range and ends up with the sinogram columns
// [part * n / parts, (part + 1) * n / parts).  bdrt is the mirror image: its first pass (the transpose
// of that last pass) runs on the rank's columns, the exchange goes the other way, the remaining passes
// run on the rank's blocks.
struct YRange {
    int y_off, y_cnt;   // groups (blockIdx.y range) of a pass that belong to a rank
};

// stages-per-pass of the sharded plans: the first K - m_last stages split as evenly as possible, then m_last
inline std::vector<int> part_split(int K, size_t elem_size, int m_last)
{
    std::vector<int> ms;
    const int head = K - m_last;
    if (head < 1 || m_last < 1) return ms;
    const int cap = max_stages_per_pass(elem_size);
    const int np = (head + cap - 1) / cap;
    int left = head;
    for (int i = 0; i < np; ++i) {
        const int m = (left + (np - i) - 1) / (np - i);
        ms.push_back(m);
        left -= m;
    }
    ms.push_back(m_last);
    return ms;
}

// a pass that works on whole blocks (every pass but the one that fuses the last m_last stages):
// groups g = k0 * e + a_g with k0 in the rank's share of the n / (e * G) block groups
inline YRange part_block_range(const Pass &p, int n, int part, int parts)
{
    const int e = 1 << p.s, K0 = n / (e << p.M);
    YRange r;
    r.y_off = (K0 / parts) * part * e;
    r.y_cnt = (K0 / parts) * e;
    return r;
}

// the pass that fuses the last m_last stages (one block group, k0 = 0): the rank's share of the angles a_g
inline YRange part_angle_range(const Pass &p, int part, int parts)
{
    const int e = 1 << p.s;
    YRange r;
    r.y_off = (e / parts) * part;
    r.y_cnt = e / parts;
    return r;
}

}  // namespace plan
}  // namespace adrt_b200
