// adrt.iadrt as fused multi-stage passes (iadrt_tile.h): one warp sweeps a group of up to 32 columns
// through up to 5 stages at once, so the sinogram crosses HBM once per pass (3 passes for n = 2048)
// instead of once per stage (11).  Reference: adrt_cdefs_iadrt.hpp:52-176.
#include "common.cuh"
#include "iadrt_tile.h"

namespace adrt_b200 {

namespace {

// Warps per block: ADRT_B200_IADRT_WARPS = 1 / 2 / 4 / 8 (A/B knob; the warps are independent, the block
// size only sets the granularity at which the SMs fill up)
constexpr int kDefaultWarps = 4;
// Launch bound: threads only.  Asking for 5 / 6 / 7 / 8 resident blocks (102 / 80 / 72 / 64 registers)
// measured 8.0 / 8.9 / 7.8 / 8.8 ms against 7.5 ms at 16 x 2048^2 fp32: the sweep is issue bound, and
// the spills of the tighter bounds cost more than the extra warps give.

// A/B knobs (tools/build_variant.sh): resident blocks the register allocation aims at for the five- / four-level
// passes (0: threads only)
#ifndef ADRT_IADRT_MINB5
#define ADRT_IADRT_MINB5 0
#endif
#ifndef ADRT_IADRT_MINB4
#define ADRT_IADRT_MINB4 0
#endif
template <int M, int W> constexpr int iadrt_min_blocks()
{
    return W != 4 ? 0 : (M == 5 ? ADRT_IADRT_MINB5 : (M == 4 ? ADRT_IADRT_MINB4 : 0));
}

template <typename T, int M, bool kInQ, bool kOutQ, int kWarpsPerBlock>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, iadrt_min_blocks<M, kWarpsPerBlock>())
iadrt_pass_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t planes, int n, int s0)
{
    using G = itile::Geo<M>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *ring = reinterpret_cast<T *>(smem_raw) + (size_t)warp * G::ROWS * itile::kLanes;
    const int team = lane / G::G, k = lane % G::G, team_lane0 = team * G::G;
    const int tp0 = (blockIdx.x * kWarpsPerBlock + warp) * G::TEAMS;
    const itile::Team tm = itile::make_team<M>(n, s0, tp0 + team);
    // the warp's first team has the largest c0, i.e. the longest sweep
    const itile::Team t0 = itile::make_team<M>(n, s0, tp0);
    if (!t0.active) return;
    const int top = itile::sweep_top(tm.D, t0.c0 * (G::G - 1));
    const long long in_plane = kInQ ? (long long)tm.D * n : (long long)n * 2 * n;
    const long long out_plane = kOutQ ? (long long)tm.D * n : (long long)n * 2 * n;
    itile::LaneConst<M> lc;
    itile::setup_levels<M>(tm, team_lane0, k, lane, lc);
    const int psi_out = tm.c0 * k;
    // base rows for which no lane of the warp needs a guard
    int ilo, ihi;
    itile::interior_range<M>(lc, tm.active, ilo, ihi);
    ilo = __reduce_max_sync(0xffffffffu, ilo);
    ihi = __reduce_min_sync(0xffffffffu, ihi);
    for (int64_t plane = blockIdx.y; plane < planes; plane += gridDim.y) {
        // the lane's input column and output column / workspace row
        const T *ip = in + plane * in_plane + (kInQ ? tm.in_col + k : (tm.in_col + k) * (long long)(2 * n));
        T *op = out + plane * out_plane + (kOutQ ? tm.out_col + k : (tm.out_col + (long long)k * tm.out_stride) * (long long)(2 * n));
        itile::LaneState<T, M> st;
#pragma unroll
        for (int t = 0; t <= M; ++t) st.prev[t] = T(0);
        st.h1 = st.h2 = T(0);
        // the other lane of the pair's h2 (last level from registers, iadrt_tile.h Geo::kRegLast)
        auto partner = [](T h2) -> T {
            if constexpr (G::kRegLast) return __shfl_xor_sync(0xffffffffu, h2, 1);
            else return h2;
        };
        itile::fetch_inputs<T, kInQ>(ip, tm, top, st.v);
        for (int X0 = top; X0 >= -M; X0 -= 8) {
            itile::commit_inputs<T, M>(ring, tm, lane, X0, st.v);
            itile::fetch_inputs<T, kInQ>(ip, tm, X0 - 8, st.v);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int X4 = X0 - 4 * h;      // the two trips of the pair
                itile::TripAddr<M> ta;
                itile::trip_setup<M, kOutQ>(lc, X4, ta);
                __syncwarp();
                if (X4 - 3 >= ilo && X4 <= ihi) {
                    itile::all_levels_interior<T, M, kOutQ, 0>(ring, lc, ta, n, X4, st, op, partner(st.h2));
                    __syncwarp();
                    itile::all_levels_interior<T, M, kOutQ, 1>(ring, lc, ta, n, X4, st, op, partner(st.h2));
                    __syncwarp();
                    itile::all_levels_interior<T, M, kOutQ, 2>(ring, lc, ta, n, X4, st, op, partner(st.h2));
                    __syncwarp();
                    itile::all_levels_interior<T, M, kOutQ, 3>(ring, lc, ta, n, X4, st, op, partner(st.h2));
                    __syncwarp();
                } else {
                    itile::all_levels<T, M, kOutQ, 0>(ring, lc, ta, n, X4, st, op, partner(st.h2));
                    __syncwarp();
                    itile::all_levels<T, M, kOutQ, 1>(ring, lc, ta, n, X4, st, op, partner(st.h2));
                    __syncwarp();
                    itile::all_levels<T, M, kOutQ, 2>(ring, lc, ta, n, X4, st, op, partner(st.h2));
                    __syncwarp();
                    itile::all_levels<T, M, kOutQ, 3>(ring, lc, ta, n, X4, st, op, partner(st.h2));
                    __syncwarp();
                }
            }
            if (!kOutQ) itile::flush_outputs<T, M>(ring, tm, psi_out, lane, X0, op);
        }
        __syncwarp();
    }
}

template <typename T, int M, bool kInQ, bool kOutQ, int kWarpsPerBlock>
int launch_iadrt_pass_w(const T *in, T *out, int64_t planes, int n, int s0, cudaStream_t s)
{
    using G = itile::Geo<M>;
    const int teams = n >> M;                                       // per plane
    const int warps = (teams + G::TEAMS - 1) / G::TEAMS;
    const int blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const size_t smem = (size_t)kWarpsPerBlock * G::ROWS * itile::kLanes * sizeof(T);
    auto kern = iadrt_pass_kernel<T, M, kInQ, kOutQ, kWarpsPerBlock>;
    ADRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)blocks, (unsigned)(planes < 65535 ? planes : 65535));
    kern<<<grid, kWarpsPerBlock * 32, smem, s>>>(in, out, planes, n, s0);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T, int M, bool kInQ, bool kOutQ>
int launch_iadrt_pass(const T *in, T *out, int64_t planes, int n, int s0, cudaStream_t s)
{
    int w = kDefaultWarps;
    if (const char *e = getenv("ADRT_B200_IADRT_WARPS")) w = atoi(e);
    switch (w) {
    case 1: return launch_iadrt_pass_w<T, M, kInQ, kOutQ, 1>(in, out, planes, n, s0, s);
    case 2: return launch_iadrt_pass_w<T, M, kInQ, kOutQ, 2>(in, out, planes, n, s0, s);
    case 8: return launch_iadrt_pass_w<T, M, kInQ, kOutQ, 8>(in, out, planes, n, s0, s);
    default: return launch_iadrt_pass_w<T, M, kInQ, kOutQ, 4>(in, out, planes, n, s0, s);
    }
}

template <typename T, bool kInQ, bool kOutQ>
int dispatch_m(int M, const T *in, T *out, int64_t planes, int n, int s0, cudaStream_t s)
{
    switch (M) {
    case 1: return launch_iadrt_pass<T, 1, kInQ, kOutQ>(in, out, planes, n, s0, s);
    case 2: return launch_iadrt_pass<T, 2, kInQ, kOutQ>(in, out, planes, n, s0, s);
    case 3: return launch_iadrt_pass<T, 3, kInQ, kOutQ>(in, out, planes, n, s0, s);
    case 4: return launch_iadrt_pass<T, 4, kInQ, kOutQ>(in, out, planes, n, s0, s);
    case 5: return launch_iadrt_pass<T, 5, kInQ, kOutQ>(in, out, planes, n, s0, s);
    }
    set_error("internal: bad iadrt stages per pass %d", M);
    return ADRT_B200_EINVAL;
}

}  // namespace

// workspace: one column-major buffer (planes x n x 2n) per intermediate, at most two
template <typename T>
size_t fused_iadrt_workspace_elems(int64_t B, int64_t n)
{
    const int K = num_iters(n);
    if (K < 1 || n > kMaxN) return 0;
    int ms[8];
    const int np = itile::iadrt_split(K, ms, (int)sizeof(T));
    const size_t w = (size_t)(B * 4) * (size_t)n * (size_t)(2 * n);
    return np <= 1 ? 0 : (np == 2 ? w : 2 * w);
}

template <typename T>
int fused_iadrt(const T *in, T *out, int64_t B, int64_t n, T *ws, size_t ws_elems, cudaStream_t s)
{
    const int K = num_iters(n);
    int ms[8];
    const int np = itile::iadrt_split(K, ms, (int)sizeof(T));
    const size_t need = fused_iadrt_workspace_elems<T>(B, n);
    if (need > 0 && (!ws || ws_elems < need)) {
        set_error("iadrt workspace too small: need %zu elements, got %zu", need, ws_elems);
        return ADRT_B200_EWORKSPACE;
    }
    const int64_t planes = B * 4;
    const size_t w = (size_t)planes * (size_t)n * (size_t)(2 * n);
    T *wbuf[2] = {ws, ws + w};
    const T *src = in;
    int s0 = 0, rc = ADRT_B200_OK;
    for (int i = 0; i < np && rc == ADRT_B200_OK; ++i) {
        const bool first = i == 0, last = i == np - 1;
        T *dst = last ? out : wbuf[i & 1];
        if (first && last) rc = dispatch_m<T, true, true>(ms[i], src, dst, planes, (int)n, s0, s);
        else if (first) rc = dispatch_m<T, true, false>(ms[i], src, dst, planes, (int)n, s0, s);
        else if (last) rc = dispatch_m<T, false, true>(ms[i], src, dst, planes, (int)n, s0, s);
        else rc = dispatch_m<T, false, false>(ms[i], src, dst, planes, (int)n, s0, s);
        src = dst;
        s0 += ms[i];
    }
    return rc;
}

template size_t fused_iadrt_workspace_elems<float>(int64_t, int64_t);
template size_t fused_iadrt_workspace_elems<double>(int64_t, int64_t);
template int fused_iadrt<float>(const float *, float *, int64_t, int64_t, float *, size_t, cudaStream_t);
template int fused_iadrt<double>(const double *, double *, int64_t, int64_t, double *, size_t, cudaStream_t);

}  // namespace adrt_b200
