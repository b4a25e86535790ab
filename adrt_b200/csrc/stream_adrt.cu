// Streaming fused passes for sm_100a (fp32, 5 or 6 stages per pass): the CUDA
// kernels around the tile programs of stream_tile.h.  One CTA = one tile of one
// group of one (image, quadrant) plane, 128 threads; the butterfly steps keep a
// radix-8 butterfly and its history in registers (few, fat threads: the tile, not
// the register file, limits occupancy).
#include "pass_args.h"
#include "sched.cuh"

#include <mutex>

namespace adrt_b200 {

namespace {

// the barrier-separated phases of a full tile, unrolled at compile time
template <typename Prog, int PH>
__device__ __forceinline__ void run_phases(int mode, float *buf, typename Prog::State &st, const float *sp, float *dp,
                                           const tile::TileCtx &c, int tid)
{
    if constexpr (PH < Prog::kPhases) {
        Prog::template phase_ct<PH>(mode, buf, st, sp, dp, c, tid);
        if (Prog::barrier_after(PH)) __syncthreads();
        run_phases<Prog, PH + 1>(mode, buf, st, sp, dp, c, tid);
    }
}

template <typename Prog>
__global__ void __launch_bounds__(Prog::NT, Prog::MIN_CTAS)
stream_kernel(const float *__restrict__ src, float *__restrict__ dst, PassArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *buf = reinterpret_cast<float *>(smem_raw);

    tile::TileCtx c;
    c.n = a.n;
    c.D = a.D;
    c.e = a.e;
    c.g = blockIdx.y + a.y_off;
    c.k0 = c.g >> a.loge;
    c.a_g = c.g & (a.e - 1);
    c.d0 = (blockIdx.x + a.x_off) * Prog::TD;
    c.next_g = a.next_g;
    c.d_need = a.d_need;
    c.sup_loge = a.sup_loge;
    c.sup_gmask = a.sup_gmask;
    c.in_pitch = a.in_pitch;
    c.out_pitch = a.out_pitch;
    c.q = 0;
    const int mode = Prog::classify(c);
    if (!Prog::runs(mode) || (mode == tile::TILE_ZERO && a.skip_zero)) return;
    const int tid = threadIdx.x;
    __shared__ unsigned long long bulk_bar;
    typename Prog::State st;
    if (mode != tile::TILE_ZERO) stile::bulk_init(st.bar, &bulk_bar, Prog::NT, tid);

    for (int plane = blockIdx.z; plane < a.planes; plane += gridDim.z) {
        const float *sp;
        if (Prog::kImage) {
            const int gp = plane + a.plane0;
            c.q = a.q_first + gp % a.q_count;
            sp = src + (long long)(gp / a.q_count) * a.src_plane_stride;
        } else {
            sp = src + (long long)plane * a.src_plane_stride;
        }
        float *dp = dst + (long long)plane * a.dst_plane_stride;
        if (mode == tile::TILE_ZERO) {
            Prog::zero_tile(buf, dp, c, tid);
        } else {
            run_phases<Prog, 0>(mode, buf, st, sp, dp, c, tid);
        }
    }
}

template <typename Prog>
int launch(const float *src, float *dst, PassArgs a, int x_first, int grid_x, int grid_y, cudaStream_t s)
{
    // d-tiles x_first .. x_first + grid_x - 1
    if (grid_x <= 0) return ADRT_B200_OK;
    a.x_off = x_first;
    auto kern = stream_kernel<Prog>;
    // + 32 bytes: the top segment of a transposed step reads a few cells past the last row
    const size_t smem = (size_t)Prog::G * stile::P * sizeof(float) + 32;
    ADRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)grid_x, (unsigned)grid_y, (unsigned)(a.planes < 65535 ? a.planes : 65535));
    kern<<<grid, Prog::NT, smem, s>>>(src, dst, a);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

// Persistent variant (sched.cuh): work items in plane-major order from a global counter, acquire of the
// producing pass's per-plane counter before the first tile of a plane, release after every tile.
template <typename Prog>
__global__ void __launch_bounds__(Prog::NT, Prog::MIN_CTAS)
stream_kernel_p(const float *__restrict__ src, float *__restrict__ dst, PassArgs a, SchedArgs sc)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *buf = reinterpret_cast<float *>(smem_raw);
    __shared__ unsigned long long bulk_bar;
    __shared__ unsigned slot[2];
    const int tid = threadIdx.x;

    tile::TileCtx c;
    c.n = a.n;
    c.D = a.D;
    c.e = a.e;
    c.next_g = a.next_g;
    c.d_need = a.d_need;
    c.sup_loge = a.sup_loge;
    c.sup_gmask = a.sup_gmask;
    c.in_pitch = a.in_pitch;
    c.out_pitch = a.out_pitch;
    c.q = 0;
    typename Prog::State st;
    stile::bulk_init(st.bar, &bulk_bar, Prog::NT, tid);
    SchedIter iter;
    iter.begin(sc, slot, tid);
    int plane, y, x, ready_plane = -1;
    while (iter.current(sc, slot, tid, plane, y, x)) {
        c.g = y + a.y_off;
        c.k0 = c.g >> a.loge;
        c.a_g = c.g & (a.e - 1);
        c.d0 = (x + a.x_off) * Prog::TD;
        const int mode = Prog::classify(c);
        if (Prog::runs(mode) && !(mode == tile::TILE_ZERO && a.skip_zero)) {
            const float *sp;
            if (Prog::kImage) {
                const int gp = plane + a.plane0;
                c.q = a.q_first + gp % a.q_count;
                sp = src + (long long)(gp / a.q_count) * a.src_plane_stride;
            } else {
                sp = src + (long long)plane * a.src_plane_stride;
            }
            float *dp = dst + (long long)plane * a.dst_plane_stride;
            if (mode == tile::TILE_ZERO) {
                Prog::zero_tile(buf, dp, c, tid);
            } else {
                if (sc.dep && plane != ready_plane) {
                    sched_wait_plane(sc, plane, tid);
                    ready_plane = plane;
                }
                run_phases<Prog, 0>(mode, buf, st, sp, dp, c, tid);
            }
        }
        if (sc.done) sched_signal_plane(sc, plane, tid);
        iter.advance();
    }
}

// d-tiles x_first .. x_first + tiles_x - 1 as one persistent launch with its own work counter
template <typename Prog>
int launch_p(const float *src, float *dst, PassArgs a, SchedArgs sc, int x_first, int tiles_x, unsigned *next, int ctas,
             cudaStream_t s)
{
    if (tiles_x <= 0) return ADRT_B200_OK;
    a.x_off = x_first;
    sc.next = next;
    sc.tiles_x = tiles_x;
    sc.total = (unsigned)a.planes * (unsigned)tiles_x * (unsigned)sc.tiles_y;
    sc.ctas = ctas;
    auto kern = stream_kernel_p<Prog>;
    ADRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    static std::atomic<size_t> cached[8] = {};
    const size_t base = (size_t)Prog::G * stile::P * sizeof(float) + 32;
    const int ci = sc.cap_per_sm > 0 && sc.cap_per_sm < 8 ? sc.cap_per_sm : 0;
    size_t smem = cached[ci].load();
    if (!smem) cached[ci].store(smem = capped_smem(kern, Prog::NT, base, ci));
    kern<<<(unsigned)ctas, Prog::NT, smem, s>>>(src, dst, a, sc);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <int M>
int dispatch_fwd_p(const plan::Pass &p, const float *src, float *dst, const PassArgs &a, const SchedArgs &sc, cudaStream_t s)
{
    using namespace tile;
    if (p.load == LOAD_IMAGE && p.store == STORE_WROWS) return launch_p<stile::FwdStream<M, LOAD_IMAGE, STORE_WROWS>>(src, dst, a, sc, 0, p.grid_x, sc.next, sc.ctas, s);
    if (p.load == LOAD_WROWS && p.store == STORE_QCOLS) return launch_p<stile::FwdStream<M, LOAD_WROWS, STORE_QCOLS>>(src, dst, a, sc, 0, p.grid_x, sc.next, sc.ctas, s);
    set_error("internal: bad co-scheduled streaming pass kinds %d/%d", p.load, p.store);
    return ADRT_B200_EINVAL;
}

// transposed: interior tiles on `s`, the tiles that reach offset D (masked program) on `side`, each a
// persistent launch with its own work counter (sc.next, sc.next + 1); both honour dep / done
template <int M, int LOADK, int STOREK>
int launch_bwd_p(const plan::Pass &p, const float *src, float *dst, const PassArgs &a, const SchedArgs &sc, cudaStream_t s,
                 cudaStream_t side)
{
    int xm = stile::BwdStream<M, LOADK, STOREK, false>::first_masked_tile(a.D);
    if (xm > p.grid_x) xm = p.grid_x;
    const int nmask = p.grid_x - xm;
    int side_ctas = sc.ctas / 4;
    if (side_ctas < 1) side_ctas = 1;
    int rc = launch_p<stile::BwdStream<M, LOADK, STOREK, false>>(src, dst, a, sc, 0, xm, sc.next, sc.ctas, s);
    if (rc != ADRT_B200_OK) return rc;
    return launch_p<stile::BwdStream<M, LOADK, STOREK, true>>(src, dst, a, sc, xm, nmask, sc.next + 1, side_ctas, side ? side : s);
}

template <int M>
int dispatch_bwd_p(const plan::Pass &p, const float *src, float *dst, const PassArgs &a, const SchedArgs &sc, cudaStream_t s,
                   cudaStream_t side)
{
    using namespace tile;
    if (p.load == LOAD_QCOLS && p.store == STORE_WROWS) return launch_bwd_p<M, LOAD_QCOLS, STORE_WROWS>(p, src, dst, a, sc, s, side);
    if (p.load == LOAD_WROWS && p.store == STORE_QCOLS) return launch_bwd_p<M, LOAD_WROWS, STORE_QCOLS>(p, src, dst, a, sc, s, side);
    set_error("internal: bad co-scheduled streaming pass kinds %d/%d", p.load, p.store);
    return ADRT_B200_EINVAL;
}

template <int M>
int dispatch_fwd(const plan::Pass &p, const float *src, float *dst, const PassArgs &a, cudaStream_t s)
{
    using namespace tile;
    if (p.load == LOAD_IMAGE && p.store == STORE_WROWS) return launch<stile::FwdStream<M, LOAD_IMAGE, STORE_WROWS>>(src, dst, a, 0, p.grid_x, p.grid_y, s);
    if (p.load == LOAD_IMAGE && p.store == STORE_QCOLS) return launch<stile::FwdStream<M, LOAD_IMAGE, STORE_QCOLS>>(src, dst, a, 0, p.grid_x, p.grid_y, s);
    if (p.load == LOAD_WROWS && p.store == STORE_WROWS) return launch<stile::FwdStream<M, LOAD_WROWS, STORE_WROWS>>(src, dst, a, 0, p.grid_x, p.grid_y, s);
    if (p.load == LOAD_WROWS && p.store == STORE_QCOLS) {
        if (M == 6 && plan::stream_split2()) return launch<stile::FwdStream<6, LOAD_WROWS, STORE_QCOLS, 2>>(src, dst, a, 0, p.grid_x, p.grid_y, s);
        return launch<stile::FwdStream<M, LOAD_WROWS, STORE_QCOLS>>(src, dst, a, 0, p.grid_x, p.grid_y, s);
    }
    set_error("internal: bad streaming pass kinds %d/%d", p.load, p.store);
    return ADRT_B200_EINVAL;
}

// A per-device helper stream: the few, slow boundary tiles run beside the interior tiles
// instead of after them (fork / join with events around the two launches).

// interior tiles and the tiles that reach offset D run as two launches (see BwdStream::phase_ct)
template <int M, int LOADK, int STOREK>
int launch_bwd(const plan::Pass &p, const float *src, float *dst, const PassArgs &a, cudaStream_t s)
{
    int xm = stile::BwdStream<M, LOADK, STOREK, false>::first_masked_tile(a.D);
    if (xm > p.grid_x) xm = p.grid_x;
    const int nmask = p.grid_x - xm;
    cudaStream_t side = (xm > 0 && nmask > 0) ? aux_stream(a.side_idx) : nullptr;
    // the helper streams are shared per device: while the caller's stream is being captured into a graph the
    // boundary tiles simply follow on the caller's stream (a fork into a shared stream would pull every other
    // user of that stream into the capture)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (side && cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) side = nullptr;
    (void)cudaGetLastError();
    cudaEvent_t fork = nullptr, join = nullptr;
    if (side) {
        if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&join, cudaEventDisableTiming) != cudaSuccess) {
            if (fork) cudaEventDestroy(fork);
            side = nullptr;
            (void)cudaGetLastError();
        }
    }
    int rc;
    if (side) {
        ADRT_CUDA_CHECK(cudaEventRecord(fork, s));
        ADRT_CUDA_CHECK(cudaStreamWaitEvent(side, fork, 0));
        rc = launch<stile::BwdStream<M, LOADK, STOREK, true>>(src, dst, a, xm, nmask, p.grid_y, side);
        if (rc == ADRT_B200_OK) rc = launch<stile::BwdStream<M, LOADK, STOREK, false>>(src, dst, a, 0, xm, p.grid_y, s);
        cudaEventRecord(join, side);
        cudaStreamWaitEvent(s, join, 0);
        cudaEventDestroy(fork);
        cudaEventDestroy(join);
        return rc;
    }
    rc = launch<stile::BwdStream<M, LOADK, STOREK, false>>(src, dst, a, 0, xm, p.grid_y, s);
    if (rc != ADRT_B200_OK) return rc;
    return launch<stile::BwdStream<M, LOADK, STOREK, true>>(src, dst, a, xm, nmask, p.grid_y, s);
}

template <int M>
int dispatch_bwd(const plan::Pass &p, const float *src, float *dst, const PassArgs &a, cudaStream_t s)
{
    using namespace tile;
    if (p.load == LOAD_QCOLS && p.store == STORE_WROWS) return launch_bwd<M, LOAD_QCOLS, STORE_WROWS>(p, src, dst, a, s);
    if (p.load == LOAD_QCOLS && p.store == STORE_QCOLS) return launch_bwd<M, LOAD_QCOLS, STORE_QCOLS>(p, src, dst, a, s);
    if (p.load == LOAD_WROWS && p.store == STORE_WROWS) return launch_bwd<M, LOAD_WROWS, STORE_WROWS>(p, src, dst, a, s);
    if (p.load == LOAD_WROWS && p.store == STORE_QCOLS) return launch_bwd<M, LOAD_WROWS, STORE_QCOLS>(p, src, dst, a, s);
    set_error("internal: bad streaming pass kinds %d/%d", p.load, p.store);
    return ADRT_B200_EINVAL;
}

}  // namespace

cudaStream_t aux_stream(int idx)
{
    static std::mutex mu;
    static cudaStream_t streams[64][4] = {};
    int dev = 0;
    if (idx < 0 || idx > 3 || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[dev][idx] && cudaStreamCreateWithFlags(&streams[dev][idx], cudaStreamNonBlocking) != cudaSuccess)
        streams[dev][idx] = nullptr;
    return streams[dev][idx];
}

int launch_stream_pass_sched(const plan::Pass &p, bool forward, const float *src, float *dst, const PassArgs &a,
                             const SchedArgs &sc, cudaStream_t s, cudaStream_t side)
{
    if (forward) {
        if (p.M == 6) return dispatch_fwd_p<6>(p, src, dst, a, sc, s);
        if (p.M == 5) return dispatch_fwd_p<5>(p, src, dst, a, sc, s);
    } else {
        if (p.M == 6) return dispatch_bwd_p<6>(p, src, dst, a, sc, s, side);
        if (p.M == 5) return dispatch_bwd_p<5>(p, src, dst, a, sc, s, side);
    }
    set_error("internal: no co-scheduled streaming kernel for M=%d forward=%d", p.M, (int)forward);
    return ADRT_B200_EINVAL;
}

int launch_stream_pass(const plan::Pass &p, bool forward, const float *src, float *dst, const PassArgs &a, cudaStream_t s)
{
    if (forward) {
        if (p.M == 6) return dispatch_fwd<6>(p, src, dst, a, s);
        if (p.M == 5) return dispatch_fwd<5>(p, src, dst, a, s);
    }
    if (!forward) {
        if (p.M == 6) return dispatch_bwd<6>(p, src, dst, a, s);
        if (p.M == 5) return dispatch_bwd<5>(p, src, dst, a, s);
    }
    set_error("internal: no streaming kernel for M=%d forward=%d", p.M, (int)forward);
    return ADRT_B200_EINVAL;
}

}  // namespace adrt_b200
