// Fused multi-stage ADRT / bdrt passes for sm_100a.
//
// One CTA = one tile of one group of one (image, quadrant) plane; the tile
// algorithms are in fused_tile.h, the pass planning in fused_plan.h.  HBM
// traffic per transform drops from 2*K sinogram sweeps (per-stage kernels) to
// one sweep per pass (2 passes up to n = 4096 in fp32).
#include "pass_args.h"

#include <type_traits>

namespace adrt_b200 {

namespace {


// The phases of a full tile, unrolled at compile time with a barrier after each.
template <typename Prog, typename T, int PH>
__device__ __forceinline__ void all_phases(int mode, T *buf, T (&regs)[tile::NREG], const T *sp, T *dp,
                                           const tile::TileCtx &c, int tid)
{
    if constexpr (PH < Prog::kPhases) {
        Prog::template phase_ct<PH>(mode, buf, regs, sp, dp, c, tid);
        __syncthreads();
        all_phases<Prog, T, PH + 1>(mode, buf, regs, sp, dp, c, tid);
    }
}

template <typename T, int M, int LOADK, int STOREK, bool kForward>
__global__ void __launch_bounds__(tile::Geo<M>::NT, sizeof(T) == 8 ? tile::Geo<M>::MIN_CTAS / 2 : tile::Geo<M>::MIN_CTAS)
pass_kernel(const T *__restrict__ src, T *__restrict__ dst, PassArgs a)
{
    using Prog = typename std::conditional<kForward, tile::FwdProgram<T, M, LOADK, STOREK>,
                                           tile::BwdProgram<T, M, LOADK, STOREK>>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *buf = reinterpret_cast<T *>(smem_raw);
    T regs[tile::NREG];

    tile::TileCtx c;
    c.n = a.n;
    c.D = a.D;
    c.e = a.e;
    c.g = blockIdx.y;
    c.k0 = c.g >> a.loge;
    c.a_g = c.g & (a.e - 1);
    c.d0 = blockIdx.x * Prog::TD;
    c.next_g = a.next_g;
    c.d_need = a.d_need;
    c.in_pitch = a.in_pitch;
    c.out_pitch = a.out_pitch;
    c.q = 0;
    const int mode = Prog::classify(c);
    if (mode == tile::TILE_SKIP) return;

    for (int plane = blockIdx.z; plane < a.planes; plane += gridDim.z) {
        const T *sp;
        if (LOADK == tile::LOAD_IMAGE) {
            c.q = a.q_first + plane % a.q_count;
            sp = src + (long long)(plane / a.q_count) * a.src_plane_stride;
        } else {
            sp = src + (long long)plane * a.src_plane_stride;
        }
        T *dp = dst + (long long)plane * a.dst_plane_stride;
        if (mode == tile::TILE_ZERO) {
            Prog::zero_tile(buf, dp, c, threadIdx.x);
        } else {
            all_phases<Prog, T, 0>(mode, buf, regs, sp, dp, c, threadIdx.x);
        }
    }
}

template <typename T, int M, int LOADK, int STOREK, bool kForward>
int launch_pass(const T *src, T *dst, const PassArgs &a, int grid_x, int grid_y, cudaStream_t s)
{
    auto kern = pass_kernel<T, M, LOADK, STOREK, kForward>;
    size_t smem = (size_t)tile::Geo<M>::G * tile::Pitch<T>::value * sizeof(T);
    if (const char *e = getenv("ADRT_B200_SMEM_PAD_KB")) smem += (size_t)atoi(e) * 1024;  // occupancy experiments
    // per device, so not cached in a static: a process may drive several GPUs
    ADRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)grid_x, (unsigned)grid_y, (unsigned)(a.planes < 65535 ? a.planes : 65535));
    kern<<<grid, tile::Geo<M>::NT, smem, s>>>(src, dst, a);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T, int M, bool kForward>
int dispatch_kinds(int load, int store, const T *src, T *dst, const PassArgs &a, int gx, int gy, cudaStream_t s)
{
    using namespace tile;
    if constexpr (kForward) {
        if (load == LOAD_IMAGE && store == STORE_WROWS) return launch_pass<T, M, LOAD_IMAGE, STORE_WROWS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_IMAGE && store == STORE_QCOLS) return launch_pass<T, M, LOAD_IMAGE, STORE_QCOLS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_WROWS && store == STORE_WROWS) return launch_pass<T, M, LOAD_WROWS, STORE_WROWS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_WROWS && store == STORE_QCOLS) return launch_pass<T, M, LOAD_WROWS, STORE_QCOLS, kForward>(src, dst, a, gx, gy, s);
    } else {
        if (load == LOAD_QCOLS && store == STORE_WROWS) return launch_pass<T, M, LOAD_QCOLS, STORE_WROWS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_QCOLS && store == STORE_QCOLS) return launch_pass<T, M, LOAD_QCOLS, STORE_QCOLS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_WROWS && store == STORE_WROWS) return launch_pass<T, M, LOAD_WROWS, STORE_WROWS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_WROWS && store == STORE_QCOLS) return launch_pass<T, M, LOAD_WROWS, STORE_QCOLS, kForward>(src, dst, a, gx, gy, s);
    }
    set_error("internal: bad pass kinds %d/%d", load, store);
    return ADRT_B200_EINVAL;
}

template <typename T, bool kForward>
int dispatch_pass(const plan::Pass &p, const T *src, T *dst, const PassArgs &a, cudaStream_t s)
{
    if (p.stream) {
        if constexpr (std::is_same<T, float>::value) return launch_stream_pass(p, kForward, src, dst, a, s);
        set_error("internal: streaming passes are fp32 only");
        return ADRT_B200_EINVAL;
    }
    switch (p.M) {
    case 1: return dispatch_kinds<T, 1, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 2: return dispatch_kinds<T, 2, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 3: return dispatch_kinds<T, 3, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 4: return dispatch_kinds<T, 4, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 5: return dispatch_kinds<T, 5, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 6: return dispatch_kinds<T, 6, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    }
    set_error("internal: unsupported stages per pass %d", p.M);
    return ADRT_B200_EINVAL;
}

// Images per wave: the batch is processed in waves so that the R-layout
// workspace of a wave can stay resident in the 126 MB L2 between the pass
// that writes it and the pass that reads it.  0 = whole batch at once.
int wave_images(int64_t B)
{
    const char *e = getenv("ADRT_B200_WAVE");  // read per call: tunable at run time
    const int w = e ? atoi(e) : 0;
    if (w <= 0 || w > B) return (int)B;
    return w;
}

template <typename T, bool kForward>
int run_plan(const plan::Plan &pl, const T *in, T *out, int64_t B, int q_first, int q_count, T *ws, size_t ws_elems,
             cudaStream_t s)
{
    // B images of q_count planes each (forward: quadrants q_first .. q_first+q_count-1 of
    // every image; transposed: any plane count, the quadrant identity does not matter)
    const int n = pl.n, D = pl.D;
    const int wave = wave_images(B);
    const size_t slot0 = pl.ws_slot_elems[0] * q_count * (size_t)wave;
    const size_t slot1 = pl.ws_slot_elems[1] * q_count * (size_t)wave;
    if (slot0 + slot1 > ws_elems) {
        set_error("fused workspace too small: need %zu elements, got %zu", slot0 + slot1, ws_elems);
        return ADRT_B200_EWORKSPACE;
    }
    T *slot[2] = {ws, ws + slot0};
    const long long img_elems = (long long)n * n, sino_plane = (long long)D * n;
    for (int64_t b0 = 0; b0 < B; b0 += wave) {
        const int nb = (int)((B - b0) < wave ? (B - b0) : wave);
        for (int i = 0; i < pl.npass; ++i) {
            const plan::Pass &p = pl.pass[i];
            PassArgs a;
            a.n = n; a.D = D; a.e = 1 << p.s; a.loge = p.s; a.next_g = p.next_g; a.d_need = p.d_need;
            a.in_pitch = p.in_pitch; a.out_pitch = p.out_pitch;
            a.planes = nb * q_count;
            a.x_off = 0;
            a.q_first = q_first; a.q_count = q_count;
            const T *src;
            T *dst;
            if (p.src_buf < 0) {
                if (kForward) { src = in + b0 * img_elems; a.src_plane_stride = img_elems; }
                else { src = in + b0 * q_count * sino_plane; a.src_plane_stride = sino_plane; }
            } else {
                src = slot[p.src_buf];
                a.src_plane_stride = (long long)n * p.in_pitch;
            }
            if (p.dst_buf < 0) {
                dst = out + b0 * q_count * sino_plane;
                a.dst_plane_stride = sino_plane;
            } else {
                dst = slot[p.dst_buf];
                a.dst_plane_stride = (long long)n * p.out_pitch;
            }
            int rc = dispatch_pass<T, kForward>(p, src, dst, a, s);
            if (rc != ADRT_B200_OK) return rc;
        }
    }
    return ADRT_B200_OK;
}

}  // namespace

template <typename T>
size_t fused_adrt_workspace_elems(int64_t B, int64_t n, int q_count)
{
    plan::Plan pl;
    if (n > kMaxN || !plan::make_forward_plan(n, sizeof(T), &pl)) return (size_t)-1;
    return (pl.ws_slot_elems[0] + pl.ws_slot_elems[1]) * q_count * (size_t)wave_images(B);
}

template <typename T>
size_t fused_bdrt_workspace_elems(int64_t B, int64_t n, int q_count)
{
    plan::Plan pl;
    if (n > kMaxN || !plan::make_transposed_plan(n, sizeof(T), &pl)) return (size_t)-1;
    return (pl.ws_slot_elems[0] + pl.ws_slot_elems[1]) * q_count * (size_t)wave_images(B);
}

template <typename T>
int fused_adrt(const T *in, T *out, int64_t B, int64_t n, int q_first, int q_count, T *ws, size_t ws_elems,
               cudaStream_t s, bool *handled)
{
    plan::Plan pl;
    *handled = false;
    if (n > kMaxN || !plan::make_forward_plan(n, sizeof(T), &pl)) return ADRT_B200_OK;
    *handled = true;
    return run_plan<T, true>(pl, in, out, B, q_first, q_count, ws, ws_elems, s);
}

template <typename T>
int fused_bdrt(const T *in, T *out, int64_t B, int64_t n, int q_count, int64_t rows, T *ws, size_t ws_elems,
               cudaStream_t s, bool *handled)
{
    // rows < 2n-1: only offsets d < rows of every output plane are computed, the rest of
    // `out` is left as it was
    plan::Plan pl;
    *handled = false;
    if (n > kMaxN || !plan::make_transposed_plan(n, sizeof(T), &pl, rows)) return ADRT_B200_OK;
    *handled = true;
    return run_plan<T, false>(pl, in, out, B, 0, q_count, ws, ws_elems, s);
}

#define INSTANTIATE(T)                                                      \
    template size_t fused_adrt_workspace_elems<T>(int64_t, int64_t, int);   \
    template size_t fused_bdrt_workspace_elems<T>(int64_t, int64_t, int);   \
    template int fused_adrt<T>(const T *, T *, int64_t, int64_t, int, int, T *, size_t, cudaStream_t, bool *); \
    template int fused_bdrt<T>(const T *, T *, int64_t, int64_t, int, int64_t, T *, size_t, cudaStream_t, bool *);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace adrt_b200
