// Shared host/device helpers for the adrt_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/adrt_b200.h"

namespace adrt_b200 {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launch_count;
extern std::atomic<int> g_mode;

#define ADRT_CUDA_CHECK(expr)                                                          \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            ::adrt_b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                   __FILE__, __LINE__);                                \
            return ADRT_B200_ECUDA;                                                    \
        }                                                                              \
    } while (0)

#define ADRT_LAUNCH_CHECK()                                                            \
    do {                                                                               \
        ::adrt_b200::g_launch_count.fetch_add(1, std::memory_order_relaxed);           \
        ADRT_CUDA_CHECK(cudaGetLastError());                                           \
    } while (0)

#define ADRT_REQUIRE(cond, ...)                                                        \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            ::adrt_b200::set_error(__VA_ARGS__);                                       \
            return ADRT_B200_EINVAL;                                                   \
        }                                                                              \
    } while (0)

// ---- shape helpers (adrt_cdefs_common.cpp:140-142, 176-192) -------------------
inline bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }
inline int num_iters(int64_t n)
{
    if (n <= 0) return 0;
    int bw = 0;
    for (uint64_t v = (uint64_t)n; v; v >>= 1) ++bw;
    return bw - (is_pow2(n) ? 1 : 0);
}
inline size_t dtype_size(int dtype) { return dtype == ADRT_B200_F64 ? 8 : 4; }
inline bool dtype_ok(int dtype) { return dtype == ADRT_B200_F32 || dtype == ADRT_B200_F64; }
// Largest image side the kernels index with 32-bit in-plane offsets: D*n < 2^31.
constexpr int64_t kMaxN = 16384;

inline int64_t sino_elems(int64_t B, int64_t n) { return B * 4 * (2 * n - 1) * n; }

// ---- per-dtype launchers implemented in the .cu files --------------------------
template <typename T> int launch_adrt_init(const T *in, T *out, int64_t B, int64_t n, cudaStream_t s);
template <typename T> int launch_adrt_step(const T *in, T *out, int64_t B, int64_t n, int step, cudaStream_t s);
template <typename T> int launch_bdrt_step(const T *in, T *out, int64_t B, int64_t n, int step, bool core_semantics, cudaStream_t s);
template <typename T> int launch_iadrt_stage(const T *in, T *out, int64_t B, int64_t n, int stage, cudaStream_t s);
template <typename T> int launch_fmg_restriction(const T *in, T *out, int64_t B, int64_t n, cudaStream_t s);
template <typename T> int launch_fmg_prolongation(const T *in, T *out, int64_t B, int64_t h, int64_t w, cudaStream_t s);
template <typename T> int launch_fmg_highpass(const T *in, T *out, int64_t B, int64_t h, int64_t w, cudaStream_t s);
template <typename T> int launch_interp_to_cart(const T *in, T *out, const float *t, const int32_t *base, const float *h_base, const float *cosv, const int32_t *sgn, const T *factor, int64_t B, int64_t n, cudaStream_t s);
template <typename T> int launch_truncate(const T *in, T *out, int64_t B, int64_t n, cudaStream_t s);
template <typename T> int launch_truncate_mean(const T *in, T *out, int64_t B, int64_t n, T divisor, cudaStream_t s);
template <typename T> int launch_truncate_mean_shares(const T *in, T *out, int64_t B, int64_t n, int per, int parts, T divisor, cudaStream_t s);
template <typename T> int launch_stitch(const T *in, T *out, int64_t B, int64_t n, bool remove_repeated, cudaStream_t s);
template <typename T> int launch_unstitch(const T *in, T *out, int64_t B, int64_t n, bool trimmed, cudaStream_t s);
template <typename T> int launch_binary(const T *a, const T *b, T *out, int64_t count, int op, cudaStream_t s);
// CG vector updates (step_kernels.cu): state = {r.r, p.Ap, new r.r, ...} in double on the device
size_t cg_workspace_bytes();
template <typename T> int launch_cg_dot(const T *a, const T *b, double *state, int slot, int64_t count, double *partials, cudaStream_t s);
template <typename T> int launch_cg_update(T *x, T *r, const T *p, const T *ap, double *state, int64_t count, double *partials, cudaStream_t s);
template <typename T> int launch_cg_direction(T *p, const T *r, double *state, int64_t count, cudaStream_t s);

// Fused multi-stage paths (fused_adrt.cu).  Return ADRT_B200_OK or an error;
// `handled` is false when the shape is left to the per-stage path.
// B images with q_count planes each (forward: quadrants q_first..q_first+q_count-1).
template <typename T> size_t fused_adrt_workspace_elems(int64_t B, int64_t n, int q_count);
template <typename T> size_t fused_bdrt_workspace_elems(int64_t B, int64_t n, int q_count);
// rows_out / rows_in: the caller-side array is in R-layout rows (planes x n x round4(2n-1)) instead of the
// public (d, column) layout: how the fused normal operator hands adrt's result to bdrt
template <typename T> int fused_adrt(const T *in, T *out, int64_t B, int64_t n, int q_first, int q_count, T *ws, size_t ws_elems, cudaStream_t s, bool *handled, bool rows_out = false);
// `sub` (same shape as `in`, public layout): the transform of in - sub, the subtraction done by the first
// pass's loader (element-wise, so bit-identical to a separate subtraction); only where fused_bdrt_sub_ok
template <typename T> int fused_bdrt(const T *in, T *out, int64_t B, int64_t n, int q_count, int64_t rows, T *ws, size_t ws_elems, cudaStream_t s, bool *handled, bool rows_in = false, const T *sub = nullptr);
template <typename T> bool fused_bdrt_sub_ok(int64_t n, int64_t rows);

// adrt.iadrt as fused multi-stage passes (iadrt_fused.cu); workspace in elements
template <typename T> size_t fused_iadrt_workspace_elems(int64_t B, int64_t n);
template <typename T> int fused_iadrt(const T *in, T *out, int64_t B, int64_t n, T *ws, size_t ws_elems, cudaStream_t s);

// Angle-block sharding of single large images over `parts` ranks (fused_plan.h part_*).
// phase 0 writes / phase 1 reads the exchange buffer `xbuf` (planes x n rows x part_exchange_pitch elements).
template <typename T> size_t part_exchange_pitch(int64_t n, int m_last, bool forward);
template <typename T> size_t part_exchange_cols(int64_t n, int m_last, int64_t rows);
template <typename T> size_t part_workspace_elems(int64_t planes, int64_t n, int m_last);
template <typename T> int fused_adrt_part(const T *img, T *xbuf, T *sino, int64_t B, int64_t n, int q_first, int q_count, int part, int parts, int m_last, int phase, T *ws, size_t ws_elems, cudaStream_t s);
template <typename T> int fused_bdrt_part(const T *sino, T *xbuf, T *out, int64_t planes, int64_t n, int64_t rows, int part, int parts, int m_last, int phase, T *ws, size_t ws_elems, cudaStream_t s);

}  // namespace adrt_b200
