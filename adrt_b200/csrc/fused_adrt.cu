// Fused multi-stage ADRT / bdrt kernels (placeholder: per-stage path only).
#include "common.cuh"

namespace adrt_b200 {

template <typename T> size_t fused_adrt_workspace_elems(int64_t, int64_t) { return (size_t)-1; }
template <typename T> size_t fused_bdrt_workspace_elems(int64_t, int64_t) { return (size_t)-1; }
template <typename T> int fused_adrt(const T *, T *, int64_t, int64_t, T *, size_t, cudaStream_t, bool *handled) { *handled = false; return ADRT_B200_OK; }
template <typename T> int fused_bdrt(const T *, T *, int64_t, int64_t, T *, size_t, cudaStream_t, bool *handled) { *handled = false; return ADRT_B200_OK; }

#define INSTANTIATE(T)                                                      \
    template size_t fused_adrt_workspace_elems<T>(int64_t, int64_t);        \
    template size_t fused_bdrt_workspace_elems<T>(int64_t, int64_t);        \
    template int fused_adrt<T>(const T *, T *, int64_t, int64_t, T *, size_t, cudaStream_t, bool *); \
    template int fused_bdrt<T>(const T *, T *, int64_t, int64_t, T *, size_t, cudaStream_t, bool *);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace adrt_b200
