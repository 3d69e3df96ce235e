// common.cuh -- shared helpers for libvtts_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/vtts_b200.h"

namespace vtts {

// thread-local last-error buffer behind vtts_last_error()
char *error_buffer();
int set_error(int code, const char *fmt, ...);

#define VTTS_CHECK_CUDA(expr)                                                               \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::vtts::set_error(VTTS_E_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__,    \
                                     #expr, cudaGetErrorString(_e));                        \
    } while (0)

#define VTTS_CHECK_LAUNCH()                                                                 \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess)                                                              \
            return ::vtts::set_error(VTTS_E_CUDA, "%s:%d: kernel launch -> %s", __FILE__,   \
                                     __LINE__, cudaGetErrorString(_e));                     \
    } while (0)

#define VTTS_REQUIRE(cond, ...)                                                             \
    do {                                                                                    \
        if (!(cond)) return ::vtts::set_error(VTTS_E_INVALID, __VA_ARGS__);                 \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- programmatic dependent launch -------------------------------------------------------------------
// The generator is a chain of ~60 dependent kernels.  Each one is launched with the programmatic-serialization
// attribute: its CTAs may become resident while the previous kernel drains and run their prologue (barrier init, TMEM
// allocation, tensor-map prefetch, resident weight loads - nothing the previous kernel produces); grid_dep_wait()
// then blocks until the previous grid has completed and its writes are visible.  grid_dep_launch() is issued only
// AFTER the wait, so by induction everything older than the previous kernel is complete before any prologue runs.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // VTTS_PDL=0 disables (conv_tc.cu)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                           bool pdl, unsigned cluster_x, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        attrs[n].id = cudaLaunchAttributeClusterDimension;
        attrs[n].val.clusterDim.x = cluster_x;
        attrs[n].val.clusterDim.y = 1;
        attrs[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl) {
        attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// ---------------------------------------------------------------------------------------
// fp32 direct convolution (conv_fp32.cu) -- channels-first (B, C, L)
// ---------------------------------------------------------------------------------------
struct ConvFp32Params {
    const float *x;      // (B, cin, L_in)
    const float *w;      // packed [phase][ci][tap][co]  (see pack.cu)
    const float *bias;   // (cout) or null
    const float *bias_b; // (B, cout) per-batch extra bias (global conditioning) or null
    const float *res;    // (B, cout, L_out) or null
    float *y;            // (B, cout, L_out)
    int B, cin, cout, L_in, L_out;
    int taps;            // taps per phase
    int tap_off0;        // x index = i + tap_off0 + j * tap_step
    int tap_step;
    int phases;          // 1 for Conv1d; stride s for ConvTranspose1d (polyphase)
    int out_stride;      // output index = i * out_stride + out_off0 + phase
    int out_off0;
    int n_pos;           // number of positions i per phase
    float slope_in;      // LeakyReLU slope applied to x on load (1 = identity)
    int accumulate;      // y = y_old + value (MRF sum, generator.py:152)
    float divide_by;     // > 0: value = value / divide_by (generator.py:153)
    int apply_tanh;      // generator.py:119
};
int launch_conv_fp32(const ConvFp32Params &p, cudaStream_t stream);

// weight packing (pack.cu)
// fold weight norm: w[i, :, :] = g[i] * v[i, :, :] / ||v[i, :, :]||   (rows = dim0)
int launch_fold_weight_norm(const float *v, const float *g, float *w, int dim0, int inner,
                            cudaStream_t stream);
// Conv1d (cout,cin,k) -> [1][ci][tap][co]
int launch_pack_conv_fp32(const float *w, float *packed, int cout, int cin, int k,
                          cudaStream_t stream);
// ConvTranspose1d (cin,cout,k), stride s, k % s == 0 -> [phase q][ci][j][co], tap kk = q + j*s
int launch_pack_convT_fp32(const float *w, float *packed, int cin, int cout, int k, int s,
                           cudaStream_t stream);

// layout helpers (pack.cu)
// (B, C, L) fp32 channels-first <-> (B, L, C) channels-last
// 16-bit operand formats of the tensor-core path
enum { VTTS_FMT_BF16 = 0, VTTS_FMT_FP16 = 1 };
// fp16 conversions saturate to the largest finite value instead of overflowing to inf (one F2FP.SATFINITE instruction)
__device__ __forceinline__ uint16_t cvt16(float v, int fmt) {
    if (fmt == VTTS_FMT_BF16) return __bfloat16_as_ushort(__float2bfloat16(v));
    uint16_t h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return h;
}
// two values -> one 32-bit register, `lo` in the low half (same rounding/saturation as cvt16)
__device__ __forceinline__ uint32_t cvt16x2(float lo, float hi, int fmt) {
    uint32_t r;
    if (fmt == VTTS_FMT_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float cvt16_to_f32(uint16_t u, int fmt) {
    return fmt == VTTS_FMT_BF16 ? __bfloat162float(__ushort_as_bfloat16(u)) : __half2float(__ushort_as_half(u));
}
int launch_cf_to_cl_16(const float *x, uint16_t *y, int B, int C, int L, int Cpad, float slope, int fmt,
                       cudaStream_t stream);
int launch_cl_to_cf_f32(const float *x, float *y, int B, int C, int L, cudaStream_t stream);
int launch_cl_16_to_cf_f32(const uint16_t *x, float *y, int B, int C, int L, int Cld, int fmt,
                           cudaStream_t stream);

}  // namespace vtts
