// dwconv.cu -- depthwise Conv1d of the Conformer convolution module, fused with the gate in front of it and the
// normalisation + activation behind it (conformer variant of the acoustic decoder).
//
// Replaces, in ConformerConvModule.sequential (models/tts/fastspeech2/blocks/conformer.py:470-480):
//     GLU(dim=1)                      u[c] = pw[c] * sigmoid(pw[C + c])                       (blocks/utils.py:75-86)
//     DepthwiseConv1d(C, C, k, pad)   v[c, t] = sum_j w[c, j] * u[c, t + j - (k-1)/2]         (:532-570, groups = C, no bias)
//     BatchNorm1d (eval)              folded by the caller into w and `bias`
//     Swish()                         y = v * sigmoid(v)                                      (blocks/utils.py:63-72)
// Input: the pointwise conv's fp32 output, channels-last (B, L, 2C); output: the next pointwise conv's 16-bit operand,
// channels-last (B, L, C).  HBM-bound: 8 B read + 2 B written per output element; fp32 arithmetic.
// A CTA owns DW_TT positions of one batch row and all channels: the gated tile (+ halo) is staged in shared memory
// [position][channel] (lanes = channels: coalesced and conflict-free), each thread keeps the k taps of its channel in
// registers and slides over the positions.
#include "common.cuh"

namespace vtts {
namespace {

constexpr int DW_THREADS = 256;
constexpr int DW_TT = 64;        // output positions per CTA
constexpr int DW_KMAX = 31;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

template <int FMT>
__global__ void __launch_bounds__(DW_THREADS)
dwconv_glu_swish_kernel(const float *__restrict__ pw, const float *__restrict__ w, const float *__restrict__ bias,
                        uint16_t *__restrict__ out, int L, int C, int k) {
    extern __shared__ float s_u[];                     // [DW_TT + k - 1][C]
    const int b = blockIdx.y, t0 = blockIdx.x * DW_TT, half = (k - 1) / 2;
    const int rows = DW_TT + k - 1;
    const float *pb = pw + (size_t)b * L * 2 * C;
    for (int idx = threadIdx.x; idx < rows * C; idx += DW_THREADS) {
        const int r = idx / C, c = idx - r * C;
        const int t = t0 - half + r;
        float u = 0.f;                                  // zero padding of the depthwise conv
        if (t >= 0 && t < L) {
            const float a = __ldg(pb + (size_t)t * 2 * C + c), g = __ldg(pb + (size_t)t * 2 * C + C + c);
            u = a * sigmoidf_(g);
        }
        s_u[idx] = u;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += DW_THREADS) {
        float wk[DW_KMAX];
#pragma unroll
        for (int j = 0; j < DW_KMAX; ++j) wk[j] = j < k ? __ldg(w + (size_t)c * k + j) : 0.f;
        const float bv = bias ? __ldg(bias + c) : 0.f;
        for (int i = 0; i < DW_TT; ++i) {
            const int t = t0 + i;
            if (t >= L) break;
            float acc = bv;
#pragma unroll
            for (int j = 0; j < DW_KMAX; ++j)
                if (j < k) acc = fmaf(wk[j], s_u[(i + j) * C + c], acc);
            const float y = acc * sigmoidf_(acc);
            out[((size_t)b * L + t) * C + c] = cvt16(y, FMT);
        }
    }
}

}  // namespace
}  // namespace vtts

using namespace vtts;

extern "C" int vtts_dwconv_glu_swish(const float *pw, const float *w, const float *bias, void *out16, int precision, int B,
                                     int L, int C, int ksize, vtts_stream_t stream) {
    VTTS_REQUIRE(B >= 0 && L >= 0 && C >= 1, "vtts_dwconv_glu_swish: bad shape");
    if (B == 0 || L == 0) return VTTS_OK;
    VTTS_REQUIRE(pw && w && out16, "vtts_dwconv_glu_swish: null pointer");
    VTTS_REQUIRE(ksize >= 1 && ksize % 2 == 1 && ksize <= DW_KMAX, "vtts_dwconv_glu_swish: kernel size must be odd and <= 31 (got %d)", ksize);
    if (precision != VTTS_PRECISION_BF16 && precision != VTTS_PRECISION_FP16)
        return set_error(VTTS_E_INVALID, "vtts_dwconv_glu_swish: precision must be bf16 or fp16");
    const size_t smem = (size_t)(DW_TT + ksize - 1) * C * sizeof(float);
    if (smem > 200 * 1024) return set_error(VTTS_E_UNSUPPORTED, "vtts_dwconv_glu_swish: %d channels need %zu B shared memory", C, smem);
    dim3 grid((unsigned)ceil_div(L, DW_TT), (unsigned)B);
    if (grid.y > 65535) return set_error(VTTS_E_UNSUPPORTED, "vtts_dwconv_glu_swish: batch %d too large", B);
    if (precision == VTTS_PRECISION_BF16) {
        if (smem > 48 * 1024) VTTS_CHECK_CUDA(cudaFuncSetAttribute(dwconv_glu_swish_kernel<VTTS_FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dwconv_glu_swish_kernel<VTTS_FMT_BF16><<<grid, DW_THREADS, smem, (cudaStream_t)stream>>>(pw, w, bias, (uint16_t *)out16, L, C, ksize);
    } else {
        if (smem > 48 * 1024) VTTS_CHECK_CUDA(cudaFuncSetAttribute(dwconv_glu_swish_kernel<VTTS_FMT_FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dwconv_glu_swish_kernel<VTTS_FMT_FP16><<<grid, DW_THREADS, smem, (cudaStream_t)stream>>>(pw, w, bias, (uint16_t *)out16, L, C, ksize);
    }
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}
