// conv_fp32.cu -- CUDA-core fp32 direct convolution (channels-first), the in-repo fp32 path.
//
// One kernel serves every layer of the generator (generator.py:132-156, layers.py:93-97):
//   Conv1d            x index = i + (j - (k-1)/2) * dilation            (1 phase)
//   ConvTranspose1d   polyphase: phase q = (t_out + p) mod s, taps kk = q + j*s read x[i0 - j]
//                     (SURVEY.md appendix 9.1), output t_out = i0*s + q - p
// with LeakyReLU fused on the operand load and bias / per-batch bias / residual / MRF
// accumulate / divide / tanh fused in the epilogue.  Register tile 8 (co) x 4 (t) per thread,
// 64 x 128 per CTA, operands staged in shared memory.  This path exists for parity (fp32,
// ~1e-6 of the CPU oracle) and for configurations the tcgen05 path does not cover; it is not
// the performance path.
#include "common.cuh"

namespace vtts {

constexpr int F32_TILE_CO = 64;
constexpr int F32_TILE_T = 128;
constexpr int F32_CI = 8;
constexpr int F32_THREADS = 256;

__global__ void __launch_bounds__(F32_THREADS)
conv_fp32_kernel(ConvFp32Params p) {
    extern __shared__ __align__(16) float smem[];
    const int span = (p.taps - 1) * (p.tap_step < 0 ? -p.tap_step : p.tap_step);
    const int XW = F32_TILE_T + span;
    float *sx = smem;                              // [F32_CI][XW]
    float *sw = smem + F32_CI * XW;                // [F32_CI][taps][F32_TILE_CO]

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int b = blockIdx.z / p.phases, phase = blockIdx.z % p.phases;
    const int i0 = blockIdx.x * F32_TILE_T;
    const int co0 = blockIdx.y * F32_TILE_CO;
    const int last_off = p.tap_off0 + (p.taps - 1) * p.tap_step;
    const int min_off = p.tap_off0 < last_off ? p.tap_off0 : last_off;

    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[a][r] = 0.f;

    const float *xb = p.x + (size_t)b * p.cin * p.L_in;
    const float *wp = p.w + (size_t)phase * p.cin * p.taps * p.cout;

    for (int c0 = 0; c0 < p.cin; c0 += F32_CI) {
        // stage activations (LeakyReLU fused, zero "same" padding outside [0, L_in))
        for (int idx = threadIdx.x; idx < F32_CI * XW; idx += F32_THREADS) {
            int ci = idx / XW, o = idx - ci * XW;
            int gx = i0 + min_off + o;
            float v = 0.f;
            if (c0 + ci < p.cin && gx >= 0 && gx < p.L_in)
                v = lrelu(__ldg(xb + (size_t)(c0 + ci) * p.L_in + gx), p.slope_in);
            sx[idx] = v;
        }
        // stage weights [ci][tap][co]
        const int wn = F32_CI * p.taps * F32_TILE_CO;
        for (int idx = threadIdx.x; idx < wn; idx += F32_THREADS) {
            int co = idx % F32_TILE_CO;
            int rest = idx / F32_TILE_CO;
            int j = rest % p.taps, ci = rest / p.taps;
            float v = 0.f;
            if (c0 + ci < p.cin && co0 + co < p.cout)
                v = __ldg(wp + ((size_t)(c0 + ci) * p.taps + j) * p.cout + co0 + co);
            sw[idx] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < F32_CI; ++ci) {
            const float *sxr = sx + ci * XW + tx;
            const float *swr = sw + (ci * p.taps) * F32_TILE_CO + ty * 8;
            for (int j = 0; j < p.taps; ++j) {
                const int o = p.tap_off0 + j * p.tap_step - min_off;
                const float4 w0 = *reinterpret_cast<const float4 *>(swr + j * F32_TILE_CO);
                const float4 w1 = *reinterpret_cast<const float4 *>(swr + j * F32_TILE_CO + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                float xv[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) xv[r] = sxr[o + 32 * r];
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[a][r] = fmaf(wv[a], xv[r], acc[a][r]);
            }
        }
        __syncthreads();
    }

    // fused epilogue
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int co = co0 + ty * 8 + a;
        if (co >= p.cout) continue;
        float bsum = p.bias ? __ldg(p.bias + co) : 0.f;
        const float bb = p.bias_b ? __ldg(p.bias_b + (size_t)b * p.cout + co) : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + tx + 32 * r;
            if (i >= p.n_pos) continue;
            const int t = i * p.out_stride + p.out_off0 + phase;
            if (t < 0 || t >= p.L_out) continue;
            const size_t o = ((size_t)b * p.cout + co) * p.L_out + t;
            float v = acc[a][r] + bsum;
            if (p.bias_b) v = v + bb;
            if (p.res) v = v + __ldg(p.res + o);
            if (p.accumulate) v = p.y[o] + v;
            if (p.divide_by > 0.f) v = __fdiv_rn(v, p.divide_by);
            if (p.apply_tanh) v = tanhf(v);
            p.y[o] = v;
        }
    }
}

int launch_conv_fp32(const ConvFp32Params &p, cudaStream_t stream) {
    if (p.B <= 0 || p.n_pos <= 0 || p.cout <= 0) return VTTS_OK;
    const int span = (p.taps - 1) * (p.tap_step < 0 ? -p.tap_step : p.tap_step);
    const size_t smem = sizeof(float) * ((size_t)F32_CI * (F32_TILE_T + span) +
                                         (size_t)F32_CI * p.taps * F32_TILE_CO);
    if (smem > 200 * 1024)
        return set_error(VTTS_E_UNSUPPORTED, "conv_fp32: taps=%d step=%d needs %zu B smem", p.taps,
                         p.tap_step, smem);
    if (smem > 48 * 1024)
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(conv_fp32_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(p.n_pos, F32_TILE_T), (unsigned)ceil_div(p.cout, F32_TILE_CO),
              (unsigned)(p.B * p.phases));
    if (grid.z > 65535 || grid.y > 65535)
        return set_error(VTTS_E_UNSUPPORTED, "conv_fp32: grid too large (B*phases=%u)", grid.z);
    conv_fp32_kernel<<<grid, F32_THREADS, smem, stream>>>(p);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

}  // namespace vtts

using namespace vtts;

// test hook: plain Conv1d, weights in the reference layout (cout, cin, k) -- packed on the fly
extern "C" int vtts_dbg_conv1d_fp32(const float *x, const float *w, const float *bias,
                                    const float *res, float *y, int B, int cin, int cout, int L,
                                    int ksize, int dilation, float slope_in, int apply_tanh,
                                    vtts_stream_t stream) {
    VTTS_REQUIRE(x && w && y, "vtts_dbg_conv1d_fp32: null pointer");
    VTTS_REQUIRE(ksize % 2 == 1 && dilation >= 1, "vtts_dbg_conv1d_fp32: odd kernel, dilation>=1");
    cudaStream_t st = (cudaStream_t)stream;
    float *packed = nullptr;
    VTTS_CHECK_CUDA(cudaMallocAsync(&packed, sizeof(float) * (size_t)cin * cout * ksize, st));
    int rc = launch_pack_conv_fp32(w, packed, cout, cin, ksize, st);
    if (rc == VTTS_OK) {
        ConvFp32Params p{};
        p.x = x; p.w = packed; p.bias = bias; p.bias_b = nullptr; p.res = res; p.y = y;
        p.B = B; p.cin = cin; p.cout = cout; p.L_in = L; p.L_out = L;
        p.taps = ksize; p.tap_off0 = -(ksize - 1) / 2 * dilation; p.tap_step = dilation;
        p.phases = 1; p.out_stride = 1; p.out_off0 = 0; p.n_pos = L;
        p.slope_in = slope_in; p.accumulate = 0; p.divide_by = 0.f; p.apply_tanh = apply_tanh;
        rc = launch_conv_fp32(p, st);
    }
    cudaFreeAsync(packed, st);
    return rc;
}
