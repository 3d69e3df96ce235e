// chain_tc.cuh -- host interface of the fused ResidualBlock chain kernel (chain_tc.cu).
#pragma once
#include "generator.cuh"

namespace vtts {
namespace tc {

// One launch: x (B, L, C) fp32 channels-last -> block output, combined into the MRF sum.
struct ChainRun {
    const float *x = nullptr;
    float *cs = nullptr;        // running MRF sum (B, L, C) fp32 channels-last
    float *out_x = nullptr;     // wr == 2: fp32 result or null
    uint16_t *out_a = nullptr;  // wr == 2: 16-bit LeakyReLU'd copy (B, L, C) or null
    int rd_cs = 0;              // 1: value += cs (last of several blocks)
    int wr = 0;                 // 0: cs = value; 1: cs += value (vector reduction); 2: out_x / out_a = value * scale
    float scale = 1.f, slope = 0.1f, slope_out = 0.1f;
    int B = 0, L = 0;
    const long long *lens = nullptr;   // padding trim (see TcConvParams)
    int len_margin = 0, len_rate = 0;
    bool pdl = false;
    long long *trace = nullptr;        // debug: device buffer for block 0's pipeline stamps (chain_tc.cu)
};

bool chain_spec_usable(const ChainSpec &s);
// extra look-ahead (positions) of the chain kernel beyond the block's receptive field (zero-weight taps of the
// phase-packed MMAs touch up to PH-1 further dilated steps per conv: garbage there must not be NaN-propagated)
int chain_extra_reach(const ChainSpec &s);
int chain_pack_raw(const ChainSpec &s, const float *const *w, const float *const *bias, ChainWeights &cw, cudaStream_t st);
void chain_free(ChainWeights &cw);
int chain_launch(const ChainWeights &cw, int fmt, const ChainRun &r, cudaStream_t st);

// (B, L, C) fp32 channels-last output conv + tanh (generator.py:108-120)
int launch_conv_post_cl(const float *x, const float *w_kc, const float *bias, float *y, int B, int C, int L, int ksize,
                        float slope, int out_channels, int oc, const long long *lens, int len_margin, int len_rate,
                        bool pdl, cudaStream_t st);

}  // namespace tc
}  // namespace vtts
