// generator.cuh -- VttsGen handle (packed weights) shared by the fp32 and tcgen05 paths.
#pragma once
#include "common.cuh"
#include <vector>

namespace vtts {

struct Layer {
    VttsLayerInfo info{};
    int stride = 1;          // ConvTranspose1d stride (upsample scale)
    int padding = 0;         // ConvTranspose1d padding
    int output_padding = 0;
    float *w_fold = nullptr; // folded fp32 weight, reference layout
    float *w_f32 = nullptr;  // fp32 packed [phase][ci][tap][co]
    float *bias = nullptr;   // fp32 (cout) or null
    uint16_t *w16[2] = {nullptr, nullptr};  // tcgen05 packing [tap][n_pad][ci_pad], [0] bf16 / [1] fp16 (conv_tc.cu)
    float *w_aux = nullptr;  // output conv only: fp32 [oc][k][ci] for the channels-last kernel
    std::vector<float> w_aux_host;  // the same on the host (passed to conv_post_cl_kernel as kernel parameters)
    int ci_pad = 0;          // bf16 packing: padded input channels
    int n_total = 0;         // bf16 packing: rows per tap (cout, or s*cout for polyphase)
    int qperm = 0;           // polyphase rows ordered ((q / 4) * cout + co) * 4 + q % 4: four consecutive output samples sit in four neighbouring lanes
    bool has_bias = false;
    bool loaded = false;
};

namespace tc {
constexpr int CH_MAX_CONVS = 6;     // 3 units x (conv1, conv2)
// Shape of one ResidualBlock (models/gan_tts/hifigan/layers.py:16-98; vits2 ResBlock1/2 sublayers.py:215-354)
struct ChainSpec {
    int C = 0;              // channels (32 or 64)
    int k = 0;              // kernel size of every conv of the block
    int n_units = 0;        // len(dilations), <= 3
    int has2 = 1;           // use_additional_convs: unit = conv1(d) -> conv2(1); else unit = conv1(d)
    int dil[3] = {1, 1, 1};
};
// Packed weights of one block for the fused chain kernel (chain_tc.cu), owned by the handle.
struct ChainWeights {
    uint16_t *w16[2] = {nullptr, nullptr};   // [blocks_total][C][C], tap-reversed with zero pad blocks; [0] bf16 / [1] fp16
    float *bias = nullptr;                   // [n_convs][C] effective biases (cumulative over the residual stream)
    ChainSpec spec;
    int n_convs = 0, blocks_total = 0;
    bool valid = false;
};
}  // namespace tc

struct StageDims {
    int C;      // channels after the upsample of this stage
    int scale;  // upsample factor
};

}  // namespace vtts

struct VttsGen {
    VttsGenConfig cfg{};
    std::vector<vtts::Layer> layers;
    int idx_pre = 0, idx_post = 0, idx_global = -1;
    std::vector<int> idx_up;                              // per stage
    std::vector<std::vector<std::vector<int>>> idx_c1;    // [stage][block][unit]
    std::vector<std::vector<std::vector<int>>> idx_c2;    // [stage][block][unit] (-1 if none)
    int upsample_factor = 1;
    int launch_count = 0;
    int device = 0;
    void *tc_state = nullptr;  // tensor-map cache etc. owned by conv_tc.cu
    const int64_t *trim_lens = nullptr;  // device (B) valid mel frames for the NEXT forward (vtts_gen_set_valid_lengths)
    int trim_margin = 0;
    std::vector<vtts::tc::ChainWeights> chain;   // [stage * num_blocks + block], packed lazily by tc_forward
    float *range_probe = nullptr;                // device (num_layers + 1) floats or null (vtts_gen_set_range_probe)
    bool chain_dirty = true;                     // a layer was (re)loaded since the last packing
};

namespace vtts {
// tcgen05 path (conv_tc.cu)
int tc_pack_layer(VttsGen *h, int layer, cudaStream_t stream);
int tc_workspace_bytes(const VttsGen *h, int B, int T, size_t *bytes);
int tc_forward(VttsGen *h, int fmt, const float *c, const float *g, float *wav, int B, int T,
               void *workspace, size_t workspace_bytes, int dump_stage, float *dump_out,
               cudaStream_t stream);
void tc_destroy(VttsGen *h);
int tc_supported(const VttsGen *h, char *why, size_t why_len);
}  // namespace vtts
