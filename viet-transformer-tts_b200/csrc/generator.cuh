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
    int ci_pad = 0;          // bf16 packing: padded input channels
    int n_total = 0;         // bf16 packing: rows per tap (cout, or s*cout for polyphase)
    bool has_bias = false;
    bool loaded = false;
};

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
