// generator.cu -- VttsGen handle: layer table, weight upload and the fp32 forward pass.
//
// Mirrors HiFiGAN.forward (models/gan_tts/hifigan/generator.py:132-156) and
// ResidualBlock.forward (models/gan_tts/hifigan/layers.py:83-98); the vits2 skin
// (models/gan_tts/vits2/layers.py:159-177) is the same function with different parameter
// names and maps onto the same handle.
#include "generator.cuh"

#include <new>
#include <string.h>

namespace vtts {

static thread_local char g_err[512] = "";
char *error_buffer() { return g_err; }
int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int convT_out_len(int L, int s, int k, int p, int op) { return (L - 1) * s - 2 * p + k + op; }

static void free_layer(Layer &l) {
    cudaFree(l.w_fold); cudaFree(l.w_f32); cudaFree(l.bias); cudaFree(l.w16[0]); cudaFree(l.w16[1]); cudaFree(l.w_aux);
    l.w_fold = l.w_f32 = l.bias = l.w_aux = nullptr; l.w16[0] = l.w16[1] = nullptr;
}

}  // namespace vtts

using namespace vtts;

extern "C" int vtts_version(void) { return VTTS_VERSION; }
extern "C" const char *vtts_last_error(void) { return error_buffer(); }

extern "C" int vtts_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    VTTS_CHECK_CUDA(cudaGetDevice(&dev));
    VTTS_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    VTTS_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    return major * 10 + minor;
}

extern "C" int vtts_gen_create(const VttsGenConfig *cfg, VttsGen **out) {
    VTTS_REQUIRE(cfg && out, "vtts_gen_create: null pointer");
    const VttsGenConfig &c = *cfg;
    VTTS_REQUIRE(c.in_channels > 0 && c.out_channels > 0 && c.channels > 0, "vtts_gen_create: channels must be > 0");
    VTTS_REQUIRE(c.kernel_size % 2 == 1, "vtts_gen_create: Kernel size must be odd number. (generator.py:62)");
    VTTS_REQUIRE(c.num_upsamples >= 1 && c.num_upsamples <= VTTS_MAX_STAGES, "vtts_gen_create: 1..%d upsample stages", VTTS_MAX_STAGES);
    VTTS_REQUIRE(c.num_blocks >= 1 && c.num_blocks <= VTTS_MAX_BLOCKS, "vtts_gen_create: 1..%d resblocks per stage", VTTS_MAX_BLOCKS);
    VTTS_REQUIRE((c.channels >> c.num_upsamples) >= 1, "vtts_gen_create: channels %d too small for %d halvings", c.channels, c.num_upsamples);
    for (int i = 0; i < c.num_upsamples; ++i) {
        VTTS_REQUIRE(c.upsample_scales[i] >= 1 && c.upsample_kernel_sizes[i] >= 1, "vtts_gen_create: bad upsample %d", i);
        if (c.upsample_kernel_sizes[i] % c.upsample_scales[i] != 0)
            return set_error(VTTS_E_UNSUPPORTED, "vtts_gen_create: upsample kernel %d is not a multiple of scale %d "
                             "(polyphase lowering needs k %% s == 0; the reference asserts k == 2s, generator.py:80)",
                             c.upsample_kernel_sizes[i], c.upsample_scales[i]);
        VTTS_REQUIRE(c.upsample_paddings[i] >= 0 && c.upsample_output_paddings[i] >= 0, "vtts_gen_create: negative padding");
    }
    for (int j = 0; j < c.num_blocks; ++j) {
        VTTS_REQUIRE(c.resblock_kernel_sizes[j] % 2 == 1, "vtts_gen_create: Kernel size must be odd number. (layers.py:46)");
        VTTS_REQUIRE(c.num_dilations[j] >= 1 && c.num_dilations[j] <= VTTS_MAX_DILATIONS, "vtts_gen_create: 1..%d dilations", VTTS_MAX_DILATIONS);
        for (int m = 0; m < c.num_dilations[j]; ++m)
            VTTS_REQUIRE(c.resblock_dilations[j][m] >= 1, "vtts_gen_create: dilation must be >= 1");
    }
    VttsGen *h = new (std::nothrow) VttsGen();
    VTTS_REQUIRE(h, "vtts_gen_create: out of host memory");
    h->cfg = c;
    cudaGetDevice(&h->device);

    auto add = [&](int kind, int cin, int cout, int k, int dil, int stage, int block, int unit, int which) {
        Layer l;
        l.info.kind = kind; l.info.cin = cin; l.info.cout = cout; l.info.ksize = k; l.info.dilation = dil;
        l.info.stage = stage; l.info.block = block; l.info.unit = unit; l.info.which = which;
        h->layers.push_back(l);
        return (int)h->layers.size() - 1;
    };
    h->idx_pre = add(0, c.in_channels, c.channels, c.kernel_size, 1, -1, -1, -1, -1);
    int ch = c.channels;
    h->upsample_factor = c.out_channels;
    for (int i = 0; i < c.num_upsamples; ++i) {
        int up = add(1, ch, ch / 2, c.upsample_kernel_sizes[i], 1, i, -1, -1, -1);
        h->layers[up].stride = c.upsample_scales[i];
        h->layers[up].padding = c.upsample_paddings[i];
        h->layers[up].output_padding = c.upsample_output_paddings[i];
        h->idx_up.push_back(up);
        h->upsample_factor *= c.upsample_scales[i];
        ch /= 2;
        h->idx_c1.emplace_back(); h->idx_c2.emplace_back();
        for (int j = 0; j < c.num_blocks; ++j) {
            h->idx_c1[i].emplace_back(); h->idx_c2[i].emplace_back();
            for (int m = 0; m < c.num_dilations[j]; ++m) {
                h->idx_c1[i][j].push_back(add(0, ch, ch, c.resblock_kernel_sizes[j], c.resblock_dilations[j][m], i, j, m, 1));
                h->idx_c2[i][j].push_back(c.use_additional_convs
                    ? add(0, ch, ch, c.resblock_kernel_sizes[j], 1, i, j, m, 2) : -1);
            }
        }
    }
    h->idx_post = add(0, ch, c.out_channels, c.kernel_size, 1, c.num_upsamples, -1, -1, -1);
    if (c.global_channels > 0)
        h->idx_global = add(0, c.global_channels, c.channels, 1, 1, c.num_upsamples, -1, -1, -1);
    *out = h;
    return VTTS_OK;
}

extern "C" int vtts_gen_destroy(VttsGen *h) {
    if (!h) return VTTS_OK;
    tc_destroy(h);
    for (auto &l : h->layers) free_layer(l);
    delete h;
    return VTTS_OK;
}

extern "C" int vtts_gen_num_layers(const VttsGen *h) {
    VTTS_REQUIRE(h, "vtts_gen_num_layers: null handle");
    return (int)h->layers.size();
}

extern "C" int vtts_gen_layer_info(const VttsGen *h, int layer, VttsLayerInfo *info) {
    VTTS_REQUIRE(h && info, "vtts_gen_layer_info: null pointer");
    VTTS_REQUIRE(layer >= 0 && layer < (int)h->layers.size(), "vtts_gen_layer_info: layer %d out of range", layer);
    *info = h->layers[layer].info;
    return VTTS_OK;
}

extern "C" int vtts_gen_load_layer(VttsGen *h, int layer, const float *weight_v, const float *weight_g,
                                   const float *bias, vtts_stream_t stream) {
    VTTS_REQUIRE(h && weight_v, "vtts_gen_load_layer: null pointer");
    VTTS_REQUIRE(layer >= 0 && layer < (int)h->layers.size(), "vtts_gen_load_layer: layer %d out of range", layer);
    cudaStream_t st = (cudaStream_t)stream;
    Layer &l = h->layers[layer];
    const int cin = l.info.cin, cout = l.info.cout, k = l.info.ksize;
    const size_t n = (size_t)cin * cout * k;
    if (!l.w_fold) VTTS_CHECK_CUDA(cudaMalloc(&l.w_fold, n * sizeof(float)));
    if (!l.w_f32) VTTS_CHECK_CUDA(cudaMalloc(&l.w_f32, n * sizeof(float)));
    const int dim0 = l.info.kind == 0 ? cout : cin;  // weight_norm dim=0
    if (weight_g) {
        int rc = launch_fold_weight_norm(weight_v, weight_g, l.w_fold, dim0, (int)(n / dim0), st);
        if (rc) return rc;
    } else {
        VTTS_CHECK_CUDA(cudaMemcpyAsync(l.w_fold, weight_v, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    int rc = l.info.kind == 0 ? launch_pack_conv_fp32(l.w_fold, l.w_f32, cout, cin, k, st)
                              : launch_pack_convT_fp32(l.w_fold, l.w_f32, cin, cout, k, l.stride, st);
    if (rc) return rc;
    if (bias) {
        if (!l.bias) VTTS_CHECK_CUDA(cudaMalloc(&l.bias, cout * sizeof(float)));
        VTTS_CHECK_CUDA(cudaMemcpyAsync(l.bias, bias, cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
        l.has_bias = true;
    } else {
        l.has_bias = false;
    }
    h->chain_dirty = true;   // chain-kernel packing (all convs of a ResidualBlock together) is redone by the next forward
    rc = tc_pack_layer(h, layer, st);
    if (rc) return rc;
    l.loaded = true;
    return VTTS_OK;
}

// --------------------------------------------------------------------------------------------
// fp32 forward (channels-first)
// --------------------------------------------------------------------------------------------
namespace {

// max |x| over n floats, folded into *out (non-negative floats order like their bit patterns)
__global__ void absmax_kernel(const float *__restrict__ x, size_t n, float *out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int *>(out), __float_as_uint(m));
}

struct Plan {
    std::vector<int> C, L;  // per stage, after upsample
    size_t max_elems = 0;
};

static Plan make_plan(const VttsGen *h, int B, int T) {
    Plan p;
    int ch = h->cfg.channels, L = T;
    p.max_elems = (size_t)B * ch * L;
    for (int i = 0; i < h->cfg.num_upsamples; ++i) {
        const Layer &u = h->layers[h->idx_up[i]];
        L = convT_out_len(L, u.stride, u.info.ksize, u.padding, u.output_padding);
        ch /= 2;
        p.C.push_back(ch); p.L.push_back(L);
        size_t e = (size_t)B * ch * (size_t)(L > 0 ? L : 0);
        if (e > p.max_elems) p.max_elems = e;
    }
    return p;
}

static ConvFp32Params conv_params(const Layer &l, const float *x, float *y, int B, int L, float slope_in) {
    ConvFp32Params p{};
    p.x = x; p.w = l.w_f32; p.bias = l.has_bias ? l.bias : nullptr; p.y = y;
    p.B = B; p.cin = l.info.cin; p.cout = l.info.cout; p.L_in = L; p.L_out = L;
    p.taps = l.info.ksize; p.tap_off0 = -(l.info.ksize - 1) / 2 * l.info.dilation; p.tap_step = l.info.dilation;
    p.phases = 1; p.out_stride = 1; p.out_off0 = 0; p.n_pos = L; p.slope_in = slope_in;
    return p;
}

static int forward_fp32(VttsGen *h, const float *c, const float *g, float *wav, int B, int T, void *workspace,
                        size_t workspace_bytes, int dump_stage, float *dump_out, cudaStream_t st) {
    const VttsGenConfig &cfg = h->cfg;
    Plan plan = make_plan(h, B, T);
    const size_t buf = align_up(plan.max_elems * sizeof(float), 256);
    const size_t need = 5 * buf + align_up((size_t)B * cfg.channels * sizeof(float), 256);
    if (workspace_bytes < need) return set_error(VTTS_E_WORKSPACE, "vtts_gen_forward: workspace %zu < %zu", workspace_bytes, need);
    char *ws = (char *)workspace;
    float *U = (float *)(ws), *P = (float *)(ws + buf), *Q = (float *)(ws + 2 * buf), *Tt = (float *)(ws + 3 * buf),
          *CS = (float *)(ws + 4 * buf), *GB = (float *)(ws + 5 * buf);
    int rc;
    auto dump = [&](int stage_id, const float *src, int C, int L) -> int {
        if (dump_stage == stage_id && dump_out)
            VTTS_CHECK_CUDA(cudaMemcpyAsync(dump_out, src, (size_t)B * C * L * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return VTTS_OK;
    };

    // optional range probe (vtts_gen_set_range_probe): max |layer output| per layer
    auto probe = [&](int layer, const float *y, size_t n) {
        if (!h->range_probe || n == 0) return;
        size_t blocks = (n + 1023) / 1024;
        if (blocks > 1184) blocks = 1184;
        absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, n, h->range_probe + layer);
    };
    probe((int)h->layers.size(), c, (size_t)B * cfg.in_channels * T);
    // global conditioning: c += global_conv(g)  (generator.py:146-147) -> per-batch bias
    const float *bias_b = nullptr;
    if (g) {
        VTTS_REQUIRE(h->idx_global >= 0, "vtts_gen_forward: g given but global_channels <= 0");
        ConvFp32Params p = conv_params(h->layers[h->idx_global], g, GB, B, 1, 1.f);
        if ((rc = launch_conv_fp32(p, st))) return rc;
        h->launch_count++;
        bias_b = GB;
    }
    // input_conv (generator.py:145)
    {
        ConvFp32Params p = conv_params(h->layers[h->idx_pre], c, P, B, T, 1.f);
        p.bias_b = bias_b;
        if ((rc = launch_conv_fp32(p, st))) return rc;
        h->launch_count++;
        probe(h->idx_pre, P, (size_t)B * cfg.channels * T);
        if ((rc = dump(0, P, cfg.channels, T))) return rc;
    }
    const float *cur = P;
    int L = T;
    for (int i = 0; i < cfg.num_upsamples; ++i) {
        const Layer &u = h->layers[h->idx_up[i]];
        const int Lo = plan.L[i], C = plan.C[i];
        VTTS_REQUIRE(Lo > 0, "vtts_gen_forward: stage %d output length %d <= 0", i, Lo);
        {   // upsamples[i]: LeakyReLU + ConvTranspose1d (generator.py:149), polyphase
            ConvFp32Params p{};
            p.x = cur; p.w = u.w_f32; p.bias = u.has_bias ? u.bias : nullptr; p.y = U;
            p.B = B; p.cin = u.info.cin; p.cout = u.info.cout; p.L_in = L; p.L_out = Lo;
            p.taps = u.info.ksize / u.stride; p.tap_off0 = 0; p.tap_step = -1;
            p.phases = u.stride; p.out_stride = u.stride; p.out_off0 = -u.padding;
            p.n_pos = L + p.taps - 1; p.slope_in = cfg.lrelu_slope;
            if ((rc = launch_conv_fp32(p, st))) return rc;
            h->launch_count++;
            probe(h->idx_up[i], U, (size_t)B * C * Lo);
            if ((rc = dump(2 * i + 1, U, C, Lo))) return rc;
        }
        for (int j = 0; j < cfg.num_blocks; ++j) {
            const float *y = U;
            const int nu = cfg.num_dilations[j];
            for (int m = 0; m < nu; ++m) {
                const bool last = (m == nu - 1);
                float *ynew = last ? CS : ((m & 1) ? Q : P);
                const Layer &l1 = h->layers[h->idx_c1[i][j][m]];
                ConvFp32Params p1 = conv_params(l1, y, cfg.use_additional_convs ? Tt : ynew, B, Lo, cfg.lrelu_slope);
                ConvFp32Params *fin = &p1;
                ConvFp32Params p2{};
                if (cfg.use_additional_convs) {
                    if ((rc = launch_conv_fp32(p1, st))) return rc;
                    h->launch_count++;
                    probe(h->idx_c1[i][j][m], Tt, (size_t)B * C * Lo);
                    p2 = conv_params(h->layers[h->idx_c2[i][j][m]], Tt, ynew, B, Lo, cfg.lrelu_slope);
                    fin = &p2;
                }
                fin->res = y;  // x = xt + x (layers.py:97)
                if (last) {    // cs += block(c); c = cs / num_blocks (generator.py:150-153)
                    fin->accumulate = (j > 0);
                    fin->divide_by = (j == cfg.num_blocks - 1) ? (float)cfg.num_blocks : 0.f;
                }
                if ((rc = launch_conv_fp32(*fin, st))) return rc;
                h->launch_count++;
                // (the last unit of a block writes the running MRF sum: an upper bound of the block output's range / num_blocks)
                probe(cfg.use_additional_convs ? h->idx_c2[i][j][m] : h->idx_c1[i][j][m], ynew, (size_t)B * C * Lo);
                y = ynew;
            }
        }
        if ((rc = dump(2 * i + 2, CS, C, Lo))) return rc;
        cur = CS;
        L = Lo;
    }
    {   // output_conv: LeakyReLU(0.01) + Conv1d + Tanh (generator.py:108-120)
        ConvFp32Params p = conv_params(h->layers[h->idx_post], cur, wav, B, L, cfg.final_lrelu_slope);
        p.apply_tanh = 1;
        if ((rc = launch_conv_fp32(p, st))) return rc;
        h->launch_count++;
    }
    return VTTS_OK;
}

}  // namespace

extern "C" int vtts_gen_workspace_bytes(const VttsGen *h, int B, int T, int precision, size_t *bytes) {
    VTTS_REQUIRE(h && bytes, "vtts_gen_workspace_bytes: null pointer");
    VTTS_REQUIRE(B >= 1 && T >= 1, "vtts_gen_workspace_bytes: B and T must be >= 1");
    if (precision == VTTS_PRECISION_FP32) {
        Plan plan = make_plan(h, B, T);
        *bytes = 5 * align_up(plan.max_elems * sizeof(float), 256) + align_up((size_t)B * h->cfg.channels * sizeof(float), 256);
        return VTTS_OK;
    }
    if (precision == VTTS_PRECISION_BF16 || precision == VTTS_PRECISION_FP16) return tc_workspace_bytes(h, B, T, bytes);
    return set_error(VTTS_E_INVALID, "vtts_gen_workspace_bytes: unknown precision %d", precision);
}

extern "C" int vtts_gen_forward(VttsGen *h, const float *c, const float *g, float *wav, int B, int T, void *workspace,
                                size_t workspace_bytes, int precision, int dump_stage, float *dump_out,
                                vtts_stream_t stream) {
    VTTS_REQUIRE(h && c && wav && workspace, "vtts_gen_forward: null pointer");
    VTTS_REQUIRE(B >= 1 && T >= 1, "vtts_gen_forward: B and T must be >= 1");
    for (size_t i = 0; i < h->layers.size(); ++i)
        if (!h->layers[i].loaded) return set_error(VTTS_E_STATE, "vtts_gen_forward: layer %zu has no weights", i);
    h->launch_count = 0;
    if (precision == VTTS_PRECISION_FP32)
        return forward_fp32(h, c, g, wav, B, T, workspace, workspace_bytes, dump_stage, dump_out, (cudaStream_t)stream);
    if (precision == VTTS_PRECISION_BF16 || precision == VTTS_PRECISION_FP16)
        return tc_forward(h, precision == VTTS_PRECISION_BF16 ? VTTS_FMT_BF16 : VTTS_FMT_FP16, c, g, wav, B, T, workspace,
                          workspace_bytes, dump_stage, dump_out, (cudaStream_t)stream);
    return set_error(VTTS_E_INVALID, "vtts_gen_forward: unknown precision %d", precision);
}

extern "C" int vtts_gen_set_valid_lengths(VttsGen *h, const int64_t *mel_len, int margin_frames) {
    VTTS_REQUIRE(h, "vtts_gen_set_valid_lengths: null handle");
    if (margin_frames < 0) margin_frames = 0;   // auto: the per-layer margins derived from the receptive fields
    h->trim_lens = mel_len;
    h->trim_margin = margin_frames;
    return VTTS_OK;
}

extern "C" int vtts_gen_set_range_probe(VttsGen *h, float *absmax) {
    VTTS_REQUIRE(h, "vtts_gen_set_range_probe: null handle");
    h->range_probe = absmax;
    return VTTS_OK;
}

extern "C" int vtts_gen_last_launch_count(const VttsGen *h) {
    VTTS_REQUIRE(h, "vtts_gen_last_launch_count: null handle");
    return h->launch_count;
}
