/* vtts_b200.h -- C ABI of libvtts_b200.so (B200 / sm_100a synthesis hot path).
 *
 * The reference (ducnt18121997/Viet-Transformer-TTS) is pure Python/PyTorch and has no FFI
 * layer; its boundary for this path is the nn.Module surface.  Each entry point below names
 * the reference interface it replaces (paths relative to the reference root).  The Python
 * module shells in viet-transformer-tts_b200/vtts_b200/ bind these with ctypes; the stub a
 * reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns 0 on success, a negative VTTS_E_* code on failure;
 *    vtts_last_error() returns a thread-local message for the last failure on this thread;
 *  - nothing throws, nothing allocates activation memory: inputs, outputs and workspaces are
 *    raw device pointers owned by the caller (PyTorch); handles own only packed weights;
 *  - every launch goes to the caller's stream (pass torch.cuda.current_stream().cuda_stream);
 *    no hidden synchronisation;
 *  - one VttsGen handle per (device, module); a handle is not thread-safe, distinct handles are.
 */
#ifndef VTTS_B200_H_
#define VTTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTTS_VERSION 100 /* 0.1.0 */

#define VTTS_OK 0
#define VTTS_E_INVALID (-1)     /* bad argument / unsupported configuration */
#define VTTS_E_CUDA (-2)        /* CUDA runtime / driver error */
#define VTTS_E_WORKSPACE (-3)   /* workspace too small */
#define VTTS_E_STATE (-4)       /* handle not ready (weights missing) */
#define VTTS_E_UNSUPPORTED (-5) /* valid configuration this build has no kernel for */

typedef void *vtts_stream_t; /* cudaStream_t */

int vtts_version(void);
const char *vtts_last_error(void);
/* Compute capability major*10+minor of the current device, or a negative code. */
int vtts_device_arch(void);

/* ------------------------------------------------------------------------------------------
 * LengthRegulator -- replaces models/tts/fastspeech2/layers.py:434-462
 * (LengthRegulator.forward) and the pad_list it calls (fastspeech2/function.py:97-124).
 * ---------------------------------------------------------------------------------------- */

/* layers.py:446-448  ds = torch.round(ds.float() * alpha).long()   (n = B*Tmax elements). */
int vtts_lr_scale_durations(const int64_t *ds, int64_t n, float alpha, int64_t *out,
                            vtts_stream_t stream);

/* Caller-side mel_lens = torch.sum(ds, dim=1) (layers.py:209) plus the three scalars the
 * host needs for pad_list's max_len and the ds.sum()==0 test (layers.py:450):
 *   stats[0] = max_b mel_len[b], stats[1] = sum_b mel_len[b], stats[2] = #negative durations.
 * mel_len may be NULL.  stats (device, 3 x int64) is zeroed by this call. */
int vtts_lr_rowsum(const int64_t *ds, int B, int Tmax, int64_t *mel_len, int64_t *stats,
                   vtts_stream_t stream);

/* layers.py:458  ds[ds.sum(dim=1).eq(0)] = 1   (in place on the caller's tensor). */
int vtts_lr_fix_zero_rows(int64_t *ds, int B, int Tmax, vtts_stream_t stream);

/* layers.py:460-462  repeat_interleave each row, pad_list to T_out.
 * xs (B,Tmax,D) and out (B,T_out,D) are contiguous arrays of elem_size-byte elements
 * (1,2,4 or 8); pad points at ONE element's bytes on the host.  Bit-exact copy. */
int vtts_lr_gather(const void *xs, const int64_t *ds, void *out, int B, int Tmax, int D,
                   int64_t T_out, int elem_size, const void *pad, vtts_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GaussianUpsampling -- replaces models/tts/fastspeech2/layers.py:476-520 (GaussianUpsampling.forward;
 * same body in models/gan_tts/jets/alignments.py:168-222).  hs (B,T_text,D) fp32, ds (B,T_text) int64,
 * h_mask (B,T_feats) / d_mask (B,T_text) torch.bool bytes or NULL, out (B,T_feats,D) fp32.
 * The all-zero-batch fix-up (layers.py:492-499) and T_feats (layers.py:501-504) are the host's job
 * (vtts_lr_rowsum / vtts_lr_fix_zero_rows).
 * ---------------------------------------------------------------------------------------- */
int vtts_gauss_upsample(const float *hs, const int64_t *ds, const unsigned char *h_mask,
                        const unsigned char *d_mask, float *out, int B, int T_text, int D, int T_feats,
                        float delta, vtts_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * vits2 monotonic duration path -- replaces models/gan_tts/vits2/utils.py:111-126 (generate_path) and the
 * attn matmuls at models/gan_tts/vits2/generator.py:256-259.
 * duration (B,t_x) fp32 (the [b,1,t_x] tensor, integer valued at the call site: w_ceil), mask (B,t_y,t_x) fp32 or NULL
 * (the [b,1,t_y,t_x] attn_mask), path (B,t_y,t_x) fp32.
 * vtts_path_expand computes matmul(path, x^T)^T without materialising the path: x (B,D,t_x) -> out (B,D,t_y).
 * ---------------------------------------------------------------------------------------- */
int vtts_path_generate(const float *duration, const float *mask, float *path, int B, int t_y, int t_x,
                       vtts_stream_t stream);
int vtts_path_expand(const float *x, const float *duration, const float *mask, float *out, int B, int D, int t_y,
                     int t_x, vtts_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * HiFi-GAN generator -- replaces models/gan_tts/hifigan/generator.py:132-156
 * (HiFiGAN.forward), layers.py:83-98 (ResidualBlock.forward) and the vits2 skin
 * models/gan_tts/vits2/layers.py:159-177 (Generator.forward), sublayers.py:293-303,341-349.
 * ---------------------------------------------------------------------------------------- */

#define VTTS_MAX_STAGES 8
#define VTTS_MAX_BLOCKS 8
#define VTTS_MAX_DILATIONS 8

typedef struct VttsGenConfig {
    int32_t in_channels;      /* generator.py:21  (80; 384 for JETS; 192 for vits2) */
    int32_t out_channels;     /* generator.py:22  (1) */
    int32_t channels;         /* generator.py:23  (512) */
    int32_t global_channels;  /* generator.py:24  (-1/0 = none) */
    int32_t kernel_size;      /* generator.py:25  input/output conv kernel (7) */
    int32_t num_upsamples;    /* len(upsample_scales) */
    int32_t upsample_scales[VTTS_MAX_STAGES];
    int32_t upsample_kernel_sizes[VTTS_MAX_STAGES];
    int32_t upsample_paddings[VTTS_MAX_STAGES];        /* ESPnet: s/2+s%2; vits2: (k-u)/2 */
    int32_t upsample_output_paddings[VTTS_MAX_STAGES]; /* ESPnet: s%2;     vits2: 0 */
    int32_t num_blocks;       /* len(resblock_kernel_sizes) */
    int32_t resblock_kernel_sizes[VTTS_MAX_BLOCKS];
    int32_t num_dilations[VTTS_MAX_BLOCKS];
    int32_t resblock_dilations[VTTS_MAX_BLOCKS][VTTS_MAX_DILATIONS];
    int32_t use_additional_convs; /* 1: ResidualBlock/ResBlock1 (conv pairs); 0: ResBlock2 */
    float lrelu_slope;            /* 0.1 */
    float final_lrelu_slope;      /* 0.01 (nn.LeakyReLU() default, generator.py:111) */
} VttsGenConfig;

typedef struct VttsGen VttsGen;

/* Layer enumeration.  Layers are numbered in reference construction order:
 *   0                      input_conv            (generator.py:70)
 *   per stage i:           upsamples[i][1]       (generator.py:86)  kind 1
 *     per block j, unit m: blocks[n].convs1[m][1] then convs2[m][1] (layers.py:48-81)
 *   then                   output_conv[1]        (generator.py:113)
 *   then (if any)          global_conv           (generator.py:123) */
typedef struct VttsLayerInfo {
    int32_t kind;        /* 0 Conv1d, 1 ConvTranspose1d */
    int32_t cin, cout, ksize, dilation;
    int32_t stage;       /* -1 pre, num_upsamples post/global */
    int32_t block, unit, which; /* which: 1 = convs1, 2 = convs2; -1 n/a */
} VttsLayerInfo;

int vtts_gen_create(const VttsGenConfig *cfg, VttsGen **out);
int vtts_gen_destroy(VttsGen *h);
int vtts_gen_num_layers(const VttsGen *h);
int vtts_gen_layer_info(const VttsGen *h, int layer, VttsLayerInfo *info);

/* Upload one layer.  `weight_v` is the layer's weight in the reference's own layout
 * (Conv1d (cout,cin,k); ConvTranspose1d (cin,cout,k)), fp32, device memory.  If `weight_g`
 * is non-NULL the weight-norm is folded on the device, w = g * v / ||v||_(1,2) per dim-0
 * index (torch.nn.utils.weight_norm, generator.py:192); NULL means `weight_v` is the plain
 * weight (after remove_weight_norm, generator.py:173-183).  bias may be NULL (vits2 conv_post).
 * Packs the fp32 and the bf16 tensor-core copies. */
int vtts_gen_load_layer(VttsGen *h, int layer, const float *weight_v, const float *weight_g,
                        const float *bias, vtts_stream_t stream);

#define VTTS_PRECISION_FP32 0 /* CUDA-core fp32 direct convolution (in-repo reference path) */
#define VTTS_PRECISION_BF16 1 /* tcgen05, bf16 operands, fp32 accumulate/residual, fp32 output conv */
#define VTTS_PRECISION_FP16 2 /* same kernels and rate with fp16 operands (8x smaller rounding error) */

int vtts_gen_workspace_bytes(const VttsGen *h, int B, int T, int precision, size_t *bytes);

/* c (B,in_channels,T) fp32 channels-first, g (B,global_channels) fp32 or NULL,
 * wav (B,out_channels,T*upsample_factor) fp32.  dump_stage >= 0 additionally copies an
 * intermediate tensor to dump_out as (B,C,L) fp32 channels-first:
 *   0 = input_conv output, 2i+1 = upsamples[i] output, 2i+2 = MRF mean of stage i. */
int vtts_gen_forward(VttsGen *h, const float *c, const float *g, float *wav, int B, int T,
                     void *workspace, size_t workspace_bytes, int precision, int dump_stage,
                     float *dump_out, vtts_stream_t stream);

/* Extension (not in the reference): padding trim for batched synthesis.  mel_len (device, B x int64, or
 * NULL to disable) gives each row's valid mel frames; until reset, every layer of the 16-bit path of
 * vtts_gen_forward skips the tiles that cannot reach the first mel_len[b] * upsample_factor samples of row b:
 * it computes mel_len[b] frames plus the look-ahead its successors need, derived from the handle's kernel
 * sizes, dilations and scales.  margin_frames adds extra frames on top (a lower bound per layer, never a cap;
 * pass 0 or a negative value for the derived margins alone).  The valid samples are bit-identical to the
 * untrimmed call; samples beyond mel_len[b] * upsample_factor are zero or undefined.  The pointer must stay
 * valid until the forward has run. */
int vtts_gen_set_valid_lengths(VttsGen *h, const int64_t *mel_len, int margin_frames);

/* Diagnostic for the 16-bit paths (not in the reference): activation range probe.  absmax (device, vtts_gen_num_layers + 1
 * floats zeroed by the caller, or NULL to disable): until reset, the FP32 path of vtts_gen_forward records
 * absmax[l] = max |output of layer l| (after the residual add, i.e. the tensor whose LeakyReLU the next conv consumes) and
 * absmax[num_layers] = max |c|.  These are exactly the tensors the 16-bit paths round to fp16 / bf16 operands
 * (cvt.rn.satfinite clips silently at 65504 for fp16): a checkpoint whose probe stays below 65504 cannot saturate. */
int vtts_gen_set_range_probe(VttsGen *h, float *absmax);

/* Number of kernel launches the last vtts_gen_forward on this handle issued. */
int vtts_gen_last_launch_count(const VttsGen *h);

/* ------------------------------------------------------------------------------------------
 * Test hooks (used by tests/ only): single-layer entry points.
 * ---------------------------------------------------------------------------------------- */

/* Plain fp32 Conv1d on channels-first tensors with the fused options the generator uses.
 * y = [tanh]( (conv(lrelu_in(x)) + bias [+ res]) ) ; slope_in == 1 disables the activation. */
int vtts_dbg_conv1d_fp32(const float *x, const float *w, const float *bias, const float *res,
                         float *y, int B, int cin, int cout, int L, int ksize, int dilation,
                         float slope_in, int apply_tanh, vtts_stream_t stream);

/* tcgen05 self-test: D[M,N] = A[M,K] * B[N,K]^T with bf16 operands (K-major, TMA SW128),
 * optional row shift of the A descriptor (exercises the halo-reuse addressing used by the
 * convolution kernels).  variant selects descriptor construction options. */
int vtts_dbg_umma_gemm(const void *a_bf16, const void *b_bf16, float *d, int M, int N, int K,
                       int a_rows_total, int row_shift, int variant, vtts_stream_t stream);

/* One Conv1d layer through the tcgen05 kernel (bf16 or fp16 operands, fp32 accumulate), channels-first
 * fp32 in/out: y = conv(bf16(lrelu_in(x))) + bias [+ res]; y_act (optional) = bf16(lrelu_out(y)). */
int vtts_dbg_conv1d_tc(const float *x, const float *w, const float *bias, const float *res, float *y,
                       float *y_act, int B, int cin, int cout, int L, int ksize, int dilation,
                       float slope_in, float slope_out, int fp16, vtts_stream_t stream);

/* Debug: enable per-tile clock64() tracing in vtts_dbg_conv1d_tc (block 0) and/or read the trace
 * buffer (16 stamps per tile, first 64 tiles) back to host memory. */
int vtts_dbg_trace(int enable, long long *host_out, int n);
/* Debug: raw tcgen05.mma issue/completion cycles (host_out[0] = issue loop, [1] = until complete). */
int vtts_dbg_umma_bench(int N, int rowb, int row_shift, int reps, int M, int two_acc, long long *host_out);

/* One ResidualBlock (layers.py:83-98; n_units x [conv1(dil[u]) -> conv2(1)] with identity skips, or conv1 only when
 * has2 == 0) through the fused chain kernel, channels-first fp32 in / out.  w / bias: HOST arrays of n_units * (1 + has2)
 * device pointers in reference order (convs1[0], convs2[0], convs1[1], ...), weights (C, C, k).  reps > 0 additionally
 * times `reps` launches with CUDA events (*ms_out = average milliseconds per launch). */
int vtts_dbg_resblock_chain(const float *x, const void *const *w, const void *const *bias, float *y, int B, int C,
                            int L, int k, const int *dil, int n_units, int has2, float slope, int fp16, int reps,
                            float *ms_out, vtts_stream_t stream);

/* ---- acoustic decoder convolutions (the layer between the LengthRegulator and the generator) -----------------
 * One Conv1d layer ("same" zero padding, odd kernel, stride 1) on the tcgen05 kernel with CHANNELS-LAST activations, the
 * layout the transformer blocks already use: replaces nn.Conv1d in PositionwiseFeedForward.w_1 / w_2
 * (models/tts/fastspeech2/blocks/transformer.py:273-286, called :290-292 between two transposes) and ConvNorm.conv + eval-mode
 * BatchNorm1d of the Postnet (models/tts/fastspeech2/layers.py:579-612, forward :614-621; the caller folds the BatchNorm
 * into weight / bias).  weight (cout, cin, k) and bias (cout, or NULL) are fp32 device arrays, packed once per load. */
typedef struct VttsConv VttsConv;
int vtts_conv_create(int cin, int cout, int ksize, int dilation, VttsConv **out);
void vtts_conv_destroy(VttsConv *c);
/* channels of the 16-bit input operand rows (cin rounded up to the kernel's K chunk: 64, or 32 for cin == 32) */
int vtts_conv_padded_channels(const VttsConv *c);
int vtts_conv_load(VttsConv *c, const float *weight, const float *bias, vtts_stream_t stream);
/* act16: (B, L, padded_channels) 16-bit operand (bf16 / fp16 per `precision`), zero in the padding channels.
 * value = conv(act) + bias (+ res[b, t, co] when res != NULL, fp32 (B, L, cout));
 * out_x (B, L, cout) fp32 or NULL receives value; out_a16 (B, L, cout) 16-bit or NULL receives
 * tanh(value) when act_tanh != 0, else LeakyReLU(value, slope_out) (slope 1 = identity, 0 = ReLU). */
int vtts_conv_forward(VttsConv *c, const void *act16, int precision, int B, int L, const float *res, float *out_x,
                      void *out_a16, float slope_out, int act_tanh, vtts_stream_t stream);

/* Conformer convolution module, the part between its two pointwise convs (models/tts/fastspeech2/blocks/conformer.py:470-480):
 * GLU over the channel halves of pw (B, L, 2C) fp32 channels-last (blocks/utils.py:75-86), depthwise Conv1d(C, C, ksize,
 * groups = C, "same" zero padding, :532-570) with weights w (C, ksize) and the eval-mode BatchNorm1d folded by the caller into
 * w / bias (C, or NULL), Swish (blocks/utils.py:63-72); writes the 16-bit operand (B, L, C) of the second pointwise conv. */
int vtts_dwconv_glu_swish(const float *pw, const float *w, const float *bias, void *out16, int precision, int B, int L, int C,
                          int ksize, vtts_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VTTS_B200_H_ */
