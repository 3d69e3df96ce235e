// conv_tc.cu -- tcgen05 / TMEM / TMA implicit-GEMM convolutions for the HiFi-GAN generator.
//
// Every convolution of the generator (generator.py:132-156, layers.py:93-97) is lowered to
//     D[n, i] = sum_j sum_ci  Wp[j][n][ci] * X[i + off_j][ci]          (16-bit x 16-bit -> fp32)
// with n = output row (out channel, or (phase, out channel) for the polyphase transposed conv,
// SURVEY.md appendix 9.1), i = time position, off_j = tap_off0 + j * tap_step.
//
// Data layout in HBM: the 16-bit operand copy (fp16 by default, bf16 optional) is channels-last (B, L, C) and
// already has the next layer's LeakyReLU applied; the fp32 residual stream is time-packed (B, L/4, C, 4).
// Weights are packed [tap][n_pad][ci_pad] (K-major rows).
//
// Three persistent, warp-specialised kernels share the building blocks below (TMA producers, tcgen05.mma issuers,
// tcgen05.ld epilogues, mbarrier pipelines):
//   conv_tc_kernel     one conv per launch: 128 output rows x 256 positions per tile, double-buffered accumulator;
//                      the activation tile is loaded once per 64-channel chunk with its halo and every tap reads it
//                      through a row-shifted shared-memory descriptor (input conv, upsamples, 256-channel stage)
//   unit_tc_kernel     conv1 -> LeakyReLU -> conv2 -> + x of a ResidualBlock unit in one launch (128 channels): the
//                      intermediate never leaves the SM (accumulator -> stmatrix -> swizzled operand tile)
//   unit64_tc_kernel   the same unit for 32 / 64 channels: M = 64 MMAs, weights resident in shared memory, four
//                      accumulators, two MMA-issuing warps
// plus conv_post_tp4_kernel (fp32 output conv + tanh) and the weight packers.  DESIGN.md section 4 has the reasoning
// and the measurements behind each choice.
#include "generator.cuh"
#include "chain_tc.cuh"
#include "tc_common.cuh"

#include <map>
#include <string>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace vtts {
namespace tc {

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    });
    return fn;
}

int make_tmap_bf16(CUtensorMap *out, const void *base, int rank, const uint64_t *dims,
                   const uint64_t *strides_bytes, const uint32_t *box, int swizzle_bytes) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return set_error(VTTS_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];  // strides of dims 1..rank-1
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                                 : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr,
                     bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(VTTS_E_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu box %u,%u,%u",
                         (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                         (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0,
                         rank > 2 ? box[2] : 0);
    return VTTS_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel (persistent: one CTA per SM loops over tiles; TMEM accumulator double-buffered so the
// epilogue of tile t overlaps the MMAs of tile t+1)
// ---------------------------------------------------------------------------------------------
constexpr int TN = 256;            // time positions per tile (UMMA N, TMEM columns per accumulator)
constexpr int TM = 128;            // output rows per tile (UMMA M, TMEM lanes)
constexpr int HALO_MAX = 64;       // (k-1)*d <= 64
constexpr int ACT_ROWS = TN + HALO_MAX;
constexpr int BOX_ROWS = 16;       // activation TMA box height (small boxes: the tile is fetched to the nearest 16 rows)
constexpr int ACC_STAGES = 2;      // TMEM accumulators
constexpr int W_PRODUCERS = 1;     // weight-producer warps (each owns the stages == its index mod 4)
constexpr int PRODUCER_WARPS = 2 + W_PRODUCERS;  // activation producer, weight producers, MMA issuer
constexpr int EPI_WARPS = 16;      // 4 per TMEM lane quarter
constexpr int MAX_TRIM_BATCH = 256; // padding trim keeps per-batch limits in shared memory

static int g_trace_on = 0;   // vtts_dbg_trace: 1 = debug conv entry, 100 + n = n-th launch of the next forward
// debug trace (vtts_dbg_trace_*): 16 stamps per tile for the first TRACE_TILES tiles of block 0
constexpr int TRACE_TILES = 64;
__device__ long long g_trace[TRACE_TILES * 16];
#define VTTS_TRACE(slot) do { if ((p.trace & 1) && blockIdx.x == 0 && tl < TRACE_TILES) g_trace[tl * 16 + (slot)] = clock64(); } while (0)
constexpr int TC_THREADS = (PRODUCER_WARPS + EPI_WARPS) * 32;

struct TcConvParams {
    // epilogue tensors (channels-last)
    const float *bias;      // (cout) or null
    const float *bias_b;    // (B, cout) or null -- per-batch global-conditioning bias (phases == 1 only)
    const float *res;       // (B, L_out, cout) fp32 or null
    float *out_x;           // (B, L_out, cout) fp32 or null
    uint16_t *out_a;        // (B, L_out, out_a_ld) bf16/fp16 or null
    int out_a_ld;
    float slope_out;        // LeakyReLU slope applied to the bf16 copy (1 = identity)
    int accumulate;         // out_x = out_x_old + value
    float divide_by;        // > 0: value /= divide_by (applied as * inv_div on this path)
    float inv_div;
    // geometry
    int n_total;            // valid output rows (cout * phases)
    int cout;               // channels per phase
    int L_out;
    int n_pos;              // time positions (GEMM N extent)
    int out_stride, out_off0;  // t_out = i * out_stride + out_off0 + phase
    int chunks;             // K chunks (ci_pad / chunk channels)
    int taps, tap_off0, tap_step;
    int act_stages, w_stages;  // pipeline depths (shared memory is carved at run time)
    int tps;                   // taps per pipeline step (2 for single-chunk layers: halves barrier ops per MMA)
    int w_rows;                // weight rows actually loaded per tile (<= 128; the rest of the A tile is don't-care)
    int epi_quarters;          // TMEM lane quarters holding real output rows in EVERY tile (1..4)
    int rep;                   // weight rows replicated `rep` times across the 128 lanes (narrow layers: 128 / n_total)
    int L4;                    // ceil(L_out / 4): fp32 streams are stored time-packed [b][t/4][c][4]
    int trace;                 // debug: block 0 records per-tile clock64() stamps into g_trace
    int stream_hint;           // 1: activation / residual reads carry the L2 evict-first policy
    int x_cl;                  // out_x is channels-last fp32 [b][t][c] (input of the chain kernel) instead of time-packed
    int act_tanh;              // the 16-bit copy is tanh(value) instead of LeakyReLU(value) (Postnet, layers.py:615-617)
    int qperm;                 // polyphase rows in quad order (see row_to_phase): 128-bit stores of the time-packed stream
    int reverse;               // walk the tiles last-to-first (alternates per launch: the tail the previous kernel just
                               // wrote is still in L2 when this kernel starts reading there)
    // optional padding trim: tiles whose first position is >= (lens[b] + len_margin) * len_rate + len_extra
    // are skipped by every role (their outputs are never needed for the valid part of utterance b)
    const long long *lens;
    int len_margin, len_rate, len_extra, batch;
    // tile schedule: work item -> (m block fastest, then time-tile group, then batch); a cluster of
    // `cluster` CTAs takes one work item: same m block (weights multicast), consecutive time tiles
    int m_blocks, t_tiles, total_tiles;   // total_tiles = work items
    int cluster, groups_per_batch;
};

// ---- epilogue helpers -------------------------------------------------------------------------------
// fp32 streams (residual x, MRF sum) use a time-packed layout [b][t/4][c][4]: a thread (one channel)
// owns 4 consecutive time steps = 16 bytes, consecutive lanes (channels) are 16 bytes apart, so a
// warp's 128-bit access covers 512 contiguous bytes.  The 16-bit operand copy stays [b][t][c]
// (K-major rows for TMA).
// max(v, v*slope) == LeakyReLU for 0 <= slope <= 1 (checked on the host); slope 1 = identity
__device__ __forceinline__ float lrelu_max(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ long long tp4_off(long long b, long long L4, long long t, int C, int c) {
    return ((b * L4 + (t >> 2)) * C + c) * 4 + (t & 3);
}

// Epilogue modes (kernel-uniform):            res  acc  div   X    A
enum { EPI_A = 0,      //                       -    -    -    -    x   conv1 of a pair, input conv
       EPI_RXA = 1,    //                       x    -    -    x    x   conv2 (x = xt + x), layers.py:97
       EPI_RX = 2,     //                       x    -    -    x    -   last unit, first block (cs = ...)
       EPI_RCX = 3,    //                       x    x    -    x    -   last unit, middle blocks (cs += ...)
       EPI_RCDXA = 4,  //                       x    x    x    x    x   last unit, last block (c = cs / n) -> next upsample
       EPI_RCDX = 5,   //                       x    x    x    x    -   same, last stage (fp32 output conv follows)
       EPI_GENERIC = 6 };
template <int MODE> struct EpiFlags {
    static constexpr bool RES = MODE >= EPI_RXA && MODE <= EPI_RCDX;
    static constexpr bool ACC = MODE == EPI_RCX || MODE == EPI_RCDXA || MODE == EPI_RCDX;
    static constexpr bool DIV = MODE == EPI_RCDXA || MODE == EPI_RCDX;
    static constexpr bool X = MODE != EPI_A;
    static constexpr bool A = MODE == EPI_A || MODE == EPI_RXA || MODE == EPI_RCDXA;
};

// residual (and running-sum) values of one 16-column group: 4 x 128-bit loads per stream
struct EpiLoads { float4 r[4]; };
template <int C_CT>
__device__ __forceinline__ void epi_load16(EpiLoads &d, const float *base, int C_rt, uint64_t policy) {
    const int C = C_CT ? C_CT : C_rt;
#pragma unroll
    for (int m = 0; m < 4; ++m) d.r[m] = ldg_f4_hint(base + (size_t)m * C * 4, policy);
}

// One fully valid 16-column group of a unit-stride layer.  C_CT > 0: compile-time channel count.
template <int FMT, int C_CT, int MODE>
__device__ __forceinline__ void epi_group16(const uint32_t (&v)[16], float bias, const TcConvParams &p,
                                            const EpiLoads &res, float *px /* tp4 base */, uint16_t *pa /* [t][c] base */) {
    using F = EpiFlags<MODE>;
    const int C = C_CT ? C_CT : p.cout;
    const int lda = C_CT ? C_CT : p.out_a_ld;
    // the running sum is read and rewritten in place: issue all four loads before the first store, otherwise every
    // load waits behind the previous stores (possible aliasing) and exposes its full latency
    float4 accv[4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
        accv[m] = F::ACC ? *reinterpret_cast<const float4 *>(px + (size_t)m * C * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        float val[4];
        const float rr[4] = {res.r[m].x, res.r[m].y, res.r[m].z, res.r[m].w};
        const float aa[4] = {accv[m].x, accv[m].y, accv[m].z, accv[m].w};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            float x = __uint_as_float(v[m * 4 + d]) + bias;
            if (F::RES) x = x + rr[d];
            if (F::ACC) x = aa[d] + x;
            if (F::DIV) x = x * p.inv_div;
            val[d] = x;
        }
        if (F::X) *reinterpret_cast<float4 *>(px + (size_t)m * C * 4) = make_float4(val[0], val[1], val[2], val[3]);
        if (F::A) {
#pragma unroll
            for (int d = 0; d < 4; ++d) pa[(size_t)(m * 4 + d) * lda] = cvt16(lrelu_max(val[d], p.slope_out), FMT);
        }
    }
}

// Generic path: per-element validity and addressing (tile edges, polyphase upsample outputs, padded rows).
template <int FMT>
__device__ __forceinline__ void epi_group16_edge(const uint32_t (&v)[16], float bias, const TcConvParams &p, bool row_ok,
                                                 int b, int ibase, int phase, int co) {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const long long t = (long long)(ibase + e) * p.out_stride + p.out_off0 + phase;
        if (row_ok && (ibase + e) < p.n_pos && t >= 0 && t < p.L_out) {
            const long long xo = p.x_cl ? ((long long)b * p.L_out + t) * p.cout + co : tp4_off(b, p.L4, t, p.cout, co);
            float val = __uint_as_float(v[e]) + bias;
            if (p.res) val = val + __ldg(p.res + xo);
            if (p.accumulate) val = p.out_x[xo] + val;
            if (p.divide_by > 0.f) val = val * p.inv_div;
            if (p.out_x) p.out_x[xo] = val;
            if (p.out_a)
                p.out_a[((long long)b * p.L_out + t) * p.out_a_ld + co] = cvt16(p.act_tanh ? tanhf(val) : lrelu_max(val, p.slope_out), FMT);
        }
    }
}

// Fully valid 16-column group of a polyphase upsample (out_stride > 1; writes fp32 x and the 16-bit copy, no
// residual): cheap 32-bit index arithmetic relative to per-batch base pointers, no per-element bounds checks.
// XCL / HAS_A are compile-time so that the 16 stores are straight-line code (one IMAD.WIDE + STG per element): with the
// flags tested per element the x2 upsamples were ISSUE-bound in this loop (trace: 4.2 k cycles of epilogue per 256-position
// tile against 1 k cycles of MMAs).
template <int FMT, bool XCL, bool HAS_A>
__device__ __forceinline__ void epi_group16_poly_t(const uint32_t (&v)[16], float bias, const TcConvParams &p, float *px_b,
                                                   uint16_t *pa_b, int t_first) {
    const int C4 = p.cout * 4;
    if (XCL) {
        float *px = px_b + (size_t)t_first * p.cout;              // channels-last: the warp's 32 channels are 128 contiguous bytes
        const int step = p.out_stride * p.cout;
#pragma unroll
        for (int e = 0; e < 16; ++e) px[e * step] = __uint_as_float(v[e]) + bias;
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int t = t_first + e * p.out_stride;
            px_b[(t >> 2) * C4 + (t & 3)] = __uint_as_float(v[e]) + bias;
        }
    }
    if (HAS_A) {
        uint16_t *pa = pa_b + (size_t)t_first * p.out_a_ld;
        const int astep = p.out_stride * p.out_a_ld;
#pragma unroll
        for (int e = 0; e < 16; ++e) pa[e * astep] = cvt16(lrelu_max(__uint_as_float(v[e]) + bias, p.slope_out), FMT);
    }
}
template <int FMT>
__device__ __forceinline__ void epi_group16_poly(const uint32_t (&v)[16], float bias, const TcConvParams &p,
                                                 float *px_b /* out_x + (b*L4*C + co)*4, or + b*L_out*C + co */,
                                                 uint16_t *pa_b /* out_a + b*L_out*lda + co */, int t_first) {
    if (p.x_cl) {
        if (p.out_a) epi_group16_poly_t<FMT, true, true>(v, bias, p, px_b, pa_b, t_first);
        else epi_group16_poly_t<FMT, true, false>(v, bias, p, px_b, pa_b, t_first);
    } else {
        if (p.out_a) epi_group16_poly_t<FMT, false, true>(v, bias, p, px_b, pa_b, t_first);
        else epi_group16_poly_t<FMT, false, false>(v, bias, p, px_b, pa_b, t_first);
    }
}

// Quad row order (TcConvParams::qperm): the four lanes of a quad hold the four output samples t = 4 u .. 4 u + 3 of one
// channel.  A 4 x 4 transpose over the quad (two butterfly steps of shuffles) gives lane r the whole 16-byte slot for
// positions 4 m + r of the group: four 128-bit stores per thread instead of sixteen 4-byte stores 16 bytes apart (which
// wrote every 32-byte L2 sector four times: 56 M write sectors for 0.6 GB in the 256 -> 128 upsample).
__device__ __forceinline__ void quad_transpose4(float &a0, float &a1, float &a2, float &a3, int r) {
    const bool up = (r & 2) != 0, odd = (r & 1) != 0;
    float s0 = up ? a0 : a2, s1 = up ? a1 : a3;
    float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (up) { a0 = r0; a1 = r1; } else { a2 = r0; a3 = r1; }
    s0 = odd ? a0 : a1; s1 = odd ? a2 : a3;
    r0 = __shfl_xor_sync(0xffffffffu, s0, 1); r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    if (odd) { a0 = r0; a2 = r1; } else { a1 = r0; a3 = r1; }
}
// px_c: out_x + ((b * L4) * C + co) * 4; slot0: time slot (t / 4) of position ibase at phase group q_hi; slot_step = stride / 4
template <int FMT>
__device__ __forceinline__ void epi_group16_poly_quad(const uint32_t (&v)[16], float bias, const TcConvParams &p, float *px_c,
                                                      uint16_t *pa_b /* out_a + b*L_out*lda + co */, int t_first, int slot0,
                                                      int slot_step, int lane) {
    float f[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]) + bias;
    if (p.out_a) {
        uint16_t *pa = pa_b + (size_t)t_first * p.out_a_ld;
        const int astep = p.out_stride * p.out_a_ld;
#pragma unroll
        for (int e = 0; e < 16; ++e) pa[e * astep] = cvt16(lrelu_max(f[e], p.slope_out), FMT);
    }
    const int r = lane & 3;
    const int C4 = p.cout * 4;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        quad_transpose4(f[4 * m], f[4 * m + 1], f[4 * m + 2], f[4 * m + 3], r);
        // this lane: position 4 m + r of the group, samples q_lo = 0 .. 3
        *reinterpret_cast<float4 *>(px_c + (size_t)(slot0 + (4 * m + r) * slot_step) * C4) =
            make_float4(f[4 * m], f[4 * m + 1], f[4 * m + 2], f[4 * m + 3]);
    }
}

template <int FMT, int MODE>
__device__ __forceinline__ void epi_group16_c(int c_ct, const uint32_t (&v)[16], float bias, const TcConvParams &p,
                                              const EpiLoads &res, float *px, uint16_t *pa) {
    switch (c_ct) {
        case 32: epi_group16<FMT, 32, MODE>(v, bias, p, res, px, pa); return;
        case 64: epi_group16<FMT, 64, MODE>(v, bias, p, res, px, pa); return;
        case 128: epi_group16<FMT, 128, MODE>(v, bias, p, res, px, pa); return;
        case 256: epi_group16<FMT, 256, MODE>(v, bias, p, res, px, pa); return;
        default: epi_group16<FMT, 0, MODE>(v, bias, p, res, px, pa); return;
    }
}
template <int FMT>
__device__ __forceinline__ void epi_group16_dispatch(int mode, int c_ct, const uint32_t (&v)[16], float bias,
                                                     const TcConvParams &p, const EpiLoads &res, float *px, uint16_t *pa) {
    switch (mode) {
        case EPI_A: epi_group16_c<FMT, EPI_A>(c_ct, v, bias, p, res, px, pa); return;
        case EPI_RXA: epi_group16_c<FMT, EPI_RXA>(c_ct, v, bias, p, res, px, pa); return;
        case EPI_RX: epi_group16_c<FMT, EPI_RX>(c_ct, v, bias, p, res, px, pa); return;
        case EPI_RCX: epi_group16_c<FMT, EPI_RCX>(c_ct, v, bias, p, res, px, pa); return;
        case EPI_RCDXA: epi_group16_c<FMT, EPI_RCDXA>(c_ct, v, bias, p, res, px, pa); return;
        default: epi_group16_c<FMT, EPI_RCDX>(c_ct, v, bias, p, res, px, pa); return;
    }
}

// ROWB: bytes per operand row: 128 (64 channels, SW128) or 64 (32 channels, SW64); FMT: 0 bf16, 1 fp16
template <int ROWB, int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w,
               const TcConvParams p) {
    constexpr int CH = ROWB / 2;                 // bf16 channels per chunk
    constexpr int KSTEPS = CH / 16;              // UMMA K = 16
    constexpr int ACT_BYTES = ACT_ROWS * ROWB;
    constexpr int W_BYTES = TM * ROWB;
    extern __shared__ __align__(1024) uint8_t smem[];
    // swizzle atoms need 1024-byte aligned stage bases; every stage size is a multiple of 1024
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("vtts: dynamic shared memory base not 1024-byte aligned\n");
        __trap();
    }
    if ((p.trace & 1) && blockIdx.x == 0 && threadIdx.x == 0) g_trace[15] = clock64();          // debug trace: kernel entry
    uint8_t *s_act = smem;
    const uint32_t ACT_STAGES = (uint32_t)p.act_stages, W_STAGES = (uint32_t)p.w_stages;
    uint8_t *s_w = smem + (size_t)ACT_STAGES * ACT_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_w + (size_t)W_STAGES * W_BYTES * p.tps);
    uint64_t *act_full = bars, *act_empty = act_full + ACT_STAGES;
    uint64_t *w_full = act_empty + ACT_STAGES, *w_empty = w_full + W_STAGES;
    uint64_t *acc_full = w_empty + W_STAGES, *acc_empty = acc_full + ACC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + ACC_STAGES);
    int *s_lim = reinterpret_cast<int *>(tmem_slot + 2);   // [MAX_TRIM_BATCH] per-batch tile-start limits (padding trim)
    int *s_ioff = s_lim + MAX_TRIM_BATCH;                  // [MAX_TRIM_BATCH + 1] first live work item of every batch row

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t pol = l2_policy(p.stream_hint != 0);   // eviction policy of the activation / residual reads
    const int last_off = p.tap_off0 + (p.taps - 1) * p.tap_step;
    const int min_off = p.tap_off0 < last_off ? p.tap_off0 : last_off;
    const int span = (p.tap_off0 < last_off ? last_off : p.tap_off0) - min_off;
    const int nbox = (TN + span + BOX_ROWS - 1) / BOX_ROWS;

    // cluster geometry: CL CTAs share every weight tile (each loads 1/CL of its rows and multicasts)
    const int CL = p.cluster;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int cid = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    // work item -> (first output row, first time position, batch); time tiles past the end are dummies
    // (i0 >= n_pos: loads are zero-filled, the epilogue skips them) that keep the barrier protocol uniform
    // Work items are walked with a mixed-radix counter (m block, time group, batch) advanced by `ncl` with
    // carries: no div/mod on the per-tile path of the single-thread roles.
    // With the padding trim the item space is COMPACT: only live tiles are numbered (per-batch offsets in s_ioff), so
    // the static round-robin over CTAs stays balanced however the utterance lengths fall (a strided walk over
    // (m, tile, batch) with dead tiles skipped left some CTAs with twice the work of others on the short stages).
    struct TileIter {
        int item, m, g, b, dm, dg, db, mb, gpb, nb, total, rev;
        const int *ioff;
        __device__ void init(int first, int step, int m_blocks, int groups, const int *offsets = nullptr, int batch = 0,
                             int n_total = 0, int reverse = 0) {
            mb = m_blocks; gpb = groups; ioff = offsets; nb = batch; total = n_total; rev = reverse;
            item = first;
            if (ioff) { b = rev ? nb - 1 : 0; locate(); return; }
            m = first % mb; int r = first / mb; g = r % gpb; b = r / gpb;
            dm = step % mb; r = step / mb; dg = r % gpb; db = r / gpb;
        }
        __device__ void locate() {                 // compact numbering; rev walks it last-to-first (see TcConvParams::reverse)
            if (item >= total) return;
            const int li = rev ? total - 1 - item : item;
            if (!rev) { while (b < nb && li >= ioff[b + 1]) ++b; } else { while (b > 0 && li < ioff[b]) --b; }
            const int r = li - ioff[b];
            g = r / mb; m = r - g * mb;
        }
        __device__ void next(int step) {
            item += step;
            if (ioff) { locate(); return; }
            m += dm; int c = 0;
            if (m >= mb) { m -= mb; c = 1; }
            g += dg + c; c = 0;
            if (g >= gpb) { g -= gpb; c = 1; }
            b += db + c;
        }
    };
    auto decode = [&](const TileIter &it, int &n0, int &i0, int &b) {
        n0 = it.m * TM;
        b = it.b;
        i0 = (it.g * CL + crank) * TN;
    };
    // padding trim: per-batch position limits live in shared memory (filled below); cluster launches keep every
    // tile because both CTAs of a cluster must stay in lock step
    const bool trimming = p.lens != nullptr && CL == 1;
    auto tile_live = [&](int i0, int b) -> bool { return !trimming || i0 < s_lim[b]; };

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < ACT_STAGES; ++s) { mbar_init(&act_full[s], 1); mbar_init(&act_empty[s], 1); }
        for (uint32_t s = 0; s < W_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], (uint32_t)CL); }
        for (int s = 0; s < ACC_STAGES; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], (uint32_t)(p.epi_quarters * (EPI_WARPS / 4)));  // one arrival per participating warp
        }
        fence_barrier_init();
    }
    // Warp roles.  The three single-thread roles sit in the HIGHEST warp ids: the warp scheduler favours high
    // warp ids, and these threads are the critical path (a starved MMA issuer idles the tensor pipe).
    constexpr int WARP_ACT = EPI_WARPS, WARP_W = EPI_WARPS + 1, WARP_MMA = EPI_WARPS + 1 + W_PRODUCERS;
    if (warp == WARP_MMA) {  // MMA warp owns the TMEM allocation (all 512 columns: 2 accumulators)
        tmem_alloc(tmem_slot, ACC_STAGES * TN);
        tmem_relinquish();
    }
    if (p.lens != nullptr)
        for (int i = threadIdx.x; i < p.batch; i += blockDim.x) {
            const long long lim = (__ldg(p.lens + i) + p.len_margin) * (long long)p.len_rate + p.len_extra;
            s_lim[i] = lim > 0x7fffffffLL ? 0x7fffffff : (int)lim;
        }
    if (trimming) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int i = 0; i < p.batch; ++i) {
                s_ioff[i] = acc;
                const int lim = s_lim[i] < 0 ? 0 : s_lim[i];
                int live = (lim + TN - 1) / TN;                       // tiles that start before the limit
                if (live > p.groups_per_batch) live = p.groups_per_batch;
                acc += live * p.m_blocks;
            }
            s_ioff[p.batch] = acc;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (CL > 1) cluster_sync_all();   // peer barriers are initialised before any multicast / remote arrive
    const int n_items = trimming ? s_ioff[p.batch] : p.total_tiles;
    const int *ioff = trimming ? s_ioff : nullptr;
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: everything above overlapped the previous kernel's tail.  The weight producer never
    // touches data of the previous kernel, so it does not wait at all and fills its ring early.
    if (warp != WARP_W) { grid_dep_wait(); grid_dep_launch(); }
    if ((p.trace & 1) && blockIdx.x == 0 && threadIdx.x == 0) g_trace[14] = clock64();          // debug trace: previous kernel done

    if (warp == WARP_ACT) {
        // ===== activation producer: one (TN + span)-row tile per K chunk, reused by every tap =====
        if (lane == 0) {
            tma_prefetch_desc(&tm_act);
            uint32_t s = 0, ph = 0, tl = 0;                     // stage index and its parity, kept incrementally
            TileIter ti;
            for (ti.init(cid, ncl, p.m_blocks, p.groups_per_batch, ioff, p.batch, n_items, p.reverse); ti.item < n_items; ti.next(ncl), ++tl) {
                int n0, i0, b;
                decode(ti, n0, i0, b);
                (void)n0;
                if (!tile_live(i0, b)) continue;
                for (int c = 0; c < p.chunks; ++c) {
                    if (c == 0) VTTS_TRACE(4);
                    mbar_wait_producer(&act_empty[s], ph ^ 1u);
                    if (c == 0) VTTS_TRACE(5);
                    mbar_arrive_expect_tx(&act_full[s], (uint32_t)(nbox * BOX_ROWS * ROWB));
                    for (int bx = 0; bx < nbox; ++bx)
                        tma_load_3d_hint(s_act + (size_t)s * ACT_BYTES + (size_t)bx * BOX_ROWS * ROWB, &tm_act,
                                         &act_full[s], c * CH, i0 + min_off + bx * BOX_ROWS, b, pol);
                    if (++s == ACT_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp >= WARP_W && warp < WARP_W + W_PRODUCERS) {
        // ===== weight producers: one 128-row tile per (chunk, tap) step.  A producer iteration (empty-barrier
        // probe + expect_tx + TMA issue) has ~1 us of serial latency, far more than the MMA time of a step, so
        // W_PRODUCERS warps run round-robin over the steps; W_STAGES is a multiple of W_PRODUCERS, so each warp
        // owns a fixed subset of the stages and the barriers need no cross-warp coordination. =====
        if (lane == 0) {
            tma_prefetch_desc(&tm_w);
            const int my_rows = p.w_rows / CL;                  // this CTA's share of every weight tile
            uint32_t s = 0, ph = 0, tl = 0;                     // stage index and its parity, kept incrementally
            const int tps = p.tps;                               // taps per stage (the box depth of tm_w)
            const size_t stage_bytes = (size_t)W_BYTES * tps;
            TileIter ti;
            for (ti.init(cid, ncl, p.m_blocks, p.groups_per_batch, ioff, p.batch, n_items, p.reverse); ti.item < n_items; ti.next(ncl), ++tl) {
                int n0, i0w, bw;
                decode(ti, n0, i0w, bw);
                if (!tile_live(i0w, bw)) continue;
                for (int c = 0; c < p.chunks; ++c)
                    for (int j = 0; j < p.taps; j += tps) {
                        if ((c | j) == 0) VTTS_TRACE(6);
                        mbar_wait(&w_empty[s], ph ^ 1u);   // hot poll: the weight ring is short and on the critical path
                        if ((c | j) == 0) VTTS_TRACE(7);
                        // a box past the last tap is zero-filled by TMA: the padded tap contributes nothing
                        mbar_arrive_expect_tx(&w_full[s], (uint32_t)(p.w_rows * ROWB * tps));   // own + peer shares
                        if (CL > 1)
                            tma_load_3d_mc(s_w + (size_t)s * stage_bytes + (size_t)crank * my_rows * ROWB, &tm_w, &w_full[s],
                                           c * CH, n0 + crank * my_rows, j, cmask);
                        else
                            tma_load_3d(s_w + (size_t)s * stage_bytes, &tm_w, &w_full[s], c * CH, n0, j);
                        if (++s == W_STAGES) { s = 0; ph ^= 1u; }
                    }
            }
        }
    } else if (warp == WARP_MMA) {
        // ===== MMA issuer =====
        // One thread issues every tcgen05.mma.  ncu showed this thread to be instruction-latency bound: the
        // MMA queue is shallow, so every dependent scalar instruction between two issues idles the tensor pipe.
        // The loop therefore keeps all state incrementally (no div/mod, descriptors advanced by adds), and each
        // step probes the NEXT stage's barrier before issuing the current MMAs (umma_step*).
        // (the whole warp runs this loop converged; elect.sync inside the step picks the issuing lane)
        {
            constexpr uint32_t idesc = make_idesc_16(TM, TN, FMT);
            const uint32_t wfull0 = smem_u32(w_full), wempty0 = smem_u32(w_empty);
            const uint64_t adesc_first = make_smem_desc(smem_u32(s_w), ROWB, 0);       // weight stage 0
            const uint64_t bdesc_first = make_smem_desc(smem_u32(s_act), ROWB, 0);     // activation stage 0, row 0
            constexpr uint64_t A_STAGE_STEP = (uint64_t)(W_BYTES >> 4), B_STAGE_STEP = (uint64_t)(ACT_BYTES >> 4);
            const long long tap0 = (long long)(p.tap_off0 - min_off) * (ROWB >> 4);    // row of tap 0, in 16-byte units
            const long long tap_step = (long long)p.tap_step * (ROWB >> 4);
            uint32_t sa = 0, aph = 0, sw = 0, wph = 0, tl = 0;   // tl counts PROCESSED tiles (accumulator ring)
            uint64_t adesc = adesc_first, bstage = bdesc_first;
            uint32_t w_ready = 0;
            const int tps = p.tps;
            const uint64_t a_stage_step = A_STAGE_STEP * (uint64_t)tps;
            TileIter ti;
            for (ti.init(cid, ncl, p.m_blocks, p.groups_per_batch, ioff, p.batch, n_items, p.reverse); ti.item < n_items; ti.next(ncl)) {
                {
                    int n0m, i0m, bm;
                    decode(ti, n0m, i0m, bm);
                    if (!tile_live(i0m, bm)) continue;
                }
                const uint32_t buf = tl % ACC_STAGES;
                if (lane == 0) VTTS_TRACE(0);
                mbar_wait(&acc_empty[buf], ((tl / ACC_STAGES) & 1u) ^ 1u);   // epilogue drained this accumulator
                if (lane == 0) VTTS_TRACE(1);
                const uint32_t tmem_d = tmem_base + buf * TN;
                uint32_t acc = 0;                                            // first MMA of the tile overwrites
                for (int c = 0; c < p.chunks; ++c) {
                    mbar_wait(&act_full[sa], aph);
                    if (c == 0 && lane == 0) VTTS_TRACE(2);
                    uint64_t bdesc = bstage + (uint64_t)tap0;
                    for (int j = 0; j < p.taps; j += tps) {
                        if (!w_ready) mbar_wait_addr(wfull0 + sw * 8u, wph);
                        tc_fence_after();
                        // next weight stage (after the very last step the probe simply reports "not ready")
                        uint32_t sn = sw + 1, pn = wph;
                        if (sn == W_STAGES) { sn = 0; pn ^= 1u; }
                        if (CL > 1) {
                            if (lane == 0) {
#pragma unroll
                                for (int ks = 0; ks < KSTEPS; ++ks)
                                    umma_bf16(tmem_d, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc, acc | (uint32_t)ks);
                                umma_commit_mc(&w_empty[sw], cmask);   // stage reusable in every CTA of the cluster
                            }
                            __syncwarp();
                            w_ready = 0;
                        } else if (tps == 2) {
                            // second tap of the step; for an odd tap count its weights are TMA zero-fill, any row works
                            const uint64_t bdesc1 = (j + 1 < p.taps) ? bdesc + (uint64_t)tap_step : bdesc;
                            if (KSTEPS == 4)
                                w_ready = umma_step4x2_warp(tmem_d, adesc, bdesc, idesc, acc, wfull0 + sn * 8u, pn,
                                                            wempty0 + sw * 8u, bdesc1, A_STAGE_STEP);
                            else
                                w_ready = umma_step2x2_warp(tmem_d, adesc, bdesc, idesc, acc, wfull0 + sn * 8u, pn,
                                                            wempty0 + sw * 8u, bdesc1, A_STAGE_STEP);
                            bdesc += (uint64_t)tap_step;
                        } else if (KSTEPS == 4) {
                            w_ready = umma_step4_warp(tmem_d, adesc, bdesc, idesc, acc, wfull0 + sn * 8u, pn, wempty0 + sw * 8u);
                        } else {
                            w_ready = umma_step2_warp(tmem_d, adesc, bdesc, idesc, acc, wfull0 + sn * 8u, pn, wempty0 + sw * 8u);
                        }
                        acc = 1;
                        bdesc += (uint64_t)tap_step;
                        adesc += a_stage_step;
                        sw = sn; wph = pn;
                        if (sn == 0) adesc = adesc_first;
                    }
                    umma_commit_elect(&act_empty[sa]);     // activation stage reusable
                    bstage += B_STAGE_STEP;
                    if (++sa == ACT_STAGES) { sa = 0; aph ^= 1u; bstage = bdesc_first; }
                }
                if (lane == 0) VTTS_TRACE(3);
                umma_commit_elect(&acc_full[buf]);         // accumulator complete -> epilogue
                ++tl;
            }
        }
    } else {
        // ===== epilogue warps: thread = output row (channel), registers = time positions =====
        const int ew = warp;                                // epilogue warps are warps 0 .. EPI_WARPS-1
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
        constexpr int SHARERS = EPI_WARPS / 4;              // warps sharing a lane quarter split the columns
        // narrow layers replicate their weight rows `rep` times over the 128 lanes; replica r is read by
        // the warps of its lane quarters and covers columns [r*TN/rep, (r+1)*TN/rep)
        const int qpc = 4 / p.rep;                          // lane quarters per replica
        const int cols_per_warp = TN / p.rep / SHARERS;     // 64 / 32 / 16
        const int col_lo = (quarter / qpc) * (TN / p.rep) + (ew / 4) * cols_per_warp;
        const int r_in_copy = (quarter % qpc) * 32 + lane;  // output row of this thread within the m-block
        // kernel-uniform epilogue specialisation
        const bool R = p.res != nullptr, C = p.accumulate != 0, D = p.divide_by > 0.f, X = p.out_x != nullptr,
                   A = p.out_a != nullptr;
        const int mode = (!R && !C && !D && !X && A) ? EPI_A
                       : (R && !C && !D && X && A) ? EPI_RXA
                       : (R && !C && !D && X && !A) ? EPI_RX
                       : (R && C && !D && X && !A) ? EPI_RCX
                       : (R && C && D && X && A) ? EPI_RCDXA
                       : (R && C && D && X && !A) ? EPI_RCDX : EPI_GENERIC;
        const bool unit = p.out_stride == 1 && p.out_off0 == 0 && p.n_total == p.cout && mode != EPI_GENERIC && !p.x_cl && !p.act_tanh;
        const int c_ct = (!A || p.out_a_ld == p.cout) ? p.cout : 0;
        // polyphase upsample: fp32 x + 16-bit copy, nothing read; 32-bit index math is safe below 2^31 elements per row
        const bool poly = !R && !C && !D && X && (A || p.x_cl) && p.out_stride > 1 && (long long)p.L4 * p.cout * 4 < 0x7fffffffLL &&
                          (long long)p.L_out * (A ? p.out_a_ld : 1) < 0x7fffffffLL;
        // L2 prefetch of the fp32 streams the epilogue will read (residual, running MRF sum), one tile ahead,
        // spread over all epilogue threads
        const int et = threadIdx.x;                                // 0 .. EPI_WARPS*32-1
        auto prefetch_tile = [&](const TileIter &tile) {
            if (!(R || C) || !unit) return;
            int n0p, i0p, bp;
            decode(tile, n0p, i0p, bp);
            if (!tile_live(i0p, bp)) return;
            const int rows_ch = (p.n_total - n0p) < TM ? (p.n_total - n0p) : TM;   // channels of this m-block
            const int lines_per_row = (rows_ch * 16 + 127) / 128;                  // one t/4 row = rows_ch * 16 bytes
            const int n_lines = (TN / 4) * lines_per_row;
            for (int l = et; l < n_lines; l += EPI_WARPS * 32) {
                const int row4 = l / lines_per_row, seg = l - row4 * lines_per_row;
                const int t = i0p + row4 * 4;
                if (t < p.n_pos) {
                    const long long off = (((long long)bp * p.L4 + (t >> 2)) * p.cout + n0p) * 4 + seg * 32;
                    if (R) prefetch_l2(p.res + off);
                    if (C) prefetch_l2(p.out_x + off);
                }
            }
        };
        TileIter ti, tnext;
        ti.init(cid, ncl, p.m_blocks, p.groups_per_batch, ioff, p.batch, n_items, p.reverse);
        tnext = ti;
        if (cid < n_items) prefetch_tile(ti);
        // with a single m block and no per-batch bias the thread's bias never changes: load it once
        const bool bias_fixed = p.m_blocks == 1 && p.bias_b == nullptr;
        float bias_const = 0.f;
        if (bias_fixed && p.bias && r_in_copy < p.n_total) bias_const = __ldg(p.bias + (p.qperm ? (r_in_copy >> 2) % p.cout : r_in_copy % p.cout));
        uint32_t tl = 0;                                    // counts PROCESSED tiles (accumulator ring)
        for (; ti.item < n_items; ti.next(ncl)) {
            tnext.next(ncl);
            if (tnext.item < n_items) prefetch_tile(tnext);
            if (quarter >= p.epi_quarters) continue;        // this warp's TMEM lanes never hold real rows: prefetch duty only
            int n0, i0, b;
            decode(ti, n0, i0, b);
            if (!tile_live(i0, b)) continue;
            const uint32_t buf = tl % ACC_STAGES;
            const int nq = n0 + (quarter % qpc) * 32;       // first output row of this warp
            const int n = n0 + r_in_copy;                   // global output row of this thread
            const bool row_ok = n < p.n_total;
            // output row -> (phase q, channel co).  Phase-major rows: the phase is warp-uniform (cout is a multiple of 32);
            // quad order: q = 4 q_hi + lane % 4 with a warp-uniform q_hi, 8 channels per warp
            int phase = p.n_total == p.cout ? 0 : nq / p.cout;
            int co = n - phase * p.cout;
            const int q_hi4 = p.qperm ? ((nq >> 2) / p.cout) * 4 : 0;
            if (p.qperm) { co = (n >> 2) - (q_hi4 >> 2) * p.cout; phase = q_hi4 + (n & 3); }
            float bias = bias_const;
            if (!bias_fixed) {
                bias = 0.f;
                if (row_ok && p.bias) bias = __ldg(p.bias + co);
                if (row_ok && p.bias_b) bias = bias + __ldg(p.bias_b + (size_t)b * p.cout + co);
            }
            const bool quarter_used = nq < p.n_total;       // warp-uniform
            const bool rows_full = nq + 32 <= p.n_total;    // warp-uniform
            const int n_valid = p.n_pos < p.L_out ? p.n_pos : p.L_out;
            // fully valid unit-stride group: 128-bit fast path
            auto group_fast = [&](int ibase) { return unit && rows_full && ibase + 16 <= n_valid; };
            auto res_ptr = [&](int ibase) { return p.res + (((long long)b * p.L4 + (ibase >> 2)) * p.cout + co) * 4; };
            // issue the first group's residual loads before waiting for the accumulator
            EpiLoads cur{}, nxt{};
            if (quarter_used && R && group_fast(i0 + col_lo)) epi_load16<0>(nxt, res_ptr(i0 + col_lo), p.cout, pol);
            if (ew == 0 && lane == 0) VTTS_TRACE(8);
            mbar_wait_relaxed(&acc_full[buf], (tl / ACC_STAGES) & 1u);
            if (ew == 0 && lane == 0) VTTS_TRACE(9);
            tc_fence_after();
            if (quarter_used) {
                for (int cg = 0; cg < cols_per_warp; cg += 16) {
                    const int col = col_lo + cg;
                    const int ibase = i0 + col;
                    if (ibase >= p.n_pos) break;            // warp-uniform: nothing valid beyond
                    uint32_t v[16];
                    tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * TN + (uint32_t)col, v);
                    cur = nxt;
                    const bool fast = group_fast(ibase);
                    if (R && cg + 16 < cols_per_warp && group_fast(ibase + 16)) epi_load16<0>(nxt, res_ptr(ibase + 16), p.cout, pol);
                    tmem_ld_wait();
                    if (fast) {
                        float *px = X ? p.out_x + (((long long)b * p.L4 + (ibase >> 2)) * p.cout + co) * 4 : nullptr;
                        uint16_t *pa = A ? p.out_a + ((long long)b * p.L_out + ibase) * p.out_a_ld + co : nullptr;
                        epi_group16_dispatch<FMT>(mode, c_ct, v, bias, p, cur, px, pa);
                    } else if (poly && p.qperm && !p.x_cl && rows_full && ibase + 16 <= p.n_pos &&
                               (long long)ibase * p.out_stride + p.out_off0 + q_hi4 >= 0 &&
                               (long long)(ibase + 15) * p.out_stride + p.out_off0 + q_hi4 + 3 < p.L_out) {
                        // (warp-uniform test over the whole quad: the shuffles inside need all 32 lanes)
                        epi_group16_poly_quad<FMT>(v, bias, p, p.out_x + ((long long)b * p.L4 * p.cout + co) * 4,
                                                   A ? p.out_a + (long long)b * p.L_out * p.out_a_ld + co : nullptr,
                                                   ibase * p.out_stride + p.out_off0 + phase,
                                                   (ibase * p.out_stride + p.out_off0 + q_hi4) >> 2, p.out_stride >> 2, lane);
                    } else if (poly && rows_full && ibase + 16 <= p.n_pos &&
                               (long long)ibase * p.out_stride + p.out_off0 + phase >= 0 &&
                               (long long)(ibase + 15) * p.out_stride + p.out_off0 + phase < p.L_out) {
                        epi_group16_poly<FMT>(v, bias, p,
                                              p.x_cl ? p.out_x + (long long)b * p.L_out * p.cout + co
                                                     : p.out_x + ((long long)b * p.L4 * p.cout + co) * 4,
                                              A ? p.out_a + (long long)b * p.L_out * p.out_a_ld + co : nullptr,
                                              ibase * p.out_stride + p.out_off0 + phase);
                    } else {
                        epi_group16_edge<FMT>(v, bias, p, row_ok, b, ibase, phase, co);
                    }
                }
            }
            if (ew == 0 && lane == 0) VTTS_TRACE(10);
            // all of this warp's tcgen05.ld have completed (wait::ld above): release the accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            ++tl;
        }
    }
    tc_fence_before();
    __syncthreads();
    if ((p.trace & 1) && blockIdx.x == 0 && threadIdx.x == 0) g_trace[16 + 15] = clock64();     // debug trace: all roles done
    if (CL > 1) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
    if (warp == WARP_MMA) tmem_dealloc(tmem_base, ACC_STAGES * TN);
}

struct TcLaunch {
    CUtensorMap tm_act, tm_w;
    TcConvParams p;
    int rowb, fmt;
    bool pdl = false;          // launch with programmatic stream serialization (forward chain only)
    dim3 grid;
    size_t smem;
};

static size_t tc_smem_bytes(int rowb, int act_stages, int w_stages) {
    return (size_t)act_stages * ACT_ROWS * rowb + (size_t)w_stages * TM * rowb +
           (size_t)(2 * act_stages + 2 * w_stages + 2 * ACC_STAGES) * 8 + 16 + (2 * MAX_TRIM_BATCH + 1) * sizeof(int);
}

static bool tc_cluster_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("VTTS_TC_CLUSTER"); on = (e && e[0] == '1') ? 1 : 0; }  // off by default: no measured benefit
    return on == 1;
}

static int tc_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <int ROWB, int FMT>
static int tc_launch_t(const TcLaunch &L, cudaStream_t st) {
    // the opt-in shared-memory size is a per-device function attribute: set it once per device, not once per process
    static bool attr[64] = {};
    int dev = 0;
    VTTS_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<ROWB, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    VTTS_CHECK_CUDA(launch_kernel_ex(conv_tc_kernel<ROWB, FMT>, L.grid, dim3(TC_THREADS), L.smem, st, L.pdl, (unsigned)L.p.cluster, L.tm_act,
                                     L.tm_w, L.p));
    return VTTS_OK;
}

static int tc_launch(const TcLaunch &L, cudaStream_t st) {
    if (L.rowb == 128) return L.fmt == VTTS_FMT_BF16 ? tc_launch_t<128, 0>(L, st) : tc_launch_t<128, 1>(L, st);
    return L.fmt == VTTS_FMT_BF16 ? tc_launch_t<64, 0>(L, st) : tc_launch_t<64, 1>(L, st);
}

// Build a launch for one layer.  act: (B, L_in, ci_pad) bf16; w: [taps][n_pad][ci_pad] bf16.
static int tc_prepare(TcLaunch &L, int fmt, const uint16_t *act, int B, int L_in, int ci_pad,
                      const uint16_t *w, int n_pad, TcConvParams p) {
    const int rowb = (ci_pad % 64 == 0) ? 128 : 64;
    const int ch = rowb / 2;
    if (ci_pad % ch != 0) return set_error(VTTS_E_UNSUPPORTED, "tc: ci_pad %d not a multiple of %d", ci_pad, ch);
    const int last_off = p.tap_off0 + (p.taps - 1) * p.tap_step;
    const int span = last_off > p.tap_off0 ? last_off - p.tap_off0 : p.tap_off0 - last_off;
    if (span > HALO_MAX) return set_error(VTTS_E_UNSUPPORTED, "tc: tap span %d > %d", span, HALO_MAX);
    if (p.bias_b && p.n_total != p.cout) return set_error(VTTS_E_INVALID, "tc: per-batch bias needs a single phase");
    if (p.slope_out < 0.f || p.slope_out > 1.f) return set_error(VTTS_E_UNSUPPORTED, "tc: LeakyReLU slope %g outside [0,1]", p.slope_out);
    if (p.out_a && p.cout % 32 != 0) return set_error(VTTS_E_UNSUPPORTED, "tc: cout %d not a multiple of 32", p.cout);
    L.rowb = rowb;
    L.fmt = fmt;
    p.inv_div = p.divide_by > 0.f ? 1.f / p.divide_by : 1.f;
    p.chunks = ci_pad / ch;
    p.m_blocks = n_pad / TM;
    p.t_tiles = ceil_div(p.n_pos, TN);
    // 2-CTA clusters multicast the weight tiles (halves the L2 -> SMEM weight traffic that bounds the
    // MMA-heavy layers); not worth it when a batch row has a single time tile
    p.cluster = (p.t_tiles >= 2 && tc_cluster_enabled()) ? 2 : 1;
    p.groups_per_batch = ceil_div(p.t_tiles, p.cluster);
    const long long total = (long long)p.m_blocks * p.groups_per_batch * B;
    if (total > 0x7fffffffLL) return set_error(VTTS_E_UNSUPPORTED, "tc: too many tiles");
    p.total_tiles = (int)total;
    // pipeline depths: narrow (HBM-bound) layers need many activation bytes in flight, wide
    // (tensor-bound) layers a deep weight pipeline
    p.batch = B;
    if (p.lens && B > MAX_TRIM_BATCH) p.lens = nullptr;   // limits table is in shared memory
    // narrow layers (32 / 64 output rows): the packed weights repeat the rows 4x / 2x over the 128 lanes so
    // that every TMEM lane quarter (= every SM sub-partition's epilogue warps) holds a replica
    p.rep = (p.m_blocks == 1 && (p.n_total == 32 || p.n_total == 64)) ? TM / p.n_total : 1;
    p.w_rows = (p.rep > 1 || p.n_total >= TM) ? TM : ((p.n_total + 31) / 32) * 32;
    // quarters that hold real rows in every tile (a partial last m-block keeps all warps in the handshake)
    p.epi_quarters = (p.rep == 1 && p.m_blocks == 1 && p.n_total < TM) ? (p.n_total + 31) / 32 : 4;
    p.L4 = (p.L_out + 3) / 4;
    // single-chunk (narrow) layers: two taps per pipeline step (needs full 128-row weight tiles in the stage)
    p.tps = (p.chunks == 1 && p.cluster == 1 && p.taps >= 2 && p.w_rows == TM) ? 2 : 1;
    if (rowb == 64) { p.act_stages = 6; p.w_stages = 8 / p.tps; }
    else if (p.chunks == 1) { p.act_stages = p.tps == 2 ? 3 : 4; p.w_stages = p.tps == 2 ? 3 : 4; }
    else { p.act_stages = 2; p.w_stages = 8; }
    L.p = p;
    L.smem = tc_smem_bytes(rowb, p.act_stages, p.w_stages * p.tps);
    if (L.smem > 227 * 1024) return set_error(VTTS_E_UNSUPPORTED, "tc: %zu B shared memory", L.smem);
    const int sms = tc_num_sms();
    const int max_clusters = sms / p.cluster;
    L.grid = dim3((unsigned)((p.total_tiles < max_clusters ? p.total_tiles : max_clusters) * p.cluster));
    {
        uint64_t dims[3] = {(uint64_t)ci_pad, (uint64_t)L_in, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)ci_pad * 2, (uint64_t)ci_pad * 2 * (uint64_t)L_in};
        uint32_t box[3] = {(uint32_t)ch, BOX_ROWS, 1};
        int rc = make_tmap_bf16(&L.tm_act, act, 3, dims, str, box, rowb);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)ci_pad, (uint64_t)n_pad, (uint64_t)p.taps};
        uint64_t str[2] = {(uint64_t)ci_pad * 2, (uint64_t)ci_pad * 2 * (uint64_t)n_pad};
        uint32_t box[3] = {(uint32_t)ch, (uint32_t)(p.w_rows / p.cluster), (uint32_t)p.tps};
        int rc = make_tmap_bf16(&L.tm_w, w, 3, dims, str, box, rowb);
        if (rc) return rc;
    }
    return VTTS_OK;
}

// ---------------------------------------------------------------------------------------------
// fused ResidualBlock unit:  x_new = conv2(lrelu(conv1(a))) + x     (layers.py:93-97, one (convs1[i], convs2[i]) pair)
//
// Same persistent warp-specialised structure as conv_tc_kernel, but a tile runs two GEMM phases and the
// intermediate xt never leaves the SM: phase A (conv1, dilated) accumulates 256 positions into TMEM
// accumulator A; the epilogue warps turn it into the 16-bit LeakyReLU'd operand and write it straight into a
// swizzled shared-memory tile (zero outside [0, L): conv2 pads ITS input with zeros);
// phase B (conv2, dilation 1) reads that tile through row-shifted descriptors and accumulates 240 positions
// into accumulator B, which gets the usual fused epilogue (bias, residual, MRF sum/mean, 16-bit copy).
// Saves the xt round trip through HBM (4 of the 16 bytes per element of a unit), one launch and one tile pass.
// ---------------------------------------------------------------------------------------------
constexpr int UN2 = 240;           // conv2 output positions per tile (UMMA N of phase B)
constexpr int UXT_OFF = 8;         // xt row 0 is position i0 - 8 (covers conv2 half-widths up to 8)

struct TcUnitParams {
    TcConvParams e;                // phase-B epilogue + shared geometry (taps/tap_off0/tap_step describe conv1)
    const float *bias1;            // conv1 bias (C) or null
    float slope_mid;               // LeakyReLU between conv1 and conv2
    int taps2;                     // conv2 taps (dilation 1)
    int xt_chunks;                 // C / chunk channels
    int ea_warps;                  // epilogue warps that build the operand tile (4 or 8); the other 16 - ea_warps drain B
};

template <int ROWB, int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1)
unit_tc_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w1,
               const __grid_constant__ CUtensorMap tm_w2, const TcUnitParams u) {
    const TcConvParams &p = u.e;
    constexpr int CH = ROWB / 2;
    constexpr int KSTEPS = CH / 16;
    constexpr int ACT_BYTES = ACT_ROWS * ROWB;
    constexpr int W_BYTES = TM * ROWB;
    constexpr int XT_BYTES = TN * ROWB;          // one 64/32-channel chunk of the xt tile
    extern __shared__ __align__(1024) uint8_t smem[];
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) { printf("vtts: smem base not 1024-byte aligned\n"); __trap(); }
    const uint32_t ACT_STAGES = (uint32_t)p.act_stages, W_STAGES = (uint32_t)p.w_stages;
    uint8_t *s_act = smem;
    uint8_t *s_w = s_act + (size_t)ACT_STAGES * ACT_BYTES;
    uint8_t *s_xt = s_w + (size_t)W_STAGES * W_BYTES * p.tps;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_xt + (size_t)u.xt_chunks * XT_BYTES);
    uint64_t *act_full = bars, *act_empty = act_full + ACT_STAGES;
    uint64_t *w_full = act_empty + ACT_STAGES, *w_empty = w_full + W_STAGES;
    uint64_t *accA_full = w_empty + W_STAGES, *xt_full = accA_full + 1, *accB_full = xt_full + 1, *accB_empty = accB_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accB_empty + 1);
    int *s_lim = reinterpret_cast<int *>(tmem_slot + 2);
    int *s_ioff = s_lim + MAX_TRIM_BATCH;          // [MAX_TRIM_BATCH + 1] first live tile index of every batch row

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t pol = l2_policy(p.stream_hint != 0);   // eviction policy of the activation / residual reads
    const int last_off = p.tap_off0 + (p.taps - 1) * p.tap_step;
    const int min_off = p.tap_off0 < last_off ? p.tap_off0 : last_off;
    const int span = (p.tap_off0 < last_off ? last_off : p.tap_off0) - min_off;
    const int nbox = (TN + span + BOX_ROWS - 1) / BOX_ROWS;
    const int h2 = (u.taps2 - 1) / 2;
    const int ncta = (int)gridDim.x;
    // epilogue warps: the first ea_warps (4 or 8) build the operand tile, the rest drain the output accumulator.  The
    // HBM-bound k=3 units give the output side 12 warps (more loads in flight), the MMA-bound ones keep 8 + 8.
    const int ea_warps = u.ea_warps, eb_warps = EPI_WARPS - u.ea_warps;

    // work items: (time tile of UN2 outputs, batch); m_blocks == 1.  With the padding trim only LIVE tiles are numbered
    // (per-batch offsets in s_ioff, binary search per tile), so the round-robin over CTAs stays balanced.
    const bool trimming = p.lens != nullptr;
    int n_items = 0;                                       // set once the per-batch offsets exist (below)
    auto tile_i0 = [&](int item, int &i0, int &b) {
        if (p.reverse) item = n_items - 1 - item;          // last-to-first (see TcConvParams::reverse)
        if (!trimming) { b = item / p.t_tiles; i0 = (item - b * p.t_tiles) * UN2; return; }
        int lo = 0, hi = p.batch - 1;                      // last b with s_ioff[b] <= item
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_ioff[mid] <= item) lo = mid; else hi = mid - 1;
        }
        b = lo; i0 = (item - s_ioff[lo]) * UN2;
    };
    auto tile_live = [&](int, int) -> bool { return true; };   // every numbered tile is live

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < ACT_STAGES; ++s) { mbar_init(&act_full[s], 1); mbar_init(&act_empty[s], 1); }
        for (uint32_t s = 0; s < W_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(accA_full, 1); mbar_init(xt_full, (uint32_t)ea_warps);
        mbar_init(accB_full, 1); mbar_init(accB_empty, (uint32_t)eb_warps);
        fence_barrier_init();
    }
    constexpr int WARP_ACT = EPI_WARPS, WARP_W = EPI_WARPS + 1, WARP_MMA = EPI_WARPS + 2;
    if (warp == WARP_MMA) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    if (p.lens != nullptr)
        for (int i = threadIdx.x; i < p.batch; i += blockDim.x) {
            const long long lim = (__ldg(p.lens + i) + p.len_margin) * (long long)p.len_rate + p.len_extra;
            s_lim[i] = lim > 0x7fffffffLL ? 0x7fffffff : (int)lim;
        }
    if (trimming) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int i = 0; i < p.batch; ++i) {
                s_ioff[i] = acc;
                const int lim = s_lim[i] < 0 ? 0 : s_lim[i];
                int live = (lim + UN2 - 1) / UN2;
                acc += live > p.t_tiles ? p.t_tiles : live;
            }
            s_ioff[p.batch] = acc;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tps = p.tps;
    n_items = trimming ? s_ioff[p.batch] : p.total_tiles;
    // the prologue above overlapped the previous kernel's tail; the weight producer (constant data only) does not wait
    if (warp != EPI_WARPS + 1) { grid_dep_wait(); grid_dep_launch(); }

    if (warp == WARP_ACT) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_act);
            uint32_t s = 0, ph = 0;
            for (int item = blockIdx.x; item < n_items; item += ncta) {
                int i0, b;
                tile_i0(item, i0, b);
                if (!tile_live(i0, b)) continue;
                for (int c = 0; c < p.chunks; ++c) {
                    mbar_wait_producer(&act_empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&act_full[s], (uint32_t)(nbox * BOX_ROWS * ROWB));
                    for (int bx = 0; bx < nbox; ++bx)
                        tma_load_3d_hint(s_act + (size_t)s * ACT_BYTES + (size_t)bx * BOX_ROWS * ROWB, &tm_act, &act_full[s],
                                         c * CH, i0 - UXT_OFF + min_off + bx * BOX_ROWS, b, pol);
                    if (++s == ACT_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == WARP_W) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_w1);
            tma_prefetch_desc(&tm_w2);
            uint32_t s = 0, ph = 0;
            const size_t stage_bytes = (size_t)W_BYTES * tps;
            for (int item = blockIdx.x; item < n_items; item += ncta) {
                int i0, b;
                tile_i0(item, i0, b);
                if (!tile_live(i0, b)) continue;
                for (int phase = 0; phase < 2; ++phase) {
                    const CUtensorMap *tm = phase == 0 ? &tm_w1 : &tm_w2;
                    const int ntaps = phase == 0 ? p.taps : u.taps2;
                    for (int c = 0; c < p.chunks; ++c)
                        for (int j = 0; j < ntaps; j += tps) {
                            mbar_wait(&w_empty[s], ph ^ 1u);   // hot poll: the weight ring is short and on the critical path
                            mbar_arrive_expect_tx(&w_full[s], (uint32_t)(p.w_rows * ROWB * tps));
                            tma_load_3d(s_w + (size_t)s * stage_bytes, tm, &w_full[s], c * CH, 0, j);
                            if (++s == W_STAGES) { s = 0; ph ^= 1u; }
                        }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // whole warp converged; elect.sync inside the step picks the issuing lane
        constexpr uint32_t idescA = make_idesc_16(TM, TN, FMT), idescB = make_idesc_16(TM, UN2, FMT);
        const uint32_t wfull0 = smem_u32(w_full), wempty0 = smem_u32(w_empty);
        const uint64_t adesc_first = make_smem_desc(smem_u32(s_w), ROWB, 0);
        const uint64_t bdesc_first = make_smem_desc(smem_u32(s_act), ROWB, 0);
        const uint64_t xdesc_first = make_smem_desc(smem_u32(s_xt), ROWB, 0);
        constexpr uint64_t A_STAGE_STEP = (uint64_t)(W_BYTES >> 4), B_STAGE_STEP = (uint64_t)(ACT_BYTES >> 4),
                           X_CHUNK_STEP = (uint64_t)(XT_BYTES >> 4);
        const uint64_t a_stage_step = A_STAGE_STEP * (uint64_t)tps;
        const long long tap0 = (long long)(p.tap_off0 - min_off) * (ROWB >> 4);
        const long long tap_step = (long long)p.tap_step * (ROWB >> 4);
        const uint64_t xtap0 = (uint64_t)((UXT_OFF - h2) * (ROWB >> 4));   // conv2 tap 0 row, 16-byte units
        constexpr uint64_t XTAP_STEP = (uint64_t)(ROWB >> 4);
        uint32_t sa = 0, aph = 0, sw = 0, wph = 0, tl = 0;
        uint64_t adesc = adesc_first, bstage = bdesc_first;
        uint32_t w_ready = 0;
        auto step = [&](uint32_t tmem_d, uint64_t bdesc, uint64_t bdesc1, uint32_t idesc, uint32_t acc) {
            if (!w_ready) mbar_wait_addr(wfull0 + sw * 8u, wph);
            tc_fence_after();
            uint32_t sn = sw + 1, pn = wph;
            if (sn == W_STAGES) { sn = 0; pn ^= 1u; }
            if (tps == 2) {
                if (KSTEPS == 4) w_ready = umma_step4x2_warp(tmem_d, adesc, bdesc, idesc, acc, wfull0 + sn * 8u, pn, wempty0 + sw * 8u, bdesc1, A_STAGE_STEP);
                else w_ready = umma_step2x2_warp(tmem_d, adesc, bdesc, idesc, acc, wfull0 + sn * 8u, pn, wempty0 + sw * 8u, bdesc1, A_STAGE_STEP);
            } else {
                if (KSTEPS == 4) w_ready = umma_step4_warp(tmem_d, adesc, bdesc, idesc, acc, wfull0 + sn * 8u, pn, wempty0 + sw * 8u);
                else w_ready = umma_step2_warp(tmem_d, adesc, bdesc, idesc, acc, wfull0 + sn * 8u, pn, wempty0 + sw * 8u);
            }
            adesc += a_stage_step;
            sw = sn; wph = pn;
            if (sn == 0) adesc = adesc_first;
        };
        for (int item = blockIdx.x; item < n_items; item += ncta) {
            {
                int i0, b;
                tile_i0(item, i0, b);
                if (!tile_live(i0, b)) continue;
            }
            // ---- phase A: conv1 -> accumulator A (columns 0..255).  A is free: xt_full of the previous tile was
            // observed before its phase B was issued, i.e. every epilogue warp had finished reading A.
            if (lane == 0) VTTS_TRACE(0);
            uint32_t acc = 0;
            for (int c = 0; c < p.chunks; ++c) {
                mbar_wait(&act_full[sa], aph);
                uint64_t bdesc = bstage + (uint64_t)tap0;
                for (int j = 0; j < p.taps; j += tps) {
                    const uint64_t bdesc1 = (j + 1 < p.taps) ? bdesc + (uint64_t)tap_step : bdesc;
                    step(tmem_base, bdesc, bdesc1, idescA, acc);
                    acc = 1;
                    bdesc += (uint64_t)(tap_step * tps);
                }
                umma_commit_elect(&act_empty[sa]);
                bstage += B_STAGE_STEP;
                if (++sa == ACT_STAGES) { sa = 0; aph ^= 1u; bstage = bdesc_first; }
            }
            umma_commit_elect(accA_full);
            if (lane == 0) VTTS_TRACE(1);
            // ---- phase B: conv2 on the xt tile -> accumulator B (columns 256..495)
            mbar_wait(xt_full, tl & 1u);                      // epilogue wrote the operand tile (and drained A)
            if (lane == 0) VTTS_TRACE(2);
            mbar_wait(accB_empty, (tl & 1u) ^ 1u);            // previous tile's output epilogue drained B
            if (lane == 0) VTTS_TRACE(3);
            tc_fence_after();
            acc = 0;
            uint64_t xchunk = xdesc_first;
            for (int c = 0; c < p.chunks; ++c) {
                uint64_t xdesc = xchunk + xtap0;
                for (int j = 0; j < u.taps2; j += tps) {
                    const uint64_t xdesc1 = (j + 1 < u.taps2) ? xdesc + XTAP_STEP : xdesc;
                    step(tmem_base + TN, xdesc, xdesc1, idescB, acc);
                    acc = 1;
                    xdesc += XTAP_STEP * (uint64_t)tps;
                }
                xchunk += X_CHUNK_STEP;
            }
            umma_commit_elect(accB_full);
            if (lane == 0) VTTS_TRACE(4);
            ++tl;
        }
    } else {
        // ===== epilogue warps: two groups so that the operand epilogue of tile t+1 (accumulator A -> xt tile in
        // shared memory, warps 0-7) overlaps the output epilogue of tile t (accumulator B -> HBM, warps 8-15) =====
        const int ew = warp, quarter = warp & 3;
        const bool is_ea = ew < ea_warps;
        const int SHARERS = (is_ea ? ea_warps : eb_warps) / 4;       // warps per lane quarter within this group
        const int qpc = 4 / p.rep;
        const int worker = (quarter / qpc) * SHARERS + (is_ea ? ew : ew - ea_warps) / 4;   // slice of the 16-column groups
        const int n_workers = p.rep * SHARERS;
        const int ch = (quarter % qpc) * 32 + lane;                  // channel of this thread (m_blocks == 1)
        const bool row_ok = ch < p.n_total;
        const bool rows_full = (quarter % qpc) * 32 + 32 <= p.n_total;
        uint32_t tl = 0;
        if (is_ea) {
            // Accumulator A -> LeakyReLU'd 16-bit xt tile.  mma-style fragments (tcgen05.ld 16x256b: a thread holds two
            // consecutive positions of one channel) are packed in pairs and stored transposed with stmatrix, so one
            // instruction writes 8 positions x 32 channels as swizzled 16-byte chunks: the K-major rows conv2 reads.
            const int chbase = (quarter % qpc) * 32;                 // first channel of this warp's 32 TMEM lanes
            const int fr = lane >> 2, fc = (lane & 3) * 2;           // fragment row (channel) / first column (position)
            float b1[4];                                             // conv1 bias of channels chbase + 16*L + 8*h + fr
#pragma unroll
            for (int q = 0; q < 4; ++q) b1[q] = u.bias1 ? __ldg(u.bias1 + chbase + (q >> 1) * 16 + (q & 1) * 8 + fr) : 0.f;
            const uint32_t xt_row0 = smem_u32(s_xt) + (uint32_t)(chbase / CH) * XT_BYTES;
            const uint32_t cblk = (uint32_t)((chbase % CH) / 8 + (lane >> 3));   // 16-byte chunk of matrix lane/8
            const int mrow = lane & 7;                               // row of that matrix this thread addresses
            const uint32_t t_lo = tmem_base + ((uint32_t)(quarter * 32) << 16), t_hi = t_lo + (16u << 16);
            for (int item = blockIdx.x; item < n_items; item += ncta) {
                int i0, b;
                tile_i0(item, i0, b);
                if (!tile_live(i0, b)) continue;
                mbar_wait_relaxed(accA_full, tl & 1u);
                if (ew == 0 && lane == 0) VTTS_TRACE(5);
                tc_fence_after();
                const int pos0 = i0 - UXT_OFF;
                const bool interior = pos0 >= 0 && pos0 + TN <= p.n_pos;      // no conv2 zero padding inside this tile
                uint32_t lo[8], hi[8];
                int col = worker * 16;
                if (col < TN) { tmem_ld_16x256_x2(t_lo + (uint32_t)col, lo); tmem_ld_16x256_x2(t_hi + (uint32_t)col, hi); }
                for (; col < TN; col += n_workers * 16) {
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        // e: bit 2 = columns +8, bit 1 = lanes +8, bit 0 = column +1
                        v[e] = lrelu_max(__uint_as_float(lo[e]) + b1[(e >> 1) & 1], u.slope_mid);
                        v[8 + e] = lrelu_max(__uint_as_float(hi[e]) + b1[2 + ((e >> 1) & 1)], u.slope_mid);
                    }
                    const int col_n = col + n_workers * 16;
                    if (col_n < TN) { tmem_ld_16x256_x2(t_lo + (uint32_t)col_n, lo); tmem_ld_16x256_x2(t_hi + (uint32_t)col_n, hi); }
                    if (!interior) {                                  // conv2 zero-pads its own input
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int pos = pos0 + col + ((e >> 2) & 1) * 8 + fc + (e & 1);
                            if (pos < 0 || pos >= p.n_pos) v[e] = 0.f;
                        }
                    }
#pragma unroll
                    for (int s8 = 0; s8 < 2; ++s8) {                  // positions col + 8*s8 .. +7
                        const int r = col + s8 * 8 + mrow;
                        const uint32_t swz = ROWB == 128 ? (uint32_t)(r & 7) : (uint32_t)((r >> 1) & 3);
                        const uint32_t addr = xt_row0 + (uint32_t)r * ROWB + ((cblk ^ swz) << 4);
                        const int o = s8 * 4;
                        stmatrix_x4_trans(addr, cvt16x2(v[o], v[o + 1], FMT), cvt16x2(v[o + 2], v[o + 3], FMT),
                                          cvt16x2(v[8 + o], v[8 + o + 1], FMT), cvt16x2(v[8 + o + 2], v[8 + o + 3], FMT));
                    }
                }
                fence_proxy_async_smem();        // generic-proxy stores -> visible to the tensor-core (async) proxy
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(xt_full);
                if (ew == 0 && lane == 0) VTTS_TRACE(6);
                ++tl;
            }
        } else {
            const bool R = p.res != nullptr, Cc = p.accumulate != 0, D = p.divide_by > 0.f, X = p.out_x != nullptr,
                       A = p.out_a != nullptr;
            const int mode = (R && !Cc && !D && X && A) ? EPI_RXA : (R && !Cc && !D && X && !A) ? EPI_RX
                           : (R && Cc && !D && X && !A) ? EPI_RCX : (R && Cc && D && X && A) ? EPI_RCDXA
                           : (R && Cc && D && X && !A) ? EPI_RCDX : EPI_GENERIC;
            const int c_ct = (!A || p.out_a_ld == p.cout) ? p.cout : 0;
            const float bias2 = (row_ok && p.bias) ? __ldg(p.bias + ch) : 0.f;
            const int n_valid = p.n_pos < p.L_out ? p.n_pos : p.L_out;
            for (int item = blockIdx.x; item < n_items; item += ncta) {
                int i0, b;
                tile_i0(item, i0, b);
                if (!tile_live(i0, b)) continue;
                {   // L2 prefetch of the fp32 streams the next live tile reads (residual, running MRF sum): in the
                    // time-packed layout a tile's rows are one contiguous block of (positions / 4) * C * 16 bytes
                    int nx = item + ncta, i0n = 0, bn = 0;
                    for (; nx < n_items; nx += ncta) { tile_i0(nx, i0n, bn); if (tile_live(i0n, bn)) break; }
                    if (nx < n_items && (R || Cc)) {
                        const int rows4 = min(UN2 / 4, p.L4 - i0n / 4);
                        const long long off = ((long long)bn * p.L4 + i0n / 4) * p.cout * 4;
                        const int n_lines = rows4 * p.cout / 8;                 // 128-byte lines
                        for (int l = threadIdx.x - ea_warps * 32; l < n_lines; l += eb_warps * 32) {
                            if (R) prefetch_l2(p.res + off + (long long)l * 32);
                            if (Cc) prefetch_l2(p.out_x + off + (long long)l * 32);
                        }
                    }
                }
                // residual loads of this warp's first output group go out before waiting for the accumulator
                EpiLoads cur{}, nxt{};
                auto group_fast = [&](int ibase) { return mode != EPI_GENERIC && rows_full && ibase + 16 <= n_valid; };
                auto res_ptr = [&](int ibase) { return p.res + (((long long)b * p.L4 + (ibase >> 2)) * p.cout + ch) * 4; };
                if (R && worker * 16 < UN2 && group_fast(i0 + worker * 16)) epi_load16<0>(nxt, res_ptr(i0 + worker * 16), p.cout, pol);
                if (ew == ea_warps && lane == 0) VTTS_TRACE(7);
                mbar_wait_relaxed(accB_full, tl & 1u);
                if (ew == ea_warps && lane == 0) VTTS_TRACE(8);
                tc_fence_after();
                for (int col = worker * 16; col < UN2; col += n_workers * 16) {
                    const int ibase = i0 + col;
                    if (ibase >= p.n_pos) break;
                    uint32_t v[16];
                    tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(TN + col), v);
                    cur = nxt;
                    const bool fast = group_fast(ibase);
                    const int col_n = col + n_workers * 16;
                    if (R && col_n < UN2 && group_fast(i0 + col_n)) epi_load16<0>(nxt, res_ptr(i0 + col_n), p.cout, pol);
                    tmem_ld_wait();
                    if (fast) {
                        float *px = X ? p.out_x + (((long long)b * p.L4 + (ibase >> 2)) * p.cout + ch) * 4 : nullptr;
                        uint16_t *pa = A ? p.out_a + ((long long)b * p.L_out + ibase) * p.out_a_ld + ch : nullptr;
                        epi_group16_dispatch<FMT>(mode, c_ct, v, bias2, p, cur, px, pa);
                    } else {
                        epi_group16_edge<FMT>(v, bias2, p, row_ok, b, ibase, 0, ch);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(accB_empty);
                if (ew == ea_warps && lane == 0) VTTS_TRACE(9);
                ++tl;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tmem_base, 512);
}

struct TcUnitLaunch {
    CUtensorMap tm_act, tm_w1, tm_w2;
    TcUnitParams u;
    int rowb, fmt;
    bool pdl = false;
    dim3 grid;
    size_t smem;
};

template <int ROWB, int FMT>
static int unit_launch_t(const TcUnitLaunch &L, cudaStream_t st) {
    // the opt-in shared-memory size is a per-device function attribute: set it once per device, not once per process
    static bool attr[64] = {};
    int dev = 0;
    VTTS_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(unit_tc_kernel<ROWB, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    VTTS_CHECK_CUDA(launch_kernel_ex(unit_tc_kernel<ROWB, FMT>, L.grid, dim3(TC_THREADS), L.smem, st, L.pdl, 1u, L.tm_act, L.tm_w1, L.tm_w2, L.u));
    return VTTS_OK;
}

static bool tc_fuse_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("VTTS_TC_FUSE"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

// Can (conv1, conv2) of one unit run fused?  C in {32, 64, 128}, conv2 dilation 1 and half-width <= UXT_OFF.
static bool unit_fusable(int C, int k1, int d1, int k2) {
    if (!tc_fuse_enabled()) return false;
    if (C != 32 && C != 64 && C != 128) return false;
    if ((k1 - 1) * d1 > HALO_MAX || (k2 - 1) / 2 > UXT_OFF || k2 < 1) return false;
    return true;
}

// act: (B, L, C) 16-bit operand of conv1; w1/w2: packed [taps][128][C]; p: phase-B epilogue (res/out_x/out_a/...).
static int unit_prepare(TcUnitLaunch &L, int fmt, const uint16_t *act, int B, int Lpos, int C, const uint16_t *w1, int k1,
                        int d1, const float *bias1, float slope_mid, const uint16_t *w2, int k2, TcConvParams p) {
    const int rowb = (C % 64 == 0) ? 128 : 64;
    const int ch = rowb / 2;
    L.rowb = rowb; L.fmt = fmt;
    p.n_total = C; p.cout = C; p.L_out = Lpos; p.n_pos = Lpos; p.out_stride = 1; p.out_off0 = 0;
    p.taps = k1; p.tap_off0 = -(k1 - 1) / 2 * d1; p.tap_step = d1;
    p.chunks = C / ch;
    p.m_blocks = 1; p.cluster = 1; p.groups_per_batch = 0;
    p.t_tiles = ceil_div(Lpos, UN2);
    const long long total = (long long)p.t_tiles * B;
    if (total > 0x7fffffffLL) return set_error(VTTS_E_UNSUPPORTED, "tc: too many tiles");
    p.total_tiles = (int)total;
    p.inv_div = p.divide_by > 0.f ? 1.f / p.divide_by : 1.f;
    p.rep = (C == 32 || C == 64) ? TM / C : 1;
    p.w_rows = TM;
    p.epi_quarters = 4;
    p.L4 = (Lpos + 3) / 4;
    p.batch = B;
    if (p.lens && B > MAX_TRIM_BATCH) p.lens = nullptr;
    p.tps = p.chunks == 1 ? 2 : 1;
    if (rowb == 64) { p.act_stages = 4; p.w_stages = 4; }          // 80 + 64 + 16 KB
    else if (p.chunks == 1) { p.act_stages = 2; p.w_stages = 3; }  // 80 + 96 + 32 KB
    else { p.act_stages = 2; p.w_stages = 4; }                     // 80 + 64 + 64 KB
    L.u.e = p; L.u.bias1 = bias1; L.u.slope_mid = slope_mid; L.u.taps2 = k2; L.u.xt_chunks = p.chunks;
    {
        static int forced = -1;
        if (forced < 0) { const char *e = getenv("VTTS_UNIT_EA_WARPS"); forced = e ? atoi(e) : 0; }
        // measured (tools/exp_ea.sh): k=3 units and the accumulating last units of a block are bound by the output
        // epilogue (-9 % with 12 warps on it); the k=7 / k=11 units want 8 warps on the operand tile (critical path)
        L.u.ea_warps = (forced == 4 || forced == 8) ? forced : ((k1 + k2 <= 6 || p.accumulate) ? 4 : 8);
    }
    L.smem = (size_t)p.act_stages * ACT_ROWS * rowb + (size_t)p.w_stages * p.tps * TM * rowb + (size_t)p.chunks * TN * rowb +
             (size_t)(2 * p.act_stages + 2 * p.w_stages + 4) * 8 + 16 + (2 * MAX_TRIM_BATCH + 1) * sizeof(int);
    if (L.smem > 227 * 1024) return set_error(VTTS_E_UNSUPPORTED, "tc unit: %zu B shared memory", L.smem);
    const int sms = tc_num_sms();
    L.grid = dim3((unsigned)(p.total_tiles < sms ? p.total_tiles : sms));
    {
        uint64_t dims[3] = {(uint64_t)C, (uint64_t)Lpos, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)C * 2 * (uint64_t)Lpos};
        uint32_t box[3] = {(uint32_t)ch, BOX_ROWS, 1};
        int rc = make_tmap_bf16(&L.tm_act, act, 3, dims, str, box, rowb);
        if (rc) return rc;
    }
    for (int w = 0; w < 2; ++w) {
        uint64_t dims[3] = {(uint64_t)C, (uint64_t)TM, (uint64_t)(w == 0 ? k1 : k2)};
        uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)C * 2 * (uint64_t)TM};
        uint32_t box[3] = {(uint32_t)ch, TM, (uint32_t)p.tps};
        int rc = make_tmap_bf16(w == 0 ? &L.tm_w1 : &L.tm_w2, w == 0 ? w1 : w2, 3, dims, str, box, rowb);
        if (rc) return rc;
    }
    return VTTS_OK;
}

static int unit_launch(const TcUnitLaunch &L, cudaStream_t st) {
    if (L.rowb == 128) return L.fmt == VTTS_FMT_BF16 ? unit_launch_t<128, 0>(L, st) : unit_launch_t<128, 1>(L, st);
    return L.fmt == VTTS_FMT_BF16 ? unit_launch_t<64, 0>(L, st) : unit_launch_t<64, 1>(L, st);
}

// ---------------------------------------------------------------------------------------------
// Narrow fused unit (C = 32 or 64): the same conv1 -> LeakyReLU -> conv2 (+ residual epilogue) as unit_tc_kernel, laid
// out for layers whose 64 (or 32, replicated twice) output channels fill only half of a 128-row MMA:
//   * M = 64 MMAs (same cycles as M = 128, half the weight bytes); accumulator row r lives in TMEM lane 32*(r/16)+r%16,
//     so every warp quarter owns 16 channels and both epilogues work on 16-lane mma-style fragments;
//   * the weights of both convs stay resident in shared memory for the whole kernel (only conv1 taps that do not fit
//     are streamed through a small ring) - no per-tile weight traffic, so small tiles cost nothing extra;
//   * four accumulators of 128 columns (A0, A1, B0, B1): conv1 of tile t+1 is issued before conv2 of tile t, so the
//     tensor pipe keeps running while the operand epilogue builds the xt tile, and the output epilogue of tile t has
//     until conv2 of tile t+2 to drain.
// ---------------------------------------------------------------------------------------------
constexpr int VN_A = 128;          // conv1 positions per tile (UMMA N of phase A)
constexpr int VN_B = 112;          // conv2 output positions per tile (UMMA N of phase B)
constexpr int V_XT_OFF = 8;        // xt row 0 is position i0 - 8
constexpr int V_ACT_ROWS = 192;    // VN_A + dilation halo (<= 64)
constexpr int V_M = 64;            // accumulator rows
constexpr int V_BOX = 16;          // activation rows per TMA box (small boxes: the tile is fetched to the nearest 16 rows)
constexpr int V_MAXG = 4;          // residual prefetch slots per output-epilogue warp (16-column groups in flight)
constexpr int V_THREADS = (EPI_WARPS + 4) * 32;   // 16 epilogue warps, activation + weight producers, two MMA issuers

struct TcUnit64Params {
    TcConvParams e;                // phase-B epilogue + shared geometry (taps/tap_off0/tap_step describe conv1)
    const float *bias1;
    float slope_mid;
    int taps2;
    int n_stream;                  // conv1 taps [0, n_stream) go through the ring, everything else is resident
    int n_res;                     // resident taps: conv1 [n_stream, taps) then conv2 [0, taps2)
    int ea_warps;                  // epilogue warps building the operand tile (4 or 8); the other 16 - ea_warps drain B
    int xt_bufs;                   // xt tiles in shared memory: 2 (one per tile parity) or 1 (when the weights need the room)
};

struct EpiLoads2 { float4 r[2]; };

// two float4 (4 consecutive positions each, 8 positions apart) of one channel
template <int FMT, int C_CT, int MODE>
__device__ __forceinline__ void epi64_pair(const float (&q)[8], float bias, const TcConvParams &p, const EpiLoads2 &res,
                                           float *px, uint16_t *pa) {
    using F = EpiFlags<MODE>;
    const int C = C_CT ? C_CT : p.cout;
    const int lda = C_CT ? C_CT : p.out_a_ld;
    // running MRF sum: the middle blocks only add to it (vector reduction, nothing to wait for); the last block needs
    // the value (mean + 16-bit copy) and loads it - both loads before the first store
    constexpr bool RED = MODE == EPI_RCX;
    float4 accv[2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
        accv[m] = (F::ACC && !RED) ? *reinterpret_cast<const float4 *>(px + (size_t)m * 2 * C * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        float val[4];
        const float rr[4] = {res.r[m].x, res.r[m].y, res.r[m].z, res.r[m].w};
        const float aa[4] = {accv[m].x, accv[m].y, accv[m].z, accv[m].w};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            float x = q[m * 4 + d] + bias;
            if (F::RES) x = x + rr[d];
            if (F::ACC && !RED) x = aa[d] + x;
            if (F::DIV) x = x * p.inv_div;
            val[d] = x;
        }
        if (RED) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(px + (size_t)m * 2 * C * 4), "f"(val[0]),
                         "f"(val[1]), "f"(val[2]), "f"(val[3])
                         : "memory");
        } else if (F::X) {
            *reinterpret_cast<float4 *>(px + (size_t)m * 2 * C * 4) = make_float4(val[0], val[1], val[2], val[3]);
        }
        if (F::A) {
#pragma unroll
            for (int d = 0; d < 4; ++d) pa[(size_t)(m * 8 + d) * lda] = cvt16(lrelu_max(val[d], p.slope_out), FMT);
        }
    }
}
// per-element path (tile edges, unusual epilogue combinations)
template <int FMT>
__device__ __noinline__ void epi64_pair_edge(const float (&q)[8], float bias, const TcConvParams &p, int b, int pos0, int ch) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int t = pos0 + (e >> 2) * 8 + (e & 3);
        if (t < p.n_pos && t < p.L_out) {
            const long long xo = tp4_off(b, p.L4, t, p.cout, ch);
            float val = q[e] + bias;
            if (p.res) val = val + __ldg(p.res + xo);
            if (p.accumulate) val = p.out_x[xo] + val;
            if (p.divide_by > 0.f) val = val * p.inv_div;
            if (p.out_x) p.out_x[xo] = val;
            if (p.out_a) p.out_a[((long long)b * p.L_out + t) * p.out_a_ld + ch] = cvt16(lrelu_max(val, p.slope_out), FMT);
        }
    }
}

// MODE: the output epilogue variant (EPI_*), compile-time so that each instantiation carries one epilogue body - the
// instruction-cache footprint of the 20 concurrently running warps matters.
template <int ROWB, int FMT, int MODE>
__global__ void __launch_bounds__(V_THREADS, 1)
unit64_tc_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w1,
                 const __grid_constant__ CUtensorMap tm_w2, const TcUnit64Params u) {
    const TcConvParams &p = u.e;
    constexpr int CH = ROWB / 2;                  // channels per operand row == C (one K chunk)
    constexpr int KSTEPS = CH / 16;
    constexpr int ACT_BYTES = V_ACT_ROWS * ROWB;
    constexpr int TAPB = V_M * ROWB;              // one tap of weights: 64 rows
    constexpr int XT_BYTES = VN_A * ROWB;
    constexpr uint32_t COL_A1 = 128, COL_B0 = 256, COL_B1 = 384;
    extern __shared__ __align__(1024) uint8_t smem[];
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) { printf("vtts: smem base not 1024-byte aligned\n"); __trap(); }
    const uint32_t ACT_STAGES = (uint32_t)p.act_stages, W_STAGES = (uint32_t)p.w_stages;
    uint8_t *s_act = smem;
    uint8_t *s_wres = s_act + (size_t)ACT_STAGES * ACT_BYTES;
    uint8_t *s_wring = s_wres + (size_t)u.n_res * TAPB;
    uint8_t *s_xt = s_wring + (size_t)W_STAGES * TAPB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_xt + (size_t)u.xt_bufs * XT_BYTES);     // one or two xt buffers
    uint64_t *act_full = bars, *act_empty = act_full + ACT_STAGES;
    uint64_t *w_full = act_empty + ACT_STAGES, *w_empty = w_full + W_STAGES;
    uint64_t *wres_full = w_empty + W_STAGES;
    // per buffer (tile parity): accA_full  conv1 MMAs done          -> operand epilogue may read A[b]
    //                            xt_full    operand epilogue done    -> conv2 may read xt[b]; conv1 may overwrite A[b]
    //                            accB_full  conv2 MMAs done          -> output epilogue may read B[b]; xt[b] may be rewritten
    //                            accB_empty output epilogue done     -> conv2 may overwrite B[b]
    uint64_t *accA_full = wres_full + 1, *xt_full = accA_full + 2, *accB_full = xt_full + 2, *accB_empty = accB_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accB_empty + 2);
    int *s_lim = reinterpret_cast<int *>(tmem_slot + 2);
    int *s_ioff = s_lim + MAX_TRIM_BATCH;          // [MAX_TRIM_BATCH + 1] first live tile index of every batch row
    volatile int *s_a_issued = s_ioff + MAX_TRIM_BATCH + 1;   // tiles whose conv1 MMAs have been issued (issue-order hand-off)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t pol = l2_policy(p.stream_hint != 0);   // eviction policy of the activation / residual reads
    const int last_off = p.tap_off0 + (p.taps - 1) * p.tap_step;
    const int min_off = p.tap_off0 < last_off ? p.tap_off0 : last_off;
    const int span = (p.tap_off0 < last_off ? last_off : p.tap_off0) - min_off;
    const int nbox = (VN_A + span + V_BOX - 1) / V_BOX;
    const int ncta = (int)gridDim.x;
    const int ea_warps = u.ea_warps, eb_warps = EPI_WARPS - u.ea_warps;   // operand / output epilogue warps (4 + 12 or 8 + 8)

    // work items: (time tile of VN_B outputs, batch).  With the padding trim only LIVE tiles are numbered (per-batch
    // offsets in s_ioff), so the round-robin over CTAs is balanced; every role walks the same list incrementally - no
    // division on the MMA issuers' path.
    const bool trimming = p.lens != nullptr;
    const bool rev = p.reverse != 0;              // walk the numbered tiles last-to-first
    int n_items = 0;                              // set once the per-batch offsets exist (below)
    struct Walk { int item, t, b; };              // item counts this CTA's steps through the list; (t, b) = tile, batch
    auto walk_begin = [&]() {
        Walk w; w.item = (int)blockIdx.x; w.t = 0; w.b = 0;
        if (w.item >= n_items) return w;
        const int li = rev ? n_items - 1 - w.item : w.item;
        if (trimming) {
            int lo = 0, hi = p.batch - 1;                              // last b with s_ioff[b] <= li
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_ioff[mid] <= li) lo = mid; else hi = mid - 1; }
            w.b = lo; w.t = li - s_ioff[lo];
        } else { w.b = li / p.t_tiles; w.t = li - w.b * p.t_tiles; }
        return w;
    };
    auto walk_next = [&](Walk &w) {
        w.item += ncta;
        if (w.item >= n_items) return;
        const int li = rev ? n_items - 1 - w.item : w.item;
        if (trimming) {
            if (!rev) { while (li >= s_ioff[w.b + 1]) ++w.b; } else { while (li < s_ioff[w.b]) --w.b; }
            w.t = li - s_ioff[w.b];
        } else if (p.t_tiles >= ncta) {                                // incremental: at most one wrap per step
            if (!rev) { w.t += ncta; if (w.t >= p.t_tiles) { w.t -= p.t_tiles; ++w.b; } }
            else { w.t -= ncta; if (w.t < 0) { w.t += p.t_tiles; --w.b; } }
        } else { w.b = li / p.t_tiles; w.t = li - w.b * p.t_tiles; }
    };

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < ACT_STAGES; ++s) { mbar_init(&act_full[s], 1); mbar_init(&act_empty[s], 1); }
        for (uint32_t s = 0; s < W_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(wres_full, 1);
        *s_a_issued = 0;
        mbar_init(&accA_full[0], 1); mbar_init(&accA_full[1], 1);
        mbar_init(&xt_full[0], (uint32_t)ea_warps); mbar_init(&xt_full[1], (uint32_t)ea_warps);
        mbar_init(&accB_full[0], 1); mbar_init(&accB_full[1], 1);
        mbar_init(&accB_empty[0], (uint32_t)eb_warps); mbar_init(&accB_empty[1], (uint32_t)eb_warps);
        fence_barrier_init();
    }
    constexpr int WARP_ACT = EPI_WARPS, WARP_W = EPI_WARPS + 1, WARP_MMA = EPI_WARPS + 2, WARP_MMA_B = EPI_WARPS + 3;
    if (warp == WARP_MMA) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    if (p.lens != nullptr)
        for (int i = threadIdx.x; i < p.batch; i += blockDim.x) {
            const long long lim = (__ldg(p.lens + i) + p.len_margin) * (long long)p.len_rate + p.len_extra;
            s_lim[i] = lim > 0x7fffffffLL ? 0x7fffffff : (int)lim;
        }
    if (trimming) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int i = 0; i < p.batch; ++i) {
                s_ioff[i] = acc;
                const int lim = s_lim[i] < 0 ? 0 : s_lim[i];
                int live = (lim + VN_B - 1) / VN_B;
                acc += live > p.t_tiles ? p.t_tiles : live;
            }
            s_ioff[p.batch] = acc;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    n_items = trimming ? s_ioff[p.batch] : p.total_tiles;
    // programmatic dependent launch: the prologue above - and the resident weight loads below, which do not depend on
    // the previous kernel - overlap that kernel's tail; everything else waits for it here
    if (warp != WARP_W) { grid_dep_wait(); grid_dep_launch(); }

    if (warp == WARP_ACT) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_act);
            uint32_t s = 0, ph = 0;
            for (Walk w = walk_begin(); w.item < n_items; walk_next(w)) {
                const int i0 = w.t * VN_B, b = w.b;
                {   // the fp32 streams the output epilogue of this tile will read (residual, running MRF sum) are one
                    // contiguous block in the time-packed layout: pull it into L2 now, several tiles ahead of its use
                    const int rows4 = min(VN_B / 4, p.L4 - i0 / 4);
                    const long long off = ((long long)b * p.L4 + i0 / 4) * p.cout * 4;
                    const uint32_t bytes = (uint32_t)rows4 * (uint32_t)p.cout * 16u;
                    if (p.res) bulk_prefetch_l2(p.res + off, bytes);
                    if (p.accumulate && p.divide_by > 0.f) bulk_prefetch_l2(p.out_x + off, bytes);
                }
                mbar_wait_producer(&act_empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&act_full[s], (uint32_t)(nbox * V_BOX * ROWB));
                for (int bx = 0; bx < nbox; ++bx)
                    tma_load_3d_hint(s_act + (size_t)s * ACT_BYTES + (size_t)bx * V_BOX * ROWB, &tm_act, &act_full[s], 0,
                                     i0 - V_XT_OFF + min_off + bx * V_BOX, b, pol);
                if (++s == ACT_STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == WARP_W) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_w1);
            tma_prefetch_desc(&tm_w2);
            // resident weights, loaded once: conv1 taps [n_stream, taps), then all conv2 taps
            mbar_arrive_expect_tx(wres_full, (uint32_t)(u.n_res * TAPB));
            int slot = 0;
            for (int j = u.n_stream; j < p.taps; ++j, ++slot) tma_load_3d(s_wres + (size_t)slot * TAPB, &tm_w1, wres_full, 0, 0, j);
            for (int j = 0; j < u.taps2; ++j, ++slot) tma_load_3d(s_wres + (size_t)slot * TAPB, &tm_w2, wres_full, 0, 0, j);
            // (no grid-dependency wait in this warp: weights are constant data)
            if (u.n_stream > 0) {
                uint32_t s = 0, ph = 0;
                for (Walk w = walk_begin(); w.item < n_items; walk_next(w))
                    for (int j = 0; j < u.n_stream; ++j) {
                        mbar_wait(&w_empty[s], ph ^ 1u);   // hot poll: the weight ring is short and on the critical path
                        mbar_arrive_expect_tx(&w_full[s], (uint32_t)TAPB);
                        tma_load_3d(s_wring + (size_t)s * TAPB, &tm_w1, &w_full[s], 0, 0, j);
                        if (++s == W_STAGES) { s = 0; ph ^= 1u; }
                    }
            }
        }
    } else if (warp == WARP_MMA || warp == WARP_MMA_B) {
        // Two MMA issuers share the tensor pipe: one issues conv1 of every tile, the other conv2.  With 64-cycle MMAs
        // the per-tile barrier probes, commits and tile bookkeeping of a single issuer (well over 1000 cycles) would
        // starve the pipe; split like this, one issuer's overhead hides behind the other's MMAs.
        if (lane == 0) {
            constexpr uint32_t idescA = make_idesc_16(V_M, VN_A, FMT), idescB = make_idesc_16(V_M, VN_B, FMT);
            constexpr uint64_t ROW16 = (uint64_t)(ROWB >> 4);                // one operand row, in 16-byte units
            constexpr uint64_t TAP16 = (uint64_t)(TAPB >> 4), ACT16 = (uint64_t)(ACT_BYTES >> 4);
            const uint64_t wres_desc = make_smem_desc(smem_u32(s_wres), ROWB, 0);
            mbar_wait(wres_full, 0);
            tc_fence_after();
            if (warp == WARP_MMA) {
                const uint64_t wring_desc = make_smem_desc(smem_u32(s_wring), ROWB, 0);
                const uint64_t act_desc = make_smem_desc(smem_u32(s_act), ROWB, 0);
                const long long tap0 = (long long)(p.tap_off0 - min_off) * (long long)ROW16;
                const uint64_t tap_step = (uint64_t)((long long)p.tap_step * (long long)ROW16);
                const uint64_t w1res_desc = wres_desc - (uint64_t)u.n_stream * TAP16;   // resident conv1 tap j at + j * TAP16
                const int n_stream = u.n_stream, taps1 = p.taps;
                uint32_t sa = 0, aph = 0, sw = 0, wph = 0, tl = 0;
                for (Walk w = walk_begin(); w.item < n_items; walk_next(w), ++tl) {
                    VTTS_TRACE(0);
                    // A[tl & 1] is free once the operand epilogue of tile tl-2 has read it
                    mbar_wait(&xt_full[tl & 1u], ((tl >> 1) & 1u) ^ 1u);
                    mbar_wait(&act_full[sa], aph);
                    VTTS_TRACE(10);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + ((tl & 1u) ? COL_A1 : 0u);
                    uint64_t bdesc = act_desc + (uint64_t)sa * ACT16 + (uint64_t)tap0;
                    uint32_t acc = 0;
                    // Streamed taps [0, n_stream) come through the weight ring, then the resident ones.  (Interleaving the
                    // two kinds to spread the ring's refills measured slower.)
                    // The issuer is latency critical (64-cycle MMAs): descriptors advance by adds only.
                    uint64_t a_res = w1res_desc + (uint64_t)n_stream * TAP16;            // resident tap n_stream
                    uint64_t b_res = bdesc + (uint64_t)n_stream * tap_step, b_str = bdesc;
                    int js = 0, jr = n_stream;
                    bool ready = n_stream > 0 && mbar_try_wait(&w_full[sw], wph);
                    while (js < n_stream) {
                        if (!ready) mbar_wait(&w_full[sw], wph);
                        tc_fence_after();
                        uint32_t sn = sw + 1, pn = wph;
                        if (sn == W_STAGES) { sn = 0; pn ^= 1u; }
                        const bool ready_next = mbar_try_wait(&w_full[sn], pn);   // probe early, read after the issue
                        const uint64_t a_str = wring_desc + (uint64_t)sw * TAP16;
#pragma unroll
                        for (int ks = 0; ks < KSTEPS; ++ks) {
                            umma_bf16(tmem_d, a_str + (uint64_t)(ks * 2), b_str + (uint64_t)(ks * 2), idescA, acc);
                            acc = 1;
                        }
                        umma_commit(&w_empty[sw]);
                        sw = sn; wph = pn;
                        ready = ready_next;
                        ++js;
                        b_str += tap_step;
                    }
                    for (; jr < taps1; ++jr) {                        // remaining (or all) resident taps
#pragma unroll
                        for (int ks = 0; ks < KSTEPS; ++ks) {
                            umma_bf16(tmem_d, a_res + (uint64_t)(ks * 2), b_res + (uint64_t)(ks * 2), idescA, acc);
                            acc = 1;
                        }
                        a_res += TAP16;
                        b_res += tap_step;
                    }
                    VTTS_TRACE(11);
                    umma_commit(&act_empty[sa]);
                    if (++sa == ACT_STAGES) { sa = 0; aph ^= 1u; }
                    umma_commit(&accA_full[tl & 1u]);
                    *s_a_issued = (int)tl + 1;
                    VTTS_TRACE(1);
                }
            } else {
                const int h2 = (u.taps2 - 1) / 2;
                const uint64_t xt_desc = make_smem_desc(smem_u32(s_xt), ROWB, 0) + (uint64_t)(V_XT_OFF - h2) * ROW16;
                constexpr uint64_t XT16 = (uint64_t)(XT_BYTES >> 4);
                const uint64_t w2_desc = wres_desc + (uint64_t)(p.taps - u.n_stream) * TAP16;
                const int taps2 = u.taps2;
                uint32_t tl = 0;
                for (Walk w = walk_begin(); w.item < n_items; walk_next(w), ++tl) {
                    mbar_wait(&xt_full[tl & 1u], (tl >> 1) & 1u);     // operand epilogue wrote xt[tl & 1]
                    VTTS_TRACE(2);
                    mbar_wait(&accB_empty[tl & 1u], ((tl >> 1) & 1u) ^ 1u);   // output epilogue of tile tl-2 drained B[tl & 1]
                    if (ACT_STAGES == 1) {
                        // Single activation stage (weights fill the rest of shared memory): the stage reloads only after
                        // conv1 of tile tl+1 has EXECUTED, so that conv1 goes into the MMA queue before this conv2 instead
                        // of interleaved with it - otherwise both finish together and the pipe idles for the reload.
                        Walk wn = w;
                        walk_next(wn);
                        if (wn.item < n_items) {
                            const long long t0 = clock64();
                            while (*s_a_issued < (int)tl + 2)
                                if (clock64() - t0 > 4000000000LL) mbar_timeout();
                        }
                    }
                    VTTS_TRACE(3);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + ((tl & 1u) ? COL_B1 : COL_B0);
                    uint64_t adesc = w2_desc, bdesc = xt_desc + (((tl & 1u) && u.xt_bufs == 2) ? XT16 : 0);
                    uint32_t acc = 0;
                    for (int j = 0; j < taps2; ++j) {
#pragma unroll
                        for (int ks = 0; ks < KSTEPS; ++ks) {
                            umma_bf16(tmem_d, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idescB, acc);
                            acc = 1;
                        }
                        adesc += TAP16;
                        bdesc += ROW16;
                    }
                    umma_commit(&accB_full[tl & 1u]);
                    VTTS_TRACE(4);
                }
            }
        }
    } else {
        // ===== epilogue warps.  Quarter q of the TMEM lanes holds accumulator rows 16q .. 16q+15 in its first 16 lanes;
        // row r is channel r % C (C = 32: two copies of every channel -> two warps per channel share the columns).
        const int ew = warp, quarter = warp & 3;
        const bool is_ea = ew < ea_warps;
        const int sharers = (is_ea ? ea_warps : eb_warps) / 4;        // warps of this group per lane quarter
        const int chb = (quarter * 16) % p.cout;                      // first channel of this warp's 16 rows
        const int copy = (quarter * 16) / p.cout;
        const int n_workers = (V_M / p.cout) * sharers;               // warps sharing one channel set
        const int worker = copy * sharers + (is_ea ? ew : ew - ea_warps) / 4;
        const int fr = lane >> 2, fc = (lane & 3) * 2;                // fragment row (channel) / first column (position)
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint32_t tl = 0;
        if (is_ea) {
            float b1[2];
            b1[0] = u.bias1 ? __ldg(u.bias1 + chb + fr) : 0.f;
            b1[1] = u.bias1 ? __ldg(u.bias1 + chb + fr + 8) : 0.f;
            const uint32_t xt_base = smem_u32(s_xt);
            const uint32_t cblk = (uint32_t)(chb / 8 + ((lane >> 3) & 1));   // 16-byte chunk written by matrix lane/8
            const int mrow = (lane & 7) + 8 * (lane >> 4);                   // row (position) this thread addresses
            for (Walk w = walk_begin(); w.item < n_items; walk_next(w), ++tl) {
                const int i0 = w.t * VN_B;
                mbar_wait_relaxed(&accA_full[tl & 1u], (tl >> 1) & 1u);
                if (ew == 0 && lane == 0) VTTS_TRACE(5);
                if (u.xt_bufs == 2) {
                    mbar_wait_relaxed(&accB_full[tl & 1u], ((tl >> 1) & 1u) ^ 1u);   // conv2 of tile tl-2 finished reading xt[tl & 1]
                } else if (tl > 0) {                              // single xt tile: conv2 of the PREVIOUS tile must be done with it
                    mbar_wait_relaxed(&accB_full[(tl - 1u) & 1u], ((tl - 1u) >> 1) & 1u);
                }
                tc_fence_after();
                const uint32_t t_acc = t_lane + ((tl & 1u) ? COL_A1 : 0u);
                const uint32_t xt0 = xt_base + (((tl & 1u) && u.xt_bufs == 2) ? (uint32_t)XT_BYTES : 0u);
                const int pos0 = i0 - V_XT_OFF;
                const bool interior = pos0 >= 0 && pos0 + VN_A <= p.n_pos;
                uint32_t r[8];
                int col = worker * 16;
                if (col < VN_A) tmem_ld_16x256_x2(t_acc + (uint32_t)col, r);
                for (; col < VN_A; col += n_workers * 16) {
                    tmem_ld_wait();
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = lrelu_max(__uint_as_float(r[e]) + b1[(e >> 1) & 1], u.slope_mid);
                    const int col_n = col + n_workers * 16;
                    if (col_n < VN_A) tmem_ld_16x256_x2(t_acc + (uint32_t)col_n, r);
                    if (!interior) {                                  // conv2 zero-pads its own input
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int pos = pos0 + col + ((e >> 2) & 1) * 8 + fc + (e & 1);
                            if (pos < 0 || pos >= p.n_pos) v[e] = 0.f;
                        }
                    }
                    const int row = col + mrow;
                    const uint32_t swz = ROWB == 128 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
                    stmatrix_x4_trans(xt0 + (uint32_t)row * ROWB + ((cblk ^ swz) << 4), cvt16x2(v[0], v[1], FMT),
                                      cvt16x2(v[2], v[3], FMT), cvt16x2(v[4], v[5], FMT), cvt16x2(v[6], v[7], FMT));
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&xt_full[tl & 1u]);
                if (ew == 0 && lane == 0) VTTS_TRACE(6);
            }
        } else {
            using F = EpiFlags<MODE>;
            constexpr bool R = MODE != EPI_GENERIC && F::RES, X = F::X, A = F::A;   // GENERIC: per-element path only
            const int odd = lane & 1;
            const int ch = chb + fr + 8 * odd;                        // after the pair exchange a thread owns one channel
            const int pg = (lane & 3) >> 1;                           // which group of 4 positions inside 8 columns
            const float bias2 = p.bias ? __ldg(p.bias + ch) : 0.f;
            const int n_valid = p.n_pos < p.L_out ? p.n_pos : p.L_out;
            auto group_fast = [&](int ibase) { return MODE != EPI_GENERIC && ibase + 16 <= n_valid; };
            auto load2 = [&](EpiLoads2 &d, const float *ptr) {
                d.r[0] = ldg_f4_hint(ptr, pol);
                d.r[1] = ldg_f4_hint(ptr + (size_t)2 * p.cout * 4, pol);
            };
            const int col0 = worker * 16, cstep = n_workers * 16;
            // residual of the 16-column group at `col` of the tile at (i0, b): 4 positions + the 4 positions 8 further
            auto res_ptr = [&](int i0, int b, int col) {
                return p.res + (((long long)b * p.L4 + (((i0 + col) >> 2) + pg)) * p.cout + ch) * 4;
            };
            // 32 channels (at most two groups per warp): the residual of a tile is loaded ONE TILE AHEAD - slot g is refilled
            // for the next tile right after group g has consumed it, so the L2/HBM latency of these loads never sits
            // between the accumulator becoming ready and the stores (this epilogue is the pace-maker of the HBM-bound
            // units).  64 channels (up to four groups per warp): that many loads in flight next to the stores measured
            // slower, so there the tile's loads go out just before the accumulator wait.  Compile-time split, so neither
            // variant pays for the other.
            constexpr bool EARLY = ROWB == 64;
            EpiLoads2 rl[V_MAXG];
            Walk w = walk_begin();
#pragma unroll
            for (int g = 0; g < V_MAXG; ++g) {
                const int col = col0 + g * cstep;
                rl[g] = EpiLoads2{};
                if (EARLY && R && w.item < n_items && col < VN_B && group_fast(w.t * VN_B + col))
                    load2(rl[g], res_ptr(w.t * VN_B, w.b, col));
            }
            for (; w.item < n_items; ++tl) {
                const int i0 = w.t * VN_B, b = w.b;
                Walk wn = w;
                walk_next(wn);
                const bool has_next = wn.item < n_items;
                const int i0n = wn.t * VN_B, bn = wn.b;
                if (!EARLY) {
#pragma unroll
                    for (int g = 0; g < V_MAXG; ++g) {
                        const int col = col0 + g * cstep;
                        rl[g] = EpiLoads2{};
                        if (R && col < VN_B && group_fast(i0 + col)) load2(rl[g], res_ptr(i0, b, col));
                    }
                }
                if (ew == ea_warps && lane == 0) VTTS_TRACE(7);
                mbar_wait_relaxed(&accB_full[tl & 1u], (tl >> 1) & 1u);
                if (ew == ea_warps && lane == 0) VTTS_TRACE(8);
                tc_fence_after();
                const uint32_t t_acc = t_lane + ((tl & 1u) ? COL_B1 : COL_B0);
                uint32_t r[8];
                if (col0 < VN_B) tmem_ld_16x256_x2(t_acc + (uint32_t)col0, r);
#pragma unroll
                for (int g = 0; g < V_MAXG; ++g) {
                    const int col = col0 + g * cstep;
                    if (col >= VN_B) break;
                    const int ibase = i0 + col;
                    tmem_ld_wait();
                    // pair exchange: even lanes keep row fr and take the partner's two columns, odd lanes keep row
                    // fr + 8; afterwards a thread holds 4 consecutive positions (twice, 8 apart) of one channel
                    const uint32_t s0 = odd ? r[0] : r[2], s1 = odd ? r[1] : r[3], s2 = odd ? r[4] : r[6], s3 = odd ? r[5] : r[7];
                    const uint32_t x0 = __shfl_xor_sync(0xffffffffu, s0, 1), y0 = __shfl_xor_sync(0xffffffffu, s1, 1);
                    const uint32_t x1 = __shfl_xor_sync(0xffffffffu, s2, 1), y1 = __shfl_xor_sync(0xffffffffu, s3, 1);
                    float q[8];
                    if (odd) {
                        q[0] = __uint_as_float(x0); q[1] = __uint_as_float(y0); q[2] = __uint_as_float(r[2]); q[3] = __uint_as_float(r[3]);
                        q[4] = __uint_as_float(x1); q[5] = __uint_as_float(y1); q[6] = __uint_as_float(r[6]); q[7] = __uint_as_float(r[7]);
                    } else {
                        q[0] = __uint_as_float(r[0]); q[1] = __uint_as_float(r[1]); q[2] = __uint_as_float(x0); q[3] = __uint_as_float(y0);
                        q[4] = __uint_as_float(r[4]); q[5] = __uint_as_float(r[5]); q[6] = __uint_as_float(x1); q[7] = __uint_as_float(y1);
                    }
                    if (col + cstep < VN_B) tmem_ld_16x256_x2(t_acc + (uint32_t)(col + cstep), r);   // next group's accumulators
                    if (ibase < p.n_pos) {
                        const int pos = ibase + pg * 4;               // first of this thread's 4 positions
                        if (group_fast(ibase)) {
                            float *px = X ? p.out_x + (((long long)b * p.L4 + (pos >> 2)) * p.cout + ch) * 4 : nullptr;
                            uint16_t *pa = A ? p.out_a + ((long long)b * p.L_out + pos) * p.out_a_ld + ch : nullptr;
                            epi64_pair<FMT, CH, MODE == EPI_GENERIC ? EPI_RX : MODE>(q, bias2, p, rl[g], px, pa);
                        } else {
                            epi64_pair_edge<FMT>(q, bias2, p, b, pos, ch);
                        }
                    }
                    if (EARLY && R && has_next && group_fast(i0n + col)) load2(rl[g], res_ptr(i0n, bn, col));   // next tile
                }
                w = wn;
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&accB_empty[tl & 1u]);
                if (ew == ea_warps && lane == 0) VTTS_TRACE(9);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tmem_base, 512);
}

struct TcUnit64Launch {
    CUtensorMap tm_act, tm_w1, tm_w2;
    TcUnit64Params u;
    int rowb, fmt, mode;
    bool pdl = false;
    dim3 grid;
    size_t smem;
};

template <int ROWB, int FMT, int MODE>
static int unit64_launch_m(const TcUnit64Launch &L, cudaStream_t st) {
    // the opt-in shared-memory size is a per-device function attribute: set it once per device, not once per process
    static bool attr[64] = {};
    int dev = 0;
    VTTS_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(unit64_tc_kernel<ROWB, FMT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    VTTS_CHECK_CUDA(launch_kernel_ex(unit64_tc_kernel<ROWB, FMT, MODE>, L.grid, dim3(V_THREADS), L.smem, st, L.pdl, 1u, L.tm_act, L.tm_w1, L.tm_w2,
                                     L.u));
    return VTTS_OK;
}
template <int ROWB, int FMT>
static int unit64_launch_t(const TcUnit64Launch &L, cudaStream_t st) {
    switch (L.mode) {
        case EPI_RXA: return unit64_launch_m<ROWB, FMT, EPI_RXA>(L, st);
        case EPI_RX: return unit64_launch_m<ROWB, FMT, EPI_RX>(L, st);
        case EPI_RCX: return unit64_launch_m<ROWB, FMT, EPI_RCX>(L, st);
        case EPI_RCDXA: return unit64_launch_m<ROWB, FMT, EPI_RCDXA>(L, st);
        case EPI_RCDX: return unit64_launch_m<ROWB, FMT, EPI_RCDX>(L, st);
        default: return unit64_launch_m<ROWB, FMT, EPI_GENERIC>(L, st);
    }
}
static int unit64_launch(const TcUnit64Launch &L, cudaStream_t st) {
    if (L.rowb == 128) return L.fmt == VTTS_FMT_BF16 ? unit64_launch_t<128, 0>(L, st) : unit64_launch_t<128, 1>(L, st);
    return L.fmt == VTTS_FMT_BF16 ? unit64_launch_t<64, 0>(L, st) : unit64_launch_t<64, 1>(L, st);
}

static bool tc_unit64_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("VTTS_TC_UNIT64"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

// shared-memory plan of the narrow unit: returns false when conv2 cannot be fully resident
static bool unit64_plan(int C, int k1, int k2, int &act_stages, int &w_stages, int &n_stream, int &n_res, int &xt_bufs,
                        size_t &smem) {
    const int rowb = C * 2;
    const size_t tapb = (size_t)V_M * rowb, actb = (size_t)V_ACT_ROWS * rowb, xtb = (size_t)VN_A * rowb;
    const size_t fixed = 512 + (2 * MAX_TRIM_BATCH + 2) * sizeof(int);
    const size_t avail = 227 * 1024;
    const int total = k1 + k2;
    w_stages = 0; n_stream = 0; n_res = total;
    // 1. everything resident, one xt tile per parity, 4..2 activation stages
    xt_bufs = 2;
    for (act_stages = 4; act_stages >= 2; --act_stages) {
        smem = fixed + 2 * xtb + act_stages * actb + (size_t)total * tapb;
        if (smem <= avail) return true;
    }
    // 2. everything resident with a single xt tile and 2..1 activation stages (64 channels x k=11: 176 KB of weights).
    //    The two MMA issuers keep the pipe busy while the single stage reloads, and no weight ring has to keep up.
    xt_bufs = 1;
    for (act_stages = 2; act_stages >= 1; --act_stages) {
        smem = fixed + xtb + act_stages * actb + (size_t)total * tapb;
        if (smem <= avail) return true;
    }
    // 3. conv2 resident, the conv1 taps that do not fit streamed through a ring
    xt_bufs = 2; act_stages = 2; w_stages = 4;
    const size_t base = fixed + 2 * xtb + act_stages * actb + w_stages * tapb;
    if (base + (size_t)k2 * tapb > avail) return false;
    n_res = (int)((avail - base) / tapb);
    if (n_res > total) n_res = total;
    n_stream = total - n_res;
    if (n_stream > k1) return false;
    smem = base + (size_t)n_res * tapb;
    return true;
}

static bool unit64_usable(int C, int k1, int d1, int k2) {
    if (!tc_unit64_enabled() || (C != 32 && C != 64)) return false;
    if ((k1 - 1) * d1 > HALO_MAX || (k2 - 1) / 2 > V_XT_OFF || k2 < 1) return false;
    int a, w, ns, nr, xb;
    size_t smem;
    return unit64_plan(C, k1, k2, a, w, ns, nr, xb, smem);
}

static int unit64_prepare(TcUnit64Launch &L, int fmt, const uint16_t *act, int B, int Lpos, int C, const uint16_t *w1, int k1,
                          int d1, const float *bias1, float slope_mid, const uint16_t *w2, int k2, TcConvParams p) {
    const int rowb = C * 2;
    L.rowb = rowb; L.fmt = fmt;
    p.n_total = C; p.cout = C; p.L_out = Lpos; p.n_pos = Lpos; p.out_stride = 1; p.out_off0 = 0;
    p.taps = k1; p.tap_off0 = -(k1 - 1) / 2 * d1; p.tap_step = d1;
    p.chunks = 1; p.m_blocks = 1; p.cluster = 1; p.groups_per_batch = 0;
    p.t_tiles = ceil_div(Lpos, VN_B);
    const long long total = (long long)p.t_tiles * B;
    if (total > 0x7fffffffLL) return set_error(VTTS_E_UNSUPPORTED, "tc: too many tiles");
    p.total_tiles = (int)total;
    p.inv_div = p.divide_by > 0.f ? 1.f / p.divide_by : 1.f;
    p.rep = V_M / C; p.w_rows = V_M; p.epi_quarters = 4; p.tps = 1;
    p.L4 = (Lpos + 3) / 4;
    p.batch = B;
    if (p.lens && B > MAX_TRIM_BATCH) p.lens = nullptr;
    int act_stages, w_stages, n_stream, n_res, xt_bufs;
    if (!unit64_plan(C, k1, k2, act_stages, w_stages, n_stream, n_res, xt_bufs, L.smem))
        return set_error(VTTS_E_UNSUPPORTED, "tc unit64: weights of conv2 do not fit in shared memory");
    p.act_stages = act_stages; p.w_stages = w_stages;
    L.u.e = p; L.u.bias1 = bias1; L.u.slope_mid = slope_mid; L.u.taps2 = k2; L.u.n_stream = n_stream; L.u.n_res = n_res;
    L.u.xt_bufs = xt_bufs;
    {
        static int forced = -1;
        if (forced < 0) { const char *e = getenv("VTTS_UNIT64_EA_WARPS"); forced = e ? atoi(e) : 0; }
        // measured (tools/exp_ea.sh): 12 output-epilogue warps pay off where that side is the bottleneck (64 ch: k=7 and
        // the accumulating last unit of a block; 32 ch: k=11), 8 + 8 elsewhere
        const bool wide_out = (C == 64 && (k1 == 7 || p.accumulate)) || (C == 32 && k1 == 11);
        L.u.ea_warps = (forced == 4 || forced == 8) ? forced : (wide_out ? 4 : 8);
    }
    {
        const bool R = p.res != nullptr, Cc = p.accumulate != 0, D = p.divide_by > 0.f, X = p.out_x != nullptr, A = p.out_a != nullptr;
        L.mode = (R && !Cc && !D && X && A) ? EPI_RXA : (R && !Cc && !D && X && !A) ? EPI_RX : (R && Cc && !D && X && !A) ? EPI_RCX
               : (R && Cc && D && X && A) ? EPI_RCDXA : (R && Cc && D && X && !A) ? EPI_RCDX : EPI_GENERIC;
        if (A && p.out_a_ld != C) L.mode = EPI_GENERIC;        // the fast epilogue assumes a dense 16-bit copy
    }
    const int sms = tc_num_sms();
    L.grid = dim3((unsigned)(p.total_tiles < sms ? p.total_tiles : sms));
    {
        uint64_t dims[3] = {(uint64_t)C, (uint64_t)Lpos, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)C * 2 * (uint64_t)Lpos};
        uint32_t box[3] = {(uint32_t)C, (uint32_t)V_BOX, 1};
        int rc = make_tmap_bf16(&L.tm_act, act, 3, dims, str, box, rowb);
        if (rc) return rc;
    }
    for (int w = 0; w < 2; ++w) {
        uint64_t dims[3] = {(uint64_t)C, (uint64_t)TM, (uint64_t)(w == 0 ? k1 : k2)};
        uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)C * 2 * (uint64_t)TM};
        uint32_t box[3] = {(uint32_t)C, (uint32_t)V_M, 1};
        int rc = make_tmap_bf16(w == 0 ? &L.tm_w1 : &L.tm_w2, w == 0 ? w1 : w2, 3, dims, str, box, rowb);
        if (rc) return rc;
    }
    return VTTS_OK;
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 reference layout -> bf16 [tap][n_pad][ci_pad]
// ---------------------------------------------------------------------------------------------
// Conv1d (cout,cin,k): n = co, tap j = kernel index.
// ConvTranspose1d (cin,cout,k), stride s: n = q*cout + co, tap j reads x[i0 - j], weight index q + j*s.
// qperm (stride and padding multiples of 4): n = ((q / 4) * cout + co) * 4 + q % 4 instead -- the four output samples of one
// 16-byte slot of the time-packed fp32 stream then live in four neighbouring accumulator lanes (epi_group16_poly_quad).
__global__ void pack_tc_kernel(const float *__restrict__ w, uint16_t *__restrict__ out, int fmt, int kind, int cin,
                               int cout, int k, int s, int taps, int n_pad, int ci_pad, int n_rows_real, int qperm) {
    const size_t total = (size_t)taps * n_pad * ci_pad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int ci = (int)(idx % ci_pad);
        const size_t rest = idx / ci_pad;
        int n = (int)(rest % n_pad);
        const int j = (int)(rest / n_pad);
        if (n_rows_real == 32 || n_rows_real == 64) n %= n_rows_real;   // replicate narrow layers over all 128 lanes
        float v = 0.f;
        if (ci < cin) {
            if (kind == 0) {
                if (n < cout) v = w[((size_t)n * cin + ci) * k + j];
            } else {
                int q = n / cout, co = n - q * cout;
                if (qperm) { const int g = n >> 2, qh = g / cout; co = g - qh * cout; q = qh * 4 + (n & 3); }
                if (q < s) v = w[((size_t)ci * cout + co) * k + q + j * s];
            }
        }
        out[idx] = cvt16(v, fmt);
    }
}

// conv_post on the time-packed fp32 stream ([b][t/4][c][4], see the epilogue helpers):
// y[b, oc, t] = tanh(bias + sum_k sum_ci w[k][ci] * lrelu(x[b, t+k-h, ci]))
// (generator.py:108-120; kept in fp32 -- the output conv dominates the 16-bit error).
// HBM-bound: reads C*4 bytes per sample once (neighbouring taps come from shared memory), writes 4.
template <int C>
__global__ void __launch_bounds__(256)
conv_post_tp4_kernel(const float *__restrict__ x, const float *__restrict__ w /* [k][C] */, const float *__restrict__ bias,
                     float *__restrict__ y, int L, int L4, int ksize, float slope, int out_channels, int oc,
                     const long long *__restrict__ lens, int len_margin, int len_rate) {
    extern __shared__ float s_tile[];  // [C][256 + 8 + 4]: time t0-4 .. t0+259 per channel
    const int b = blockIdx.y, t0 = blockIdx.x * 256, h = (ksize - 1) / 2;   // h <= 4
    grid_dep_wait();
    grid_dep_launch();
    const long long t_lim = lens != nullptr ? (lens[b] + len_margin) * (long long)len_rate : (long long)L;
    if ((long long)t0 >= t_lim) {
        // trimmed padding: defined (zero) output, no work
        const int t = t0 + threadIdx.x;
        if (t < L) y[((size_t)b * out_channels + oc) * L + t] = 0.f;
        return;
    }
    constexpr int ROWS = 256 + 8, PITCH = ROWS + 4;   // channel-major tile: float4 rows, conflict-free 128-bit stores
    const float *xb = x + (size_t)b * L4 * C * 4;
    // memory order: (t/4, c, t%4); tile covers t in [t0-4, t0+260): one 128-bit load per (t/4, c)
    for (int m = threadIdx.x; m < (ROWS / 4) * C; m += 256) {
        const int c = m % C, r4 = m / C;
        const int tb = t0 - 4 + r4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tb >= 0 && tb < L) {
            v = __ldg(reinterpret_cast<const float4 *>(xb + ((size_t)(tb >> 2) * C + c) * 4));
            v.x = lrelu(v.x, slope);
            v.y = tb + 1 < L ? lrelu(v.y, slope) : 0.f;
            v.z = tb + 2 < L ? lrelu(v.z, slope) : 0.f;
            v.w = tb + 3 < L ? lrelu(v.w, slope) : 0.f;
        }
        *reinterpret_cast<float4 *>(s_tile + c * PITCH + r4 * 4) = v;
    }
    __syncthreads();
    // 64 threads x 4 consecutive outputs: per channel three 128-bit shared loads (t-4 .. t+7) feed 4 x ksize FMAs
    // (one load per FMA made this kernel shared-memory bound); the other threads only helped with the tile load
    if (threadIdx.x >= 64) return;
    const int tl4 = threadIdx.x * 4;                         // first output of this thread inside the tile
    if (t0 + tl4 >= L) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) {
        const float4 *col = reinterpret_cast<const float4 *>(s_tile + c * PITCH + tl4);   // tile index 0 is t0-4
        const float4 a = col[0], bq = col[1], cq = col[2];
        const float v[12] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w, cq.x, cq.y, cq.z, cq.w};   // t-4 .. t+7
#pragma unroll
        for (int j = -4; j <= 4; ++j) {                      // tap k = j + h reads t + j: static register indices
            const int k = j + h;
            if (k >= 0 && k < ksize) {
                const float wk = __ldg(w + k * C + c);
#pragma unroll
                for (int o = 0; o < 4; ++o) acc[o] = fmaf(wk, v[o + 4 + j], acc[o]);
            }
        }
    }
    const float bv = bias ? __ldg(bias + oc) : 0.f;
    float *yo = y + ((size_t)b * out_channels + oc) * L + t0 + tl4;
#pragma unroll
    for (int o = 0; o < 4; ++o)
        if (t0 + tl4 + o < L) yo[o] = (long long)(t0 + tl4 + o) < t_lim ? tanhf(acc[o] + bv) : 0.f;   // padding: defined zeros
}

// debug / test layout converters for the time-packed fp32 stream
__global__ void tp4_to_cf_kernel(const float *__restrict__ x, float *__restrict__ y, int C, int L, int L4) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < L) y[((size_t)b * C + c) * L + t] = x[(((size_t)b * L4 + (t >> 2)) * C + c) * 4 + (t & 3)];
}
__global__ void cf_to_tp4_kernel(const float *__restrict__ x, float *__restrict__ y, int C, int L, int L4) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < L) y[(((size_t)b * L4 + (t >> 2)) * C + c) * 4 + (t & 3)] = x[((size_t)b * C + c) * L + t];
}
static int launch_tp4_to_cf(const float *x, float *y, int B, int C, int L, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(L, 256), (unsigned)C, (unsigned)B);
    tp4_to_cf_kernel<<<grid, 256, 0, st>>>(x, y, C, L, (L + 3) / 4);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}
static int launch_cf_to_tp4(const float *x, float *y, int B, int C, int L, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(L, 256), (unsigned)C, (unsigned)B);
    cf_to_tp4_kernel<<<grid, 256, 0, st>>>(x, y, C, L, (L + 3) / 4);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

}  // namespace tc
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("VTTS_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

// ---------------------------------------------------------------------------------------------
// handle integration
// ---------------------------------------------------------------------------------------------
using namespace tc;

static int pad_to(int v, int m) { return (v + m - 1) / m * m; }
static int ci_pad_of(int cin) { return cin % 64 == 0 ? cin : (cin == 32 ? 32 : pad_to(cin, 64)); }

// output conv weights (cout, cin, k) -> [oc][k][ci] fp32 for conv_post_cl_kernel
__global__ void pack_post_kernel(const float *__restrict__ w, float *__restrict__ out, int cout, int cin, int k) {
    const int n = cout * cin * k;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int ci = idx % cin, kk = (idx / cin) % k, oc = idx / (cin * k);
        out[idx] = w[((size_t)oc * cin + ci) * k + kk];
    }
}

int tc_pack_layer(VttsGen *h, int layer, cudaStream_t st) {
    Layer &l = h->layers[layer];
    const int cin = l.info.cin, cout = l.info.cout, k = l.info.ksize;
    const bool transposed = l.info.kind == 1;
    const int s = transposed ? l.stride : 1;
    const int taps = transposed ? k / s : k;
    l.ci_pad = ci_pad_of(cin);
    l.n_total = transposed ? s * cout : cout;
    // quad row order: needs t % 4 == q % 4 (stride and padding multiples of 4) and whole channel octets per warp
    static int quad_on = -1;
    if (quad_on < 0) { const char *e = getenv("VTTS_TC_QUAD"); quad_on = (e && e[0] == '0') ? 0 : 1; }
    l.qperm = (quad_on && transposed && s % 4 == 0 && l.padding % 4 == 0 && cout % 8 == 0 && (s * cout) % 128 == 0) ? 1 : 0;
    const int n_pad = pad_to(l.n_total, TM);
    const size_t n = (size_t)taps * n_pad * l.ci_pad;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 8192) blocks = 8192;
    for (int fmt = 0; fmt < 2; ++fmt) {
        if (!l.w16[fmt]) VTTS_CHECK_CUDA(cudaMalloc(&l.w16[fmt], n * sizeof(uint16_t)));
        pack_tc_kernel<<<blocks, 256, 0, st>>>(l.w_fold, l.w16[fmt], fmt, l.info.kind, cin, cout, k, s, taps, n_pad, l.ci_pad, l.n_total, l.qperm);
        VTTS_CHECK_LAUNCH();
    }
    if (layer == h->idx_post) {
        if (!l.w_aux) VTTS_CHECK_CUDA(cudaMalloc(&l.w_aux, (size_t)cin * cout * k * sizeof(float)));
        pack_post_kernel<<<ceil_div(cin * cout * k, 256), 256, 0, st>>>(l.w_fold, l.w_aux, cout, cin, k);
        VTTS_CHECK_LAUNCH();
        l.w_aux_host.resize((size_t)cin * cout * k);   // conv_post_cl_kernel takes the weights as kernel parameters
        VTTS_CHECK_CUDA(cudaMemcpyAsync(l.w_aux_host.data(), l.w_aux, l.w_aux_host.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
        VTTS_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    return VTTS_OK;
}

void tc_destroy(VttsGen *h) {
    for (auto &cw : h->chain) chain_free(cw);
    h->chain.clear();
}



int tc_supported(const VttsGen *h, char *why, size_t why_len) {
    const VttsGenConfig &c = h->cfg;
    auto fail = [&](const char *msg) { if (why) snprintf(why, why_len, "%s", msg); return 0; };
    int ch = c.channels;
    for (int i = 0; i < c.num_upsamples; ++i) {
        ch /= 2;
        if (ch != 32 && ch % 64 != 0) return fail("bf16 path needs every stage width to be 32 or a multiple of 64 channels");
        if ((c.upsample_kernel_sizes[i] / c.upsample_scales[i]) - 1 > HALO_MAX) return fail("upsample taps too wide");
    }
    if (c.channels % 64 != 0) return fail("bf16 path needs channels to be a multiple of 64");
    for (int j = 0; j < c.num_blocks; ++j)
        for (int m = 0; m < c.num_dilations[j]; ++m)
            if ((c.resblock_kernel_sizes[j] - 1) * c.resblock_dilations[j][m] > HALO_MAX)
                return fail("bf16 path needs (kernel-1)*dilation <= 64");
    if (c.kernel_size - 1 > HALO_MAX) return fail("input conv too wide");
    if (c.kernel_size > 9) return fail("16-bit path: output conv kernel size must be <= 9");
    const int cl = c.channels >> c.num_upsamples;
    if (cl != 32 && cl != 64 && cl != 128) return fail("bf16 path: last stage width must be 32/64/128 for the fp32 output conv");
    return 1;
}

namespace {
struct TcPlan {
    std::vector<int> C, L;
    size_t max_rows_c = 0;  // max over stages of L*C (per batch)
};
static int convT_len(int L, int s, int k, int p, int op) { return (L - 1) * s - 2 * p + k + op; }
static TcPlan tc_plan(const VttsGen *h, int T) {
    TcPlan p;
    int ch = h->cfg.channels, L = T;
    p.max_rows_c = (size_t)ch * ((L + 3) / 4 * 4);
    for (int i = 0; i < h->cfg.num_upsamples; ++i) {
        const Layer &u = h->layers[h->idx_up[i]];
        L = convT_len(L, u.stride, u.info.ksize, u.padding, u.output_padding);
        ch /= 2;
        p.C.push_back(ch); p.L.push_back(L);
        size_t e = (size_t)ch * (size_t)(L > 0 ? (L + 3) / 4 * 4 : 0);
        if (e > p.max_rows_c) p.max_rows_c = e;
    }
    return p;
}
struct TcBuffers {
    uint16_t *a_in;   // input operand (B, T, ci_pad)
    uint16_t *a_u, *a_p, *a_q, *a_t, *a_c;  // 16-bit operand copies (max stage size)
    float *x_u, *x_p, *x_q, *x_cs;                // fp32 residual stream
    float *gb;                                    // (B, channels) global conditioning bias
    float *tmp;                                   // fp32 staging for dumps / conditioning input
    size_t total;
};
static TcBuffers tc_carve(const VttsGen *h, int B, int T, void *ws) {
    TcPlan pl = tc_plan(h, T);
    const size_t elems = (size_t)B * pl.max_rows_c;
    const size_t in_pad = (size_t)B * T * ci_pad_of(h->cfg.in_channels);
    char *p = (char *)ws;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    TcBuffers b{};
    b.a_in = (uint16_t *)take(in_pad * 2);
    b.a_u = (uint16_t *)take(elems * 2);
    b.a_p = (uint16_t *)take(elems * 2);
    b.a_q = (uint16_t *)take(elems * 2);
    b.a_t = (uint16_t *)take(elems * 2);
    b.a_c = (uint16_t *)take(elems * 2);
    b.x_u = (float *)take(elems * 4);
    b.x_p = (float *)take(elems * 4);
    b.x_q = (float *)take(elems * 4);
    b.x_cs = (float *)take(elems * 4);
    b.gb = (float *)take((size_t)B * h->cfg.channels * 4);
    b.tmp = (float *)take(elems * 4);
    b.total = off;
    return b;
}
}  // namespace

int tc_workspace_bytes(const VttsGen *h, int B, int T, size_t *bytes) {
    char why[128];
    if (!tc_supported(h, why, sizeof(why))) return set_error(VTTS_E_UNSUPPORTED, "%s", why);
    *bytes = tc_carve(h, B, T, nullptr).total;
    return VTTS_OK;
}

namespace {
// VTTS_PROFILE=1: per-launch CUDA-event timing of the tcgen05 path, printed to stderr after the forward
struct ProfRec { cudaEvent_t a, b; int kind, cin, cout, k, d, B, L; bool trimmed; };
static std::vector<ProfRec> g_prof;
static bool prof_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("VTTS_PROFILE"); on = (e && e[0] == '1') ? 1 : 0; }
    return on == 1;
}
// VTTS_STAGE_PROFILE=1: CUDA events at the stage boundaries only (programmatic dependent launch stays on, unlike the
// per-launch VTTS_PROFILE): where the forward's time goes in the pipelined run; printed to stderr after the forward
struct StageProf {
    bool on = false;
    std::vector<std::pair<std::string, cudaEvent_t>> ev;
    void mark(const char *name, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.emplace_back(name, e);
    }
    void report() {
        if (!on || ev.size() < 2) return;
        cudaEventSynchronize(ev.back().second);
        float total = 0.f;
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0.f; cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
            fprintf(stderr, "[vtts-stage] %-22s %8.3f ms\n", ev[i].first.c_str(), ms);
            total += ms;
        }
        fprintf(stderr, "[vtts-stage] %-22s %8.3f ms\n", "total", total);
        for (auto &e : ev) cudaEventDestroy(e.second);
        ev.clear();
    }
};
static bool stage_prof_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("VTTS_STAGE_PROFILE"); on = (e && e[0] == '1') ? 1 : 0; }
    return on == 1;
}
static void prof_report() {
    if (g_prof.empty()) return;
    cudaDeviceSynchronize();
    double total = 0, tflop = 0;
    for (auto &r : g_prof) {
        float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
        const double fl = 2.0 * r.cin * r.cout * r.k * (double)r.B * r.L;   // conv: per output step; convT: per input step
        // (with the padding trim active the launch computes fewer positions than B * L: TFLOP/s is then an upper bound)
        fprintf(stderr, "[vtts-prof] kind=%d cin=%4d cout=%4d k=%2d d=%d L=%7d  %8.3f ms  %7.1f TFLOP/s%s\n", r.kind, r.cin,
                r.cout, r.k, r.d, r.L, ms, fl / (ms * 1e-3) / 1e12, r.trimmed ? " (full-length FLOPs / trimmed launch)" : "");
        total += ms; tflop += fl;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    fprintf(stderr, "[vtts-prof] total %.3f ms over %zu conv launches, %.1f TFLOP/s\n", total, g_prof.size(),
            tflop / (total * 1e-3) / 1e12);
    g_prof.clear();
}
// one conv layer on the tensor cores
static int run_conv(VttsGen *h, int fmt, const Layer &l, const uint16_t *act, int B, int L_in, int L_out, TcConvParams p,
                    cudaStream_t st, int in_rate = 0, int margin = -1) {
    const bool transposed = l.info.kind == 1;
    const int s = transposed ? l.stride : 1;
    p.bias = l.has_bias ? l.bias : nullptr;
    p.n_total = l.n_total;
    p.qperm = l.qperm;
    p.cout = l.info.cout;
    p.L_out = L_out;
    if (transposed) {
        p.taps = l.info.ksize / s; p.tap_off0 = 0; p.tap_step = -1;
        p.out_stride = s; p.out_off0 = -l.padding; p.n_pos = L_in + p.taps - 1;
    } else {
        p.taps = l.info.ksize; p.tap_off0 = -(l.info.ksize - 1) / 2 * l.info.dilation; p.tap_step = l.info.dilation;
        p.out_stride = 1; p.out_off0 = 0; p.n_pos = L_in;
    }
    // padding trim: positions of this GEMM run at `in_rate` positions per mel frame
    if (h->trim_lens && in_rate > 0) {
        p.lens = (const long long *)h->trim_lens;
        p.len_margin = margin >= 0 ? margin : h->trim_margin;
        p.len_rate = in_rate;
        p.len_extra = transposed ? p.taps : 0;
    }
    {
        static int alt = -1;
        if (alt < 0) { const char *e = getenv("VTTS_TC_ALTERNATE"); alt = (e && e[0] == '0') ? 0 : 1; }
        p.reverse = alt ? (h->launch_count & 1) : 0;
        static int hint = -1;
        if (hint < 0) { const char *e = getenv("VTTS_TC_STREAM_HINT"); hint = (e && e[0] == '0') ? 0 : 1; }
        p.stream_hint = hint;
    }
    if (g_trace_on >= 100 && h->launch_count == g_trace_on - 100) p.trace = 1;   // debug: pipeline trace of this launch
    TcLaunch L;
    int rc = tc_prepare(L, fmt, act, B, L_in, l.ci_pad, l.w16[fmt], pad_to(l.n_total, TM), p);
    if (rc) return rc;
    L.pdl = pdl_enabled() && !prof_enabled();
    ProfRec pr{};
    if (prof_enabled()) {
        cudaEventCreate(&pr.a); cudaEventCreate(&pr.b);
        pr.kind = l.info.kind; pr.cin = l.info.cin; pr.cout = l.info.cout; pr.k = l.info.ksize; pr.d = l.info.dilation;
        pr.B = B; pr.L = transposed ? L_in : L_out; pr.trimmed = p.lens != nullptr;
        cudaEventRecord(pr.a, st);
    }
    if ((rc = tc_launch(L, st))) return rc;
    if (prof_enabled()) { cudaEventRecord(pr.b, st); g_prof.push_back(pr); }
    h->launch_count++;
    return VTTS_OK;
}
// fused (conv1, conv2) unit of a ResidualBlock on the tensor cores
static int run_unit(VttsGen *h, int fmt, const Layer &l1, const Layer &l2, const uint16_t *act, int B, int Lpos,
                    TcConvParams p, float slope_mid, cudaStream_t st, int rate, int margin = -1) {
    p.bias = l2.has_bias ? l2.bias : nullptr;
    if (h->trim_lens && rate > 0) {
        p.lens = (const long long *)h->trim_lens;
        p.len_margin = margin >= 0 ? margin : h->trim_margin;
        p.len_rate = rate;
        p.len_extra = 0;
    }
    {
        static int alt = -1;
        if (alt < 0) { const char *e = getenv("VTTS_TC_ALTERNATE"); alt = (e && e[0] == '0') ? 0 : 1; }
        p.reverse = alt ? (h->launch_count & 1) : 0;
        static int hint = -1;
        if (hint < 0) { const char *e = getenv("VTTS_TC_STREAM_HINT"); hint = (e && e[0] == '0') ? 0 : 1; }
        p.stream_hint = hint;
    }
    const bool narrow = unit64_usable(l1.info.cout, l1.info.ksize, l1.info.dilation, l2.info.ksize);
    TcUnitLaunch L;
    TcUnit64Launch L64;
    int rc = narrow ? unit64_prepare(L64, fmt, act, B, Lpos, l1.info.cout, l1.w16[fmt], l1.info.ksize, l1.info.dilation,
                                     l1.has_bias ? l1.bias : nullptr, slope_mid, l2.w16[fmt], l2.info.ksize, p)
                    : unit_prepare(L, fmt, act, B, Lpos, l1.info.cout, l1.w16[fmt], l1.info.ksize, l1.info.dilation,
                                   l1.has_bias ? l1.bias : nullptr, slope_mid, l2.w16[fmt], l2.info.ksize, p);
    if (rc) return rc;
    ProfRec pr{};
    if (prof_enabled()) {
        cudaEventCreate(&pr.a); cudaEventCreate(&pr.b);
        pr.kind = 2; pr.cin = l1.info.cin; pr.cout = l1.info.cout; pr.k = l1.info.ksize + l2.info.ksize; pr.d = l1.info.dilation;
        pr.B = B; pr.L = Lpos; pr.trimmed = p.lens != nullptr;
        cudaEventRecord(pr.a, st);
    }
    L.pdl = L64.pdl = pdl_enabled() && !prof_enabled();
    if (g_trace_on >= 100 && h->launch_count == g_trace_on - 100) { L.u.e.trace = 1; L64.u.e.trace = 1; }   // debug: trace this launch
    if ((rc = narrow ? unit64_launch(L64, st) : unit_launch(L, st))) return rc;
    if (prof_enabled()) { cudaEventRecord(pr.b, st); g_prof.push_back(pr); }
    h->launch_count++;
    return VTTS_OK;
}
}  // namespace

int tc_forward(VttsGen *h, int fmt, const float *c, const float *g, float *wav, int B, int T, void *workspace,
               size_t workspace_bytes, int dump_stage, float *dump_out, cudaStream_t st) {
    const VttsGenConfig &cfg = h->cfg;
    char why[128];
    if (!tc_supported(h, why, sizeof(why))) return set_error(VTTS_E_UNSUPPORTED, "%s", why);
    TcBuffers bf = tc_carve(h, B, T, workspace);
    if (workspace_bytes < bf.total) return set_error(VTTS_E_WORKSPACE, "vtts_gen_forward: workspace %zu < %zu", workspace_bytes, bf.total);
    TcPlan plan = tc_plan(h, T);
    int rc;
    // padding trim is only exact when every stage length is T * (product of scales)
    int rate = 1;
    {
        int r = 1;
        bool exact = true;
        for (int i = 0; i < cfg.num_upsamples; ++i) { r *= cfg.upsample_scales[i]; exact = exact && plan.L[i] == T * r; }
        if (!exact) h->trim_lens = nullptr;
    }
    const long long *trim_lens = (const long long *)h->trim_lens;
    // Stages whose every ResidualBlock fits the fused chain kernel (chain_tc.cu) run ONE launch per block on a
    // channels-last fp32 stream: the upsample writes x (B, L, C) fp32 only, the blocks combine into the MRF sum.
    auto chain_spec_of = [&](int C, int j) {
        ChainSpec s;
        s.C = C; s.k = cfg.resblock_kernel_sizes[j]; s.n_units = cfg.num_dilations[j]; s.has2 = cfg.use_additional_convs ? 1 : 0;
        for (int m = 0; m < 3 && m < cfg.num_dilations[j]; ++m) s.dil[m] = cfg.resblock_dilations[j][m];
        return s;
    };
    bool stage_chain[VTTS_MAX_STAGES];
    for (int i = 0; i < cfg.num_upsamples; ++i) {
        stage_chain[i] = true;
        for (int j = 0; j < cfg.num_blocks; ++j)
            if (cfg.num_dilations[j] > 3 || !chain_spec_usable(chain_spec_of(plan.C[i], j))) stage_chain[i] = false;
    }
    if (h->chain_dirty) {
        h->chain.resize((size_t)cfg.num_upsamples * cfg.num_blocks);
        for (int i = 0; i < cfg.num_upsamples; ++i) {
            if (!stage_chain[i]) continue;
            for (int j = 0; j < cfg.num_blocks; ++j) {
                const float *wp[CH_MAX_CONVS], *bp[CH_MAX_CONVS];
                int n = 0;
                for (int m = 0; m < cfg.num_dilations[j]; ++m) {
                    const Layer &c1 = h->layers[h->idx_c1[i][j][m]];
                    wp[n] = c1.w_fold; bp[n] = c1.has_bias ? c1.bias : nullptr; ++n;
                    if (cfg.use_additional_convs) {
                        const Layer &c2 = h->layers[h->idx_c2[i][j][m]];
                        wp[n] = c2.w_fold; bp[n] = c2.has_bias ? c2.bias : nullptr; ++n;
                    }
                }
                if ((rc = chain_pack_raw(chain_spec_of(plan.C[i], j), wp, bp, h->chain[(size_t)i * cfg.num_blocks + j], st))) return rc;
            }
        }
        h->chain_dirty = false;
    }
    // Per-layer trim margins (mel frames beyond mel_len that a layer still computes).  Walking the generator backwards
    // from the last valid sample: a layer's output is needed E positions past the valid end, its input therefore
    // E + (taps reach) past it.  The caller's margin (vtts_gen_set_valid_lengths) is a LOWER bound (extra safety frames),
    // never a cap: a configuration with a longer look-ahead than V1 must not stop early and read stale workspace.
    int mg_pre = h->trim_margin, mg_post = h->trim_margin, mg_up[VTTS_MAX_STAGES], mg_mrf[VTTS_MAX_STAGES];
    for (int i = 0; i < VTTS_MAX_STAGES; ++i) mg_up[i] = mg_mrf[i] = h->trim_margin;
    if (trim_lens) {
        auto cap = [&](long long frames) { return (int)(frames > h->trim_margin ? frames : h->trim_margin); };
        auto frames_of = [](long long ext, long long r) { return (ext + r) / r + 1; };   // ceil((ext + 1) / r) + 1 spare frame
        long long R[VTTS_MAX_STAGES + 1];
        R[0] = 1;
        for (int i = 0; i < cfg.num_upsamples; ++i) R[i + 1] = R[i] * cfg.upsample_scales[i];
        mg_post = 0;                                                 // output samples beyond mel_len are never needed
        long long E = (h->layers[h->idx_post].info.ksize - 1) / 2;  // conv_post reads this far past the end
        for (int i = cfg.num_upsamples - 1; i >= 0; --i) {
            long long reach = 0;                                     // deepest look-ahead of any ResidualBlock of the stage
            for (int j = 0; j < cfg.num_blocks; ++j) {
                long long rj = 0;
                for (int m = 0; m < cfg.num_dilations[j]; ++m) {
                    const Layer &c1 = h->layers[h->idx_c1[i][j][m]];
                    rj += (long long)(c1.info.ksize - 1) / 2 * c1.info.dilation;
                    if (cfg.use_additional_convs) rj += (h->layers[h->idx_c2[i][j][m]].info.ksize - 1) / 2;
                }
                if (stage_chain[i]) rj += chain_extra_reach(chain_spec_of(plan.C[i], j));   // zero-weight taps of the packed MMAs
                reach = rj > reach ? rj : reach;
            }
            const long long E_in = E + reach;                        // every unit of the stage is computed this far
            mg_mrf[i] = cap(frames_of(E_in, R[i + 1]));
            const Layer &u = h->layers[h->idx_up[i]];
            const long long in_ext = (E_in + u.padding) / u.stride + 1;   // upsample input positions past the end
            mg_up[i] = cap(frames_of(in_ext, R[i]));
            E = in_ext;
        }
        mg_pre = cap(E + 2);
    }
    auto dump_f32 = [&](int id, const float *src_cl, int C, int L) -> int {
        if (dump_stage != id || !dump_out) return VTTS_OK;
        h->launch_count++;
        return launch_tp4_to_cf(src_cl, dump_out, B, C, L, st);
    };

    // global conditioning bias (tiny 1x1 conv, fp32 CUDA cores)
    const float *bias_b = nullptr;
    if (g) {
        VTTS_REQUIRE(h->idx_global >= 0, "vtts_gen_forward: g given but global_channels <= 0");
        const Layer &gl = h->layers[h->idx_global];
        ConvFp32Params p{};
        p.x = g; p.w = gl.w_f32; p.bias = gl.has_bias ? gl.bias : nullptr; p.y = bf.gb;
        p.B = B; p.cin = gl.info.cin; p.cout = gl.info.cout; p.L_in = 1; p.L_out = 1; p.taps = 1;
        p.tap_off0 = 0; p.tap_step = 1; p.phases = 1; p.out_stride = 1; p.out_off0 = 0; p.n_pos = 1; p.slope_in = 1.f;
        if ((rc = launch_conv_fp32(p, st))) return rc;
        h->launch_count++;
        bias_b = bf.gb;
    }
    StageProf sp;
    sp.on = stage_prof_enabled();
    sp.mark("start", st);
    // input layout change: (B, Cin, T) fp32 -> (B, T, ci_pad) bf16
    const Layer &pre = h->layers[h->idx_pre];
    if ((rc = launch_cf_to_cl_16(c, bf.a_in, B, cfg.in_channels, T, pre.ci_pad, 1.f, fmt, st))) return rc;
    h->launch_count++;
    {   // input_conv -> bf16 LeakyReLU'd operand for upsample 0 (fp32 copy only when dumped)
        TcConvParams p{};
        p.bias_b = bias_b;
        p.out_a = bf.a_c; p.out_a_ld = cfg.channels; p.slope_out = cfg.lrelu_slope;
        p.out_x = (dump_stage == 0) ? bf.x_cs : nullptr;
        if ((rc = run_conv(h, fmt, pre, bf.a_in, B, T, T, p, st, rate, mg_pre))) return rc;
        if ((rc = dump_f32(0, bf.x_cs, cfg.channels, T))) return rc;
    }
    sp.mark("layout + input conv", st);
    const uint16_t *cur_a = bf.a_c;
    int L = T;
    for (int i = 0; i < cfg.num_upsamples; ++i) {
        const Layer &u = h->layers[h->idx_up[i]];
        const int Lo = plan.L[i], C = plan.C[i];
        VTTS_REQUIRE(Lo > 0, "vtts_gen_forward: stage %d output length %d <= 0", i, Lo);
        const bool last_stage = (i == cfg.num_upsamples - 1);
        if (stage_chain[i]) {
            {   // upsample: channels-last fp32 x_u only (the chain kernel builds its own 16-bit operands)
                TcConvParams p{};
                p.out_x = bf.x_u; p.x_cl = 1; p.slope_out = 1.f;
                if ((rc = run_conv(h, fmt, u, cur_a, B, L, Lo, p, st, rate, mg_up[i]))) return rc;
                rate *= u.stride;
                if (dump_stage == 2 * i + 1 && dump_out) {
                    h->launch_count++;
                    if ((rc = launch_cl_to_cf_f32(bf.x_u, dump_out, B, C, Lo, st))) return rc;
                }
            }
            sp.mark((std::string("upsample ") + std::to_string(i)).c_str(), st);
            const int nb = cfg.num_blocks;
            for (int j = 0; j < nb; ++j) {
                ChainRun r;
                r.x = bf.x_u; r.cs = bf.x_cs;
                if (nb == 1) { r.wr = 2; }
                else if (j == 0) { r.wr = 0; }
                else if (j < nb - 1) { r.wr = 1; }
                else { r.wr = 2; r.rd_cs = 1; r.scale = 1.f / (float)nb; }
                if (r.wr == 2) {   // c = cs / num_blocks (generator.py:153): fp32 for the output conv, 16-bit for the next upsample
                    r.out_x = (last_stage || dump_stage == 2 * i + 2) ? bf.x_cs : nullptr;
                    r.out_a = last_stage ? nullptr : bf.a_c;
                }
                r.slope = cfg.lrelu_slope; r.slope_out = cfg.lrelu_slope; r.B = B; r.L = Lo;
                if (h->trim_lens) { r.lens = (const long long *)h->trim_lens; r.len_margin = mg_mrf[i]; r.len_rate = rate; }
                r.pdl = pdl_enabled() && !prof_enabled();
                ProfRec pr{};
                if (prof_enabled()) {
                    cudaEventCreate(&pr.a); cudaEventCreate(&pr.b);
                    pr.kind = 3; pr.cin = C; pr.cout = C; pr.k = cfg.resblock_kernel_sizes[j] * cfg.num_dilations[j] * (cfg.use_additional_convs ? 2 : 1);
                    pr.d = 0; pr.B = B; pr.L = Lo; pr.trimmed = r.lens != nullptr;
                    cudaEventRecord(pr.a, st);
                }
                if ((rc = chain_launch(h->chain[(size_t)i * nb + j], fmt, r, st))) return rc;
                if (prof_enabled()) { cudaEventRecord(pr.b, st); g_prof.push_back(pr); }
                h->launch_count++;
            }
            if (dump_stage == 2 * i + 2 && dump_out) {
                h->launch_count++;
                if ((rc = launch_cl_to_cf_f32(bf.x_cs, dump_out, B, C, Lo, st))) return rc;
            }
            sp.mark((std::string("blocks of stage ") + std::to_string(i) + " (chain)").c_str(), st);
            cur_a = bf.a_c;
            L = Lo;
            continue;
        }
        {   // upsample: fp32 residual stream x_u + bf16 operand a_u = lrelu(x_u)
            TcConvParams p{};
            p.out_x = bf.x_u; p.out_a = bf.a_u; p.out_a_ld = C; p.slope_out = cfg.lrelu_slope;
            if ((rc = run_conv(h, fmt, u, cur_a, B, L, Lo, p, st, rate, mg_up[i]))) return rc;
            rate *= u.stride;
            if ((rc = dump_f32(2 * i + 1, bf.x_u, C, Lo))) return rc;
        }
        sp.mark((std::string("upsample ") + std::to_string(i)).c_str(), st);
        for (int j = 0; j < cfg.num_blocks; ++j) {
            const float *yx = bf.x_u;
            const uint16_t *ya = bf.a_u;
            const int nu = cfg.num_dilations[j];
            for (int m = 0; m < nu; ++m) {
                const bool last = (m == nu - 1);
                float *nx = last ? bf.x_cs : ((m & 1) ? bf.x_q : bf.x_p);
                uint16_t *na = (m & 1) ? bf.a_q : bf.a_p;
                const Layer &l1 = h->layers[h->idx_c1[i][j][m]];
                const Layer *fin = &l1;
                const uint16_t *fin_in = ya;
                const bool fuse = cfg.use_additional_convs &&
                                  unit_fusable(C, l1.info.ksize, l1.info.dilation, h->layers[h->idx_c2[i][j][m]].info.ksize);
                if (cfg.use_additional_convs && !fuse) {
                    TcConvParams p1{};  // xt = conv1(lrelu(x)); only its LeakyReLU'd bf16 copy is needed
                    p1.out_a = bf.a_t; p1.out_a_ld = C; p1.slope_out = cfg.lrelu_slope;
                    if ((rc = run_conv(h, fmt, l1, ya, B, Lo, Lo, p1, st, rate, mg_mrf[i]))) return rc;
                    fin = &h->layers[h->idx_c2[i][j][m]];
                    fin_in = bf.a_t;
                }
                TcConvParams p2{};
                p2.res = yx;               // x = xt + x (layers.py:97)
                p2.out_x = nx;
                if (last) {                // cs += block(c); c = cs / num_blocks (generator.py:150-153)
                    p2.accumulate = (j > 0);
                    if (j == cfg.num_blocks - 1) {
                        p2.divide_by = (float)cfg.num_blocks;
                        if (!last_stage) {  // next consumer: LeakyReLU + upsample
                            p2.out_a = bf.a_c; p2.out_a_ld = C; p2.slope_out = cfg.lrelu_slope;
                        }
                    }
                } else {
                    p2.out_a = na; p2.out_a_ld = C; p2.slope_out = cfg.lrelu_slope;
                }
                if (fuse) {
                    if ((rc = run_unit(h, fmt, l1, h->layers[h->idx_c2[i][j][m]], ya, B, Lo, p2, cfg.lrelu_slope, st, rate, mg_mrf[i]))) return rc;
                } else if ((rc = run_conv(h, fmt, *fin, fin_in, B, Lo, Lo, p2, st, rate, mg_mrf[i]))) return rc;
                yx = nx; ya = na;
            }
        }
        if ((rc = dump_f32(2 * i + 2, bf.x_cs, C, Lo))) return rc;
        sp.mark((std::string("blocks of stage ") + std::to_string(i)).c_str(), st);
        cur_a = bf.a_c;
        L = Lo;
    }
    {   // output_conv in fp32 on the fp32 MRF mean (channels-last), + tanh
        const Layer &post = h->layers[h->idx_post];
        const int C = post.info.cin, k = post.info.ksize;
        const float *wt = post.w_aux;  // [oc][k][ci], packed at load time
        const bool pdl = pdl_enabled() && !prof_enabled();
        dim3 grid((unsigned)ceil_div(L, 256), (unsigned)B);
        if (k > 9) return set_error(VTTS_E_UNSUPPORTED, "conv_post: kernel size %d > 9", k);
        const size_t smem = (size_t)C * (256 + 8 + 4) * sizeof(float);
        for (int oc = 0; oc < post.info.cout; ++oc) {
            const float *w_oc = wt + (size_t)oc * k * C;
            const float *bias = post.has_bias ? post.bias : nullptr;
            if (stage_chain[cfg.num_upsamples - 1]) {   // the last stage left its mean channels-last
                if (post.w_aux_host.size() != (size_t)post.info.cout * k * C) return set_error(VTTS_E_STATE, "conv_post: host weights missing");
                if ((rc = launch_conv_post_cl(bf.x_cs, post.w_aux_host.data() + (size_t)oc * k * C, bias, wav, B, C, L, k,
                                              cfg.final_lrelu_slope, post.info.cout, oc, trim_lens, mg_post, rate, pdl, st))) return rc;
                h->launch_count++;
                continue;
            }
#define VTTS_POST(CC)                                                                                         \
    do {                                                                                                      \
        if (smem > 48 * 1024)                                                                                 \
            VTTS_CHECK_CUDA(cudaFuncSetAttribute(conv_post_tp4_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        VTTS_CHECK_CUDA(launch_kernel_ex(conv_post_tp4_kernel<CC>, grid, dim3(256), smem, st, pdl, 1u, (const float *)bf.x_cs, w_oc, bias, wav, L, \
                                         (L + 3) / 4, k, cfg.final_lrelu_slope, post.info.cout, oc, trim_lens, mg_post, rate)); \
    } while (0)
            if (C == 32) VTTS_POST(32);
            else if (C == 64) VTTS_POST(64);
            else if (C == 128) VTTS_POST(128);
            else return set_error(VTTS_E_UNSUPPORTED, "conv_post: %d input channels", C);
#undef VTTS_POST
            VTTS_CHECK_LAUNCH();
            h->launch_count++;
        }
    }
    sp.mark("output conv", st);
    sp.report();
    if (prof_enabled()) prof_report();
    return VTTS_OK;
}

}  // namespace vtts

// ---------------------------------------------------------------------------------------------
// test hooks
// ---------------------------------------------------------------------------------------------
using namespace vtts;
using namespace vtts::tc;

namespace vtts {
namespace probe {
using namespace vtts::tc;
// Minimal single-CTA GEMM used to pin descriptor semantics on real hardware:
// D[m, n] = sum_k A[m, k] * B[n + row_shift, k];  M = 128 (TMEM lanes), N columns.
template <int ROWB>
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float *d,
                  int N, int kblocks, int b_rows, int row_shift, int variant) {
    constexpr int CH = ROWB / 2;
    constexpr int KSTEPS = CH / 16;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *s_a = smem;                       // 128 rows
    uint8_t *s_b = smem + 128 * ROWB;          // b_rows rows (multiple of 64)
    uint64_t *bar_ld = reinterpret_cast<uint64_t *>(s_b + (size_t)b_rows * ROWB);
    uint64_t *bar_mma = bar_ld + 1;
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar_ld, 1); mbar_init(bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_16((variant & 4) ? 64 : 128, N, 0);
        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_arrive_expect_tx(bar_ld, (uint32_t)((128 + b_rows) * ROWB));
            tma_load_2d(s_a, &tm_a, bar_ld, kb * CH, 0);
            tma_load_2d(s_a + 64 * ROWB, &tm_a, bar_ld, kb * CH, 64);
            for (int r = 0; r < b_rows; r += 64) tma_load_2d(s_b + (size_t)r * ROWB, &tm_b, bar_ld, kb * CH, r);
            mbar_wait(bar_ld, (uint32_t)kb & 1u);
            tc_fence_after();
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const uint32_t a_addr = smem_u32(s_a) + ks * 32;
                const uint32_t b_addr = smem_u32(s_b) + row_shift * ROWB + ks * 32;
                const uint32_t bo = (variant & 1) ? ((b_addr >> 7) & 7u) : 0u;
                umma_bf16(tmem, make_smem_desc(a_addr, ROWB, 0), make_smem_desc(b_addr, ROWB, bo), idesc,
                          (uint32_t)((kb | ks) != 0));
            }
            umma_commit(bar_mma);
            mbar_wait(bar_mma, (uint32_t)kb & 1u);
        }
    }
    __syncthreads();
    tc_fence_after();
    if (variant & 8) {
        // readout through 16-lane mma-style fragments (tcgen05.ld 16x256b.x2), written back under the assumed layout
        for (int h = 0; h < 2; ++h)
            for (int col = 0; col < N; col += 16) {
                uint32_t v[8];
                tmem_ld_16x256_x2(tmem + ((uint32_t)(warp * 32 + h * 16) << 16) + (uint32_t)col, v);
                tmem_ld_wait();
                for (int e = 0; e < 8; ++e) {
                    const int m = warp * 32 + h * 16 + (lane >> 2) + ((e >> 1) & 1) * 8;
                    const int c = col + ((e >> 2) & 1) * 8 + (lane & 3) * 2 + (e & 1);
                    d[(size_t)m * N + c] = __uint_as_float(v[e]);
                }
            }
    } else {
    for (int col = 0; col < N; col += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)col, v);
        tmem_ld_wait();
        const int m = warp * 32 + lane;
        for (int e = 0; e < 32; ++e) d[(size_t)m * N + col + e] = __uint_as_float(v[e]);
    }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}
}  // namespace probe
}  // namespace vtts
using namespace vtts::probe;

extern "C" int vtts_dbg_umma_gemm(const void *a_bf16, const void *b_bf16, float *d, int M, int N, int K,
                                  int b_rows_total, int row_shift, int variant, vtts_stream_t stream) {
    VTTS_REQUIRE(a_bf16 && b_bf16 && d, "vtts_dbg_umma_gemm: null pointer");
    VTTS_REQUIRE(M == 128, "vtts_dbg_umma_gemm: M must be 128 (variant bit 2 issues M=64 MMAs on the first 64 rows)");
    VTTS_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256, "vtts_dbg_umma_gemm: N must be a multiple of 32 in [32,256]");
    const int rowb = (variant & 2) ? 64 : 128;
    const int ch = rowb / 2;
    VTTS_REQUIRE(K % ch == 0 && K > 0, "vtts_dbg_umma_gemm: K must be a multiple of %d", ch);
    VTTS_REQUIRE(b_rows_total % 64 == 0 && b_rows_total >= N + row_shift && b_rows_total <= 512 && row_shift >= 0,
                 "vtts_dbg_umma_gemm: b_rows_total must be a multiple of 64 covering N + row_shift");
    CUtensorMap ta, tb;
    {
        uint64_t dims[2] = {(uint64_t)K, 128};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {(uint32_t)ch, 64};
        int rc = make_tmap_bf16(&ta, a_bf16, 2, dims, str, box, rowb);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)b_rows_total};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {(uint32_t)ch, 64};
        int rc = make_tmap_bf16(&tb, b_bf16, 2, dims, str, box, rowb);
        if (rc) return rc;
    }
    const size_t smem = 1024 + (size_t)(128 + b_rows_total) * rowb + 64;
    cudaStream_t st = (cudaStream_t)stream;
    if (rowb == 128) {
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        umma_probe_kernel<128><<<1, 128, smem, st>>>(ta, tb, d, N, K / ch, b_rows_total, row_shift, variant);
    } else {
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        umma_probe_kernel<64><<<1, 128, smem, st>>>(ta, tb, d, N, K / ch, b_rows_total, row_shift, variant);
    }
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

// Raw tcgen05.mma throughput: `reps` x KSTEPS MMAs (M=128, N, K=16) issued back to back by one thread on
// resident operands; cycles from first issue to commit completion are written to out[0..1].
template <int ROWB>
__global__ void __launch_bounds__(128, 1)
umma_bench_kernel(long long *out, int N, int row_shift, int reps, int a_rows, int two_acc) {
    constexpr int KSTEPS = ROWB / 32;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *s_a = smem;                         // 128 rows
    uint8_t *s_b = smem + 128 * ROWB;            // 320 rows
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_b + 320 * ROWB);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    for (int i = threadIdx.x; i < (128 + 320) * ROWB / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_16(a_rows, N, 1);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const uint32_t a_addr = smem_u32(s_a) + ks * 32;
                const uint32_t b_addr = smem_u32(s_b) + ((r % 11) * row_shift) * ROWB + ks * 32;
                umma_bf16(tmem + ((two_acc && (r & 1)) ? 256u : 0u), make_smem_desc(a_addr, ROWB, 0),
                          make_smem_desc(b_addr, ROWB, 0), idesc, 1u);
            }
        }
        const long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

extern "C" int vtts_dbg_umma_bench(int N, int rowb, int row_shift, int reps, int M, int two_acc, long long *host_out) {
    long long *d = nullptr;
    VTTS_CHECK_CUDA(cudaMalloc(&d, 16));
    const size_t smem = (size_t)(128 + 320) * rowb + 64;
    if (rowb == 128) {
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(umma_bench_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        umma_bench_kernel<128><<<148, 128, smem>>>(d, N, row_shift, reps, M, two_acc);
    } else {
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(umma_bench_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        umma_bench_kernel<64><<<148, 128, smem>>>(d, N, row_shift, reps, M, two_acc);
    }
    VTTS_CHECK_LAUNCH();
    VTTS_CHECK_CUDA(cudaDeviceSynchronize());
    VTTS_CHECK_CUDA(cudaMemcpy(host_out, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return VTTS_OK;
}

extern "C" int vtts_dbg_trace(int enable, long long *host_out, int n) {
    g_trace_on = enable;
    if (host_out && n > 0) {
        if (n > TRACE_TILES * 16) n = TRACE_TILES * 16;
        VTTS_CHECK_CUDA(cudaDeviceSynchronize());
        VTTS_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * n));
    }
    return VTTS_OK;
}

// Single Conv1d layer through the tensor-core kernel, channels-first fp32 in/out (test hook).
extern "C" int vtts_dbg_conv1d_tc(const float *x, const float *w, const float *bias, const float *res, float *y,
                                  float *y_act, int B, int cin, int cout, int L, int ksize, int dilation,
                                  float slope_in, float slope_out, int fp16, vtts_stream_t stream) {
    VTTS_REQUIRE(x && w && (y || y_act), "vtts_dbg_conv1d_tc: null pointer");
    const int fmt = fp16 ? VTTS_FMT_FP16 : VTTS_FMT_BF16;
    cudaStream_t st = (cudaStream_t)stream;
    const int ci_pad = ci_pad_of(cin), n_pad = pad_to(cout, TM);
    uint16_t *a = nullptr, *wp = nullptr, *oa = nullptr;
    float *ox = nullptr, *rcl = nullptr;
    VTTS_CHECK_CUDA(cudaMalloc(&a, (size_t)B * L * ci_pad * 2));
    VTTS_CHECK_CUDA(cudaMalloc(&wp, (size_t)ksize * n_pad * ci_pad * 2));
    VTTS_CHECK_CUDA(cudaMalloc(&oa, (size_t)B * L * cout * 2));
    const size_t L4p = (size_t)(L + 3) / 4 * 4;
    VTTS_CHECK_CUDA(cudaMalloc(&ox, (size_t)B * L4p * cout * 4));
    VTTS_CHECK_CUDA(cudaMalloc(&rcl, (size_t)B * L4p * cout * 4));
    int rc = launch_cf_to_cl_16(x, a, B, cin, L, ci_pad, slope_in, fmt, st);
    if (!rc) {
        size_t n = (size_t)ksize * n_pad * ci_pad;
        pack_tc_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(w, wp, fmt, 0, cin, cout, ksize, 1, ksize, n_pad, ci_pad, cout, 0);
    }
    if (!rc && res) {
        rc = launch_cf_to_tp4(res, rcl, B, cout, L, st);
    }
    if (!rc) {
        TcConvParams p{};
        p.bias = bias; p.res = res ? rcl : nullptr; p.out_x = y ? ox : nullptr; p.out_a = y_act ? oa : nullptr; p.out_a_ld = cout;
        p.slope_out = slope_out; p.n_total = cout; p.cout = cout; p.L_out = L; p.n_pos = L; p.out_stride = 1;
        p.out_off0 = 0; p.taps = ksize; p.tap_off0 = -(ksize - 1) / 2 * dilation; p.tap_step = dilation;
        p.trace = g_trace_on;
        TcLaunch Lc;
        rc = tc_prepare(Lc, fmt, a, B, L, ci_pad, wp, n_pad, p);
        if (!rc) rc = tc_launch(Lc, st);
    }
    if (!rc && y) rc = launch_tp4_to_cf(ox, y, B, cout, L, st);
    if (!rc && y_act) rc = launch_cl_16_to_cf_f32(oa, y_act, B, cout, L, cout, fmt, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(a); cudaFree(wp); cudaFree(oa); cudaFree(ox); cudaFree(rcl);
    if (!rc && e != cudaSuccess) return set_error(VTTS_E_CUDA, "vtts_dbg_conv1d_tc: %s", cudaGetErrorString(e));
    return rc;
}


// ---------------------------------------------------------------------------------------------
// Stand-alone Conv1d layer on the tcgen05 kernel (channels-last activations): the convolutions of the acoustic decoder
// between the LengthRegulator and the generator -- PositionwiseFeedForward w_1 / w_2 (blocks/transformer.py:265-298) and
// the Postnet's five Conv1d + eval-mode BatchNorm (fastspeech2/layers.py:571-625, BatchNorm folded by the caller).
// ---------------------------------------------------------------------------------------------
struct VttsConv {
    int cin = 0, cout = 0, k = 0, dil = 1, ci_pad = 0, n_pad = 0;
    uint16_t *w16[2] = {nullptr, nullptr};
    float *bias = nullptr;
    bool has_bias = false;   // (the allocation is kept when a later load passes no bias: cudaFree would synchronise the device)
    bool loaded = false;
};

extern "C" int vtts_conv_create(int cin, int cout, int ksize, int dilation, VttsConv **out) {
    VTTS_REQUIRE(out, "vtts_conv_create: null pointer");
    VTTS_REQUIRE(cin >= 1 && cout >= 1 && ksize >= 1 && ksize % 2 == 1 && dilation >= 1, "vtts_conv_create: bad shape");
    if ((ksize - 1) * dilation > HALO_MAX) return set_error(VTTS_E_UNSUPPORTED, "vtts_conv_create: (kernel-1)*dilation > %d", HALO_MAX);
    VttsConv *c = new (std::nothrow) VttsConv();
    if (!c) return set_error(VTTS_E_INVALID, "vtts_conv_create: out of host memory");
    c->cin = cin; c->cout = cout; c->k = ksize; c->dil = dilation;
    c->ci_pad = ci_pad_of(cin); c->n_pad = pad_to(cout, TM);
    *out = c;
    return VTTS_OK;
}

extern "C" void vtts_conv_destroy(VttsConv *c) {
    if (!c) return;
    cudaFree(c->w16[0]); cudaFree(c->w16[1]); cudaFree(c->bias);
    delete c;
}

extern "C" int vtts_conv_padded_channels(const VttsConv *c) {
    VTTS_REQUIRE(c, "vtts_conv_padded_channels: null handle");
    return c->ci_pad;
}

extern "C" int vtts_conv_load(VttsConv *c, const float *weight, const float *bias, vtts_stream_t stream) {
    VTTS_REQUIRE(c && weight, "vtts_conv_load: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)c->k * c->n_pad * c->ci_pad;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 8192) blocks = 8192;
    for (int fmt = 0; fmt < 2; ++fmt) {
        if (!c->w16[fmt]) VTTS_CHECK_CUDA(cudaMalloc(&c->w16[fmt], n * sizeof(uint16_t)));
        pack_tc_kernel<<<blocks, 256, 0, st>>>(weight, c->w16[fmt], fmt, 0, c->cin, c->cout, c->k, 1, c->k, c->n_pad, c->ci_pad, c->cout, 0);
        VTTS_CHECK_LAUNCH();
    }
    if (bias) {
        if (!c->bias) VTTS_CHECK_CUDA(cudaMalloc(&c->bias, (size_t)c->cout * sizeof(float)));
        VTTS_CHECK_CUDA(cudaMemcpyAsync(c->bias, bias, (size_t)c->cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    c->has_bias = bias != nullptr;
    c->loaded = true;
    return VTTS_OK;
}

extern "C" int vtts_conv_forward(VttsConv *c, const void *act16, int precision, int B, int L, const float *res, float *out_x,
                                 void *out_a16, float slope_out, int act_tanh, vtts_stream_t stream) {
    VTTS_REQUIRE(c && act16 && (out_x || out_a16), "vtts_conv_forward: null pointer");
    VTTS_REQUIRE(c->loaded, "vtts_conv_forward: no weights loaded");
    VTTS_REQUIRE(B >= 1 && L >= 1, "vtts_conv_forward: B and L must be >= 1");
    if (out_a16 && c->cout % 32 != 0) return set_error(VTTS_E_UNSUPPORTED, "vtts_conv_forward: 16-bit output needs cout %% 32 == 0 (cout = %d)", c->cout);
    if (precision != VTTS_PRECISION_BF16 && precision != VTTS_PRECISION_FP16)
        return set_error(VTTS_E_INVALID, "vtts_conv_forward: precision must be bf16 or fp16");
    const int fmt = precision == VTTS_PRECISION_BF16 ? VTTS_FMT_BF16 : VTTS_FMT_FP16;
    TcConvParams p{};
    p.bias = c->has_bias ? c->bias : nullptr; p.res = res; p.out_x = out_x; p.out_a = (uint16_t *)out_a16; p.out_a_ld = c->cout;
    p.slope_out = slope_out; p.act_tanh = act_tanh; p.x_cl = 1;
    p.n_total = c->cout; p.cout = c->cout; p.L_out = L; p.n_pos = L; p.out_stride = 1; p.out_off0 = 0;
    p.taps = c->k; p.tap_off0 = -(c->k - 1) / 2 * c->dil; p.tap_step = c->dil;
    TcLaunch Lc;
    int rc = tc_prepare(Lc, fmt, (const uint16_t *)act16, B, L, c->ci_pad, c->w16[fmt], c->n_pad, p);
    if (rc) return rc;
    Lc.pdl = false;
    return tc_launch(Lc, (cudaStream_t)stream);
}
