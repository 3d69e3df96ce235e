// gauss.cu -- GaussianUpsampling (the reference's default regulator, `use_gaussian: true`).
//
// Replaces models/tts/fastspeech2/layers.py:465-520 (same body: models/gan_tts/jets/alignments.py:168-222):
//   c_j   = cumsum(ds)_j - ds_j / 2
//   e_tj  = -delta * (t - c_j)^2            (t multiplied by h_mask; masked tokens -> -inf)
//   out_t = softmax_j(e_t.) @ hs            fp32 throughout
// One fused kernel: a CTA owns FR consecutive output frames of one batch row, builds the token centres with
// a block prefix sum in shared memory, each warp runs the softmax of FRW frames at once (probabilities kept
// in shared memory) and accumulates the weighted sum of token rows with every hs row read once per FRW
// frames (coalesced, 128 bytes per warp access), only over the tokens whose probability is non-zero for one of those
// frames.  Floating point: parity within 1e-5 of the fp32 oracle.
// Work per call: B * T_feats * (window ~ 64 frames / mean duration) * D FMA; bytes: hs rows of the window re-read per 4
// frames (L1/L2 hits), out B * T_feats * D * 4 written once.
#include "common.cuh"

namespace vtts {

constexpr int GU_THREADS = 256;
constexpr int GU_WARPS = GU_THREADS / 32;
constexpr int GU_FRW = 4;                    // frames per warp
constexpr int GU_FR = GU_WARPS * GU_FRW;     // frames per CTA
constexpr int GU_DCH = 8;                    // feature chunks of 32 per pass (256 features)
constexpr int GU_TOK = 32;                   // token rows staged in shared memory per pass (32 x 256 x 4 B = 32 KB)

__global__ void __launch_bounds__(GU_THREADS)
gauss_upsample_kernel(const float *__restrict__ hs, const long long *__restrict__ ds,
                      const unsigned char *__restrict__ h_mask, const unsigned char *__restrict__ d_mask,
                      float *__restrict__ out, int T_text, int D, int T_feats, float neg_delta) {
    extern __shared__ float sm[];
    float *s_c = sm;                                   // [T_text] token centres
    float *s_p = sm + T_text;                          // [GU_WARPS][GU_FRW][T_text] probabilities
    __shared__ long long warp_tot[GU_WARPS];
    __shared__ int s_jlo, s_jhi;                       // union of the warps' token windows (ordered by the barriers below)
    if (threadIdx.x == 0) { s_jlo = T_text; s_jhi = 0; }
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long *drow = ds + (size_t)b * T_text;

    // centres: c_j = float(cumsum_j) - float(d_j) / 2   (layers.py:509)
    long long carry = 0;
    for (int base = 0; base < T_text; base += GU_THREADS) {
        const int j = base + threadIdx.x;
        const long long d = j < T_text ? drow[j] : 0;
        long long s = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += n;
        }
        if (lane == 31) warp_tot[warp] = s;
        __syncthreads();
        long long off = carry, tile = 0;
#pragma unroll
        for (int w = 0; w < GU_WARPS; ++w) { const long long t = warp_tot[w]; if (w < warp) off += t; tile += t; }
        if (j < T_text) s_c[j] = __fsub_rn(__ll2float_rn(s + off), __fdiv_rn(__ll2float_rn(d), 2.0f));
        carry += tile;
        __syncthreads();
    }

    const int t0 = blockIdx.x * GU_FR + warp * GU_FRW;
    const bool active = t0 < T_feats;                  // (inactive warps of the last frame tile only keep the barriers)
    float *p = s_p + (size_t)warp * GU_FRW * T_text;
    const unsigned char *dm = d_mask ? d_mask + (size_t)b * T_text : nullptr;

    float tt[GU_FRW], mx[GU_FRW], sum[GU_FRW];
    int jlo = T_text, jhi = 0;
    if (active) {
        // softmax per frame (layers.py:505-516): lanes stride over tokens
#pragma unroll
        for (int f = 0; f < GU_FRW; ++f) {
            const int t = t0 + f;
            float tv = (float)t;
            if (h_mask && t < T_feats) tv = tv * (h_mask[(size_t)b * T_feats + t] ? 1.0f : 0.0f);   // t = t * h_masks.float()
            tt[f] = tv;
            mx[f] = -INFINITY;
        }
        for (int j = lane; j < T_text; j += 32) {
            const float c = s_c[j];
            const bool keep = !dm || dm[j];
#pragma unroll
            for (int f = 0; f < GU_FRW; ++f) {
                const float d = tt[f] - c;
                const float e = keep ? neg_delta * (d * d) : -INFINITY;
                p[f * T_text + j] = e;
                mx[f] = fmaxf(mx[f], e);
            }
        }
#pragma unroll
        for (int f = 0; f < GU_FRW; ++f) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx[f] = fmaxf(mx[f], __shfl_xor_sync(0xffffffffu, mx[f], o));
            sum[f] = 0.f;
        }
        __syncwarp();
        for (int j = lane; j < T_text; j += 32) {
#pragma unroll
            for (int f = 0; f < GU_FRW; ++f) {
                const float ex = expf(p[f * T_text + j] - mx[f]);   // all -inf row -> NaN, like torch.softmax
                p[f * T_text + j] = ex;
                sum[f] += ex;
            }
        }
#pragma unroll
        for (int f = 0; f < GU_FRW; ++f)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum[f] += __shfl_xor_sync(0xffffffffu, sum[f], o);
        __syncwarp();
        // normalise; remember the token window outside of which every probability of these frames is exactly 0 (exp
        // underflows to 0 beyond |t - c_j| of about 32 frames at delta = 0.1): the weighted sum below skips those tokens,
        // which adds exact zeros in the reference's matmul - bit-identical for finite hs
        for (int j = lane; j < T_text; j += 32) {
            bool nz = false;
#pragma unroll
            for (int f = 0; f < GU_FRW; ++f) {
                const float v = __fdiv_rn(p[f * T_text + j], sum[f]);
                p[f * T_text + j] = v;
                nz |= (v != 0.f);                          // (NaN rows - everything masked - stay in the window)
            }
            if (nz) { jlo = min(jlo, j); jhi = max(jhi, j + 1); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            jlo = min(jlo, __shfl_xor_sync(0xffffffffu, jlo, o));
            jhi = max(jhi, __shfl_xor_sync(0xffffffffu, jhi, o));
        }
        __syncwarp();

        if (lane == 0 && jlo < jhi) { atomicMin(&s_jlo, jlo); atomicMax(&s_jhi, jhi); }
    }
    __syncthreads();
    // weighted sum (layers.py:517).  The 32 frames of this CTA see nearly the same token window: its hs rows are staged in
    // shared memory by all threads (coalesced, many loads in flight) and every warp accumulates its own sub-window from
    // there in ascending token order - the same order as before, so the result is bit-identical.  (Per-warp global loads
    // left the kernel latency-bound: ~1 us per token and warp, 29 us for a single CTA.)
    const int J0 = s_jlo, J1 = s_jhi;
    float *s_h = s_p + (size_t)GU_WARPS * GU_FRW * T_text;           // [GU_TOK][dw] staged rows
    const float *hb = hs + (size_t)b * T_text * D;
    for (int d0 = 0; d0 < D; d0 += 32 * GU_DCH) {
        const int dw = D - d0 < 32 * GU_DCH ? D - d0 : 32 * GU_DCH;  // feature columns of this pass
        float acc[GU_FRW][GU_DCH];
#pragma unroll
        for (int f = 0; f < GU_FRW; ++f)
#pragma unroll
            for (int q = 0; q < GU_DCH; ++q) acc[f][q] = 0.f;
        for (int jc = J0; jc < J1; jc += GU_TOK) {
            const int nt = J1 - jc < GU_TOK ? J1 - jc : GU_TOK;
            for (int idx = threadIdx.x; idx < nt * dw; idx += GU_THREADS) {
                const int r = idx / dw, c = idx - r * dw;
                s_h[r * dw + c] = __ldg(hb + (size_t)(jc + r) * D + d0 + c);
            }
            __syncthreads();
            if (active) {
                const int ja = jlo > jc ? jlo : jc, jb = jhi < jc + nt ? jhi : jc + nt;
                for (int j = ja; j < jb; ++j) {
                    float pj[GU_FRW];
#pragma unroll
                    for (int f = 0; f < GU_FRW; ++f) pj[f] = p[f * T_text + j];
                    const float *row = s_h + (j - jc) * dw + lane;
#pragma unroll
                    for (int q = 0; q < GU_DCH; ++q) {
                        const float h = (q * 32 + lane) < dw ? row[q * 32] : 0.f;
#pragma unroll
                        for (int f = 0; f < GU_FRW; ++f) acc[f][q] = fmaf(pj[f], h, acc[f][q]);
                    }
                }
            }
            __syncthreads();
        }
        if (active) {
#pragma unroll
            for (int f = 0; f < GU_FRW; ++f) {
                const int t = t0 + f;
                if (t >= T_feats) continue;
                float *o = out + ((size_t)b * T_feats + t) * D + d0 + lane;
#pragma unroll
                for (int q = 0; q < GU_DCH; ++q)
                    if (q * 32 + lane < dw) o[q * 32] = acc[f][q];
            }
        }
    }
}

}  // namespace vtts

using namespace vtts;

extern "C" int vtts_gauss_upsample(const float *hs, const int64_t *ds, const unsigned char *h_mask,
                                   const unsigned char *d_mask, float *out, int B, int T_text, int D, int T_feats,
                                   float delta, vtts_stream_t stream) {
    VTTS_REQUIRE(B >= 0 && T_text >= 0 && D >= 0 && T_feats >= 0, "vtts_gauss_upsample: negative shape");
    if (B == 0 || T_feats == 0 || D == 0) return VTTS_OK;
    VTTS_REQUIRE(hs && ds && out, "vtts_gauss_upsample: null pointer");
    VTTS_REQUIRE(T_text >= 1, "vtts_gauss_upsample: T_text must be >= 1");
    const size_t smem = sizeof(float) * ((size_t)T_text + (size_t)GU_WARPS * GU_FRW * T_text + (size_t)GU_TOK * 32 * GU_DCH);
    if (smem > 200 * 1024)
        return set_error(VTTS_E_UNSUPPORTED, "vtts_gauss_upsample: T_text %d needs %zu B shared memory", T_text, smem);
    if (smem > 48 * 1024)
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(gauss_upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(T_feats, GU_FR), (unsigned)B);
    if (grid.y > 65535) return set_error(VTTS_E_UNSUPPORTED, "vtts_gauss_upsample: batch %d too large", B);
    gauss_upsample_kernel<<<grid, GU_THREADS, smem, (cudaStream_t)stream>>>(
        hs, (const long long *)ds, h_mask, d_mask, out, T_text, D, T_feats, -1.0f * delta);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}
