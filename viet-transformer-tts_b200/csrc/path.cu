// path.cu -- vits2 monotonic duration path: generate_path and the attn matmuls that expand the
// prior statistics from token rate to frame rate.
//
// Reference: models/gan_tts/vits2/utils.py:111-126 (generate_path), call site
// models/gan_tts/vits2/generator.py:251-259:
//     attn = generate_path(w_ceil, attn_mask)                                   # (B,1,t_y,t_x), 0/1 floats
//     m_p  = torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2) # (B,D,t_x) -> (B,D,t_y)
// generate_path: cum = cumsum(duration); path[b,y,x] = ((y < cum[x]) - (y < cum[x-1])) * mask[b,y,x].
// Each row y of the path holds at most one 1 for non-negative durations, so the matmul is a gather: no GEMM, no
// (B,t_y,t_x) attention tensor in HBM.  HBM-bound: reads D*t_x, writes D*t_y floats per utterance.
//
// The running sum is a sequential fp32 scan like torch.cumsum on the CPU; durations at the call site are ceil()ed,
// i.e. integer valued, so every summation order gives the same partial sums.
#include "common.cuh"

namespace vtts {
namespace {

constexpr int PATH_THREADS = 256;

// cum[x] = duration[0] + ... + duration[x] for one batch row, into shared memory (all threads return after the sync)
__device__ void row_cumsum(const float *__restrict__ d, float *s_cum, int t_x) {
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int x = 0; x < t_x; ++x) { acc += d[x]; s_cum[x] = acc; }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PATH_THREADS)
path_generate_kernel(const float *__restrict__ duration, const float *__restrict__ mask, float *__restrict__ path, int t_y, int t_x) {
    extern __shared__ float s_cum[];
    const int b = blockIdx.y;
    row_cumsum(duration + (size_t)b * t_x, s_cum, t_x);
    const size_t base = (size_t)b * t_y * t_x;
    const long long n = (long long)t_y * t_x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / t_x), x = (int)(i - (long long)y * t_x);
        const float fy = (float)y;
        // sequence_mask(cum, t_y) minus the same shifted by one token (utils.py:121-124)
        const float cur = fy < s_cum[x] ? 1.f : 0.f;
        const float prev = (x > 0 && fy < s_cum[x - 1]) ? 1.f : 0.f;
        const float v = cur - prev;
        path[base + i] = mask ? v * mask[base + i] : v;
    }
}

// out[b,d,y] = sum_x path[b,y,x] * x_in[b,d,x]; lanes own consecutive frames y (coalesced writes)
__global__ void __launch_bounds__(PATH_THREADS)
path_expand_kernel(const float *__restrict__ x_in, const float *__restrict__ duration, const float *__restrict__ mask,
                   float *__restrict__ out, int D, int t_y, int t_x) {
    extern __shared__ float s_cum[];
    const int b = blockIdx.y;
    row_cumsum(duration + (size_t)b * t_x, s_cum, t_x);
    // 32 consecutive frames per CTA (lanes: coalesced writes) x 8 feature groups (warps): the first version ran one thread
    // per frame over all D features on 3 x B CTAs - 71 us at (16, 192, 120 -> 760), 7 % issue-slot use
    const int y = blockIdx.x * 32 + (threadIdx.x & 31);
    const int dg = threadIdx.x >> 5, DG = PATH_THREADS / 32;
    if (y >= t_y) return;
    const float fy = (float)y;
    const float *xb = x_in + (size_t)b * D * t_x;
    const float *mb = mask ? mask + ((size_t)b * t_y + y) * t_x : nullptr;
    float *ob = out + (size_t)b * D * t_y + y;
    // nonzero entries of this path row
    int cnt = 0, x1 = 0;
    float w1 = 0.f;
    for (int x = 0; x < t_x; ++x) {
        const float v = (fy < s_cum[x] ? 1.f : 0.f) - ((x > 0 && fy < s_cum[x - 1]) ? 1.f : 0.f);
        if (v != 0.f) {
            const float w = mb ? v * mb[x] : v;
            if (cnt == 0) { x1 = x; w1 = w; }
            ++cnt;
        }
    }
    if (cnt <= 1) {  // the monotonic case: a gather (0.0f + ... normalises -0 like a sum of products does)
        for (int d = dg; d < D; d += DG) ob[(size_t)d * t_y] = cnt ? 0.f + w1 * xb[(size_t)d * t_x + x1] : 0.f;
        return;
    }
    for (int d = dg; d < D; d += DG) {  // negative durations: several +-1 entries per row; sum them in token order
        float acc = 0.f;
        for (int x = 0; x < t_x; ++x) {
            const float v = (fy < s_cum[x] ? 1.f : 0.f) - ((x > 0 && fy < s_cum[x - 1]) ? 1.f : 0.f);
            if (v != 0.f) acc += (mb ? v * mb[x] : v) * xb[(size_t)d * t_x + x];
        }
        ob[(size_t)d * t_y] = acc;
    }
}

}  // namespace
}  // namespace vtts

using namespace vtts;

extern "C" int vtts_path_generate(const float *duration, const float *mask, float *path, int B, int t_y, int t_x,
                                  vtts_stream_t stream) {
    VTTS_REQUIRE(B >= 0 && t_y >= 0 && t_x >= 0, "vtts_path_generate: negative size");
    if (B == 0 || t_y == 0 || t_x == 0) return VTTS_OK;
    VTTS_REQUIRE(duration && path, "vtts_path_generate: null pointer");
    VTTS_REQUIRE((size_t)t_x * sizeof(float) <= 160 * 1024, "vtts_path_generate: t_x %d too large", t_x);
    const size_t smem = (size_t)t_x * sizeof(float);
    if (smem > 48 * 1024) VTTS_CHECK_CUDA(cudaFuncSetAttribute(path_generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n = (long long)t_y * t_x;
    const unsigned gx = (unsigned)((n + PATH_THREADS * 8LL - 1) / (PATH_THREADS * 8LL));
    path_generate_kernel<<<dim3(gx < 1 ? 1 : (gx > 4096 ? 4096 : gx), (unsigned)B), PATH_THREADS, smem, (cudaStream_t)stream>>>(duration, mask, path, t_y, t_x);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

extern "C" int vtts_path_expand(const float *x, const float *duration, const float *mask, float *out, int B, int D, int t_y,
                                int t_x, vtts_stream_t stream) {
    VTTS_REQUIRE(B >= 0 && D >= 0 && t_y >= 0 && t_x >= 0, "vtts_path_expand: negative size");
    if (B == 0 || D == 0 || t_y == 0) return VTTS_OK;
    VTTS_REQUIRE(out && (t_x == 0 || (x && duration)), "vtts_path_expand: null pointer");
    VTTS_REQUIRE((size_t)t_x * sizeof(float) <= 160 * 1024, "vtts_path_expand: t_x %d too large", t_x);
    const size_t smem = (size_t)(t_x > 0 ? t_x : 1) * sizeof(float);
    if (smem > 48 * 1024) VTTS_CHECK_CUDA(cudaFuncSetAttribute(path_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    path_expand_kernel<<<dim3((unsigned)ceil_div(t_y, 32), (unsigned)B), PATH_THREADS, smem, (cudaStream_t)stream>>>(x, duration, mask, out, D, t_y, t_x);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}
