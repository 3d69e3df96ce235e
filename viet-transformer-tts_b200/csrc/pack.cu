// pack.cu -- weight-norm folding, weight packing and layout conversion kernels.
//
// Weight norm: old-style torch.nn.utils.weight_norm (generator.py:185-195, vits2/layers.py:134)
// stores weight_g (dim0,1,1) and weight_v; effective w = v * (g / ||v||_(1,2)) per dim-0 index
// (out-channel for Conv1d, in-channel for ConvTranspose1d).
#include "common.cuh"

namespace vtts {

__global__ void fold_weight_norm_kernel(const float *__restrict__ v, const float *__restrict__ g,
                                        float *__restrict__ w, int inner) {
    __shared__ double s_part[8];
    __shared__ float s_scale;
    const int r = blockIdx.x;
    const float *vr = v + (size_t)r * inner;
    double s = 0.0;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) {
        double t = (double)vr[i];
        s += t * t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_part[k];
        float norm = (float)sqrt(t);
        s_scale = __fdiv_rn(g[r], norm);
    }
    __syncthreads();
    const float sc = s_scale;
    float *wr = w + (size_t)r * inner;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) wr[i] = vr[i] * sc;
}

int launch_fold_weight_norm(const float *v, const float *g, float *w, int dim0, int inner,
                            cudaStream_t stream) {
    fold_weight_norm_kernel<<<dim0, 256, 0, stream>>>(v, g, w, inner);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

// Conv1d (cout,cin,k) -> [ci][tap][co]
__global__ void pack_conv_fp32_kernel(const float *__restrict__ w, float *__restrict__ packed,
                                      int cout, int cin, int k) {
    size_t n = (size_t)cout * cin * k;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        int co = (int)(idx % cout);
        size_t rest = idx / cout;
        int j = (int)(rest % k), ci = (int)(rest / k);
        packed[idx] = w[((size_t)co * cin + ci) * k + j];
    }
}

int launch_pack_conv_fp32(const float *w, float *packed, int cout, int cin, int k,
                          cudaStream_t stream) {
    size_t n = (size_t)cout * cin * k;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    pack_conv_fp32_kernel<<<blocks, 256, 0, stream>>>(w, packed, cout, cin, k);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

// ConvTranspose1d (cin,cout,k) -> [q][ci][j][co] with kk = q + j*s
__global__ void pack_convT_fp32_kernel(const float *__restrict__ w, float *__restrict__ packed,
                                       int cin, int cout, int k, int s) {
    const int taps = k / s;
    size_t n = (size_t)cin * cout * k;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        int co = (int)(idx % cout);
        size_t rest = idx / cout;
        int j = (int)(rest % taps);
        rest /= taps;
        int ci = (int)(rest % cin), q = (int)(rest / cin);
        packed[idx] = w[((size_t)ci * cout + co) * k + q + j * s];
    }
}

int launch_pack_convT_fp32(const float *w, float *packed, int cin, int cout, int k, int s,
                           cudaStream_t stream) {
    size_t n = (size_t)cin * cout * k;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    pack_convT_fp32_kernel<<<blocks, 256, 0, stream>>>(w, packed, cin, cout, k, s);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

// ---------------------------------------------------------------------------------------------
// layout conversion through a 32x32 shared-memory transpose tile
// (B, C, L) fp32 -> (B, L, Cpad) bf16/fp16 with LeakyReLU(slope) fused, zero channel padding
__global__ void cf_to_cl_16_kernel(const float *__restrict__ x, uint16_t *__restrict__ y,
                                   int C, int L, int Cpad, float slope, int fmt) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *xb = x + (size_t)b * C * L;
    uint16_t *yb = y + (size_t)b * L * Cpad;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, l = l0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && l < L) ? xb[(size_t)c * L + l] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int l = l0 + r, c = c0 + threadIdx.x;
        if (l < L && c < Cpad) yb[(size_t)l * Cpad + c] = cvt16(lrelu(tile[threadIdx.x][r], slope), fmt);
    }
}

int launch_cf_to_cl_16(const float *x, uint16_t *y, int B, int C, int L, int Cpad, float slope, int fmt,
                       cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div(L, 32), (unsigned)ceil_div(Cpad, 32), (unsigned)B);
    cf_to_cl_16_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, y, C, L, Cpad, slope, fmt);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

// (B, L, C) fp32 -> (B, C, L) fp32
__global__ void cl_to_cf_f32_kernel(const float *__restrict__ x, float *__restrict__ y, int C, int L) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *xb = x + (size_t)b * C * L;
    float *yb = y + (size_t)b * C * L;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int l = l0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && l < L) ? xb[(size_t)l * C + c] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, l = l0 + threadIdx.x;
        if (c < C && l < L) yb[(size_t)c * L + l] = tile[threadIdx.x][r];
    }
}

int launch_cl_to_cf_f32(const float *x, float *y, int B, int C, int L, cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div(L, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
    cl_to_cf_f32_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, y, C, L);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

// (B, L, Cld) 16-bit (first C channels) -> (B, C, L) fp32
__global__ void cl_16_to_cf_f32_kernel(const uint16_t *__restrict__ x, float *__restrict__ y,
                                       int C, int L, int Cld, int fmt) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const uint16_t *xb = x + (size_t)b * Cld * L;
    float *yb = y + (size_t)b * C * L;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int l = l0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && l < L) ? cvt16_to_f32(xb[(size_t)l * Cld + c], fmt) : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, l = l0 + threadIdx.x;
        if (c < C && l < L) yb[(size_t)c * L + l] = tile[threadIdx.x][r];
    }
}

int launch_cl_16_to_cf_f32(const uint16_t *x, float *y, int B, int C, int L, int Cld, int fmt,
                           cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div(L, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
    cl_16_to_cf_f32_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, y, C, L, Cld, fmt);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

}  // namespace vtts
