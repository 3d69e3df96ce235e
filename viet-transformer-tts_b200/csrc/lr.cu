// lr.cu -- LengthRegulator kernels (duration prefix-sum + vectorised frame gather).
//
// Replaces models/tts/fastspeech2/layers.py:434-462 (LengthRegulator.forward) and pad_list
// (models/tts/fastspeech2/function.py:97-124).  Integer/byte work only: results are bit-exact.
//
// HBM roofline: read B*Tmax*D*s + B*Tmax*8, write B*T_out*D*s bytes; each source row is read
// once from HBM (re-reads for repeated frames hit L1/L2), each output row is written once with
// 128-bit stores.  Grid = (frame tiles, B); every CTA re-scans its row's durations (<= a few
// KB, L2 resident) with a warp-shuffle prefix sum into shared memory, then each warp resolves
// frame -> token by binary search on the inclusive prefix sum and copies whole rows.
#include "common.cuh"

namespace vtts {

constexpr int LR_THREADS = 256;
constexpr int LR_WARPS = LR_THREADS / 32;
constexpr int LR_FRAMES_PER_CTA = 64;

__device__ __forceinline__ long long warp_inclusive_scan(long long v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// inclusive prefix sum of ds[0..Tmax) into cum[] (shared); returns the row total to all threads
__device__ long long block_prefix_sum(const long long *__restrict__ ds, int Tmax, long long *cum,
                                      long long *warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long carry = 0;
    for (int base = 0; base < Tmax; base += LR_THREADS) {
        int i = base + threadIdx.x;
        long long v = (i < Tmax) ? ds[i] : 0;
        long long s = warp_inclusive_scan(v, lane);
        if (lane == 31) warp_tot[warp] = s;
        __syncthreads();
        long long off = carry;
#pragma unroll
        for (int w = 0; w < LR_WARPS; ++w) {
            long long t = warp_tot[w];
            if (w < warp) off += t;
        }
        long long tile_total = 0;
#pragma unroll
        for (int w = 0; w < LR_WARPS; ++w) tile_total += warp_tot[w];
        if (i < Tmax) cum[i] = s + off;
        carry += tile_total;
        __syncthreads();
    }
    return carry;
}

// ---------------------------------------------------------------------------------------------
// layers.py:446-448
__global__ void lr_scale_kernel(const long long *__restrict__ ds, long long n, float alpha,
                                long long *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float f = __ll2float_rn(ds[i]);          // ds.float()
        f = __fmul_rn(f, alpha);                 // * alpha (fp32 op-math)
        out[i] = __float2ll_rz(rintf(f));        // torch.round (half-to-even) then .long()
    }
}

// layers.py:209 / :450 -- per-row sums, batch max/total, negative count
__global__ void lr_rowsum_kernel(const long long *__restrict__ ds, int Tmax,
                                 long long *__restrict__ mel_len, long long *__restrict__ stats) {
    __shared__ long long s_sum[LR_WARPS];
    __shared__ int s_neg[LR_WARPS];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long *row = ds + (long long)b * Tmax;
    long long s = 0;
    int neg = 0;
    for (int i = threadIdx.x; i < Tmax; i += blockDim.x) {
        long long v = row[i];
        s += v;
        neg += (v < 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        neg += __shfl_down_sync(0xffffffffu, neg, o);
    }
    if (lane == 0) { s_sum[warp] = s; s_neg[warp] = neg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        int n = 0;
        for (int w = 0; w < LR_WARPS; ++w) { t += s_sum[w]; n += s_neg[w]; }
        if (mel_len) mel_len[b] = t;
        atomicMax(&stats[0], t);
        atomicAdd(reinterpret_cast<unsigned long long *>(&stats[1]), (unsigned long long)t);
        if (n) atomicAdd(reinterpret_cast<unsigned long long *>(&stats[2]), (unsigned long long)n);
    }
}

// layers.py:458
__global__ void lr_fix_zero_rows_kernel(long long *__restrict__ ds, int Tmax) {
    __shared__ long long s_sum[LR_WARPS];
    __shared__ long long s_total;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long *row = ds + (long long)b * Tmax;
    long long s = 0;
    for (int i = threadIdx.x; i < Tmax; i += blockDim.x) s += row[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) s_sum[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < LR_WARPS; ++w) t += s_sum[w];
        s_total = t;
    }
    __syncthreads();
    if (s_total == 0)
        for (int i = threadIdx.x; i < Tmax; i += blockDim.x) row[i] = 1;
}

// layers.py:460-462.  VEC = bytes moved per lane per access (16, 8, 4, 2, 1).
template <typename V>
__global__ void __launch_bounds__(LR_THREADS)
lr_gather_kernel(const unsigned char *__restrict__ xs, const long long *__restrict__ ds,
                 unsigned char *__restrict__ out, int Tmax, long long row_bytes,
                 long long T_out, V pad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    long long *cum = reinterpret_cast<long long *>(smem_raw);
    __shared__ long long warp_tot[LR_WARPS];

    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long len = block_prefix_sum(ds + (long long)b * Tmax, Tmax, cum, warp_tot);

    const long long nvec = row_bytes / (long long)sizeof(V);
    const long long t0 = (long long)blockIdx.x * LR_FRAMES_PER_CTA;
    const unsigned char *xrow = xs + (long long)b * Tmax * row_bytes;
    unsigned char *orow = out + (long long)b * T_out * row_bytes;

    // Each warp owns LR_FRAMES_PER_CTA / LR_WARPS CONSECUTIVE frames and walks the tokens that cover them: a token row is
    // loaded once (one 512-byte chunk per pass, kept in registers) and stored once per frame it expands to - the reads
    // drop by the mean duration, and the kernel is bound by its streaming stores.
    constexpr int FPW = LR_FRAMES_PER_CTA / LR_WARPS;
    const long long ta = t0 + (long long)warp * FPW;
    long long tb = ta + FPW;
    if (tb > T_out) tb = T_out;
    if (ta >= tb) return;
    const long long live_end = tb < len ? tb : len;              // frames [ta, live_end) are copies, [live_end, tb) padding
    if (ta < live_end) {
        // j = #{i : cum[i] <= ta}  (upper bound); warp-uniform binary search in smem
        int lo = 0, hi = Tmax;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (cum[mid] <= ta) lo = mid + 1; else hi = mid;
        }
        long long t = ta;
        for (int j = lo; t < live_end; ++j) {
            long long te = cum[j];                                // frames [t, te) repeat token j (te == t: duration 0)
            if (te > live_end) te = live_end;
            if (te <= t) continue;
            const V *src = reinterpret_cast<const V *>(xrow + (long long)j * row_bytes);
            for (long long i0 = 0; i0 < nvec; i0 += 128) {        // 4 independent 128-bit requests in flight per lane
                const long long i = i0 + lane;
                V a = pad, c = pad, d = pad, e = pad;
                if (i < nvec) a = __ldg(src + i);
                if (i + 32 < nvec) c = __ldg(src + i + 32);
                if (i + 64 < nvec) d = __ldg(src + i + 64);
                if (i + 96 < nvec) e = __ldg(src + i + 96);
                for (long long tt = t; tt < te; ++tt) {
                    V *dst = reinterpret_cast<V *>(orow + tt * row_bytes);
                    if (i < nvec) __stcs(dst + i, a);
                    if (i + 32 < nvec) __stcs(dst + i + 32, c);
                    if (i + 64 < nvec) __stcs(dst + i + 64, d);
                    if (i + 96 < nvec) __stcs(dst + i + 96, e);
                }
            }
            t = te;
        }
    }
    for (long long t = (ta > live_end ? ta : live_end); t < tb; ++t) {
        V *dst = reinterpret_cast<V *>(orow + t * row_bytes);
        for (long long i = lane; i < nvec; i += 32) __stcs(dst + i, pad);
    }
}

template <typename V>
static int launch_gather(const void *xs, const int64_t *ds, void *out, int B, int Tmax,
                         long long row_bytes, int64_t T_out, V pad, cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div64(T_out, LR_FRAMES_PER_CTA), (unsigned)B);
    size_t smem = (size_t)Tmax * sizeof(long long);
    if (smem > 48 * 1024) {
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(lr_gather_kernel<V>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    lr_gather_kernel<V><<<grid, LR_THREADS, smem, stream>>>(
        (const unsigned char *)xs, (const long long *)ds, (unsigned char *)out, Tmax, row_bytes,
        (long long)T_out, pad);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

}  // namespace vtts

using namespace vtts;

extern "C" int vtts_lr_scale_durations(const int64_t *ds, int64_t n, float alpha, int64_t *out,
                                       vtts_stream_t stream) {
    VTTS_REQUIRE(ds && out && n >= 0, "vtts_lr_scale_durations: null pointer or negative n");
    VTTS_REQUIRE(alpha > 0.f, "vtts_lr_scale_durations: alpha must be > 0 (layers.py:447)");
    if (n == 0) return VTTS_OK;
    lr_scale_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(
        (const long long *)ds, (long long)n, alpha, (long long *)out);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

extern "C" int vtts_lr_rowsum(const int64_t *ds, int B, int Tmax, int64_t *mel_len,
                              int64_t *stats, vtts_stream_t stream) {
    VTTS_REQUIRE(stats, "vtts_lr_rowsum: stats is null");
    VTTS_REQUIRE(B >= 0 && Tmax >= 0, "vtts_lr_rowsum: negative shape");
    VTTS_CHECK_CUDA(cudaMemsetAsync(stats, 0, 3 * sizeof(int64_t), (cudaStream_t)stream));
    if (B == 0) return VTTS_OK;
    VTTS_REQUIRE(ds || Tmax == 0, "vtts_lr_rowsum: ds is null");
    lr_rowsum_kernel<<<B, LR_THREADS, 0, (cudaStream_t)stream>>>(
        (const long long *)ds, Tmax, (long long *)mel_len, (long long *)stats);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

extern "C" int vtts_lr_fix_zero_rows(int64_t *ds, int B, int Tmax, vtts_stream_t stream) {
    VTTS_REQUIRE(B >= 0 && Tmax >= 0, "vtts_lr_fix_zero_rows: negative shape");
    if (B == 0 || Tmax == 0) return VTTS_OK;
    VTTS_REQUIRE(ds, "vtts_lr_fix_zero_rows: ds is null");
    lr_fix_zero_rows_kernel<<<B, LR_THREADS, 0, (cudaStream_t)stream>>>((long long *)ds, Tmax);
    VTTS_CHECK_LAUNCH();
    return VTTS_OK;
}

extern "C" int vtts_lr_gather(const void *xs, const int64_t *ds, void *out, int B, int Tmax,
                              int D, int64_t T_out, int elem_size, const void *pad,
                              vtts_stream_t stream) {
    VTTS_REQUIRE(B >= 0 && Tmax >= 0 && D >= 0 && T_out >= 0, "vtts_lr_gather: negative shape");
    VTTS_REQUIRE(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8,
                 "vtts_lr_gather: elem_size must be 1, 2, 4 or 8 (got %d)", elem_size);
    if (B == 0 || T_out == 0 || D == 0) return VTTS_OK;
    VTTS_REQUIRE(xs && ds && out && pad, "vtts_lr_gather: null pointer");
    VTTS_REQUIRE((size_t)Tmax * 8 <= 200 * 1024, "vtts_lr_gather: Tmax %d too large for the "
                 "shared-memory prefix sum (max 25600)", Tmax);
    const long long row_bytes = (long long)D * elem_size;
    cudaStream_t st = (cudaStream_t)stream;
    // build a 16-byte pad pattern by repeating the element
    unsigned char pat[16];
    for (int i = 0; i < 16; ++i) pat[i] = ((const unsigned char *)pad)[i % elem_size];
    const uintptr_t align = (uintptr_t)xs | (uintptr_t)out | (uintptr_t)row_bytes;
    if (align % 16 == 0) {
        uint4 p; memcpy(&p, pat, 16);
        return launch_gather<uint4>(xs, ds, out, B, Tmax, row_bytes, T_out, p, st);
    } else if (align % 8 == 0) {
        uint2 p; memcpy(&p, pat, 8);
        return launch_gather<uint2>(xs, ds, out, B, Tmax, row_bytes, T_out, p, st);
    } else if (align % 4 == 0) {
        unsigned p; memcpy(&p, pat, 4);
        return launch_gather<unsigned>(xs, ds, out, B, Tmax, row_bytes, T_out, p, st);
    } else if (align % 2 == 0) {
        unsigned short p; memcpy(&p, pat, 2);
        return launch_gather<unsigned short>(xs, ds, out, B, Tmax, row_bytes, T_out, p, st);
    }
    unsigned char p = pat[0];
    return launch_gather<unsigned char>(xs, ds, out, B, Tmax, row_bytes, T_out, p, st);
}
