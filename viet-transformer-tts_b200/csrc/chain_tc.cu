// chain_tc.cu -- one launch per ResidualBlock of the narrow (32 / 64 channel) MRF stages.
//
// ResidualBlock.forward (models/gan_tts/hifigan/layers.py:83-98) is
//     for u in units:  x = convs2[u](lrelu(convs1[u](lrelu(x)))) + x          (conv1 dilated, conv2 dilation 1)
// This kernel runs the whole chain (6 convs for V1) of one block on a time tile without leaving the SM:
//
//   * TIME PHASES PACKED INTO THE MMA ROWS.  A 32-channel conv fills a quarter of a 128-row tcgen05.mma.  Here accumulator
//     row (p, co) is output channel co at time phase p (PH = 128 / C phases), accumulator column n is the time group:
//     position = d * (PH * n + p) (+ residue, see below).  The B operand (activations, K-major rows [position][channel]) is
//     kept as PH planes (plane r = positions with phase r), and for the input offset s' = PH * q + r the MMA reads plane r
//     shifted by q rows.  The A operand for offset s' is rows [tap(c + s'), tap(c + s' - 1), ... tap(c + s' - PH + 1)] x
//     C output channels: with the taps stored in REVERSED order and PH-1 zero blocks on either side that is simply a
//     128-row window of the weight array - the Toeplitz expansion costs no memory.  A k-tap conv takes k + PH - 1 MMA
//     slices of K = C instead of k slices that waste 128 - C rows: 11 taps at 32 channels run 3.1x faster.
//   * DILATION d: positions are split into (position div d) mod PH planes in position order, so a dilated conv is the same
//     computation with row shifts of q * d; accumulator column c then holds position d * PH * (c div d) + d * p + c mod d.
//   * THE RESIDUAL STREAM LIVES IN TMEM.  x is loaded once (fp32, channels-last) into accumulator X with tcgen05.st;
//     every conv2 accumulates straight onto it (x_new = x + conv2(...), biases are added when X is read); conv1 writes
//     accumulator XT.  After each conv the epilogue warps turn the accumulator into the next conv's operand planes
//     (tcgen05.ld fragment -> bias, LeakyReLU, 16-bit -> stmatrix into the swizzled planes, zero outside [0, L) as
//     every conv of the reference zero-pads its own input), in place: the conv that read the planes has completed.
//   * TWO TILES IN FLIGHT (TMEM: 2 x (X 128 + XT 128 columns)): the epilogue of one overlaps the MMAs of the other.
//   * Tile = 120 columns x PH positions; the block's receptive field (halo h = sum over convs of (k-1)/2 * d per side)
//     is recomputed: 480 - 2 * 60 = 360 valid of 480 positions for k = 11 at 32 channels.
//   * Weights: the whole block resident in shared memory when it fits (k = 3, 7 at 32 channels), else streamed per conv
//     through a TMA ring shared by both tiles.
//
// HBM traffic per element and block: 4 B in (x), 4-8 B out (MRF sum) - the unit-per-launch kernels moved 12-20 B per
// element and UNIT (3 units per block).
#include "chain_tc.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <vector>

namespace vtts {
namespace tc {

// Two tile geometries (template parameter NS = tile slots in flight):
//   NS = 2: two tiles of 120 accumulator columns (UMMA N = 128), the epilogue of one overlaps the MMAs of the other;
//   NS = 1: one tile of 240 columns (UMMA N = 256, all 512 TMEM columns): no overlap, but half the recomputed halo per valid
//           position and half the MMA count - wins where the halo is a large part of a 120-column tile (64 channels, k = 7 / 11).
constexpr int CH_G = 16;           // zero guard rows in front of / behind the data rows of an operand plane
__host__ __device__ constexpr int ch_cols(int ns) { return ns == 2 ? 120 : 240; }   // live accumulator columns per tile (divisible by the dilations)
__host__ __device__ constexpr int ch_n(int ns) { return ns == 2 ? 128 : 256; }      // UMMA N (the last 8 / 16 columns are scratch)
__host__ __device__ constexpr int ch_nr(int ns) { return CH_G + ch_n(ns) + CH_G; }  // rows per operand plane
constexpr int CH_EPI_WARPS = 16;   // 4 per TMEM lane quarter, each owns 32 accumulator columns
constexpr int CH_THREADS = (CH_EPI_WARPS + 3) * 32;   // + weight loader + one MMA issuer per tile slot
constexpr int CH_MAX_SLOTS = 8;    // weight ring slots
constexpr int CH_MAX_TRIM_BATCH = 256;
constexpr int CH_SMEM_MAX = 227 * 1024;

struct ChainParams {
    const float *x;
    float *cs, *out_x;
    uint16_t *out_a;
    const float *bias;             // [n_convs][C]
    int rd_cs, wr;
    float scale, slope, slope_out;
    int B, L, k, n_convs;
    int cd[CH_MAX_CONVS];          // dilation of conv i
    int cx[CH_MAX_CONVS];          // 1: conv i accumulates onto the residual accumulator X, 0: writes XT
    uint32_t cmagic[CH_MAX_CONVS]; // ceil(65536 / cd[i])
    int halo, V, t_tiles, total_items;
    int sps, spc, n_slots, resident, blocks_total;
    const long long *lens;
    int len_margin, len_rate;
    long long *trace;              // debug (VTTS_CHAIN_TRACE=1 in vtts_dbg_resblock_chain): clock64 stamps of block 0
};
constexpr int CH_TRACE_PAIRS = 12, CH_TRACE_STRIDE = 32;   // per (pair, slot): conv i -> 4 stamps; 24.. = init / final
#define CH_TRACE(pair_, s_, e_) do { if (p.trace && blockIdx.x == 0 && (pair_) < CH_TRACE_PAIRS) \
    p.trace[(((pair_) * 2) + (s_)) * CH_TRACE_STRIDE + (e_)] = clock64(); } while (0)

// max(v, v*slope) == LeakyReLU for 0 <= slope <= 1 (checked on the host)
__device__ __forceinline__ float lrelu_fast(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int C, int FMT, int NS>
__global__ void __launch_bounds__(CH_THREADS, 1)
chain_tc_kernel(const __grid_constant__ CUtensorMap tm_w, const ChainParams p) {
    constexpr int CH_COLS = ch_cols(NS), CH_N = ch_n(NS), CH_NR = ch_nr(NS);
    constexpr int CPW = CH_N / 4;                   // accumulator columns per epilogue warp
    constexpr int PH = 128 / C, LOGPH = PH == 4 ? 2 : 1;
    constexpr int ROWB = C * 2;                     // bytes per operand row (one position, C channels)
    constexpr int KSTEPS = C / 16;
    constexpr int W = PH * CH_COLS;                 // positions per tile
    constexpr int PLANEB = CH_NR * ROWB;
    constexpr int OPNDB = PH * PLANEB;              // operand planes of one tile slot
    constexpr int BLKB = C * ROWB;                  // one tap block of weights (C x C)
    extern __shared__ __align__(1024) uint8_t smem[];
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) { printf("vtts: smem base not 1024-byte aligned\n"); __trap(); }
    uint8_t *s_op = smem;
    uint8_t *s_w = smem + NS * OPNDB;
    const int slot_blocks = p.sps + PH - 1;
    const size_t wbytes = p.resident ? (size_t)p.blocks_total * BLKB : (size_t)p.n_slots * slot_blocks * BLKB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_w + wbytes);
    uint64_t *opnd_full = bars, *acc_full = bars + 2, *w_full = bars + 4, *w_empty = w_full + CH_MAX_SLOTS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_empty + CH_MAX_SLOTS);
    float *s_bias = reinterpret_cast<float *>(tmem_slot + 2);
    int *s_lim = reinterpret_cast<int *>(s_bias + CH_MAX_CONVS * C);
    int *s_ioff = s_lim + CH_MAX_TRIM_BATCH;
    // s_tab[conv i][phase][accumulator column]: operand row (plane * CH_NR + row) that column's value goes to when the
    // operand planes of conv i are built (source: x for i == 0, else the output of conv i-1)
    uint16_t *s_tab = reinterpret_cast<uint16_t *>(s_ioff + CH_MAX_TRIM_BATCH + 2);

    // (shuffle broadcast: tells the compiler the warp index is warp-uniform, so role-specific state can live in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    constexpr int WARP_TMA = CH_EPI_WARPS, WARP_MMA = CH_EPI_WARPS + 1;   // MMA issuers: warps WARP_MMA (slot 0), WARP_MMA + 1 (slot 1)
    const bool trimming = p.lens != nullptr;
    const int nsl = p.k + PH - 1;                   // MMA slices per conv
    const int ch_half = (p.k - 1) / 2;

    if (threadIdx.x == 0) {
        mbar_init(&opnd_full[0], CH_EPI_WARPS); mbar_init(&opnd_full[1], CH_EPI_WARPS);
        mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
        for (int s = 0; s < CH_MAX_SLOTS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], NS); }   // every issuer releases a stage
        fence_barrier_init();
    }
    if (warp == WARP_MMA) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    // operand planes start as zeros: the guard rows are never written again
    for (int i = threadIdx.x; i < NS * OPNDB / 16; i += CH_THREADS) reinterpret_cast<uint4 *>(s_op)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < p.n_convs * C; i += CH_THREADS) s_bias[i] = __ldg(p.bias + i);
    for (int e = threadIdx.x; e < p.n_convs * PH * CH_N; e += CH_THREADS) {
        const int col = e & (CH_N - 1), php = (e / CH_N) & (PH - 1), i = e / (CH_N * PH);
        const int d_src = (i == 0 || p.cx[i - 1]) ? 1 : p.cd[i - 1], d_dst = p.cd[i];
        const int n = col / d_src;
        const int tau = d_src * (PH * n + php) + (col - n * d_src);
        const int u = tau / d_dst, rho = tau - u * d_dst;
        const int row = (u & (PH - 1)) * CH_NR + CH_G + rho + d_dst * (u >> LOGPH);
        s_tab[e] = (uint16_t)(col < CH_COLS ? row : 0);
    }
    if (trimming)
        for (int i = threadIdx.x; i < p.B; i += CH_THREADS) {
            const long long lim = (__ldg(p.lens + i) + p.len_margin) * (long long)p.len_rate;
            s_lim[i] = lim > 0x7fffffffLL ? 0x7fffffff : (int)lim;
        }
    fence_proxy_async_smem();
    __syncthreads();
    if (trimming && threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < p.B; ++i) {
            s_ioff[i] = acc;
            const int lim = s_lim[i] < 0 ? 0 : s_lim[i];
            const int live = (lim + p.V - 1) / p.V;
            acc += live > p.t_tiles ? p.t_tiles : live;
        }
        s_ioff[p.B] = acc;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_items = trimming ? s_ioff[p.B] : p.total_items;
    const int G2 = NS * (int)gridDim.x;
    // this CTA's j-th item (NS = 2: pairs of neighbouring tiles share their halo lines in L2)
    auto item_li = [&](int j) { return (j / NS) * G2 + NS * (int)blockIdx.x + (j % NS); };
    // TMEM: NS = 2: slot s owns columns [256 s, 256 s + 256) = X (128) + XT (128); NS = 1: X = [0, 256), XT = [256, 512)
    auto acc_col = [&](int s, bool is_x) { return tmem_base + (uint32_t)(NS == 2 ? s * 256 + (is_x ? 0 : 128) : (is_x ? 0 : 256)); };
    struct Item { int b, T0; };
    auto locate = [&](int li, int &bhint) {
        Item it;
        if (trimming) {
            while (li >= s_ioff[bhint + 1]) ++bhint;
            it.b = bhint;
            it.T0 = (li - s_ioff[bhint]) * p.V - p.halo;
        } else {
            it.b = li / p.t_tiles;
            it.T0 = (li - it.b * p.t_tiles) * p.V - p.halo;
        }
        return it;
    };
    // weights are constants: the loader does not wait for the previous kernel
    if (warp != WARP_TMA) { grid_dep_wait(); grid_dep_launch(); }

    if (warp == WARP_TMA) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_w);
            if (p.resident) {
                mbar_arrive_expect_tx(&w_full[0], (uint32_t)(p.blocks_total * BLKB));
                for (int q = 0; q < p.blocks_total; ++q) tma_load_2d(s_w + (size_t)q * BLKB, &tm_w, &w_full[0], 0, q * C);
            } else {
                uint32_t slot = 0, par = 0;
                for (int j = 0; item_li(j) < n_items; j += NS)
                    for (int i = 0; i < p.n_convs; ++i)
                        for (int st = 0; st < p.spc; ++st) {
                            mbar_wait(&w_empty[slot], par ^ 1u);
                            const int rest = nsl - st * p.sps;
                            const int nb = (rest < p.sps ? rest : p.sps) + PH - 1;
                            const int gb0 = i * nsl + st * p.sps;
                            mbar_arrive_expect_tx(&w_full[slot], (uint32_t)(nb * BLKB));
                            for (int q = 0; q < nb; ++q)
                                tma_load_2d(s_w + ((size_t)slot * slot_blocks + q) * BLKB, &tm_w, &w_full[slot], 0, (gb0 + q) * C);
                            if (++slot == (uint32_t)p.n_slots) { slot = 0; par ^= 1u; }
                        }
            }
        }
    } else if (warp == WARP_MMA || (NS == 2 && warp == WARP_MMA + 1)) {
        // ===== MMA issuers: one warp per tile slot.  A single warp issued the MMAs of both tiles back to back and was
        // issue-bound (~95 cycles per N=128 MMA against 64 on the tensor pipe); two warps on two schedulers overlap their
        // issue overhead.  Each warp runs its loop converged with warp-uniform operands (a single-lane loop pays an R2UR
        // round trip per descriptor word); elect.sync inside the helpers picks the issuing lane.  Streamed weight stages are
        // shared by the two tiles of a pair: both issuers wait for a stage and both release it. =====
        const int s = warp - WARP_MMA;
        constexpr uint32_t idesc = make_idesc_16(128, CH_N, FMT);
        constexpr uint64_t ROW16 = ROWB >> 4, PLANE16 = PLANEB >> 4, BLK16 = BLKB >> 4;
        const uint64_t op_desc = make_smem_desc(smem_u32(s_op + (size_t)s * OPNDB), ROWB, 0);
        const uint64_t w_desc = make_smem_desc(smem_u32(s_w), ROWB, 0);
        uint32_t nf = 0u;
        uint32_t slot = 0, par = 0;
        if (p.resident) { mbar_wait(&w_full[0], 0); tc_fence_after(); }
        // first slice: s' = PH - 1 + half, descending
        const int sp0 = PH - 1 + ch_half;
        const int r0 = sp0 & (PH - 1), q0 = sp0 >> LOGPH;
        for (int j = 0; item_li(j) < n_items; j += NS) {
            const bool has = s == 0 || item_li(j + 1) < n_items;   // an odd last pair: slot 1 only keeps the ring in step
            if (!has && p.resident) break;
            for (int i = 0; i < p.n_convs; ++i) {
                const uint64_t dstep = (uint64_t)p.cd[i] * ROW16;
                if (has) {
                    mbar_wait(&opnd_full[s], nf & 1u);
                    ++nf;
                    tc_fence_after();
                    if (lane == 0) CH_TRACE(j / NS, s, i * 4 + 2);
                }
                const uint32_t tmem_d = acc_col(s, p.cx[i] != 0);
                uint32_t acc = p.cx[i] ? 1u : 0u;
                uint64_t b = op_desc + (uint64_t)r0 * PLANE16 + (uint64_t)CH_G * ROW16 + (uint64_t)q0 * dstep;
                int r = r0;
                for (int st = 0; st < p.spc; ++st) {
                    uint64_t a;
                    if (p.resident) {
                        a = w_desc + (uint64_t)(i * nsl + st * p.sps) * BLK16;
                    } else {
                        mbar_wait(&w_full[slot], par);
                        tc_fence_after();
                        a = w_desc + (uint64_t)slot * (uint64_t)slot_blocks * BLK16;
                    }
                    if (has) {
                        const int rest = nsl - st * p.sps;
                        const int ns = rest < p.sps ? rest : p.sps;
                        for (int ul = 0; ul < ns; ++ul) {
                            umma_slice_warp<KSTEPS>(tmem_d, a, b, idesc, acc);
                            acc = 1u;
                            a += BLK16;
                            if (r == 0) { r = PH - 1; b += (uint64_t)(PH - 1) * PLANE16; b -= dstep; }
                            else { --r; b -= PLANE16; }
                        }
                    }
                    if (!p.resident) {
                        if (has) umma_commit_elect(&w_empty[slot]);
                        else if (lane == 0) mbar_arrive(&w_empty[slot]);
                        if (++slot == (uint32_t)p.n_slots) { slot = 0; par ^= 1u; }
                    }
                }
                if (has) {
                    umma_commit_elect(&acc_full[s]);
                    if (lane == 0) CH_TRACE(j / NS, s, i * 4 + 3);
                }
            }
        }
        __syncwarp();
    } else if (warp < CH_EPI_WARPS) {
        // ===== epilogue warps: quarter q of the TMEM lanes = rows 32q .. 32q+31 = (phase, channel) =====
        // (instruction-lean on purpose: 16 warps x 6 convs per tile made this kernel issue-bound - interior tiles take
        // paths without per-element bounds checks, the position -> operand-row map comes from a table)
        const int quarter = warp & 3, part = warp >> 2;
        const int ph = C == 32 ? quarter : (quarter >> 1);
        const int chq = C == 32 ? 0 : (quarter & 1) * 32;          // first channel of this quarter
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const int col_lo = part * CPW;
        const int ncol = col_lo + CPW <= CH_COLS ? CPW : CH_COLS - col_lo;   // live accumulator columns of this warp
        const int fc = (lane & 3) * 2;                              // fragment: first column
        const int mrow = lane & 7;                                  // stmatrix: operand row this thread addresses
        const uint32_t mchunk = (uint32_t)(chq / 8 + (lane >> 3));  // 16-byte chunk (8 channels) of matrix lane/8
        const uint32_t op_base[2] = {smem_u32(s_op), smem_u32(s_op + (NS - 1) * OPNDB)};
        const int bfr = chq + (lane >> 2);                          // fragment row -> channel (rows fr, fr+8, fr+16, fr+24)
        uint32_t na[2] = {0u, 0u};
        Item cur[2] = {{0, 0}, {0, 0}};
        bool has[2] = {false, false};
        int bhint = 0;

        // accumulator (column col, this warp's phase) -> tile-relative position, for an accumulator written by a conv of
        // dilation d (d == 1 also describes the residual accumulator X)
        auto tau_of = [&](int col, int d, uint32_t magic) {
            const int n = (int)(((uint32_t)col * magic) >> 16);
            return d * (PH * n + ph) + (col - n * d);
        };
        // TMEM accumulator -> operand planes of conv `nxt` (s_tab[nxt]: accumulator column -> operand row)
        auto build_operand = [&](int s, uint32_t t_acc, int d_src, uint32_t m_src, const float *bias4, int nxt, const Item &it) {
            const bool edge = it.T0 < 0 || it.T0 + W > p.L;
            const float b0 = bias4 ? bias4[bfr] : 0.f, b1 = bias4 ? bias4[bfr + 8] : 0.f;
            const float b2 = bias4 ? bias4[bfr + 16] : 0.f, b3 = bias4 ? bias4[bfr + 24] : 0.f;
            const uint16_t *tab = s_tab + (nxt * PH + ph) * CH_N + col_lo + mrow;
            const float slope = p.slope;
#pragma unroll
            for (int g8 = 0; g8 < CPW / 8; ++g8) {
                if (g8 * 8 >= ncol) break;
                const int col = col_lo + g8 * 8;
                uint32_t ra[4], rb[4];
                tmem_ld_16x256_x1(t_acc + (uint32_t)col, ra);
                tmem_ld_16x256_x1(t_acc + (16u << 16) + (uint32_t)col, rb);
                const uint32_t off = (uint32_t)tab[g8 * 8] * (uint32_t)ROWB;
                const uint32_t swz = ROWB == 128 ? ((off >> 7) & 7u) : ((off >> 7) & 3u);
                const uint32_t dst = op_base[s] + off + ((mchunk ^ swz) << 4);
                tmem_ld_wait();
                float v[8];
                v[0] = lrelu_fast(__uint_as_float(ra[0]) + b0, slope); v[1] = lrelu_fast(__uint_as_float(ra[1]) + b0, slope);
                v[2] = lrelu_fast(__uint_as_float(ra[2]) + b1, slope); v[3] = lrelu_fast(__uint_as_float(ra[3]) + b1, slope);
                v[4] = lrelu_fast(__uint_as_float(rb[0]) + b2, slope); v[5] = lrelu_fast(__uint_as_float(rb[1]) + b2, slope);
                v[6] = lrelu_fast(__uint_as_float(rb[2]) + b3, slope); v[7] = lrelu_fast(__uint_as_float(rb[3]) + b3, slope);
                if (edge) {                                         // every conv zero-pads its own input
                    const int t0 = it.T0 + tau_of(col + fc, d_src, m_src), t1 = it.T0 + tau_of(col + fc + 1, d_src, m_src);
                    if (t0 < 0 || t0 >= p.L) v[0] = v[2] = v[4] = v[6] = 0.f;
                    if (t1 < 0 || t1 >= p.L) v[1] = v[3] = v[5] = v[7] = 0.f;
                }
                stmatrix_x4_trans(dst, cvt16x2(v[0], v[1], FMT), cvt16x2(v[2], v[3], FMT), cvt16x2(v[4], v[5], FMT),
                                  cvt16x2(v[6], v[7], FMT));
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&opnd_full[s]);
        };
        // x tile (fp32 channels-last) -> residual accumulator X of slot s; thread = TMEM lane = (phase, channel),
        // register c = accumulator column col_lo + c = position T0 + PH * (col_lo + c) + ph
        auto init_x = [&](int s, const Item &it) {
#pragma unroll 1
            for (int h = 0; h < CPW / 32; ++h) {
                const int c0 = col_lo + 32 * h;                       // this pass: accumulator columns c0 .. c0 + 31
                const int nc32 = ncol - 32 * h < 32 ? ncol - 32 * h : 32;
                if (nc32 <= 0) break;
                const float *xb = p.x + ((long long)it.b * p.L + it.T0 + ph + PH * c0) * C + chq + lane;
                uint32_t v[32];
                if (it.T0 >= 0 && it.T0 + W <= p.L && nc32 == 32) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(ld_stream_f32(xb + c * (PH * C)));
                } else {
                    // live columns: c < nc32 and 0 <= T0 + PH * (c0 + c) + ph < L
                    const int base_t = it.T0 + ph + PH * c0;
                    int c_lo = base_t >= 0 ? 0 : (-base_t + PH - 1) >> LOGPH;
                    int c_hi = (p.L - base_t + PH - 1) >> LOGPH;
                    c_hi = p.L <= base_t ? 0 : (c_hi < nc32 ? c_hi : nc32);
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = (c >= c_lo && c < c_hi) ? __float_as_uint(ld_stream_f32(xb + c * (PH * C))) : 0u;
                }
                tmem_st_32x32(acc_col(s, true) + lane_off + (uint32_t)c0, v);
            }
            tmem_st_wait();
        };
        // residual accumulator X of slot s -> global (the block's output, combined into the MRF sum)
        auto final_out = [&](int s, const Item &it, const float *bias_last) {
          const float bv = bias_last[chq + lane];
#pragma unroll 1
          for (int h = 0; h < CPW / 32; ++h) {
            const int c0 = col_lo + 32 * h;                           // this pass: accumulator columns c0 .. c0 + 31
            const int nc32 = ncol - 32 * h < 32 ? ncol - 32 * h : 32;
            if (nc32 <= 0) break;
            // valid columns: tile-relative position PH * (c0 + c) + ph in [halo, min(halo + V, L - T0))
            const int hi_tau = p.halo + p.V < p.L - it.T0 ? p.halo + p.V : p.L - it.T0;
            const int rel = ph + PH * c0;
            int c_lo = p.halo <= rel ? 0 : (p.halo - rel + PH - 1) >> LOGPH;
            int c_hi = hi_tau <= rel ? 0 : (hi_tau - rel + PH - 1) >> LOGPH;
            if (c_hi > nc32) c_hi = nc32;
            if (c_lo >= c_hi) continue;
            uint32_t v[32];
            tmem_ld_32x32(acc_col(s, true) + lane_off + (uint32_t)c0, v);
            const long long base = ((long long)it.b * p.L + it.T0 + rel) * C + chq + lane;
            const bool full = c_lo == 0 && c_hi == 32;
            constexpr int CS = PH * C;                              // floats between two columns of one phase
            float old[32];
            if (p.rd_cs) {
                const float *cp = p.cs + base;
#pragma unroll
                for (int c = 0; c < 32; ++c) old[c] = (full || (c >= c_lo && c < c_hi)) ? ld_stream_f32(cp + c * CS) : 0.f;
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) old[c] = 0.f;
            }
            tmem_ld_wait();
            if (p.wr == 0) {
                float *o = p.cs + base;
                if (full) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) o[c * CS] = __uint_as_float(v[c]) + bv + old[c];
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) if (c >= c_lo && c < c_hi) o[c * CS] = __uint_as_float(v[c]) + bv + old[c];
                }
            } else if (p.wr == 1) {
                float *o = p.cs + base;
                if (full) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(o + c * CS), "f"(__uint_as_float(v[c]) + bv + old[c]) : "memory");
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (c >= c_lo && c < c_hi)
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(o + c * CS), "f"(__uint_as_float(v[c]) + bv + old[c]) : "memory");
                }
            } else {
                // (interior passes: straight-line stores, the output selection hoisted out of the element loop)
                float *ox = p.out_x ? p.out_x + base : nullptr;
                uint16_t *oa = p.out_a ? p.out_a + base : nullptr;
                float val[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) val[c] = (__uint_as_float(v[c]) + bv + old[c]) * p.scale;
                if (full) {
                    if (ox) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) ox[c * CS] = val[c];
                    }
                    if (oa) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) oa[c * CS] = cvt16(lrelu_fast(val[c], p.slope_out), FMT);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (c >= c_lo && c < c_hi) {
                            if (ox) ox[c * CS] = val[c];
                            if (oa) oa[c * CS] = cvt16(lrelu_fast(val[c], p.slope_out), FMT);
                        }
                    }
                }
            }
          }
        };
        // L2 prefetch of the x (and MRF-sum) ranges of this CTA's NEXT pair of tiles: their first touch is init_x, whose
        // latency nothing hides (both slots reach the end of their chains together)
        auto prefetch_pair = [&](int j) {
            int bh = bhint;
            for (int s = 0; s < NS; ++s) {
                const int li = item_li(j + s);
                if (li >= n_items) break;
                const Item it = locate(li, bh);
                const int t0 = it.T0 < 0 ? 0 : it.T0, t1 = it.T0 + W < p.L ? it.T0 + W : p.L;
                if (t1 <= t0) continue;
                const long long o = ((long long)it.b * p.L + t0) * C;
                bulk_prefetch_l2(p.x + o, (uint32_t)(t1 - t0) * C * 4u);
                if (p.rd_cs) bulk_prefetch_l2(p.cs + o, (uint32_t)(t1 - t0) * C * 4u);
            }
        };

        const int last = p.n_convs - 1;
        for (int j = 0;; j += NS) {
            const bool more = item_li(j) < n_items;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (has[s]) {                                       // the previous item of this slot: its last conv is done
                    mbar_wait_relaxed(&acc_full[s], na[s] & 1u);
                    ++na[s];
                    tc_fence_after();
                    if (warp == 0 && lane == 0) CH_TRACE(j / NS - 1, s, 26);
                    final_out(s, cur[s], s_bias + last * C);
                    if (warp == 0 && lane == 0) CH_TRACE(j / NS - 1, s, 27);
                    has[s] = false;
                }
                if (more && item_li(j + s) < n_items) {
                    cur[s] = locate(item_li(j + s), bhint);
                    has[s] = true;
                    if (warp == 0 && lane == 0) CH_TRACE(j / NS, s, 24);
                    init_x(s, cur[s]);
                    if (warp == 0 && lane == 0) CH_TRACE(j / NS, s, 25);
                    build_operand(s, acc_col(s, true) + lane_off, 1, 65536u, nullptr, 0, cur[s]);
                    if (warp == 0 && lane == 0) CH_TRACE(j / NS, s, 28);
                }
            }
            if (!more) break;
            if (warp == 0 && lane == 0) prefetch_pair(j + NS);
            for (int i = 0; i < last; ++i) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (!has[s]) continue;
                    mbar_wait_relaxed(&acc_full[s], na[s] & 1u);
                    ++na[s];
                    tc_fence_after();
                    const bool from_x = p.cx[i] != 0;
                    if (warp == 0 && lane == 0) CH_TRACE(j / NS, s, i * 4 + 0);
                    build_operand(s, acc_col(s, from_x) + lane_off, from_x ? 1 : p.cd[i],
                                  from_x ? 65536u : p.cmagic[i], s_bias + i * C, i + 1, cur[s]);
                    if (warp == 0 && lane == 0) CH_TRACE(j / NS, s, i * 4 + 1);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// weight packing: [global block gb][co][ci]; conv i occupies blocks [i * (k + PH - 1), ...) : PH - 1 zero blocks, then the
// taps in REVERSED order; one more run of PH - 1 zero blocks closes the array.  Effective biases: a conv that feeds the
// residual accumulator carries the sum of all residual-side biases so far (they are added when the accumulator is read).
// ---------------------------------------------------------------------------------------------
struct ChainPackArgs {
    const float *w[CH_MAX_CONVS];
    const float *bias[CH_MAX_CONVS];
    int cx[CH_MAX_CONVS];
};
__global__ void chain_pack_kernel(ChainPackArgs a, uint16_t *out0, uint16_t *out1, float *bias_out, int C, int k, int PH,
                                  int n_convs, int blocks_total) {
    const int per = k + PH - 1;
    const size_t total = (size_t)blocks_total * C * C;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int ci = (int)(idx % C);
        const int co = (int)((idx / C) % C);
        const int gb = (int)(idx / ((size_t)C * C));
        const int i = gb / per, beta = gb - i * per;
        float v = 0.f;
        if (i < n_convs && beta >= PH - 1) {
            const int j = (k - 1) - (beta - (PH - 1));
            v = a.w[i][((size_t)co * C + ci) * k + j];
        }
        out0[idx] = cvt16(v, VTTS_FMT_BF16);
        out1[idx] = cvt16(v, VTTS_FMT_FP16);
    }
    if (blockIdx.x == 0)
        for (int t = threadIdx.x; t < n_convs * C; t += blockDim.x) {
            const int i = t / C, c = t - i * C;
            float b = 0.f;
            if (a.cx[i]) {
                for (int i2 = 0; i2 <= i; ++i2)
                    if (a.cx[i2] && a.bias[i2]) b += a.bias[i2][c];
            } else if (a.bias[i]) {
                b = a.bias[i][c];
            }
            bias_out[t] = b;
        }
}

static int chain_ph(const ChainSpec &s) { return 128 / s.C; }

// shared-memory plan: resident weights when the whole block fits next to the operand planes, else a ring of
// (sps + PH - 1)-block stages with at least one stage of prefetch beyond a whole conv
struct ChainPlan { int sps, spc, n_slots, resident, blocks_total; size_t smem; bool ok; };
static ChainPlan chain_plan(const ChainSpec &s, int ns) {
    ChainPlan pl{};
    const int CH_NR = ch_nr(ns), CH_N = ch_n(ns);
    const int PH = chain_ph(s), rowb = s.C * 2;
    const int n_convs = s.n_units * (s.has2 ? 2 : 1);
    const int nsl = s.k + PH - 1;
    const size_t blkb = (size_t)s.C * rowb;
    const size_t opnd = (size_t)ns * PH * CH_NR * rowb;
    const size_t fixed = (4 + 2 * CH_MAX_SLOTS) * 8 + 16 + (size_t)CH_MAX_CONVS * s.C * 4 + (2 * CH_MAX_TRIM_BATCH + 2) * 4 +
                         (size_t)CH_MAX_CONVS * PH * CH_N * 2 + 64;
    pl.blocks_total = n_convs * nsl + PH - 1;
    if (opnd + fixed > (size_t)CH_SMEM_MAX) return pl;
    const size_t avail = (size_t)CH_SMEM_MAX - opnd - fixed;
    static int force_stream = -1;
    if (force_stream < 0) { const char *e = getenv("VTTS_CHAIN_STREAM"); force_stream = (e && e[0] == '1') ? 1 : 0; }
    if (!force_stream && (size_t)pl.blocks_total * blkb <= avail) {
        pl.resident = 1; pl.sps = nsl; pl.spc = 1; pl.n_slots = 1;
        pl.smem = opnd + (size_t)pl.blocks_total * blkb + fixed;
        pl.ok = true;
        return pl;
    }
    for (int sps = nsl; sps >= 1; --sps) {
        const int spc = (nsl + sps - 1) / sps;
        const size_t slotb = (size_t)(sps + PH - 1) * blkb;
        int n_slots = (int)(avail / slotb);
        if (n_slots > CH_MAX_SLOTS) n_slots = CH_MAX_SLOTS;
        // (the two MMA issuers walk the stages of a conv side by side and release each stage together: a ring of three
        // stages keeps one load in flight; it need not hold a whole conv)
        if (n_slots >= 3 || (n_slots >= 2 && (sps == 1 || ns == 1))) {
            pl.resident = 0; pl.sps = sps; pl.spc = spc; pl.n_slots = n_slots;
            pl.smem = opnd + (size_t)n_slots * slotb + fixed;
            pl.ok = true;
            return pl;
        }
    }
    return pl;
}

static bool chain_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("VTTS_TC_CHAIN"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

static bool chain_geometry_usable(const ChainSpec &s, int ns) {
    if (!chain_enabled()) return false;
    if (s.C != 32 && s.C != 64) return false;
    if (s.k < 1 || s.k % 2 == 0 || s.n_units < 1 || s.n_units > 3) return false;
    const int PH = chain_ph(s), half = (s.k - 1) / 2, cols = ch_cols(ns);
    const int qmax = (half + PH - 1) / PH;                                  // |q| of every slice: s' in [-half, half + PH - 1]
    int halo = 0;
    for (int u = 0; u < s.n_units; ++u) {
        const int d = s.dil[u];
        if (d < 1 || cols % d != 0 || qmax * d > CH_G) return false;
        // without the second conv the dilated conv itself would accumulate onto the residual accumulator, whose columns
        // are laid out for dilation 1: such blocks (vits2 ResBlock2, use_additional_convs=False) keep the per-unit kernels
        if (!s.has2 && d != 1) return false;
        halo += half * d + (s.has2 ? half : 0);
    }
    if (qmax > CH_G) return false;
    if (PH * cols - 2 * halo < PH * cols / 4) return false;                // recompute would dominate
    return chain_plan(s, ns).ok;
}

// Tile geometry for a block: 2 = two 120-column tiles in flight, 1 = one 240-column tile, 0 = the chain kernel cannot run it.
// Cost model per valid position (cycles, from the in-kernel trace): conv hand-over E = 1100 per 128 columns (TMEM read
// bound), M = 64 cycles per N=128 MMA, 5300 per tile of drain / reload / first build.  With two tiles in flight E and M
// overlap, with one wide tile they add up but the halo is paid once per 240 columns.  VTTS_CHAIN_NS=1|2 forces a geometry.
static int chain_pick_ns(const ChainSpec &s) {
    static int force = -1;
    if (force < 0) { const char *e = getenv("VTTS_CHAIN_NS"); force = e ? atoi(e) : 0; }
    const bool ok2 = chain_geometry_usable(s, 2), ok1 = chain_geometry_usable(s, 1);
    if (force == 1 || force == 2) return (force == 1 ? ok1 : ok2) ? force : (ok2 ? 2 : (ok1 ? 1 : 0));
    if (!ok1 || !ok2) return ok2 ? 2 : (ok1 ? 1 : 0);
    const int PH = chain_ph(s), half = (s.k - 1) / 2;
    const int n_convs = s.n_units * (s.has2 ? 2 : 1);
    int halo = 0;
    for (int u = 0; u < s.n_units; ++u) halo += half * s.dil[u] + (s.has2 ? half : 0);
    const double E = 1100.0, M = (double)(s.k + PH - 1) * (s.C / 16) * 64.0, Bd = 5300.0;
    const double v2 = PH * 120 - 2 * halo, v1 = PH * 240 - 2 * halo;
    const double c2 = (n_convs * (2 * E > 2.5 * M ? 2 * E : 2.5 * M) + 2 * Bd) / (2 * v2);
    const double c1 = (n_convs * (2 * M + 2 * E) + 2 * Bd) / v1;
    return c1 < 0.95 * c2 ? 1 : 2;
}

bool chain_spec_usable(const ChainSpec &s) { return chain_pick_ns(s) != 0; }

int chain_extra_reach(const ChainSpec &s) {
    const int PH = chain_ph(s);
    int e = 0;
    for (int u = 0; u < s.n_units; ++u) e += (PH - 1) * s.dil[u] + (s.has2 ? PH - 1 : 0);
    return e;
}

void chain_free(ChainWeights &cw) {
    cudaFree(cw.w16[0]); cudaFree(cw.w16[1]); cudaFree(cw.bias);
    cw.w16[0] = cw.w16[1] = nullptr; cw.bias = nullptr; cw.valid = false;
}

int chain_pack_raw(const ChainSpec &s, const float *const *w, const float *const *bias, ChainWeights &cw, cudaStream_t st) {
    if (!chain_spec_usable(s)) return set_error(VTTS_E_UNSUPPORTED, "chain: unsupported block shape");
    const int PH = chain_ph(s);
    const int n_convs = s.n_units * (s.has2 ? 2 : 1);
    const int blocks_total = n_convs * (s.k + PH - 1) + PH - 1;
    const size_t n = (size_t)blocks_total * s.C * s.C;
    if (cw.w16[0] && (cw.blocks_total != blocks_total || cw.spec.C != s.C)) chain_free(cw);
    for (int f = 0; f < 2; ++f)
        if (!cw.w16[f]) VTTS_CHECK_CUDA(cudaMalloc(&cw.w16[f], n * sizeof(uint16_t)));
    if (!cw.bias) VTTS_CHECK_CUDA(cudaMalloc(&cw.bias, (size_t)CH_MAX_CONVS * 64 * sizeof(float)));
    ChainPackArgs a{};
    for (int i = 0; i < n_convs; ++i) {
        a.w[i] = w[i];
        a.bias[i] = bias ? bias[i] : nullptr;
        a.cx[i] = s.has2 ? (i & 1) : 1;
    }
    int blocks = (int)((n + 255) / 256);
    if (blocks > 1024) blocks = 1024;
    chain_pack_kernel<<<blocks, 256, 0, st>>>(a, cw.w16[0], cw.w16[1], cw.bias, s.C, s.k, PH, n_convs, blocks_total);
    VTTS_CHECK_LAUNCH();
    cw.spec = s; cw.n_convs = n_convs; cw.blocks_total = blocks_total; cw.valid = true;
    return VTTS_OK;
}

static int chain_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <int C, int FMT, int NS>
static int chain_launch_t(const CUtensorMap &tm, const ChainParams &p, size_t smem, dim3 grid, bool pdl, cudaStream_t st) {
    static bool attr[64] = {};
    int dev = 0;
    VTTS_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        VTTS_CHECK_CUDA(cudaFuncSetAttribute(chain_tc_kernel<C, FMT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_MAX));
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    VTTS_CHECK_CUDA(launch_kernel_ex(chain_tc_kernel<C, FMT, NS>, grid, dim3(CH_THREADS), smem, st, pdl, 1u, tm, p));
    return VTTS_OK;
}

int chain_launch(const ChainWeights &cw, int fmt, const ChainRun &r, cudaStream_t st) {
    if (!cw.valid) return set_error(VTTS_E_STATE, "chain: weights not packed");
    const ChainSpec &s = cw.spec;
    const int PH = chain_ph(s);
    const int ns = chain_pick_ns(s);
    if (!ns) return set_error(VTTS_E_UNSUPPORTED, "chain: unsupported block shape");
    const ChainPlan pl = chain_plan(s, ns);
    if (!pl.ok) return set_error(VTTS_E_UNSUPPORTED, "chain: no shared-memory plan");
    if (r.slope < 0.f || r.slope > 1.f || r.slope_out < 0.f || r.slope_out > 1.f)
        return set_error(VTTS_E_UNSUPPORTED, "chain: LeakyReLU slope outside [0,1]");
    ChainParams p{};
    p.x = r.x; p.cs = r.cs; p.out_x = r.out_x; p.out_a = r.out_a; p.bias = cw.bias;
    p.rd_cs = r.rd_cs; p.wr = r.wr; p.scale = r.scale; p.slope = r.slope; p.slope_out = r.slope_out;
    p.B = r.B; p.L = r.L; p.k = s.k; p.n_convs = cw.n_convs;
    const int half = (s.k - 1) / 2;
    int halo = 0;
    for (int i = 0; i < cw.n_convs; ++i) {
        const int u = s.has2 ? i / 2 : i;
        const bool is2 = s.has2 && (i & 1);
        p.cd[i] = is2 ? 1 : s.dil[u];
        p.cx[i] = s.has2 ? (i & 1) : 1;
        p.cmagic[i] = (uint32_t)((65536 + p.cd[i] - 1) / p.cd[i]);
        halo += half * p.cd[i];
    }
    p.halo = halo;
    p.V = PH * ch_cols(ns) - 2 * halo;
    p.t_tiles = ceil_div(r.L, p.V);
    const long long total = (long long)p.t_tiles * r.B;
    if (total > 0x3fffffffLL) return set_error(VTTS_E_UNSUPPORTED, "chain: too many tiles");
    p.total_items = (int)total;
    p.sps = pl.sps; p.spc = pl.spc; p.n_slots = pl.n_slots; p.resident = pl.resident; p.blocks_total = pl.blocks_total;
    p.trace = r.trace;
    if (r.lens && r.B <= CH_MAX_TRIM_BATCH) { p.lens = r.lens; p.len_margin = r.len_margin; p.len_rate = r.len_rate; }
    CUtensorMap tm;
    {
        uint64_t dims[2] = {(uint64_t)s.C, (uint64_t)cw.blocks_total * s.C};
        uint64_t str[1] = {(uint64_t)s.C * 2};
        uint32_t box[2] = {(uint32_t)s.C, (uint32_t)s.C};
        int rc = make_tmap_bf16(&tm, cw.w16[fmt], 2, dims, str, box, s.C * 2);
        if (rc) return rc;
    }
    const int sms = chain_num_sms();
    const int pairs = (p.total_items + ns - 1) / ns;
    dim3 grid((unsigned)(pairs < sms ? pairs : sms));
#define VTTS_CHAIN_GO(CC, FF) (ns == 2 ? chain_launch_t<CC, FF, 2>(tm, p, pl.smem, grid, r.pdl, st) : chain_launch_t<CC, FF, 1>(tm, p, pl.smem, grid, r.pdl, st))
    if (s.C == 32) return fmt == VTTS_FMT_BF16 ? VTTS_CHAIN_GO(32, 0) : VTTS_CHAIN_GO(32, 1);
    return fmt == VTTS_FMT_BF16 ? VTTS_CHAIN_GO(64, 0) : VTTS_CHAIN_GO(64, 1);
#undef VTTS_CHAIN_GO
}

// ---------------------------------------------------------------------------------------------
// output conv on the channels-last fp32 stream: y[b, oc, t] = tanh(bias + sum_k sum_c w[k][c] * lrelu(x[b, t+k-h, c]))
// (generator.py:108-120; fp32 - the output conv dominates the 16-bit error budget).  HBM-bound: C*4 bytes in, 4 out.
// ---------------------------------------------------------------------------------------------
template <int C> struct PostW { float w[9 * C]; };   // [k][C], passed by value: FFMA reads them as constant-bank operands
template <int C>
__global__ void __launch_bounds__(256)
conv_post_cl_kernel(const float *__restrict__ x, const PostW<C> pw, const float *__restrict__ bias,
                    float *__restrict__ y, int L, int ksize, float slope, int out_channels, int oc,
                    const long long *__restrict__ lens, int len_margin, int len_rate) {
    constexpr int TT = 256, HMAX = 4, ROWS = TT + 2 * HMAX, PITCH = C + 1;
    extern __shared__ float s_t[];                 // [ROWS][C + 1]
    const int b = blockIdx.y, t0 = blockIdx.x * TT, h = (ksize - 1) / 2;
    grid_dep_wait();
    grid_dep_launch();
    const long long t_lim = lens != nullptr ? (lens[b] + len_margin) * (long long)len_rate : (long long)L;
    if ((long long)t0 >= t_lim) {
        const int t = t0 + threadIdx.x;
        if (t < L) y[((size_t)b * out_channels + oc) * L + t] = 0.f;
        return;
    }
    const float *xb = x + (size_t)b * L * C;
    // tile rows t0-4 .. t0+259; coalesced: consecutive threads read consecutive channels
    for (int m = threadIdx.x; m < ROWS * (C / 4); m += 256) {
        const int r = m / (C / 4), c4 = (m - r * (C / 4)) * 4;
        const int t = t0 - HMAX + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < L) {
            v = *reinterpret_cast<const float4 *>(xb + (size_t)t * C + c4);
            v.x = lrelu(v.x, slope); v.y = lrelu(v.y, slope); v.z = lrelu(v.z, slope); v.w = lrelu(v.w, slope);
        }
        float *d = s_t + r * PITCH + c4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= L) return;
    float acc = 0.f;
#pragma unroll
    for (int kk = 0; kk < 9; ++kk) {
        if (kk < ksize) {
            const float *row = s_t + (threadIdx.x + HMAX + kk - h) * PITCH;
#pragma unroll
            for (int c = 0; c < C; ++c) acc = fmaf(pw.w[kk * C + c], row[c], acc);
        }
    }
    const float bv = bias ? __ldg(bias + oc) : 0.f;
    y[((size_t)b * out_channels + oc) * L + t] = (long long)t < t_lim ? tanhf(acc + bv) : 0.f;
}

int launch_conv_post_cl(const float *x, const float *w_kc_host, const float *bias, float *y, int B, int C, int L, int ksize,
                        float slope, int out_channels, int oc, const long long *lens, int len_margin, int len_rate,
                        bool pdl, cudaStream_t st) {
    if (ksize > 9) return set_error(VTTS_E_UNSUPPORTED, "conv_post: kernel size %d > 9", ksize);
    dim3 grid((unsigned)ceil_div(L, 256), (unsigned)B);
    const size_t smem = (size_t)(256 + 8) * (C + 1) * sizeof(float);
#define VTTS_POSTCL(CC)                                                                                              \
    do {                                                                                                             \
        PostW<CC> pw{};                                                                                              \
        for (int i = 0; i < ksize * CC; ++i) pw.w[i] = w_kc_host[i];                                                 \
        if (smem > 48 * 1024)                                                                                        \
            VTTS_CHECK_CUDA(cudaFuncSetAttribute(conv_post_cl_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        VTTS_CHECK_CUDA(launch_kernel_ex(conv_post_cl_kernel<CC>, grid, dim3(256), smem, st, pdl, 1u, x, pw, bias, y, L, ksize, \
                                         slope, out_channels, oc, lens, len_margin, len_rate));                     \
    } while (0)
    if (C == 32) VTTS_POSTCL(32);
    else if (C == 64) VTTS_POSTCL(64);
    else return set_error(VTTS_E_UNSUPPORTED, "conv_post_cl: %d input channels", C);
#undef VTTS_POSTCL
    return VTTS_OK;
}

// (B, C, L) fp32 channels-first -> (B, L, C) fp32 channels-last (test hook / input staging only)
__global__ void cf_to_cl_f32_kernel(const float *__restrict__ x, float *__restrict__ y, int C, int L) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *xb = x + (size_t)b * C * L;
    float *yb = y + (size_t)b * C * L;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, l = l0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && l < L) ? xb[(size_t)c * L + l] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int l = l0 + r, c = c0 + threadIdx.x;
        if (l < L && c < C) yb[(size_t)l * C + c] = tile[threadIdx.x][r];
    }
}

}  // namespace tc
}  // namespace vtts

using namespace vtts;
using namespace vtts::tc;

// Test hook: one ResidualBlock through the chain kernel, channels-first fp32 in / out.
extern "C" int vtts_dbg_resblock_chain(const float *x, const void *const *w, const void *const *bias, float *y, int B, int C,
                                       int L, int k, const int *dil, int n_units, int has2, float slope, int fp16, int reps,
                                       float *ms_out, vtts_stream_t stream) {
    VTTS_REQUIRE(x && w && y && dil, "vtts_dbg_resblock_chain: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    ChainSpec s;
    s.C = C; s.k = k; s.n_units = n_units; s.has2 = has2;
    for (int u = 0; u < n_units && u < 3; ++u) s.dil[u] = dil[u];
    if (!chain_spec_usable(s)) return set_error(VTTS_E_UNSUPPORTED, "vtts_dbg_resblock_chain: block shape not supported by the chain kernel");
    const int n_convs = n_units * (has2 ? 2 : 1);
    const float *wp[CH_MAX_CONVS], *bp[CH_MAX_CONVS];
    for (int i = 0; i < n_convs; ++i) { wp[i] = (const float *)w[i]; bp[i] = bias ? (const float *)bias[i] : nullptr; }
    ChainWeights cw;
    float *xcl = nullptr, *ycl = nullptr;
    const size_t n = (size_t)B * C * L;
    VTTS_CHECK_CUDA(cudaMalloc(&xcl, n * 4));
    VTTS_CHECK_CUDA(cudaMalloc(&ycl, n * 4));
    int rc = chain_pack_raw(s, wp, bp, cw, st);
    if (!rc) {
        dim3 grid((unsigned)ceil_div(L, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
        cf_to_cl_f32_kernel<<<grid, dim3(32, 8), 0, st>>>(x, xcl, C, L);
        if (cudaGetLastError() != cudaSuccess) rc = set_error(VTTS_E_CUDA, "vtts_dbg_resblock_chain: layout kernel launch failed");
    }
    ChainRun r;
    r.x = xcl; r.cs = ycl; r.out_x = ycl; r.wr = 2; r.scale = 1.f; r.slope = slope; r.slope_out = slope; r.B = B; r.L = L;
    if (!rc) rc = chain_launch(cw, fp16 ? VTTS_FMT_FP16 : VTTS_FMT_BF16, r, st);
    if (!rc) rc = launch_cl_to_cf_f32(ycl, y, B, C, L, st);
    if (!rc && getenv("VTTS_CHAIN_TRACE")) {   // debug: per-conv pipeline stamps of block 0 (second launch: warm caches)
        const int n = CH_TRACE_PAIRS * 2 * CH_TRACE_STRIDE;
        long long *dt = nullptr;
        std::vector<long long> ht(n);
        cudaMalloc(&dt, n * sizeof(long long));
        cudaMemsetAsync(dt, 0, n * sizeof(long long), st);
        r.trace = dt;
        rc = chain_launch(cw, fp16 ? VTTS_FMT_FP16 : VTTS_FMT_BF16, r, st);
        r.trace = nullptr;
        cudaMemcpyAsync(ht.data(), dt, n * sizeof(long long), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        cudaFree(dt);
        const long long t0 = ht[24];
        for (int pr = 0; pr < CH_TRACE_PAIRS; ++pr)
            for (int sl = 0; sl < 2; ++sl) {
                const long long *e = ht.data() + (pr * 2 + sl) * CH_TRACE_STRIDE;
                if (!e[24]) continue;
                fprintf(stderr, "[chain-trace] pair %2d slot %d: init %lld ld+st %lld build0 %lld |", pr, sl, e[24] - t0, e[25] - e[24], e[28] - e[25]);
                for (int i = 0; i < n_convs; ++i)
                    fprintf(stderr, " c%d mma[wait@%lld issue %lld] epi[acc@%lld build %lld]", i, e[i * 4 + 2] - t0, e[i * 4 + 3] - e[i * 4 + 2],
                            i < n_convs - 1 ? e[i * 4 + 0] - t0 : e[26] - t0, i < n_convs - 1 ? e[i * 4 + 1] - e[i * 4 + 0] : e[27] - e[26]);
                fprintf(stderr, "\n");
            }
    }
    if (!rc && reps > 0 && ms_out) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        for (int i = 0; i < reps && !rc; ++i) rc = chain_launch(cw, fp16 ? VTTS_FMT_FP16 : VTTS_FMT_BF16, r, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_out = ms / (float)reps;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(xcl); cudaFree(ycl);
    chain_free(cw);
    if (!rc && e != cudaSuccess) return set_error(VTTS_E_CUDA, "vtts_dbg_resblock_chain: %s", cudaGetErrorString(e));
    return rc;
}
