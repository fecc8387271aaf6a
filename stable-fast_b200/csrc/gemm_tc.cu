// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M, N] = A[M, K] * W[N, K]^T   (fp16/bf16 operands, fp32 accumulation in TMEM)
//
// One 128 x 160 output tile per CTA.  Warp roles (64 + 32 * kEpiWarps threads):
//   warp 0      TMA producer: A tile (128 rows x 64 K) + W tile (160 rows x 64 K) per stage,
//               both landing in 128-byte-swizzled K-major shared memory;
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 per stage);
//   warps 2..   epilogue: tcgen05.ld the accumulator (thread = row), fuse bias / time-embedding
//               row bias / residual / GEGLU / QKV head scatter, store 16-byte vectors.
// For the convolution the A tile of filter tap (kh, kw) is a *shifted NHWC box*: the tile's
// 128 output pixels are a [box_n, box_h, W] block, so one 4-D TMA box load at
// (c, kw-1, h0*stride+kh-1, n0) is exactly the im2col slice, and TMA's out-of-bounds zero fill
// is the convolution's zero padding.  No im2col buffer, no index tables.
//
// Replaces: cudnn_convolution_bias(_add) (/root/reference/src/sfast/csrc/operators/cudnn/
// cudnn_convolution_impl.cc:890-987), cublas_lowp_linear(_add) (csrc/operators/cublas/
// cublas_gemm.cpp:798-853,900-948) and cutlass_linear_geglu (csrc/operators/cutlass/
// cutlass_dual_linear_kernel.cu:442-525).
#include "common.cuh"
#include "host.h"

#include <stdlib.h>
#include <string.h>

namespace sfb {

constexpr int BM = 128;
constexpr int BK = 64;
// epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, each converting half of the
// tile's columns -- the epilogue is latency-bound with a single warp per SM sub-partition)
#ifndef SFB_EPI_WARPS
#define SFB_EPI_WARPS 8
#endif
constexpr int kEpiWarps = SFB_EPI_WARPS;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kColSplit = kEpiWarps / 4;
static_assert(kEpiWarps == 4 || kEpiWarps == 8, "epilogue warps");
constexpr int kGemmThreads = 64 + kEpiThreads;
__device__ __forceinline__ void epi_bar() {
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
}

struct EpiArgs {
    int epi;
    int dtype;
    int M, N;
    void* out;
    int ldo;
    const float* bias;
    const float* rowbias;
    int rows_per_img;
    int ld_rowbias;
    const void* residual;
    int ldr;
    int geglu_n_out;
    void* q;
    void* k;
    void* vt;
    int heads, head_dim, which_base, seq, q_pitch, q_rows, k_rows, vt_rows, vt_pitch;
    // GroupNorm statistics of the tensor this GEMM writes, accumulated for up to two consumers
    float* gn_stats[2];
    int gn_cpg[2], gn_choff[2];
    int gn_groups, gn_rpi, gn_shard_stride;  // floats between two of the 8 accumulation shards
    // LayerNorm folded around the GEMM (see sfb200.h): producer side / consumer side
    float* rowstats_out;
    const float* ln_rowstats;
    const float* ln_colsum;
    float ln_eps;
    int ln_dim;
};

// LayerNorm(x) W^T == rstd * (x W'^T - mean * colsum(W')) + (beta W^T + b), W' = W * gamma.
// (mean, rstd) of row m from the (sum, sum of squares) its producer GEMMs accumulated.
__device__ __forceinline__ float2 ln_row_params(const EpiArgs& e, int m) {
    const float2 st = __ldcg(reinterpret_cast<const float2*>(e.ln_rowstats) + m);
    const float inv = 1.0f / (float)e.ln_dim;
    const float mean = st.x * inv;
    const float var = fmaxf(st.y * inv - mean * mean, 0.f);
    return make_float2(mean, rsqrtf(var + e.ln_eps));
}

struct GemmArgs {
    int a_mode;
    int nkb_total;  // K / 64
    int splits;
    int pdl;        // launched with programmatic stream serialization
    long long* dbg; // optional: per-CTA %globaltimer stamps (8 per CTA) for latency breakdowns
    int* split_sync; // [tiles][2] arrive / done counters of the fused split-K reduction (or null)
    // cluster split-K: the `splits` CTAs of one output tile form a cluster along grid.z and sum their
    // fp32 partial tiles through distributed shared memory (no workspace, no second kernel)
    int cluster_k;
    float* ws;
    // conv geometry
    int img_n, img_h, img_w, cpb /* cin / 64 */, conv_stride, box_h, box_n, tiles_per_img;
    int box_w, tiles_per_row;  // M tile = [box_n, box_h, box_w] pixels; box_w < img_w: 2-D patches
    int up_tiles;   // SFB_A_UPCONV2X: M tiles per output phase
    int up_ntiles;  // ... and N tiles per phase in the phase-concatenated weight matrix
    // thread-block cluster (cx along N: the cx CTAs of one M-tile each load 1/cx of the A tile and
    // multicast it; cy along M: the cy CTAs of one N-tile each load 1/cy of the weight tile)
    int cx, cy;
    int a_part_dim;  // conv: which box dim the A tile is split along (1 = w, 2 = h, 3 = n)
    int a_part_ext;  // extent of one part along that dim (output pixels / rows / images)
    EpiArgs e;
};

// ---------------------------------------------------------------------------------------
// epilogue building blocks (shared by the GEMM kernel and the split-K reduction kernel)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void add_bias8(const float* __restrict__ b, int n, float (&acc)[8]) {
    const float4 b0 = *reinterpret_cast<const float4*>(b + n);
    const float4 b1 = *reinterpret_cast<const float4*>(b + n + 4);
    acc[0] += b0.x; acc[1] += b0.y; acc[2] += b0.z; acc[3] += b0.w;
    acc[4] += b1.x; acc[5] += b1.y; acc[6] += b1.z; acc[7] += b1.w;
}

__device__ __forceinline__ uint4 pack8(const float (&acc)[8], int bf16) {
    uint4 o;
    o.x = pack2(acc[0], acc[1], bf16);
    o.y = pack2(acc[2], acc[3], bf16);
    o.z = pack2(acc[4], acc[5], bf16);
    o.w = pack2(acc[6], acc[7], bf16);
    return o;
}

__device__ __forceinline__ void add_res8(uint4 r, int dtype, float (&acc)[8]) {
    float2 f;
    f = unpack2(r.x, dtype); acc[0] += f.x; acc[1] += f.y;
    f = unpack2(r.y, dtype); acc[2] += f.x; acc[3] += f.y;
    f = unpack2(r.z, dtype); acc[4] += f.x; acc[5] += f.y;
    f = unpack2(r.w, dtype); acc[6] += f.x; acc[7] += f.y;
}

template <int BF16>
__device__ __forceinline__ float round16(float v) {
    if (BF16) return __bfloat162float(__float2bfloat16_rn(v));
    return __half2float(__float2half_rn(v));
}

// GroupNorm partial sums of 8 stored values (row m, columns n..n+7) into block-shared accumulators
// sacc[2 targets][2 images][groups][2] (images img0, img0+1; anything else goes straight to global
// memory).  Used by the split-K reduction kernel, where a thread owns one 8-column slice.
template <int BF16>
__device__ __forceinline__ void gn_accumulate8(const EpiArgs& e, int m, int n, const float (&acc)[8],
                                               float* sacc, int img0, int shard) {
    const int img = m / e.gn_rpi;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        if (!e.gn_stats[t]) continue;
        int g_run = (e.gn_choff[t] + n) / e.gn_cpg[t];
        float s = 0.f, ss = 0.f;
        auto flush = [&]() {
            float* d = (img == img0 || img == img0 + 1)
                ? sacc + ((t * 2 + (img - img0)) * e.gn_groups + g_run) * 2
                : e.gn_stats[t] + (size_t)shard * e.gn_shard_stride + ((size_t)img * e.gn_groups + g_run) * 2;
            atomicAdd(d, s);
            atomicAdd(d + 1, ss);
        };
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int g = (e.gn_choff[t] + n + i) / e.gn_cpg[t];
            if (g != g_run) { flush(); g_run = g; s = 0.f; ss = 0.f; }
            const float v = round16<BF16>(acc[i]);
            s += v; ss += v * v;
        }
        flush();
    }
}

// (sum, sum of squares) of 8 values as they will be stored (rounded to the 16-bit type)
__device__ __forceinline__ void row_stats8(const float (&acc)[8], int dtype, float& s, float& ss) {
    const uint4 p = pack8(acc, dtype);
    const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = unpack2(w[i], dtype);
        s += f.x + f.y;
        ss += f.x * f.x + f.y * f.y;
    }
}

// 8 consecutive output columns [n, n+8) of row m; `acc` already holds bias / row bias / residual.
template <int BF16>
__device__ __forceinline__ void epi_store8(const EpiArgs& e, int m, int n, const float (&acc)[8]) {
    if (e.epi == SFB_EPI_STORE) {
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.out) + (size_t)m * e.ldo + n) =
            pack8(acc, BF16);
    } else {  // SFB_EPI_QKV
        const int C = e.heads * e.head_dim;
        const int which = n / C + e.which_base;
        const int nn = n % C;
        const int h = nn / e.head_dim;
        const int d = nn % e.head_dim;
        const int b = m / e.seq;
        const int s = m % e.seq;
        const size_t bh = (size_t)b * e.heads + h;
        if (which == 0) {
            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.q) +
                                      (bh * e.q_rows + s) * e.q_pitch + d) = pack8(acc, BF16);
        } else if (which == 1) {
            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.k) +
                                      (bh * e.k_rows + s) * e.q_pitch + d) = pack8(acc, BF16);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                store1(e.vt, (bh * e.vt_rows + d + i) * e.vt_pitch + s, acc[i], BF16);
        }
    }
}

// GEGLU: value / gate (bias already added) of 8 output columns [nout, nout+8)
template <int BF16>
__device__ __forceinline__ void epi_geglu8(const EpiArgs& e, int m, int nout, const float (&v)[8],
                                           const float (&g)[8]) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = v[i] * gelu_erf_f(g[i]);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.out) + (size_t)m * e.ldo + nout) =
        pack8(o, BF16);
}

// origin (image, row, column) of conv M-tile `tile` (already reduced to one up-conv phase)
__device__ __forceinline__ void conv_tile_origin(const GemmArgs& a, int tile, int& n0, int& h0, int& w0) {
    if (a.box_n == 1) {
        n0 = tile / a.tiles_per_img;
        const int rem = tile - n0 * a.tiles_per_img;
        const int trow = rem / a.tiles_per_row;
        h0 = trow * a.box_h;
        w0 = (rem - trow * a.tiles_per_row) * a.box_w;
    } else {
        n0 = tile * a.box_n;
        h0 = 0;
        w0 = 0;
    }
}

// tile-local row r (0..127) of M-tile `tile` -> global row m; false if the row is padding
__device__ __forceinline__ bool tile_row_to_m(const GemmArgs& a, int tile, int r, int& m) {
    if (a.a_mode == SFB_A_MATRIX) {
        m = tile * BM + r;
        return m < a.e.M;
    }
    int phase = 0;
    if (a.a_mode == SFB_A_UPCONV2X) {
        phase = tile / a.up_tiles;
        tile -= phase * a.up_tiles;
    }
    int n0, h0, w0;
    conv_tile_origin(a, tile, n0, h0, w0);
    const int w = w0 + r % a.box_w;
    const int t = r / a.box_w;
    const int dh = t % a.box_h;
    const int dn = t / a.box_h;
    const int n = n0 + dn, h = h0 + dh;
    if (a.a_mode == SFB_A_UPCONV2X)  // source pixel (h, w) -> output pixel (2h + py, 2w + px)
        m = (n * 2 * a.img_h + 2 * h + (phase >> 1)) * 2 * a.img_w + 2 * w + (phase & 1);
    else
        m = (n * a.img_h + h) * a.img_w + w;
    return (n < a.img_n) && (h < a.img_h);
}

// ---------------------------------------------------------------------------------------
// split-K reduction + epilogue
// ---------------------------------------------------------------------------------------
// Sum the `splits` fp32 partials of 8 output columns of row m and run the fused epilogue on them.
// `n` is the output column (GEGLU: output column of the gated product).  Partials were written by
// other SMs: read them through L2 (ld.global.cg).
// `sum8(col, acc)` yields the summed partials of global columns [col, col+8) of row m.
template <int BN, int BF16, typename Sum8>
__device__ __forceinline__ void reduce_epilogue8(Sum8&& sum8, const EpiArgs& e, int m, int n,
                                                 float* gn_sacc = nullptr, int gn_img0 = 0,
                                                 int gn_shard = 0) {
    if (e.epi == SFB_EPI_GEGLU) {
        const int tile = n / (BN / 2);
        const int nv = tile * BN + (n - tile * (BN / 2));
        const int ng = nv + BN / 2;
        float v[8], g[8];
        sum8(nv, v);
        sum8(ng, g);
        if (e.ln_rowstats) {
            const float2 ln = ln_row_params(e, m);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] = ln.y * (v[i] - ln.x * e.ln_colsum[nv + i]);
                g[i] = ln.y * (g[i] - ln.x * e.ln_colsum[ng + i]);
            }
        }
        if (e.bias) {
            add_bias8(e.bias, nv, v);
            add_bias8(e.bias, ng, g);
        }
        epi_geglu8<BF16>(e, m, n, v, g);
    } else {
        float acc[8];
        sum8(n, acc);
        if (e.ln_rowstats) {
            const float2 ln = ln_row_params(e, m);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = ln.y * (acc[i] - ln.x * e.ln_colsum[n + i]);
        }
        if (e.bias) add_bias8(e.bias, n, acc);
        if (e.epi == SFB_EPI_STORE) {
            if (e.rowbias) add_bias8(e.rowbias + (size_t)(m / e.rows_per_img) * e.ld_rowbias, n, acc);
            if (e.residual)
                add_res8(*reinterpret_cast<const uint4*>(
                             reinterpret_cast<const uint16_t*>(e.residual) + (size_t)m * e.ldr + n),
                         BF16, acc);
        }
        if (e.rowstats_out) {
            float rs = 0.f, rss = 0.f;
            row_stats8(acc, BF16, rs, rss);
            atomicAdd(e.rowstats_out + 2 * (size_t)m, rs);
            atomicAdd(e.rowstats_out + 2 * (size_t)m + 1, rss);
        }
        if (gn_sacc) gn_accumulate8<BF16>(e, m, n, acc, gn_sacc, gn_img0, gn_shard);
        epi_store8<BF16>(e, m, n, acc);
    }
}

template <int BN, int BF16>
__device__ __forceinline__ void splitk_reduce8(const float* __restrict__ ws, int splits, const EpiArgs& e,
                                               int m, int n, float* gn_sacc = nullptr, int gn_img0 = 0,
                                               int gn_shard = 0) {
    auto sum8 = [&](int col, float (&acc)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        // batches of 6 partials with every load in flight before the first add: this kernel is
        // one L2 round trip per batch, not per partial
        constexpr int kU = 6;
        const float* p0 = ws + (size_t)m * e.N + col;
        const size_t stride = (size_t)e.M * e.N;
        for (int s0 = 0; s0 < splits; s0 += kU) {
            float4 a[kU], b[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (s0 + u < splits) {
                    const float* p = p0 + (size_t)(s0 + u) * stride;
                    a[u] = __ldcg(reinterpret_cast<const float4*>(p));
                    b[u] = __ldcg(reinterpret_cast<const float4*>(p + 4));
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (s0 + u < splits) {
                    acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
                    acc[4] += b[u].x; acc[5] += b[u].y; acc[6] += b[u].z; acc[7] += b[u].w;
                }
            }
        }
    };
    reduce_epilogue8<BN, BF16>(sum8, e, m, n, gn_sacc, gn_img0, gn_shard);
}

// Stand-alone split-K reduction kernel (the default: measured faster than the in-kernel variants,
// DESIGN.md section 4.1).  One thread per (row, 8 columns); convs that feed a GroupNorm skip it
// (defer_finish) and let that kernel sum the partials.
template <int BN, int BF16>
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const float* __restrict__ ws, int splits, const EpiArgs e) {
    pdl_launch_dependents();
    pdl_wait();
    const int ncols = (e.epi == SFB_EPI_GEGLU) ? e.geglu_n_out : e.N;
    const int groups = ncols / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = idx < (long long)e.M * groups;
    __shared__ float sacc[2 * 2 * 64 * 2];
    const bool gn = e.gn_stats[0] != nullptr && e.epi == SFB_EPI_STORE;
    int img0 = 0;
    if (gn) {
        for (int i = threadIdx.x; i < 2 * 2 * e.gn_groups * 2; i += blockDim.x) sacc[i] = 0.f;
        const long long first = (long long)blockIdx.x * blockDim.x;
        img0 = (int)(first / groups) / e.gn_rpi;
        __syncthreads();
    }
    if (active)
        splitk_reduce8<BN, BF16>(ws, splits, e, (int)(idx / groups), (int)(idx % groups) * 8,
                                 gn ? sacc : nullptr, img0, blockIdx.x & 7);
    if (gn) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * 2 * e.gn_groups; i += blockDim.x) {
            const int t = i / (2 * e.gn_groups), im = (i / e.gn_groups) & 1, g = i % e.gn_groups;
            const float a0 = sacc[i * 2], a1 = sacc[i * 2 + 1];
            if (e.gn_stats[t] && (a0 != 0.f || a1 != 0.f)) {
                float* d = e.gn_stats[t] + (size_t)(blockIdx.x & 7) * e.gn_shard_stride +
                           ((size_t)(img0 + im) * e.gn_groups + g) * 2;
                atomicAdd(d, a0);
                atomicAdd(d + 1, a1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// the GEMM kernel
// ---------------------------------------------------------------------------------------
// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA PAIR (cluster 1x2 along M) computes a 256 x BN
// tile with tcgen05.mma.cta_group::2 -- each CTA stages its own 128 A rows but only HALF of the
// weight tile, which cuts the shared-memory traffic per MMA (the 1-CTA kernel is smem-bandwidth
// bound: 36 KB written + 36 KB read per 320 MMA cycles at 128 B/clk).
template <int BN, int STAGES, int CG = 1>
struct GemmSmem {
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2 / CG;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kBiasOffset = kBarOffset + 256;              // fp32 [kBiasSlots][BN]
    // images per conv M-tile whose row bias is staged; 3 stages + 4 slots keeps 2 CTAs per SM
    static constexpr int kBiasSlots = 4;
    static constexpr int kRowMOffset = kBiasOffset + kBiasSlots * BN * 4;  // int [128]: row -> m
    static constexpr int kTotal = kRowMOffset + BM * 4 + 1024;                // + alignment slack
    // fp32 staging tile of the epilogue, aliased onto the (by then idle) pipeline stages; the
    // +4 float pad makes the thread-per-row float4 writes of phase A bank-conflict free
    static constexpr int kStagePitch = BN + 4;
    // QKV scatter tables (row -> (batch, position), column slice -> (q/k/v, offset)), also aliased
    // onto the idle stage buffers, behind the GroupNorm accumulators
    static constexpr int kQkvRowOffset = BM * kStagePitch * 4 + BM * 4 + 2 * 8 * (BN / 2 + 2) * 2 * 4;
    static constexpr int kQkvColOffset = kQkvRowOffset + BM * 8;
    static_assert(kQkvColOffset + (BN / 8) * 16 <= kBarOffset,
                  "staging tile + GroupNorm accumulators + QKV tables must fit in the stage buffers");
    static_assert(STAGES > 4 || 2 * (kTotal + 1024) <= 228 * 1024, "shallow configs must fit twice per SM");
    static_assert(BN == 160, "the epilogue's 32+32+16 TMEM load split assumes 80-column halves");
};

template <int BN, int STAGES, int BF16, int CG>
__global__ void __launch_bounds__(kGemmThreads, STAGES <= 4 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const GemmArgs args) {
    using L = GemmSmem<BN, STAGES, CG>;
    constexpr uint32_t kTmemCols = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * L::kABytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    float* sBias = reinterpret_cast<float*>(smem + L::kBiasOffset);
    int* sRowM = reinterpret_cast<int*>(smem + L::kRowMOffset);
    float* sStage = reinterpret_cast<float*>(smem);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // pair mode: the two CTAs of a cta_group::2 pair must be neighbours along the cluster's x axis,
    // so the M tiles run along grid.x there (grid = (m_tiles, n_tiles, splits), cluster (2,1,1))
    const int n_tile = CG == 2 ? blockIdx.y : blockIdx.x;
    const int m_tile = CG == 2 ? blockIdx.x : blockIdx.y;
    const int n_tiles_grid = CG == 2 ? gridDim.y : gridDim.x;
    const int split = blockIdx.z;
    long long* dbg = args.dbg ? args.dbg + 8 * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = globaltimer_ns();
    const int kb_begin = (int)(((long long)args.nkb_total * split) / args.splits);
    const int kb_end = (int)(((long long)args.nkb_total * (split + 1)) / args.splits);
    const int nkb = kb_end - kb_begin;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            // released by every CTA whose stage this CTA's multicasts write into (pair mode: by
            // the leader's multicast commit only)
            mbar_init(&empty_bar[i], CG == 2 ? 1 : args.cx + args.cy - 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    const bool mcast = (CG == 1) && args.cx * args.cy > 1;
    const bool clustered = (CG == 2) || mcast || args.cluster_k;
    // pair mode inside a larger (split-K) cluster: the pair is ranks (2j, 2j+1)
    const uint16_t pair_mask = (uint16_t)(0b11u << (clustered ? (cluster_ctarank() & ~1u) : 0u));
    const int cix = mcast ? (int)cluster_ctaid_x() : 0;
    const int ciy = CG == 2 ? (int)cluster_ctaid_x() : (mcast ? (int)cluster_ctaid_y() : 0);
    const bool leader = (CG == 1) || ciy == 0;  // pair mode: the even CTA issues every MMA
    // CTAs sharing this CTA's A tile (same M-tile: all cix) / weight tile (same N-tile: all ciy)
    const uint16_t mask_a = (uint16_t)(((1u << args.cx) - 1u) << (ciy * args.cx));
    uint16_t mask_b = 0;
    for (int y = 0; y < args.cy; ++y) mask_b |= (uint16_t)(1u << (y * args.cx + cix));
    if (warp == 1) {
        if (CG == 2) tmem_alloc_pair<kTmemCols>(tmem_slot);
        else tmem_alloc<kTmemCols>(tmem_slot);
    }
    // Everything up to each role's pdl_wait() overlaps the previous kernel's tail (programmatic
    // dependent launch); global memory produced by it is only touched after that wait.
    pdl_launch_dependents();
    tc_fence_before();
    if (clustered) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (dbg && threadIdx.x == 0) dbg[1] = globaltimer_ns();

    if (warp == 0) {
        if (lane == 0) {
            pdl_wait();
            int n0 = 0, h0 = 0, w0 = 0;
            // up-conv: output phase (py, px) of this M tile; its 2x2 taps sit at source offsets
            // (py - 1 + ty, px - 1 + tx).  The weight tile comes from the phase's row block.
            int up_py = 0, up_px = 0, b_ntile = n_tile;
            if (args.a_mode != SFB_A_MATRIX) {
                int mt = m_tile;
                if (args.a_mode == SFB_A_UPCONV2X) {
                    const int ph = m_tile / args.up_tiles;
                    mt = m_tile - ph * args.up_tiles;
                    up_py = ph >> 1; up_px = ph & 1;
                    b_ntile = ph * args.up_ntiles + n_tile;
                }
                conv_tile_origin(args, mt, n0, h0, w0);
            }
            // conv tap of K block kb -> (w, h) source offset of the A box
            auto tap_offset = [&](int tap, int& dw, int& dh) {
                if (args.a_mode == SFB_A_UPCONV2X) {
                    const int ty = tap >> 1, tx = tap & 1;
                    dw = up_px - 1 + tx;
                    dh = up_py - 1 + ty;
                } else {
                    const int kh = tap / 3, kw = tap - kh * 3;
                    dw = kw - 1;
                    dh = kh - 1;
                }
            };
            for (int i = 0; i < nkb; ++i) {
                const int stage = i % STAGES;
                const uint32_t phase = (i / STAGES) & 1;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const int kb = kb_begin + i;
                if (CG == 2) {
                    // pair mode: both CTAs' bytes are counted on the LEADER's barrier
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                    uint8_t* dA = sA + stage * L::kABytes;
                    uint8_t* dB = sB + stage * L::kBBytes;
                    const int b_row2 = (b_ntile * args.nkb_total + kb) * BN + ciy * (BN / 2);
                    if (args.a_mode == SFB_A_MATRIX) {
                        tma_load_2d_pair(dA, &tma_a, &full_bar[stage], kb * BK, m_tile * BM);
                    } else {
                        const int tap = kb / args.cpb;
                        const int cc = kb - tap * args.cpb;
                        int dw, dh;
                        tap_offset(tap, dw, dh);
                        tma_load_4d_pair(dA, &tma_a, &full_bar[stage], cc * BK, w0 * args.conv_stride + dw,
                                         h0 * args.conv_stride + dh, n0);
                    }
                    tma_load_2d_pair(dB, &tma_b, &full_bar[stage], 0, b_row2);
                    continue;
                }
                mbar_expect_tx(&full_bar[stage], L::kStageBytes);
                uint8_t* dstA = sA + stage * L::kABytes + cix * (L::kABytes / args.cx);
                uint8_t* dstB = sB + stage * L::kBBytes + ciy * (L::kBBytes / args.cy);
                // weights are pre-tiled in HBM: tile (n_tile, kb) is one contiguous BN x 64 block
                const int b_row = (b_ntile * args.nkb_total + kb) * BN + ciy * (BN / args.cy);
                if (args.a_mode == SFB_A_MATRIX) {
                    const int a_row = m_tile * BM + cix * (BM / args.cx);
                    if (args.cx > 1) tma_load_2d_mc(dstA, &tma_a, &full_bar[stage], kb * BK, a_row, mask_a);
                    else tma_load_2d(dstA, &tma_a, &full_bar[stage], kb * BK, a_row);
                } else {
                    const int tap = kb / args.cpb;
                    const int cc = kb - tap * args.cpb;
                    int dw, dh;
                    tap_offset(tap, dw, dh);
                    int c1 = w0 * args.conv_stride + dw, c2 = h0 * args.conv_stride + dh, c3 = n0;
                    if (args.cx > 1) {
                        const int off = cix * args.a_part_ext;
                        if (args.a_part_dim == 1) c1 += off * args.conv_stride;
                        else if (args.a_part_dim == 2) c2 += off * args.conv_stride;
                        else c3 += off;
                        tma_load_4d_mc(dstA, &tma_a, &full_bar[stage], cc * BK, c1, c2, c3, mask_a);
                    } else {
                        tma_load_4d(dstA, &tma_a, &full_bar[stage], cc * BK, c1, c2, c3);
                    }
                }
                if (args.cy > 1) tma_load_2d_mc(dstB, &tma_b, &full_bar[stage], 0, b_row, mask_b);
                else tma_load_2d(dstB, &tma_b, &full_bar[stage], 0, b_row);
                if (dbg && i == 0) dbg[2] = globaltimer_ns();
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            const uint32_t idesc = umma_idesc_f16(BM * CG, BN, BF16 != 0);
            for (int i = 0; i < nkb; ++i) {
                const int stage = i % STAGES;
                const uint32_t phase = (i / STAGES) & 1;
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (dbg && i == 0) dbg[3] = globaltimer_ns();
                const uint64_t da = umma_desc_k_sw128(smem_u32(sA + stage * L::kABytes));
                const uint64_t db = umma_desc_k_sw128(smem_u32(sB + stage * L::kBBytes));
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // +32 bytes along K inside the 128-byte swizzle atom = +2 in the >>4 field
                    if (CG == 2)
                        umma_f16_ss_pair(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                                         (i | k) != 0);
                    else
                        umma_f16_ss(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                                    (i | k) != 0);
                }
                if (CG == 2) umma_commit_pair(&empty_bar[stage], pair_mask);
                else if (mcast) umma_commit_mc(&empty_bar[stage], (uint16_t)(mask_a | mask_b));
                else umma_commit(&empty_bar[stage]);
            }
            if (CG == 2) umma_commit_pair(tmem_full_bar, pair_mask);
            else umma_commit(tmem_full_bar);
            if (dbg) dbg[4] = globaltimer_ns();
        }
        __syncwarp();
    } else {
        pdl_wait();
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int et = threadIdx.x - 64;  // index among the epilogue threads
        const int chalf = (warp - 2) >> 2;  // which part of the tile's columns this warp converts
        int m;
        const bool valid = tile_row_to_m(args, m_tile, r, m);
        sRowM[r] = valid ? m : -1;
        if (args.splits == 1) {
            // stage bias (+ per-image time-embedding row bias) for this tile's columns in smem;
            // only the four epilogue warps take part (named barrier 1), the TMA / MMA warps are
            // already streaming
            const EpiArgs& e = args.e;
            int img0 = 0, nslots = 1;
            const bool rb_staged = e.rowbias && args.box_n <= L::kBiasSlots;
            if (rb_staged) {
                nslots = args.box_n;
                img0 = (args.box_n == 1) ? (m_tile / args.tiles_per_img) : m_tile * args.box_n;
            }
            for (int i = et; i < nslots * BN; i += kEpiThreads) {
                const int slot = i / BN, c = i - slot * BN;
                const int n = n_tile * BN + c;
                float v = 0.f;
                if (n < e.N) {
                    if (e.bias) v = e.bias[n];
                    if (rb_staged && img0 + slot < args.img_n)
                        v += e.rowbias[(size_t)(img0 + slot) * e.ld_rowbias + n];
                }
                sBias[i] = v;
            }
            if (e.ln_rowstats) {  // slot 1: column sums of the gamma-scaled weight
                for (int c = et; c < BN; c += kEpiThreads) {
                    const int n = n_tile * BN + c;
                    sBias[BN + c] = (n < e.N) ? e.ln_colsum[n] : 0.f;
                }
            }
        }
        epi_bar();
        // ---- epilogue.  Phase A: thread = accumulator row (warp w may only touch TMEM lanes
        // [32*(w%4), +32)): TMEM -> registers -> (+bias / LayerNorm fold) -> fp32 staging tile in the
        // now-idle pipeline buffers.  Phase B: threads re-partition the tile so that every global
        // access (residual load, output store) is a coalesced 16-byte slice of a row segment
        // instead of 32 rows x 16 bytes per warp instruction.
        const EpiArgs& e = args.e;
        const int ncol0 = n_tile * BN;
        const bool partial = args.splits > 1;
        constexpr int kGroups = BN / 8;  // 16-byte output slices per row
        constexpr int kItems = BM * kGroups / kEpiThreads;  // slices per epilogue thread
        // Residual in the coalesced phase-B ownership (thread <-> 16-byte slice), software-pipelined
        // in batches of 5 slices; the first batch is issued before the accumulator is even ready.
        // (Loops here are deliberately ROLLED: the epilogue runs once per CTA, so straight-line
        // unrolled code is all instruction-cache misses -- measured 5 us per tile on B200.)
        const bool has_res = (e.residual != nullptr) && (e.epi == SFB_EPI_STORE) && !partial;
        constexpr int kBatch = 5;
        auto item_addr = [&](int it, int& row, int& grp, int& mm, int& n) {
            const int idx = et + it * kEpiThreads;
            row = idx / kGroups;
            grp = idx - row * kGroups;
            mm = sRowM[row];
            n = ncol0 + grp * 8;
        };
        auto load_res = [&](int it) -> uint4 {
            int row, grp, mm, n;
            item_addr(it, row, grp, mm, n);
            return (mm >= 0 && n < e.N)
                ? *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(e.residual) + (size_t)mm * e.ldr + n)
                : make_uint4(0, 0, 0, 0);
        };
        uint4 rcur[kBatch], rnxt[kBatch];
        if (has_res) {
#pragma unroll
            for (int j = 0; j < kBatch; ++j) rcur[j] = load_res(j);
        }
        float2 ln = make_float2(0.f, 1.f);
        if (e.ln_rowstats && valid && !partial) ln = ln_row_params(e, m);
        const float* brow = sBias;
        const bool rb_staged = e.rowbias && args.box_n <= L::kBiasSlots;
        if (rb_staged) brow += (r / (args.box_h * args.img_w)) * BN;
        // many tiny images per tile (4x4 feature maps): row bias straight from global memory
        const float* rb_global = nullptr;
        if (e.rowbias && !rb_staged && !partial)
            rb_global = e.rowbias + (size_t)(m_tile * args.box_n + r / (args.box_h * args.img_w)) * e.ld_rowbias;

        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        if (dbg && threadIdx.x == 64) dbg[5] = globaltimer_ns();
        // QKV scatter: the integer divisions of the address computation once per row / per column
        // slice (tables in the now idle stage buffers) instead of five per stored 16-byte slice
        int2* sQkvRow = reinterpret_cast<int2*>(smem + L::kQkvRowOffset);
        longlong2* sQkvCol = reinterpret_cast<longlong2*>(smem + L::kQkvColOffset);
        if (args.e.epi == SFB_EPI_QKV && !partial) {
            const EpiArgs& q = args.e;
            if (chalf == 0) sQkvRow[r] = valid ? make_int2(m / q.seq, m % q.seq) : make_int2(0, 0);
            if (et < BN / 8) {
                const int n = ncol0 + et * 8;
                const int C = q.heads * q.head_dim;
                const int which = n / C + q.which_base;
                const int nn = n % C;
                const int h = nn / q.head_dim, d = nn % q.head_dim;
                long long off;
                if (which == 0) off = (long long)h * q.q_rows * q.q_pitch + d;
                else if (which == 1) off = (long long)h * q.k_rows * q.q_pitch + d;
                else off = ((long long)h * q.vt_rows + d) * q.vt_pitch;
                sQkvCol[et] = make_longlong2(which, off);
            }
        }
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float* srow = sStage + r * L::kStagePitch;
        constexpr int kColsPer = BN / kColSplit;             // columns converted per thread
        constexpr int kChunk = kColSplit == 1 ? 32 : 16;     // columns per TMEM load
        // accumulator chunk (already in registers) -> bias / LayerNorm fold -> fp32 staging tile
        auto stage_chunk = [&](const uint32_t* v, int c0) {
#pragma unroll
            for (int j = 0; j < kChunk / 8; ++j) {
                const int cl = c0 + j * 8;  // column inside the tile
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float acc = __uint_as_float(v[j * 8 + i]);
                    if (!partial) {
                        if (e.ln_rowstats) acc = ln.y * (acc - ln.x * sBias[BN + cl + i]);
                        acc += brow[cl + i];
                    }
                    f[i] = acc;
                }
                if (rb_global && ncol0 + cl < e.N) add_bias8(rb_global, ncol0 + cl, f);
                *reinterpret_cast<float4*>(srow + cl) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(srow + cl + 4) = make_float4(f[4], f[5], f[6], f[7]);
            }
        };
        constexpr int kChunks = kColsPer / kChunk;
        // (software-pipelining these TMEM reads -- next chunk in flight while this one is converted --
        // was measured: no gain, 4.87 vs 4.83 ms per step)
#pragma unroll 1
        for (int cb = 0; cb < kChunks; ++cb) {
            uint32_t v[kChunk];
            const int c0 = chalf * kColsPer + cb * kChunk;
            if constexpr (kChunk == 32) tmem_ld32(trow + c0, v);
            else tmem_ld16(trow + c0, v);
            tmem_wait_ld();
            stage_chunk(v, c0);
        }
        epi_bar();

        // ---- phase B
        auto load8 = [&](int row, int col, float (&f)[8]) {
            const float4 a = *reinterpret_cast<const float4*>(sStage + row * L::kStagePitch + col);
            const float4 b = *reinterpret_cast<const float4*>(sStage + row * L::kStagePitch + col + 4);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        };
        if (partial && args.cluster_k) {
            // the reduction happens after the cluster barrier below (all warps take part in it)
        } else if (partial) {
#pragma unroll 1
            for (int it = 0; it < kItems; ++it) {
                const int idx = et + it * kEpiThreads;
                const int row = idx / kGroups, grp = idx - row * kGroups;
                const int mm = sRowM[row], n = ncol0 + grp * 8;
                if (mm >= 0 && n < e.N) {
                    float f[8];
                    load8(row, grp * 8, f);
                    float* dst = args.ws + ((size_t)split * e.M + mm) * e.N + n;
                    *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
                }
            }
            if (args.split_sync) {
                // Fused reduction, no waiting: every split CTA publishes its partial tile and bumps
                // the tile's counter; whoever arrives LAST (its own partial is still in shared
                // memory) adds the other partials from L2 and runs the epilogue.  No co-residency
                // requirement, no second kernel.  The last CTA re-arms the counter.
                int* cnt = args.split_sync + (m_tile * n_tiles_grid + n_tile);
                int* s_last = reinterpret_cast<int*>(tmem_slot + 1);
                __threadfence();
                epi_bar();
                if (et == 0) {
                    const int old = atomicAdd(cnt, 1);
                    const int last = old == args.splits - 1;
                    if (last) *reinterpret_cast<volatile int*>(cnt) = 0;
                    *s_last = last;
                }
                epi_bar();
                if (*s_last) {
                    __threadfence();
                    for (int ps = 0; ps < args.splits; ++ps) {
                        if (ps == split) continue;
                        const float* wsp = args.ws + (size_t)ps * e.M * e.N;
#pragma unroll 1
                        for (int b = 0; b < kItems / kBatch; ++b) {
                            float4 lo[kBatch], hi[kBatch];
#pragma unroll
                            for (int j = 0; j < kBatch; ++j) {
                                int row, grp, mm, n;
                                item_addr(b * kBatch + j, row, grp, mm, n);
                                if (mm >= 0 && n < e.N) {
                                    const float* src = wsp + (size_t)mm * e.N + n;
                                    lo[j] = __ldcg(reinterpret_cast<const float4*>(src));
                                    hi[j] = __ldcg(reinterpret_cast<const float4*>(src + 4));
                                } else {
                                    lo[j] = hi[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < kBatch; ++j) {
                                int row, grp, mm, n;
                                item_addr(b * kBatch + j, row, grp, mm, n);
                                float4* d = reinterpret_cast<float4*>(sStage + row * L::kStagePitch + grp * 8);
                                float4 x = d[0], y = d[1];
                                x.x += lo[j].x; x.y += lo[j].y; x.z += lo[j].z; x.w += lo[j].w;
                                y.x += hi[j].x; y.y += hi[j].y; y.z += hi[j].z; y.w += hi[j].w;
                                d[0] = x; d[1] = y;
                            }
                        }
                    }
                    epi_bar();  // the epilogue below re-partitions the tile among the threads
                    const bool geglu = e.epi == SFB_EPI_GEGLU;
                    const int gpr = geglu ? kGroups / 2 : kGroups;  // items per row
#pragma unroll 1
                    for (int idx = et; idx < BM * gpr; idx += kEpiThreads) {
                        const int row = idx / gpr, grp = idx - row * gpr;
                        const int mm = sRowM[row];
                        const int n = (geglu ? n_tile * (BN / 2) : ncol0) + grp * 8;
                        if (mm < 0 || n >= (geglu ? e.geglu_n_out : e.N)) continue;
                        auto sum8 = [&](int col, float (&acc)[8]) {
                            load8(row, col - ncol0, acc);
                        };
                        reduce_epilogue8<BN, BF16>(sum8, e, mm, n);
                    }
                }
            }
        } else if (e.epi == SFB_EPI_GEGLU) {
#pragma unroll 1
            for (int it = 0; it < kItems / 2; ++it) {
                const int idx = et + it * kEpiThreads;
                const int row = idx / (kGroups / 2), og = idx - row * (kGroups / 2);
                const int mm = sRowM[row], nout = n_tile * (BN / 2) + og * 8;
                if (mm >= 0 && nout < e.geglu_n_out) {
                    float fv[8], fg[8];
                    load8(row, og * 8, fv);
                    load8(row, BN / 2 + og * 8, fg);
                    epi_geglu8<BF16>(e, mm, nout, fv, fg);
                }
            }
        } else {
            if (e.epi == SFB_EPI_STORE) {
#pragma unroll 1
                for (int b = 0; b < kItems / kBatch; ++b) {
                    if (has_res && b + 1 < kItems / kBatch) {
#pragma unroll
                        for (int j = 0; j < kBatch; ++j) rnxt[j] = load_res((b + 1) * kBatch + j);
                    }
#pragma unroll
                    for (int j = 0; j < kBatch; ++j) {
                        int row, grp, mm, n;
                        item_addr(b * kBatch + j, row, grp, mm, n);
                        if (mm >= 0 && n < e.N) {
                            float f[8];
                            load8(row, grp * 8, f);
                            if (has_res) add_res8(rcur[j], BF16, f);
                            if (e.rowstats_out || e.gn_stats[0]) {  // final values back to the tile (row / column sums)
                                float* d = sStage + row * L::kStagePitch + grp * 8;
                                *reinterpret_cast<float4*>(d) = make_float4(f[0], f[1], f[2], f[3]);
                                *reinterpret_cast<float4*>(d + 4) = make_float4(f[4], f[5], f[6], f[7]);
                            }
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.out) + (size_t)mm * e.ldo + n) =
                                pack8(f, BF16);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < kBatch; ++j) rcur[j] = rnxt[j];
                }
            } else {
                // QKV scatter.  V^T columns want consecutive lanes = consecutive rows (2-byte stores
                // along seq); Q / K want consecutive lanes = consecutive 16-byte slices of a row.
                const int C = e.heads * e.head_dim;
                const int n_last = min(ncol0 + BN, e.N) - 1;
                const bool row_fastest = (ncol0 / C + e.which_base == 2) && (n_last / C + e.which_base == 2);
#pragma unroll 1
                for (int it = 0; it < kItems; ++it) {
                    const int idx = et + it * kEpiThreads;
                    int row, grp;
                    if (row_fastest) { grp = idx >> 7; row = idx & 127; }
                    else { row = idx / kGroups; grp = idx - row * kGroups; }
                    const int mm = sRowM[row], n = ncol0 + grp * 8;
                    if (mm >= 0 && n < e.N) {
                        float f[8];
                        load8(row, grp * 8, f);
                        const int2 bs = sQkvRow[row];          // (batch, position)
                        const longlong2 ci = sQkvCol[grp];     // (q / k / v, column offset)
                        if (ci.x == 0) {
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.q) +
                                ((size_t)bs.x * e.heads * e.q_rows + bs.y) * e.q_pitch + ci.y) = pack8(f, BF16);
                        } else if (ci.x == 1) {
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.k) +
                                ((size_t)bs.x * e.heads * e.k_rows + bs.y) * e.q_pitch + ci.y) = pack8(f, BF16);
                        } else {
                            const size_t base = (size_t)bs.x * e.heads * e.vt_rows * e.vt_pitch + ci.y + bs.y;
#pragma unroll
                            for (int i = 0; i < 8; ++i) store1(e.vt, base + (size_t)i * e.vt_pitch, f[i], BF16);
                        }
                    }
                }
            }
            if (e.gn_stats[0] && e.epi == SFB_EPI_STORE) {
                // GroupNorm statistics of the finished tile for the consumer(s): per-column sums over
                // the tile's rows (split at image boundaries), merged per group in shared memory, then
                // one fire-and-forget global atomic per (consumer, image, group, moment).
                constexpr int kMaxImg = 8, kMaxGrp = BN / 2 + 2;
                int* sRowImg = reinterpret_cast<int*>(smem + BM * L::kStagePitch * 4);
                float* sGn = reinterpret_cast<float*>(sRowImg + BM);  // [2][kMaxImg][kMaxGrp][2]
                sRowImg[r] = valid ? m / e.gn_rpi : -1;
                for (int i = et; i < 2 * kMaxImg * kMaxGrp * 2; i += kEpiThreads) sGn[i] = 0.f;
                epi_bar();
                const int img_base = sRowImg[0];
                int gfirst[2] = {0, 0};
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (e.gn_stats[t]) gfirst[t] = (e.gn_choff[t] + ncol0) / e.gn_cpg[t];
                for (int c = et; c < BN; c += kEpiThreads) {
                    const int n = ncol0 + c;
                    if (n >= e.N) continue;
                    int gl[2] = {0, 0};
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                        if (e.gn_stats[t]) gl[t] = (e.gn_choff[t] + n) / e.gn_cpg[t] - gfirst[t];
                    float cs = 0.f, css = 0.f;
                    int cur = -1;
                    auto flush = [&]() {
                        if (cur < 0) return;
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            if (!e.gn_stats[t]) continue;
                            float* d = sGn + ((t * kMaxImg + (cur - img_base)) * kMaxGrp + gl[t]) * 2;
                            atomicAdd(d, cs);
                            atomicAdd(d + 1, css);
                        }
                    };
#pragma unroll 4
                    for (int row = 0; row < BM; ++row) {
                        const int img = sRowImg[row];
                        if (img < 0) continue;
                        if (img != cur) { flush(); cur = img; cs = 0.f; css = 0.f; }
                        const float v = round16<BF16>(sStage[row * L::kStagePitch + c]);
                        cs += v;
                        css += v * v;
                    }
                    flush();
                }
                epi_bar();
                for (int i = et; i < 2 * kMaxImg * kMaxGrp; i += kEpiThreads) {
                    const int t = i / (kMaxImg * kMaxGrp);
                    const int slot = (i / kMaxGrp) % kMaxImg, g = i % kMaxGrp;
                    if (!e.gn_stats[t]) continue;
                    const float a0 = sGn[i * 2], a1 = sGn[i * 2 + 1];
                    if (a0 != 0.f || a1 != 0.f) {
                        // 8 accumulation shards (by M tile) keep same-address atomic contention low
                        float* d = e.gn_stats[t] + (size_t)(m_tile & 7) * e.gn_shard_stride +
                                   ((size_t)(img_base + slot) * e.gn_groups + gfirst[t] + g) * 2;
                        atomicAdd(d, a0);
                        atomicAdd(d + 1, a1);
                    }
                }
            }
            if (e.rowstats_out) {
                epi_bar();
                if (valid) {
                    float rs_sum = 0.f, rs_sq = 0.f;
                    for (int g = chalf * (kGroups / kColSplit); g < (chalf + 1) * (kGroups / kColSplit); ++g) {
                        if (ncol0 + g * 8 < e.N) {
                            float f[8];
                            load8(r, g * 8, f);
                            row_stats8(f, BF16, rs_sum, rs_sq);
                        }
                    }
                    atomicAdd(e.rowstats_out + 2 * (size_t)m, rs_sum);
                    atomicAdd(e.rowstats_out + 2 * (size_t)m + 1, rs_sq);
                }
            }
        }
    }

    if (args.cluster_k) {
        // ---- cluster split-K: every CTA's fp32 partial tile now sits in its own shared memory.
        // CTA `split` reduces rows [split * rows_per, +rows_per) of the tile over all peers' tiles
        // (distributed shared memory) and runs the fused epilogue on them.
        cluster_sync_all();
        if (warp >= 2) {
            const EpiArgs& e = args.e;
            const int et = threadIdx.x - 64;
            constexpr int kGroups = BN / 8;
            const int rows_per = (BM + args.splits - 1) / args.splits;
            const int r0 = split * rows_per;
            const int nrows = max(0, min(rows_per, BM - r0));
            const bool geglu = e.epi == SFB_EPI_GEGLU;
            const int gpr = geglu ? kGroups / 2 : kGroups;  // items per row
            const int ncol0 = n_tile * BN;
            // peers: same position inside the pair (CG == 2: rank bit 0), every split
            const uint32_t my_rank = cluster_ctarank();
            const uint32_t rank0 = CG == 2 ? (my_rank & 1u) : 0u;
            const uint32_t stage_u32 = smem_u32(sStage);
#pragma unroll 1
            for (int idx = et; idx < nrows * gpr; idx += kEpiThreads) {
                const int row = r0 + idx / gpr, grp = idx % gpr;
                const int mm = sRowM[row];
                const int n = (geglu ? n_tile * (BN / 2) : ncol0) + grp * 8;
                if (mm < 0 || n >= (geglu ? e.geglu_n_out : e.N)) continue;
                auto sum8 = [&](int col, float (&acc)[8]) {
                    const uint32_t off = stage_u32 + (uint32_t)(row * L::kStagePitch + (col - ncol0)) * 4u;
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
                    for (int s = 0; s < args.splits; ++s) {
                        const uint32_t pa = dsmem_map(off, rank0 + (uint32_t)(s * CG));
                        const float4 a = dsmem_ld_f4(pa);
                        const float4 b = dsmem_ld_f4(pa + 16);
                        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
                        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
                    }
                };
                reduce_epilogue8<BN, BF16>(sum8, e, mm, n);
            }
        }
    }
    if (dbg && threadIdx.x == 64) dbg[6] = globaltimer_ns();
    tc_fence_before();
    // no CTA may exit while cluster peers can still signal its barriers / read its shared memory
    if (clustered) cluster_sync_all();
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair<kTmemCols>(tmem_base);
        else tmem_dealloc<kTmemCols>(tmem_base);
    }
    if (dbg && threadIdx.x == 32) dbg[7] = globaltimer_ns();
}

}  // namespace sfb

using namespace sfb;

template <int BN, int STAGES, int BF16, int CG>
static int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, dim3 grid,
                         cudaStream_t stream) {
    using L = GemmSmem<BN, STAGES, CG>;
    static PerDeviceOnce attr_once;
    bool& attr_set = attr_once.flag();
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, BF16, CG>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "sfb_gemm: smem attribute: %s", cudaGetErrorString(err));
        attr_set = true;
    }
    const int cz = a.cluster_k ? a.splits : 1;
    if (cz * CG > 8) {
        static PerDeviceOnce np_once;
        bool& np_set = np_once.flag();
        if (!np_set) {
            cudaError_t err = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, BF16, CG>,
                                                   cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "sfb_gemm: non-portable cluster attribute: %s", cudaGetErrorString(err));
            np_set = true;
        }
    }
    cudaError_t err = launch_cluster_pdl(gemm_tc_kernel<BN, STAGES, BF16, CG>, grid, dim3(kGemmThreads),
                                         CG == 2 ? dim3(2, 1, cz) : dim3(a.cx, a.cy, cz), L::kTotal, stream,
                                         ta, tb, a);
    if (err != cudaSuccess) {
        // diagnostics: can the requested cluster be scheduled at all?
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = L::kTotal; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CG == 2 ? 2 : a.cx; at[0].val.clusterDim.y = CG == 2 ? 1 : a.cy; at[0].val.clusterDim.z = cz;
        cfg.attrs = at; cfg.numAttrs = 1;
        int max_clusters = -1;
        cudaError_t e2 = cudaOccupancyMaxActiveClusters(&max_clusters, gemm_tc_kernel<BN, STAGES, BF16, CG>, &cfg);
        cudaGetLastError();
        return fail(SFB_ERR_CUDA, "sfb_gemm: launch: %s (%s) grid=(%u,%u,%u) cluster=(%d,%d) smem=%d stages=%d cg=%d maxActiveClusters=%d (%s)",
                    cudaGetErrorString(err), cudaGetErrorName(err), grid.x, grid.y, grid.z,
                    CG == 2 ? 1 : a.cx, CG == 2 ? 2 : a.cy, (int)L::kTotal, STAGES, CG, max_clusters,
                    cudaGetErrorName(e2));
    }
    return check_launch("sfb_gemm");
}

template <int BN, int STAGES, int CG>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, dim3 grid,
                       cudaStream_t stream) {
    return a.e.dtype == SFB_BF16 ? launch_gemm_t<BN, STAGES, 1, CG>(ta, tb, a, grid, stream)
                                 : launch_gemm_t<BN, STAGES, 0, CG>(ta, tb, a, grid, stream);
}

extern "C" int sfb_gemm(const sfb_gemm_params* p, sfb_stream_t stream_) {
    constexpr int BN = 160;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!p || !p->tmap_a || !p->tmap_b) return fail(SFB_ERR_INVALID, "sfb_gemm: null argument");
    if (p->K <= 0 || p->K % BK) return fail(SFB_ERR_INVALID, "sfb_gemm: K=%d must be a multiple of 64", p->K);
    if (p->N <= 0 || p->N % 8 || p->M <= 0) return fail(SFB_ERR_INVALID, "sfb_gemm: bad M=%d N=%d", p->M, p->N);
    if (p->dtype != SFB_F16 && p->dtype != SFB_BF16) return fail(SFB_ERR_INVALID, "sfb_gemm: dtype");
    GemmArgs a{};
    a.a_mode = p->a_mode;
    a.nkb_total = p->K / BK;
    a.splits = p->splits < 1 ? 1 : p->splits;
    a.pdl = g_pdl;
    a.dbg = reinterpret_cast<long long*>(p->debug_stamps);
    a.split_sync = nullptr;
    if (a.splits > a.nkb_total) return fail(SFB_ERR_INVALID, "sfb_gemm: splits > K blocks");
    a.cluster_k = (p->cluster_k && a.splits > 1) ? 1 : 0;
    if (a.cluster_k && (a.splits * (p->cta_pair ? 2 : 1) > 16 || p->gn_stats[0] || p->cluster_n > 1 || p->cluster_m > 1))
        return fail(SFB_ERR_INVALID, "sfb_gemm: cluster_k needs splits * pair <= 16, no gn_stats, no multicast cluster");
    if (a.splits > 1 && !a.cluster_k && !p->ws) return fail(SFB_ERR_INVALID, "sfb_gemm: split-K needs a workspace");
    a.ws = p->ws;
    int m_tiles;
    const bool upconv = p->a_mode == SFB_A_UPCONV2X;
    if (p->a_mode == SFB_A_CONV3X3 || upconv) {
        if (p->cin <= 0 || p->cin % BK || p->K != (upconv ? 4 : 9) * p->cin)
            return fail(SFB_ERR_INVALID, "sfb_gemm: conv cin=%d K=%d", p->cin, p->K);
        const int box_w = p->box_w > 0 ? p->box_w : p->img_w;
        if (p->box_n * p->box_h * box_w != BM || box_w > p->img_w || p->img_w % box_w)
            return fail(SFB_ERR_INVALID, "sfb_gemm: conv M-tile box %dx%dx%d != 128 pixels / does not tile width %d",
                        p->box_n, p->box_h, box_w, p->img_w);
        if (p->box_n > 1 && (p->box_h != p->img_h || box_w != p->img_w))
            return fail(SFB_ERR_INVALID, "sfb_gemm: multi-image box needs box_h == img_h and box_w == img_w");
        if (box_w != p->img_w && (p->cluster_n > 1 || p->cluster_m > 1))
            return fail(SFB_ERR_INVALID, "sfb_gemm: patch tiles (box_w < img_w) do not combine with multicast clusters");
        if (p->M != (upconv ? 4 : 1) * p->img_n * p->img_h * p->img_w) return fail(SFB_ERR_INVALID, "sfb_gemm: conv M mismatch");
        if (upconv && (p->conv_stride != 1 || p->epi != SFB_EPI_STORE || p->cluster_n > 1 || p->cluster_m > 1))
            return fail(SFB_ERR_INVALID, "sfb_gemm: up-conv needs stride 1, the STORE epilogue and no multicast cluster");
        a.img_n = p->img_n; a.img_h = p->img_h; a.img_w = p->img_w;
        a.cpb = p->cin / BK;
        a.conv_stride = p->conv_stride;
        a.box_h = p->box_h; a.box_n = p->box_n; a.box_w = box_w;
        a.tiles_per_row = p->img_w / box_w;
        a.tiles_per_img = ((p->img_h + p->box_h - 1) / p->box_h) * a.tiles_per_row;
        m_tiles = (p->box_n == 1) ? p->img_n * a.tiles_per_img : (p->img_n + p->box_n - 1) / p->box_n;
        if (upconv) {  // 4 output phases, each its own set of M tiles and its own weight row block
            a.up_tiles = m_tiles;
            a.up_ntiles = (p->N + BN - 1) / BN;
            m_tiles *= 4;
        }
    } else if (p->a_mode == SFB_A_MATRIX) {
        m_tiles = (p->M + BM - 1) / BM;
    } else {
        return fail(SFB_ERR_INVALID, "sfb_gemm: a_mode");
    }
    if (p->rowbias && p->a_mode != SFB_A_CONV3X3)
        return fail(SFB_ERR_INVALID, "sfb_gemm: rowbias (time-embedding add) needs conv mode");
    EpiArgs& e = a.e;
    e.epi = p->epi; e.dtype = p->dtype; e.M = p->M; e.N = p->N;
    e.out = p->out; e.ldo = p->ldo; e.bias = p->bias; e.rowbias = p->rowbias;
    e.rows_per_img = p->rows_per_img > 0 ? p->rows_per_img : 1;
    e.ld_rowbias = p->ld_rowbias; e.residual = p->residual; e.ldr = p->ldr;
    e.q = p->q; e.k = p->k; e.vt = p->vt; e.heads = p->heads; e.head_dim = p->head_dim;
    e.which_base = p->which_base; e.seq = p->seq; e.q_pitch = p->q_pitch; e.q_rows = p->q_rows;
    e.k_rows = p->k_rows; e.vt_rows = p->vt_rows; e.vt_pitch = p->vt_pitch;
    for (int t = 0; t < 2; ++t) {
        e.gn_stats[t] = p->gn_stats[t]; e.gn_cpg[t] = p->gn_cpg[t]; e.gn_choff[t] = p->gn_choff[t];
    }
    e.gn_groups = p->gn_groups; e.gn_rpi = p->gn_rows_per_img; e.gn_shard_stride = p->gn_shard_stride;
    if (p->gn_stats[0]) {
        if (p->epi != SFB_EPI_STORE || p->gn_groups <= 0 || p->gn_groups > 64 || p->gn_shard_stride <= 0 ||
            p->gn_rows_per_img < 16 || p->gn_cpg[0] < 2 ||
            (p->gn_stats[1] && p->gn_cpg[1] < 2) ||
            (BM % p->gn_rows_per_img != 0 && p->gn_rows_per_img % BM != 0))
            return fail(SFB_ERR_INVALID, "sfb_gemm: unsupported GroupNorm statistics geometry");
    } else if (p->gn_stats[1]) {
        return fail(SFB_ERR_INVALID, "sfb_gemm: gn_stats[1] without gn_stats[0]");
    }
    e.rowstats_out = p->rowstats_out; e.ln_rowstats = p->ln_rowstats; e.ln_colsum = p->ln_colsum;
    e.ln_eps = p->ln_eps; e.ln_dim = p->ln_dim;
    if (p->ln_rowstats && (!p->ln_colsum || p->ln_dim <= 0 || p->rowbias))
        return fail(SFB_ERR_INVALID, "sfb_gemm: LayerNorm fold needs ln_colsum / ln_dim and no rowbias");
    if (p->rowstats_out && p->epi != SFB_EPI_STORE)
        return fail(SFB_ERR_INVALID, "sfb_gemm: rowstats_out needs the STORE epilogue");
    if (p->epi == SFB_EPI_STORE) {
        if (!p->out || p->ldo % 8) return fail(SFB_ERR_INVALID, "sfb_gemm: out/ldo");
        if (p->residual && p->ldr % 8) return fail(SFB_ERR_INVALID, "sfb_gemm: ldr");
    } else if (p->epi == SFB_EPI_GEGLU) {
        if (!p->out || p->ldo % 8 || p->N % BN) return fail(SFB_ERR_INVALID, "sfb_gemm: geglu needs N %% 160 == 0");
        if (p->geglu_n_out <= 0 || p->geglu_n_out % 8 || p->geglu_n_out > p->N / 2)
            return fail(SFB_ERR_INVALID, "sfb_gemm: geglu_n_out");
        e.geglu_n_out = p->geglu_n_out;
    } else if (p->epi == SFB_EPI_QKV) {
        if (p->head_dim % 8 || p->heads <= 0 || p->seq <= 0 || p->N % (p->heads * p->head_dim))
            return fail(SFB_ERR_INVALID, "sfb_gemm: qkv geometry");
    } else {
        return fail(SFB_ERR_INVALID, "sfb_gemm: epilogue mode");
    }
    dim3 grid((p->N + BN - 1) / BN, m_tiles, a.splits);
    a.cx = p->cluster_n > 0 ? p->cluster_n : 1;
    a.cy = p->cluster_m > 0 ? p->cluster_m : 1;
    a.a_part_dim = p->a_part_dim;
    a.a_part_ext = p->a_part_ext;
    if ((a.cx != 1 && a.cx != 2) || (a.cy != 1 && a.cy != 2 && a.cy != 4) || grid.x % a.cx || grid.y % a.cy)
        return fail(SFB_ERR_INVALID, "sfb_gemm: cluster %dx%d does not divide the %ux%u tile grid", a.cx, a.cy, grid.x, grid.y);
    if (a.cx > 1 && p->a_mode == SFB_A_CONV3X3 && (a.a_part_dim < 1 || a.a_part_dim > 3 || a.a_part_ext <= 0))
        return fail(SFB_ERR_INVALID, "sfb_gemm: conv cluster needs a_part_dim / a_part_ext");
    CUtensorMap ta, tb;
    memcpy(&ta, p->tmap_a, sizeof(CUtensorMap));
    memcpy(&tb, p->tmap_b, sizeof(CUtensorMap));
    // <= one CTA per SM anyway: take the deep 6-stage pipeline; otherwise 3 stages x 2 CTAs/SM
    const long long ctas = (long long)grid.x * grid.y * grid.z;
    // fused split-K reduction (last-arriving CTA of a tile finishes it); GroupNorm statistics only
    // exist in the stand-alone reduction kernel
    if (a.splits > 1 && !a.cluster_k && p->split_sync && !p->gn_stats[0]) a.split_sync = reinterpret_cast<int*>(p->split_sync);
    static const int force_stages = [] { const char* v = getenv("SFB_GEMM_STAGES"); return v ? atoi(v) : 0; }();
    const bool deep = force_stages ? (force_stages == 6) : (ctas <= 148);
    int rc;
    if (p->cta_pair) {
        // CTA pairs along M (cluster 1x2, tcgen05.mma.cta_group::2): tmap_b box = 80 rows
        if (grid.y % 2 || a.cx != 1 || a.cy != 1 || (upconv && a.up_tiles % 2))
            return fail(SFB_ERR_INVALID, "sfb_gemm: cta_pair needs an even number of M tiles (per up-conv phase) and no multicast cluster");
        const dim3 pgrid(grid.y, grid.x, grid.z);  // M tiles along x: pairs are x-neighbours
        rc = deep ? launch_gemm<BN, 8, 2>(ta, tb, a, pgrid, stream) : launch_gemm<BN, 4, 2>(ta, tb, a, pgrid, stream);
    } else {
        rc = deep ? launch_gemm<BN, 6, 1>(ta, tb, a, grid, stream) : launch_gemm<BN, 3, 1>(ta, tb, a, grid, stream);
    }
    if (rc) return rc;
    if (p->defer_finish && a.splits > 1) {
        if (a.split_sync || a.cluster_k || e.epi != SFB_EPI_STORE || e.rowstats_out || e.ln_rowstats || e.gn_stats[0])
            return fail(SFB_ERR_INVALID, "sfb_gemm: defer_finish needs a plain STORE epilogue and the workspace split-K path");
        return rc;  // the consumer (sfb_group_norm_fused with part_ws) finishes the tensor
    }
    if (a.splits > 1 && !a.split_sync && !a.cluster_k) {
        const int ncols = (e.epi == SFB_EPI_GEGLU) ? e.geglu_n_out : e.N;
        const long long items = (long long)e.M * (ncols / 8);
        const int blocks = (int)((items + 255) / 256);
        cudaError_t err = (e.dtype == SFB_BF16)
            ? launch_pdl(splitk_finish_kernel<BN, 1>, dim3(blocks), dim3(256), 0, stream, (const float*)a.ws, a.splits, e)
            : launch_pdl(splitk_finish_kernel<BN, 0>, dim3(blocks), dim3(256), 0, stream, (const float*)a.ws, a.splits, e);
        if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "sfb_gemm: finish launch: %s", cudaGetErrorString(err));
        rc = check_launch("sfb_gemm(split-K finish)");
    }
    return rc;
}
