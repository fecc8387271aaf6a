// GroupNorm (+SiLU) and LayerNorm for NHWC / token-major 16-bit activations (HBM-bound kernels).
//
// GroupNorm is two passes, like the reference's NHWC path (/root/reference/src/sfast/triton/ops/
// group_norm.py:126-165 stats, :272-320 apply) but with the statistics kept in fp32 end to end
// (the reference round-trips mean/rstd through the input dtype, group_norm.py:405-416) and with
// the stats grid sized to fill 148 SMs instead of groups*batch CTAs.
// LayerNorm is one warp per row, row held in registers, two-pass mean/variance in fp32
// (reference: triton/ops/layer_norm.py:51-133).
#include "common.cuh"
#include "host.h"

#include <stdlib.h>

namespace sfb {

constexpr int kGnThreads = 512;

struct GnArgs {
    const uint16_t* x;
    uint16_t* y;
    const float* gamma;
    const float* beta;
    float* stats;
    int n, hw, c, ldx, ldy, groups, cpg, nvec, rows_per_block;
    int shards, shard_stride;  // statistics arrive as `shards` partial copies, `shard_stride` floats apart
    float eps;
    int silu, dtype;
    // deferred split-K finish of the GEMM producing channels [0, part_c) (fused kernel only)
    const float* part_ws;
    int part_splits, part_c, part_ld;
    const float* part_bias;
    const float* part_rowbias;
    int part_ld_rowbias;
    const uint16_t* part_residual;
    int part_ldr;
};

// Finished 16-bit values of channels [tx*8, tx*8+8) of pixel `row` of image `img` from the
// producer GEMM's fp32 split-K partials: sum + bias + per-image row bias + residual (the STORE
// epilogue of gemm_tc.cu's reduction kernel), also written to x.
__device__ __forceinline__ uint4 gn_finish_partials(const GnArgs& a, int img, int row, int tx) {
    const int m = img * a.hw + row;
    const int col = tx * 8;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    constexpr int kU = 6;  // partial loads in flight per batch
    const size_t stride = (size_t)a.n * a.hw * a.part_ld;
    const float* p0 = a.part_ws + (size_t)m * a.part_ld + col;
    for (int s0 = 0; s0 < a.part_splits; s0 += kU) {
        float4 lo[kU], hi[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (s0 + u < a.part_splits) {
                const float* p = p0 + (size_t)(s0 + u) * stride;
                lo[u] = __ldcg(reinterpret_cast<const float4*>(p));
                hi[u] = __ldcg(reinterpret_cast<const float4*>(p + 4));
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (s0 + u < a.part_splits) {
                acc[0] += lo[u].x; acc[1] += lo[u].y; acc[2] += lo[u].z; acc[3] += lo[u].w;
                acc[4] += hi[u].x; acc[5] += hi[u].y; acc[6] += hi[u].z; acc[7] += hi[u].w;
            }
        }
    }
    auto add8 = [&](const float* b) {
        const float4 b0 = *reinterpret_cast<const float4*>(b);
        const float4 b1 = *reinterpret_cast<const float4*>(b + 4);
        acc[0] += b0.x; acc[1] += b0.y; acc[2] += b0.z; acc[3] += b0.w;
        acc[4] += b1.x; acc[5] += b1.y; acc[6] += b1.z; acc[7] += b1.w;
    };
    if (a.part_bias) add8(a.part_bias + col);
    if (a.part_rowbias) add8(a.part_rowbias + (size_t)img * a.part_ld_rowbias + col);
    if (a.part_residual) {
        const uint4 r = *reinterpret_cast<const uint4*>(a.part_residual + (size_t)m * a.part_ldr + col);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = unpack2(w[i], a.dtype);
            acc[2 * i] += f.x; acc[2 * i + 1] += f.y;
        }
    }
    uint4 v;
    v.x = pack2(acc[0], acc[1], a.dtype); v.y = pack2(acc[2], acc[3], a.dtype);
    v.z = pack2(acc[4], acc[5], a.dtype); v.w = pack2(acc[6], acc[7], a.dtype);
    *reinterpret_cast<uint4*>(const_cast<uint16_t*>(a.x) + (size_t)m * a.ldx + col) = v;
    return v;
}

// grid (blocks_per_img, n).  Thread (tx = vector column, ty = row slot) keeps per-channel
// partial sums for its 8 channels over rows ty, ty+BY, ...; one shared-memory atomic per
// (thread, group run) at the end, then one global atomic per (block, group, moment).
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const GnArgs a) {
    __shared__ float acc[2 * 64];
    pdl_launch_dependents();
    pdl_wait();
    const int img = blockIdx.y;
    const int by = kGnThreads / a.nvec;
    const int tx = threadIdx.x % a.nvec;
    const int ty = threadIdx.x / a.nvec;
    for (int i = threadIdx.x; i < 2 * a.groups; i += kGnThreads) acc[i] = 0.f;
    __syncthreads();
    if (ty < by) {
        float s[8], ss[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
        const int row0 = blockIdx.x * a.rows_per_block;
        const int row1 = min(row0 + a.rows_per_block, a.hw);
        const uint16_t* base = a.x + (size_t)img * a.hw * a.ldx + tx * 8;
        constexpr int kU = 8;
        for (int rb = row0 + ty; rb < row1; rb += by * kU) {
            uint4 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int row = rb + u * by;
                if (row < row1) v[u] = *reinterpret_cast<const uint4*>(base + (size_t)row * a.ldx);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (rb + u * by < row1) {
                    const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = unpack2(w[i], a.dtype);
                        s[2 * i] += f.x; ss[2 * i] += f.x * f.x;
                        s[2 * i + 1] += f.y; ss[2 * i + 1] += f.y * f.y;
                    }
                }
            }
        }
        // merge runs of equal group id among the 8 channels
        int g_run = (tx * 8) / a.cpg;
        float rs = 0.f, rss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int g = (tx * 8 + i) / a.cpg;
            if (g != g_run) {
                atomicAdd(&acc[2 * g_run], rs);
                atomicAdd(&acc[2 * g_run + 1], rss);
                g_run = g; rs = 0.f; rss = 0.f;
            }
            rs += s[i]; rss += ss[i];
        }
        atomicAdd(&acc[2 * g_run], rs);
        atomicAdd(&acc[2 * g_run + 1], rss);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * a.groups; i += kGnThreads)
        atomicAdd(&a.stats[(size_t)img * a.groups * 2 + i], acc[i]);
}

// grid (blocks_per_img, n): per-channel scale/shift for this image staged in shared memory, then
// a streaming y = act(x * scale + shift) over the block's rows with 16-byte accesses.
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const GnArgs a) {
    extern __shared__ float sm[];
    pdl_launch_dependents();
    pdl_wait();
    float* scale = sm;
    float* shift = sm + a.c;
    const int img = blockIdx.y;
    const float inv_cnt = 1.0f / ((float)a.hw * (float)a.cpg);
    for (int ch = threadIdx.x; ch < a.c; ch += kGnThreads) {
        const int g = ch / a.cpg;
        float sum = 0.f, sq = 0.f;
        for (int s = 0; s < a.shards; ++s) {
            sum += a.stats[(size_t)s * a.shard_stride + ((size_t)img * a.groups + g) * 2];
            sq += a.stats[(size_t)s * a.shard_stride + ((size_t)img * a.groups + g) * 2 + 1];
        }
        const float mean = sum * inv_cnt;
        const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + a.eps);
        const float sc = rstd * a.gamma[ch];
        scale[ch] = sc;
        shift[ch] = a.beta[ch] - mean * sc;
    }
    __syncthreads();
    const int row0 = blockIdx.x * a.rows_per_block;
    const int row1 = min(row0 + a.rows_per_block, a.hw);
    const int items = (row1 - row0) * a.nvec;
    const uint16_t* xb = a.x + ((size_t)img * a.hw + row0) * a.ldx;
    uint16_t* yb = a.y + ((size_t)img * a.hw + row0) * a.ldy;
    constexpr int kU = 4;
    for (int it0 = threadIdx.x; it0 < items; it0 += kGnThreads * kU) {
        uint4 v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int it = it0 + u * kGnThreads;
            if (it < items) {
                const int row = it / a.nvec, vec = it - row * a.nvec;
                v[u] = *reinterpret_cast<const uint4*>(xb + (size_t)row * a.ldx + vec * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int it = it0 + u * kGnThreads;
            if (it >= items) continue;
            const int row = it / a.nvec, vec = it - row * a.nvec;
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack2(w[i], a.dtype);
                const int ch = vec * 8 + 2 * i;
                float y0 = f.x * scale[ch] + shift[ch];
                float y1 = f.y * scale[ch + 1] + shift[ch + 1];
                if (a.silu) { y0 = silu_f(y0); y1 = silu_f(y1); }
                o[i] = pack2(y0, y1, a.dtype);
            }
            *reinterpret_cast<uint4*>(yb + (size_t)row * a.ldy + vec * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}


// ---------------------------------------------------------------------------------------
// Fused GroupNorm for tensors that fit in the machine's shared memory (the B <= 2 regime the
// SD-1.5 bs=1 config lives in): ONE launch, x is read from HBM/L2 once.
//   phase 1  each CTA pulls its slab of rows into shared memory and accumulates group sums
//            (fp32, one global atomic per (CTA, group, moment));
//   barrier  grid-wide arrive/spin on a global counter (grid <= 148 CTAs, all co-resident);
//   phase 2  scale/shift from the finished statistics, y = act(x * scale + shift) from smem.
// ---------------------------------------------------------------------------------------
template <bool kPart>
__global__ void __launch_bounds__(kGnThreads) gn_fused_kernel(const GnArgs a, unsigned* sync_counter) {
    extern __shared__ __align__(16) uint8_t fsm[];
    __shared__ float acc[2 * 64];
    float* scale = reinterpret_cast<float*>(fsm);
    float* shift = scale + a.c;
    uint4* slab = reinterpret_cast<uint4*>(fsm + 2 * a.c * sizeof(float));
    pdl_launch_dependents();
    pdl_wait();
    const int img = blockIdx.y;
    const int by = kGnThreads / a.nvec;
    const int tx = threadIdx.x % a.nvec;
    const int ty = threadIdx.x / a.nvec;
    for (int i = threadIdx.x; i < 2 * a.groups; i += kGnThreads) acc[i] = 0.f;
    __syncthreads();
    const int row0 = blockIdx.x * a.rows_per_block;
    const int row1 = min(row0 + a.rows_per_block, a.hw);
    if (ty < by) {
        float s[8], ss[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
        const uint16_t* base = a.x + (size_t)img * a.hw * a.ldx + tx * 8;
        auto accumulate = [&](const uint4& v) {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack2(w[i], a.dtype);
                s[2 * i] += f.x; ss[2 * i] += f.x * f.x;
                s[2 * i + 1] += f.y; ss[2 * i + 1] += f.y * f.y;
            }
        };
        if constexpr (kPart) {
            // channels [0, part_c) come from the producer GEMM's split-K partials (a few rows per
            // thread, each with all its partial loads in flight)
            const bool from_partials = tx * 8 < a.part_c;
            for (int row = row0 + ty; row < row1; row += by) {
                const uint4 v = from_partials ? gn_finish_partials(a, img, row, tx)
                                              : *reinterpret_cast<const uint4*>(base + (size_t)row * a.ldx);
                slab[(row - row0) * a.nvec + tx] = v;
                accumulate(v);
            }
        } else {
            // all of a thread's loads of one batch are in flight together: the slab is a handful of
            // rows per thread, so a load-use-load chain would be nothing but exposed L2 latency
            constexpr int kU = 8;
            for (int rb = row0 + ty; rb < row1; rb += by * kU) {
                uint4 v[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int row = rb + u * by;
                    if (row < row1) v[u] = *reinterpret_cast<const uint4*>(base + (size_t)row * a.ldx);
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int row = rb + u * by;
                    if (row < row1) {
                        slab[(row - row0) * a.nvec + tx] = v[u];
                        accumulate(v[u]);
                    }
                }
            }
        }
        int g_run = (tx * 8) / a.cpg;
        float rs = 0.f, rss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int g = (tx * 8 + i) / a.cpg;
            if (g != g_run) {
                atomicAdd(&acc[2 * g_run], rs);
                atomicAdd(&acc[2 * g_run + 1], rss);
                g_run = g; rs = 0.f; rss = 0.f;
            }
            rs += s[i]; rss += ss[i];
        }
        atomicAdd(&acc[2 * g_run], rs);
        atomicAdd(&acc[2 * g_run + 1], rss);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * a.groups; i += kGnThreads)
        atomicAdd(&a.stats[(size_t)img * a.groups * 2 + i], acc[i]);
    // affine parameters do not depend on the statistics: fetch them (cold, from HBM) while the
    // grid barrier is pending
    for (int ch = threadIdx.x; ch < a.c; ch += kGnThreads) {
        scale[ch] = a.gamma[ch];
        shift[ch] = a.beta[ch];
    }
    // ---- grid barrier
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        atomicAdd(sync_counter, 1u);
        long long t0 = clock64();
        while (*reinterpret_cast<volatile unsigned*>(sync_counter) < total) {
            if (clock64() - t0 > 4000000000LL) __trap();  // not co-resident: fail, do not hang
        }
        __threadfence();
    }
    __syncthreads();
    // ---- phase 2
    const float inv_cnt = 1.0f / ((float)a.hw * (float)a.cpg);
    for (int ch = threadIdx.x; ch < a.c; ch += kGnThreads) {
        const int g = ch / a.cpg;
        const float sum = __ldcg(&a.stats[((size_t)img * a.groups + g) * 2]);
        const float sq = __ldcg(&a.stats[((size_t)img * a.groups + g) * 2 + 1]);
        const float mean = sum * inv_cnt;
        const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + a.eps);
        const float sc = rstd * scale[ch];
        scale[ch] = sc;
        shift[ch] = shift[ch] - mean * sc;
    }
    __syncthreads();
    if (ty < by) {
        uint16_t* yb = a.y + (size_t)img * a.hw * a.ldy + tx * 8;
        float sc[8], sh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { sc[i] = scale[tx * 8 + i]; sh[i] = shift[tx * 8 + i]; }
        for (int row = row0 + ty; row < row1; row += by) {
            const uint4 v = slab[(row - row0) * a.nvec + tx];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack2(w[i], a.dtype);
                float y0 = f.x * sc[2 * i] + sh[2 * i];
                float y1 = f.y * sc[2 * i + 1] + sh[2 * i + 1];
                if (a.silu) { y0 = silu_f(y0); y1 = silu_f(y1); }
                o[i] = pack2(y0, y1, a.dtype);
            }
            *reinterpret_cast<uint4*>(yb + (size_t)row * a.ldy) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

constexpr int kGnFusedMaxSmem = 200 * 1024;

// ---------------------------------------------------------------------------------------
// GroupNorm with one CTA per (image, group): the group's [hw, cpg] slab (strided 2*cpg-byte
// segments of the NHWC rows) is pulled into shared memory once, reduced inside the CTA and
// normalised from shared memory.  No grid barrier, no global atomics, no statistics buffer:
// the dependent-latency chain of the barrier kernel above (atomics -> fence -> counter -> spin
// -> statistics reload) disappears, which is what a small-batch step is made of.  Used whenever
// hw * cpg * 2 bytes fit in shared memory (every SD-1.5 / SDXL GroupNorm at 64x64 latents
// except the 960-channel concat at full resolution).
// ---------------------------------------------------------------------------------------
constexpr int kGnGroupThreads = 512;

// channels (col, col+1) of pixel `px` of image `img` from the producer GEMM's split-K partials
__device__ __forceinline__ uint32_t gn_finish_partials2(const GnArgs& a, int img, int px, int col) {
    const int m = img * a.hw + px;
    const size_t stride = (size_t)a.n * a.hw * a.part_ld;
    const float* p0 = a.part_ws + (size_t)m * a.part_ld + col;
    float ax = 0.f, ay = 0.f;
#pragma unroll 6
    for (int s = 0; s < a.part_splits; ++s) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(p0 + (size_t)s * stride));
        ax += v.x; ay += v.y;
    }
    if (a.part_bias) { ax += a.part_bias[col]; ay += a.part_bias[col + 1]; }
    if (a.part_rowbias) {
        const float* rb = a.part_rowbias + (size_t)img * a.part_ld_rowbias + col;
        ax += rb[0]; ay += rb[1];
    }
    if (a.part_residual) {
        const float2 r = unpack2(*reinterpret_cast<const uint32_t*>(a.part_residual + (size_t)m * a.part_ldr + col), a.dtype);
        ax += r.x; ay += r.y;
    }
    const uint32_t v = pack2(ax, ay, a.dtype);
    *reinterpret_cast<uint32_t*>(const_cast<uint16_t*>(a.x) + (size_t)m * a.ldx + col) = v;
    return v;
}

template <bool kPart>
__global__ void __launch_bounds__(kGnGroupThreads) gn_group_kernel(const GnArgs a) {
    extern __shared__ __align__(16) uint8_t gsm[];
    uint32_t* slab = reinterpret_cast<uint32_t*>(gsm);  // [hw][cpg / 2] pairs of 16-bit values
    __shared__ float s_gamma[128], s_beta[128];
    __shared__ float s_red[2][kGnGroupThreads / 32];
    __shared__ float s_stat[2];
    const int g = blockIdx.x, img = blockIdx.y;
    const int hp = a.cpg >> 1;
    const int items = a.hw * hp;
    const int ch0 = g * a.cpg;
    pdl_launch_dependents();
    // the affine parameters are weights: fetch them before waiting for the producer kernel
    for (int i = threadIdx.x; i < a.cpg; i += kGnGroupThreads) {
        s_gamma[i] = a.gamma[ch0 + i];
        s_beta[i] = a.beta[ch0 + i];
    }
    pdl_wait();
    const uint16_t* xb = a.x + (size_t)img * a.hw * a.ldx + ch0;
    float s = 0.f, ss = 0.f;
    constexpr int kU = kPart ? 2 : 8;  // items whose loads are in flight together
    for (int i0 = threadIdx.x; i0 < items; i0 += kGnGroupThreads * kU) {
        uint32_t v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int i = i0 + u * kGnGroupThreads;
            if (i < items) {
                const int px = i / hp, pr = i - px * hp;
                if (kPart && ch0 + 2 * pr < a.part_c)
                    v[u] = gn_finish_partials2(a, img, px, ch0 + 2 * pr);
                else
                    v[u] = *reinterpret_cast<const uint32_t*>(xb + (size_t)px * a.ldx + 2 * pr);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int i = i0 + u * kGnGroupThreads;
            if (i < items) {
                slab[i] = v[u];
                const float2 f = unpack2(v[u], a.dtype);
                s += f.x + f.y;
                ss += f.x * f.x + f.y * f.y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_red[0][warp] = s; s_red[1][warp] = ss; }
    __syncthreads();
    if (warp == 0) {
        float t = lane < kGnGroupThreads / 32 ? s_red[0][lane] : 0.f;
        float tt = lane < kGnGroupThreads / 32 ? s_red[1][lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t += __shfl_xor_sync(0xffffffffu, t, o);
            tt += __shfl_xor_sync(0xffffffffu, tt, o);
        }
        if (lane == 0) {
            const float inv_cnt = 1.0f / ((float)a.hw * (float)a.cpg);
            const float mean = t * inv_cnt;
            const float var = fmaxf(tt * inv_cnt - mean * mean, 0.f);
            s_stat[0] = mean;
            s_stat[1] = rsqrtf(var + a.eps);
        }
    }
    __syncthreads();
    const float mean = s_stat[0], rstd = s_stat[1];
    uint16_t* yb = a.y + (size_t)img * a.hw * a.ldy + ch0;
    for (int i = threadIdx.x; i < items; i += kGnGroupThreads) {
        const int px = i / hp, pr = i - px * hp;
        const float2 f = unpack2(slab[i], a.dtype);
        const float sc0 = rstd * s_gamma[2 * pr], sc1 = rstd * s_gamma[2 * pr + 1];
        float y0 = (f.x - mean) * sc0 + s_beta[2 * pr];
        float y1 = (f.y - mean) * sc1 + s_beta[2 * pr + 1];
        if (a.silu) { y0 = silu_f(y0); y1 = silu_f(y1); }
        *reinterpret_cast<uint32_t*>(yb + (size_t)px * a.ldy + 2 * pr) = pack2(y0, y1, a.dtype);
    }
}

// whether the per-(image, group) kernel applies: the group slab must fit in shared memory
static bool gn_group_fits(const sfb_gn_params* p, size_t& smem) {
    // measured on B200: 4-byte strided accesses from only n * groups CTAs cost more than the grid
    // barrier they avoid (5.60 vs 4.88 ms per step at B = 2), so this kernel is opt-in
    static const bool enabled = [] { const char* v = getenv("SFB_GN_GROUP"); return v && v[0] == '1'; }();
    if (!enabled || p->groups <= 0 || p->c % p->groups || p->n <= 0) return false;
    const int cpg = p->c / p->groups;
    if (cpg % 2 || cpg > 128 || p->ldx % 2 || p->ldy % 2) return false;
    smem = (size_t)p->hw * cpg * 2;
    return smem <= (size_t)kGnFusedMaxSmem;
}

// grid geometry of the fused kernel; returns false if the tensor does not fit in shared memory
static bool gn_fused_geometry(const sfb_gn_params* p, int& blocks_per_img, int& rows_per_block,
                              size_t& smem) {
    if (p->n <= 0 || p->n > 148) return false;
    blocks_per_img = 148 / p->n;
    if (blocks_per_img > p->hw) blocks_per_img = p->hw;
    if (blocks_per_img < 1) return false;
    rows_per_block = (p->hw + blocks_per_img - 1) / blocks_per_img;
    blocks_per_img = (p->hw + rows_per_block - 1) / rows_per_block;
    smem = (size_t)rows_per_block * p->c * 2 + 2 * (size_t)p->c * sizeof(float);
    return smem <= (size_t)kGnFusedMaxSmem;
}

struct LnArgs {
    const uint16_t* x;
    uint16_t* y;
    const float* gamma;
    const float* beta;
    int rows, c, ldx, ldy, nvec;
    float eps;
    int dtype;
};

constexpr int kLnMaxVec = 8;  // c <= 8 * 32 * 8 = 2048

__global__ void __launch_bounds__(256) layer_norm_kernel(const LnArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= a.rows) return;
    const uint16_t* xr = a.x + (size_t)warp * a.ldx;
    float v[kLnMaxVec][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
        const int vec = lane + j * 32;
        if (vec < a.nvec) {
            const uint4 u = *reinterpret_cast<const uint4*>(xr + vec * 8);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack2(w[i], a.dtype);
                v[j][2 * i] = f.x; v[j][2 * i + 1] = f.y;
                sum += f.x + f.y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)a.c;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
        if (lane + j * 32 < a.nvec) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; sq += d * d; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)a.c + a.eps);
    uint16_t* yr = a.y + (size_t)warp * a.ldy;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
        const int vec = lane + j * 32;
        if (vec < a.nvec) {
            const float4 g0 = *reinterpret_cast<const float4*>(a.gamma + vec * 8);
            const float4 g1 = *reinterpret_cast<const float4*>(a.gamma + vec * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(a.beta + vec * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(a.beta + vec * 8 + 4);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                o[i] = pack2((v[j][2 * i] - mean) * rstd * g[2 * i] + b[2 * i],
                             (v[j][2 * i + 1] - mean) * rstd * g[2 * i + 1] + b[2 * i + 1], a.dtype);
            *reinterpret_cast<uint4*>(yr + vec * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

static int make_gn_args(const sfb_gn_params* p, GnArgs& a, int& blocks_per_img) {
    if (!p || !p->x || !p->stats) return fail(SFB_ERR_INVALID, "group_norm: null argument");
    if (p->c % 8 || p->ldx % 8 || p->groups <= 0 || p->groups > 64 || p->c % p->groups ||
        p->c / 8 > kGnThreads || p->c > 4096)
        return fail(SFB_ERR_INVALID, "group_norm: unsupported geometry c=%d groups=%d ldx=%d", p->c, p->groups, p->ldx);
    a.x = reinterpret_cast<const uint16_t*>(p->x);
    a.y = reinterpret_cast<uint16_t*>(p->y);
    a.gamma = p->gamma; a.beta = p->beta; a.stats = p->stats;
    a.n = p->n; a.hw = p->hw; a.c = p->c; a.ldx = p->ldx; a.ldy = p->ldy; a.groups = p->groups;
    a.cpg = p->c / p->groups; a.nvec = p->c / 8; a.eps = p->eps; a.silu = p->silu; a.dtype = p->dtype;
    a.shards = p->stat_shards > 0 ? p->stat_shards : 1;
    a.shard_stride = p->stat_shard_stride;
    const int by = kGnThreads / a.nvec;
    // aim for >= 2 waves of 148 SMs while giving each thread a few rows
    int want = (2 * 148 + p->n - 1) / p->n;
    int max_blocks = (p->hw + by - 1) / by;
    blocks_per_img = want < max_blocks ? want : max_blocks;
    if (blocks_per_img < 1) blocks_per_img = 1;
    a.rows_per_block = (p->hw + blocks_per_img - 1) / blocks_per_img;
    blocks_per_img = (p->hw + a.rows_per_block - 1) / a.rows_per_block;
    return SFB_OK;
}

}  // namespace sfb

using namespace sfb;

extern "C" int sfb_group_norm_stats(const sfb_gn_params* p, sfb_stream_t stream) {
    GnArgs a{};
    int bpi = 1;
    int rc = make_gn_args(p, a, bpi);
    if (rc) return rc;
    cudaError_t err = launch_pdl(gn_stats_kernel, dim3(bpi, p->n), dim3(kGnThreads), 0,
                                 static_cast<cudaStream_t>(stream), a);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "group_norm_stats: %s", cudaGetErrorString(err));
    return check_launch("sfb_group_norm_stats");
}

extern "C" int sfb_group_norm_apply(const sfb_gn_params* p, sfb_stream_t stream) {
    GnArgs a{};
    int bpi = 1;
    int rc = make_gn_args(p, a, bpi);
    if (rc) return rc;
    if (!p->y || !p->gamma || !p->beta || p->ldy % 8) return fail(SFB_ERR_INVALID, "group_norm_apply: null/ldy");
    cudaError_t err = launch_pdl(gn_apply_kernel, dim3(bpi, p->n), dim3(kGnThreads),
                                 2 * p->c * sizeof(float), static_cast<cudaStream_t>(stream), a);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "group_norm_apply: %s", cudaGetErrorString(err));
    return check_launch("sfb_group_norm_apply");
}


extern "C" int sfb_group_norm_fused_fits(const sfb_gn_params* p) {
    int bpi, rpb;
    size_t smem;
    if (!p || p->c % 8 || p->c <= 0 || p->groups <= 0 || p->c % p->groups || p->c / 8 > kGnThreads) return 0;
    if (gn_group_fits(p, smem)) return 1;
    return gn_fused_geometry(p, bpi, rpb, smem) ? 1 : 0;
}

extern "C" int sfb_group_norm_fused(const sfb_gn_params* p, sfb_stream_t stream) {
    GnArgs a{};
    int bpi = 1;
    int rc = make_gn_args(p, a, bpi);
    if (rc) return rc;
    if (!p->y || !p->gamma || !p->beta || p->ldy % 8 || !p->sync_counter)
        return fail(SFB_ERR_INVALID, "group_norm_fused: null/ldy/sync_counter");
    int rpb = 0;
    size_t smem;
    const bool per_group = gn_group_fits(p, smem);
    if (!per_group && !gn_fused_geometry(p, bpi, rpb, smem))
        return fail(SFB_ERR_INVALID, "group_norm_fused: tensor does not fit in shared memory");
    a.rows_per_block = rpb;
    if (p->part_splits > 1) {
        if (!p->part_ws || p->part_c <= 0 || p->part_c % 8 || p->part_c > p->c || p->part_ld % 4 ||
            p->part_ld < p->part_c || (p->part_residual && p->part_ldr % 8))
            return fail(SFB_ERR_INVALID, "group_norm_fused: bad deferred split-K description");
        a.part_ws = p->part_ws; a.part_splits = p->part_splits; a.part_c = p->part_c; a.part_ld = p->part_ld;
        a.part_bias = p->part_bias; a.part_rowbias = p->part_rowbias; a.part_ld_rowbias = p->part_ld_rowbias;
        a.part_residual = reinterpret_cast<const uint16_t*>(p->part_residual); a.part_ldr = p->part_ldr;
    }
    static PerDeviceOnce attr_once;
    bool& attr_set = attr_once.flag();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gn_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kGnFusedMaxSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gn_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kGnFusedMaxSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gn_group_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kGnFusedMaxSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gn_group_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kGnFusedMaxSmem);
        if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "group_norm_fused: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    if (per_group) {
        // one CTA per (image, group): no grid barrier, no statistics buffer
        const dim3 grid(p->groups, p->n);
        cudaError_t err = a.part_splits > 1
            ? launch_pdl(gn_group_kernel<true>, grid, dim3(kGnGroupThreads), smem, static_cast<cudaStream_t>(stream), a)
            : launch_pdl(gn_group_kernel<false>, grid, dim3(kGnGroupThreads), smem, static_cast<cudaStream_t>(stream), a);
        if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "group_norm_fused(per group): %s", cudaGetErrorString(err));
        return check_launch("sfb_group_norm_fused");
    }
    cudaError_t err = a.part_splits > 1
        ? launch_pdl(gn_fused_kernel<true>, dim3(bpi, p->n), dim3(kGnThreads), smem,
                     static_cast<cudaStream_t>(stream), a, p->sync_counter)
        : launch_pdl(gn_fused_kernel<false>, dim3(bpi, p->n), dim3(kGnThreads), smem,
                     static_cast<cudaStream_t>(stream), a, p->sync_counter);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "group_norm_fused: %s", cudaGetErrorString(err));
    return check_launch("sfb_group_norm_fused");
}

extern "C" int sfb_layer_norm(const sfb_ln_params* p, sfb_stream_t stream) {
    if (!p || !p->x || !p->y || !p->gamma || !p->beta) return fail(SFB_ERR_INVALID, "layer_norm: null argument");
    if (p->c % 8 || p->c > kLnMaxVec * 256 || p->ldx % 8 || p->ldy % 8 || p->rows <= 0)
        return fail(SFB_ERR_INVALID, "layer_norm: unsupported geometry c=%d", p->c);
    LnArgs a{};
    a.x = reinterpret_cast<const uint16_t*>(p->x);
    a.y = reinterpret_cast<uint16_t*>(p->y);
    a.gamma = p->gamma; a.beta = p->beta; a.rows = p->rows; a.c = p->c; a.ldx = p->ldx; a.ldy = p->ldy;
    a.nvec = p->c / 8; a.eps = p->eps; a.dtype = p->dtype;
    const int blocks = (p->rows + 7) / 8;
    cudaError_t err = launch_pdl(layer_norm_kernel, dim3(blocks), dim3(256), 0,
                                 static_cast<cudaStream_t>(stream), a);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "layer_norm: %s", cudaGetErrorString(err));
    return check_launch("sfb_layer_norm");
}
