// tcgen05 flash attention for sm_100a:  O = softmax(Q K^T * scale) V, one (batch, head, 128-query
// tile) per CTA, online softmax over 128-wide KV tiles.
//
// Warp roles (192 threads):
//   warp 0      TMA producer: Q once, then a ring of K tiles and a ring of V^T tiles;
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer:
//                 S = Q K^T  (128 x 128, fp32, TMEM columns [0,128))
//                 O += P V   (128 x DV,  fp32, TMEM columns [128,128+DV))
//               QK^T of tile j+1 is issued as soon as the softmax warps have pulled S_j into
//               registers, so it overlaps the exponentials of tile j;
//   warps 2..5  softmax: thread = query row; tcgen05.ld S, running max in fp32, one FFMA
//               (s*scale*log2e - max) + packed 16-bit ex2 per element, P written to shared memory
//               in the 128-byte-swizzled K-major layout tcgen05.mma reads as its A operand, O
//               rescaled in TMEM only when the running max moved.  The softmax denominator costs
//               nothing: row `head_dim` of V^T is all ones, so column `head_dim` of O accumulates
//               sum(P) on the tensor core, consistently with the rounded P and the O rescales.
// All three operands are K-major: Q/K as [rows, head_dim] and V pre-transposed as V^T
// [head_dim, seq] -- the QKV projection's epilogue (gemm_tc.cu, SFB_EPI_QKV) writes those
// layouts directly, so no permute/copy kernels exist between projection and attention.
//
// Replaces xformers.ops.memory_efficient_attention as wrapped at
// /root/reference/src/sfast/libs/xformers/xformers_attention.py:26-63.
#include "common.cuh"
#include "host.h"

#include <stdlib.h>
#include <string.h>

namespace sfb {

constexpr int kAttnThreads = 192;
constexpr int kTileQ = 128;
constexpr int kTileKV = 128;

struct AttnArgs {
    void* out;
    int batch, heads, head_dim;
    int seq_q, seq_kv;
    int q_rows, k_rows, vt_rows;
    int dtype;
    float scale_log2;  // scale * log2(e)
    int causal;        // v2 only: key j is visible to query i iff j <= i
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA / ALU pipes instead of the MUFU (special-function) pipe: Cody-Waite split
// x = n + f, f in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial (max relative error 1.0e-4,
// below the half-ulp of the 16-bit probability it is rounded to), 2^n by an integer add into the
// exponent.  The softmax of these kernels is bound by MUFU.EX2 (16 lanes / clk / SM against 128 for
// FMA): one exponential per score with head_dim 40 / 64 means one MUFU op per 160 / 256 tensor FLOPs.
// Sending a fixed fraction of every row's exponentials through this routine (kPolyMask: which of each
// 8 consecutive scores) trades 8 extra issue slots per element for a free MUFU slot; the balance
// point of the two pipes was estimated at ~30 %, measured lower (below).  x <= ~8 here (lazy maximum).
__device__ __forceinline__ float poly_exp2(float x) {
    x = fmaxf(x, -125.0f);                       // 2^-125: a zero probability after rounding
    const float t = x + 12582912.0f;             // 1.5 * 2^23: round(x) lands in the low mantissa bits
    const float f = x - (t - 12582912.0f);       // [-0.5, 0.5]
    float p = fmaf(f, 0.05500893f, 0.24221096f);
    p = fmaf(p, f, 0.69328293f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// Measured on B200 (profiles/r02_attention_poly_exp_split.jsonl, S = 4096): the 128-key-tile kernel
// (head_dim 40 / 80 / 160) gains 8.8 % with 1 of every 8 exponentials on the FMA pipe (119.4 -> 108.9 us
// at head_dim 40, B = 2; 25 %: 112.1, 37.5 %: 113.5 -- the extra issue slots then cost more than the
// freed MUFU slots); the 64-key-tile kernel (head_dim 64) LOSES with any split (596 -> 618 / 646 / 670 us):
// with two score tiles in flight it is bound by issue slots, not by MUFU.  Hence per-kernel masks.
#ifndef SFB_EXP_POLY_MASK
#define SFB_EXP_POLY_MASK 0x08   // element 3 of every 8 scores: 12.5 % of the exponentials
#endif
constexpr unsigned kPolyMaskV1 = SFB_EXP_POLY_MASK;  // attention_tc_kernel
constexpr unsigned kPolyMaskV2 = 0x00;               // attention_v2_kernel
template <int I, unsigned MASK>
__device__ __forceinline__ float softmax_exp2(float x) {
    if constexpr ((MASK >> (I & 7)) & 1u) return poly_exp2(x);
    else return fast_exp2(x);
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int DC, int DK, int DV, int KVS>
struct AttnSmem {
    static constexpr int kQBytes = DC * kTileQ * 128;
    static constexpr int kKStage = DC * kTileKV * 128;
    static constexpr int kVChunk = DV * 128;  // [DV rows x 64 kv columns]
    static constexpr int kVStage = 2 * kVChunk;
    static constexpr int kPBytes = 2 * kTileQ * 128;
    static constexpr int kOffK = kQBytes;
    static constexpr int kOffV = kOffK + KVS * kKStage;
    static constexpr int kOffP = kOffV + KVS * kVStage;
    static constexpr int kOffBar = kOffP + kPBytes;
    static constexpr int kTotal = kOffBar + 256 + 1024;
    static_assert(kVChunk % 1024 == 0, "V^T chunk must keep 1024-byte swizzle-atom alignment");
};

template <int DC, int DK, int DV, int KVS, int BF16>
__global__ void __launch_bounds__(kAttnThreads, (DC == 1) ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tma_q,
                    const __grid_constant__ CUtensorMap tma_k,
                    const __grid_constant__ CUtensorMap tma_vt, const AttnArgs a) {
    using L = AttnSmem<DC, DK, DV, KVS>;
    constexpr uint32_t kTmemCols = (128 + DV <= 256) ? 256 : 512;
    constexpr uint32_t kColS = 0, kColO = 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + L::kOffK;
    uint8_t* sV = smem + L::kOffV;
    uint8_t* sP = smem + L::kOffP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + KVS;
    uint64_t* v_full = k_empty + KVS;
    uint64_t* v_empty = v_full + KVS;
    uint64_t* s_full = v_empty + KVS;
    uint64_t* s_free = s_full + 1;
    uint64_t* p_full = s_free + 1;
    uint64_t* pv_done = p_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x;
    const int bh = blockIdx.y;
    const int n_kv = (a.seq_kv + kTileKV - 1) / kTileKV;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_k);
        tma_prefetch_desc(&tma_vt);
        mbar_init(q_full, 1);
        for (int i = 0; i < KVS; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 128);
        mbar_init(p_full, 128);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // only this thread reads the predecessor's output (Q/K/V via TMA); everything the other
            // warps do is ordered behind these loads
            pdl_wait();
            mbar_expect_tx(q_full, L::kQBytes);
#pragma unroll
            for (int c = 0; c < DC; ++c)
                tma_load_2d(sQ + c * (kTileQ * 128), &tma_q, q_full, c * 64,
                            bh * a.q_rows + q_tile * kTileQ);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % KVS;
                const uint32_t ph = (j / KVS) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], L::kKStage);
#pragma unroll
                for (int c = 0; c < DC; ++c)
                    tma_load_2d(sK + st * L::kKStage + c * (kTileKV * 128), &tma_k, &k_full[st],
                                c * 64, bh * a.k_rows + j * kTileKV);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], L::kVStage);
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(sV + st * L::kVStage + c * L::kVChunk, &tma_vt, &v_full[st],
                                j * kTileKV + c * 64, bh * a.vt_rows);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr bool bf16 = BF16 != 0;
            const uint32_t idesc_s = umma_idesc_f16(kTileQ, kTileKV, bf16);
            const uint32_t idesc_o = umma_idesc_f16(kTileQ, DV, bf16);
            const uint32_t tS = tmem_base + kColS;
            const uint32_t tO = tmem_base + kColO;
            auto issue_qk = [&](int j) {
                const int st = j % KVS;
#pragma unroll
                for (int kk = 0; kk < DK / 16; ++kk) {
                    const int c = kk / 4, k4 = kk % 4;
                    const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ + c * (kTileQ * 128)));
                    const uint64_t dk = umma_desc_k_sw128(
                        smem_u32(sK + st * L::kKStage + c * (kTileKV * 128)));
                    umma_f16_ss(tS, dq + (uint64_t)(k4 * 2), dk + (uint64_t)(k4 * 2), idesc_s,
                                kk != 0);
                }
                umma_commit(&k_empty[st]);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_qk(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) {
                    const int st1 = (j + 1) % KVS;
                    mbar_wait(&k_full[st1], ((j + 1) / KVS) & 1);
                    mbar_wait(s_free, j & 1);  // S_j is in the softmax warps' registers
                    tc_fence_after();
                    issue_qk(j + 1);
                }
                const int st = j % KVS;
                mbar_wait(&v_full[st], (j / KVS) & 1);
                mbar_wait(p_full, j & 1);  // P_j in smem, O rescaled
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < kTileKV / 16; ++kk) {
                    const int c = kk / 4, k4 = kk % 4;
                    const uint64_t dp = umma_desc_k_sw128(smem_u32(sP + c * (kTileQ * 128)));
                    const uint64_t dv =
                        umma_desc_k_sw128(smem_u32(sV + st * L::kVStage + c * L::kVChunk));
                    umma_f16_ss(tO, dp + (uint64_t)(k4 * 2), dv + (uint64_t)(k4 * 2), idesc_o,
                                (j > 0) || (kk != 0));
                }
                umma_commit(&v_empty[st]);
                umma_commit(pv_done);
            }
        }
        __syncwarp();
    } else {
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + kColS;
        const uint32_t tO = tmem_base + lane_base + kColO;
        float m_run = -INFINITY;
        const float sl2 = a.scale_log2;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t sraw[128];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t tmp[32];
                tmem_ld32(tS + c * 32, tmp);
#pragma unroll
                for (int i = 0; i < 32; ++i) sraw[c * 32 + i] = tmp[i];
            }
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(s_free);

            // ragged tail: only the last KV tile can hold columns >= seq_kv
            const int n_valid = a.seq_kv - j * kTileKV;  // >= 1
            if (n_valid < kTileKV) {
#pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i >= n_valid) sraw[i] = 0xff800000u;  // -inf
            }
            // four independent FMNMX3 chains: a single 128-long max chain is pure latency
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
                mx0 = fmax3(mx0, __uint_as_float(sraw[i]), __uint_as_float(sraw[i + 1]));
                mx1 = fmax3(mx1, __uint_as_float(sraw[i + 2]), __uint_as_float(sraw[i + 3]));
                mx2 = fmax3(mx2, __uint_as_float(sraw[i + 4]), __uint_as_float(sraw[i + 5]));
                mx3 = fmax3(mx3, __uint_as_float(sraw[i + 6]), __uint_as_float(sraw[i + 7]));
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            // Lazy reference maximum: the softmax is invariant to the reference (O and the
            // denominator, which is column head_dim of O, carry the same factor), so the stale
            // reference is kept until a row's maximum exceeds it by more than 2^kLazyLog2 -- the
            // probabilities then stay <= 2^kLazyLog2 (exact in 16-bit floating point, fp32
            // accumulation) and the O rescale in TMEM, which an exact running maximum triggers in
            // ~80 % of the tiles of a 32-row warp, almost never runs after the first tile.
            constexpr float kLazyLog2 = 8.0f;
            const float m_cand = mx * sl2;  // scale > 0
            const bool moved = m_cand > m_run + kLazyLog2;  // always true on the first tile (-inf)
            const float m_new = moved ? m_cand : m_run;
            const float alpha = moved ? fast_exp2(m_run - m_new) : 1.0f;  // 0 on the first tile
            m_run = m_new;
            const float neg_m = -m_new;

            // all 128 exponentials first (into 64 packed registers) ...
            uint4 pk[16];
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaf(__uint_as_float(sraw[g * 8 + i]), sl2, neg_m);
                pk[g].x = pack2(softmax_exp2<0, kPolyMaskV1>(x[0]), softmax_exp2<1, kPolyMaskV1>(x[1]), BF16);
                pk[g].y = pack2(softmax_exp2<2, kPolyMaskV1>(x[2]), softmax_exp2<3, kPolyMaskV1>(x[3]), BF16);
                pk[g].z = pack2(softmax_exp2<4, kPolyMaskV1>(x[4]), softmax_exp2<5, kPolyMaskV1>(x[5]), BF16);
                pk[g].w = pack2(softmax_exp2<6, kPolyMaskV1>(x[6]), softmax_exp2<7, kPolyMaskV1>(x[7]), BF16);
            }
            // ... and only then wait for PV(j-1): the P buffer / O accumulator are not touched before,
            // so the previous tile's second MMA overlaps this tile's MUFU work
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);  // P buffer free, O stable
                tc_fence_after();
            }
#pragma unroll
            for (int g = 0; g < 16; ++g) {  // 16 groups of 8 kv columns -> one 16-byte store each
                const int chunk = g >> 3;  // which 64-column half
                const int g8 = g & 7;
                uint8_t* dst = sP + chunk * (kTileQ * 128) + r * 128 + ((g8 ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(dst) = pk[g];
            }

            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
                for (int c = 0; c < DV / 16; ++c) {
                    uint32_t o[16];
                    tmem_ld16(tO + c * 16, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st16(tO + c * 16, o);
                }
                tmem_wait_st();
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // final: O / l, with l = O[:, head_dim] (the ones row of V^T)
        mbar_wait(pv_done, (n_kv - 1) & 1);
        tc_fence_after();
        const int srow = q_tile * kTileQ + r;
        const bool valid = srow < a.seq_q;
        const int b = bh / a.heads, h = bh % a.heads;
        float inv_l;
        {
            uint32_t o[16];
            tmem_ld16(tO + (a.head_dim & ~15), o);
            tmem_wait_ld();
            float l = 1.f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i == (a.head_dim & 15)) l = __uint_as_float(o[i]);
            inv_l = 1.0f / l;
        }
        uint16_t* orow = reinterpret_cast<uint16_t*>(a.out) +
                         ((size_t)b * a.seq_q + (valid ? srow : 0)) * (a.heads * a.head_dim) +
                         h * a.head_dim;
#pragma unroll 1
        for (int c = 0; c < DV / 16; ++c) {
            uint32_t o[16];
            tmem_ld16(tO + c * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int jv = 0; jv < 2; ++jv) {
                const int d = c * 16 + jv * 8;
                if (valid && d < a.head_dim) {
                    uint4 pk;
                    pk.x = pack2(__uint_as_float(o[jv * 8 + 0]) * inv_l, __uint_as_float(o[jv * 8 + 1]) * inv_l, BF16);
                    pk.y = pack2(__uint_as_float(o[jv * 8 + 2]) * inv_l, __uint_as_float(o[jv * 8 + 3]) * inv_l, BF16);
                    pk.z = pack2(__uint_as_float(o[jv * 8 + 4]) * inv_l, __uint_as_float(o[jv * 8 + 5]) * inv_l, BF16);
                    pk.w = pack2(__uint_as_float(o[jv * 8 + 6]) * inv_l, __uint_as_float(o[jv * 8 + 7]) * inv_l, BF16);
                    *reinterpret_cast<uint4*>(orow + d) = pk;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

template <int DC, int DK, int DV, int KVS, int BF16>
static int launch_attention_t(const sfb_attn_params* p, const AttnArgs& a, cudaStream_t stream) {
    using L = AttnSmem<DC, DK, DV, KVS>;
    static PerDeviceOnce attr_once;
    bool& attr_set = attr_once.flag();
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(attention_tc_kernel<DC, DK, DV, KVS, BF16>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (err != cudaSuccess)
            return fail(SFB_ERR_CUDA, "sfb_attention: smem attribute: %s", cudaGetErrorString(err));
        attr_set = true;
    }
    CUtensorMap tq, tk, tv;
    memcpy(&tq, p->tmap_q, sizeof(CUtensorMap));
    memcpy(&tk, p->tmap_k, sizeof(CUtensorMap));
    memcpy(&tv, p->tmap_vt, sizeof(CUtensorMap));
    dim3 grid((p->seq_q + kTileQ - 1) / kTileQ, p->batch * p->heads);
    cudaError_t err = launch_pdl(attention_tc_kernel<DC, DK, DV, KVS, BF16>, grid, dim3(kAttnThreads),
                                 L::kTotal, stream, tq, tk, tv, a);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "sfb_attention: launch: %s", cudaGetErrorString(err));
    return check_launch("sfb_attention");
}

template <int DC, int DK, int DV, int KVS>
static int launch_attention(const sfb_attn_params* p, const AttnArgs& a, cudaStream_t stream) {
    return a.dtype == SFB_BF16 ? launch_attention_t<DC, DK, DV, KVS, 1>(p, a, stream)
                               : launch_attention_t<DC, DK, DV, KVS, 0>(p, a, stream);
}


// ---------------------------------------------------------------------------------------
// v2 (head_dim <= 64): 64-key tiles with the score tile DOUBLE-BUFFERED in TMEM (S0 | S1 | O) and
// the probability tile double-buffered in shared memory.  In v1 the four softmax warps of a CTA
// meet the tensor core twice per tile through single buffers (S: QK^T(j+1) cannot start before all
// of them pulled S_j; P: nobody may write P_{j+1} before PV_j has read P_j), so the slowest warp
// paces the others and the MUFU pipe idles ~40 % of the time.  With two buffers each the MMA
// thread runs up to two score tiles ahead and the softmax warps one probability tile ahead.
// ---------------------------------------------------------------------------------------
constexpr int kTileKV2 = 64;

template <int DV, int KVS>
struct AttnSmem2 {
    static constexpr int kQBytes = kTileQ * 128;
    static constexpr int kKStage = kTileKV2 * 128;  // 64 keys x 64 (padded) head dims
    static constexpr int kVStage = DV * 128;        // [DV rows x 64 kv columns]
    static constexpr int kPBytes = kTileQ * 128;    // 128 rows x 64 probabilities
    static constexpr int kOffK = kQBytes;
    static constexpr int kOffV = kOffK + KVS * kKStage;
    static constexpr int kOffP = kOffV + KVS * kVStage;
    static constexpr int kOffBar = kOffP + 2 * kPBytes;
    static constexpr int kTotal = kOffBar + 512 + 1024;
    static_assert(kVStage % 1024 == 0, "V^T stage must keep 1024-byte swizzle-atom alignment");
    static_assert(2 * (kTotal + 1024) <= 228 * 1024, "two CTAs per SM");
};

template <int DK, int DV, int KVS, int BF16>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_v2_kernel(const __grid_constant__ CUtensorMap tma_q,
                    const __grid_constant__ CUtensorMap tma_k,
                    const __grid_constant__ CUtensorMap tma_vt, const AttnArgs a) {
    using L = AttnSmem2<DV, KVS>;
    constexpr uint32_t kTmemCols = 256;
    constexpr uint32_t kColS = 0, kColO = 128;  // S0 = [0, 64), S1 = [64, 128), O = [128, 128 + DV)
    static_assert(128 + DV <= 256, "TMEM budget");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + L::kOffK;
    uint8_t* sV = smem + L::kOffV;
    uint8_t* sP = smem + L::kOffP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + KVS;
    uint64_t* v_full = k_empty + KVS;
    uint64_t* v_empty = v_full + KVS;
    uint64_t* s_full = v_empty + KVS;  // [2] score buffer b holds QK^T of its current tile
    uint64_t* s_free = s_full + 2;     // [2] all 128 softmax threads pulled it into registers
    uint64_t* p_full = s_free + 2;     // [2] probability buffer b written (and O rescaled)
    uint64_t* p_free = p_full + 2;     // [2] the PV MMA that read buffer b has completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x;
    const int bh = blockIdx.y;
    const int n_kv = (a.seq_kv + kTileKV2 - 1) / kTileKV2;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_k);
        tma_prefetch_desc(&tma_vt);
        mbar_init(q_full, 1);
        for (int i = 0; i < KVS; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 128);
            mbar_init(&p_full[i], 128);
            mbar_init(&p_free[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            pdl_wait();
            mbar_expect_tx(q_full, L::kQBytes);
            tma_load_2d(sQ, &tma_q, q_full, 0, bh * a.q_rows + q_tile * kTileQ);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % KVS;
                const uint32_t ph = (j / KVS) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], L::kKStage);
                tma_load_2d(sK + st * L::kKStage, &tma_k, &k_full[st], 0, bh * a.k_rows + j * kTileKV2);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], L::kVStage);
                tma_load_2d(sV + st * L::kVStage, &tma_vt, &v_full[st], j * kTileKV2, bh * a.vt_rows);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr bool bf16 = BF16 != 0;
            const uint32_t idesc_s = umma_idesc_f16(kTileQ, kTileKV2, bf16);
            const uint32_t idesc_o = umma_idesc_f16(kTileQ, DV, bf16);
            const uint32_t tO = tmem_base + kColO;
            const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ));
            auto issue_qk = [&](int j) {  // S[j & 1] = Q K_j^T
                const int st = j % KVS;
                const uint32_t tS = tmem_base + kColS + (uint32_t)(j & 1) * kTileKV2;
                const uint64_t dk = umma_desc_k_sw128(smem_u32(sK + st * L::kKStage));
#pragma unroll
                for (int kk = 0; kk < DK / 16; ++kk)
                    umma_f16_ss(tS, dq + (uint64_t)(kk * 2), dk + (uint64_t)(kk * 2), idesc_s, kk != 0);
                umma_commit(&k_empty[st]);
                umma_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_qk(0);
            if (n_kv > 1) {
                mbar_wait(&k_full[1 % KVS], (1 / KVS) & 1);
                tc_fence_after();
                issue_qk(1);
            }
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % KVS, b = j & 1;
                const uint32_t ph2 = (j >> 1) & 1;  // phase of the per-buffer barriers for tile j
                mbar_wait(&v_full[st], (j / KVS) & 1);
                mbar_wait(&p_full[b], ph2);  // P_j in shared memory, O rescaled if needed
                tc_fence_after();
                const uint64_t dp = umma_desc_k_sw128(smem_u32(sP + b * L::kPBytes));
                const uint64_t dv = umma_desc_k_sw128(smem_u32(sV + st * L::kVStage));
#pragma unroll
                for (int kk = 0; kk < kTileKV2 / 16; ++kk)
                    umma_f16_ss(tO, dp + (uint64_t)(kk * 2), dv + (uint64_t)(kk * 2), idesc_o,
                                (j > 0) || (kk != 0));
                umma_commit(&v_empty[st]);
                umma_commit(&p_free[b]);
                if (j + 2 < n_kv) {  // score buffer b was pulled into registers long ago: refill it
                    const int st2 = (j + 2) % KVS;
                    mbar_wait(&k_full[st2], ((j + 2) / KVS) & 1);
                    mbar_wait(&s_free[b], ph2);
                    tc_fence_after();
                    issue_qk(j + 2);
                }
            }
        }
        __syncwarp();
    } else {
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const uint32_t tO = tmem_base + lane_base + kColO;
        float m_run = -INFINITY;
        const float sl2 = a.scale_log2;
        for (int j = 0; j < n_kv; ++j) {
            const int b = j & 1;
            const uint32_t ph2 = (j >> 1) & 1;
            const uint32_t tS = tmem_base + lane_base + kColS + (uint32_t)b * kTileKV2;
            mbar_wait(&s_full[b], ph2);
            tc_fence_after();
            uint32_t sraw[64];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t tmp[32];
                tmem_ld32(tS + c * 32, tmp);
#pragma unroll
                for (int i = 0; i < 32; ++i) sraw[c * 32 + i] = tmp[i];
            }
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(&s_free[b]);

            int n_valid = a.seq_kv - j * kTileKV2;  // >= 1; < 64 only on the last tile
            // causal (text encoders): this row sees keys 0 .. its own position.  Key 0 is visible to
            // every row, so tile 0 always moves the running maximum off -inf; a later tile that is
            // entirely masked for a row leaves exp2(-inf) = 0 everywhere.
            if (a.causal) n_valid = min(n_valid, q_tile * kTileQ + r - j * kTileKV2 + 1);
            if (n_valid < kTileKV2) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i >= n_valid) sraw[i] = 0xff800000u;  // -inf
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 64; i += 8) {
                mx0 = fmax3(mx0, __uint_as_float(sraw[i]), __uint_as_float(sraw[i + 1]));
                mx1 = fmax3(mx1, __uint_as_float(sraw[i + 2]), __uint_as_float(sraw[i + 3]));
                mx2 = fmax3(mx2, __uint_as_float(sraw[i + 4]), __uint_as_float(sraw[i + 5]));
                mx3 = fmax3(mx3, __uint_as_float(sraw[i + 6]), __uint_as_float(sraw[i + 7]));
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            // lazy reference maximum (see v1)
            constexpr float kLazyLog2 = 8.0f;
            const float m_cand = mx * sl2;
            const bool moved = m_cand > m_run + kLazyLog2;
            const float m_new = moved ? m_cand : m_run;
            const float alpha = moved ? fast_exp2(m_run - m_new) : 1.0f;
            m_run = m_new;
            const float neg_m = -m_new;

            uint4 pk[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaf(__uint_as_float(sraw[g * 8 + i]), sl2, neg_m);
                pk[g].x = pack2(softmax_exp2<0, kPolyMaskV2>(x[0]), softmax_exp2<1, kPolyMaskV2>(x[1]), BF16);
                pk[g].y = pack2(softmax_exp2<2, kPolyMaskV2>(x[2]), softmax_exp2<3, kPolyMaskV2>(x[3]), BF16);
                pk[g].z = pack2(softmax_exp2<4, kPolyMaskV2>(x[4]), softmax_exp2<5, kPolyMaskV2>(x[5]), BF16);
                pk[g].w = pack2(softmax_exp2<6, kPolyMaskV2>(x[6]), softmax_exp2<7, kPolyMaskV2>(x[7]), BF16);
            }
            // probability buffer b was last read by PV(j-2)
            if (j >= 2) {
                mbar_wait(&p_free[b], ((j >> 1) - 1) & 1);
                tc_fence_after();
            }
            uint8_t* prow = sP + b * L::kPBytes + r * 128;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                *reinterpret_cast<uint4*>(prow + ((g ^ (r & 7)) << 4)) = pk[g];

            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
                // O must be quiescent: PV(j-1) complete (PV(j) is only issued after p_full below)
                mbar_wait(&p_free[b ^ 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < DV / 16; ++c) {
                    uint32_t o[16];
                    tmem_ld16(tO + c * 16, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st16(tO + c * 16, o);
                }
                tmem_wait_st();
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&p_full[b]);
        }
        // final: O / l, with l = O[:, head_dim] (the ones row of V^T)
        mbar_wait(&p_free[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
        tc_fence_after();
        const int srow = q_tile * kTileQ + r;
        const bool valid = srow < a.seq_q;
        const int bb = bh / a.heads, h = bh % a.heads;
        float inv_l;
        {
            uint32_t o[16];
            tmem_ld16(tO + (a.head_dim & ~15), o);
            tmem_wait_ld();
            float l = 1.f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i == (a.head_dim & 15)) l = __uint_as_float(o[i]);
            inv_l = 1.0f / l;
        }
        uint16_t* orow = reinterpret_cast<uint16_t*>(a.out) +
                         ((size_t)bb * a.seq_q + (valid ? srow : 0)) * (a.heads * a.head_dim) +
                         h * a.head_dim;
#pragma unroll 1
        for (int c = 0; c < DV / 16; ++c) {
            uint32_t o[16];
            tmem_ld16(tO + c * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int jv = 0; jv < 2; ++jv) {
                const int d = c * 16 + jv * 8;
                if (valid && d < a.head_dim) {
                    uint4 w;
                    w.x = pack2(__uint_as_float(o[jv * 8 + 0]) * inv_l, __uint_as_float(o[jv * 8 + 1]) * inv_l, BF16);
                    w.y = pack2(__uint_as_float(o[jv * 8 + 2]) * inv_l, __uint_as_float(o[jv * 8 + 3]) * inv_l, BF16);
                    w.z = pack2(__uint_as_float(o[jv * 8 + 4]) * inv_l, __uint_as_float(o[jv * 8 + 5]) * inv_l, BF16);
                    w.w = pack2(__uint_as_float(o[jv * 8 + 6]) * inv_l, __uint_as_float(o[jv * 8 + 7]) * inv_l, BF16);
                    *reinterpret_cast<uint4*>(orow + d) = w;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

template <int DK, int DV, int KVS, int BF16>
static int launch_attention_v2_t(const sfb_attn_params* p, const AttnArgs& a, cudaStream_t stream) {
    using L = AttnSmem2<DV, KVS>;
    static PerDeviceOnce attr_once;
    bool& attr_set = attr_once.flag();
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(attention_v2_kernel<DK, DV, KVS, BF16>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (err != cudaSuccess)
            return fail(SFB_ERR_CUDA, "sfb_attention(v2): smem attribute: %s", cudaGetErrorString(err));
        attr_set = true;
    }
    CUtensorMap tq, tk, tv;
    memcpy(&tq, p->tmap_q, sizeof(CUtensorMap));
    memcpy(&tk, p->tmap_k, sizeof(CUtensorMap));
    memcpy(&tv, p->tmap_vt, sizeof(CUtensorMap));
    dim3 grid((p->seq_q + kTileQ - 1) / kTileQ, p->batch * p->heads);
    cudaError_t err = launch_pdl(attention_v2_kernel<DK, DV, KVS, BF16>, grid, dim3(kAttnThreads),
                                 L::kTotal, stream, tq, tk, tv, a);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "sfb_attention(v2): launch: %s", cudaGetErrorString(err));
    return check_launch("sfb_attention");
}

template <int DK, int DV, int KVS>
static int launch_attention_v2(const sfb_attn_params* p, const AttnArgs& a, cudaStream_t stream) {
    return a.dtype == SFB_BF16 ? launch_attention_v2_t<DK, DV, KVS, 1>(p, a, stream)
                               : launch_attention_v2_t<DK, DV, KVS, 0>(p, a, stream);
}

}  // namespace sfb

using namespace sfb;

extern "C" int sfb_attention(const sfb_attn_params* p, sfb_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!p || !p->tmap_q || !p->tmap_k || !p->tmap_vt || !p->out)
        return fail(SFB_ERR_INVALID, "sfb_attention: null argument");
    if (p->head_dim % 8 || p->head_dim <= 0 || p->seq_q <= 0 || p->seq_kv <= 0)
        return fail(SFB_ERR_INVALID, "sfb_attention: bad geometry");
    // V^T carries one extra all-ones row at index head_dim (softmax denominator on the tensor
    // core), so its row count is (head_dim + 1) rounded up to 16
    const int dv = (p->head_dim + 1 + 15) / 16 * 16;
    if (p->vt_rows != dv)
        return fail(SFB_ERR_INVALID, "sfb_attention: vt_rows=%d must be (head_dim+1) rounded to 16 (%d)", p->vt_rows, dv);
    AttnArgs a{};
    a.out = p->out; a.batch = p->batch; a.heads = p->heads; a.head_dim = p->head_dim;
    a.seq_q = p->seq_q; a.seq_kv = p->seq_kv; a.q_rows = p->q_rows; a.k_rows = p->k_rows;
    a.vt_rows = p->vt_rows; a.dtype = p->dtype;
    a.scale_log2 = p->scale * 1.4426950408889634f;
    a.causal = p->causal ? 1 : 0;
    if (a.causal && (p->kv_tile != 64 || p->seq_q != p->seq_kv))
        return fail(SFB_ERR_INVALID, "sfb_attention: causal needs kv_tile 64 (head_dim <= 64) and seq_q == seq_kv");
    if (p->kv_tile == 64) {  // v2: the caller built tmap_k with a 64-row box
        switch (p->head_dim) {
            case 32: return launch_attention_v2<32, 48, 4>(p, a, stream);
            case 40: return launch_attention_v2<48, 48, 4>(p, a, stream);
            case 64: return launch_attention_v2<64, 80, 3>(p, a, stream);
            default:
                return fail(SFB_ERR_INVALID, "sfb_attention: kv_tile 64 supports head_dim 32 / 40 / 64, not %d", p->head_dim);
        }
    }
    if (p->kv_tile != 0 && p->kv_tile != 128) return fail(SFB_ERR_INVALID, "sfb_attention: kv_tile must be 0, 64 or 128");
    switch (p->head_dim) {
        case 32: return launch_attention<1, 32, 48, 2>(p, a, stream);
        case 40: return launch_attention<1, 48, 48, 2>(p, a, stream);
        case 64: return launch_attention<1, 64, 80, 2>(p, a, stream);
        case 80: return launch_attention<2, 80, 96, 2>(p, a, stream);
        case 128: return launch_attention<2, 128, 144, 1>(p, a, stream);
        case 160: return launch_attention<3, 160, 176, 1>(p, a, stream);
        default:
            return fail(SFB_ERR_INVALID, "sfb_attention: unsupported head_dim %d", p->head_dim);
    }
}
