// Kernels of the SVD (UNetSpatioTemporalConditionModel) temporal path that are not GEMMs:
//
//   temporal_attention_kernel   self-attention ACROSS FRAMES: for every (batch, pixel, head) a
//                               sequence of F <= 32 frame tokens (SVD-XT: 25) of head_dim 64.
//                               Reads the fused QKV projection [rows, 3C] in the spatial row order
//                               m = (b * F + f) * S + p -- the diffusers reshape / permute to
//                               [B * S, F, C] (TemporalBasicTransformerBlock.forward) is only an
//                               index map here, no tensor is transposed.  92 160 tiny attentions of
//                               25 x 25 x 64 at the 72 x 128 level: CUDA-core FMA work
//                               (2 * F * F * D MACs per head), one warp per (b, p, head), the K / V
//                               rows of the sequence staged in shared memory.
//   row_broadcast_add_kernel    x_out[m, :] = x_in[m, :] + vec[idx(m), :]  (+ per-row LayerNorm
//                               statistics of the stored row).  Three uses: the frame position
//                               embedding (idx = frame), and cross-attention over a context of ONE
//                               token, where softmax over a single key is 1 and the layer reduces
//                               to adding to_out(to_v(context)) -- spatial blocks (idx = image)
//                               and temporal blocks (idx = diffusers' [H*W, B]-interleaved
//                               time_context, kept as published).
//   alpha_blend_kernel          AlphaBlender: out = alpha * x_spatial + (1 - alpha) * x_temporal
//                               (+ row statistics), alpha = sigmoid(mix_factor) read from memory.
//
// The reference has no kernel for any of these: it traces the diffusers module and leaves them to
// aten (/root/reference/src/sfast/compilers/diffusion_pipeline_compiler.py:101-103).
#include "common.cuh"
#include "host.h"

namespace sfb {

// ---------------------------------------------------------------------------------------
// temporal attention
// ---------------------------------------------------------------------------------------
constexpr int kTaWarps = 4;       // (b, p, head) sequences per CTA (K/V staging: 36 KB static smem)
constexpr int kTaMaxF = 32;

struct TemporalAttnArgs {
    const uint16_t* qkv;  // [B * F * S, 3 * C]: q | k | v, head h at columns h * D
    uint16_t* out;        // [B * F * S, ldo]
    int batch, frames, seq /* S = pixels per frame */, heads, ldq, ldo, dtype;
    float scale;
};

template <int D>
__global__ void __launch_bounds__(kTaWarps * 32) temporal_attention_kernel(const TemporalAttnArgs a) {
    // K and V of each warp's sequence: [frames][D] halves, padded to avoid bank conflicts on the
    // per-lane row reads of the P V step
    __shared__ __align__(16) uint16_t sK[kTaWarps][kTaMaxF][D + 8];
    __shared__ __align__(16) uint16_t sV[kTaWarps][kTaMaxF][D + 8];
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nseq = (long long)a.batch * a.seq * a.heads;
    const long long sid = (long long)blockIdx.x * kTaWarps + warp;  // ((b * S + p) * heads + h)
    if (sid >= nseq) return;
    const int h = (int)(sid % a.heads);
    const long long bp = sid / a.heads;
    const int p = (int)(bp % a.seq), b = (int)(bp / a.seq);
    const int C = a.heads * D;
    const size_t row0 = ((size_t)b * a.frames) * a.seq + p;  // row of frame 0; frames are seq rows apart
    // ---- stage K, V: lane l copies 16-byte chunks (D / 8 chunks per row and operand)
    constexpr int kChunks = D / 8;
    for (int i = lane; i < a.frames * kChunks; i += 32) {
        const int f = i / kChunks, c = i - f * kChunks;
        const uint16_t* src = a.qkv + (row0 + (size_t)f * a.seq) * a.ldq + h * D + c * 8;
        *reinterpret_cast<uint4*>(&sK[warp][f][c * 8]) = *reinterpret_cast<const uint4*>(src + C);
        *reinterpret_cast<uint4*>(&sV[warp][f][c * 8]) = *reinterpret_cast<const uint4*>(src + 2 * C);
    }
    __syncwarp();
    if (lane >= a.frames) return;  // lane = query frame
    // ---- q row in registers (fp32)
    float q[D];
    {
        const uint16_t* src = a.qkv + (row0 + (size_t)lane * a.seq) * a.ldq + h * D;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + c * 8);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f2 = unpack2(w[i], a.dtype);
                q[c * 8 + 2 * i] = f2.x * a.scale;
                q[c * 8 + 2 * i + 1] = f2.y * a.scale;
            }
        }
    }
    // ---- scores (all lanes read the same K row: shared-memory broadcast), online softmax not
    // needed: F <= 32 scores live in registers
    float s[kTaMaxF];
    float mx = -INFINITY;
    // (loops over keys are fully unrolled with a uniform `j < frames` guard so that s[] stays in
    // registers; a dynamically indexed array would live in local memory)
#pragma unroll
    for (int j = 0; j < kTaMaxF; ++j) {
        s[j] = -INFINITY;
        if (j < a.frames) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const uint4 v = *reinterpret_cast<const uint4*>(&sK[warp][j][c * 8]);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f2 = unpack2(w[i], a.dtype);
                    acc += q[c * 8 + 2 * i] * f2.x + q[c * 8 + 2 * i + 1] * f2.y;
                }
            }
            s[j] = acc;
            mx = fmaxf(mx, acc);
        }
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < kTaMaxF; ++j) {
        s[j] = (j < a.frames) ? __expf(s[j] - mx) : 0.f;
        den += s[j];
    }
    const float inv = 1.0f / den;
    // ---- O = P V (q registers reused as the output accumulator)
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = 0.f;
#pragma unroll
    for (int j = 0; j < kTaMaxF; ++j) {
        if (j < a.frames) {
            const float pj = s[j] * inv;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const uint4 v = *reinterpret_cast<const uint4*>(&sV[warp][j][c * 8]);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f2 = unpack2(w[i], a.dtype);
                    q[c * 8 + 2 * i] += pj * f2.x;
                    q[c * 8 + 2 * i + 1] += pj * f2.y;
                }
            }
        }
    }
    uint16_t* dst = a.out + (row0 + (size_t)lane * a.seq) * a.ldo + h * D;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        uint4 o;
        o.x = pack2(q[c * 8 + 0], q[c * 8 + 1], a.dtype);
        o.y = pack2(q[c * 8 + 2], q[c * 8 + 3], a.dtype);
        o.z = pack2(q[c * 8 + 4], q[c * 8 + 5], a.dtype);
        o.w = pack2(q[c * 8 + 6], q[c * 8 + 7], a.dtype);
        *reinterpret_cast<uint4*>(dst + c * 8) = o;
    }
}

// ---------------------------------------------------------------------------------------
// row-wise broadcast add / AlphaBlender (one warp per row, 16-byte accesses)
// ---------------------------------------------------------------------------------------
struct RowOpArgs {
    const uint16_t* x;   // [rows, ldx]
    const uint16_t* x2;  // blend: the temporal stream [rows, ldx2]; else unused
    const uint16_t* vec; // broadcast add: [nvec, ldv] 16-bit
    uint16_t* y;         // [rows, ldy]
    float* rowstats;     // [rows, rs_slots, 2] fp32: slot 0 = (sum, sum of squares) of the stored row, WRITTEN; or null
    int rs_slots;
    const float* mix;    // blend: pointer to mix_factor (alpha = sigmoid(*mix))
    int rows, c, ldx, ldx2, ldv, ldy, dtype;
    int mode, div, mod, frames, seq, batch;  // index mode (sfb_row_index_mode) and its parameters
};

__device__ __forceinline__ int row_vec_index(const RowOpArgs& a, int m) {
    if (a.mode == SFB_ROW_IDX_DIV_MOD) return (m / a.div) % a.mod;
    // SFB_ROW_IDX_TEMPORAL_CTX: token (b, p) of the temporal block is sequence j = b * S + p; the
    // diffusers time_context tensor is laid out [S, B] (first frame's context of video j % B), so
    // sequence j reads the context row of video j % B -- kept as published.  vec holds one row per
    // (video, frame): the first frame of video v is row v * F.
    const int b = m / (a.frames * a.seq), p = m % a.seq;
    return (int)(((long long)b * a.seq + p) % a.batch) * a.frames;
}

template <bool kBlend>
__global__ void __launch_bounds__(256) row_op_kernel(const RowOpArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= a.rows) return;
    const uint16_t* xr = a.x + (size_t)m * a.ldx;
    const uint16_t* br = kBlend ? a.x2 + (size_t)m * a.ldx2 : a.vec + (size_t)row_vec_index(a, m) * a.ldv;
    uint16_t* yr = a.y + (size_t)m * a.ldy;
    float alpha = 1.f, beta = 1.f;
    if (kBlend) {
        alpha = 1.0f / (1.0f + __expf(-__ldg(a.mix)));
        beta = 1.0f - alpha;
    }
    float s = 0.f, ss = 0.f;
    for (int v = lane; v < a.c / 8; v += 32) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xr + v * 8);
        const uint4 bv = *reinterpret_cast<const uint4*>(br + v * 8);
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 xf = unpack2(xw[i], a.dtype), bf = unpack2(bw[i], a.dtype);
            o[i] = pack2(alpha * xf.x + beta * bf.x, alpha * xf.y + beta * bf.y, a.dtype);
            const float2 r = unpack2(o[i], a.dtype);  // statistics of the values as stored
            s += r.x + r.y;
            ss += r.x * r.x + r.y * r.y;
        }
        *reinterpret_cast<uint4*>(yr + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (a.rowstats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
        }
        if (lane == 0) *reinterpret_cast<float2*>(a.rowstats + 2 * (size_t)m * a.rs_slots) = make_float2(s, ss);
    }
}

}  // namespace sfb

using namespace sfb;

extern "C" int sfb_temporal_attention(const sfb_temporal_attn_params* p, sfb_stream_t stream) {
    if (!p || !p->qkv || !p->out) return fail(SFB_ERR_INVALID, "temporal_attention: null argument");
    if (p->frames <= 0 || p->frames > kTaMaxF || p->head_dim != 64 || p->heads <= 0 || p->batch <= 0 ||
        p->seq <= 0 || p->ld_qkv % 8 || p->ld_out % 8 || p->ld_qkv < 3 * p->heads * p->head_dim)
        return fail(SFB_ERR_INVALID, "temporal_attention: needs frames <= 32, head_dim 64 (got frames=%d head_dim=%d)",
                    p->frames, p->head_dim);
    TemporalAttnArgs a{reinterpret_cast<const uint16_t*>(p->qkv), reinterpret_cast<uint16_t*>(p->out),
                       p->batch, p->frames, p->seq, p->heads, p->ld_qkv, p->ld_out, p->dtype,
                       p->scale};
    const long long nseq = (long long)p->batch * p->seq * p->heads;
    const long long blocks = (nseq + kTaWarps - 1) / kTaWarps;
    if (blocks > 0x7fffffffLL) return fail(SFB_ERR_INVALID, "temporal_attention: too many sequences");
    cudaError_t err = launch_pdl(temporal_attention_kernel<64>, dim3((unsigned)blocks), dim3(kTaWarps * 32), 0,
                                 static_cast<cudaStream_t>(stream), a);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "temporal_attention: %s", cudaGetErrorString(err));
    return check_launch("sfb_temporal_attention");
}

static int make_row_args(const sfb_row_op_params* p, RowOpArgs& a, bool blend) {
    if (!p || !p->x || !p->y || (blend ? (!p->x2 || !p->mix_factor) : !p->vec))
        return fail(SFB_ERR_INVALID, "row op: null argument");
    if (p->rows <= 0 || p->c <= 0 || p->c % 8 || p->ldx % 8 || p->ldy % 8 || (blend ? p->ldx2 % 8 : p->ldv % 8))
        return fail(SFB_ERR_INVALID, "row op: c / pitches must be multiples of 8");
    if (!blend && (p->mode == SFB_ROW_IDX_DIV_MOD ? (p->div <= 0 || p->mod <= 0)
                                                  : (p->mode != SFB_ROW_IDX_TEMPORAL_CTX || p->frames <= 0 ||
                                                     p->seq <= 0 || p->batch <= 0)))
        return fail(SFB_ERR_INVALID, "row op: bad index mode / geometry");
    a.x = reinterpret_cast<const uint16_t*>(p->x); a.x2 = reinterpret_cast<const uint16_t*>(p->x2);
    a.vec = reinterpret_cast<const uint16_t*>(p->vec); a.y = reinterpret_cast<uint16_t*>(p->y);
    a.rowstats = p->rowstats_out; a.mix = p->mix_factor;
    a.rs_slots = p->rowstats_slots > 0 ? p->rowstats_slots : 1;
    a.rows = p->rows; a.c = p->c; a.ldx = p->ldx; a.ldx2 = p->ldx2; a.ldv = p->ldv; a.ldy = p->ldy;
    a.dtype = p->dtype; a.mode = p->mode; a.div = p->div; a.mod = p->mod;
    a.frames = p->frames; a.seq = p->seq; a.batch = p->batch;
    return SFB_OK;
}

extern "C" int sfb_row_broadcast_add(const sfb_row_op_params* p, sfb_stream_t stream) {
    RowOpArgs a{};
    int rc = make_row_args(p, a, false);
    if (rc) return rc;
    cudaError_t err = launch_pdl(row_op_kernel<false>, dim3((p->rows + 7) / 8), dim3(256), 0,
                                 static_cast<cudaStream_t>(stream), a);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "row_broadcast_add: %s", cudaGetErrorString(err));
    return check_launch("sfb_row_broadcast_add");
}

extern "C" int sfb_alpha_blend(const sfb_row_op_params* p, sfb_stream_t stream) {
    RowOpArgs a{};
    int rc = make_row_args(p, a, true);
    if (rc) return rc;
    cudaError_t err = launch_pdl(row_op_kernel<true>, dim3((p->rows + 7) / 8), dim3(256), 0,
                                 static_cast<cudaStream_t>(stream), a);
    if (err != cudaSuccess) return fail(SFB_ERR_CUDA, "alpha_blend: %s", cudaGetErrorString(err));
    return check_launch("sfb_alpha_blend");
}
