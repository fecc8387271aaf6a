// Microbenchmark: which exp2 / max / convert mix is fastest for the attention softmax on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_pipes softmax_pipes.cu
// Each variant processes a 128-element fp32 row held in registers per thread (like one KV tile),
// 8 warps per SM (2 per SMSP, as with 2 attention CTAs per SM); prints cycles per element per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(float a, float b) {
    uint32_t r; asm("{\n\t.reg .b32 t;\n\tcvt.rn.f16x2.f32 t, %2, %1;\n\tex2.approx.f16x2 %0, t;\n\t}" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint32_t ex2_bf16x2(float a, float b) {
    uint32_t r; asm("{\n\t.reg .b32 t;\n\tcvt.rn.bf16x2.f32 t, %2, %1;\n\tex2.approx.ftz.bf16x2 %0, t;\n\t}" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
    uint32_t r; asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
// Cody-Waite + degree-3 polynomial 2^x for x <= 0 (fp32 FMA pipe only, no MUFU)
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    float fl = floorf(x);
    float f = x - fl;                       // [0,1)
    float p = fmaf(f, 0.0555054f, 0.2402265f);
    p = fmaf(p, f, 0.6931472f);
    p = fmaf(p, f, 1.0f);
    int e = (int)fl;
    return __int_as_float(__float_as_int(p) + (e << 23));
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(const float* __restrict__ in, uint32_t* __restrict__ out, long long* cyc, int iters) {
    float s[128];
    for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 128 + i) & 4095];
    float m_run = -1e30f, l = 0.f;
    uint32_t sink = 0;
    const float sl2 = 0.2280f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // v1 kernel: serial max chain, fmul, fsub, ex2.f32, add, pack
            float mx = -1e30f;
#pragma unroll
            for (int i = 0; i < 128; ++i) { s[i] = s[i] * sl2; mx = fmaxf(mx, s[i]); }
            float mn = fmaxf(m_run, mx); m_run = mn;
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
                float p0 = ex2f(s[i] - mn), p1 = ex2f(s[i + 1] - mn);
                l += p0 + p1; sink ^= pack_f16x2(p0, p1);
                s[i] = p0 + 1.0f; s[i + 1] = p1 + 1.0f;
            }
        } else if (MODE == 1) {  // v2 kernel: fmax3 chain, ffma, cvt+ex2.f16x2
            float mx = -1e30f;
#pragma unroll
            for (int i = 0; i < 128; i += 2) mx = fmax3(mx, s[i], s[i + 1]);
            float mn = fmaxf(m_run, mx * sl2); m_run = mn; float ng = -mn;
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
                uint32_t p = ex2_f16x2(fmaf(s[i], sl2, ng), fmaf(s[i + 1], sl2, ng));
                sink ^= p; s[i] += 1.0f; s[i + 1] += 1.0f;
            }
        } else if (MODE == 2) {  // 4 independent fmax3 chains, ffma, ex2.f32, pack (no sum)
            float m0 = -1e30f, m1 = m0, m2 = m0, m3 = m0;
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
                m0 = fmax3(m0, s[i], s[i + 1]); m1 = fmax3(m1, s[i + 2], s[i + 3]);
                m2 = fmax3(m2, s[i + 4], s[i + 5]); m3 = fmax3(m3, s[i + 6], s[i + 7]);
            }
            float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            float mn = fmaxf(m_run, mx * sl2); m_run = mn; float ng = -mn;
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
                float p0 = ex2f(fmaf(s[i], sl2, ng)), p1 = ex2f(fmaf(s[i + 1], sl2, ng));
                sink ^= pack_f16x2(p0, p1); s[i] += 1.0f; s[i + 1] += 1.0f;
            }
        } else if (MODE == 3) {  // like 2 but ex2.f16x2
            float m0 = -1e30f, m1 = m0, m2 = m0, m3 = m0;
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
                m0 = fmax3(m0, s[i], s[i + 1]); m1 = fmax3(m1, s[i + 2], s[i + 3]);
                m2 = fmax3(m2, s[i + 4], s[i + 5]); m3 = fmax3(m3, s[i + 6], s[i + 7]);
            }
            float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            float mn = fmaxf(m_run, mx * sl2); m_run = mn; float ng = -mn;
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
                uint32_t p = ex2_f16x2(fmaf(s[i], sl2, ng), fmaf(s[i + 1], sl2, ng));
                sink ^= p; s[i] += 1.0f; s[i + 1] += 1.0f;
            }
        } else if (MODE == 4) {  // like 2 but 1/4 of the elements use the FMA-pipe polynomial
            float m0 = -1e30f, m1 = m0, m2 = m0, m3 = m0;
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
                m0 = fmax3(m0, s[i], s[i + 1]); m1 = fmax3(m1, s[i + 2], s[i + 3]);
                m2 = fmax3(m2, s[i + 4], s[i + 5]); m3 = fmax3(m3, s[i + 6], s[i + 7]);
            }
            float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            float mn = fmaxf(m_run, mx * sl2); m_run = mn; float ng = -mn;
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
                float p0 = ex2f(fmaf(s[i], sl2, ng)), p1 = ex2f(fmaf(s[i + 1], sl2, ng));
                float p2 = ex2f(fmaf(s[i + 2], sl2, ng)), p3 = ex2_poly(fmaf(s[i + 3], sl2, ng));
                sink ^= pack_f16x2(p0, p1) ^ pack_f16x2(p2, p3);
                s[i] += 1.0f; s[i + 1] += 1.0f; s[i + 2] += 1.0f; s[i + 3] += 1.0f;
            }
        } else if (MODE == 5) {  // pure MUFU f32 rate
#pragma unroll
            for (int i = 0; i < 128; ++i) s[i] = ex2f(s[i]);
        } else if (MODE == 6) {  // pure MUFU f16x2 rate (convert outside the loop cost excluded)
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
                uint32_t v = __float_as_uint(s[i]);
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v));
                s[i] = __uint_as_float(v);
            }
        } else if (MODE == 7) {  // pure bf16x2
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
                uint32_t v = __float_as_uint(s[i]);
                asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v));
                s[i] = __uint_as_float(v);
            }
        } else if (MODE == 8) {  // 1/2 poly
            float m0 = -1e30f, m1 = m0, m2 = m0, m3 = m0;
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
                m0 = fmax3(m0, s[i], s[i + 1]); m1 = fmax3(m1, s[i + 2], s[i + 3]);
                m2 = fmax3(m2, s[i + 4], s[i + 5]); m3 = fmax3(m3, s[i + 6], s[i + 7]);
            }
            float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            float mn = fmaxf(m_run, mx * sl2); m_run = mn; float ng = -mn;
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
                float p0 = ex2f(fmaf(s[i], sl2, ng)), p1 = ex2_poly(fmaf(s[i + 1], sl2, ng));
                sink ^= pack_f16x2(p0, p1); s[i] += 1.0f; s[i + 1] += 1.0f;
            }
        }
    }
    long long t1 = clock64();
    float acc = l + m_run;
    for (int i = 0; i < 128; ++i) acc += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink ^ __float_as_uint(acc);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char* name, const float* in, uint32_t* out, long long* cyc) {
    const int iters = 200;
    bench<MODE><<<148, 256>>>(in, out, cyc, 10);
    cudaDeviceSynchronize();
    bench<MODE><<<148, 256>>>(in, out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // 8 warps/SM, each 128 elements per iteration per thread
    printf("{\"mode\": %d, \"name\": \"%s\", \"cycles_per_tile_row128\": %.1f, \"cycles_per_elem_per_warp\": %.3f, \"sm_cycles_per_128x128_tile\": %.1f}\n",
           MODE, name, (double)c / iters, (double)c / iters / 128.0, (double)c / iters / 2.0);
}

int main() {
    float* in; uint32_t* out; long long* cyc;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = -(float)(i % 97) * 0.11f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("v1: serial fmax, fmul, fsub, ex2.f32, fadd sum, pack", in, out, cyc);
    run<1>("v2: serial fmax3, ffma, cvt+ex2.f16x2", in, out, cyc);
    run<2>("4x fmax3 chains, ffma, ex2.f32, pack", in, out, cyc);
    run<3>("4x fmax3 chains, ffma, cvt+ex2.f16x2", in, out, cyc);
    run<4>("as 2 with 1/4 FMA-pipe polynomial exp2", in, out, cyc);
    run<8>("as 2 with 1/2 FMA-pipe polynomial exp2", in, out, cyc);
    run<5>("pure ex2.approx.ftz.f32", in, out, cyc);
    run<6>("pure ex2.approx.f16x2 (per 2 elems)", in, out, cyc);
    run<7>("pure ex2.approx.ftz.bf16x2 (per 2 elems)", in, out, cyc);
    return 0;
}
