// ubench2 — register-operand throughput of the packed fp32x2 instructions as the fused kernel really issues them (B200).
// The round-1 microbenchmark fed FFMA2 uniform / immediate operands; the TD update reads THREE distinct register operands
// per FFMA2 (weight pair x staged scalar + accumulator pair).  This measures cycles per warp instruction per SMSP for each
// operand form, alone and interleaved with MUFU / FSEL / LDS, at 1..4 warps per scheduler (in-kernel clock64).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(512) k(const float* __restrict__ in, float* out, long long* clk, int iters) {
    __shared__ float4 sm[64];
    float2 acc[16], w[8];
    float s[8];
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(in[t + i], in[t + 32 + i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = make_float2(in[t + 64 + i], in[t + 96 + i]); s[i] = in[t + 128 + i]; }
    if (t < 64) sm[t] = make_float4(in[t], in[t + 1], in[t + 2], in[t + 3]);
    __syncthreads();
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = in[t + 160 + i];
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int j = (i + r) & 7, q = (i * 3 + r) & 7;
                if (MODE == 0) acc[i] = __ffma2_rn(w[j], make_float2(s[q], s[q]), acc[i]);                   // pair x scalar + pair
                if (MODE == 1) acc[i] = __ffma2_rn(w[j], w[q], acc[i]);                                      // pair x pair + pair
                if (MODE == 2) acc[i] = __ffma2_rn(acc[i], make_float2(0.999f, 0.999f), make_float2(1e-3f, 1e-3f));   // immediates
                if (MODE == 3) acc[i] = __fmul2_rn(acc[i], w[j]);                                            // FMUL2 pair x pair
                if (MODE == 4) acc[i] = __fadd2_rn(acc[i], w[j]);                                            // FADD2 pair + pair
                if (MODE == 5) { acc[i].x = fmaf(w[j].x, s[q], acc[i].x); acc[i].y = fmaf(w[j].y, s[q], acc[i].y); }   // 2 scalar FFMA
                if (MODE == 6) { acc[i] = __ffma2_rn(w[j], make_float2(s[q], s[q]), acc[i]); if ((i & 1) == 0) m[i >> 1] = ex2a(m[i >> 1]); }   // FFMA2 : EX2 = 2 : 1
                if (MODE == 7) { acc[i] = __ffma2_rn(w[j], make_float2(s[q], s[q]), acc[i]); if ((i & 3) == 0) m[i >> 2] = (s[q] > m[i >> 2]) ? s[j] : m[i >> 2]; }   // + FSETP/FSEL 4:1
                if (MODE == 8) { acc[i] = __ffma2_rn(w[j], make_float2(s[q], s[q]), acc[i]); if ((i & 7) == 0) { const float4 v = sm[(i + r + it) & 63]; m[0] += v.x; } }   // + LDS.128 8:1
                if (MODE == 9) m[i & 7] = ex2a(m[i & 7]);                                                   // MUFU.EX2 alone
                if (MODE == 10) { acc[i] = __fmul2_rn(acc[i], w[j]); if ((i & 1) == 0) m[i >> 1] = ex2a(m[i >> 1]); }      // FMUL2 : EX2 = 2 : 1
                if (MODE == 11) acc[i] = __ffma2_rn(w[j], make_float2(s[q], s[q]), acc[(i + 1) & 15]);       // accumulator != destination
                if (MODE == 12) { acc[i] = __ffma2_rn(w[j], make_float2(s[q], s[q]), acc[i]); if ((i & 3) == 0) m[i >> 2] = ex2a(m[i >> 2]); }   // FFMA2 : EX2 = 4 : 1
                if (MODE == 13) { acc[i] = __ffma2_rn(w[j], make_float2(s[q], s[q]), acc[i]); if ((i & 7) == 0) m[i >> 3] = ex2a(m[i >> 3]); }   // FFMA2 : EX2 = 8 : 1
                if (MODE == 14) { acc[i] = __fmul2_rn(acc[i], w[j]); if ((i & 3) == 0) m[i >> 2] = ex2a(m[i >> 2]); }      // FMUL2 : EX2 = 4 : 1
                if (MODE == 15) { acc[i] = __fmul2_rn(acc[i], w[j]); if ((i & 7) == 0) m[i >> 3] = ex2a(m[i >> 3]); }      // FMUL2 : EX2 = 8 : 1
                if (MODE == 16) acc[i & 1] = __ffma2_rn(w[j], make_float2(s[q], s[q]), acc[i & 1]);         // two dependent chains
                if (MODE == 17) acc[i] = __ffma2_rn(w[r], make_float2(s[q], s[q]), acc[i]);                 // weight-stationary: 16 instrs share w
            }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += acc[i].x + acc[i].y;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += m[i];
    if (r == 123.456f) out[0] = r;
    if ((t & 31) == 0) clk[blockIdx.x * 16 + (t >> 5)] = t1 - t0;
}

template <int MODE>
void run(const float* in, float* out, long long* clk, int sms, const char* name, double instr_per_body) {
    const int iters = 512;
    printf("%-34s", name);
    for (int wps = 1; wps <= 4; ++wps) {
        const int threads = wps * 4 * 32;
        k<MODE><<<sms, threads>>>(in, out, clk, iters);
        cudaDeviceSynchronize();
        k<MODE><<<sms, threads>>>(in, out, clk, iters);
        cudaDeviceSynchronize();
        long long h[16 * 256];
        cudaMemcpy(h, clk, sizeof(long long) * 16 * sms, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (int b = 0; b < sms; ++b) for (int w = 0; w < wps * 4; ++w) mx = h[b * 16 + w] > mx ? (double)h[b * 16 + w] : mx;
        const double n_instr = (double)iters * 64.0 * instr_per_body;      // per warp
        printf("  %dw: %5.2f", wps, mx / (n_instr * wps));               // cycles per warp instruction per SMSP
    }
    printf("   clk/instr/SMSP\n");
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float *in, *out; long long* clk;
    CK(cudaMalloc(&in, 4096 * 4)); CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&clk, sizeof(long long) * 16 * 256));
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = 0.5f + 1e-4f * (i % 97);
    CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    run<0>(in, out, clk, sms, "FFMA2 pair*scalar+pair", 1);
    run<1>(in, out, clk, sms, "FFMA2 pair*pair+pair", 1);
    run<2>(in, out, clk, sms, "FFMA2 pair*imm+imm", 1);
    run<11>(in, out, clk, sms, "FFMA2 pair*scalar+pair (d != c)", 1);
    run<3>(in, out, clk, sms, "FMUL2 pair*pair", 1);
    run<4>(in, out, clk, sms, "FADD2 pair+pair", 1);
    run<5>(in, out, clk, sms, "FFMA reg*reg+reg (x2)", 2);
    run<9>(in, out, clk, sms, "MUFU.EX2", 1);
    run<6>(in, out, clk, sms, "FFMA2 + EX2 2:1", 1.5);
    run<10>(in, out, clk, sms, "FMUL2 + EX2 2:1", 1.5);
    run<7>(in, out, clk, sms, "FFMA2 + FSETP/FSEL 4:1 (x2)", 1.5);
    run<8>(in, out, clk, sms, "FFMA2 + LDS.128+FADD 8:1", 1.25);
    run<12>(in, out, clk, sms, "FFMA2 + EX2 4:1", 1.25);
    run<13>(in, out, clk, sms, "FFMA2 + EX2 8:1", 1.125);
    run<14>(in, out, clk, sms, "FMUL2 + EX2 4:1", 1.25);
    run<15>(in, out, clk, sms, "FMUL2 + EX2 8:1", 1.125);
    run<16>(in, out, clk, sms, "FFMA2 two dependent chains", 1);
    run<17>(in, out, clk, sms, "FFMA2 weight-stationary (reuse)", 1);
    CK(cudaDeviceSynchronize());
    return 0;
}
