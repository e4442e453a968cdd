// divcheck.cu — compares the branch-free fast-path sequences of div.rn.f32 / sqrt.rn.f32 used by LaneCore::adam_polyak
// with the IEEE intrinsics over random operands inside the guarded ranges.  nvcc -arch=sm_100a -o divcheck divcheck.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../learning_environments_b200/csrc/le_lane.cuh"

__device__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ float rnd_mag(uint32_t h, int emin, int emax) {   // random sign-less float with exponent in [emin, emax]
    const int e = emin + (int)(hash(h) % (uint32_t)(emax - emin + 1));
    const uint32_t mant = hash(h ^ 0x9e3779b9U) & 0x7fffffU;
    return __uint_as_float(((uint32_t)(e + 127) << 23) | mant);
}
__global__ void check(unsigned long long* bad_div, unsigned long long* bad_sqrt, int n_per_thread) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bd = 0, bs = 0;
    for (int i = 0; i < n_per_thread; ++i) {
        const uint32_t h = tid * 7919u + i * 104729u;
        // Adam's operand ranges: numerator |lr/bc1 * m| in [1e-20, 6e29] (guarded), denominator sqrt(v)/sqrt(bc2) + eps in [1e-8, 1e15];
        // every third sample: the first division sqrt(v) / sqrt(bc2) with sqrt(bc2) in [0.03, 1]
        float a = rnd_mag(h, -66, 99), b = rnd_mag(h + 1, -27, 50);
        if (i % 3 == 0) { a = rnd_mag(h, -50, 50); b = rnd_mag(h + 1, -5, 0); }
        if (hash(h + 2) & 1) a = -a;
        if (le::div_rn_core(a, b) != __fdiv_rn(a, b)) bd++;
        const float x = rnd_mag(h + 3, -99, 99);
        if (le::sqrt_rn_core(x) != __fsqrt_rn(x)) bs++;
    }
    atomicAdd(bad_div, bd);
    atomicAdd(bad_sqrt, bs);
}
int main() {
    unsigned long long *d, h[2] = {0, 0};
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    check<<<1024, 256>>>(d, d + 1, 256);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("div_rn_core mismatches vs __fdiv_rn: %llu of %d; sqrt_rn_core mismatches vs __fsqrt_rn: %llu\n", h[0], 1024 * 256 * 256, h[1]);
    return (h[0] || h[1]) ? 1 : 0;
}
