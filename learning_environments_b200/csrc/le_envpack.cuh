// le_envpack.cuh — SE / RN weights repacked for the one-warp-per-lane kernels.
//
// The member's SE (envs/virtual_env.py: three MLPs over cat(one_hot(a), s)) or RN (envs/reward_env.py)
// parameter vector arrives in torch state_dict order (include/le_b200.h "parameter vectors").  The env step of
// a lane is a single-row mat-vec, so every hidden unit of every net becomes one RECORD owned by thread
// (rec % 32); records are stored as float4 planes [k][q][lane] so that a warp's load of plane (k, q) is one
// fully coalesced 512-byte request.
//
//   SE record (REC floats = 4*RQ):  W1[IN] | b1 | W2ext[SD+2] | slope | pad      (IN = AD + SD)
//       W2ext = column of the unit in the concatenated output [next_state(SD), reward, done], zero elsewhere
//   RN record:                      W1[SD] | b1 | w2 | slope | pad
//   tail (one extra group of float4): output biases  b2_state[SD], b2_reward, b2_done   (RN: b2)
#pragma once
#include "le_common.cuh"

namespace le {

template <int SD, int AD>
struct SePack {
    static constexpr int IN = AD + SD;
    static constexpr int NOUT = SD + 2;
    static constexpr int REC_RAW = IN + 1 + NOUT + 1;
    static constexpr int RQ = (REC_RAW + 3) / 4;
    static constexpr int REC = RQ * 4;
    static constexpr int OFF_B1 = IN, OFF_W2 = IN + 1, OFF_SLOPE = IN + 1 + NOUT;
    static constexpr int TAILQ = (NOUT + 3) / 4;
    __host__ __device__ static int records(int H) { return 3 * H; }
    __host__ __device__ static int K(int H) { return (3 * H + 31) / 32; }
    __host__ __device__ static int64_t pack_vec4(int H) { return (int64_t)K(H) * RQ * 32 + TAILQ; }  // float4 count
};

template <int SD>
struct RnPack {
    static constexpr int REC_RAW = SD + 3;
    static constexpr int RQ = (REC_RAW + 3) / 4;
    static constexpr int REC = RQ * 4;
    static constexpr int OFF_B1 = SD, OFF_W2 = SD + 1, OFF_SLOPE = SD + 2;
    static constexpr int TAILQ = 1;
    __host__ __device__ static int K(int H) { return (H + 31) / 32; }
    __host__ __device__ static int64_t pack_vec4(int H) { return (int64_t)K(H) * RQ * 32 + TAILQ; }
};

__host__ __device__ inline int mlp_params(int in, int H, int out) { return H * in + H + out * H + out; }

// value of float `f` of SE record `rec` from the canonical theta
template <int SD, int AD>
__device__ inline float se_record_value(const float* __restrict__ th, int H, int rec, int f, const float* slopes) {
    using P = SePack<SD, AD>;
    if (rec >= 3 * H) return 0.f;
    const int n = rec / H, j = rec % H;
    const int out_n = n == 0 ? SD : 1;
    const int base = n == 0 ? 0 : (mlp_params(P::IN, H, SD) + (n - 1) * mlp_params(P::IN, H, 1));
    const float* W1 = th + base;
    const float* b1 = W1 + H * P::IN;
    const float* W2 = b1 + H;
    if (f < P::IN) return W1[j * P::IN + f];
    if (f == P::OFF_B1) return b1[j];
    if (f < P::OFF_W2 + P::NOUT) {
        const int o = f - P::OFF_W2;
        if (n == 0) return o < SD ? W2[o * H + j] : 0.f;
        return (o == SD + (n - 1)) ? W2[j] : 0.f;
    }
    if (f == P::OFF_SLOPE) return slopes[n];
    (void)out_n;
    return 0.f;
}

template <int SD, int AD>
__global__ void pack_se_kernel(const float* __restrict__ theta, int P_env, int n_env, int H, float s0, float s1, float s2,
                               float* __restrict__ out, int64_t out_stride_f) {
    using P = SePack<SD, AD>;
    const int K = P::K(H);
    const int64_t per_env = (int64_t)K * P::RQ * 32 * 4 + P::TAILQ * 4;
    const float slopes[3] = {s0, s1, s2};
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < per_env * n_env; t += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(t / per_env);
        const int64_t r = t % per_env;
        const float* th = theta + (int64_t)e * P_env;
        float v;
        const int64_t body = (int64_t)K * P::RQ * 32 * 4;
        if (r < body) {
            // float index inside float4 plane layout [k][q][lane][4]
            const int c = (int)(r & 3);
            const int lane = (int)((r >> 2) & 31);
            const int q = (int)((r >> 7) % P::RQ);
            const int k = (int)((r >> 7) / P::RQ);
            v = se_record_value<SD, AD>(th, H, k * 32 + lane, q * 4 + c, slopes);
        } else {
            const int o = (int)(r - body);
            const int b2s = H * P::IN + H + SD * H;                                  // state_net b2
            const int b2r = mlp_params(P::IN, H, SD) + H * P::IN + H + H;            // reward_net b2
            const int b2d = b2r + mlp_params(P::IN, H, 1);                           // done_net b2
            v = o < SD ? th[b2s + o] : (o == SD ? th[b2r] : (o == SD + 1 ? th[b2d] : 0.f));
        }
        out[(int64_t)e * out_stride_f + r] = v;
    }
}

template <int SD>
__global__ void pack_rn_kernel(const float* __restrict__ theta, int P_env, int n_env, int H, float slope,
                               float* __restrict__ out, int64_t out_stride_f) {
    using P = RnPack<SD>;
    const int K = P::K(H);
    const int64_t body = (int64_t)K * P::RQ * 32 * 4;
    const int64_t per_env = body + 4;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < per_env * n_env; t += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(t / per_env);
        const int64_t r = t % per_env;
        const float* th = theta + (int64_t)e * P_env;
        float v = 0.f;
        if (r < body) {
            const int c = (int)(r & 3);
            const int lane = (int)((r >> 2) & 31);
            const int q = (int)((r >> 7) % P::RQ);
            const int k = (int)((r >> 7) / P::RQ);
            const int j = k * 32 + lane, f = q * 4 + c;
            if (j < H) {
                if (f < SD) v = th[j * SD + f];
                else if (f == P::OFF_B1) v = th[H * SD + j];
                else if (f == P::OFF_W2) v = th[H * SD + H + j];
                else if (f == P::OFF_SLOPE) v = slope;
            }
        } else if (r == body) {
            v = th[H * SD + H + H];
        }
        out[(int64_t)e * out_stride_f + r] = v;
    }
}

// ---- VirtualEnv.step for one row held replicated in registers (envs/virtual_env.py:43-54) ---------------
// SH: `pack` points to a shared-memory copy of the pack (multi-warp lanes stage it once per lane with a TMA bulk copy): plain loads
// instead of ld.global.nc
template <bool SH>
__device__ __forceinline__ float4 pack_ld4(const float4* p) { if constexpr (SH) return *p; else return __ldg(p); }
template <bool SH>
__device__ __forceinline__ float pack_ld1(const float* p) { if constexpr (SH) return *p; else return __ldg(p); }

template <int SD, int AD, bool SH = false>
__device__ __forceinline__ void se_step_row(const float4* __restrict__ pack, int H, bool is_tanh, const float (&s)[SD],
                                            int action, int lane, float (&ns)[SD], float& reward, float& done) {
    using P = SePack<SD, AD>;
    const int K = P::K(H);
    float acc[P::NOUT];
#pragma unroll
    for (int o = 0; o < P::NOUT; ++o) acc[o] = 0.f;
    for (int k = 0; k < K; ++k) {
        float rec[P::REC];
#pragma unroll
        for (int q = 0; q < P::RQ; ++q) {
            const float4 v = pack_ld4<SH>(pack + ((int64_t)k * P::RQ + q) * 32 + lane);
            rec[4 * q + 0] = v.x; rec[4 * q + 1] = v.y; rec[4 * q + 2] = v.z; rec[4 * q + 3] = v.w;
        }
        // input = cat(one_hot(action), state): the one-hot picks one action column of W1
        float wa = rec[0];
#pragma unroll
        for (int a = 1; a < AD; ++a) wa = (action == a) ? rec[a] : wa;
        float z = rec[P::OFF_B1] + wa;
#pragma unroll
        for (int i = 0; i < SD; ++i) z = fmaf(rec[AD + i], s[i], z);
        const float h = env_act(z, is_tanh, rec[P::OFF_SLOPE]);
#pragma unroll
        for (int o = 0; o < P::NOUT; ++o) acc[o] = fmaf(h, rec[P::OFF_W2 + o], acc[o]);
    }
    const float* tail = reinterpret_cast<const float*>(pack + (int64_t)K * P::RQ * 32);
#pragma unroll
    for (int o = 0; o < P::NOUT; ++o) acc[o] = warp_allreduce_sum(acc[o]) + pack_ld1<SH>(tail + o);
#pragma unroll
    for (int i = 0; i < SD; ++i) ns[i] = acc[i];
    reward = acc[SD];
    done = acc[SD + 1];
}

// ---- the RN potential Phi(s), Phi(s') (envs/reward_env.py:84-110) ---------------------------------------
template <int SD, bool SH = false>
__device__ __forceinline__ void rn_phi2(const float4* __restrict__ pack, int H, bool is_tanh, const float (&s)[SD],
                                        const float (&s2)[SD], int lane, float& phi_s, float& phi_s2) {
    using P = RnPack<SD>;
    const int K = P::K(H);
    float a0 = 0.f, a1 = 0.f;
    for (int k = 0; k < K; ++k) {
        float rec[P::REC];
#pragma unroll
        for (int q = 0; q < P::RQ; ++q) {
            const float4 v = pack_ld4<SH>(pack + ((int64_t)k * P::RQ + q) * 32 + lane);
            rec[4 * q + 0] = v.x; rec[4 * q + 1] = v.y; rec[4 * q + 2] = v.z; rec[4 * q + 3] = v.w;
        }
        float z0 = rec[P::OFF_B1], z1 = rec[P::OFF_B1];
#pragma unroll
        for (int i = 0; i < SD; ++i) { z0 = fmaf(rec[i], s[i], z0); z1 = fmaf(rec[i], s2[i], z1); }
        a0 = fmaf(env_act(z0, is_tanh, rec[P::OFF_SLOPE]), rec[P::OFF_W2], a0);
        a1 = fmaf(env_act(z1, is_tanh, rec[P::OFF_SLOPE]), rec[P::OFF_W2], a1);
    }
    const float b2 = pack_ld1<SH>(reinterpret_cast<const float*>(pack + (int64_t)K * P::RQ * 32));
    phi_s = warp_allreduce_sum(a0) + b2;
    phi_s2 = warp_allreduce_sum(a1) + b2;
}

// RewardEnv._calc_reward, state-only types (envs/reward_env.py:81-110)
__device__ __forceinline__ float rn_combine(int rn_type, float gamma, float real_reward, float phi_s, float phi_s2) {
    switch (rn_type) {
        case 1: return gamma * phi_s2 - phi_s;
        case 2: return (real_reward + gamma * phi_s2) - phi_s;
        case 5: return phi_s2;
        case 6: return real_reward + phi_s2;
        default: return real_reward;
    }
}

}  // namespace le
