// le_common.cuh — device helpers shared by every kernel of the hot path (sm_100a).
//   * Philox4x32-10 (same streams as oracle/philox.py; Random123 constants)
//   * activations: accurate rational tanh on the FMA pipe (1 MUFU), leaky family with runtime slope
//   * warp reductions (butterfly all-reduce, recursive-halving row reduce)
//   * gym 0.17.3 CartPole / Acrobot dynamics in fp64 without FMA contraction
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/le_b200.h"

#define LE_FULL_MASK 0xffffffffu

// Philox stream purposes (oracle/philox.py)
#define LE_P_ACT 1u
#define LE_P_SAMPLE 2u
#define LE_P_RESET_TRAIN 3u
#define LE_P_RESET_TEST 4u
#define LE_P_QINIT 5u
#define LE_P_NOISE 6u

namespace le {

// ---------------------------------------------------------------------------------------------------
struct u32x4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return u32x4{c0, c1, c2, c3};
}

__device__ __forceinline__ uint32_t pick(const u32x4& w, int k) {
    return k == 0 ? w.x : (k == 1 ? w.y : (k == 2 ? w.z : w.w));
}

// ---------------------------------------------------------------------------------------------------
// Activations.  ACT template ids used by the Q-net kernels:
constexpr int QACT_TANH = 0;   // nn.Tanh
constexpr int QACT_LEAKY = 1;  // nn.ReLU (slope 0) / nn.LeakyReLU (slope 0.01): z > 0 ? z : slope*z

// tanh as a 13/6 rational minimax on the FMA pipe + one MUFU.RCP: max relative error 4e-7 (6 ulp) against
// fp64 tanh over the whole range — inside the 1e-5 parity budget.  Used by the SE / RN mat-vecs (env_act); the Q-net
// kernels use tanh_pair / tanh_one of le_lane.cuh (1 - 2 / (1 + 2^(2x log2 e)) on MUFU.EX2 / MUFU.RCP, abs error
// <= 2.4e-7).  tanh.approx.f32 (MUFU.TANH, 2^-11) is NOT accurate enough.
__device__ __forceinline__ float tanh_rational(float x) {
    x = fminf(fmaxf(x, -7.90531110763549805f), 7.90531110763549805f);
    const float x2 = x * x;
    float p = -2.76076847742355e-16f;
    p = fmaf(p, x2, 2.00018790482477e-13f);
    p = fmaf(p, x2, -8.60467152213735e-11f);
    p = fmaf(p, x2, 5.12229709037114e-08f);
    p = fmaf(p, x2, 1.48572235717979e-05f);
    p = fmaf(p, x2, 6.37261928875436e-04f);
    p = fmaf(p, x2, 4.89352455891786e-03f);
    p = p * x;
    float q = 1.19825839466702e-06f;
    q = fmaf(q, x2, 1.18534705686654e-04f);
    q = fmaf(q, x2, 2.26843463243900e-03f);
    q = fmaf(q, x2, 4.89352518554385e-03f);
    return __fdividef(p, q);
}

template <int ACT>
__device__ __forceinline__ float q_act(float z, float slope) {
    if (ACT == QACT_TANH) return tanh_rational(z);
    return z > 0.f ? z : slope * z;
}
// derivative from the activation OUTPUT h (sign(h) == sign(z) for the leaky family; relu: h > 0 <=> z > 0)
template <int ACT>
__device__ __forceinline__ float q_act_grad(float h, float slope) {
    if (ACT == QACT_TANH) return fmaf(-h, h, 1.f);
    return h > 0.f ? 1.f : slope;
}

// runtime-selected activation of the SE / RN nets (forward only): tanh flag or leaky-family slope
// (relu 0, leakyrelu 0.01, prelu a, identity 1).
__device__ __forceinline__ float env_act(float z, bool is_tanh, float slope) {
    return is_tanh ? tanh_rational(z) : (z > 0.f ? z : slope * z);
}

__host__ __device__ inline float act_slope(int act, float prelu_slope) {
    switch (act) {
        case LE_ACT_RELU: return 0.f;
        case LE_ACT_LEAKYRELU: return 0.01f;
        case LE_ACT_PRELU: return prelu_slope;
        case LE_ACT_IDENTITY: return 1.f;
        default: return 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------
// Warp reductions
__device__ __forceinline__ float warp_allreduce_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(LE_FULL_MASK, v, m);
    return v;  // bit-identical on all lanes (IEEE add is commutative)
}
__device__ __forceinline__ double warp_allreduce_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(LE_FULL_MASK, v, m);
    return v;
}

// Recursive-halving reduction of R per-row partials across the 32 lanes (R in {2,4,8,16,32}).
// Returns on lane L the total of row  (L / (32/R)) ; costs R-1 + log2(32/R) shuffles instead of 5R.
template <int R>
__device__ __forceinline__ float warp_reduce_rows(float (&v)[R], int lane) {
    int n = R;
#pragma unroll
    for (int m = 16; m >= 32 / R && n > 1; m >>= 1) {
        const bool hi = (lane & m) != 0;
        const int half = n >> 1;
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            if (k < half) {
                const float send = hi ? v[k] : v[k + half];
                const float keep = hi ? v[k + half] : v[k];
                v[k] = keep + __shfl_xor_sync(LE_FULL_MASK, send, m);
            }
        }
        n = half;
    }
    float t = v[0];
#pragma unroll
    for (int m = (32 / R) >> 1; m > 0; m >>= 1) t += __shfl_xor_sync(LE_FULL_MASK, t, m);
    return t;
}

// ---------------------------------------------------------------------------------------------------
// gym 0.17.3 classic_control, fp64, operation order of the python sources; __d*_rn intrinsics forbid FMA
// contraction so the result differs from the CPU restatement only through sin/cos (<= 1-2 ulp fp64).
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ void cartpole_step(double (&st)[4], int action, float& reward, bool& done) {
    const double gravity = 9.8, masspole = 0.1, total_mass = 0.1 + 1.0, length = 0.5, polemass_length = 0.1 * 0.5;
    const double force_mag = 10.0, tau = 0.02;
    const double theta_thr = 12 * 2 * 3.141592653589793 / 360, x_thr = 2.4;
    double x = st[0], x_dot = st[1], theta = st[2], theta_dot = st[3];
    const double force = action == 1 ? force_mag : -force_mag;
    double sintheta, costheta;
    sincos(theta, &sintheta, &costheta);
    const double temp = ddiv(dadd(force, dmul(dmul(polemass_length, dmul(theta_dot, theta_dot)), sintheta)), total_mass);
    const double thetaacc =
        ddiv(dsub(dmul(gravity, sintheta), dmul(costheta, temp)),
             dmul(length, dsub(4.0 / 3.0, ddiv(dmul(masspole, dmul(costheta, costheta)), total_mass))));
    const double xacc = dsub(temp, ddiv(dmul(dmul(polemass_length, thetaacc), costheta), total_mass));
    x = dadd(x, dmul(tau, x_dot));
    x_dot = dadd(x_dot, dmul(tau, xacc));
    theta = dadd(theta, dmul(tau, theta_dot));
    theta_dot = dadd(theta_dot, dmul(tau, thetaacc));
    st[0] = x; st[1] = x_dot; st[2] = theta; st[3] = theta_dot;
    done = (x < -x_thr) || (x > x_thr) || (theta < -theta_thr) || (theta > theta_thr);
    reward = 1.0f;
}

__device__ __forceinline__ void acrobot_dsdt(const double (&s)[5], double (&d)[5]) {
    const double g = 9.8, pi = 3.141592653589793;
    // m1 = m2 = l1 = 1, lc1 = lc2 = 0.5, I1 = I2 = 1: products with these constants are exact
    const double a = s[4], theta1 = s[0], theta2 = s[1], dtheta1 = s[2], dtheta2 = s[3];
    double s2, c2;
    sincos(theta2, &s2, &c2);
    // d1 = m1*lc1**2 + m2*(l1**2 + lc2**2 + 2*l1*lc2*cos(theta2)) + I1 + I2
    const double d1 = dadd(dadd(dadd(0.25, dadd(dadd(1.0, 0.25), dmul(1.0, c2))), 1.0), 1.0);
    // d2 = m2*(lc2**2 + l1*lc2*cos(theta2)) + I2
    const double d2 = dadd(dadd(0.25, dmul(0.5, c2)), 1.0);
    // phi2 = m2*lc2*g*cos(theta1 + theta2 - pi/2)
    const double phi2 = dmul(dmul(0.5, g), cos(dsub(dadd(theta1, theta2), pi / 2.)));
    // phi1 = -m2*l1*lc2*dtheta2**2*sin(theta2) - 2*m2*l1*lc2*dtheta2*dtheta1*sin(theta2)
    //        + (m1*lc1 + m2*l1)*g*cos(theta1 - pi/2) + phi2
    const double t1 = dmul(dmul(-0.5, dmul(dtheta2, dtheta2)), s2);
    const double t2 = dmul(dmul(dmul(1.0, dtheta2), dtheta1), s2);
    const double t3 = dmul(dmul(1.5, g), cos(dsub(theta1, pi / 2)));
    const double phi1 = dadd(dadd(dsub(t1, t2), t3), phi2);
    // ddtheta2 = (a + d2/d1*phi1 - m2*l1*lc2*dtheta1**2*sin(theta2) - phi2) / (m2*lc2**2 + I2 - d2**2/d1)
    const double num = dsub(dsub(dadd(a, dmul(ddiv(d2, d1), phi1)), dmul(dmul(0.5, dmul(dtheta1, dtheta1)), s2)), phi2);
    const double den = dsub(dadd(0.25, 1.0), ddiv(dmul(d2, d2), d1));
    const double ddtheta2 = ddiv(num, den);
    const double ddtheta1 = ddiv(-dadd(dmul(d2, ddtheta2), phi1), d1);
    d[0] = dtheta1; d[1] = dtheta2; d[2] = ddtheta1; d[3] = ddtheta2; d[4] = 0.;
}

__device__ __forceinline__ double wrap_pi(double x) {
    const double M = 3.141592653589793, m = -3.141592653589793, diff = M - m;
    while (x > M) x = dsub(x, diff);
    while (x < m) x = dadd(x, diff);
    return x;
}

__device__ __forceinline__ void acrobot_step(double (&st)[4], int action, float& reward, bool& done) {
    const double dt = .2, dt2 = dt / 2.0, pi = 3.141592653589793;
    double y0[5] = {st[0], st[1], st[2], st[3], (double)(action - 1)};
    double k1[5], k2[5], k3[5], k4[5], y[5];
    acrobot_dsdt(y0, k1);
#pragma unroll
    for (int i = 0; i < 5; ++i) y[i] = dadd(y0[i], dmul(dt2, k1[i]));
    acrobot_dsdt(y, k2);
#pragma unroll
    for (int i = 0; i < 5; ++i) y[i] = dadd(y0[i], dmul(dt2, k2[i]));
    acrobot_dsdt(y, k3);
#pragma unroll
    for (int i = 0; i < 5; ++i) y[i] = dadd(y0[i], dmul(dt, k3[i]));
    acrobot_dsdt(y, k4);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        y[i] = dadd(y0[i], dmul(dt / 6.0, dadd(dadd(dadd(k1[i], dmul(2, k2[i])), dmul(2, k3[i])), k4[i])));
    y[0] = wrap_pi(y[0]);
    y[1] = wrap_pi(y[1]);
    const double v1 = 4 * pi, v2 = 9 * pi;
    y[2] = fmin(fmax(y[2], -v1), v1);
    y[3] = fmin(fmax(y[3], -v2), v2);
#pragma unroll
    for (int i = 0; i < 4; ++i) st[i] = y[i];
    const bool terminal = dsub(-cos(st[0]), cos(dadd(st[1], st[0]))) > 1.;
    done = terminal;
    reward = terminal ? 0.f : -1.f;
}

template <int SD>
__device__ __forceinline__ void real_obs(int real_env, const double (&st)[4], float (&obs)[SD]) {
    if (SD == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) obs[i] = (float)st[i];
    } else {
        double s0, c0, s1, c1;
        sincos(st[0], &s0, &c0);
        sincos(st[1], &s1, &c1);
        obs[0] = (float)c0; obs[1] = (float)s0; obs[2] = (float)c1; obs[3] = (float)s1;
        obs[4 % SD] = (float)st[2]; obs[5 % SD] = (float)st[3];
    }
    (void)real_env;
}

// gym step + TimeLimit (gym/wrappers/time_limit.py): done when elapsed >= max_steps
template <int SD>
__device__ __forceinline__ void real_step(int real_env, int max_steps, double (&st)[4], int& elapsed, int action,
                                          float (&obs)[SD], float& reward, float& done) {
    bool d;
    if (SD == 4) cartpole_step(st, action, reward, d);
    else acrobot_step(st, action, reward, d);
    elapsed += 1;
    if (elapsed >= max_steps) d = true;
    real_obs<SD>(real_env, st, obs);
    done = d ? 1.f : 0.f;
}

__device__ __forceinline__ void real_reset(int real_env, const u32x4& w, double (&st)[4]) {
    const double half = real_env == LE_REAL_CARTPOLE ? 0.05 : 0.1;
    const double span = dsub(half, -half);
    st[0] = dadd(-half, dmul(span, dmul((double)w.x, 1.0 / 4294967296.0)));
    st[1] = dadd(-half, dmul(span, dmul((double)w.y, 1.0 / 4294967296.0)));
    st[2] = dadd(-half, dmul(span, dmul((double)w.z, 1.0 / 4294967296.0)));
    st[3] = dadd(-half, dmul(span, dmul((double)w.w, 1.0 / 4294967296.0)));
}

}  // namespace le

// ---------------------------------------------------------------------------------------------------
// host-side error plumbing (le_api.cu owns the thread-local message)
void le_set_error(const char* fmt, ...);
#define LE_CUDA_CHECK(expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            le_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return LE_ECUDA;                                                                        \
        }                                                                                           \
    } while (0)
