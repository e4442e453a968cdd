// le_td3.cu — TD3_discrete_vary lanes (agents/TD3_discrete_vary.py, models/actor_critic.py:22-36,69-76) on the CTA-per-lane
// machinery of le_general.cuh: one 256-thread CTA owns one agent — actor, two critics, three target nets, Adam state, the
// replay ring and the minibatch activations live in the lane slot's HBM workspace; every dense layer goes through the
// strided GEMM / thin paths of le_general.cuh.  Control flow mirrors general_loop_kernel (BaseAgent.train / test,
// agents/base_agent.py:64-227) with the TD3 pieces:
//   select_train_action / select_test_action (:155-166): actor -> * max_action -> F.gumbel_softmax(tau, hard) + N(0,1)*action_std,
//       the environment receives argmax (agents/base_agent.py:111-114), the replay ring the action VECTOR;
//   learn (:62-119): target policy smoothing, twin target critics (min), two Adam optimizers, delayed policy update through
//       critic_1 with the straight-through Gumbel-softmax gradient, Polyak on the three targets.
// Random draws: Philox streams P_TD3_EXPO / P_TD3_NORMAL (documented with the CPU restatement's stream table).
#include "le_general.cuh"
#include "le_td3_api.h"

namespace le {

constexpr uint32_t LE_P_TD3_EXPO = 8u, LE_P_TD3_NORMAL = 9u;

// Exp(1) draw i of (phase, c0, sub): -ln((w + 0.5) * 2^-32) in fp64, rounded to fp32
__device__ __forceinline__ float td3_expo(uint32_t k0, uint32_t k1, uint32_t phase, uint32_t c0, uint32_t sub, int i) {
    const u32x4 w = philox4x32_10(c0, phase, LE_P_TD3_EXPO, (uint32_t)(i >> 2) + (sub << 16), k0, k1);
    return (float)(-log(((double)pick(w, i & 3) + 0.5) * (1.0 / 4294967296.0)));
}
// N(0,1) draw i: Box-Muller in fp64 on the word pair (2h, 2h+1) of its block (same transform as the NES noise)
__device__ __forceinline__ float td3_normal(uint32_t k0, uint32_t k1, uint32_t phase, uint32_t c0, uint32_t sub, int i) {
    const u32x4 w = philox4x32_10(c0, phase, LE_P_TD3_NORMAL, (uint32_t)(i >> 2) + (sub << 16), k0, k1);
    const int h = (i & 3) >> 1;
    const double u1 = ((double)pick(w, 2 * h) + 1.0) * (1.0 / 4294967296.0), u2 = (double)pick(w, 2 * h + 1) * (1.0 / 4294967296.0);
    const double r = sqrt(-2.0 * log(u1)), t = (2.0 * 3.14159265358979323846) * u2;
    return (float)((i & 1) ? r * sin(t) : r * cos(t));
}

// F.gumbel_softmax for one row (AD <= 4): ret = hard ? (onehot(argmax y) - y) + y : y
template <int AD>
__device__ __forceinline__ void gumbel_softmax_row(const float (&logits)[AD], const float (&expo)[AD], float tau, int hard,
                                                   float (&y)[AD], float (&ret)[AD]) {
    float z[AD], mx = -INFINITY, sum = 0.f;
#pragma unroll
    for (int k = 0; k < AD; ++k) { z[k] = __fdiv_rn(logits[k] + (-logf(expo[k])), tau); mx = fmaxf(mx, z[k]); }
#pragma unroll
    for (int k = 0; k < AD; ++k) { y[k] = expf(z[k] - mx); sum += y[k]; }
#pragma unroll
    for (int k = 0; k < AD; ++k) y[k] = __fdiv_rn(y[k], sum);
    int idx = 0;
#pragma unroll
    for (int k = 1; k < AD; ++k) if (y[k] > y[idx]) idx = k;
#pragma unroll
    for (int k = 0; k < AD; ++k) ret[k] = hard ? ((k == idx ? 1.f : 0.f) - y[k]) + y[k] : y[k];
}

// an MLP in -> H (x L) -> out as a GNet with feature layers only (torch state_dict order)
__host__ __device__ inline void tnet_build(GNet* n, int in, int H, int L, int out, int q_act) {
    int p = 0, y = 0;
    const int act = q_act == LE_ACT_TANH ? 1 : 2;
    n->kind = LE_Q_DQN; n->sd = in; n->ad = out; n->nfeat = 0;
    n->slope = q_act == LE_ACT_LEAKYRELU ? 0.01f : 0.f;
    gnet_add(&n->feat[n->nfeat++], in, H, act, &p, &y);
    for (int i = 1; i < (L > 1 ? L : 1); ++i) gnet_add(&n->feat[n->nfeat++], H, H, act, &p, &y);
    gnet_add(&n->feat[n->nfeat++], H, out, 0, &p, &y);
    n->P = p; n->sum_out = y;
}

// backward of all layers for B rows: dact holds dL/d(output layer) at its y_off (everything else is overwritten);
// grad <- parameter gradients; dX (may be null) <- dL/d(input rows), row stride dxs
static __device__ void t_net_backward(const GNet& n, const float* th, float* grad, const float* X, int xs, const float* acts, float* dact, int B,
                               float* dX, int dxs, float* sm) {
    const int S = n.sum_out;
    for (int i = n.nfeat - 1; i >= 0; --i) {
        const GLayer& l = n.feat[i];
        g_layer_bwd(l, th, grad, i > 0 ? acts + n.feat[i - 1].y_off : X, i > 0 ? S : xs, acts, dact, S, B,
                    i > 0 ? dact + n.feat[i - 1].y_off : dX, i > 0 ? S : dxs, false, n.slope, sm);
    }
}

// torch.optim.Adam step of one parameter vector (same op order as g_td_update); t = step count AFTER the increment
static __device__ void t_adam(float* th, float* m, float* v, const float* grad, int P, const LearnScalars& ls, double b1pow, double b2pow) {
    const double bc1 = 1.0 - b1pow, bc2 = 1.0 - b2pow;
    const float neg_step = (float)(-(ls.lr / bc1)), bc2s = (float)sqrt(bc2);
    for (int p = threadIdx.x; p < P; p += kGThreads) {
        const float g = __ldcg(grad + p);
        float mm = __ldcg(m + p), vv = __ldcg(v + p);
        mm = mm + ls.w1 * (g - mm);
        vv = vv * ls.beta2;
        vv = vv + (ls.w2 * g) * g;
        __stcg(m + p, mm);
        __stcg(v + p, vv);
        bool exact = true;
        float up = adam_update_core(mm, vv, neg_step, bc2s, ls.eps, &exact);
        if (!exact) up = __fdiv_rn(neg_step * mm, __fdiv_rn(__fsqrt_rn(vv), bc2s) + ls.eps);
        __stcg(th + p, __ldcg(th + p) + up);
    }
    __syncthreads();
}
static __device__ void t_polyak(const float* th, float* thT, int P, const LearnScalars& ls) {
    for (int p = threadIdx.x; p < P; p += kGThreads) __stcg(thT + p, ls.tau * __ldcg(th + p) + ls.one_minus_tau * __ldcg(thT + p));
    __syncthreads();
}

struct TSlot {
    float *ring, *actor, *actorT, *m_a, *v_a, *g_a, *c1, *c1T, *m_c1, *v_c1, *c2, *c2T, *m_c2, *v_c2, *g_c, *X, *X2, *Xpi, *S2, *misc, *Y, *DX,
        *Aa, *A1, *D, *obs;
};

struct TRunParams {
    RunParams rp;
    le_td3_cfg tc;
    GNet na, nc;
    float* slots; int64_t slot_stride;
    int bmax, rowf;
    const float *actor_init, *c1_init, *c2_init;   // [n_lanes][P] or the same for all lanes when init_stride == 0
    int init_stride_a, init_stride_c;
    float* actor_final;
};

__host__ __device__ inline int64_t tslot_floats(const GNet& na, const GNet& nc, int sd, int ad, int ring_cap, int rowf, int bmax, int64_t* offs /* [26] */) {
    const int Pa = (na.P + 3) / 4 * 4, Pc = (nc.P + 3) / 4 * 4, XI = sd + ad;
    const int SM = na.sum_out > nc.sum_out ? na.sum_out : nc.sum_out;
    int64_t o = 0;
    int k = 0;
    auto take = [&](int64_t nfl) { offs[k++] = o; o += (nfl + 3) / 4 * 4; };
    take((int64_t)ring_cap * rowf);                                     // 0 ring
    take(Pa); take(Pa); take(Pa); take(Pa); take(Pa);                   // 1..5 actor, actorT, m, v, grad
    take(Pc); take(Pc); take(Pc); take(Pc);                             // 6..9 c1, c1T, m, v
    take(Pc); take(Pc); take(Pc); take(Pc); take(Pc);                   // 10..14 c2, c2T, m, v, grad (shared critic gradient buffer)
    take((int64_t)bmax * XI); take((int64_t)bmax * XI); take((int64_t)bmax * XI);   // 15 X, 16 X2, 17 Xpi
    take((int64_t)bmax * sd); take((int64_t)bmax * 4); take((int64_t)bmax * 4); take((int64_t)bmax * XI);   // 18 S2, 19 misc, 20 Y, 21 DX
    take((int64_t)bmax * na.sum_out); take((int64_t)bmax * nc.sum_out); take((int64_t)bmax * SM);          // 22 Aa, 23 A1, 24 D
    take((int64_t)64 * sd);                                             // 25 obs (test rollouts)
    return o;
}

template <int SD, int AD>
__global__ void __launch_bounds__(kGThreads, 2) td3_loop_kernel(const TRunParams G) {
    static_assert(AD <= 4, "action vectors are handled as register arrays");
    __shared__ __align__(16) float sm[kGSmemFloats];
    __shared__ float red[32];
    __shared__ __align__(16) float box[16];
    __shared__ int ibox[4];
    __shared__ double dred[kGThreads / 32];
    __shared__ int sred[kGThreads / 32];
    const RunParams& P = G.rp;
    const le_lane_cfg& c = G.tc.base;
    const GNet& na = G.na;
    const GNet& nc = G.nc;
    constexpr int XI = SD + AD;
    const int ROWF = G.rowf;     // ring row: [s(SD) | a(AD) | s'(SD) | r | d | pad]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Sa = na.sum_out, Sc = nc.sum_out, yo_a = na.feat[na.nfeat - 1].y_off, yo_c = nc.feat[nc.nfeat - 1].y_off;
    const float max_action = (float)G.tc.max_action, action_std = (float)G.tc.action_std;
    const float policy_std = (float)G.tc.policy_std, policy_clip = (float)G.tc.policy_std_clip;
    const int hard = G.tc.gumbel_hard;
    TSlot w;
    {
        int64_t offs[26];
        tslot_floats(na, nc, SD, AD, P.ring_cap, ROWF, G.bmax, offs);
        float* base = G.slots + (int64_t)blockIdx.x * G.slot_stride;
        w.ring = base + offs[0]; w.actor = base + offs[1]; w.actorT = base + offs[2]; w.m_a = base + offs[3]; w.v_a = base + offs[4];
        w.g_a = base + offs[5]; w.c1 = base + offs[6]; w.c1T = base + offs[7]; w.m_c1 = base + offs[8]; w.v_c1 = base + offs[9];
        w.c2 = base + offs[10]; w.c2T = base + offs[11]; w.m_c2 = base + offs[12]; w.v_c2 = base + offs[13]; w.g_c = base + offs[14];
        w.X = base + offs[15]; w.X2 = base + offs[16]; w.Xpi = base + offs[17]; w.S2 = base + offs[18]; w.misc = base + offs[19];
        w.Y = base + offs[20]; w.DX = base + offs[21]; w.Aa = base + offs[22]; w.A1 = base + offs[23]; w.D = base + offs[24];
        w.obs = base + offs[25];
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) ibox[0] = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int lane_id = ibox[0];
        if (lane_id >= P.n_lanes) break;
        const uint32_t k0 = P.keys[2 * lane_id], k1 = P.keys[2 * lane_id + 1];
        const float4* pack = P.env_pack + (int64_t)(P.env_index ? P.env_index[lane_id] : 0) * P.env_pack_stride;
        const bool env_tanh = c.env_act == LE_ACT_TANH;
        double* rewards = P.rewards + (int64_t)lane_id * P.rew_stride;
        int32_t* lengths = P.lengths + (int64_t)lane_id * P.rew_stride;
        double* test_rewards = P.test_rewards + (int64_t)lane_id * P.test_stride;
        const bool tracing = P.trace.cap > 0 && lane_id == P.trace_lane;
        // ---- agent construction: targets = copies, Adam state zero (agents/TD3_discrete_vary.py:41-52)
        for (int p = tid; p < na.P; p += kGThreads) {
            const float v0 = G.actor_init[(int64_t)lane_id * G.init_stride_a + p];
            w.actor[p] = v0; w.actorT[p] = v0; w.m_a[p] = 0.f; w.v_a[p] = 0.f;
        }
        for (int p = tid; p < nc.P; p += kGThreads) {
            const float a1 = G.c1_init[(int64_t)lane_id * G.init_stride_c + p], a2 = G.c2_init[(int64_t)lane_id * G.init_stride_c + p];
            w.c1[p] = a1; w.c1T[p] = a1; w.m_c1[p] = 0.f; w.v_c1[p] = 0.f;
            w.c2[p] = a2; w.c2T[p] = a2; w.m_c2[p] = 0.f; w.v_c2[p] = 0.f;
        }
        __syncthreads();
        LearnScalars ls;
        fill_learn_scalars(ls, c);
        double b1pow_a = 1.0, b2pow_a = 1.0, b1pow_c = 1.0, b2pow_c = 1.0;
        const double T0 = G.tc.gumbel_temp, Tstep = (T0 / 20.0 - T0) / 1999.0;      // np.linspace(T0, T0/20, 2000)
        float temp = (float)T0;                                                     // self.gumbel_temp_annealed
        int total_it = 0;
        const int K = c.same_action_num > 1 ? c.same_action_num : 1;

        // actor(rows) + Gumbel-softmax + action noise for M rows held in `Xrows` (row stride xs): thread m < M handles row m.
        // Returns this thread's argmax; avec <- the action vector.
        auto act_rows = [&](const float* Xrows, int xs, int M, uint32_t phase, uint32_t c0, uint32_t sub_of_thread, bool active,
                            float (&avec)[AD]) -> int {
            g_net_forward(na, w.actor, Xrows, xs, M, w.Aa, sm);
            int best = 0;
            if (active) {
                float logits[AD], expo[AD], y[AD], ret[AD];
#pragma unroll
                for (int k = 0; k < AD; ++k) {
                    logits[k] = __ldcg(w.Aa + (int64_t)tid * Sa + yo_a + k) * max_action;
                    expo[k] = td3_expo(k0, k1, phase, c0, sub_of_thread, k);
                }
                gumbel_softmax_row<AD>(logits, expo, temp, hard, y, ret);
#pragma unroll
                for (int k = 0; k < AD; ++k) {
                    avec[k] = ret[k] + td3_normal(k0, k1, phase, c0, sub_of_thread, k) * action_std;
                    if (avec[k] > avec[best]) best = k;
                }
            }
            return best;
        };

        // greedy-with-noise test rollouts on the real env, episodes = rows of one batched actor forward per step
        auto run_test = [&](int test_call, double* ep_out, int64_t& test_steps) -> double {
            double total = 0.0;
            for (int ep0 = 0; ep0 < c.test_episodes; ep0 += 64) {
                const int M = min(64, c.test_episodes - ep0);
                double st[4] = {0, 0, 0, 0};
                float obs[SD];
                int elapsed = 0, ep_steps = 0;
                float ep_rew = 0.f;
                bool running = tid < M;
                if (tid < M) {
                    real_reset(c.real_env, philox4x32_10((uint32_t)test_call, (uint32_t)(ep0 + tid), LE_P_RESET_TEST, 0u, k0, k1), st);
                    real_obs<SD>(c.real_env, st, obs);
#pragma unroll
                    for (int i = 0; i < SD; ++i) w.obs[tid * SD + i] = obs[i];
                }
                for (int t = 0; t < c.max_steps; t += K) {
                    if (!__syncthreads_or(running ? 1 : 0)) break;
                    float avec[AD];
                    const int a = act_rows(w.obs, SD, M, 1u, ((uint32_t)test_call << 16) | (uint32_t)(t / K), (uint32_t)(ep0 + tid), running, avec);
                    if (running) {
                        float r, d;
                        real_step<SD>(c.real_env, c.max_steps, st, elapsed, a, obs, r, d);
                        if (K > 1) {
                            double rsum = (double)r;
                            for (int k = 1; k < K && !(d > 0.5f); ++k) { float rk; real_step<SD>(c.real_env, c.max_steps, st, elapsed, a, obs, rk, d); rsum += (double)rk; }
                            r = (float)rsum;
                        }
#pragma unroll
                        for (int i = 0; i < SD; ++i) w.obs[tid * SD + i] = obs[i];
                        ep_rew += r;
                        ep_steps += 1;
                        if (d > 0.5f) running = false;
                    }
                }
                __syncthreads();
                if (tid < M && ep_out) ep_out[ep0 + tid] = (double)ep_rew;
                double dv = warp_allreduce_sum((tid < M) ? (double)ep_rew : 0.0);
                int sv = (tid < M) ? ep_steps : 0;
#pragma unroll
                for (int mm = 16; mm > 0; mm >>= 1) sv += __shfl_xor_sync(LE_FULL_MASK, sv, mm);
                __syncthreads();
                if (lane == 0) { dred[warp] = dv; sred[warp] = sv; }
                __syncthreads();
                for (int q = 0; q < kGThreads / 32; ++q) { total += dred[q]; test_steps += sred[q]; }
                __syncthreads();
            }
            return total / (double)c.test_episodes;
        };

        int rb_ptr = 0, rb_size = 0;
        int64_t train_steps = 0, learn_iters = 0, test_steps = 0;
        int test_calls = 0, n_ep = 0, timed_out = 0;
        const bool rule_virtual = (!c.use_test_env) && c.env_kind == LE_ENV_SE;
        for (int episode = 0; episode < c.train_episodes; ++episode) {
            if (c.step_budget > 0 && train_steps >= c.step_budget) { timed_out = 1; break; }
            double st[4];
            real_reset(c.real_env, philox4x32_10((uint32_t)episode, 0u, LE_P_RESET_TRAIN, 0u, k0, k1), st);
            float state[SD];
            real_obs<SD>(c.real_env, st, state);
            int elapsed = 0, ep_len = 0;
            float ep_rew = 0.f;
            for (int t = 0; t < c.max_steps; t += K) {
                // ---- select_train_action (:155-162)
                float avec[AD];
                int action;
                if (episode < c.init_episodes) {
                    const u32x4 wa = philox4x32_10((uint32_t)train_steps, 0u, LE_P_ACT, 0u, k0, k1);
                    action = (int)__umulhi(wa.y, (uint32_t)AD);
#pragma unroll
                    for (int k = 0; k < AD; ++k) avec[k] = k == action ? 1.f : 0.f;
                } else {
                    __syncthreads();
                    if (tid < SD) __stcg(w.X + tid, state[tid]);      // row 0 of the staging area (free between learn() calls)
                    __syncthreads();
                    float av0[AD];
                    const int a0 = act_rows(w.X, XI, 1, 0u, (uint32_t)train_steps, 0u, tid == 0, av0);
                    if (tid == 0) {
                        ibox[1] = a0;
#pragma unroll
                        for (int k = 0; k < AD; ++k) box[8 + k] = av0[k];
                    }
                    __syncthreads();
                    action = ibox[1];
#pragma unroll
                    for (int k = 0; k < AD; ++k) avec[k] = box[8 + k];
                    __syncthreads();
                }
                // ---- env.step(action.argmax())
                float ns[SD], r = 0.f, d = 0.f;
                if (c.env_kind == LE_ENV_SE) {
                    __syncthreads();
                    if (warp == 0) {
                        float cur[SD], ns0[SD], r0 = 0.f, d0 = 0.f;
#pragma unroll
                        for (int i = 0; i < SD; ++i) cur[i] = state[i];
                        for (int k = 0; k < K; ++k) {
                            float rk;
                            se_step_row<SD, AD>(pack, c.env_hidden, env_tanh, cur, action, lane, ns0, rk, d0);
                            r0 = k == 0 ? rk : r0 + rk;
#pragma unroll
                            for (int i = 0; i < SD; ++i) cur[i] = ns0[i];
                        }
                        if (lane == 0) {
#pragma unroll
                            for (int i = 0; i < SD; ++i) box[i] = ns0[i];
                            box[SD] = r0; box[SD + 1] = d0;
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int i = 0; i < SD; ++i) ns[i] = box[i];
                    r = box[SD]; d = box[SD + 1];
                    __syncthreads();
                } else {
                    double rsum = 0.0;
                    for (int k = 0; k < K; ++k) {
                        float rr;
                        real_step<SD>(c.real_env, c.max_steps, st, elapsed, action, ns, rr, d);
                        rsum += (double)rr;
                        r = K == 1 ? rr : (float)rsum;
                        if (d > 0.5f) break;
                    }
                }
                // ---- replay_buffer.add with the action vector
                if (tid == 0) {
                    float* row = w.ring + (int64_t)rb_ptr * ROWF;
#pragma unroll
                    for (int i = 0; i < SD; ++i) { __stcg(row + i, state[i]); __stcg(row + SD + AD + i, ns[i]); }
#pragma unroll
                    for (int k = 0; k < AD; ++k) __stcg(row + SD + k, avec[k]);
                    __stcg(row + 2 * SD + AD, r);
                    __stcg(row + 2 * SD + AD + 1, d);
                }
                rb_ptr = (rb_ptr + 1 == P.ring_cap) ? 0 : rb_ptr + 1;
                rb_size = rb_size + 1 < P.ring_cap ? rb_size + 1 : P.ring_cap;
#pragma unroll
                for (int i = 0; i < SD; ++i) state[i] = ns[i];
                ep_rew += r;
                ep_len += K;
                float loss = __int_as_float(0x7fc00000);
                if (episode >= c.init_episodes) {
                    // ---- learn (:62-119)
                    __syncthreads();
                    temp = (float)(total_it >= 1999 ? T0 / 20.0 : (double)total_it * Tstep + T0);
                    total_it += 1;
                    const int B = ls.batch;
                    for (int b = tid; b < B; b += kGThreads) {     // replay_buffer.sample on the P_SAMPLE stream
                        const u32x4 wv = philox4x32_10((uint32_t)learn_iters, (uint32_t)(b >> 2), LE_P_SAMPLE, 0u, k0, k1);
                        const float* src = w.ring + (int64_t)__umulhi(pick(wv, b & 3), (uint32_t)rb_size) * ROWF;
#pragma unroll
                        for (int i = 0; i < XI; ++i) w.X[b * XI + i] = __ldcg(src + i);
#pragma unroll
                        for (int i = 0; i < SD; ++i) w.S2[b * SD + i] = __ldcg(src + XI + i);
                        w.misc[4 * b] = __ldcg(src + 2 * SD + AD);
                        w.misc[4 * b + 1] = __ldcg(src + 2 * SD + AD + 1);
                    }
                    __syncthreads();
                    // target policy smoothing: a' = gumbel_softmax(actor_target(s')) + clip(N * policy_std)
                    g_net_forward(na, w.actorT, w.S2, SD, B, w.Aa, sm);
                    for (int b = tid; b < B; b += kGThreads) {
                        float logits[AD], expo[AD], y[AD], ret[AD];
#pragma unroll
                        for (int k = 0; k < AD; ++k) {
                            logits[k] = __ldcg(w.Aa + (int64_t)b * Sa + yo_a + k) * max_action;
                            expo[k] = td3_expo(k0, k1, 2u, (uint32_t)learn_iters, 0u, b * AD + k);
                        }
                        gumbel_softmax_row<AD>(logits, expo, temp, hard, y, ret);
#pragma unroll
                        for (int i = 0; i < SD; ++i) w.X2[b * XI + i] = w.S2[b * SD + i];
#pragma unroll
                        for (int k = 0; k < AD; ++k) {
                            float nz = td3_normal(k0, k1, 2u, (uint32_t)learn_iters, 0u, b * AD + k) * policy_std;
                            nz = nz < -policy_clip ? -policy_clip : (nz > policy_clip ? policy_clip : nz);
                            w.X2[b * XI + SD + k] = ret[k] + nz;
                        }
                    }
                    __syncthreads();
                    // target_Q = r + (1 - d) * gamma * min(Q1', Q2')
                    g_net_forward(nc, w.c1T, w.X2, XI, B, w.A1, sm);
                    for (int b = tid; b < B; b += kGThreads) w.misc[4 * b + 2] = __ldcg(w.A1 + (int64_t)b * Sc + yo_c);
                    __syncthreads();
                    g_net_forward(nc, w.c2T, w.X2, XI, B, w.A1, sm);
                    for (int b = tid; b < B; b += kGThreads) {
                        const float q1 = w.misc[4 * b + 2], q2 = __ldcg(w.A1 + (int64_t)b * Sc + yo_c);
                        w.misc[4 * b + 2] = w.misc[4 * b] + ((1.f - w.misc[4 * b + 1]) * ls.gamma) * fminf(q1, q2);
                    }
                    __syncthreads();
                    // critic loss and the two critic steps (one optimizer: shared step count)
                    b1pow_c *= ls.beta1; b2pow_c *= ls.beta2d;
                    float lsum = 0.f;
                    for (int which = 0; which < 2; ++which) {
                        float* cth = which ? w.c2 : w.c1;
                        g_net_forward(nc, cth, w.X, XI, B, w.A1, sm);
                        float lpart = 0.f;
                        for (int b = tid; b < B; b += kGThreads) {
                            const float e = __ldcg(w.A1 + (int64_t)b * Sc + yo_c) - w.misc[4 * b + 2];
                            lpart = fmaf(e, e, lpart);
                            __stcg(w.D + (int64_t)b * Sc + yo_c, ls.norm * e);
                        }
                        lsum += g_block_sum(lpart, red) / (float)B;
                        __syncthreads();
                        t_net_backward(nc, cth, w.g_c, w.X, XI, w.A1, w.D, B, nullptr, 0, sm);
                        t_adam(cth, which ? w.m_c2 : w.m_c1, which ? w.v_c2 : w.v_c1, w.g_c, nc.P, ls, b1pow_c, b2pow_c);
                    }
                    loss = lsum;
                    if (total_it % G.tc.policy_delay == 0) {
                        // actor loss = -mean(critic_1(s, actor(s))) through the (updated) critic and the straight-through Gumbel-softmax
                        g_net_forward(na, w.actor, w.X, XI, B, w.Aa, sm);
                        for (int b = tid; b < B; b += kGThreads) {
                            float logits[AD], expo[AD], y[AD], ret[AD];
#pragma unroll
                            for (int k = 0; k < AD; ++k) {
                                logits[k] = __ldcg(w.Aa + (int64_t)b * Sa + yo_a + k) * max_action;
                                expo[k] = td3_expo(k0, k1, 3u, (uint32_t)learn_iters, 0u, b * AD + k);
                            }
                            gumbel_softmax_row<AD>(logits, expo, temp, hard, y, ret);
#pragma unroll
                            for (int i = 0; i < SD; ++i) w.Xpi[b * XI + i] = w.X[b * XI + i];
#pragma unroll
                            for (int k = 0; k < AD; ++k) { w.Xpi[b * XI + SD + k] = ret[k]; w.Y[4 * b + k] = y[k]; }
                        }
                        __syncthreads();
                        g_net_forward(nc, w.c1, w.Xpi, XI, B, w.A1, sm);
                        for (int b = tid; b < B; b += kGThreads) __stcg(w.D + (int64_t)b * Sc + yo_c, -1.f / (float)B);
                        __syncthreads();
                        t_net_backward(nc, w.c1, w.g_c, w.Xpi, XI, w.A1, w.D, B, w.DX, XI, sm);     // only dL/d(action) is used
                        for (int b = tid; b < B; b += kGThreads) {
                            float dot = 0.f;
#pragma unroll
                            for (int k = 0; k < AD; ++k) dot += w.Y[4 * b + k] * __ldcg(w.DX + (int64_t)b * XI + SD + k);
#pragma unroll
                            for (int k = 0; k < AD; ++k)
                                __stcg(w.D + (int64_t)b * Sa + yo_a + k,
                                       __fdiv_rn(w.Y[4 * b + k] * (__ldcg(w.DX + (int64_t)b * XI + SD + k) - dot), temp) * max_action);
                        }
                        __syncthreads();
                        t_net_backward(na, w.actor, w.g_a, w.X, XI, w.Aa, w.D, B, nullptr, 0, sm);
                        b1pow_a *= ls.beta1; b2pow_a *= ls.beta2d;
                        t_adam(w.actor, w.m_a, w.v_a, w.g_a, na.P, ls, b1pow_a, b2pow_a);
                        t_polyak(w.c1, w.c1T, nc.P, ls);
                        t_polyak(w.c2, w.c2T, nc.P, ls);
                        t_polyak(w.actor, w.actorT, na.P, ls);
                    }
                    learn_iters += 1;
                }
                if (tracing && train_steps < P.trace.cap && tid == 0) {
                    const int64_t i = train_steps;
                    P.trace.action[i] = action; P.trace.explore[i] = episode < c.init_episodes ? 1 : 0;
                    P.trace.reward[i] = r; P.trace.done[i] = d; P.trace.loss[i] = loss;
#pragma unroll
                    for (int k = 0; k < SD; ++k) P.trace.next_state[i * SD + k] = ns[k];
                }
                train_steps += 1;
                if (d > 0.5f) break;
            }
            double ep_value;
            if (c.use_test_env) ep_value = run_test(test_calls++, nullptr, test_steps);
            else ep_value = (double)ep_rew;
            __syncthreads();
            if (tid == 0) { lengths[n_ep] = ep_len; rewards[n_ep] = ep_value; __threadfence_block(); }
            n_ep += 1;
            __syncthreads();
            if (episode >= c.init_episodes) {
                const double avg = mean_window(rewards, n_ep, c.early_out_num, 0);
                bool solved;
                if (rule_virtual) {
                    const double avg_last = mean_window(rewards, n_ep, c.early_out_num, c.early_out_num);
                    solved = (fabs(avg - avg_last) / (fabs(avg_last) + 1e-9) < c.early_out_virtual_diff) &&
                             (episode >= c.init_episodes + c.early_out_num);
                } else solved = avg >= c.solved_reward;
                if (solved) break;
            }
        }
        double score = 0.0;
        if (c.final_test) score = run_test(test_calls++, test_rewards, test_steps);
        __syncthreads();
        if (G.actor_final) for (int p = tid; p < na.P; p += kGThreads) G.actor_final[(int64_t)lane_id * na.P + p] = w.actor[p];
        if (tid == 0) {
            le_lane_out o;
            o.n_episodes = n_ep; o.timed_out = timed_out; o.train_steps = train_steps; o.learn_iters = learn_iters;
            o.test_steps = test_steps; o.score = score;
            P.out[lane_id] = o;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
int td3_plan(const le_td3_cfg* tc, int n_lanes, int ring_cap, int sms, Td3Plan* tp) {
    const le_lane_cfg* c = &tc->base;
    GNet na, nc;
    tnet_build(&na, c->sd, c->q_hidden, c->q_layers, c->ad, c->q_act);
    tnet_build(&nc, c->sd + c->ad, c->q_hidden, c->q_layers, 1, c->q_act);
    int occ = 0;
    if (c->sd == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, td3_loop_kernel<4, 2>, kGThreads, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, td3_loop_kernel<6, 3>, kGThreads, 0);
    if (occ < 1) occ = 1;
    tp->grid = n_lanes < sms * occ ? n_lanes : sms * occ;
    tp->bmax = c->batch_size > 64 ? c->batch_size : 64;
    tp->rowf = (2 * c->sd + c->ad + 2 + 3) / 4 * 4;
    int64_t offs[26];
    tp->slot_floats = tslot_floats(na, nc, c->sd, c->ad, ring_cap, tp->rowf, tp->bmax, offs);
    tp->p_actor = na.P;
    tp->p_critic = nc.P;
    return LE_OK;
}

cudaError_t td3_launch(const le_td3_cfg* tc, const RunParams& rp, float* slots, const Td3Plan& tp, const float* actor_init, const float* c1_init,
                       const float* c2_init, int per_lane_init, float* actor_final, cudaStream_t st) {
    TRunParams G;
    G.rp = rp;
    G.tc = *tc;
    const le_lane_cfg* c = &tc->base;
    tnet_build(&G.na, c->sd, c->q_hidden, c->q_layers, c->ad, c->q_act);
    tnet_build(&G.nc, c->sd + c->ad, c->q_hidden, c->q_layers, 1, c->q_act);
    G.slots = slots; G.slot_stride = tp.slot_floats; G.bmax = tp.bmax; G.rowf = tp.rowf;
    G.actor_init = actor_init; G.c1_init = c1_init; G.c2_init = c2_init;
    G.init_stride_a = per_lane_init ? tp.p_actor : 0;
    G.init_stride_c = per_lane_init ? tp.p_critic : 0;
    G.actor_final = actor_final;
    if (c->sd == 4) td3_loop_kernel<4, 2><<<tp.grid, kGThreads, 0, st>>>(G);
    else td3_loop_kernel<6, 3><<<tp.grid, kGThreads, 0, st>>>(G);
    return cudaGetLastError();
}

}  // namespace le
