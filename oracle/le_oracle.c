/*
 * le_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded-per-lane CPU restatement of the reference's NES inner loop.  It is the checker
 * for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs) and
 * is never imported, linked or executed by the product package.
 *
 * Each function cites the reference file:line (under /root/reference) it follows.  The restatement is pinned
 * against golden vectors produced by the UNMODIFIED reference under RNG injection (oracle/gen_golden.py ->
 * tests/golden/, checked in tests/test_oracle_vs_golden.py).  The gym 0.17.3 dynamics are a third-party
 * dependency absent from /root/reference: restated from the published sources,
 * "parity unpinned" against upstream gym itself.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, so fp32/fp64 results follow
 * torch-eager / CPython operation-by-operation rounding).
 */
#include "le_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011; Random123 constants) — same streams as oracle/philox.py           */

void le_oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline uint32_t mulhi32(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

/* ------------------------------------------------------------------------------------------------ */
/* activations (models/model_utils.py:9-20)                                                            */

static inline float act_f(int act, float slope, float z) {
    switch (act) {
        case LE_ACT_TANH: return tanhf(z);
        case LE_ACT_RELU: return z > 0.f ? z : 0.f;
        case LE_ACT_LEAKYRELU: return z > 0.f ? z : 0.01f * z;
        case LE_ACT_PRELU: return z > 0.f ? z : slope * z;
        default: return z;
    }
}
/* derivative w.r.t. z given z and h = act(z) (autograd of the above) */
static inline float act_grad(int act, float slope, float z, float h) {
    switch (act) {
        case LE_ACT_TANH: return 1.f - h * h;
        case LE_ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case LE_ACT_LEAKYRELU: return z > 0.f ? 1.f : 0.01f;
        case LE_ACT_PRELU: return z > 0.f ? 1.f : slope;
        default: return 1.f;
    }
}

/* One-hidden-layer MLP, torch layout: W1[H][in], b1[H], W2[out][H], b2[out]
 * (models/model_utils.py:31-38 with hidden_layer <= 1: Linear, act, Linear). */
static void mlp_forward(const float* th, int in, int H, int out, int act, float slope, const float* x, float* y,
                        float* hbuf /* optional [H] */, float* zbuf /* optional [H] */) {
    const float* W1 = th;
    const float* b1 = W1 + (size_t)H * in;
    const float* W2 = b1 + H;
    const float* b2 = W2 + (size_t)out * H;
    for (int o = 0; o < out; ++o) y[o] = b2[o];
    for (int j = 0; j < H; ++j) {
        float z = b1[j];
        for (int i = 0; i < in; ++i) z += W1[(size_t)j * in + i] * x[i];
        float h = act_f(act, slope, z);
        if (hbuf) hbuf[j] = h;
        if (zbuf) zbuf[j] = z;
        for (int o = 0; o < out; ++o) y[o] += W2[(size_t)o * H + j] * h;
    }
}

int le_oracle_mlp_params(int in, int H, int out) { return H * in + H + out * H + out; }
int le_oracle_se_params(const le_lane_cfg* c) {
    int in = c->sd + c->ad, H = c->env_hidden;
    return le_oracle_mlp_params(in, H, c->sd) + 2 * le_oracle_mlp_params(in, H, 1);
}
int le_oracle_rn_params(const le_lane_cfg* c) { return le_oracle_mlp_params(c->sd, c->env_hidden, 1); }
static int is_simple_dqn(const le_lane_cfg* c);
static int general_q_params(const le_lane_cfg* c);
int le_oracle_q_params(const le_lane_cfg* c) { return is_simple_dqn(c) ? le_oracle_mlp_params(c->sd, c->q_hidden, c->ad) : general_q_params(c); }

/* VirtualEnv.step (envs/virtual_env.py:43-54): input = cat(one_hot(action), state); three nets. */
void le_oracle_se_step(const le_lane_cfg* c, const float* theta, const float* state, int action, float* next_state,
                       float* reward, float* done) {
    int in = c->sd + c->ad, H = c->env_hidden;
    float x[LE_ORACLE_MAX_IN];
    for (int a = 0; a < c->ad; ++a) x[a] = (a == action) ? 1.f : 0.f; /* utils.py:108-126 */
    for (int i = 0; i < c->sd; ++i) x[c->ad + i] = state[i];
    const float* th_s = theta;
    const float* th_r = th_s + le_oracle_mlp_params(in, H, c->sd);
    const float* th_d = th_r + le_oracle_mlp_params(in, H, 1);
    mlp_forward(th_s, in, H, c->sd, c->env_act, c->env_slope[0], x, next_state, NULL, NULL);
    mlp_forward(th_r, in, H, 1, c->env_act, c->env_slope[1], x, reward, NULL, NULL);
    mlp_forward(th_d, in, H, 1, c->env_act, c->env_slope[2], x, done, NULL, NULL);
}

/* RewardEnv._calc_reward (envs/reward_env.py:68-133), state-only types. Returns <0 for info-vector types
 * (the reference raises ValueError('No info dict…') for CartPole/Acrobot, envs/reward_env.py:91-92). */
int le_oracle_rn_reward(const le_lane_cfg* c, const float* theta, const float* s, const float* s2, float real_reward,
                        float* out) {
    float ps = 0.f, ps2 = 0.f;
    float g = (float)c->gamma;
    int H = c->env_hidden;
    switch (c->rn_type) {
        case 0: *out = real_reward; return 0;
        case 1:
        case 2:
            mlp_forward(theta, c->sd, H, 1, c->env_act, c->env_slope[0], s2, &ps2, NULL, NULL);
            mlp_forward(theta, c->sd, H, 1, c->env_act, c->env_slope[0], s, &ps, NULL, NULL);
            if (c->rn_type == 1) *out = g * ps2 - ps;
            else *out = real_reward + g * ps2 - ps; /* (r + γΦ(s')) − Φ(s) */
            return 0;
        case 5:
        case 6:
            mlp_forward(theta, c->sd, H, 1, c->env_act, c->env_slope[0], s2, &ps2, NULL, NULL);
            *out = (c->rn_type == 5) ? ps2 : real_reward + ps2;
            return 0;
        default: return -1;
    }
}

void le_oracle_q_forward_general(const le_lane_cfg* c, const float* th, const float* state, float* q, int* argmax);
void le_oracle_q_forward(const le_lane_cfg* c, const float* q_theta, const float* state, float* q, int* argmax) {
    if (!(c->q_kind == LE_Q_DQN && c->q_layers <= 1)) { le_oracle_q_forward_general(c, q_theta, state, q, argmax); return; }
    mlp_forward(q_theta, c->sd, c->q_hidden, c->ad, c->q_act, 0.f, state, q, NULL, NULL);
    int best = 0;
    for (int a = 1; a < c->ad; ++a)
        if (q[a] > q[best]) best = a; /* torch.argmax: first maximal index */
    if (argmax) *argmax = best;
}


/* ------------------------------------------------------------------------------------------------ */
/* General Q-networks: Critic_DQN with hidden_layer >= 1 and Critic_DuelingDQN (models/actor_critic.py:84-122)  */

typedef struct { int in, out, act, w_off, b_off, y_off; } olayer; /* y_off: offset of the layer output in a row's activation record */
typedef struct {
    int kind, nfeat, P, sum_out, sd, ad;
    olayer feat[4], val[2], adv[2];
} onet;

static void add_layer(olayer* l, int in, int out, int act, int* p, int* y) {
    l->in = in; l->out = out; l->act = act; l->w_off = *p; *p += in * out; l->b_off = *p; *p += out; l->y_off = *y; *y += out;
}

static void build_net(const le_lane_cfg* c, onet* n) {
    int p = 0, y = 0;
    const int L = c->q_layers > 1 ? c->q_layers : 1, H = c->q_hidden, act = c->q_act;
    n->kind = c->q_kind; n->sd = c->sd; n->ad = c->ad; n->nfeat = 0;
    add_layer(&n->feat[n->nfeat++], c->sd, H, act, &p, &y);
    for (int i = 1; i < L; ++i) add_layer(&n->feat[n->nfeat++], H, H, act, &p, &y);
    if (c->q_kind == LE_Q_DQN) {
        add_layer(&n->feat[n->nfeat++], H, c->ad, LE_ACT_IDENTITY, &p, &y);
    } else {
        const int fd = c->q_feature_dim;
        add_layer(&n->feat[n->nfeat++], H, fd, LE_ACT_IDENTITY, &p, &y); /* no activation after the feature stream */
        add_layer(&n->val[0], fd, fd, act, &p, &y);
        add_layer(&n->val[1], fd, 1, LE_ACT_IDENTITY, &p, &y);
        add_layer(&n->adv[0], fd, fd, act, &p, &y);
        add_layer(&n->adv[1], fd, c->ad, LE_ACT_IDENTITY, &p, &y);
    }
    n->P = p; n->sum_out = y;
}

static int is_simple_dqn(const le_lane_cfg* c) { return c->q_kind == LE_Q_DQN && c->q_layers <= 1; }
static int general_q_params(const le_lane_cfg* c) { onet n; build_net(c, &n); return n.P; }

static void layer_fwd(const olayer* l, const float* th, const float* x, float* y) {
    for (int o = 0; o < l->out; ++o) {
        float z = th[l->b_off + o];
        const float* w = th + l->w_off + (size_t)o * l->in;
        for (int i = 0; i < l->in; ++i) z += w[i] * x[i];
        y[o] = act_f(l->act, 0.f, z);
    }
}

/* one row through the net; acts = the row's activation record [sum_out]; returns pointers to V (or NULL) and A/q */
static void net_forward_row(const onet* n, const float* th, const float* x, float* acts, const float** v_out, const float** a_out) {
    const float* in = x;
    for (int i = 0; i < n->nfeat; ++i) { layer_fwd(&n->feat[i], th, in, acts + n->feat[i].y_off); in = acts + n->feat[i].y_off; }
    if (n->kind == LE_Q_DQN) { *v_out = NULL; *a_out = in; return; }
    layer_fwd(&n->val[0], th, in, acts + n->val[0].y_off);
    layer_fwd(&n->val[1], th, acts + n->val[0].y_off, acts + n->val[1].y_off);
    layer_fwd(&n->adv[0], th, in, acts + n->adv[0].y_off);
    layer_fwd(&n->adv[1], th, acts + n->adv[0].y_off, acts + n->adv[1].y_off);
    *v_out = acts + n->val[1].y_off; *a_out = acts + n->adv[1].y_off;
}

/* q-values of a batch: DQN -> the output layer; dueling -> V + (A - mean over ALL rows and actions of A) (:121) */
static void net_forward_batch(const onet* n, const float* th, const float* X, int x_stride, int B, float* acts, float* q) {
    double dummy = 0.0; (void)dummy;
    float asum = 0.f;
    for (int b = 0; b < B; ++b) {
        const float *v, *a;
        net_forward_row(n, th, X + (size_t)b * x_stride, acts + (size_t)b * n->sum_out, &v, &a);
        for (int k = 0; k < n->ad; ++k) { q[b * n->ad + k] = a[k]; asum += a[k]; }
    }
    if (n->kind == LE_Q_DUELING) {
        const float mean = asum / (float)(B * n->ad);
        for (int b = 0; b < B; ++b) {
            const float v = acts[(size_t)b * n->sum_out + n->val[1].y_off];
            for (int k = 0; k < n->ad; ++k) q[b * n->ad + k] = v + (q[b * n->ad + k] - mean);
        }
    }
}

static void layer_bwd(const olayer* l, const float* th, float* g, const float* x, const float* y, float* dy /* in: dL/dy, destroyed */,
                      float* dx /* accumulated into, may be NULL */) {
    for (int o = 0; o < l->out; ++o) {
        const float dz = dy[o] * act_grad(l->act, 0.f, y[o], y[o]);  /* relu/leaky: sign(y) == sign(z) */
        g[l->b_off + o] += dz;
        float* gw = g + l->w_off + (size_t)o * l->in;
        const float* w = th + l->w_off + (size_t)o * l->in;
        for (int i = 0; i < l->in; ++i) { gw[i] += dz * x[i]; if (dx) dx[i] += dz * w[i]; }
    }
}

void le_oracle_q_forward_general(const le_lane_cfg* c, const float* th, const float* state, float* q, int* argmax) {
    onet n; build_net(c, &n);
    float* acts = (float*)malloc(sizeof(float) * n.sum_out);
    net_forward_batch(&n, th, state, c->sd, 1, acts, q);
    int best = 0;
    for (int a = 1; a < c->ad; ++a) if (q[a] > q[best]) best = a;
    if (argmax) *argmax = best;
    free(acts);
}

static float td_update_general(const le_lane_cfg* c, float* th, float* thT, float* m, float* v, int32_t* adam_t, const float* rows, int B) {
    onet n; build_net(c, &n);
    const int sd = c->sd, ad = c->ad, P = n.P, ROW = 2 * sd + 3, S = n.sum_out;
    float* acts = (float*)malloc(sizeof(float) * (size_t)B * S * 2);
    float* acts2 = acts + (size_t)B * S;
    float* q = (float*)malloc(sizeof(float) * (size_t)B * ad * 3);
    float* q2 = q + B * ad; float* qT = q2 + B * ad;
    float* g = (float*)calloc((size_t)P, sizeof(float));
    float* dbuf = (float*)calloc((size_t)S + sd, sizeof(float));
    net_forward_batch(&n, th, rows, ROW, B, acts, q);                 /* q_values = model(states) */
    net_forward_batch(&n, th, rows + sd + 1, ROW, B, acts2, q2);      /* next_q_values = model(next_states) */
    net_forward_batch(&n, thT, rows + sd + 1, ROW, B, acts2, qT);     /* model_target(next_states) */
    const float gam = (float)c->gamma, norm = (float)(2.0 / (double)B);
    float* dq = (float*)calloc((size_t)B, sizeof(float));
    float loss_f = 0.f, gsum = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* r = rows + (size_t)b * ROW;
        const int a = (int)r[sd];
        int astar = 0;
        for (int k = 1; k < ad; ++k) if (q2[b * ad + k] > q2[b * ad + astar]) astar = k;
        const float y = r[2 * sd + 1] + gam * qT[b * ad + astar] * (1.f - r[2 * sd + 2]);
        const float delta = q[b * ad + a] - y;
        loss_f += delta * delta;
        dq[b] = norm * delta;
        gsum += dq[b];
    }
    const float loss = loss_f / (float)B;
    const float mean_g = gsum / (float)(B * ad);   /* d/dA of -mean(A): every element gets -sum(dq)/(B*ad) */
    for (int b = 0; b < B; ++b) {
        const float* r = rows + (size_t)b * ROW;
        const int a = (int)r[sd];
        const float* A = acts + (size_t)b * S;
        memset(dbuf, 0, sizeof(float) * (S + sd));
        float* d = dbuf; /* gradient w.r.t. each layer output, same offsets as the activation record */
        if (n.kind == LE_Q_DQN) {
            const olayer* lo = &n.feat[n.nfeat - 1];
            d[lo->y_off + a] = dq[b];
        } else {
            d[n.val[1].y_off] = dq[b];
            for (int k = 0; k < ad; ++k) d[n.adv[1].y_off + k] = (k == a ? dq[b] : 0.f) - mean_g;
            const olayer* lf = &n.feat[n.nfeat - 1];
            layer_bwd(&n.val[1], th, g, A + n.val[0].y_off, A + n.val[1].y_off, d + n.val[1].y_off, d + n.val[0].y_off);
            layer_bwd(&n.val[0], th, g, A + lf->y_off, A + n.val[0].y_off, d + n.val[0].y_off, d + lf->y_off);
            layer_bwd(&n.adv[1], th, g, A + n.adv[0].y_off, A + n.adv[1].y_off, d + n.adv[1].y_off, d + n.adv[0].y_off);
            layer_bwd(&n.adv[0], th, g, A + lf->y_off, A + n.adv[0].y_off, d + n.adv[0].y_off, d + lf->y_off);
        }
        for (int i = n.nfeat - 1; i >= 0; --i) {
            const olayer* l = &n.feat[i];
            const float* x = i > 0 ? A + n.feat[i - 1].y_off : r;
            layer_bwd(l, th, g, x, A + l->y_off, d + l->y_off, i > 0 ? d + n.feat[i - 1].y_off : NULL);
        }
    }
    *adam_t += 1;
    const double b1d = c->beta1, b2d = c->beta2;
    const float w1 = (float)(1.0 - b1d), b2f = (float)b2d, w2 = (float)(1.0 - b2d);
    const double bc1 = 1.0 - pow(b1d, (double)*adam_t), bc2 = 1.0 - pow(b2d, (double)*adam_t);
    const float neg_step = (float)(-(c->lr / bc1)), bc2s = (float)sqrt(bc2), epsf = (float)c->adam_eps;
    const float tau = (float)c->tau, omt = (float)(1.0 - c->tau);
    for (int p = 0; p < P; ++p) {
        m[p] = m[p] + w1 * (g[p] - m[p]);
        v[p] = v[p] * b2f;
        v[p] = v[p] + w2 * g[p] * g[p];
        const float denom = sqrtf(v[p]) / bc2s + epsf;
        th[p] = th[p] + neg_step * m[p] / denom;
        thT[p] = tau * th[p] + omt * thT[p];
    }
    free(acts); free(q); free(g); free(dbuf); free(dq);
    return loss;
}

/* ------------------------------------------------------------------------------------------------ */
/* TD3_discrete_vary.learn (agents/TD3_discrete_vary.py:62-119) — groundwork for the next row of SURVEY.md §8(f):
 * actor = Actor_TD3_discrete (models/actor_critic.py:22-36: MLP * max_action -> F.gumbel_softmax), two Critic_Q
 * (:69-76: MLP over cat(state, action)), three target nets, two Adam optimizers, delayed policy update.  Randomness
 * (policy noise, Gumbel noise) is an INPUT, so the restatement is independent of any stream definition. */

typedef struct { int nl; olayer l[4]; int P, S, in, out; } omlp;

static void build_mlp(omlp* n, int in, int H, int L, int out, int act) {
    int p = 0, y = 0;
    n->nl = 0; n->in = in; n->out = out;
    add_layer(&n->l[n->nl++], in, H, act, &p, &y);
    for (int i = 1; i < (L > 1 ? L : 1); ++i) add_layer(&n->l[n->nl++], H, H, act, &p, &y);
    add_layer(&n->l[n->nl++], H, out, LE_ACT_IDENTITY, &p, &y);
    n->P = p; n->S = y;
}
static const float* mlp_fwd_row(const omlp* n, const float* th, const float* x, float* acts) {
    const float* in = x;
    for (int i = 0; i < n->nl; ++i) { layer_fwd(&n->l[i], th, in, acts + n->l[i].y_off); in = acts + n->l[i].y_off; }
    return in;
}
/* backward of one row: d = gradient record (same offsets as acts, destroyed), dx_in (may be NULL) receives dL/dx */
static void mlp_bwd_row(const omlp* n, const float* th, float* g, const float* x, const float* acts, float* d, float* dx_in) {
    for (int i = n->nl - 1; i >= 0; --i) {
        const olayer* l = &n->l[i];
        layer_bwd(l, th, g, i > 0 ? acts + n->l[i - 1].y_off : x, acts + l->y_off, d + l->y_off, i > 0 ? d + n->l[i - 1].y_off : dx_in);
    }
}
/* F.gumbel_softmax(logits, tau, hard) for one row; expo = the Exp(1) samples torch draws (gumbel = -log(expo)) */
static void gumbel_softmax_row(const float* logits, const float* expo, int n, float tau, int hard, float* y_soft, float* ret) {
    float z[LE_ORACLE_MAX_AD], mx = -INFINITY, sum = 0.f;
    for (int k = 0; k < n; ++k) { z[k] = (logits[k] + (-logf(expo[k]))) / tau; if (z[k] > mx) mx = z[k]; }
    for (int k = 0; k < n; ++k) { y_soft[k] = expf(z[k] - mx); sum += y_soft[k]; }
    for (int k = 0; k < n; ++k) y_soft[k] = y_soft[k] / sum;
    if (!hard) { for (int k = 0; k < n; ++k) ret[k] = y_soft[k]; return; }
    int idx = 0;
    for (int k = 1; k < n; ++k) if (y_soft[k] > y_soft[idx]) idx = k;
    for (int k = 0; k < n; ++k) ret[k] = ((k == idx ? 1.f : 0.f) - y_soft[k]) + y_soft[k];   /* y_hard - y_soft.detach() + y_soft */
}
static void adam_vec(float* th, float* m, float* v, const float* g, int P, int t, double lr, double b1d, double b2d, double epsd) {
    const float w1 = (float)(1.0 - b1d), b2f = (float)b2d, w2 = (float)(1.0 - b2d);
    const double bc1 = 1.0 - pow(b1d, (double)t), bc2 = 1.0 - pow(b2d, (double)t);
    const float neg_step = (float)(-(lr / bc1)), bc2s = (float)sqrt(bc2), epsf = (float)epsd;
    for (int p = 0; p < P; ++p) {
        m[p] = m[p] + w1 * (g[p] - m[p]);
        v[p] = v[p] * b2f;
        v[p] = v[p] + w2 * g[p] * g[p];
        th[p] = th[p] + neg_step * m[p] / (sqrtf(v[p]) / bc2s + epsf);
    }
}

int le_oracle_td3_params(int sd, int ad, int H, int L, int* P_actor, int* P_critic) {
    omlp a, c; build_mlp(&a, sd, H, L, ad, LE_ACT_TANH); build_mlp(&c, sd + ad, H, L, 1, LE_ACT_TANH);
    *P_actor = a.P; *P_critic = c.P;
    return 0;
}

/* One learn() call.  rows [B][sd + ad + sd + 2] = [s | action vector | s' | r | d]; policy_noise [B][ad] = randn_like(actions);
 * expo_target [B][ad], expo_actor [B][ad] = the exponential_() draws of the two gumbel_softmax calls (expo_actor is only read
 * when this call updates the policy).  total_it is self.total_it AFTER the increment (1 on the first call).
 * Returns the critic loss; *actor_loss is set when the policy was updated (else NaN). */
float le_oracle_td3_learn(int sd, int ad, int H, int L, int act, double gamma, double tau, double lr, int policy_delay,
                          float max_action, float policy_std, float policy_std_clip, float gumbel_tau, int gumbel_hard,
                          float* actor, float* actorT, float* c1, float* c1T, float* c2, float* c2T,
                          float* m_a, float* v_a, float* m_c1, float* v_c1, float* m_c2, float* v_c2, int32_t* t_actor, int32_t* t_critic,
                          int total_it, const float* rows, int B, const float* policy_noise, const float* expo_target,
                          const float* expo_actor, float* actor_loss) {
    omlp na, nc; build_mlp(&na, sd, H, L, ad, act); build_mlp(&nc, sd + ad, H, L, 1, act);
    const int ROW = 2 * sd + ad + 2, XI = sd + ad;
    float* acts_a = (float*)malloc(sizeof(float) * na.S);
    float* acts_c1 = (float*)malloc(sizeof(float) * (size_t)B * nc.S);
    float* acts_c2 = (float*)malloc(sizeof(float) * (size_t)B * nc.S);
    float* acts_t = (float*)malloc(sizeof(float) * nc.S);
    float* tq = (float*)malloc(sizeof(float) * B);
    float* g1 = (float*)calloc(nc.P, sizeof(float));
    float* g2 = (float*)calloc(nc.P, sizeof(float));
    float* d = (float*)malloc(sizeof(float) * (nc.S > na.S ? nc.S : na.S));
    float x[LE_ORACLE_MAX_SD + LE_ORACLE_MAX_AD], ysoft[LE_ORACLE_MAX_AD], na_[LE_ORACLE_MAX_AD];
    /* with torch.no_grad(): target actions and target Q (:73-83) */
    for (int b = 0; b < B; ++b) {
        const float* r = rows + (size_t)b * ROW;
        const float* s2 = r + sd + ad;
        const float* out = mlp_fwd_row(&na, actorT, s2, acts_a);
        float logits[LE_ORACLE_MAX_AD];
        for (int k = 0; k < ad; ++k) logits[k] = out[k] * max_action;
        gumbel_softmax_row(logits, expo_target + (size_t)b * ad, ad, gumbel_tau, gumbel_hard, ysoft, na_);
        for (int k = 0; k < ad; ++k) {
            float nz = policy_noise[(size_t)b * ad + k] * policy_std;
            nz = nz < -policy_std_clip ? -policy_std_clip : (nz > policy_std_clip ? policy_std_clip : nz);
            x[sd + k] = na_[k] + nz;
        }
        memcpy(x, s2, sizeof(float) * sd);
        const float q1 = *mlp_fwd_row(&nc, c1T, x, acts_t);
        const float q2 = *mlp_fwd_row(&nc, c2T, x, acts_t);
        const float mn = q1 < q2 ? q1 : q2;                                /* torch.min */
        tq[b] = r[2 * sd + ad] + ((1.f - r[2 * sd + ad + 1]) * (float)gamma) * mn;   /* rewards + (1 - dones) * gamma * target_Q */
    }
    /* critic loss = mse(Q1, target) + mse(Q2, target); one Adam over both critics (:86-95) */
    float l1 = 0.f, l2 = 0.f;
    const float norm = (float)(2.0 / (double)B);
    for (int b = 0; b < B; ++b) {
        const float* r = rows + (size_t)b * ROW;     /* cat(state, action) is the first sd+ad entries of the row */
        const float q1 = *mlp_fwd_row(&nc, c1, r, acts_c1 + (size_t)b * nc.S);
        const float q2 = *mlp_fwd_row(&nc, c2, r, acts_c2 + (size_t)b * nc.S);
        const float e1 = q1 - tq[b], e2 = q2 - tq[b];
        l1 += e1 * e1; l2 += e2 * e2;
        memset(d, 0, sizeof(float) * nc.S); d[nc.l[nc.nl - 1].y_off] = norm * e1;
        mlp_bwd_row(&nc, c1, g1, r, acts_c1 + (size_t)b * nc.S, d, NULL);
        memset(d, 0, sizeof(float) * nc.S); d[nc.l[nc.nl - 1].y_off] = norm * e2;
        mlp_bwd_row(&nc, c2, g2, r, acts_c2 + (size_t)b * nc.S, d, NULL);
    }
    const float critic_loss = l1 / (float)B + l2 / (float)B;
    *t_critic += 1;
    adam_vec(c1, m_c1, v_c1, g1, nc.P, *t_critic, lr, 0.9, 0.999, 1e-8);
    adam_vec(c2, m_c2, v_c2, g2, nc.P, *t_critic, lr, 0.9, 0.999, 1e-8);
    *actor_loss = NAN;
    if (total_it % policy_delay == 0) {   /* delayed policy update (:98-119) */
        float* ga = (float*)calloc(na.P, sizeof(float));
        float* gdummy = (float*)calloc(nc.P, sizeof(float));
        float* da = (float*)malloc(sizeof(float) * na.S);
        float lsum = 0.f;
        for (int b = 0; b < B; ++b) {
            const float* r = rows + (size_t)b * ROW;
            const float* out = mlp_fwd_row(&na, actor, r, acts_a);
            float logits[LE_ORACLE_MAX_AD], ret[LE_ORACLE_MAX_AD], dx[LE_ORACLE_MAX_SD + LE_ORACLE_MAX_AD];
            for (int k = 0; k < ad; ++k) logits[k] = out[k] * max_action;
            gumbel_softmax_row(logits, expo_actor + (size_t)b * ad, ad, gumbel_tau, gumbel_hard, ysoft, ret);
            memcpy(x, r, sizeof(float) * sd);
            for (int k = 0; k < ad; ++k) x[sd + k] = ret[k];
            const float q = *mlp_fwd_row(&nc, c1, x, acts_t);          /* the critic AFTER its Adam step */
            lsum += -q;
            memset(d, 0, sizeof(float) * nc.S); d[nc.l[nc.nl - 1].y_off] = -1.f / (float)B;
            memset(dx, 0, sizeof(dx));
            mlp_bwd_row(&nc, c1, gdummy, x, acts_t, d, dx);             /* only dL/d(action) is used: the critic is not stepped */
            /* straight-through: dL/dy_soft = dL/dret; softmax backward; / tau; * max_action */
            float dot = 0.f;
            for (int k = 0; k < ad; ++k) dot += ysoft[k] * dx[sd + k];
            memset(da, 0, sizeof(float) * na.S);
            for (int k = 0; k < ad; ++k) da[na.l[na.nl - 1].y_off + k] = (ysoft[k] * (dx[sd + k] - dot) / gumbel_tau) * max_action;
            mlp_bwd_row(&na, actor, ga, r, acts_a, da, NULL);
            (void)XI;
        }
        *actor_loss = lsum / (float)B;
        *t_actor += 1;
        adam_vec(actor, m_a, v_a, ga, na.P, *t_actor, lr, 0.9, 0.999, 1e-8);
        const float tf = (float)tau, omt = (float)(1.0 - tau);
        for (int p = 0; p < nc.P; ++p) { c1T[p] = tf * c1[p] + omt * c1T[p]; c2T[p] = tf * c2[p] + omt * c2T[p]; }
        for (int p = 0; p < na.P; ++p) actorT[p] = tf * actor[p] + omt * actorT[p];
        free(ga); free(gdummy); free(da);
    }
    free(acts_a); free(acts_c1); free(acts_c2); free(acts_t); free(tq); free(g1); free(g2); free(d);
    return critic_loss;
}

/* ------------------------------------------------------------------------------------------------ */
/* gym 0.17.3 classic_control (third-party, restated)                             */

void le_oracle_cartpole_step(double st[4], int action, double* reward, int* done) {
    const double gravity = 9.8, masscart = 1.0, masspole = 0.1;
    const double total_mass = masspole + masscart;
    const double length = 0.5;
    const double polemass_length = masspole * length;
    const double force_mag = 10.0, tau = 0.02;
    const double theta_thr = 12 * 2 * M_PI / 360, x_thr = 2.4;
    double x = st[0], x_dot = st[1], theta = st[2], theta_dot = st[3];
    double force = action == 1 ? force_mag : -force_mag;
    double costheta = cos(theta), sintheta = sin(theta);
    double temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) / total_mass;
    double thetaacc = (gravity * sintheta - costheta * temp) /
                      (length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass));
    double xacc = temp - polemass_length * thetaacc * costheta / total_mass;
    x = x + tau * x_dot;
    x_dot = x_dot + tau * xacc;
    theta = theta + tau * theta_dot;
    theta_dot = theta_dot + tau * thetaacc;
    st[0] = x; st[1] = x_dot; st[2] = theta; st[3] = theta_dot;
    *done = (x < -x_thr || x > x_thr || theta < -theta_thr || theta > theta_thr);
    *reward = 1.0; /* the agent loop never steps past done (agents/base_agent.py:128) */
}

static void acrobot_dsdt(const double s[5], double d[5]) {
    const double m1 = 1., m2 = 1., l1 = 1., lc1 = 0.5, lc2 = 0.5, I1 = 1., I2 = 1., g = 9.8, pi = M_PI;
    double a = s[4], theta1 = s[0], theta2 = s[1], dtheta1 = s[2], dtheta2 = s[3];
    double d1 = m1 * (lc1 * lc1) + m2 * (l1 * l1 + lc2 * lc2 + 2 * l1 * lc2 * cos(theta2)) + I1 + I2;
    double d2 = m2 * (lc2 * lc2 + l1 * lc2 * cos(theta2)) + I2;
    double phi2 = m2 * lc2 * g * cos(theta1 + theta2 - pi / 2.);
    double phi1 = -m2 * l1 * lc2 * (dtheta2 * dtheta2) * sin(theta2) - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2) +
                  (m1 * lc1 + m2 * l1) * g * cos(theta1 - pi / 2) + phi2;
    double ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * (dtheta1 * dtheta1) * sin(theta2) - phi2) /
                      (m2 * (lc2 * lc2) + I2 - (d2 * d2) / d1);
    double ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
    d[0] = dtheta1; d[1] = dtheta2; d[2] = ddtheta1; d[3] = ddtheta2; d[4] = 0.;
}

static double wrap_pi(double x) {
    const double m = -M_PI, M = M_PI, diff = M - m;
    while (x > M) x = x - diff;
    while (x < m) x = x + diff;
    return x;
}

void le_oracle_acrobot_step(double st[4], int action, double* reward, int* done) {
    const double dt = .2, dt2 = dt / 2.0;
    double y0[5] = {st[0], st[1], st[2], st[3], (double)(action - 1)}; /* AVAIL_TORQUE = [-1, 0, +1] */
    double k1[5], k2[5], k3[5], k4[5], y[5];
    acrobot_dsdt(y0, k1);
    for (int i = 0; i < 5; ++i) y[i] = y0[i] + dt2 * k1[i];
    acrobot_dsdt(y, k2);
    for (int i = 0; i < 5; ++i) y[i] = y0[i] + dt2 * k2[i];
    acrobot_dsdt(y, k3);
    for (int i = 0; i < 5; ++i) y[i] = y0[i] + dt * k3[i];
    acrobot_dsdt(y, k4);
    for (int i = 0; i < 4; ++i) y[i] = y0[i] + dt / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    y[0] = wrap_pi(y[0]);
    y[1] = wrap_pi(y[1]);
    const double v1 = 4 * M_PI, v2 = 9 * M_PI;
    y[2] = fmin(fmax(y[2], -v1), v1);
    y[3] = fmin(fmax(y[3], -v2), v2);
    for (int i = 0; i < 4; ++i) st[i] = y[i];
    int terminal = (-cos(st[0]) - cos(st[1] + st[0]) > 1.);
    *done = terminal;
    *reward = terminal ? 0. : -1.;
}

void le_oracle_real_obs(int real_env, const double st[4], float* obs) {
    if (real_env == LE_REAL_CARTPOLE) {
        for (int i = 0; i < 4; ++i) obs[i] = (float)st[i];
    } else {
        obs[0] = (float)cos(st[0]); obs[1] = (float)sin(st[0]);
        obs[2] = (float)cos(st[1]); obs[3] = (float)sin(st[1]);
        obs[4] = (float)st[2]; obs[5] = (float)st[3];
    }
}

/* gym step + TimeLimit (elapsed counted by the caller) */
void le_oracle_real_step(int real_env, int max_steps, double st[4], int* elapsed, int action, float* obs, float* reward,
                         float* done) {
    double r; int d;
    if (real_env == LE_REAL_CARTPOLE) le_oracle_cartpole_step(st, action, &r, &d);
    else le_oracle_acrobot_step(st, action, &r, &d);
    *elapsed += 1;
    if (*elapsed >= max_steps) d = 1; /* gym/wrappers/time_limit.py */
    le_oracle_real_obs(real_env, st, obs);
    *reward = (float)r; /* envs/env_wrapper.py:63-65: torch.tensor(..., dtype=float32) */
    *done = d ? 1.f : 0.f;
}

static void real_reset(int real_env, const uint32_t w[4], double st[4]) {
    double half = real_env == LE_REAL_CARTPOLE ? 0.05 : 0.1;
    for (int i = 0; i < 4; ++i) st[i] = -half + (half - (-half)) * ((double)w[i] * (1.0 / 4294967296.0));
}

/* ------------------------------------------------------------------------------------------------ */
/* DDQN.learn (agents/DDQN.py:60-95) on an explicit minibatch of packed rows [s a s' r d]               */

float le_oracle_td_update(const le_lane_cfg* c, float* th, float* thT, float* m, float* v, int32_t* adam_t,
                          const float* rows, int B) {
    if (!is_simple_dqn(c)) return td_update_general(c, th, thT, m, v, adam_t, rows, B);
    const int sd = c->sd, ad = c->ad, H = c->q_hidden, act = c->q_act;
    const int P = le_oracle_q_params(c);
    const int ROW = 2 * sd + 3;
    float* W1 = th; float* b1 = W1 + H * sd; float* W2 = b1 + H; float* b2 = W2 + ad * H;
    float* g = (float*)calloc((size_t)P, sizeof(float));
    float* gW1 = g; float* gb1 = gW1 + H * sd; float* gW2 = gb1 + H; float* gb2 = gW2 + ad * H;
    float* h = (float*)malloc(sizeof(float) * H * 2);
    float* z = h + H;
    float q[LE_ORACLE_MAX_AD], q2[LE_ORACLE_MAX_AD], qT[LE_ORACLE_MAX_AD];
    const float gam = (float)c->gamma;
    const float norm = (float)(2.0 / (double)B); /* mse_loss backward, reduction='mean' */
    double loss_acc = 0.0;
    float loss_f = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* r = rows + (size_t)b * ROW;
        const float* s = r; int a = (int)r[sd]; const float* s2 = r + sd + 1; float rew = r[2 * sd + 1], dn = r[2 * sd + 2];
        mlp_forward(th, sd, H, ad, act, 0.f, s, q, h, z);          /* q_values = model(states)            :79 */
        int astar;
        le_oracle_q_forward(c, th, s2, q2, &astar);               /* next_q_values.max(1)[1]             :80,84 */
        mlp_forward(thT, sd, H, ad, act, 0.f, s2, qT, NULL, NULL); /* model_target(next_states)           :81 */
        float y = rew + gam * qT[astar] * (1.f - dn);             /* expected_q_value                    :85 */
        float delta = q[a] - y;
        loss_f += delta * delta;
        loss_acc += (double)delta * delta;
        float dq = norm * delta;
        gb2[a] += dq;
        for (int j = 0; j < H; ++j) {
            gW2[a * H + j] += dq * h[j];
            float dz = dq * W2[a * H + j] * act_grad(act, 0.f, z[j], h[j]);
            gb1[j] += dz;
            for (int i = 0; i < sd; ++i) gW1[j * sd + i] += dz * s[i];
        }
    }
    (void)loss_acc;
    float loss = loss_f / (float)B;
    /* torch.optim.Adam single-tensor step (torch/optim/adam.py _single_tensor_adam), defaults */
    *adam_t += 1;
    const double b1d = c->beta1, b2d = c->beta2;
    const float w1 = (float)(1.0 - b1d), b2f = (float)b2d, w2 = (float)(1.0 - b2d);
    const double bc1 = 1.0 - pow(b1d, (double)*adam_t), bc2 = 1.0 - pow(b2d, (double)*adam_t);
    const float neg_step = (float)(-(c->lr / bc1));
    const float bc2s = (float)sqrt(bc2);
    const float epsf = (float)c->adam_eps;
    for (int p = 0; p < P; ++p) {
        m[p] = m[p] + w1 * (g[p] - m[p]);   /* exp_avg.lerp_(grad, 1-beta1) */
        v[p] = v[p] * b2f;                  /* exp_avg_sq.mul_(beta2)       */
        v[p] = v[p] + w2 * g[p] * g[p];     /* .addcmul_(grad, grad, value=1-beta2) */
        float denom = sqrtf(v[p]) / bc2s + epsf;
        th[p] = th[p] + neg_step * m[p] / denom; /* param.addcdiv_(exp_avg, denom, value=-step_size) */
    }
    /* Polyak (agents/DDQN.py:93-94) */
    const float tau = (float)c->tau, omt = (float)(1.0 - c->tau);
    for (int p = 0; p < P; ++p) thT[p] = tau * th[p] + omt * thT[p];
    (void)b1; (void)b2;
    free(g); free(h);
    return loss;
}

/* ------------------------------------------------------------------------------------------------ */
/* the lane: BaseAgent.train / test (agents/base_agent.py:64-227) + calc_score (agents/GTN_worker.py:187-221) */

typedef struct {
    const le_lane_cfg* c;
    const float* env_theta;
    uint32_t k0, k1;
    float* th; float* thT; float* m; float* v; int32_t adam_t;
    float* rb; int rb_cap; int rb_ptr; int rb_size; /* utils.py:9-32 */
    int64_t train_steps, learn_iters, test_steps;
    int test_calls;
} lane_t;

static double mean_window(const double* vals, int len, int num, int ignore_last) {
    /* AverageMeter._mean (utils.py:103-105) */
    int lo = len - num - ignore_last; if (lo < 0) lo = 0;
    int hi = len - ignore_last; if (hi < 0) hi = 0;
    double s = 0.0;
    for (int i = lo; i < hi; ++i) s += vals[i];
    return s / ((double)(hi - lo) + 1e-9);
}

/* BaseAgent.test (agents/base_agent.py:155-227): greedy rollouts on the real env. */
static void lane_test(lane_t* L, double* rewards_out) {
    const le_lane_cfg* c = L->c;
    for (int ep = 0; ep < c->test_episodes; ++ep) {
        uint32_t w[4];
        le_oracle_philox((uint32_t)L->test_calls, (uint32_t)ep, LE_P_RESET_TEST, 0, L->k0, L->k1, w);
        double st[4]; real_reset(c->real_env, w, st);
        float obs[LE_ORACLE_MAX_SD]; le_oracle_real_obs(c->real_env, st, obs);
        int elapsed = 0; float ep_rew = 0.f;
        const int K = c->same_action_num > 1 ? c->same_action_num : 1;
        for (int t = 0; t < c->max_steps; t += K) {          /* agents/base_agent.py:193 */
            float q[LE_ORACLE_MAX_AD]; int a;
            le_oracle_q_forward(c, L->th, obs, q, &a);     /* select_test_action agents/DDQN.py:106-110 */
            float r = 0.f, d = 0.f;
            double rsum = 0.0;                               /* EnvWrapper.step real branch envs/env_wrapper.py:56-61 */
            for (int k = 0; k < K; ++k) {
                float rk;
                le_oracle_real_step(c->real_env, c->max_steps, st, &elapsed, a, obs, &rk, &d);
                rsum += (double)rk;
                if (d > 0.5f) break;
            }
            r = (float)rsum;
            ep_rew += r; L->test_steps++;
            if (d > 0.5f) break;
        }
        rewards_out[ep] = (double)ep_rew;
    }
    L->test_calls++;
}

int le_oracle_run_lane(const le_lane_cfg* c, const float* env_theta, uint32_t k0, uint32_t k1, const float* q_init,
                       float* q_final, le_lane_out* out, double* rewards, int32_t* lengths, double* test_rewards,
                       const le_trace* tr) {
    if (c->sd > LE_ORACLE_MAX_SD || c->ad > LE_ORACLE_MAX_AD) return -1;
    const int sd = c->sd, ad = c->ad, P = le_oracle_q_params(c), ROW = 2 * sd + 3;
    lane_t L; memset(&L, 0, sizeof(L));
    L.c = c; L.env_theta = env_theta; L.k0 = k0; L.k1 = k1;
    L.th = (float*)malloc(sizeof(float) * P * 4);
    L.thT = L.th + P; L.m = L.thT + P; L.v = L.m + P;
    if (q_init) memcpy(L.th, q_init, sizeof(float) * P);
    else le_oracle_q_init(c, k0, k1, L.th);
    memcpy(L.thT, L.th, sizeof(float) * P);           /* model_target.load_state_dict  agents/DDQN.py:36 */
    memset(L.m, 0, sizeof(float) * P * 2);
    int64_t max_total = (int64_t)c->train_episodes * c->max_steps;
    if (c->step_budget > 0 && c->step_budget + c->max_steps < max_total) max_total = c->step_budget + c->max_steps;
    L.rb_cap = c->rb_size < max_total ? c->rb_size : (int)max_total;
    if (L.rb_cap < 1) L.rb_cap = 1;
    L.rb = (float*)malloc(sizeof(float) * (size_t)L.rb_cap * ROW);
    float* batch = (float*)malloc(sizeof(float) * (size_t)c->batch_size * ROW);
    double* test_tmp = (double*)malloc(sizeof(double) * (c->test_episodes > 0 ? c->test_episodes : 1));

    double eps = c->eps_init;
    int n_ep = 0, timed_out = 0;
    const int rule_virtual = (!c->use_test_env && c->env_kind == LE_ENV_SE);
    for (int episode = 0; episode < c->train_episodes; ++episode) {
        if (c->step_budget > 0 && L.train_steps >= c->step_budget) { timed_out = 1; break; } /* time_is_up :91-96 */
        /* update_parameters_per_episode (agents/DDQN.py:112-117) */
        if (episode == 0) eps = c->eps_init;
        else { eps *= c->eps_decay; if (eps < c->eps_min) eps = c->eps_min; }
        /* env.reset(): SE -> reset_env.reset() cast to f32 (envs/virtual_env.py:35-41); RN/REAL -> gym reset */
        uint32_t w[4];
        le_oracle_philox((uint32_t)episode, 0, LE_P_RESET_TRAIN, 0, k0, k1, w);
        double st[4]; real_reset(c->real_env, w, st);
        float state[LE_ORACLE_MAX_SD]; le_oracle_real_obs(c->real_env, st, state);
        int elapsed = 0;
        float ep_rew = 0.f; int ep_len = 0;
        const int K = c->same_action_num > 1 ? c->same_action_num : 1;
        for (int t = 0; t < c->max_steps; t += K) {          /* agents/base_agent.py:104 */
            /* select_train_action (agents/DDQN.py:97-104) */
            le_oracle_philox((uint32_t)L.train_steps, 0, LE_P_ACT, 0, k0, k1, w);
            double u = (double)(w[0] >> 8) * (1.0 / 16777216.0);
            int a, explore = (u < eps);
            float qgap = NAN;
            if (explore) a = (int)mulhi32(w[1], (uint32_t)ad);
            else {
                float q[LE_ORACLE_MAX_AD]; le_oracle_q_forward(c, L.th, state, q, &a);
                float second = -3.4e38f;
                for (int k = 0; k < ad; ++k) if (k != a && q[k] > second) second = q[k];
                qgap = (q[a] - second) / fmaxf(fmaxf(fabsf(q[a]), fabsf(second)), 1e-12f);   /* le_trace.qgap */
            }
            /* env.step */
            float ns[LE_ORACLE_MAX_SD], r = 0.f, d = 0.f;
            if (c->env_kind == LE_ENV_SE) {
                /* EnvWrapper.step virtual branch (envs/env_wrapper.py:24-30): same_action_num chained SE steps, fp32 reward sum,
                 * no break on done; done of the last step */
                float cur[LE_ORACLE_MAX_SD]; memcpy(cur, state, sizeof(float) * sd);
                for (int k = 0; k < K; ++k) {
                    float rk;
                    le_oracle_se_step(c, env_theta, cur, a, ns, &rk, &d);
                    r = k == 0 ? rk : r + rk;
                    memcpy(cur, ns, sizeof(float) * sd);
                }
            } else {
                /* real branch (envs/env_wrapper.py:56-61): python-float reward sum, break on done */
                float cur[LE_ORACLE_MAX_SD]; memcpy(cur, state, sizeof(float) * sd);
                double rsum = 0.0;
                for (int k = 0; k < K; ++k) {
                    float rr, rk;
                    le_oracle_real_step(c->real_env, c->max_steps, st, &elapsed, a, ns, &rr, &d);
                    if (c->env_kind == LE_ENV_RN) {
                        /* RewardEnv.step: state/next_state are the fp64 gym states cast to f32 (:78-79) */
                        if (le_oracle_rn_reward(c, env_theta, cur, ns, rr, &rk) != 0) { free(L.th); free(L.rb); free(batch); free(test_tmp); return -2; }
                    } else rk = rr;
                    rsum += (double)rk;
                    memcpy(cur, ns, sizeof(float) * sd);
                    if (d > 0.5f) break;
                }
                r = (float)rsum;
            }
            /* replay_buffer.add (utils.py:24-32) */
            float* row = L.rb + (size_t)L.rb_ptr * ROW;
            memcpy(row, state, sizeof(float) * sd); row[sd] = (float)a; memcpy(row + sd + 1, ns, sizeof(float) * sd);
            row[2 * sd + 1] = r; row[2 * sd + 2] = d;
            L.rb_ptr = (L.rb_ptr + 1) % L.rb_cap;
            L.rb_size = L.rb_size + 1 < L.rb_cap ? L.rb_size + 1 : L.rb_cap;
            memcpy(state, ns, sizeof(float) * sd);
            ep_rew += r; ep_len += K;                      /* episode_length += same_action_num :123 */
            float loss = NAN;
            if (episode >= c->init_episodes) { /* learn (agents/DDQN.py:60-95) */
                for (int b = 0; b < c->batch_size; b += 4) {
                    le_oracle_philox((uint32_t)L.learn_iters, (uint32_t)(b >> 2), LE_P_SAMPLE, 0, k0, k1, w);
                    for (int k = 0; k < 4 && b + k < c->batch_size; ++k) {
                        uint32_t idx = mulhi32(w[k], (uint32_t)L.rb_size); /* np.random.randint(0,size) utils.py:35 */
                        memcpy(batch + (size_t)(b + k) * ROW, L.rb + (size_t)idx * ROW, sizeof(float) * ROW);
                    }
                }
                loss = le_oracle_td_update(c, L.th, L.thT, L.m, L.v, &L.adam_t, batch, c->batch_size);
                L.learn_iters++;
            }
            if (tr && tr->cap > 0 && L.train_steps < tr->cap) {
                int64_t i = L.train_steps;
                tr->action[i] = a; tr->explore[i] = explore; tr->reward[i] = r; tr->done[i] = d; tr->loss[i] = loss;
                if (tr->qgap) tr->qgap[i] = qgap;
                memcpy(tr->next_state + i * sd, ns, sizeof(float) * sd);
            }
            L.train_steps++;
            if (d > 0.5f) break;
        }
        lengths[n_ep] = ep_len;
        if (c->use_test_env) {
            lane_test(&L, test_tmp);
            double s = 0.0; for (int i = 0; i < c->test_episodes; ++i) s += test_tmp[i];
            rewards[n_ep] = s / (double)c->test_episodes; /* statistics.mean */
        } else rewards[n_ep] = (double)ep_rew;
        n_ep++;
        if (episode >= c->init_episodes) { /* env_solved (agents/base_agent.py:49-62) */
            double avg = mean_window(rewards, n_ep, c->early_out_num, 0);
            double avg_last = mean_window(rewards, n_ep, c->early_out_num, c->early_out_num);
            int solved;
            if (rule_virtual)
                solved = (fabs(avg - avg_last) / (fabs(avg_last) + 1e-9) < c->early_out_virtual_diff) &&
                         (episode >= c->init_episodes + c->early_out_num);
            else solved = (avg >= c->solved_reward);
            if (solved) break;
        }
    }
    double score = 0.0;
    if (c->final_test) {
        lane_test(&L, test_rewards);
        for (int i = 0; i < c->test_episodes; ++i) score += test_rewards[i];
        score /= (double)c->test_episodes;
    }
    out->n_episodes = n_ep; out->timed_out = timed_out; out->train_steps = L.train_steps;
    out->learn_iters = L.learn_iters; out->test_steps = L.test_steps; out->score = score;
    if (q_final) memcpy(q_final, L.th, sizeof(float) * P);
    free(L.th); free(L.rb); free(batch); free(test_tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* TD3_discrete_vary lane: BaseAgent.train / test (agents/base_agent.py:64-227) with select_train_action / select_test_action
 * (agents/TD3_discrete_vary.py:155-166) and learn (:62-119).  Streams: oracle/philox.py P_TD3_EXPO / P_TD3_NORMAL. */
void le_oracle_td3_expo(uint32_t k0, uint32_t k1, uint32_t phase, uint32_t c0, uint32_t sub, int n, float* out) {
    for (int i = 0; i < n; i += 4) {
        uint32_t w[4];
        le_oracle_philox(c0, phase, LE_P_TD3_EXPO, (uint32_t)(i >> 2) + (sub << 16), k0, k1, w);
        for (int k = 0; k < 4 && i + k < n; ++k) out[i + k] = (float)(-log(((double)w[k] + 0.5) * (1.0 / 4294967296.0)));
    }
}
void le_oracle_td3_normal(uint32_t k0, uint32_t k1, uint32_t phase, uint32_t c0, uint32_t sub, int n, float* out) {
    for (int i = 0; i < n; i += 4) {
        uint32_t w[4];
        le_oracle_philox(c0, phase, LE_P_TD3_NORMAL, (uint32_t)(i >> 2) + (sub << 16), k0, k1, w);
        double z[4];
        for (int h = 0; h < 2; ++h) {
            const double u1 = ((double)w[2 * h] + 1.0) * (1.0 / 4294967296.0), u2 = (double)w[2 * h + 1] * (1.0 / 4294967296.0);
            const double r = sqrt(-2.0 * log(u1)), t = (2.0 * M_PI) * u2;
            z[2 * h] = r * cos(t); z[2 * h + 1] = r * sin(t);
        }
        for (int k = 0; k < 4 && i + k < n; ++k) out[i + k] = (float)z[k];
    }
}

/* actor(state, temp) + randn(ad) * action_std -> action vector; returns its argmax (first maximum) */
static int td3_act(const le_oracle_td3_cfg* tc, const omlp* na, const float* actor, float* acts, const float* state, float temp,
                   uint32_t k0, uint32_t k1, uint32_t phase, uint32_t c0, uint32_t sub, float* avec) {
    const int ad = tc->base.ad;
    float logits[LE_ORACLE_MAX_AD] = {0}, expo[LE_ORACLE_MAX_AD], nz[LE_ORACLE_MAX_AD], ysoft[LE_ORACLE_MAX_AD], ret[LE_ORACLE_MAX_AD];
    const float* o = mlp_fwd_row(na, actor, state, acts);
    for (int k = 0; k < ad; ++k) logits[k] = o[k] * (float)tc->max_action;
    le_oracle_td3_expo(k0, k1, phase, c0, sub, ad, expo);
    le_oracle_td3_normal(k0, k1, phase, c0, sub, ad, nz);
    gumbel_softmax_row(logits, expo, ad, temp, tc->gumbel_hard, ysoft, ret);
    int best = 0;
    for (int k = 0; k < ad; ++k) { avec[k] = ret[k] + nz[k] * (float)tc->action_std; if (avec[k] > avec[best]) best = k; }
    return best;
}

int le_oracle_run_lane_td3(const le_oracle_td3_cfg* tc, const float* env_theta, uint32_t k0, uint32_t k1, const float* actor_init,
                           const float* c1_init, const float* c2_init, float* actor_final, le_lane_out* out, double* rewards,
                           int32_t* lengths, double* test_rewards, const le_trace* tr) {
    const le_lane_cfg* c = &tc->base;
    if (c->sd > LE_ORACLE_MAX_SD || c->ad > LE_ORACLE_MAX_AD) return -1;
    const int sd = c->sd, ad = c->ad, H = c->q_hidden, L = c->q_layers > 1 ? c->q_layers : 1, B = c->batch_size;
    const int ROW = 2 * sd + ad + 2, K = c->same_action_num > 1 ? c->same_action_num : 1;
    omlp na, nc; build_mlp(&na, sd, H, L, ad, c->q_act); build_mlp(&nc, sd + ad, H, L, 1, c->q_act);
    float* nets = (float*)calloc((size_t)2 * na.P + 4 * nc.P, sizeof(float));
    float *actor = nets, *actorT = actor + na.P, *c1 = actorT + na.P, *c1T = c1 + nc.P, *c2 = c1T + nc.P, *c2T = c2 + nc.P;
    memcpy(actor, actor_init, sizeof(float) * na.P); memcpy(actorT, actor_init, sizeof(float) * na.P);
    memcpy(c1, c1_init, sizeof(float) * nc.P); memcpy(c1T, c1_init, sizeof(float) * nc.P);
    memcpy(c2, c2_init, sizeof(float) * nc.P); memcpy(c2T, c2_init, sizeof(float) * nc.P);
    float* mom = (float*)calloc((size_t)2 * na.P + 4 * nc.P, sizeof(float));
    float *m_a = mom, *v_a = m_a + na.P, *m_c1 = v_a + na.P, *v_c1 = m_c1 + nc.P, *m_c2 = v_c1 + nc.P, *v_c2 = m_c2 + nc.P;
    int32_t t_actor = 0, t_critic = 0;
    int64_t max_total = (int64_t)c->train_episodes * c->max_steps;
    if (c->step_budget > 0 && c->step_budget + c->max_steps < max_total) max_total = c->step_budget + c->max_steps;
    int rb_cap = c->rb_size < max_total ? c->rb_size : (int)max_total; if (rb_cap < 1) rb_cap = 1;
    float* rb = (float*)malloc(sizeof(float) * (size_t)rb_cap * ROW);
    float* batch = (float*)malloc(sizeof(float) * (size_t)B * ROW);
    float* noise = (float*)malloc(sizeof(float) * (size_t)B * ad * 3);
    float* acts_a = (float*)malloc(sizeof(float) * na.S);
    double* test_tmp = (double*)malloc(sizeof(double) * (c->test_episodes > 0 ? c->test_episodes : 1));
    int rb_ptr = 0, rb_size = 0, n_ep = 0, timed_out = 0, total_it = 0, test_calls = 0;
    int64_t train_steps = 0, learn_iters = 0, test_steps = 0;
    const double T0 = tc->gumbel_temp, Tstep = (T0 / 20.0 - T0) / 1999.0;   /* np.linspace(T0, T0/20, 2000) :59 */
    float temp = (float)T0;                                                  /* self.gumbel_temp_annealed */
    const int rule_virtual = (!c->use_test_env && c->env_kind == LE_ENV_SE);
#define TD3_TEST(dst)                                                                                                              \
    do {                                                                                                                           \
        for (int ep_ = 0; ep_ < c->test_episodes; ++ep_) {                                                                         \
            uint32_t w_[4]; le_oracle_philox((uint32_t)test_calls, (uint32_t)ep_, LE_P_RESET_TEST, 0, k0, k1, w_);                 \
            double st_[4]; real_reset(c->real_env, w_, st_);                                                                       \
            float obs_[LE_ORACLE_MAX_SD]; le_oracle_real_obs(c->real_env, st_, obs_);                                              \
            int el_ = 0; float er_ = 0.f;                                                                                          \
            for (int t_ = 0; t_ < c->max_steps; t_ += K) {                                                                         \
                float av_[LE_ORACLE_MAX_AD], r_ = 0.f, d_ = 0.f; double rs_ = 0.0;                                                 \
                const int a_ = td3_act(tc, &na, actor, acts_a, obs_, temp, k0, k1, 1u, ((uint32_t)test_calls << 16) | (uint32_t)(t_ / K), (uint32_t)ep_, av_);                 \
                for (int k_ = 0; k_ < K; ++k_) { float rk_; le_oracle_real_step(c->real_env, c->max_steps, st_, &el_, a_, obs_, &rk_, &d_); \
                                                 rs_ += (double)rk_; if (d_ > 0.5f) break; }                                       \
                r_ = (float)rs_; er_ += r_; test_steps++;                                                                          \
                if (d_ > 0.5f) break;                                                                                              \
            }                                                                                                                      \
            (dst)[ep_] = (double)er_;                                                                                              \
        }                                                                                                                          \
        test_calls++;                                                                                                              \
    } while (0)
    for (int episode = 0; episode < c->train_episodes; ++episode) {
        if (c->step_budget > 0 && train_steps >= c->step_budget) { timed_out = 1; break; }
        uint32_t w[4];
        le_oracle_philox((uint32_t)episode, 0, LE_P_RESET_TRAIN, 0, k0, k1, w);
        double st[4]; real_reset(c->real_env, w, st);
        float state[LE_ORACLE_MAX_SD]; le_oracle_real_obs(c->real_env, st, state);
        int elapsed = 0, ep_len = 0; float ep_rew = 0.f;
        for (int t = 0; t < c->max_steps; t += K) {
            float avec[LE_ORACLE_MAX_AD]; int a;
            if (episode < c->init_episodes) {     /* env.get_random_action() -> one-hot (:156-159) */
                le_oracle_philox((uint32_t)train_steps, 0, LE_P_ACT, 0, k0, k1, w);
                a = (int)mulhi32(w[1], (uint32_t)ad);
                for (int k = 0; k < ad; ++k) avec[k] = k == a ? 1.f : 0.f;
            } else a = td3_act(tc, &na, actor, acts_a, state, temp, k0, k1, 0u, (uint32_t)train_steps, 0u, avec);
            float ns[LE_ORACLE_MAX_SD], r = 0.f, d = 0.f;
            if (c->env_kind == LE_ENV_SE) {
                float cur[LE_ORACLE_MAX_SD]; memcpy(cur, state, sizeof(float) * sd);
                for (int k = 0; k < K; ++k) { float rk; le_oracle_se_step(c, env_theta, cur, a, ns, &rk, &d); r = k == 0 ? rk : r + rk;
                                              memcpy(cur, ns, sizeof(float) * sd); }
            } else {
                float cur[LE_ORACLE_MAX_SD]; memcpy(cur, state, sizeof(float) * sd);
                double rsum = 0.0;
                for (int k = 0; k < K; ++k) {
                    float rr, rk;
                    le_oracle_real_step(c->real_env, c->max_steps, st, &elapsed, a, ns, &rr, &d);
                    if (c->env_kind == LE_ENV_RN) { if (le_oracle_rn_reward(c, env_theta, cur, ns, rr, &rk) != 0) return -2; } else rk = rr;
                    rsum += (double)rk; memcpy(cur, ns, sizeof(float) * sd);
                    if (d > 0.5f) break;
                }
                r = (float)rsum;
            }
            float* row = rb + (size_t)rb_ptr * ROW;      /* replay_buffer.add with the action VECTOR (agents/base_agent.py:72-77,119) */
            memcpy(row, state, sizeof(float) * sd); memcpy(row + sd, avec, sizeof(float) * ad);
            memcpy(row + sd + ad, ns, sizeof(float) * sd); row[2 * sd + ad] = r; row[2 * sd + ad + 1] = d;
            rb_ptr = (rb_ptr + 1) % rb_cap; rb_size = rb_size + 1 < rb_cap ? rb_size + 1 : rb_cap;
            memcpy(state, ns, sizeof(float) * sd);
            ep_rew += r; ep_len += K;
            float loss = NAN;
            if (episode >= c->init_episodes) {
                temp = (float)(total_it >= 1999 ? T0 / 20.0 : (double)total_it * Tstep + T0);   /* gumbel_temp_anneal_steps[total_it] :64-67 */
                total_it += 1;
                for (int b = 0; b < B; b += 4) {
                    le_oracle_philox((uint32_t)learn_iters, (uint32_t)(b >> 2), LE_P_SAMPLE, 0, k0, k1, w);
                    for (int k = 0; k < 4 && b + k < B; ++k)
                        memcpy(batch + (size_t)(b + k) * ROW, rb + (size_t)mulhi32(w[k], (uint32_t)rb_size) * ROW, sizeof(float) * ROW);
                }
                le_oracle_td3_normal(k0, k1, 2u, (uint32_t)learn_iters, 0u, B * ad, noise);
                le_oracle_td3_expo(k0, k1, 2u, (uint32_t)learn_iters, 0u, B * ad, noise + (size_t)B * ad);
                le_oracle_td3_expo(k0, k1, 3u, (uint32_t)learn_iters, 0u, B * ad, noise + (size_t)2 * B * ad);
                float aloss;
                loss = le_oracle_td3_learn(sd, ad, H, L, c->q_act, c->gamma, c->tau, c->lr, tc->policy_delay, (float)tc->max_action,
                                           (float)tc->policy_std, (float)tc->policy_std_clip, temp, tc->gumbel_hard, actor, actorT, c1, c1T,
                                           c2, c2T, m_a, v_a, m_c1, v_c1, m_c2, v_c2, &t_actor, &t_critic, total_it, batch, B, noise,
                                           noise + (size_t)B * ad, noise + (size_t)2 * B * ad, &aloss);
                learn_iters++;
            }
            if (tr && tr->cap > 0 && train_steps < tr->cap) {
                const int64_t i = train_steps;
                tr->action[i] = a; tr->explore[i] = episode < c->init_episodes; tr->reward[i] = r; tr->done[i] = d; tr->loss[i] = loss;
                memcpy(tr->next_state + i * sd, ns, sizeof(float) * sd);
            }
            train_steps++;
            if (d > 0.5f) break;
        }
        lengths[n_ep] = ep_len;
        if (c->use_test_env) {
            TD3_TEST(test_tmp);
            double s = 0.0; for (int i = 0; i < c->test_episodes; ++i) s += test_tmp[i];
            rewards[n_ep] = s / (double)c->test_episodes;
        } else rewards[n_ep] = (double)ep_rew;
        n_ep++;
        if (episode >= c->init_episodes) {
            const double avg = mean_window(rewards, n_ep, c->early_out_num, 0), avg_last = mean_window(rewards, n_ep, c->early_out_num, c->early_out_num);
            const int solved = rule_virtual ? ((fabs(avg - avg_last) / (fabs(avg_last) + 1e-9) < c->early_out_virtual_diff) &&
                                               (episode >= c->init_episodes + c->early_out_num))
                                            : (avg >= c->solved_reward);
            if (solved) break;
        }
    }
    double score = 0.0;
    if (c->final_test) {
        TD3_TEST(test_rewards);
        for (int i = 0; i < c->test_episodes; ++i) score += test_rewards[i];
        score /= (double)c->test_episodes;
    }
#undef TD3_TEST
    out->n_episodes = n_ep; out->timed_out = timed_out; out->train_steps = train_steps; out->learn_iters = learn_iters;
    out->test_steps = test_steps; out->score = score;
    if (actor_final) memcpy(actor_final, actor, sizeof(float) * na.P);
    free(nets); free(mom); free(rb); free(batch); free(noise); free(acts_a); free(test_tmp);
    return 0;
}

/* torch default nn.Linear init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias, drawn from the
 * P_QINIT stream in canonical parameter order (distribution of models/model_utils.py:31,38; the stream is ours). */
void le_oracle_q_init(const le_lane_cfg* c, uint32_t k0, uint32_t k1, float* th) {
    if (!is_simple_dqn(c)) {
        onet n; build_net(c, &n);
        const olayer* ls[8]; int nl = 0;
        for (int i = 0; i < n.nfeat; ++i) ls[nl++] = &n.feat[i];
        if (n.kind == LE_Q_DUELING) { ls[nl++] = &n.val[0]; ls[nl++] = &n.val[1]; ls[nl++] = &n.adv[0]; ls[nl++] = &n.adv[1]; }
        for (int li = 0; li < nl; ++li) {
            const double bnd = 1.0 / sqrt((double)ls[li]->in);
            const int end = ls[li]->b_off + ls[li]->out;
            for (int p = ls[li]->w_off; p < end; ++p) {
                uint32_t w[4];
                le_oracle_philox((uint32_t)(p >> 2), 0, LE_P_QINIT, 0, k0, k1, w);
                th[p] = (float)((2.0 * (((double)w[p & 3] + 0.5) * (1.0 / 4294967296.0)) - 1.0) * bnd);
            }
        }
        return;
    }
    const int sd = c->sd, H = c->q_hidden, P = le_oracle_q_params(c);
    const int n1 = H * sd + H;
    const double bnd1 = 1.0 / sqrt((double)sd), bnd2 = 1.0 / sqrt((double)H);
    for (int p = 0; p < P; p += 4) {
        uint32_t w[4];
        le_oracle_philox((uint32_t)(p >> 2), 0, LE_P_QINIT, 0, k0, k1, w);
        for (int k = 0; k < 4 && p + k < P; ++k) {
            double u = ((double)w[k] + 0.5) * (1.0 / 4294967296.0);
            th[p + k] = (float)((2.0 * u - 1.0) * ((p + k) < n1 ? bnd1 : bnd2));
        }
    }
}

/* n independent lanes on n_threads pthreads (dynamic lane queue).  The reference runs one single-threaded
 * process per member (experiments/GTN_Worker.py:8).  cfgs: n_cfg == 1 shares the configuration. */
typedef struct {
    const le_lane_cfg* cfgs; int n_cfg; const float* env_theta; int P_env; const int32_t* env_index;
    const uint32_t* keys; const float* q_init; float* q_final; int n_lanes; le_lane_out* out;
    double* rewards; int32_t* lengths; double* test_rewards;
    int next; int rc; pthread_mutex_t mu;
} lanes_job_t;

static void* lanes_worker(void* arg) {
    lanes_job_t* J = (lanes_job_t*)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int i = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (i >= J->n_lanes) break;
        const le_lane_cfg* c = &J->cfgs[J->n_cfg == 1 ? 0 : i];
        const le_lane_cfg* c0 = &J->cfgs[0];
        int Pq = le_oracle_q_params(c0); /* strides follow cfg 0 (max shapes) */
        const float* th = J->env_theta + (size_t)(J->env_index ? J->env_index[i] : 0) * J->P_env;
        int r = le_oracle_run_lane(c, th, J->keys[2 * i], J->keys[2 * i + 1], J->q_init ? J->q_init + (size_t)i * Pq : NULL,
                                   J->q_final ? J->q_final + (size_t)i * Pq : NULL, &J->out[i],
                                   J->rewards + (size_t)i * c0->train_episodes, J->lengths + (size_t)i * c0->train_episodes,
                                   J->test_rewards + (size_t)i * c0->test_episodes, NULL);
        if (r != 0) { pthread_mutex_lock(&J->mu); J->rc = r; pthread_mutex_unlock(&J->mu); }
    }
    return NULL;
}

int le_oracle_run_lanes(const le_lane_cfg* cfgs, int n_cfg, const float* env_theta, int P_env, const int32_t* env_index,
                        const uint32_t* keys, const float* q_init, float* q_final, int n_lanes, le_lane_out* out,
                        double* rewards, int32_t* lengths, double* test_rewards, int n_threads) {
    lanes_job_t J = {cfgs, n_cfg, env_theta, P_env, env_index, keys, q_init, q_final, n_lanes, out,
                     rewards, lengths, test_rewards, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t tid[256];
    for (int t = 1; t < n_threads; ++t) pthread_create(&tid[t], NULL, lanes_worker, &J);
    lanes_worker(&J);
    for (int t = 1; t < n_threads; ++t) pthread_join(tid[t], NULL);
    return J.rc;
}

int le_oracle_sizeof_cfg(void) { return (int)sizeof(le_lane_cfg); }
