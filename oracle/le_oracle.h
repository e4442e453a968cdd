/* le_oracle.h — TEST INFRASTRUCTURE ONLY: prototypes of the CPU restatement (see le_oracle.c). */
#ifndef LE_ORACLE_H
#define LE_ORACLE_H
#include <stdint.h>

#include "../include/le_b200.h" /* le_lane_cfg, le_lane_out, le_trace, LE_* ids */

#define LE_ORACLE_MAX_SD 8
#define LE_ORACLE_MAX_AD 4
#define LE_ORACLE_MAX_IN (LE_ORACLE_MAX_SD + LE_ORACLE_MAX_AD)

/* Philox stream purposes (oracle/philox.py) */
#define LE_P_ACT 1
#define LE_P_SAMPLE 2
#define LE_P_RESET_TRAIN 3
#define LE_P_RESET_TEST 4
#define LE_P_QINIT 5
#define LE_P_NOISE 6
#define LE_P_TD3_EXPO 8
#define LE_P_TD3_NORMAL 9

void le_oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]);
int le_oracle_mlp_params(int in, int H, int out);
int le_oracle_se_params(const le_lane_cfg* c);
int le_oracle_rn_params(const le_lane_cfg* c);
int le_oracle_q_params(const le_lane_cfg* c);
void le_oracle_se_step(const le_lane_cfg* c, const float* theta, const float* state, int action, float* next_state,
                       float* reward, float* done);
int le_oracle_rn_reward(const le_lane_cfg* c, const float* theta, const float* s, const float* s2, float real_reward,
                        float* out);
void le_oracle_q_forward(const le_lane_cfg* c, const float* q_theta, const float* state, float* q, int* argmax);
void le_oracle_cartpole_step(double st[4], int action, double* reward, int* done);
void le_oracle_acrobot_step(double st[4], int action, double* reward, int* done);
void le_oracle_real_obs(int real_env, const double st[4], float* obs);
void le_oracle_real_step(int real_env, int max_steps, double st[4], int* elapsed, int action, float* obs, float* reward,
                         float* done);
float le_oracle_td_update(const le_lane_cfg* c, float* th, float* thT, float* m, float* v, int32_t* adam_t,
                          const float* rows, int B);
void le_oracle_q_init(const le_lane_cfg* c, uint32_t k0, uint32_t k1, float* th);
int le_oracle_run_lane(const le_lane_cfg* c, const float* env_theta, uint32_t k0, uint32_t k1, const float* q_init,
                       float* q_final, le_lane_out* out, double* rewards, int32_t* lengths, double* test_rewards,
                       const le_trace* tr);
int le_oracle_run_lanes(const le_lane_cfg* cfgs, int n_cfg, const float* env_theta, int P_env, const int32_t* env_index,
                        const uint32_t* keys, const float* q_init, float* q_final, int n_lanes, le_lane_out* out,
                        double* rewards, int32_t* lengths, double* test_rewards, int n_threads);
int le_oracle_sizeof_cfg(void);
/* TD3_discrete_vary.learn restatement (agents/TD3_discrete_vary.py:62-119): see le_oracle.c */
int le_oracle_td3_params(int sd, int ad, int H, int L, int* P_actor, int* P_critic);
float le_oracle_td3_learn(int sd, int ad, int H, int L, int act, double gamma, double tau, double lr, int policy_delay,
                          float max_action, float policy_std, float policy_std_clip, float gumbel_tau, int gumbel_hard,
                          float* actor, float* actorT, float* c1, float* c1T, float* c2, float* c2T,
                          float* m_a, float* v_a, float* m_c1, float* v_c1, float* m_c2, float* v_c2, int32_t* t_actor, int32_t* t_critic,
                          int total_it, const float* rows, int B, const float* policy_noise, const float* expo_target,
                          const float* expo_actor, float* actor_loss);

/* TD3_discrete_vary inner loop (BaseAgent.train/test with the TD3 act/learn): groundwork, oracle only */
typedef struct le_oracle_td3_cfg {
    le_lane_cfg base;              /* env / loop / Adam fields; q_hidden, q_layers, q_act = shape of the actor and critic MLPs */
    int32_t policy_delay, gumbel_hard;
    double action_std, policy_std, policy_std_clip, gumbel_temp, max_action;
} le_oracle_td3_cfg;
void le_oracle_td3_expo(uint32_t k0, uint32_t k1, uint32_t phase, uint32_t c0, uint32_t sub, int n, float* out);
void le_oracle_td3_normal(uint32_t k0, uint32_t k1, uint32_t phase, uint32_t c0, uint32_t sub, int n, float* out);
int le_oracle_run_lane_td3(const le_oracle_td3_cfg* tc, const float* env_theta, uint32_t k0, uint32_t k1, const float* actor_init,
                           const float* c1_init, const float* c2_init, float* actor_final, le_lane_out* out, double* rewards,
                           int32_t* lengths, double* test_rewards, const le_trace* tr);

#endif
