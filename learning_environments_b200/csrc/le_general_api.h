// le_general_api.h — entry points of le_general.cu used by le_api.cu
#pragma once
#include "le_inner_loop.cuh"

namespace le {
struct GeneralPlan { int grid, bmax, params; int64_t slot_floats; };
int general_plan(const le_lane_cfg* c, int n_lanes, int ring_cap, int sms, GeneralPlan* gp);
cudaError_t general_launch(const le_lane_cfg* c, const RunParams& rp, float* slots, const GeneralPlan& gp, cudaStream_t st);
int general_td_update(const le_lane_cfg* cfg, const le_lane_cfg* cfg_dev, float* th, float* thT, float* m, float* v, int32_t* t, int n_lanes,
                      const float* rows, float* loss, cudaStream_t st);
int general_qnet_forward(const le_lane_cfg* cfg, const float* q_theta, int n_rows, const float* state, float* q_out, int32_t* argmax,
                         cudaStream_t st);
int general_q_params(const le_lane_cfg* c);
}  // namespace le
