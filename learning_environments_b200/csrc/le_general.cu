// le_general.cu — host side of the CTA-per-lane general Q-network kernels (le_general.cuh).
#include "le_general.cuh"
#include "le_general_api.h"

namespace le {

static int general_occupancy(int sd) {
    int nb = 0;
    // dynamic shared memory = the operand parts of the tcgen05 GEMM (le_tc.cuh); two CTAs per SM also bound the TMEM use
    // (2 x 128 of the SM's 512 accumulator columns)
    if (sd == 4) {
        cudaFuncSetAttribute(general_loop_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, general_loop_kernel<4, 2>, kGThreads, tc::kSmemBytes);
    } else {
        cudaFuncSetAttribute(general_loop_kernel<6, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, general_loop_kernel<6, 3>, kGThreads, tc::kSmemBytes);
    }
    if (nb > 2) nb = 2;
    return nb < 1 ? 1 : nb;
}

int general_plan(const le_lane_cfg* c, int n_lanes, int ring_cap, int sms, GeneralPlan* gp) {
    GNet n;
    gnet_build(c, &n);
    if (n.sum_out <= 0 || n.P <= 0 || c->q_hidden < 1 || (c->q_kind == LE_Q_DUELING && c->q_feature_dim < 1)) {
        le_set_error("bad Q-network shape (hidden=%d layers=%d feature_dim=%d)", c->q_hidden, c->q_layers, c->q_feature_dim);
        return LE_EINVAL;
    }
    const int occ = general_occupancy(c->sd);
    gp->grid = n_lanes < sms * occ ? n_lanes : sms * occ;
    gp->bmax = c->batch_size > 64 ? c->batch_size : 64;
    int64_t offs[18];
    gp->slot_floats = gslot_floats(n, ring_cap, RowLayout<4>::ROWF * (c->sd == 4) + RowLayout<6>::ROWF * (c->sd == 6), gp->bmax, offs);
    gp->params = n.P;
    return LE_OK;
}

cudaError_t general_launch(const le_lane_cfg* c, const RunParams& rp, float* slots, const GeneralPlan& gp, cudaStream_t st) {
    GRunParams G;
    G.rp = rp;
    gnet_build(c, &G.net);
    G.slots = slots;
    G.slot_stride = gp.slot_floats;
    G.bmax = gp.bmax;
    cudaError_t e = c->sd == 4 ? cudaFuncSetAttribute(general_loop_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes)
                               : cudaFuncSetAttribute(general_loop_kernel<6, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
    if (e != cudaSuccess) return e;
    if (c->sd == 4) general_loop_kernel<4, 2><<<gp.grid, kGThreads, tc::kSmemBytes, st>>>(G);
    else general_loop_kernel<6, 3><<<gp.grid, kGThreads, tc::kSmemBytes, st>>>(G);
    return cudaGetLastError();
}

int general_td_update(const le_lane_cfg* cfg, const le_lane_cfg* cfg_dev, float* th, float* thT, float* m, float* v, int32_t* t, int n_lanes,
                      const float* rows, float* loss, cudaStream_t st) {
    GNet n;
    gnet_build(cfg, &n);
    int64_t offs[18];
    const int64_t stride = gslot_floats(n, 0, 0, cfg->batch_size, offs);
    float* scratch = nullptr;
    LE_CUDA_CHECK(cudaMallocAsync((void**)&scratch, (size_t)stride * 4 * n_lanes, st));
    if (cfg->sd == 4) {
        LE_CUDA_CHECK(cudaFuncSetAttribute(general_td_update_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes));
        general_td_update_kernel<4, 2><<<n_lanes, kGThreads, tc::kSmemBytes, st>>>(cfg_dev, n, th, thT, m, v, t, n.P, rows, cfg->batch_size, loss, scratch, stride);
    } else {
        LE_CUDA_CHECK(cudaFuncSetAttribute(general_td_update_kernel<6, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes));
        general_td_update_kernel<6, 3><<<n_lanes, kGThreads, tc::kSmemBytes, st>>>(cfg_dev, n, th, thT, m, v, t, n.P, rows, cfg->batch_size, loss, scratch, stride);
    }
    LE_CUDA_CHECK(cudaGetLastError());
    LE_CUDA_CHECK(cudaFreeAsync(scratch, st));
    return LE_OK;
}

int general_qnet_forward(const le_lane_cfg* cfg, const float* q_theta, int n_rows, const float* state, float* q_out, int32_t* argmax,
                         cudaStream_t st) {
    GNet n;
    gnet_build(cfg, &n);
    const int64_t stride = (n.sum_out + n.ad + 7) / 4 * 4;
    float* scratch = nullptr;
    LE_CUDA_CHECK(cudaMallocAsync((void**)&scratch, (size_t)stride * 4 * n_rows, st));
    if (cfg->sd == 4) general_qnet_forward_kernel<4, 2><<<n_rows, kGThreads, 0, st>>>(n, q_theta, n.P, state, q_out, argmax, scratch, stride);
    else general_qnet_forward_kernel<6, 3><<<n_rows, kGThreads, 0, st>>>(n, q_theta, n.P, state, q_out, argmax, scratch, stride);
    LE_CUDA_CHECK(cudaGetLastError());
    LE_CUDA_CHECK(cudaFreeAsync(scratch, st));
    return LE_OK;
}

int general_q_params(const le_lane_cfg* c) {
    GNet n;
    gnet_build(c, &n);
    return n.P;
}

}  // namespace le
