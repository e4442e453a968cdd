// le_general.cu — host side of the CTA-per-lane general Q-network kernels (le_general.cuh).
#include <cstdio>
#include <cstdlib>
#include "le_general.cuh"
#include "le_general_api.h"

namespace le {

// LE_TC=1 in the environment selects the tensor-core instantiations of the CTA-per-lane kernels (dense hidden x hidden layers
// on tcgen05 / TMEM, le_tc.cuh).  Default: the FFMA instantiations — measured faster at the reference's layer sizes (60 .. 128
// wide, 128 .. 193 rows): a TMEM-allocating kernel is held to one CTA per SM and the per-GEMM operand split / staging is not
// amortised by 128^3 contractions (profiles/r02_general_tc.txt).
bool general_use_tc() {
    const char* e = getenv("LE_TC");
    return e && e[0] == '1';
}

template <int SD, int AD, bool TC>
static int occupancy_of(int* nb) {
    const size_t dyn = TC ? tc::kSmemBytes : 0;
    cudaError_t e = cudaSuccess;
    if (TC) e = cudaFuncSetAttribute(general_loop_kernel<SD, AD, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, general_loop_kernel<SD, AD, TC>, kGThreads, dyn);
    if (getenv("LE_DEBUG")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, general_loop_kernel<SD, AD, TC>);
        fprintf(stderr, "[le] general kernel<%d,%d,tc=%d> occupancy: %d CTAs/SM (%s; regs %d, static smem %zu B, dynamic %zu B, local %zu B)\n", SD, AD,
                (int)TC, *nb, cudaGetErrorString(e), fa.numRegs, fa.sharedSizeBytes, dyn, fa.localSizeBytes);
    }
    return e == cudaSuccess ? 0 : -1;
}

static int general_occupancy(int sd) {
    int nb = 0;
    const bool tcp = general_use_tc();
    if (sd == 4) { if (tcp) occupancy_of<4, 2, true>(&nb); else occupancy_of<4, 2, false>(&nb); }
    else { if (tcp) occupancy_of<6, 3, true>(&nb); else occupancy_of<6, 3, false>(&nb); }
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

template <int SD, int AD, bool TC>
static cudaError_t launch_loop(const GRunParams& G, int grid, cudaStream_t st) {
    const size_t dyn = TC ? tc::kSmemBytes : 0;
    if (TC) {
        cudaError_t e = cudaFuncSetAttribute(general_loop_kernel<SD, AD, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
    }
    general_loop_kernel<SD, AD, TC><<<grid, kGThreads, dyn, st>>>(G);
    return cudaGetLastError();
}

cudaError_t general_launch(const le_lane_cfg* c, const RunParams& rp, float* slots, const GeneralPlan& gp, cudaStream_t st) {
    GRunParams G;
    G.rp = rp;
    gnet_build(c, &G.net);
    G.slots = slots;
    G.slot_stride = gp.slot_floats;
    G.bmax = gp.bmax;
    const bool tcp = general_use_tc();
    if (c->sd == 4) return tcp ? launch_loop<4, 2, true>(G, gp.grid, st) : launch_loop<4, 2, false>(G, gp.grid, st);
    return tcp ? launch_loop<6, 3, true>(G, gp.grid, st) : launch_loop<6, 3, false>(G, gp.grid, st);
}

template <int SD, int AD, bool TC>
static cudaError_t launch_td(const le_lane_cfg* cfg_dev, const GNet& n, float* th, float* thT, float* m, float* v, int32_t* t, int n_lanes,
                             const float* rows, int B, float* loss, float* scratch, int64_t stride, cudaStream_t st) {
    const size_t dyn = TC ? tc::kSmemBytes : 0;
    if (TC) {
        cudaError_t e = cudaFuncSetAttribute(general_td_update_kernel<SD, AD, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
    }
    general_td_update_kernel<SD, AD, TC><<<n_lanes, kGThreads, dyn, st>>>(cfg_dev, n, th, thT, m, v, t, n.P, rows, B, loss, scratch, stride);
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
    const bool tcp = general_use_tc();
    const int B = cfg->batch_size;
    cudaError_t le;
    if (cfg->sd == 4) le = tcp ? launch_td<4, 2, true>(cfg_dev, n, th, thT, m, v, t, n_lanes, rows, B, loss, scratch, stride, st)
                               : launch_td<4, 2, false>(cfg_dev, n, th, thT, m, v, t, n_lanes, rows, B, loss, scratch, stride, st);
    else le = tcp ? launch_td<6, 3, true>(cfg_dev, n, th, thT, m, v, t, n_lanes, rows, B, loss, scratch, stride, st)
                  : launch_td<6, 3, false>(cfg_dev, n, th, thT, m, v, t, n_lanes, rows, B, loss, scratch, stride, st);
    LE_CUDA_CHECK(le);
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
