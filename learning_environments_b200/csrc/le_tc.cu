// le_tc.cu — unit entry point of the tcgen05 / TMEM dense-layer GEMM (le_tc.cuh): one CTA runs one strided GEMM with the
// contract of the general kernel's g_gemm, so the GPU tests can pin the tensor-core path (descriptors, operand staging,
// 3xTF32 accuracy, epilogue) against float64 numpy for the three operand forms of a dense layer (X W^T, dZ W, dZ^T X).
#include "le_common.cuh"
#include "le_lane.cuh"
#include "le_tc.cuh"

namespace le {

struct TcAct {
    __device__ __forceinline__ float operator()(int act, float slope, float z) const {
        if (act == 1) return tanh_one(z);
        if (act == 2) return fmaxf(z, slope * z);
        return z;
    }
};

__global__ void __launch_bounds__(tc::kThreads) tc_gemm_kernel(const float* A, int a_si, int a_sl, const float* B, int b_sl, int b_sj, float* C,
                                                                int c_si, int c_sj, int I, int J, int L, const float* bias, int act, float slope,
                                                                int accumulate) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    tc::Ctx ctx = tc::ctx_create(tc_smem);
    tc::gemm_3xtf32(ctx, A, a_si, a_sl, B, b_sl, b_sj, C, c_si, c_sj, I, J, L, bias, act, slope, accumulate != 0, TcAct());
    tc::ctx_destroy(ctx);
}

}  // namespace le

extern "C" int le_tc_gemm(const float* A_dev, int a_si, int a_sl, const float* B_dev, int b_sl, int b_sj, float* C_dev, int c_si, int c_sj, int I,
                          int J, int L, const float* bias_dev, int act, float slope, int accumulate, void* stream) {
    if (!A_dev || !B_dev || !C_dev || I < 1 || J < 1 || L < 1 || act < 0 || act > 2) { le_set_error("le_tc_gemm: bad arguments"); return LE_EINVAL; }
    LE_CUDA_CHECK(cudaFuncSetAttribute(le::tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)le::tc::kSmemBytes));
    le::tc_gemm_kernel<<<1, le::tc::kThreads, le::tc::kSmemBytes, (cudaStream_t)stream>>>(A_dev, a_si, a_sl, B_dev, b_sl, b_sj, C_dev, c_si, c_sj, I, J, L,
                                                                                         bias_dev, act, slope, accumulate);
    LE_CUDA_CHECK(cudaGetLastError());
    return LE_OK;
}
