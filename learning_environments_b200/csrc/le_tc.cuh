// le_tc.cuh — 5th-generation tensor-core path (tcgen05 + TMEM, sm_100a) for the dense layers of the CTA-per-lane kernels.
//
// The only genuinely dense contractions of the hot path are the hidden x hidden layers of Critic_DuelingDQN
// (models/actor_critic.py:94-122: 60x60 with B = 193 rows for the CartPole yaml, 128x128 with B = 128 rows for
// default_config_acrobot.yaml:61-80) and of two-hidden-layer Critic_DQN / TD3 nets: forward  Y = X W^T, input gradient
// dX = dZ W and weight gradient dW = dZ^T X.  They run here as
//
//     D[128 x N] (fp32, TMEM)  +=  A[128 x 8] * B[N x 8]^T      tcgen05.mma.cta_group::1.kind::tf32, issued by ONE thread
//
// with the "3xTF32" error-compensated split that keeps fp32-level accuracy (the parity budget is 1e-5 on TD losses):
// every fp32 operand x is staged in shared memory as hi = tf32(x) and lo = tf32(x - hi), and each k-step issues
// hi*lo + lo*hi + hi*hi into the same accumulator (the dropped lo*lo term is 2^-22 relative).  Operands travel
// global -> registers (split) -> shared memory in the canonical no-swizzle K-major UMMA layout
//
//     [k/4][row][k%4]   core matrix = 8 rows x 16 B,   SBO = 128 B (next 8 rows),   LBO = (128 + 1) * 16 B (next 4 k)
//
// for BOTH operands: sources that are contiguous along m/n instead of k (dZ^T and X in dW = dZ^T X, W in dX = dZ W) are
// transposed on the way into shared memory (kind::tf32 accepts MN-major operands only in a 128B-swizzled layout; a register
// pass is needed for the hi/lo split anyway).  The odd 16-byte pad of LBO makes the vector stores of k-contiguous sources and
// the scalar stores of the transposing path bank-conflict free.  X W^T, dZ W and dZ^T X all use ONE staging routine.  Completion
// is tracked with tcgen05.commit on an mbarrier; the epilogue reads the accumulator with tcgen05.ld (32 lanes x 32
// columns per warp), adds the bias, applies the activation and stores (or accumulates into) the row-major result.
//
// Descriptor layouts follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction descriptor" tables (the same
// bit positions as cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS headers; nothing of CUTLASS is compiled in).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace le {
namespace tc {

constexpr int kThreads = 256;
constexpr int kTM = 128;                 // rows of the A tile = TMEM lanes (M of the instruction)
constexpr int kTN = 128;                 // columns of the D tile (N of the instruction <= kTN, multiple of 16)
constexpr int kKC = 32;                  // k elements per staged chunk = 4 MMA k-steps of 8 (tf32: 32 bytes per k-step)
constexpr int kTmemCols = 128;           // fp32 accumulator columns (power of two >= 32)
constexpr uint32_t kLboK = (kTM + 1) * 16, kSboK = 128;          // K-major part:  [kKC/4][129] float4
constexpr uint32_t kPartBytes = (kKC / 4) * kLboK;               // 16512 B
constexpr uint32_t kSmemBytes = 4 * kPartBytes + 16;             // A_hi, A_lo, B_hi, B_lo + mbarrier + tmem address

struct Ctx {
    uint32_t smem;        // shared-memory address (u32) of the 4 operand parts; 16-byte aligned
    uint32_t bar;         // mbarrier (8 bytes) behind the parts
    uint32_t tmem;        // TMEM base address of the accumulator (lane 0, column c0)
    uint32_t phase;       // parity of the next mbarrier completion (kept identical in all threads)
};

__device__ __forceinline__ uint32_t cvt_tf32(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return u; }

// shared memory matrix descriptor, no swizzle: start address, leading / stride byte offsets (16-byte units), version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor for kind::tf32: D = F32, A = B = TF32, major bits, N >> 3, M >> 4
__device__ __forceinline__ uint32_t instr_desc(int n) {   // both operands K-major (major bits 15 / 16 = 0)
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tLE_TC_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                 "@P1 bra LE_TC_DONE;\n\tbra LE_TC_WAIT;\n\tLE_TC_DONE:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One warp allocates kTmemCols TMEM columns for the CTA (address -> shared memory), another call frees them.
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"((uint32_t)kTmemCols) : "memory");
}
// 32 lanes x 32 columns of the accumulator -> 32 registers per thread (thread t <-> TMEM lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// CTA set-up / tear-down (all kThreads threads call both).  `smem` = kSmemBytes of 16-byte aligned shared memory.
__device__ __forceinline__ Ctx ctx_create(void* smem) {
    Ctx c;
    c.smem = (uint32_t)__cvta_generic_to_shared(smem);
    c.bar = c.smem + 4 * kPartBytes;
    c.phase = 0;
    const uint32_t tslot = c.bar + 8;
    if (threadIdx.x == 0) mbar_init(c.bar, 1);
    if (threadIdx.x < 32) tmem_alloc(tslot);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c.tmem) : "r"(tslot) : "memory");
    return c;
}
__device__ __forceinline__ void ctx_destroy(const Ctx& c) {
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_free(c.tmem);
}

// Stage one operand chunk (rows [r0, r0 + 128) x k [k0, k0 + kKC) of `src`, element (row, k) at src[row * s_row + k * s_k]) as
// hi / lo tf32 parts in the K-major layout; rows >= n_rows and k >= n_k are zero.
__device__ __forceinline__ void stage_operand(const float* __restrict__ src, int64_t s_row, int64_t s_k, int r0, int n_rows, int k0, int n_k,
                                              uint32_t s_hi, uint32_t s_lo) {
    const int tid = threadIdx.x;
    auto put = [&](uint32_t off, float x) {
        const uint32_t hi = cvt_tf32(x);
        const uint32_t lo = cvt_tf32(x - __uint_as_float(hi));
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(s_hi + off), "r"(hi) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(s_lo + off), "r"(lo) : "memory");
    };
    auto put4 = [&](uint32_t off, float4 x) {
        uint32_t h[4], l[4];
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { h[e] = cvt_tf32(xv[e]); l[e] = cvt_tf32(xv[e] - __uint_as_float(h[e])); }
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(s_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(s_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
    };
    if (s_k == 1) {   // ---- k-contiguous source
        const float* base = src + (int64_t)r0 * s_row + k0;
        const bool vec = ((s_row & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
        if (vec) {   // item (row m, k quad kq): a warp reads 4 rows x 128 contiguous bytes, stores 32 distinct 16-byte slots
#pragma unroll
            for (int it = 0; it < kTM * (kKC / 4) / kThreads; ++it) {
                const int id = it * kThreads + tid, kq = id % (kKC / 4), m = id / (kKC / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < n_rows && 4 * kq < n_k) {
                    const float* p = base + (int64_t)m * s_row + 4 * kq;
                    if (4 * kq + 4 <= n_k) v = __ldcg(reinterpret_cast<const float4*>(p));
                    else { v.x = __ldcg(p); if (4 * kq + 1 < n_k) v.y = __ldcg(p + 1); if (4 * kq + 2 < n_k) v.z = __ldcg(p + 2); }
                }
                put4((uint32_t)(kq * (int)kLboK + m * 16), v);
            }
        } else {     // item (row m, k): a warp reads one row x 32 contiguous floats
#pragma unroll 4
            for (int it = 0; it < kTM * kKC / kThreads; ++it) {
                const int id = it * kThreads + tid, k = id % kKC, m = id / kKC;
                const float v = (m < n_rows && k < n_k) ? __ldcg(base + (int64_t)m * s_row + k) : 0.f;
                put((uint32_t)((k >> 2) * (int)kLboK + m * 16 + (k & 3) * 4), v);
            }
        }
    } else {          // ---- row-contiguous (or generally strided) source: transposed on the way in
        const float* base = src + (int64_t)k0 * s_k + (int64_t)r0 * s_row;
        const bool vec = (s_row == 1) && ((s_k & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
        if (vec) {   // lane <-> k (the 32 k of the chunk), one row quad per warp pass: 32 scalar stores hit 32 distinct banks
#pragma unroll
            for (int it = 0; it < kKC * (kTM / 4) / kThreads; ++it) {
                const int id = it * kThreads + tid, k = id % kKC, m4 = id / kKC;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < n_k && 4 * m4 < n_rows) {
                    const float* p = base + (int64_t)k * s_k + 4 * m4;
                    if (4 * m4 + 4 <= n_rows) v = __ldcg(reinterpret_cast<const float4*>(p));
                    else { v.x = __ldcg(p); if (4 * m4 + 1 < n_rows) v.y = __ldcg(p + 1); if (4 * m4 + 2 < n_rows) v.z = __ldcg(p + 2); }
                }
                const uint32_t off = (uint32_t)((k >> 2) * (int)kLboK + 4 * m4 * 16 + (k & 3) * 4);
                put(off, v.x); put(off + 16, v.y); put(off + 32, v.z); put(off + 48, v.w);
            }
        } else {     // item (k, row m), lanes along the rows
#pragma unroll 4
            for (int it = 0; it < kTM * kKC / kThreads; ++it) {
                const int id = it * kThreads + tid, m = id % kTM, k = id / kTM;
                const float v = (m < n_rows && k < n_k) ? __ldcg(base + (int64_t)m * s_row + (int64_t)k * s_k) : 0.f;
                put((uint32_t)((k >> 2) * (int)kLboK + m * 16 + (k & 3) * 4), v);
            }
        }
    }
}

// C[i*c_si + j*c_sj] (+)= sum_l A[i*a_si + l*a_sl] * B[l*b_sl + j*b_sj]  (+ bias[j], activation)  — whole CTA (kThreads threads),
// the same contract as g_gemm (le_general.cuh) on the tensor cores.  ActFn(act, slope, v) applies the layer activation.
template <typename ActFn>
__device__ __noinline__ void gemm_3xtf32(Ctx& c, const float* __restrict__ A, int a_si, int a_sl, const float* __restrict__ B, int b_sl, int b_sj,
                                         float* __restrict__ C, int c_si, int c_sj, int I, int J, int L, const float* __restrict__ bias,
                                         int act, float slope, bool accumulate, ActFn act_fn) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sAh = c.smem, sAl = c.smem + kPartBytes, sBh = c.smem + 2 * kPartBytes, sBl = c.smem + 3 * kPartBytes;
    for (int i0 = 0; i0 < I; i0 += kTM) {
        for (int j0 = 0; j0 < J; j0 += kTN) {
            const int nj = min(kTN, J - j0), n_inst = (nj + 15) & ~15;
            for (int l0 = 0; l0 < L; l0 += kKC) {
                const int nk = min(kKC, L - l0);
                stage_operand(A, a_si, a_sl, i0, min(kTM, I - i0), l0, nk, sAh, sAl);
                stage_operand(B, b_sj, b_sl, j0, nj, l0, nk, sBh, sBl);
                fence_async_smem();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
                __syncthreads();
                if (tid == 0) {
                    fence_after_sync();
                    const uint32_t idesc = instr_desc(n_inst);
                    const int steps = (nk + 7) >> 3;
                    for (int s = 0; s < steps; ++s) {
                        const uint32_t ko = (uint32_t)(2 * s) * kLboK;          // one k-step = 8 k = two 16-byte k-groups
                        const uint64_t ah = smem_desc(sAh + ko, kLboK, kSboK), al = smem_desc(sAl + ko, kLboK, kSboK);
                        const uint64_t bh = smem_desc(sBh + ko, kLboK, kSboK), bl = smem_desc(sBl + ko, kLboK, kSboK);
                        mma_tf32(c.tmem, ah, bl, idesc, (l0 > 0 || s > 0) ? 1u : 0u);   // small terms first
                        mma_tf32(c.tmem, al, bh, idesc, 1u);
                        mma_tf32(c.tmem, ah, bh, idesc, 1u);
                    }
                    mma_commit(c.bar);       // arrives when every MMA issued so far has completed (implies fence::before_thread_sync)
                }
                mbar_wait(c.bar, c.phase);   // operands consumed: the parts may be restaged, the accumulator read
                c.phase ^= 1;
            }
            // ---- epilogue of tile (i0, j0): warp w reads TMEM lanes 32 (w % 4) .. +31, columns 64 (w / 4) .. +63
            fence_after_sync();
            const int row = i0 + 32 * (warp & 3) + lane;
#pragma unroll 1
            for (int cb = 0; cb < 2; ++cb) {
                const int col0 = 64 * (warp >> 2) + 32 * cb;
                if (col0 >= n_inst) continue;            // warp-uniform
                float v[32];
                tmem_ld32(c.tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0, v);
                if (row < I) {
                    float* crow = C + (int64_t)row * c_si + (int64_t)(j0 + col0) * c_sj;
                    if (accumulate) {
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (col0 + q < nj) v[q] += __ldcg(crow + (int64_t)q * c_sj);
                    }
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        if (col0 + q < nj) {
                            float x = v[q];
                            if (bias) x += __ldcg(bias + j0 + col0 + q);
                            __stcg(crow + (int64_t)q * c_sj, act_fn(act, slope, x));
                        }
                    }
                }
            }
            fence_before_sync();
            __syncthreads();                 // the accumulator is free for the next tile; C is visible to the CTA
        }
    }
}

}  // namespace tc
}  // namespace le
