// mlp_tc.cuh -- the gt_mpc value term on Blackwell tensor cores (tcgen05 + TMEM), CTA-wide.
//
// Replaces the CasADi SX expansion of the value network inside the NLP cost
// (reference model.py:53-67 `get_casadi_mlp`, used at mpc.py:367-369): network 6 -> 128 -> tanh ->
// 128 -> tanh -> 1 (game_theoretic_NN/configs/sc{1,2,4,5,8}_config.yaml: hidden_size 128,
// num_layers 2).  Only s_N and v_N are decision variables (mpc.py:326-337), so one evaluation
// needs the value plus first and second forward tangents in (s_N, v_N): six rows per problem.
//
// One CTA evaluates up to 256 problems (one per thread) together:
//   layer 1 (K = 6)      CUDA cores, fp32: a1 = W1eff xN + b1eff with W1eff = W1 Wn folded on the host
//   layer 2 (128 x 128)  tcgen05.mma kind::f16, M = 128 problems, N = 64, K = 16 per instruction,
//                        bf16x3 split of both operands (6 products, fp32 accumulation in TMEM):
//                        a single bf16 / tf32 pass misses the 1e-4 cost tolerance
//   layer 3 (128 -> 1)   CUDA cores while draining TMEM (tcgen05.ld 32x32b: TMEM lane == problem)
// The weight splits (96 KB, laid out on the host in the UMMA no-swizzle K-major core-matrix
// order) are staged into shared memory once per kernel by a bulk TMA copy
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace igt {

constexpr int MLP_H = 128;                 // hidden width handled on tensor cores
constexpr int TC_TILE_BYTES = MLP_H * MLP_H * 2;      // one bf16 128x128 operand tile (32 KB)
constexpr int TC_SMEM_W = 0;                          // 3 weight splits
constexpr int TC_SMEM_A = 3 * TC_TILE_BYTES;          // 3 activation splits
constexpr int TC_SMEM_MISC = 6 * TC_TILE_BYTES;       // barriers, tmem address, small fp32 tables
constexpr int TC_MISC_BYTES = 64 + (MLP_H * 6 + MLP_H * 3 + 8) * 4;
constexpr int TC_SMEM_BYTES = TC_SMEM_MISC + TC_MISC_BYTES + 1024;   // + slack for 1024 B alignment

// device-resident description of a prepared network (built by igt_set_mlp)
struct MlpTcWeights {
    const uint8_t *w2_splits;   // [3][32 KB] bf16 hi / mid / lo of W2 in core-matrix order
    const float *w1eff;         // [128][6]  W1 * Wn
    const float *b1eff;         // [128]     b1 - W1eff * mu_f
    const float *b2;            // [128]
    const float *w3;            // [128]
    float b3, sigma_t, mu_t;
    int enabled;
};

// byte offset of element (row, k) of a 128x128 bf16 K-major operand in the no-swizzle canonical
// layout: 8x8 core matrices (8 rows x 16 B) stored contiguously; cores adjacent in K are 128 B
// apart (LBO), groups of 8 rows are 16 * 128 B apart (SBO)
__host__ __device__ inline int tc_elem_offset(int row, int k)
{
    return ((row >> 3) * 16 + (k >> 3)) * 128 + (row & 7) * 16 + (k & 7) * 2;
}
constexpr uint32_t TC_LBO = 128, TC_SBO = 2048;

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr)
{   // UMMA shared-memory matrix descriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46,
    // layout type 0 (no swizzle)
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(TC_LBO >> 4) << 16;
    d |= (uint64_t)(TC_SBO >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor, kind::f16: D = F32, A = B = BF16, both K-major, M = 128, N = 64
__device__ __forceinline__ uint32_t tc_idesc_m128_n64()
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// per-CTA context living in dynamic shared memory
struct MlpTcCtx {
    uint8_t *smem;          // 1024 B aligned base
    uint64_t *bar_w, *bar_mma;
    uint32_t *tmem_slot;
    float *w1eff, *b1eff, *b2, *w3;
    uint32_t tmem_base;
    uint32_t mma_phase;
    float b3, sigma_t, mu_t;
};

// Called once by every thread of the CTA (blockDim.x == 256) before the first evaluation.
__device__ __forceinline__ void mlp_tc_setup(MlpTcCtx &c, uint8_t *dyn_smem, const MlpTcWeights &wt)
{
    uint8_t *base = (uint8_t *)(((uintptr_t)dyn_smem + 1023) & ~(uintptr_t)1023);
    c.smem = base;
    uint8_t *misc = base + TC_SMEM_MISC;
    c.bar_w = (uint64_t *)misc; c.bar_mma = (uint64_t *)(misc + 8); c.tmem_slot = (uint32_t *)(misc + 16);
    c.w1eff = (float *)(misc + 64); c.b1eff = c.w1eff + MLP_H * 6; c.b2 = c.b1eff + MLP_H; c.w3 = c.b2 + MLP_H;
    c.b3 = wt.b3; c.sigma_t = wt.sigma_t; c.mu_t = wt.mu_t; c.mma_phase = 0;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(c.bar_w, 1);
        mbar_init(c.bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < MLP_H * 6; i += blockDim.x) c.w1eff[i] = wt.w1eff[i];
    for (int i = tid; i < MLP_H; i += blockDim.x) { c.b1eff[i] = wt.b1eff[i]; c.b2[i] = wt.b2[i]; c.w3[i] = wt.w3[i]; }
    if (tid < 32) {   // warp 0 allocates all 512 TMEM columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(c.tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {   // TMA bulk copy of the three weight splits, completion on bar_w
        const uint32_t bytes = 3 * TC_TILE_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(c.bar_w)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(base + TC_SMEM_W)), "l"(wt.w2_splits), "r"(bytes), "r"(smem_u32(c.bar_w)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    c.tmem_base = *c.tmem_slot;
    mbar_wait(c.bar_w, 0);
}

__device__ __forceinline__ void mlp_tc_teardown(MlpTcCtx &c)
{
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(c.tmem_base) : "memory");
}

// Value + tangents of the terminal value term for this thread's problem.  CTA-uniform call:
// every thread of the 256-thread CTA must call it; threads without work pass valid = false.
// out[6] = (V, dV/ds, dV/dv, d2V/dss, d2V/dsv, d2V/dvv), already scaled by sigma_t (+ mu_t on V).
__device__ __forceinline__ void mlp_tc_eval(MlpTcCtx &c, bool valid, float sN, float vN, const float *ctx, float *out)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = tid >> 7;                      // two groups of 128 problems share smem A and TMEM in turn
    const int row = tid & 127;                     // M row == TMEM lane of this thread's problem
    float xN[6] = { ctx[0], ctx[1], ctx[2], sN - ctx[0], vN - ctx[1], ctx[3] - ctx[2] };
    if (!valid) { for (int j = 0; j < 6; j++) xN[j] = 0.f; }
    float acc[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    const uint32_t idesc = tc_idesc_m128_n64();
    uint8_t *sA = c.smem + TC_SMEM_A, *sW = c.smem + TC_SMEM_W;
    // the two groups of 128 rows take turns; a group without work is skipped (CTA-uniform)
    const int g_lo = __syncthreads_or(valid && grp == 0) ? 0 : 1, g_hi = __syncthreads_or(valid && grp == 1) ? 2 : 1;
    for (int g = g_lo; g < g_hi; g++) {
        for (int nh = 0; nh < 2; nh++) {
            for (int r = 0; r < 6; r++) {
                // ---- this group's threads write row r of the layer-1 output (bf16 hi/mid/lo) ----
                // y = tanh(a1) is computed once (first round) and parked in the 128 TMEM columns the
                // accumulators leave free (this thread's TMEM lane); later rounds reload it.
                if (grp == g) {
                    const uint32_t ycache = c.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 384u;
#pragma unroll 1
                    for (int ob = 0; ob < MLP_H; ob += 16) {
                        uint32_t yv[16];
                        if (nh == 0 && r == 0) {
#pragma unroll
                            for (int i = 0; i < 16; i++) {
                                const float *w = c.w1eff + (ob + i) * 6;
                                float a = c.b1eff[ob + i];
#pragma unroll
                                for (int j = 0; j < 6; j++) a += w[j] * xN[j];
                                yv[i] = __float_as_uint(tanhf(a));
                            }
                            asm volatile(
                                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                                :: "r"(ycache + (uint32_t)ob), "r"(yv[0]), "r"(yv[1]), "r"(yv[2]), "r"(yv[3]), "r"(yv[4]), "r"(yv[5]),
                                   "r"(yv[6]), "r"(yv[7]), "r"(yv[8]), "r"(yv[9]), "r"(yv[10]), "r"(yv[11]), "r"(yv[12]), "r"(yv[13]),
                                   "r"(yv[14]), "r"(yv[15]) : "memory");
                        } else {
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                : "=r"(yv[0]), "=r"(yv[1]), "=r"(yv[2]), "=r"(yv[3]), "=r"(yv[4]), "=r"(yv[5]), "=r"(yv[6]), "=r"(yv[7]),
                                  "=r"(yv[8]), "=r"(yv[9]), "=r"(yv[10]), "=r"(yv[11]), "=r"(yv[12]), "=r"(yv[13]), "=r"(yv[14]), "=r"(yv[15])
                                : "r"(ycache + (uint32_t)ob));
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        }
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const int o = ob + i;
                            const float y = __uint_as_float(yv[i]), d1 = 1.f - y * y, d2 = -2.f * y * d1;
                            const float ts = c.w1eff[o * 6 + 3], tv = c.w1eff[o * 6 + 4];
                            float val = (r == 0) ? y : (r == 1) ? d1 * ts : (r == 2) ? d1 * tv : (r == 3) ? d2 * ts * ts
                                      : (r == 4) ? d2 * ts * tv : d2 * tv * tv;
                            __nv_bfloat16 hi = __float2bfloat16_rn(val);
                            float r1 = val - __bfloat162float(hi);
                            __nv_bfloat16 mid = __float2bfloat16_rn(r1);
                            __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
                            const int off = tc_elem_offset(row, o);
                            *(__nv_bfloat16 *)(sA + off) = hi;
                            *(__nv_bfloat16 *)(sA + TC_TILE_BYTES + off) = mid;
                            *(__nv_bfloat16 *)(sA + 2 * TC_TILE_BYTES + off) = lo;
                        }
                    }
                    if (nh == 0 && r == 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = c.tmem_base + (uint32_t)(r * 64);
                    const uint32_t aBase = smem_u32(sA), wBase = smem_u32(sW) + (uint32_t)nh * (64 / 8) * TC_SBO;
                    // products hi*hi, hi*mid, mid*hi, hi*lo, lo*hi, mid*mid
                    const int pa[6] = { 0, 0, 1, 0, 2, 1 }, pb[6] = { 0, 1, 0, 2, 0, 1 };
                    uint32_t accum = 0;
#pragma unroll 1
                    for (int pr = 0; pr < 6; pr++)
#pragma unroll 1
                        for (int ks = 0; ks < MLP_H / 16; ks++) {
                            uint64_t ad = tc_smem_desc(aBase + pa[pr] * TC_TILE_BYTES + ks * 2 * TC_LBO);
                            uint64_t bd = tc_smem_desc(wBase + pb[pr] * TC_TILE_BYTES + ks * 2 * TC_LBO);
                            tc_mma(d_tmem, ad, bd, idesc, accum);
                            accum = 1;
                        }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 :: "r"(smem_u32(c.bar_mma)) : "memory");
                }
                mbar_wait(c.bar_mma, c.mma_phase);
                c.mma_phase ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            // ---- drain: 6 rows x 64 columns of this N half; tanh, tangent rules, layer-3 dot ----
            if (grp == g) {
                const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
#pragma unroll 1
                for (int cb = 0; cb < 64; cb += 16) {
                    uint32_t v[6][16];
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        const uint32_t taddr = c.tmem_base + lane_base + (uint32_t)(r * 64 + cb);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                            : "=r"(v[r][0]), "=r"(v[r][1]), "=r"(v[r][2]), "=r"(v[r][3]), "=r"(v[r][4]), "=r"(v[r][5]),
                              "=r"(v[r][6]), "=r"(v[r][7]), "=r"(v[r][8]), "=r"(v[r][9]), "=r"(v[r][10]), "=r"(v[r][11]),
                              "=r"(v[r][12]), "=r"(v[r][13]), "=r"(v[r][14]), "=r"(v[r][15])
                            : "r"(taddr));
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int o = nh * 64 + cb + i;
                        float a0 = __uint_as_float(v[0][i]) + c.b2[o];
                        float a1 = __uint_as_float(v[1][i]), a2 = __uint_as_float(v[2][i]);
                        float a3 = __uint_as_float(v[3][i]), a4 = __uint_as_float(v[4][i]), a5 = __uint_as_float(v[5][i]);
                        float y = tanhf(a0), d1 = 1.f - y * y, d2 = -2.f * y * d1, w3 = c.w3[o];
                        acc[0] += w3 * y;
                        acc[1] += w3 * (d1 * a1);
                        acc[2] += w3 * (d1 * a2);
                        acc[3] += w3 * (d1 * a3 + d2 * a1 * a1);
                        acc[4] += w3 * (d1 * a4 + d2 * a1 * a2);
                        acc[5] += w3 * (d1 * a5 + d2 * a2 * a2);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();            // TMEM and smem A are reused by the next N half / group
        }
    }
    out[0] = (acc[0] + c.b3) * c.sigma_t + c.mu_t;
#pragma unroll
    for (int r = 1; r < 6; r++) out[r] = acc[r] * c.sigma_t;
    (void)lane;
}
#endif   // __CUDACC__

}  // namespace igt
