// mlp_coop.cuh -- the gt_mpc value term (mpc.py:326-354, :367-369; model.py:14-67) evaluated exactly
// (the solver's own precision, fp64 on the parity path) by the whole CTA: one WARP per evaluation.
//
// Value + first and second forward tangents in (s_N, v_N): six rows per evaluation through every layer.  The
// per-thread version (mlp_eval_thread) streams each thread's activations through HBM and every thread reads the
// whole weight matrices; here a warp keeps its six activation rows in shared memory, lane l owns outputs l, l + 32,
// l + 64, l + 96 of a layer (24 independent accumulators), and the weights are read transposed ([in][out]) so that
// the 32 lanes read 32 consecutive words.  Same operation order per output as mlp_eval_thread: bit-identical
// results (tests/hostsim runs that one on the CPU).  Any network with layer widths <= MLP_COOP_W.
#pragma once
#include <cstdint>

namespace igt {

constexpr int MLP_COOP_W = 128;                       // widest layer handled
constexpr int COOP_ITEM = 8, COOP_ROWS = 6;
template <typename T>
constexpr int mlp_coop_smem_bytes(int block) { return (int)sizeof(T) * (block * (COOP_ITEM + COOP_ROWS) + (block / 32) * 2 * COOP_ROWS * MLP_COOP_W); }

#ifdef __CUDACC__
// CTA-uniform call (like mlp_tc_eval): every thread passes its own request (valid = false: none) and gets
// out[6] = (V, dV/ds, dV/dv, d2V/dss, d2V/dsv, d2V/dvv), already scaled by sigma_t (+ mu_t on V).
template <typename T>
__device__ __noinline__ void mlp_coop_eval(const DevParams<T> &P, uint8_t *smem, bool valid, T sN, T vN, const T *ctx, T *out)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x, nw = nthr >> 5;
    T *items = reinterpret_cast<T *>(smem);                       // [nthr][8]: sN, vN, ctx[4], valid
    T *res = items + nthr * COOP_ITEM;                            // [nthr][6]
    T *ha = res + nthr * COOP_ROWS + warp * (2 * COOP_ROWS * MLP_COOP_W), *hb = ha + COOP_ROWS * MLP_COOP_W;
    constexpr int W = MLP_COOP_W;
    items[tid * COOP_ITEM + 0] = sN; items[tid * COOP_ITEM + 1] = vN;
#pragma unroll
    for (int i = 0; i < 4; i++) items[tid * COOP_ITEM + 2 + i] = ctx[i];
    items[tid * COOP_ITEM + 6] = valid ? T(1) : T(0);
    __syncthreads();
    for (int it = warp; it < nthr; it += nw) {
        const T *q = items + it * COOP_ITEM;
        if (q[6] == T(0)) continue;                               // warp-uniform
        T *hi = ha, *ho = hb;
        if (lane < 6) {   // whitened feature x = Wn ((s_tv, v_tv, e_tv, s_N - s_tv, v_N - v_tv, e_ego - e_tv) - mu_f), mpc.py:326-337
            const T xN[6] = { q[2], q[3], q[4], q[0] - q[2], q[1] - q[3], q[5] - q[4] };
            T acc = T(0);
            for (int j = 0; j < 6; j++) acc += P.Wn[lane * 6 + j] * (xN[j] - P.mu_f[j]);
            hi[0 * W + lane] = acc;
            hi[1 * W + lane] = P.Wn[lane * 6 + 3];
            hi[2 * W + lane] = P.Wn[lane * 6 + 4];
            hi[3 * W + lane] = hi[4 * W + lane] = hi[5 * W + lane] = T(0);
        }
        __syncwarp();
        for (int l = 0; l < P.n_layers; l++) {
            const int nin = P.dims[l], nout = P.dims[l + 1];
            const T *Wt = P.Wt[l], *b = P.b[l];
            if (l < P.n_layers - 1) {
                T acc[4][6];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int o = lane + 32 * j;
                    acc[j][0] = o < nout ? b[o] : T(0);
#pragma unroll
                    for (int r = 1; r < 6; r++) acc[j][r] = T(0);
                }
                // eight inputs per pass: 32 weight loads (L1 / L2) and 48 broadcast shared-memory loads are issued before
                // their 192 FMAs, so one memory round trip is exposed per pass, not per input
                constexpr int UI = 8;
                for (int i0 = 0; i0 < nin; i0 += UI) {
                    T h[UI][6], w[UI][4];
#pragma unroll
                    for (int ii = 0; ii < UI; ii++)
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            w[ii][j] = (i0 + ii < nin && lane + 32 * j < nout) ? Wt[(long)(i0 + ii) * nout + lane + 32 * j] : T(0);
#pragma unroll
                    for (int ii = 0; ii < UI; ii++)
#pragma unroll
                        for (int r = 0; r < 6; r++) h[ii][r] = (i0 + ii < nin) ? hi[r * W + i0 + ii] : T(0);
#pragma unroll
                    for (int ii = 0; ii < UI; ii++)                // input order within a pass: same sums as mlp_eval_thread
#pragma unroll
                        for (int j = 0; j < 4; j++)
#pragma unroll
                            for (int r = 0; r < 6; r++) acc[j][r] += w[ii][j] * h[ii][r];
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int o = lane + 32 * j;
                    if (o < nout) {
                        const T y = tanh(acc[j][0]), d1 = T(1) - y * y, d2 = T(-2) * y * d1, ts = acc[j][1], tv = acc[j][2];
                        ho[0 * W + o] = y;
                        ho[1 * W + o] = d1 * ts;
                        ho[2 * W + o] = d1 * tv;
                        ho[3 * W + o] = d1 * acc[j][3] + d2 * ts * ts;
                        ho[4 * W + o] = d1 * acc[j][4] + d2 * ts * tv;
                        ho[5 * W + o] = d1 * acc[j][5] + d2 * tv * tv;
                    }
                }
            } else if (lane < 6) {                                // output layer (nout = 1): lane r sums row r in input order
                T acc = lane == 0 ? b[0] : T(0);
                for (int i = 0; i < nin; i++) acc += Wt[i] * hi[lane * W + i];
                res[it * COOP_ROWS + lane] = lane == 0 ? acc * P.sigma_t + P.mu_t : acc * P.sigma_t;
            }
            __syncwarp();
            T *t = hi; hi = ho; ho = t;
        }
    }
    __syncthreads();
    if (valid) {
#pragma unroll
        for (int r = 0; r < 6; r++) out[r] = res[tid * COOP_ROWS + r];
    }
    __syncthreads();
}
#endif

}  // namespace igt
