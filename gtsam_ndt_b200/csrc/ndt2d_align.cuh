// The align machinery shared by the kernels that run SPEC 5 on the device (k_align, k_align_block in ndt2d_kernels.cu;
// k_pairs_fused in ndt2d_pairs.cu): the per-warp LM state, one evaluation into it, THE Levenberg-Marquardt loop and the
// result write-back. Reference file:line: none (the mount is /root/reference/README.md:1 only); arithmetic follows SPEC.md.
#pragma once
#include "ndt2d_device.cuh"

#ifndef NDT2D_EVAL_INLINE
#define NDT2D_EVAL_INLINE 1 // 1: k_align inlines the evaluation into the LM loop (r2g: 18.70 vs 18.17 M matches/s - no spills at
                            // 80 registers and both gathers of a step issue back to back); 0: a separate function, the round-1
                            // form. The fused pairs kernel always calls it as a function: inlined, its code outgrows the
                            // instruction cache (r2h: 9.7 vs 10.8 M pairs/s).
#endif
#ifndef NDT2D_PIPE
#define NDT2D_PIPE 0      // align kernel: 1 = register software pipeline (K = 1)
#endif

namespace ndt2d {

// SPEC 5, one pyramid level. Every lane carries the same f64 state (the butterfly reduction gives
// all lanes identical sums), so the control flow is warp-uniform and needs no broadcast.
// where a warp reads its scan from: its padded shared-memory slot or global memory
struct ScanView {
    const float2 *pts;
    int n;
};

// Per-warp LM state. Every lane computes the same f64 values (the butterfly gives all lanes identical sums), so one
// copy per warp in shared memory is the whole solver state: nothing but a few ints stays live in registers
// across the evaluation call (register spills of warp-uniform doubles would cost 32 lanes x 8 B of local memory each).
struct WarpState {
    double v[10];   // the last accepted evaluation (SPEC 5's E)
    double t[10];   // the trial evaluation
    double p[3];    // the current pose
    double pn[3];   // the trial pose
    double lambda;
    int count, tcount;
};

// One SPEC 4 evaluation at the trial pose ws->pn, result to ws->t / ws->tcount; lane t stores sum t.
template <int OV, bool STAGED, int TABLE>
__device__ __forceinline__ void eval_to_smem_body(const LevelDev *L, const float2 *pts, int n, WarpState *ws)
{
    const int lane = threadIdx.x & 31;
    const double *pose = ws->pn;
    Eval E;
    eval_warp<OV, true, STAGED, (STAGED && OV == 0 && TABLE == TABLE_DENSE) ? NDT2D_PIPE : 0, true, TABLE>(*L, pts, n, pose_for_level(pose[0], pose[1], pose[2], *L), lane, E);
    __syncwarp();
    ws->t[E.slot] = E.v[0]; // lanes holding the same sum store the same bits
    if (lane == 0) ws->tcount = E.count;
    __syncwarp();
}
// ... as a function of its own (the point loop then gets a register allocation independent of the f64 solver)
template <int OV, bool STAGED, int TABLE>
__device__ __noinline__ void eval_to_smem_call(const LevelDev *L, const float2 *pts, int n, WarpState *ws)
{
    eval_to_smem_body<OV, STAGED, TABLE>(L, pts, n, ws);
}
template <int OV, bool STAGED, int TABLE>
__device__ __forceinline__ void eval_to_smem(const LevelDev *L, const float2 *pts, int n, WarpState *ws)
{
    if (NDT2D_EVAL_INLINE && TABLE != TABLE_SHASH) eval_to_smem_body<OV, STAGED, TABLE>(L, pts, n, ws);
    else eval_to_smem_call<OV, STAGED, TABLE>(L, pts, n, ws);
}

// The threads that share one WarpState: a warp (k_align) or a whole block (k_align_block)
struct WarpScope {
    static __device__ __forceinline__ void sync() { __syncwarp(); }
    static __device__ __forceinline__ int id() { return threadIdx.x & 31; }
};
struct BlockScope {
    static __device__ __forceinline__ void sync() { __syncthreads(); }
    static __device__ __forceinline__ int id() { return threadIdx.x; }
};

// SPEC 5, one pyramid level, starting from and finishing in ws->p: THE Levenberg-Marquardt loop, shared by the
// warp-per-scan and the block-per-scan kernels. Every thread of the scope computes the same f64 values from the shared
// state, so the control flow is uniform. eval() evaluates the trial pose ws->pn into ws->t / ws->tcount and ends with a
// Scope::sync(). There is ONE call site: the first evaluation (at ws->p) runs as a trial at pn = p that is accepted
// unconditionally, so an inlined evaluation exists once in the kernel (round 2: the second copy had the worse register
// allocation - ten accumulator moves per 64 points - and it was the one that ran 11 times out of 12).
template <class Scope, class EvalFn>
__device__ __forceinline__ int lm_level(const ndt2d_params &P, int n, WarpState *ws, int &evals_total, EvalFn eval)
{
    const int id = Scope::id();
    if (id == 0) ws->lambda = P.lambda_init;
    if (id < 3) ws->pn[id] = ws->p[id];
    Scope::sync();
    int evals = 0, status = NDT2D_MAX_ITERATIONS;
    bool small = false;
    for (;;) {
        eval();
        evals += 1;
        double lambda = ws->lambda;
        const bool first = evals == 1;
        const bool better = first || ws->t[0] > ws->v[0];
        Scope::sync();
        if (better) {
            if (id < 10) ws->v[id] = ws->t[id];
            if (id == 10) ws->count = ws->tcount;
            if (id >= 11 && id < 14) ws->p[id - 11] = ws->pn[id - 11];
            if (!first) {
                lambda = fmax(lambda / P.lambda_down, P.lambda_min);
                if (id == 14) ws->lambda = lambda;
            }
            Scope::sync();
            if (first) {
                if (n == 0 || ws->count == 0) { status = NDT2D_NO_OVERLAP; break; }
            } else if (small) { status = NDT2D_CONVERGED; break; }
        } else {
            if (small) { status = NDT2D_CONVERGED; break; }
            lambda = lambda * P.lambda_up;
            if (id == 14) ws->lambda = lambda;
            Scope::sync();
            if (lambda > P.lambda_max) { status = NDT2D_STALLED; break; }
        }
        if (evals >= P.max_iterations) break;
        double d[3];
        bool stalled = false;
        {
            double g[3] = {ws->v[1], ws->v[2], ws->v[3]};
            double H6[6] = {ws->v[4], ws->v[5], ws->v[6], ws->v[7], ws->v[8], ws->v[9]};
            while (!solve3(g, H6, lambda, d)) {
                lambda = lambda * P.lambda_fail_up;
                if (lambda > P.lambda_max) { stalled = true; break; }
            }
        }
        if (stalled) { status = NDT2D_STALLED; break; }
        double n2 = d[0] * d[0] + d[1] * d[1]; // squared translation step: no sqrt unless the step is clamped
        if (n2 > P.max_step_trans * P.max_step_trans) {
            double sc = P.max_step_trans / sqrt(n2);
            d[0] *= sc; d[1] *= sc; d[2] *= sc; n2 = P.max_step_trans * P.max_step_trans;
        }
        if (fabs(d[2]) > P.max_step_rot) {
            double sc = P.max_step_rot / fabs(d[2]);
            d[0] *= sc; d[1] *= sc; d[2] *= sc; n2 = n2 * (sc * sc);
        }
        small = (n2 < P.eps_trans * P.eps_trans) && (fabs(d[2]) < P.eps_rot);
        Scope::sync();
        if (id == 0) {
            ws->pn[0] = ws->p[0] + d[0];
            ws->pn[1] = ws->p[1] + d[1];
            ws->pn[2] = ws->p[2] + d[2];
            ws->lambda = lambda;
        }
        Scope::sync();
    }
    evals_total += evals;
    return status;
}

// what align returns (SPEC 5): the state after the finest level, theta wrapped
__device__ __forceinline__ void write_result(const WarpState &E, int evals, int status, ndt2d_result *r)
{
    const double TWO_PI = 6.283185307179586476925286766559;
    r->pose[0] = E.p[0];
    r->pose[1] = E.p[1];
    r->pose[2] = E.p[2] - TWO_PI * rint(E.p[2] / TWO_PI);
    r->score = E.v[0];
    r->grad[0] = E.v[1]; r->grad[1] = E.v[2]; r->grad[2] = E.v[3];
    r->hessian[0] = E.v[4]; r->hessian[1] = E.v[5]; r->hessian[2] = E.v[6];
    r->hessian[3] = E.v[5]; r->hessian[4] = E.v[7]; r->hessian[5] = E.v[8];
    r->hessian[6] = E.v[6]; r->hessian[7] = E.v[8]; r->hessian[8] = E.v[9];
    r->iterations = evals;
    r->status = status;
    r->count = E.count;
    r->reserved = 0;
}

template <int OV, bool STAGED, int TABLE>
__device__ __forceinline__ int align_level(const LevelDev *L, const ndt2d_params &P, const ScanView &v, WarpState *ws,
                                           int &evals_total)
{
    return lm_level<WarpScope>(P, v.n, ws, evals_total, [&]() { eval_to_smem<OV, STAGED, TABLE>(L, v.pts, v.n, ws); });
}


} // namespace ndt2d
