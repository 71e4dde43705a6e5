// PROBE (never shipped, compiled only with -DNDT2D_PROBE_HALF): one point per lane and step, ten f32 accumulators per lane - the inner
// loop that a two-warps-per-scan layout would run. Used to measure what that layout could reach before
// building its LM coordination: bench.py --workload newton with this build evaluates every pose with the one-point loop.
#pragma once
#include "ndt2d_device.cuh"
namespace ndt2d {
struct Partials1 {
    u64 s12, s45, s68;
    float s0, s3, s7, s9;
};
template <bool FULL>
__device__ __forceinline__ void step1(const float4 *__restrict__ cells, const LatticePk &G, const PosePk &P, unsigned sa, Partials1 &S, int &cnt)
{
    float2 a;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a.x), "=f"(a.y) : "r"(sa));
    PointPk A;
    rotate_point<true>(P, a.x, a.y, A);
    const unsigned bA = locate_base(P, a.x, a.y, G, A.df);
    const Cell4 c = load_cell(cells, bA);
    A.XY = local_xy(G, A.df, 0);
    const u64 q = sub2(A.XY, c.mu);
    const u64 u = fma2(c.B0, bc(lo32(q)), mul2(c.B1, bc(hi32(q))));
    const float nh = __fmul_rn(-0.5f, hsum(mul2(q, u)));
    float e = expneg(-nh);
    e = select_count<FULL>(hi32(c.nv), nh, e, cnt);
    S.s0 = __fadd_rn(S.s0, e);
    if (FULL) {
        const float a2 = hsum(mul2(u, A.j));
        const u64 v = fma2(c.B0, bc(lo32(A.j)), mul2(c.B1, bc(hi32(A.j))));
        const float w = hsum(mul2(u, A.r));
        const float k = __fsub_rn(hsum(mul2(A.j, v)), w);
        const float c9 = __fmaf_rn(-a2, a2, k);
        const u64 c45 = fma2(bc(-lo32(u)), u, c.B0);
        const u64 c68 = fma2(bc(-a2), u, v);
        const float c7 = __fmaf_rn(-hi32(u), hi32(u), hi32(c.B1));
        fma2_acc(S.s12, u, bc(e));
        fma2_acc(S.s45, c45, bc(e));
        fma2_acc(S.s68, c68, bc(e));
        S.s3 = __fmaf_rn(e, a2, S.s3);
        S.s7 = __fmaf_rn(e, c7, S.s7);
        S.s9 = __fmaf_rn(e, c9, S.s9);
    }
}
// all points of the scan, one per lane and step (what the two warps of a pair would share between them)
template <bool FULL>
__device__ __forceinline__ void eval_warp_half_probe(const LevelDev &L, const float2 *pts, int n, const Pose32 &q, int lane, Eval &E)
{
    Partials1 S;
    S.s12 = S.s45 = S.s68 = 0ull;
    S.s0 = S.s3 = S.s7 = S.s9 = 0.0f;
    int cnt = 0;
    const float4 *__restrict__ cells = L.cells;
    const PosePk P = pose_pack(q);
    const LatticePk G = lattice_pack<0>(L, TABLE_DENSE);
    const int npad = (n + 63) & ~63;
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(pts) + 8u * (unsigned)lane, s1 = s0 + 8u * (unsigned)npad;
#pragma unroll 1
    for (unsigned sa = s0; sa < s1; sa += 256u) step1<FULL>(cells, G, P, sa, S, cnt);
    Partials T;
    T.s0 = pk(S.s0, 0.f); T.s3 = pk(S.s3, 0.f); T.s9 = pk(S.s9, 0.f);
    T.s12[0] = S.s12; T.s12[1] = 0ull; T.s45[0] = S.s45; T.s45[1] = 0ull; T.s68[0] = S.s68; T.s68[1] = 0ull;
    T.s7[0] = S.s7; T.s7[1] = 0.f;
    finish_partials<FULL, true>(T, cnt, lane, E);
}
} // namespace ndt2d
