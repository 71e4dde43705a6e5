// Device-side building blocks of the 2D NDT path: lattice index (SPEC 2), per-pair terms (SPEC 4),
// expneg (SPEC 4.1), warp evaluation with f64 shuffle reduction, and the damped 3x3 solve (SPEC 5).
// Compiled with -fmad=false: only the explicit fmaf()/fma() calls below fuse, exactly as SPEC.md
// writes them. Reference file:line: none exists (/root/reference/README.md:1 is the whole mount).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ndt2d_internal.h"

namespace ndt2d {

__device__ __forceinline__ float ld_f(const float *p) { return __ldg(p); }

// SPEC 2: lattice index. Returns false when outside (also for NaN).
__device__ __forceinline__ bool lattice(const LevelDev &L, float X, float Y, int &hx, int &hy)
{
    float fx = __fmul_rn(__fsub_rn(X, L.ox), L.inv_st);
    float fy = __fmul_rn(__fsub_rn(Y, L.oy), L.inv_st);
    bool inside = (fx >= 0.0f) && (fx < L.nhxf) && (fy >= 0.0f) && (fy < L.nhyf);
    hx = (int)fx;
    hy = (int)fy;
    return inside;
}

// SPEC 4.1: exp(-h), bit-exact sequence of f32 operations.
__device__ __forceinline__ float expneg(float h)
{
    float z = __fmul_rn(h, 1.44269502f);
    float t = __fadd_rn(z, 12582912.0f);
    float nf = __fsub_rn(t, 12582912.0f);
    int ni = __float_as_int(t) - 0x4B400000;
    float r = __fmaf_rn(nf, -0.693145752f, h);
    r = __fmaf_rn(nf, -1.42860677e-6f, r);
    float y = -r;
    float p = 1.38888889e-3f;
    p = __fmaf_rn(p, y, 8.33333333e-3f);
    p = __fmaf_rn(p, y, 4.16666667e-2f);
    p = __fmaf_rn(p, y, 1.66666667e-1f);
    p = __fmaf_rn(p, y, 0.5f);
    p = __fmaf_rn(p, y, 1.0f);
    p = __fmaf_rn(p, y, 1.0f);
    return __fmul_rn(p, __int_as_float(0x3F800000 - ni * 0x800000));
}

struct Pose32 {
    float c, s, tx, ty;
};

// SPEC 4: pose to f32 (f64 sincos, one rounding each)
__device__ __forceinline__ Pose32 pose_to_f32(double tx, double ty, double th)
{
    double sn, cs;
    sincos(th, &sn, &cs);
    Pose32 q;
    q.c = (float)cs;
    q.s = (float)sn;
    q.tx = (float)tx;
    q.ty = (float)ty;
    return q;
}

// One cell record = two 16-byte halves {mux, muy, B00, B01 | B11, det, n, valid}
struct CellRec {
    float4 a, b;
};

__device__ __forceinline__ CellRec load_cell(const float4 *__restrict__ cells, size_t idx)
{
    CellRec r;
    r.a = __ldg(cells + 2 * idx);
    r.b = __ldg(cells + 2 * idx + 1);
    return r;
}

// SPEC 4: the ten f32 terms of one (point, cell) pair. Returns false when the pair is skipped.
template <bool FULL>
__device__ __forceinline__ bool pair_terms(const CellRec &rec, float rx, float ry, float X, float Y, float T[10])
{
    if (rec.b.w == 0.0f) return false;
    const float B00 = rec.a.z, B01 = rec.a.w, B11 = rec.b.x;
    float qx = __fsub_rn(X, rec.a.x), qy = __fsub_rn(Y, rec.a.y);
    float ux = __fmaf_rn(B00, qx, __fmul_rn(B01, qy));
    float uy = __fmaf_rn(B01, qx, __fmul_rn(B11, qy));
    float mm = __fmaf_rn(qx, ux, __fmul_rn(qy, uy));
    float h = __fmul_rn(0.5f, mm);
    if (!(h < 30.0f)) return false;
    float e = expneg(h);
    T[0] = e;
    if (FULL) {
        float a2 = __fmaf_rn(uy, rx, -__fmul_rn(ux, ry));
        float vx = __fmaf_rn(B01, rx, -__fmul_rn(B00, ry));
        float vy = __fmaf_rn(B11, rx, -__fmul_rn(B01, ry));
        float w = __fmaf_rn(ux, rx, __fmul_rn(uy, ry));
        float k = __fmaf_rn(rx, vy, -__fmul_rn(ry, vx));
        k = __fsub_rn(k, w);
        k = __fmaf_rn(-a2, a2, k);
        T[1] = __fmul_rn(e, ux);
        T[2] = __fmul_rn(e, uy);
        T[3] = __fmul_rn(e, a2);
        T[4] = __fmul_rn(e, __fmaf_rn(-ux, ux, B00));
        T[5] = __fmul_rn(e, __fmaf_rn(-ux, uy, B01));
        T[6] = __fmul_rn(e, __fmaf_rn(-ux, a2, vx));
        T[7] = __fmul_rn(e, __fmaf_rn(-uy, uy, B11));
        T[8] = __fmul_rn(e, __fmaf_rn(-uy, a2, vy));
        T[9] = __fmul_rn(e, k);
    }
    return true;
}

// Result of one evaluation held redundantly by every lane of the warp.
struct Eval {
    double v[10];
    int count;
};

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x; // xor butterfly: every lane ends with the same bits
}

// SPEC 4 for one warp: lanes stride over the scan (shared or global memory), f32 terms are widened
// and summed in f64 per lane, then a fixed xor-butterfly combines the lanes (deterministic).
template <int OV, bool FULL>
__device__ __forceinline__ void eval_warp(const LevelDev &L, const float2 *pts, int n, const Pose32 &q, int lane,
                                          Eval &E)
{
    constexpr int NT = FULL ? 10 : 1;
    double acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.0;
    int cnt = 0;
    const float4 *__restrict__ cells = L.cells;
#pragma unroll 2
    for (int i = lane; i < n; i += 32) {
        float2 p = pts[i];
        float rx = __fmaf_rn(q.c, p.x, -__fmul_rn(q.s, p.y));
        float ry = __fmaf_rn(q.s, p.x, __fmul_rn(q.c, p.y));
        float X = __fadd_rn(rx, q.tx), Y = __fadd_rn(ry, q.ty);
        int hx, hy;
        if (!lattice(L, X, Y, hx, hy)) continue;
        size_t base = (size_t)hy * (size_t)L.njx + (size_t)hx;
        if (OV == 0) {
            CellRec rec = load_cell(cells, base);
            float T[10];
            if (pair_terms<FULL>(rec, rx, ry, X, Y, T)) {
#pragma unroll
                for (int t = 0; t < NT; ++t) acc[t] += (double)T[t];
                cnt += 1;
            }
        } else {
            CellRec rec[4];
            rec[0] = load_cell(cells, base);
            rec[1] = load_cell(cells, base + 1);
            rec[2] = load_cell(cells, base + L.njx);
            rec[3] = load_cell(cells, base + L.njx + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float T[10];
                if (pair_terms<FULL>(rec[k], rx, ry, X, Y, T)) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) acc[t] += (double)T[t];
                    cnt += 1;
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) E.v[t] = warp_sum(acc[t]);
    if (!FULL) {
#pragma unroll
        for (int t = 1; t < 10; ++t) E.v[t] = 0.0;
    }
    E.count = __reduce_add_sync(0xffffffffu, cnt);
}

// SPEC 5: damped Cholesky solve in f64, no contraction. g = v[1..3], H6 = v[4..9].
__device__ __forceinline__ bool solve3(const double *g, const double *H6, double lambda, double d[3])
{
    double A00 = H6[0] + lambda * fmax(fabs(H6[0]), 1e-9);
    double A11 = H6[3] + lambda * fmax(fabs(H6[3]), 1e-9);
    double A22 = H6[5] + lambda * fmax(fabs(H6[5]), 1e-9);
    double A01 = H6[1], A02 = H6[2], A12 = H6[4];
    double p0 = A00;
    if (!(p0 > 0.0)) return false;
    double L00 = sqrt(p0);
    double L10 = A01 / L00, L20 = A02 / L00;
    double p1 = A11 - L10 * L10;
    if (!(p1 > 0.0)) return false;
    double L11 = sqrt(p1);
    double L21 = (A12 - L20 * L10) / L11;
    double p2 = (A22 - L20 * L20) - L21 * L21;
    if (!(p2 > 0.0)) return false;
    double L22 = sqrt(p2);
    double y0 = -g[0] / L00;
    double y1 = (-g[1] - L10 * y0) / L11;
    double y2 = ((-g[2] - L20 * y0) - L21 * y1) / L22;
    d[2] = y2 / L22;
    d[1] = (y1 - L21 * d[2]) / L11;
    d[0] = ((y0 - L10 * d[1]) - L20 * d[2]) / L00;
    return true;
}

} // namespace ndt2d
