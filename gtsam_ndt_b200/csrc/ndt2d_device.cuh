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

// ------------------------------------------------------------------------------------------------
// Packed path: Blackwell's f32x2 instructions (FFMA2 / FMUL2 / FADD2) carry two points per issue slot.
// Every packed operation is the IEEE round-to-nearest operation of SPEC 4 applied to both halves, so
// the terms are bit-identical to the scalar path above; negations are sign-bit flips (exact).
// Identities used to place the negations (all exact under round-to-nearest-even):
//   -(a*b) == (-a)*b,   fma(n,-C,h) == -fma(n,C,-h),   z + M == M - (-z).
// ------------------------------------------------------------------------------------------------
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 pk1(float c) { return pk(c, c); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 neg2(u64 a) { return a ^ 0x8000000080000000ull; }

// 32-byte cell record in one 256-bit load (LDG.E.256 on sm_100a)
struct Cell8 {
    float mux, muy, B00, B01, B11, det, n, valid;
};
__device__ __forceinline__ Cell8 load_cell256(const float4 *__restrict__ cells, unsigned idx)
{
    Cell8 r;
    const float4 *p = cells + 2 * (size_t)idx;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.mux), "=f"(r.muy), "=f"(r.B00), "=f"(r.B01), "=f"(r.B11), "=f"(r.det), "=f"(r.n), "=f"(r.valid)
        : "l"(p));
    return r;
}

// pose constants replicated into both halves, built once per evaluation
struct Pose2 {
    u64 c, s, ns, tx, ty;
};
__device__ __forceinline__ Pose2 pose_pack(const Pose32 &q)
{
    Pose2 P;
    P.c = pk1(q.c); P.s = pk1(q.s); P.ns = pk1(-q.s); P.tx = pk1(q.tx); P.ty = pk1(q.ty);
    return P;
}

// SPEC 4.1 on two values; nh = -h. Returns exp(-h) in both halves.
__device__ __forceinline__ u64 expneg2(u64 nh)
{
    const u64 MAGIC = pk1(12582912.0f);
    u64 nz = mul2(nh, pk1(1.44269502f));
    u64 t = sub2(MAGIC, nz);
    u64 nf = sub2(t, MAGIC);
    float t0, t1;
    upk(t, t0, t1);
    int n0 = __float_as_int(t0) - 0x4B400000, n1 = __float_as_int(t1) - 0x4B400000;
    u64 y = fma2(nf, pk1(0.693145752f), nh);
    y = fma2(nf, pk1(1.42860677e-6f), y);
    u64 p = fma2(pk1(1.38888889e-3f), y, pk1(8.33333333e-3f));
    p = fma2(p, y, pk1(4.16666667e-2f));
    p = fma2(p, y, pk1(1.66666667e-1f));
    p = fma2(p, y, pk1(0.5f));
    p = fma2(p, y, pk1(1.0f));
    p = fma2(p, y, pk1(1.0f));
    return mul2(p, pk(__int_as_float(0x3F800000 - n0 * 0x800000), __int_as_float(0x3F800000 - n1 * 0x800000)));
}

// SPEC 4 for two (point, cell) pairs at once. ok0/ok1 say which halves contribute (already false for
// points outside the lattice); they are cleared for invalid cells and for h >= 30.
template <bool FULL>
__device__ __forceinline__ void pair_terms2(const Cell8 &r0, const Cell8 &r1, u64 rx, u64 ry, u64 nry, u64 X, u64 Y,
                                            bool &ok0, bool &ok1, u64 T[10])
{
    ok0 = ok0 && (r0.valid != 0.0f);
    ok1 = ok1 && (r1.valid != 0.0f);
    const u64 B00 = pk(r0.B00, r1.B00), B01 = pk(r0.B01, r1.B01), B11 = pk(r0.B11, r1.B11);
    u64 qx = sub2(X, pk(r0.mux, r1.mux)), qy = sub2(Y, pk(r0.muy, r1.muy));
    u64 ux = fma2(B00, qx, mul2(B01, qy));
    u64 uy = fma2(B01, qx, mul2(B11, qy));
    u64 mm = fma2(qx, ux, mul2(qy, uy));
    u64 nh = mul2(pk1(-0.5f), mm);
    float nh0, nh1;
    upk(nh, nh0, nh1);
    ok0 = ok0 && (nh0 > -30.0f);
    ok1 = ok1 && (nh1 > -30.0f);
    u64 e = expneg2(nh);
    T[0] = e;
    if (FULL) {
        u64 nux = neg2(ux), nuy = neg2(uy);
        u64 a2 = fma2(uy, rx, mul2(ux, nry));
        u64 vx = fma2(B01, rx, mul2(B00, nry));
        u64 vy = fma2(B11, rx, mul2(B01, nry));
        u64 w = fma2(ux, rx, mul2(uy, ry));
        u64 k = fma2(rx, vy, mul2(nry, vx));
        k = sub2(k, w);
        k = fma2(neg2(a2), a2, k);
        T[1] = mul2(e, ux);
        T[2] = mul2(e, uy);
        T[3] = mul2(e, a2);
        T[4] = mul2(e, fma2(nux, ux, B00));
        T[5] = mul2(e, fma2(nux, uy, B01));
        T[6] = mul2(e, fma2(nux, a2, vx));
        T[7] = mul2(e, fma2(nuy, uy, B11));
        T[8] = mul2(e, fma2(nuy, a2, vy));
        T[9] = mul2(e, k);
    }
}

// SPEC 4 for one warp, packed: the scan lives in shared memory as two planes xs[], ys[] padded with NaN
// to a multiple of 64 points; lane l takes points (64 j + 2 l, 64 j + 2 l + 1). Cell records arrive through
// one 256-bit load each; points outside the lattice read record 0 and are masked.
template <int OV, bool FULL>
__device__ __forceinline__ void eval_warp2(const LevelDev &L, const float *xs, const float *ys, int npad, const Pose32 &q,
                                           int lane, Eval &E)
{
    constexpr int NT = FULL ? 10 : 1;
    double acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.0;
    int cnt = 0;
    const float4 *__restrict__ cells = L.cells;
    const Pose2 P = pose_pack(q);
    const u64 ox = pk1(L.ox), oy = pk1(L.oy), inv = pk1(L.inv_st);
    const float nhxf = L.nhxf, nhyf = L.nhyf;
    const unsigned njx = (unsigned)L.njx;
#pragma unroll 1
    for (int i = 2 * lane; i < npad; i += 64) {
        u64 x = *reinterpret_cast<const u64 *>(xs + i), y = *reinterpret_cast<const u64 *>(ys + i);
        u64 rx = fma2(P.c, x, mul2(P.ns, y));
        u64 ry = fma2(P.s, x, mul2(P.c, y));
        u64 X = add2(rx, P.tx), Y = add2(ry, P.ty);
        u64 fx = mul2(sub2(X, ox), inv), fy = mul2(sub2(Y, oy), inv);
        float fx0, fx1, fy0, fy1;
        upk(fx, fx0, fx1);
        upk(fy, fy0, fy1);
        bool in0 = (fx0 >= 0.0f) && (fx0 < nhxf) && (fy0 >= 0.0f) && (fy0 < nhyf);
        bool in1 = (fx1 >= 0.0f) && (fx1 < nhxf) && (fy1 >= 0.0f) && (fy1 < nhyf);
        unsigned b0 = in0 ? (unsigned)(int)fy0 * njx + (unsigned)(int)fx0 : 0u;
        unsigned b1 = in1 ? (unsigned)(int)fy1 * njx + (unsigned)(int)fx1 : 0u;
        u64 nry = neg2(ry);
        // K = 1: one record per point. K = 4: two rows of two adjacent records (64 contiguous bytes per row);
        // rows are processed one after the other to bound the registers held by in-flight loads.
        constexpr int ROWS = OV ? 2 : 1, COLS = OV ? 2 : 1;
#pragma unroll
        for (int b = 0; b < ROWS; ++b) {
            Cell8 r0[COLS], r1[COLS];
#pragma unroll
            for (int a = 0; a < COLS; ++a) {
                r0[a] = load_cell256(cells, b0 + a + b * njx);
                r1[a] = load_cell256(cells, b1 + a + b * njx);
            }
#pragma unroll
            for (int a = 0; a < COLS; ++a) {
                bool ok0 = in0, ok1 = in1;
                u64 T[10];
                pair_terms2<FULL>(r0[a], r1[a], rx, ry, nry, X, Y, ok0, ok1, T);
                if (ok0) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        float lo, hi;
                        upk(T[t], lo, hi);
                        acc[t] += (double)lo;
                    }
                    cnt += 1;
                }
                if (ok1) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        float lo, hi;
                        upk(T[t], lo, hi);
                        acc[t] += (double)hi;
                    }
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
