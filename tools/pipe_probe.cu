// Issue cost of the instruction classes of k_align's point loop on one SM sub-partition (B200, sm_100a):
// cycles per warp-instruction at 1, 2, 4 and 8 resident warps per scheduler, for independent chains of one instruction
// class and for interleaved pairs of classes. Answers: does a packed FFMA2 cost the scheduler one issue slot or two, do
// the FP64 and conversion pipes run beside the FMA pipe, and what the loop's instruction mix could issue at best.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe tools/pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
#define CHAINS 8
#define REP 8      // the body is repeated REP times per trip, so the three loop-control instructions are 3 in 64+
#define ITERS 512

enum Op { FFMA, FFMA2, FMUL2, DFMA, DADD, F2F_WIDE, F2F_NARROW, IMAD, LOP, IMADW, LEA2, MOVS, MIX_FFMA2_FFMA, MIX_FFMA2_DFMA, MIX_FFMA2_F2F, MIX_FFMA2_LOP, MIX_DFMA_F2F, MIX_LOOP, NOPS };
static const char *kNames[NOPS] = {"FFMA", "FFMA2 (packed f32x2)", "FMUL2", "DFMA", "DADD", "F2F.F64.F32", "F2F.F32.F64", "IMAD", "LOP3", "IMAD.WIDE (address = index * 32 + base)", "shift-add pair for the same 64-bit address", "MOV (register copy)",
                                   "FFMA2 + FFMA alternating", "FFMA2 + DFMA alternating", "FFMA2 + F2F alternating", "FFMA2 + LOP3 alternating",
                                   "DFMA + F2F alternating", "loop mix: 8 FFMA2, 2 FFMA, 2 IMAD, 3 DFMA/DADD, 1 F2F, 1 LOP3 (17)"};
static const int kInstrPerIter[NOPS] = {CHAINS, CHAINS, CHAINS, CHAINS, CHAINS, CHAINS, CHAINS, CHAINS, CHAINS, CHAINS, 2 * CHAINS, CHAINS, 2 * CHAINS, 2 * CHAINS, 2 * CHAINS,
                                        2 * CHAINS, 2 * CHAINS, 17};

template <int OP>
__global__ void __launch_bounds__(1024) k_probe(u64 *out_cycles, float *sink, float seed)
{
    float f[CHAINS];
    u64 p[CHAINS];
    double d[CHAINS], w[CHAINS];
    int n[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        f[i] = seed + i;
        d[i] = seed * 0.5 + i;
        w[i] = 0.0;
        n[i] = (int)seed + i + threadIdx.x;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(seed + i), "f"(seed - i));
    }
    const int k3 = ((int)seed + (int)threadIdx.x) | 3;
    const float a = seed * 1.0001f, b = seed * 0.5f;
    const double da = seed * 1.0001, db = seed * 0.25;
    u64 pa;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(b));
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int rep = 0; rep < REP; ++rep) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == FFMA || OP == MIX_FFMA2_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
            if (OP == FFMA2 || OP == MIX_FFMA2_FFMA || OP == MIX_FFMA2_DFMA || OP == MIX_FFMA2_F2F || OP == MIX_FFMA2_LOP)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pa));
            if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            if (OP == DFMA || OP == MIX_FFMA2_DFMA || OP == MIX_DFMA_F2F) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
            if (OP == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(db));
            // the empty asm makes the source opaque: the conversion cannot be hoisted or merged with the previous one
            if (OP == F2F_WIDE || OP == MIX_FFMA2_F2F || OP == MIX_DFMA_F2F) { asm volatile("" : "+f"(f[i])); asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w[i]) : "f"(f[i])); }
            if (OP == F2F_NARROW) { asm volatile("" : "+d"(d[i])); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i])); }
            if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(n[i]) : "r"(k3));
            if (OP == IMADW) asm volatile("mad.wide.u32 %0, %1, 32, %0;" : "+l"(p[i]) : "r"(n[i]));
            if (OP == LEA2) asm volatile("{.reg .b64 t; cvt.u64.u32 t, %1; shl.b64 t, t, 5; add.s64 %0, %0, t;}" : "+l"(p[i]) : "r"(n[i]));
            if (OP == MOVS) { asm volatile("mov.b32 %0, %1;" : "=r"(n[i]) : "r"(n[(i + 1) % CHAINS])); }
            if (OP == MIX_FFMA2_LOP || OP == LOP) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(n[i]) : "r"(k3));
        }
        if (OP == MIX_LOOP) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pa));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[0]) : "f"(a), "f"(b));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[1]) : "f"(a), "f"(b));
            asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(n[0]) : "r"(k3));
            asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(n[2]) : "r"(k3));
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[0]) : "d"(da), "d"(db));
            asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[1]) : "d"(db));
            asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[2]) : "d"(db));
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w[3]) : "f"(f[0]));
            asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(n[4]) : "r"(k3));
        }
      }
    }
    const long long t1 = clock64();
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        acc += f[i] + lo + hi + (float)d[i] + (float)w[i] + (float)n[i];
    }
    if (acc == 12345.678f) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out_cycles[0] = (u64)(t1 - t0);
}

template <int OP>
static void run(u64 *d_cycles, float *d_sink)
{
    printf("%-86s", kNames[OP]);
    for (int wps = 1; wps <= 8; wps *= 2) {           // warps per scheduler: one block on one SM, 4 * wps warps
        cudaMemset(d_cycles, 0, 8);
        k_probe<OP><<<1, 128 * wps>>>(d_cycles, d_sink, 1.5f);
        cudaDeviceSynchronize();
        k_probe<OP><<<1, 128 * wps>>>(d_cycles, d_sink, 1.5f);
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaGetLastError();
        u64 c = 0;
        cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
        // cycles per warp-instruction and scheduler = cycles / (instructions per warp * warps per scheduler)
        if (e != cudaSuccess) printf("  %6s", cudaGetErrorName(e));
        else printf("  %6.3f", (double)c / ((double)ITERS * REP * kInstrPerIter[OP] * wps));
    }
    printf("\n");
}

int main()
{
    u64 *d_cycles;
    float *d_sink;
    cudaMalloc(&d_cycles, 8);
    cudaMalloc(&d_sink, 4);
    printf("cycles per warp-instruction per scheduler (SM sub-partition); columns: 1, 2, 4, 8 warps per scheduler, %d independent chains per warp\n", CHAINS);
    run<FFMA>(d_cycles, d_sink);
    run<FFMA2>(d_cycles, d_sink);
    run<FMUL2>(d_cycles, d_sink);
    run<DFMA>(d_cycles, d_sink);
    run<DADD>(d_cycles, d_sink);
    run<F2F_WIDE>(d_cycles, d_sink);
    run<F2F_NARROW>(d_cycles, d_sink);
    run<IMAD>(d_cycles, d_sink);
    run<LOP>(d_cycles, d_sink);
    run<IMADW>(d_cycles, d_sink);
    run<LEA2>(d_cycles, d_sink);
    run<MOVS>(d_cycles, d_sink);
    run<MIX_FFMA2_FFMA>(d_cycles, d_sink);
    run<MIX_FFMA2_DFMA>(d_cycles, d_sink);
    run<MIX_FFMA2_F2F>(d_cycles, d_sink);
    run<MIX_FFMA2_LOP>(d_cycles, d_sink);
    run<MIX_DFMA_F2F>(d_cycles, d_sink);
    run<MIX_LOOP>(d_cycles, d_sink);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
