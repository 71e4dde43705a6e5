// ndt2d_align_pairs: batched scan-to-scan. One grid per target scan, built by one warp per (target, level) into a hash
// table (k_pairs_build, below); the pairs are then aligned by k_align in pairs mode (ndt2d_kernels.cu). The host side
// finds the distinct targets, sizes the tables and cuts the call into chunks that fit the table budget.
// Reference interface: none citable (/root/reference/README.md:1 is the whole mount).
#include "ndt2d_device.cuh"
#include "ndt2d_host.h"

namespace ndt2d {

static constexpr unsigned FULL_MASK = 0xffffffffu;

// Same arithmetic as the dense build (SPEC 2 auto-fit lattice, SPEC 3 integer sums and finalisation), so every record
// equals the one ndt2d_set_target(scan) would produce; only the container differs: a scan touches a few hundred of
// the cells of its bounding box, so the cells live in an open-addressing hash table keyed by the dense cell index.
static constexpr int PAIRS_LIST_CAP = 2304; // claimed-slot list per warp (u16 entries, 36 KB per block); longer: full-table scan

template <int OV>
__global__ void __launch_bounds__(256) k_pairs_build(const PairBuildArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (int64_t)a.ntargets * a.nlevels) return;
    const int t = (int)(wid / a.nlevels), l = (int)(wid % a.nlevels);
    // slots this warp claimed, so that the finalisation visits the few hundred occupied slots instead of the whole table
    __shared__ unsigned short s_list[8][PAIRS_LIST_CAP];
    __shared__ int s_nlist[8];
    const int w = threadIdx.x >> 5;
    if (lane == 0) s_nlist[w] = 0;
    const int scan = __ldg(a.targets + t);
    const int64_t o0 = __ldg(a.offsets + scan), o1 = __ldg(a.offsets + scan + 1);
    const int n = (int)(o1 - o0);
    const float2 *__restrict__ src = a.xy + o0;

    // SPEC 2 auto-fit: bounding box of the finite points (min and max are order independent)
    float xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
    if (!a.explicit_grid) {
        for (int i = lane; i < n; i += 32) {
            float2 p = __ldg(src + i);
            if (!isfinite(p.x) || !isfinite(p.y)) continue;
            xmin = fminf(xmin, p.x); xmax = fmaxf(xmax, p.x);
            ymin = fminf(ymin, p.y); ymax = fmaxf(ymax, p.y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xmin = fminf(xmin, __shfl_xor_sync(FULL_MASK, xmin, o));
            ymin = fminf(ymin, __shfl_xor_sync(FULL_MASK, ymin, o));
            xmax = fmaxf(xmax, __shfl_xor_sync(FULL_MASK, xmax, o));
            ymax = fmaxf(ymax, __shfl_xor_sync(FULL_MASK, ymax, o));
        }
    }
    // geometry, the f32 expressions of setup_level() in ndt2d_capi.cu (every lane computes the same values)
    LevelDev L;
    L.res = a.res[l];
    L.ov = a.ov;
    L.st = L.ov ? __fmul_rn(L.res, 0.5f) : L.res;
    L.inv_st = __fdiv_rn(1.0f, L.st);
    if (a.explicit_grid) {
        L.ox = a.gox; L.oy = a.goy;
        L.nhx = (int)ceilf(__fdiv_rn(a.gex, L.st));
        L.nhy = (int)ceilf(__fdiv_rn(a.gey, L.st));
    } else {
        if (!(xmin <= xmax)) { xmin = xmax = ymin = ymax = 0.0f; }
        L.ox = __fsub_rn(__fmul_rn(floorf(__fdiv_rn(xmin, L.res)), L.res), L.res);
        L.oy = __fsub_rn(__fmul_rn(floorf(__fdiv_rn(ymin, L.res)), L.res), L.res);
        L.nhx = (int)ceilf(__fdiv_rn(__fsub_rn(xmax, L.ox), L.st)) + 2;
        L.nhy = (int)ceilf(__fdiv_rn(__fsub_rn(ymax, L.oy), L.st)) + 2;
    }
    bool too_big = L.nhx < 1 || L.nhy < 1 || (int64_t)(L.nhx + L.ov) * (int64_t)(L.nhy + L.ov) >= ((int64_t)1 << 31);
    if (too_big) { L.nhx = L.nhy = 0; }      // nothing is inside: every align against this target ends NO_OVERLAP
    L.njx = L.nhx + L.ov;
    L.njy = L.nhy + L.ov;
    L.inv_std = __ddiv_rn(1.0, (double)L.st);
    L.qs = __ddiv_rn(4194304.0, (double)L.res);
    L.qu = __dmul_rn((double)L.res, 1.0 / 4194304.0);
    L.hash_mask = a.cap - 1u;
    float4 *rec = a.tab + (size_t)wid * ((size_t)a.cap + 1) * 2;
    uint32_t *cnt = a.cnt + (size_t)wid * a.cap;
    unsigned long long *sums = a.sums + (size_t)wid * a.cap * 5;
    L.cells = rec;
    L.cnt = cnt;
    L.sums = sums;
    if (lane == 0) {
        a.geo[wid] = L;
        if (too_big) atomicMax(a.error, t + 1);
    }

    // clear the records: empty key in every slot's `n` word; the sentinel record after the last slot
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 empty = make_float4(0.f, 0.f, __int_as_float((int)kEmptyKey), 0.f);
    for (unsigned s = lane; s <= a.cap; s += 32) {
        rec[2 * (size_t)s] = zero;
        rec[2 * (size_t)s + 1] = s < a.cap ? empty : zero;
    }
    // (cnt and sums are zero on entry: they are zeroed when allocated and every finalisation zeroes what it consumed)
    __syncwarp();

    // SPEC 3 accumulation, as k_accumulate: runs of equal cells are combined in the warp, run heads find or claim the
    // cell's slot (compare-and-swap on the key word) and add into its integer sums
    unsigned *keyword = reinterpret_cast<unsigned *>(rec);   // key of slot s = word 8 s + 6
    float2 pnext = lane < n ? __ldg(src + lane) : make_float2(0.f, 0.f);   // the next window's point is requested one window ahead
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const float2 p = pnext;
        if (i + 32 < n) pnext = __ldg(src + i + 32);
        accumulate_window<OV>(L, i < n, p.x, p.y, lane, [&](int key, unsigned c, long long sx, long long sy, long long sxx, long long sxy, long long syy) {
            unsigned s = hash_slot((unsigned)key, L.hash_mask);
            bool placed = false;
            // the table is at most 2/3 full by construction; the probe count is bounded all the same, so that
            // inconsistent arguments (offsets on the device that differ from the host copy) cannot hang the GPU
            for (unsigned probes = 0; probes <= L.hash_mask; ++probes) {
                unsigned was = atomicCAS(keyword + 8 * (size_t)s + 6, kEmptyKey, (unsigned)key);
                if (was == kEmptyKey) {
                    const int pos = atomicAdd(&s_nlist[w], 1);
                    if (pos < PAIRS_LIST_CAP) s_list[w][pos] = (unsigned short)s;
                    placed = true;
                    break;
                }
                if (was == (unsigned)key) { placed = true; break; }
                s = (s + 1u) & L.hash_mask;
            }
            if (!placed) { atomicMax(a.error, t + 1); return; }
            atomicAdd(cnt + s, c);
            unsigned long long *q = sums + 5 * (size_t)s;
            atomicAdd(q + 0, (unsigned long long)sx);
            atomicAdd(q + 1, (unsigned long long)sy);
            atomicAdd(q + 2, (unsigned long long)sxx);
            atomicAdd(q + 3, (unsigned long long)sxy);
            atomicAdd(q + 4, (unsigned long long)syy);
        });
    }
    __threadfence();
    __syncwarp();

    // SPEC 3 finalisation of the occupied slots; the key stays in the record's `n` word for the probing reader, and the
    // slot's sums go back to zero for the next build that uses this memory
    auto finish = [&](unsigned s) {
        const unsigned key = *(volatile unsigned *)(keyword + 8 * (size_t)s + 6);
        if (key == kEmptyKey) return;
        const unsigned nn = *(volatile uint32_t *)(cnt + s);
        volatile long long *q = reinterpret_cast<volatile long long *>(sums + 5 * (size_t)s);
        const long long q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
        float4 ra, rb;
        finalize_record(nn, q0, q1, q2, q3, q4, L.qu, a.min_points, a.eig_ratio, ra, rb);
        rb.z = __int_as_float((int)key);
        rec[2 * (size_t)s] = ra;
        rec[2 * (size_t)s + 1] = rb;
        cnt[s] = 0u;
        q[0] = 0; q[1] = 0; q[2] = 0; q[3] = 0; q[4] = 0;
    };
    const int nlist = s_nlist[w];
    if (nlist <= PAIRS_LIST_CAP) {
        for (int e = lane; e < nlist; e += 32) finish(s_list[w][e]);
    } else {
        for (unsigned s = lane; s < a.cap; s += 32) finish(s);
    }
}

cudaError_t launch_pairs_build(const LaunchCfg &c, const PairBuildArgs &a, int64_t *launches)
{
    const int64_t warps = (int64_t)a.ntargets * a.nlevels;
    if (warps <= 0) return cudaSuccess;
    const int grid = (int)((warps + 7) / 8);
    if (a.ov) k_pairs_build<1><<<grid, 256, 0, c.stream>>>(a);
    else k_pairs_build<0><<<grid, 256, 0, c.stream>>>(a);
    ++*launches;
    return cudaGetLastError();
}

} // namespace ndt2d

using namespace ndt2d;

extern "C" {

// ---- batched scan-to-scan (north_star stage 3, "batched multi-scan"): many (target scan, source scan) pairs per call ----

static unsigned next_pow2(unsigned v)
{
    unsigned p = 1;
    while (p < v) p <<= 1;
    return p;
}

// d_xy / d_offsets / d_init / d_res: device; h_offsets / pairs: host (the chunking and the target list are host work).
// One chunk = as many pairs as fit the table budget. Asynchronous on the handle's stream apart from the uploads of the
// small per-chunk index lists; *err_out (optional) receives the device error flag after a synchronisation.
static int align_pairs_impl(ndt2d_matcher *m, const float *d_xy, const int64_t *d_offsets, const int64_t *h_offsets, int nscans,
                            const int32_t *pairs, int npairs, const double *d_init, ndt2d_result *d_res)
{
    int64_t max_src = 0, max_tgt = 0;
    for (int p = 0; p < npairs; ++p) {
        const int t = pairs[2 * p], s = pairs[2 * p + 1];
        if (t < 0 || t >= nscans || s < 0 || s >= nscans) return fail(m, NDT2D_EINVAL, "pair %d: scan index out of range", p);
        max_tgt = std::max(max_tgt, h_offsets[t + 1] - h_offsets[t]);
        max_src = std::max(max_src, h_offsets[s + 1] - h_offsets[s]);
    }
    const int K = m->prm.overlap ? 4 : 1, L = m->nlevels;
    // slots per table: 1.5 x the most cells a target can occupy (every point in cells of its own), rounded up to a power
    // of two: at most 2/3 full in that worst case, typically a quarter (a 1080-beam scan occupies ~500 cells)
    const uint64_t want = 3ull * (uint64_t)std::max<int64_t>(max_tgt, 16) * K / 2;
    if (want > 65536) return fail(m, NDT2D_EINVAL, "align_pairs: target scans of %lld points need more than 65536 table slots",
                                  (long long)max_tgt);
    const unsigned cap = next_pow2((unsigned)want);
    const size_t per_target = (size_t)L * (((size_t)cap + 1) * 32 + (size_t)cap * 44 + sizeof(LevelDev));
    size_t budget = (size_t)8 << 30;
    if (const char *e = getenv("NDT2D_PAIRS_BYTES")) budget = (size_t)strtoull(e, nullptr, 10);
    const int tmax = (int)std::max<size_t>(1, std::min<size_t>(budget / per_target, (size_t)nscans));
    CK(m, m->b_perr.ensure(4));
    int rc;
    CK(m, cudaMemsetAsync(m->b_perr.p, 0, 4, m->cfg.stream));
    std::vector<int32_t> slot((size_t)nscans, -1), targets, resolved;
    int p0 = 0;
    while (p0 < npairs) {
        // the next chunk: pairs in order until the chunk's distinct targets would exceed the budget
        targets.clear();
        resolved.clear();
        int p1 = p0;
        for (; p1 < npairs; ++p1) {
            const int t = pairs[2 * p1];
            if (slot[t] < 0) {
                if ((int)targets.size() == tmax) break;
                slot[t] = (int32_t)targets.size();
                targets.push_back(t);
            }
            resolved.push_back(slot[t]);
            resolved.push_back(pairs[2 * p1 + 1]);
        }
        const size_t T = targets.size(), TL = T * (size_t)L;
        CK(m, m->b_ptab.ensure(TL * ((size_t)cap + 1) * 32));
        // the accumulators are zero between builds (every finalisation zeroes what it consumed): clear them when they are (re)allocated
        if (m->b_pcnt.cap < TL * cap * 4 || m->b_psums.cap < TL * cap * 40) {
            CK(m, cudaStreamSynchronize(m->cfg.stream));
            CK(m, m->b_pcnt.ensure(TL * cap * 4));
            CK(m, m->b_psums.ensure(TL * cap * 40));
            CK(m, cudaMemsetAsync(m->b_pcnt.p, 0, m->b_pcnt.cap, m->cfg.stream));
            CK(m, cudaMemsetAsync(m->b_psums.p, 0, m->b_psums.cap, m->cfg.stream));
        }
        CK(m, m->b_pgeo.ensure(TL * sizeof(LevelDev)));
        if ((rc = upload(m, m->b_ptargets, targets.data(), T * 4))) return rc;
        if ((rc = upload(m, m->b_ppairs, resolved.data(), resolved.size() * 4))) return rc;
        CK(m, cudaStreamSynchronize(m->cfg.stream)); // the vectors are reused by the next chunk
        PairBuildArgs b;
        memset(&b, 0, sizeof(b));
        b.xy = reinterpret_cast<const float2 *>(d_xy);
        b.offsets = d_offsets;
        b.targets = m->b_ptargets.as<int32_t>();
        b.ntargets = (int)T; b.nlevels = L; b.ov = m->prm.overlap; b.explicit_grid = m->explicit_grid ? 1 : 0;
        for (int l = 0; l < L; ++l) b.res[l] = m->res[l];
        b.gox = m->gox; b.goy = m->goy; b.gex = m->gex; b.gey = m->gey;
        b.min_points = m->prm.min_points; b.eig_ratio = m->prm.eig_ratio;
        b.cap = cap;
        b.tab = m->b_ptab.as<float4>(); b.cnt = m->b_pcnt.as<uint32_t>(); b.sums = m->b_psums.as<unsigned long long>();
        b.geo = m->b_pgeo.as<LevelDev>();
        b.error = m->b_perr.as<int>();
        CK(m, launch_pairs_build(m->cfg, b, &m->launches));
        AlignArgs a;
        fill_align_args(m, a);
        a.xy = reinterpret_cast<const float2 *>(d_xy);
        if (!a.xy) a.xy = m->b_counter.as<float2>();
        a.offsets = d_offsets;
        a.pairs = m->b_ppairs.as<int32_t>();
        a.geo = m->b_pgeo.as<LevelDev>();
        a.init = d_init + 3 * (size_t)p0;
        a.res = d_res + p0;
        a.nscans = p1 - p0;
        a.cap_points = align_cap_points(m, (int)max_src);
        CK(m, launch_align(m->cfg, a, &m->launches));
        for (int32_t t : targets) slot[t] = -1;
        p0 = p1;
    }
    return NDT2D_OK;
}

static int align_pairs_check(ndt2d_matcher *m)
{
    int err = 0;
    CK(m, cudaMemcpyAsync(&err, m->b_perr.p, 4, cudaMemcpyDeviceToHost, m->cfg.stream));
    int rc = ndt2d_synchronize(m);
    if (rc) return rc;
    if (err) return fail(m, NDT2D_EINVAL, "align_pairs: the auto-fitted lattice of a target scan exceeds 2^31 cells "
                                          "(a point far from the rest?); its pairs were returned with status NO_OVERLAP");
    return NDT2D_OK;
}

static int align_pairs_validate(ndt2d_matcher *m, const int64_t *offsets, int nscans, const int32_t *pairs, int npairs)
{
    if (nscans < 0 || npairs < 0 || (nscans > 0 && !offsets) || (npairs > 0 && !pairs)) return fail(m, NDT2D_EINVAL, "bad arguments");
    if (m->nlevels < 1) return fail(m, NDT2D_EINVAL, "no resolution set");
    for (int b = 0; b < nscans; ++b) {
        int64_t nb = offsets[b + 1] - offsets[b];
        if (nb < 0 || nb > 0x7fffffff) return fail(m, NDT2D_EINVAL, "offsets not monotone at scan %d", b);
    }
    if (nscans > 0 && offsets[0] < 0) return fail(m, NDT2D_EINVAL, "bad offsets");
    return NDT2D_OK;
}

int ndt2d_align_pairs_device(ndt2d_matcher *m, const float *d_xy, const int64_t *d_offsets, const int64_t *offsets, int nscans,
                             const int32_t *pairs, int npairs, const double *d_init, ndt2d_result *d_res)
{
    if (!m) return NDT2D_EINVAL;
    int rc = align_pairs_validate(m, offsets, nscans, pairs, npairs);
    if (rc) return rc;
    if (npairs == 0) return NDT2D_OK;
    if (!d_offsets || !d_init || !d_res) return fail(m, NDT2D_EINVAL, "bad arguments");
    DeviceGuard g(m->device);
    return align_pairs_impl(m, d_xy, d_offsets, offsets, nscans, pairs, npairs, d_init, d_res);
}

int ndt2d_align_pairs(ndt2d_matcher *m, const float *xy, const int64_t *offsets, int nscans, const int32_t *pairs, int npairs,
                      const double *init, ndt2d_result *res)
{
    if (!m) return NDT2D_EINVAL;
    int rc = align_pairs_validate(m, offsets, nscans, pairs, npairs);
    if (rc) return rc;
    if (npairs == 0) return NDT2D_OK;
    if (!init || !res) return fail(m, NDT2D_EINVAL, "bad arguments");
    const int64_t total = offsets[nscans];
    if (total > 0 && !xy) return fail(m, NDT2D_EINVAL, "bad offsets / xy");
    DeviceGuard g(m->device);
    if ((rc = upload(m, m->b_xy, xy, (size_t)total * 8))) return rc;
    if ((rc = upload(m, m->b_off, offsets, (size_t)(nscans + 1) * 8))) return rc;
    if ((rc = upload(m, m->b_init, init, (size_t)npairs * 24))) return rc;
    CK(m, m->b_res.ensure((size_t)npairs * sizeof(ndt2d_result)));
    rc = align_pairs_impl(m, m->b_xy.as<float>(), m->b_off.as<int64_t>(), offsets, nscans, pairs, npairs, m->b_init.as<double>(),
                          m->b_res.as<ndt2d_result>());
    if (rc) return rc;
    CK(m, cudaMemcpyAsync(res, m->b_res.p, (size_t)npairs * sizeof(ndt2d_result), cudaMemcpyDeviceToHost, m->cfg.stream));
    return align_pairs_check(m);
}

} // extern "C"
