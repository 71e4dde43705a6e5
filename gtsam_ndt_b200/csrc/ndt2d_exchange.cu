// ndt2d_exchange_*, ndt2d_sweep_publish: best-hypothesis exchange of a sharded sweep over peer memory (CUDA IPC set-up,
// publication from the arg-max kernel, host-side poll). The kernel side is k_argmax_pass / PublishArgs in ndt2d_kernels.cu.
// Reference interface: none citable (/root/reference/README.md:1 is the whole mount).
#include "ndt2d_host.h"

#include <time.h>

using namespace ndt2d;

extern "C" {

// ---- multi-GPU best-hypothesis exchange over peer memory -------------------------------------------------------

int ndt2d_exchange_close(ndt2d_matcher *m)
{
    if (!m) return NDT2D_EINVAL;
    if (m->ex_world == 0) return NDT2D_OK;
    DeviceGuard g(m->device);
    cudaStreamSynchronize(m->cfg.stream);
    for (int r = 0; r < m->ex_world; ++r) {
        if (r != m->ex_rank && m->ex_opened[r]) cudaIpcCloseMemHandle(m->ex_table[r]);
        m->ex_opened[r] = false;
        if (r != m->ex_rank) m->ex_table[r] = nullptr;
    }
    if (m->ex_table[m->ex_rank]) cudaFree(m->ex_table[m->ex_rank]);
    m->ex_table[m->ex_rank] = nullptr;
    if (m->ex_host) cudaFreeHost(m->ex_host);
    m->ex_host = nullptr;
    m->ex_world = m->ex_rank = m->ex_slots = 0;
    return NDT2D_OK;
}

int ndt2d_exchange_create(ndt2d_matcher *m, int world, int rank, int nslots, unsigned char *handle)
{
    if (!m || !handle) return NDT2D_EINVAL;
    // the slot discipline (wait for q - nslots/2 before publishing q) needs at least two rows as soon as there is a peer
    if (world < 1 || world > NDT2D_MAX_RANKS || rank < 0 || rank >= world || nslots < (world > 1 ? 2 : 1) || nslots > 4096)
        return fail(m, NDT2D_EINVAL, "exchange: world %d (max %d), rank %d, nslots %d (2..4096 when world > 1)", world, NDT2D_MAX_RANKS, rank,
                    nslots);
    static_assert(sizeof(cudaIpcMemHandle_t) == NDT2D_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    static_assert(sizeof(ndt2d_best) == 32, "ndt2d_best is 32 bytes");
    ndt2d_exchange_close(m);
    DeviceGuard g(m->device);
    const size_t bytes = (size_t)nslots * world * sizeof(ndt2d_best);
    ndt2d_best *own = nullptr;
    CK(m, cudaMalloc(reinterpret_cast<void **>(&own), bytes)); // its own allocation: IPC handles name whole allocations
    cudaError_t e = cudaMemset(own, 0, bytes);                 // epoch 0 = nothing published
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&m->ex_host), 2 * bytes, cudaHostAllocDefault);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, own);
    if (e != cudaSuccess) {
        cudaFree(own);
        if (m->ex_host) cudaFreeHost(m->ex_host);
        m->ex_host = nullptr;
        return fail(m, NDT2D_ECUDA, "exchange_create: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof(h));
    memset(m->ex_host, 0, 2 * bytes);   // both host snapshots start as "nothing published"
    m->ex_world = world; m->ex_rank = rank; m->ex_slots = nslots;
    m->ex_table[rank] = own;
    m->ex_verified_ok.assign((size_t)nslots, 0);
    return NDT2D_OK;
}

int ndt2d_exchange_open(ndt2d_matcher *m, const unsigned char *handles)
{
    if (!m || !handles) return NDT2D_EINVAL;
    if (m->ex_world == 0) return fail(m, NDT2D_EINVAL, "exchange_open before exchange_create");
    DeviceGuard g(m->device);
    for (int r = 0; r < m->ex_world; ++r) {
        if (r == m->ex_rank || m->ex_opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * NDT2D_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(m, NDT2D_ECUDA, "exchange_open: rank %d's table: %s", r, cudaGetErrorString(e));
        m->ex_table[r] = static_cast<ndt2d_best *>(p);
        m->ex_opened[r] = true;
    }
    return NDT2D_OK;
}

int ndt2d_sweep_publish(ndt2d_matcher *m, int level, const float *d_xy, int n, const float *d_hyp, int64_t nhyp,
                        double *d_scores, int64_t index_offset, uint64_t query)
{
    int rc = check_level(m, level);
    if (rc) return rc;
    if (m->ex_world == 0) return fail(m, NDT2D_EINVAL, "sweep_publish before exchange_create");
    for (int r = 0; r < m->ex_world; ++r)
        if (!m->ex_table[r]) return fail(m, NDT2D_EINVAL, "sweep_publish: rank %d's table is not open (ndt2d_exchange_open)", r);
    if (n < 0 || nhyp < 0 || (nhyp > 0 && !d_hyp)) return fail(m, NDT2D_EINVAL, "bad arguments");
    DeviceGuard g(m->device);
    if (!d_scores) {
        CK(m, m->b_scores.ensure((size_t)(nhyp ? nhyp : 1) * 8));
        d_scores = m->b_scores.as<double>();
    }
    CK(m, m->b_tki.ensure(8));
    CK(m, m->b_tkv.ensure(8));
    CK(m, launch_eval_poses(m->cfg, m->lv[level], reinterpret_cast<const float2 *>(d_xy), n, d_hyp, 1, nhyp, 0, d_scores, 1,
                            nullptr, &m->launches));
    PublishArgs pub;
    memset(&pub, 0, sizeof(pub));
    for (int r = 0; r < m->ex_world; ++r) pub.table[r] = m->ex_table[r];
    pub.world = m->ex_world; pub.rank = m->ex_rank; pub.row = (int)(query % (uint64_t)m->ex_slots);
    pub.index_offset = index_offset;
    pub.epoch = query + 1;
    CK(m, launch_topk(m->cfg, d_scores, nhyp, 1, m->b_tki.as<int64_t>(), m->b_tkv.as<double>(), m->b_scratch.as<unsigned long long>(),
                      &m->launches, &pub));
    return NDT2D_OK;
}

// best of one complete row by (-score, index), SPEC 6
static void exchange_pick(const ndt2d_best *row, int W, int64_t *best_index, double *best_score)
{
    int64_t bi = -1;
    double bs = 0.0;
    for (int r = 0; r < W; ++r) {
        const ndt2d_best &b = row[r];
        if (b.index < 0 || b.score != b.score) continue;
        if (bi < 0 || b.score > bs || (b.score == bs && b.index < bi)) { bi = b.index; bs = b.score; }
    }
    *best_index = bi;
    *best_score = bs;
}

// The poll copies the row of the query (world x 32 B) to pinned memory on the copy stream, so that it never waits for
// kernels queued on the handle's stream, and backs off between polls. A row is trusted only from a snapshot taken after
// an earlier snapshot showed all its epochs (record, system fence, epoch: a complete epoch guarantees the record was
// written before the later copy started). The verifying snapshot copies the whole table once, so that waiting for
// several finished queries in a row costs one more small copy each at most.
int ndt2d_exchange_wait(ndt2d_matcher *m, uint64_t query, int timeout_ms, int64_t *best_index, double *best_score)
{
    if (!m || !best_index || !best_score) return NDT2D_EINVAL;
    if (m->ex_world == 0) return fail(m, NDT2D_EINVAL, "exchange_wait before exchange_create");
    DeviceGuard g(m->device);
    const int W = m->ex_world;
    const size_t rows = (size_t)m->ex_slots, bytes = rows * W * sizeof(ndt2d_best), row_bytes = (size_t)W * sizeof(ndt2d_best);
    const size_t row = (size_t)(query % (uint64_t)m->ex_slots);
    ndt2d_best *probe = m->ex_host, *verified = m->ex_host + rows * W;
    auto complete = [&](const ndt2d_best *t) {
        for (int r = 0; r < W; ++r)
            if (t[row * W + r].epoch != query + 1) return false;
        return true;
    };
    if (m->ex_verified_ok[row] && complete(verified)) {
        exchange_pick(verified + row * W, W, best_index, best_score);
        return NDT2D_OK;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int polls = 0;; ++polls) {
        CK(m, cudaMemcpyAsync(probe + row * W, m->ex_table[m->ex_rank] + row * W, row_bytes, cudaMemcpyDeviceToHost, m->copy_stream));
        CK(m, cudaStreamSynchronize(m->copy_stream));
        if (complete(probe)) break;
        const auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (us > (long long)timeout_ms * 1000)
            return fail(m, NDT2D_ETIMEOUT, "exchange_wait: query %llu not published by every rank within %d ms", (unsigned long long)query,
                        timeout_ms);
        if (polls > 64) {   // a sweep takes 0.3-2 ms: after the first polls sleep a little instead of hammering the copy engine
            struct timespec ts = {0, polls > 1024 ? 200000 : 20000};
            nanosleep(&ts, nullptr);
        }
    }
    // the verifying snapshot: the whole table, so that rows of other finished queries are verified in passing. A row of it
    // is verified if the polled row (this query) or the previous whole snapshot already showed the same complete epochs.
    std::vector<uint64_t> before(rows);
    for (size_t q = 0; q < rows; ++q) {
        bool full = verified[q * W].epoch != 0;
        for (int r = 1; r < W; ++r) full = full && verified[q * W + r].epoch == verified[q * W].epoch;
        before[q] = full ? verified[q * W].epoch : 0;
    }
    before[row] = query + 1;   // shown complete by the poll above
    CK(m, cudaMemcpyAsync(verified, m->ex_table[m->ex_rank], bytes, cudaMemcpyDeviceToHost, m->copy_stream));
    CK(m, cudaStreamSynchronize(m->copy_stream));
    for (size_t q = 0; q < rows; ++q) {
        bool same = before[q] != 0;
        for (int r = 0; r < W; ++r) same = same && verified[q * W + r].epoch == before[q];
        m->ex_verified_ok[q] = same;
    }
    if (!m->ex_verified_ok[row] || !complete(verified))   // the row moved on between the two copies: slot discipline broken
        return fail(m, NDT2D_EINVAL, "exchange_wait: row of query %llu was overwritten while waiting (see the slot discipline in ndt2d.h)",
                    (unsigned long long)query);
    exchange_pick(verified + row * W, W, best_index, best_score);
    return NDT2D_OK;
}

} // extern "C"
