// Multi-GPU exchanges over peer memory (CUDA IPC set-up, publication from kernels, host-side poll):
//   ndt2d_exchange_*, ndt2d_sweep_publish   the best hypothesis of a sharded sweep; the arg-max kernel itself stores it
//                                            into every rank's table (k_argmax_pass / PublishArgs in ndt2d_kernels.cu)
//   ndt2d_reloc_*                            a sharded relocalisation end to end: every rank sweeps its shard, refines its
//                                            own k best and stores the k {index, sweep score, refined record} candidates
//                                            into every rank's table; the global top-k is then a local merge
// Reference interface: none citable (/root/reference/README.md:1 is the whole mount).
#include "ndt2d_host.h"

#include <time.h>

using namespace ndt2d;

namespace {

void peer_close(ndt2d_matcher *m, PeerTable &t)
{
    if (t.world == 0) return;
    DeviceGuard g(m->device);
    cudaStreamSynchronize(m->cfg.stream);
    for (int r = 0; r < t.world; ++r) {
        if (r != t.rank && t.opened[r]) cudaIpcCloseMemHandle(t.table[r]);
        t.opened[r] = false;
        if (r != t.rank) t.table[r] = nullptr;
    }
    if (t.table[t.rank]) cudaFree(t.table[t.rank]);
    t.table[t.rank] = nullptr;
    if (t.host) cudaFreeHost(t.host);
    t.host = nullptr;
    t.world = t.rank = t.slots = 0;
    t.block_bytes = 0;
}

int peer_create(ndt2d_matcher *m, PeerTable &t, int world, int rank, int nslots, size_t block_bytes, unsigned char *handle)
{
    // the slot discipline (wait for q - nslots/2 before publishing q) needs at least two rows as soon as there is a peer
    if (world < 1 || world > NDT2D_MAX_RANKS || rank < 0 || rank >= world || nslots < (world > 1 ? 2 : 1) || nslots > 4096)
        return fail(m, NDT2D_EINVAL, "exchange: world %d (max %d), rank %d, nslots %d (2..4096 when world > 1)", world, NDT2D_MAX_RANKS, rank,
                    nslots);
    static_assert(sizeof(cudaIpcMemHandle_t) == NDT2D_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    peer_close(m, t);
    DeviceGuard g(m->device);
    const size_t bytes = (size_t)nslots * world * block_bytes;
    unsigned char *own = nullptr;
    CK(m, cudaMalloc(reinterpret_cast<void **>(&own), bytes)); // its own allocation: IPC handles name whole allocations
    cudaError_t e = cudaMemset(own, 0, bytes);                 // epoch 0 = nothing published
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&t.host), (size_t)world * block_bytes + bytes, cudaHostAllocDefault);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, own);
    if (e != cudaSuccess) {
        cudaFree(own);
        if (t.host) cudaFreeHost(t.host);
        t.host = nullptr;
        return fail(m, NDT2D_ECUDA, "exchange_create: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof(h));
    memset(t.host, 0, (size_t)world * block_bytes + bytes);   // both host areas start as "nothing published"
    t.world = world; t.rank = rank; t.slots = nslots; t.block_bytes = block_bytes;
    t.table[rank] = own;
    t.verified_ok.assign((size_t)nslots, 0);
    return NDT2D_OK;
}

int peer_open(ndt2d_matcher *m, PeerTable &t, const unsigned char *handles)
{
    if (t.world == 0) return fail(m, NDT2D_EINVAL, "exchange_open before exchange_create");
    DeviceGuard g(m->device);
    for (int r = 0; r < t.world; ++r) {
        if (r == t.rank || t.opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * NDT2D_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(m, NDT2D_ECUDA, "exchange_open: rank %d's table: %s", r, cudaGetErrorString(e));
        t.table[r] = static_cast<unsigned char *>(p);
        t.opened[r] = true;
    }
    return NDT2D_OK;
}

int peer_ready(ndt2d_matcher *m, const PeerTable &t, const char *what)
{
    if (t.world == 0) return fail(m, NDT2D_EINVAL, "%s before the exchange was created", what);
    for (int r = 0; r < t.world; ++r)
        if (!t.table[r]) return fail(m, NDT2D_EINVAL, "%s: rank %d's table is not open", what, r);
    return NDT2D_OK;
}

uint64_t block_epoch(const unsigned char *block) { return reinterpret_cast<const ndt2d_best *>(block)->epoch; }

// Blocks until every rank has published `query` (or timeout), then returns the verified host copy of its row.
// The poll copies the row of the query to pinned memory on the copy stream, so that it never waits for kernels queued on
// the handle's stream, and backs off between polls. A row is trusted only from a snapshot taken after an earlier snapshot
// showed all its epochs (data, system fence, epoch: a complete epoch guarantees the data was written before the later
// copy started). The verifying snapshot copies the whole table once, so that waiting for several finished queries in a
// row costs no further copy.
int peer_wait_row(ndt2d_matcher *m, PeerTable &t, uint64_t query, int timeout_ms, const unsigned char **row_out)
{
    DeviceGuard g(m->device);
    const int W = t.world;
    const size_t rows = (size_t)t.slots, rb = t.row_bytes(), bytes = rows * rb;
    const size_t row = (size_t)(query % (uint64_t)t.slots);
    unsigned char *probe = t.host, *verified = t.host + rb;
    auto complete = [&](const unsigned char *r) {
        for (int k = 0; k < W; ++k)
            if (block_epoch(r + (size_t)k * t.block_bytes) != query + 1) return false;
        return true;
    };
    if (t.verified_ok[row] && complete(verified + row * rb)) {
        *row_out = verified + row * rb;
        return NDT2D_OK;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int polls = 0;; ++polls) {
        CK(m, cudaMemcpyAsync(probe, t.table[t.rank] + row * rb, rb, cudaMemcpyDeviceToHost, m->copy_stream));
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
    // a row of the whole-table snapshot is verified if the polled row (this query) or the previous snapshot already
    // showed the same complete epochs
    std::vector<uint64_t> before(rows);
    for (size_t q = 0; q < rows; ++q) {
        const unsigned char *r = verified + q * rb;
        bool full = block_epoch(r) != 0;
        for (int k = 1; k < W; ++k) full = full && block_epoch(r + (size_t)k * t.block_bytes) == block_epoch(r);
        before[q] = full ? block_epoch(r) : 0;
    }
    before[row] = query + 1;   // shown complete by the poll above
    CK(m, cudaMemcpyAsync(verified, t.table[t.rank], bytes, cudaMemcpyDeviceToHost, m->copy_stream));
    CK(m, cudaStreamSynchronize(m->copy_stream));
    for (size_t q = 0; q < rows; ++q) {
        bool same = before[q] != 0;
        for (int k = 0; k < W; ++k) same = same && block_epoch(verified + q * rb + (size_t)k * t.block_bytes) == before[q];
        t.verified_ok[q] = same;
    }
    if (!t.verified_ok[row] || !complete(verified + row * rb))   // the row moved on between the two copies: slot discipline broken
        return fail(m, NDT2D_EINVAL, "exchange_wait: row of query %llu was overwritten while waiting (see the slot discipline in ndt2d.h)",
                    (unsigned long long)query);
    *row_out = verified + row * rb;
    return NDT2D_OK;
}

} // namespace

extern "C" {

// ---- best hypothesis of a sharded sweep ---------------------------------------------------------------------------

int ndt2d_exchange_close(ndt2d_matcher *m)
{
    if (!m) return NDT2D_EINVAL;
    peer_close(m, m->ex);
    return NDT2D_OK;
}

int ndt2d_exchange_create(ndt2d_matcher *m, int world, int rank, int nslots, unsigned char *handle)
{
    if (!m || !handle) return NDT2D_EINVAL;
    static_assert(sizeof(ndt2d_best) == 32, "ndt2d_best is 32 bytes");
    return peer_create(m, m->ex, world, rank, nslots, sizeof(ndt2d_best), handle);
}

int ndt2d_exchange_open(ndt2d_matcher *m, const unsigned char *handles)
{
    if (!m || !handles) return NDT2D_EINVAL;
    return peer_open(m, m->ex, handles);
}

int ndt2d_sweep_publish(ndt2d_matcher *m, int level, const float *d_xy, int n, const float *d_hyp, int64_t nhyp,
                        double *d_scores, int64_t index_offset, uint64_t query)
{
    int rc = check_level(m, level);
    if (rc) return rc;
    if ((rc = peer_ready(m, m->ex, "sweep_publish"))) return rc;
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
    for (int r = 0; r < m->ex.world; ++r) pub.table[r] = reinterpret_cast<ndt2d_best *>(m->ex.table[r]);
    pub.world = m->ex.world; pub.rank = m->ex.rank; pub.row = (int)(query % (uint64_t)m->ex.slots);
    pub.index_offset = index_offset;
    pub.epoch = query + 1;
    CK(m, launch_topk(m->cfg, d_scores, nhyp, 1, m->b_tki.as<int64_t>(), m->b_tkv.as<double>(), m->b_scratch.as<unsigned long long>(),
                      &m->launches, &pub));
    return NDT2D_OK;
}

int ndt2d_exchange_wait(ndt2d_matcher *m, uint64_t query, int timeout_ms, int64_t *best_index, double *best_score)
{
    if (!m || !best_index || !best_score) return NDT2D_EINVAL;
    if (m->ex.world == 0) return fail(m, NDT2D_EINVAL, "exchange_wait before exchange_create");
    const unsigned char *row = nullptr;
    int rc = peer_wait_row(m, m->ex, query, timeout_ms, &row);
    if (rc) return rc;
    // best of the complete row by (-score, index), SPEC 6
    const ndt2d_best *b = reinterpret_cast<const ndt2d_best *>(row);
    int64_t bi = -1;
    double bs = 0.0;
    for (int r = 0; r < m->ex.world; ++r) {
        if (b[r].index < 0 || b[r].score != b[r].score) continue;
        if (bi < 0 || b[r].score > bs || (b[r].score == bs && b[r].index < bi)) { bi = b[r].index; bs = b[r].score; }
    }
    *best_index = bi;
    *best_score = bs;
    return NDT2D_OK;
}

// ---- sharded relocalisation: refined candidates ---------------------------------------------------------------------

int ndt2d_reloc_close(ndt2d_matcher *m)
{
    if (!m) return NDT2D_EINVAL;
    peer_close(m, m->rx);
    m->rx_k = 0;
    return NDT2D_OK;
}

int ndt2d_reloc_create(ndt2d_matcher *m, int world, int rank, int nslots, int kmax, unsigned char *handle)
{
    if (!m || !handle) return NDT2D_EINVAL;
    if (kmax < 1 || kmax > 64) return fail(m, NDT2D_EINVAL, "reloc_create: kmax %d not in [1, 64]", kmax);
    static_assert(sizeof(ndt2d_candidate) == 176 && offsetof(ndt2d_candidate, epoch) == offsetof(ndt2d_best, epoch), "ndt2d_candidate layout");
    int rc = peer_create(m, m->rx, world, rank, nslots, (size_t)kmax * sizeof(ndt2d_candidate), handle);
    if (rc == NDT2D_OK) m->rx_k = kmax;
    return rc;
}

int ndt2d_reloc_open(ndt2d_matcher *m, const unsigned char *handles)
{
    if (!m || !handles) return NDT2D_EINVAL;
    return peer_open(m, m->rx, handles);
}

int ndt2d_relocalize_publish(ndt2d_matcher *m, int level, const float *d_xy, int n, const float *d_hyp, int64_t nhyp,
                             int64_t index_offset, int k, uint64_t query)
{
    int rc = check_level(m, level);
    if (rc) return rc;
    if ((rc = peer_ready(m, m->rx, "relocalize_publish"))) return rc;
    if (k < 1 || k > m->rx_k) return fail(m, NDT2D_EINVAL, "relocalize_publish: k %d not in [1, %d] (kmax of ndt2d_reloc_create)", k, m->rx_k);
    DeviceGuard g(m->device);
    CK(m, m->b_tki.ensure((size_t)k * 8));
    CK(m, m->b_res.ensure((size_t)k * sizeof(ndt2d_result)));
    // sweep of the shard, its top-k, the k refinements: everything queued on the stream (ndt2d_relocalize_device) ...
    rc = ndt2d_relocalize_device(m, level, d_xy, n, d_hyp, nhyp, k, m->b_tki.as<int64_t>(), m->b_res.as<ndt2d_result>());
    if (rc) return rc;
    // ... and one more small kernel stores the k candidates into every rank's table
    CandidatePublishArgs pub;
    memset(&pub, 0, sizeof(pub));
    for (int r = 0; r < m->rx.world; ++r)
        pub.table[r] = reinterpret_cast<ndt2d_candidate *>(m->rx.table[r] + ((size_t)(query % (uint64_t)m->rx.slots) * m->rx.world + m->rx.rank) * m->rx.block_bytes);
    pub.world = m->rx.world;
    pub.k = k;
    pub.index_offset = index_offset;
    pub.epoch = query + 1;
    CK(m, launch_publish_candidates(m->cfg, m->b_tki.as<int64_t>(), m->b_tkv.as<double>(), m->b_res.as<ndt2d_result>(), pub, &m->launches));
    return NDT2D_OK;
}

int ndt2d_relocalize_wait(ndt2d_matcher *m, uint64_t query, int timeout_ms, int k, int64_t *best_idx, ndt2d_result *res)
{
    if (!m || !best_idx || !res) return NDT2D_EINVAL;
    if (m->rx.world == 0) return fail(m, NDT2D_EINVAL, "relocalize_wait before reloc_create");
    if (k < 1 || k > m->rx_k) return fail(m, NDT2D_EINVAL, "relocalize_wait: k %d not in [1, %d]", k, m->rx_k);
    const unsigned char *row = nullptr;
    int rc = peer_wait_row(m, m->rx, query, timeout_ms, &row);
    if (rc) return rc;
    // the global top-k by (-sweep score, index), SPEC 6: every member of it is among its own shard's k best
    std::vector<const ndt2d_candidate *> all;
    for (int r = 0; r < m->rx.world; ++r) {
        const ndt2d_candidate *c = reinterpret_cast<const ndt2d_candidate *>(row + (size_t)r * m->rx.block_bytes);
        for (int j = 0; j < k; ++j)
            if (c[j].index >= 0 && c[j].sweep_score == c[j].sweep_score) all.push_back(c + j);
    }
    std::sort(all.begin(), all.end(), [](const ndt2d_candidate *a, const ndt2d_candidate *b) {
        return a->sweep_score > b->sweep_score || (a->sweep_score == b->sweep_score && a->index < b->index);
    });
    for (int j = 0; j < k; ++j) {
        if (j < (int)all.size()) {
            best_idx[j] = all[j]->index;
            res[j] = all[j]->refined;
        } else {
            best_idx[j] = -1;
            memset(res + j, 0, sizeof(ndt2d_result));
            res[j].status = NDT2D_NO_OVERLAP;
        }
    }
    return NDT2D_OK;
}

} // extern "C"
