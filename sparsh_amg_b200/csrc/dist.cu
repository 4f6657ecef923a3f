// dist.cu — multi-GPU solve phase: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference has no distributed path (no MPI/NCCL anywhere).  This file row-partitions the same V-cycle / PCG:
//   * operators: each rank owns a row block of A_l, P_l and R_l = P_l^T with columns relabelled to [owned | halo]
//     positions (built on the host by host/dist_plan.cpp); the per-row entry order is preserved, so every row sum is
//     bit-identical to the single-GPU one and only the reductions (dot products) see a different summation tree;
//   * halo exchange: pack kernel -> grouped ncclSend/ncclRecv straight into the halo tail of the input vector, issued
//     on a communication stream; interior rows (no halo reference) run meanwhile on the compute stream, boundary rows
//     after the receive (north_star: "halo exchange ... overlapped with interior-row SpMV");
//   * Krylov scalars: local fixed-tree partial -> ncclAllReduce(sum) of 1-2 doubles, in place in device memory;
//   * small levels: gathered once per cycle with ncclAllGather and solved redundantly on every GPU by the single-GPU
//     hierarchy code (levels whose halo would exceed their interior are latency-bound on a partitioned layout).
#include <nccl.h>

#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "hierarchy.cuh"

using namespace sparsh;

namespace {

struct Comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
};
Comm &comm() {
    static Comm c;
    return c;
}

int nccl_fail(ncclResult_t r, const char *what, int line) {
    set_error(std::string("NCCL error at dist.cu:") + std::to_string(line) + " " + what + ": " + ncclGetErrorString(r));
    return SPARSH_ERR_CUDA;
}
#define SP_NCCL(call)                                          \
    do {                                                       \
        ncclResult_t r__ = (call);                             \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call, __LINE__); \
    } while (0)

struct DistOp {
    sparsh_matrix_s *M = nullptr;
    int nrow = 0, ncol_local = 0, nhalo = 0;
    std::vector<int> send_rank, send_ptr, recv_rank, recv_ptr;
    int *d_send_idx = nullptr;
    double *d_sendbuf = nullptr;
    int ib = 0, ie = 0;
    bool needs_exchange() const { return !send_rank.empty() || !recv_rank.empty(); }
};

struct DistLevel {
    DistOp A, P, R;
    int n = 0;        // owned rows of this level
    int n_next = 0;   // owned rows of the next level
    double *xbuf = nullptr, *tbuf = nullptr;  // [owned | halo(max of A, P_{l-1})]
    double *bbuf = nullptr;                   // owned (levels >= 1)
    double *rbuf = nullptr;                   // [owned | halo(R)]
    size_t xcap = 0;
};

__global__ void __launch_bounds__(256) pack_kernel(const double *__restrict__ x, const int *__restrict__ idx, int n, double *out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = x[idx[i]];
}
// scatter the padded all-gather buffer into the replicated coarse vector: full[map[i]] = gathered[i]
__global__ void __launch_bounds__(256) scatter_map_kernel(const double *__restrict__ src, const int *__restrict__ map, int n, double *dst) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const int g = map[i];
        if (g >= 0) dst[g] = src[i];
    }
}
__global__ void __launch_bounds__(256) gather_rows_kernel(const double *__restrict__ src, const int *__restrict__ rows, int n, double *dst) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = src[rows[i]];
}

}  // namespace

struct sparsh_dist_s {
    std::vector<DistLevel> lev;      // distributed levels 0..nd-1
    sparsh_hierarchy_s *tail = nullptr;  // replicated levels nd..L
    sparsh_params prm;
    int n_tail0 = 0;                 // global rows of the first replicated level
    int tail_maxc = 0;               // padded per-rank count for the all-gather
    int n_own_tail0 = 0;             // rows of that level owned by this rank
    double *tail_send = nullptr, *tail_recv = nullptr, *tail_b = nullptr, *tail_x = nullptr;
    int *d_tail_map = nullptr, *d_tail_rows = nullptr;
    double *xtail_local = nullptr;   // owned part of X at level nd, [owned | halo(P_{nd-1})]
    double *btail_local = nullptr;
    // Krylov
    double *kv[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *d_sc = nullptr, *h_sc = nullptr;
    std::vector<GraphEntry> graphs;
};

namespace {

int make_op(const sparsh_dist_op_desc &d, DistOp &op) {
    op.nrow = d.nrow;
    op.ncol_local = d.ncol_local;
    op.nhalo = d.nhalo;
    op.ib = d.interior_begin;
    op.ie = d.interior_end;
    SP_TRY(sparsh_matrix_create(d.nrow, d.ncol_local + d.nhalo, d.nnz, d.rowptr, d.colindex, d.val, d.diag, &op.M));
    op.send_rank.assign(d.send_rank, d.send_rank + d.n_send);
    op.send_ptr.assign(d.send_ptr, d.send_ptr + d.n_send + 1);
    op.recv_rank.assign(d.recv_rank, d.recv_rank + d.n_recv);
    op.recv_ptr.assign(d.recv_ptr, d.recv_ptr + d.n_recv + 1);
    const int total = d.n_send ? d.send_ptr[d.n_send] : 0;
    if (total > 0) {
        SP_CUDA(cudaMalloc(&op.d_send_idx, sizeof(int) * (size_t)total));
        SP_CUDA(cudaMalloc(&op.d_sendbuf, sizeof(double) * (size_t)total));
        SP_CUDA(cudaMemcpy(op.d_send_idx, d.send_idx, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice));
    }
    return SPARSH_OK;
}
void free_op(DistOp &op) {
    sparsh_matrix_destroy(op.M);
    cudaFree(op.d_send_idx);
    cudaFree(op.d_sendbuf);
}

// Fill the halo tail of x = [owned | halo] for operator `op`.  Issued on the communication stream after everything
// queued so far on the compute stream; the caller decides when the compute stream waits for it (finish_exchange).
int start_exchange(const DistOp &op, double *x) {
    if (!op.needs_exchange()) return SPARSH_OK;
    Context &c = ctx();
    Comm &m = comm();
    const int total = op.send_rank.empty() ? 0 : op.send_ptr.back();
    if (total > 0) {
        pack_kernel<<<(total + 255) / 256, 256, 0, c.stream>>>(x, op.d_send_idx, total, op.d_sendbuf);
        count_launch();
    }
    SP_CUDA(cudaEventRecord(m.ev_ready, c.stream));
    SP_CUDA(cudaStreamWaitEvent(m.comm_stream, m.ev_ready, 0));
    SP_NCCL(ncclGroupStart());
    for (size_t s = 0; s < op.send_rank.size(); s++)
        SP_NCCL(ncclSend(op.d_sendbuf + op.send_ptr[s], (size_t)(op.send_ptr[s + 1] - op.send_ptr[s]), ncclDouble,
                         op.send_rank[s], m.comm, m.comm_stream));
    for (size_t r = 0; r < op.recv_rank.size(); r++)
        SP_NCCL(ncclRecv(x + op.ncol_local + op.recv_ptr[r], (size_t)(op.recv_ptr[r + 1] - op.recv_ptr[r]), ncclDouble,
                         op.recv_rank[r], m.comm, m.comm_stream));
    SP_NCCL(ncclGroupEnd());
    SP_CUDA(cudaEventRecord(m.ev_done, m.comm_stream));
    return SPARSH_OK;
}
int finish_exchange(const DistOp &op) {
    if (!op.needs_exchange()) return SPARSH_OK;
    SP_CUDA(cudaStreamWaitEvent(ctx().stream, comm().ev_done, 0));
    return SPARSH_OK;
}

// y = epi(op x): halo exchange overlapped with the interior rows
int apply(const DistOp &op, int epi, double *x, double *y, const EpiArgs &args) {
    const bool reduces = epi == EPI_SPMV_DOT || epi == EPI_RESNORM;
    if (!op.needs_exchange() || reduces || op.ie <= op.ib) {
        // (the fused reductions need one grid over all rows: exchange first, then a single launch)
        SP_TRY(start_exchange(op, x));
        SP_TRY(finish_exchange(op));
        return launch_csr(op.M, epi, x, y, args, 0, op.nrow);
    }
    SP_TRY(start_exchange(op, x));
    SP_TRY(launch_csr(op.M, epi, x, y, args, op.ib, op.ie));  // interior rows while the halo is in flight
    SP_TRY(finish_exchange(op));
    if (op.ib > 0) SP_TRY(launch_csr(op.M, epi, x, y, args, 0, op.ib));
    if (op.ie < op.nrow) SP_TRY(launch_csr(op.M, epi, x, y, args, op.ie, op.nrow));
    return SPARSH_OK;
}

int allreduce_sum(double *d_vals, int count) {
    Comm &m = comm();
    if (m.nranks == 1) return SPARSH_OK;
    SP_NCCL(ncclAllReduce(d_vals, d_vals, (size_t)count, ncclDouble, ncclSum, m.comm, ctx().stream));
    return SPARSH_OK;
}

int dist_smooth(sparsh_dist_s *h, DistLevel &L, const double *b, double *&cur, double *&other, int sweeps, bool zero) {
    if (sweeps == 0) {
        if (zero) SP_TRY(k_fill(cur, (size_t)L.n, 0.0));
        return SPARSH_OK;
    }
    for (int s = 0; s < sweeps; s++) {
        if (s == 0 && zero) {
            SP_TRY(k_jacobi_zero((size_t)L.n, b, L.A.M->diag, h->prm.omega, other));
        } else {
            EpiArgs a;
            a.b = b;
            a.xi = cur;
            a.d = L.A.M->diag;
            a.omega = h->prm.omega;
            SP_TRY(apply(L.A, EPI_JACOBI, cur, other, a));
        }
        std::swap(cur, other);
    }
    return SPARSH_OK;
}

// x must have halo capacity (lev[0].xcap doubles).  One V-cycle, enqueue only.
int enqueue_dist_vcycle(sparsh_dist_s *h, const double *b, double *x, bool x_is_zero) {
    Context &c = ctx();
    Comm &m = comm();
    const int nd = (int)h->lev.size();
    std::vector<double *> X(nd + 1), T(nd);
    std::vector<const double *> B(nd + 1);
    for (int l = 0; l < nd; l++) {
        X[l] = l == 0 ? x : h->lev[l].xbuf;
        T[l] = h->lev[l].tbuf;
        B[l] = l == 0 ? b : h->lev[l].bbuf;
    }
    X[nd] = h->xtail_local;
    B[nd] = h->btail_local;
    for (int l = 0; l < nd; l++) {
        DistLevel &L = h->lev[l];
        SP_TRY(dist_smooth(h, L, B[l], X[l], T[l], h->prm.pre_sweeps, l > 0 || x_is_zero));
        EpiArgs a;
        a.b = B[l];
        SP_TRY(apply(L.A, EPI_RESID, X[l], L.rbuf, a));
        double *bnext = l + 1 < nd ? h->lev[l + 1].bbuf : h->btail_local;
        SP_TRY(apply(L.R, EPI_SPMV, L.rbuf, bnext, EpiArgs()));
    }
    // replicated tail: all-gather the restricted right-hand side, solve redundantly, keep the owned rows
    if (m.nranks > 1) {
        SP_CUDA(cudaMemcpyAsync(h->tail_send, h->btail_local, sizeof(double) * (size_t)h->n_own_tail0, cudaMemcpyDeviceToDevice, c.stream));
        SP_NCCL(ncclAllGather(h->tail_send, h->tail_recv, (size_t)h->tail_maxc, ncclDouble, m.comm, c.stream));
        const int tot = h->tail_maxc * m.nranks;
        scatter_map_kernel<<<(tot + 255) / 256, 256, 0, c.stream>>>(h->tail_recv, h->d_tail_map, tot, h->tail_b);
        count_launch();
    } else {
        const int tot = h->n_own_tail0;
        scatter_map_kernel<<<(tot + 255) / 256, 256, 0, c.stream>>>(h->btail_local, h->d_tail_map, tot, h->tail_b);
        count_launch();
    }
    SP_TRY(enqueue_vcycle(h->tail, h->tail_b, h->tail_x, true));
    if (h->n_own_tail0 > 0) {
        gather_rows_kernel<<<(h->n_own_tail0 + 255) / 256, 256, 0, c.stream>>>(h->tail_x, h->d_tail_rows, h->n_own_tail0, h->xtail_local);
        count_launch();
    }
    for (int l = nd; l > 0; l--) {
        DistLevel &F = h->lev[l - 1];
        SP_TRY(apply(F.P, EPI_PROLONG, X[l], X[l - 1], EpiArgs()));
        SP_TRY(dist_smooth(h, F, B[l - 1], X[l - 1], T[l - 1], h->prm.post_sweeps, false));
    }
    if (X[0] != x) SP_CUDA(cudaMemcpyAsync(x, X[0], sizeof(double) * (size_t)h->lev[0].n, cudaMemcpyDeviceToDevice, c.stream));
    SP_CUDA(cudaGetLastError());
    return SPARSH_OK;
}

template <class F>
int dist_run_graphed(sparsh_dist_s *h, const void *k0, const void *k1, int tag, F body) {
    // NCCL point-to-point and collective calls are capturable; the communication stream joins the capture through
    // the ready/done events
    Context &c = ctx();
    if (!h->prm.use_graph) return body();
    GraphEntry *ent = nullptr;
    for (auto &g : h->graphs)
        if (g.k0 == k0 && g.k1 == k1 && g.tag == tag) ent = &g;
    if (!ent) {
        GraphEntry g;
        g.k0 = k0;
        g.k1 = k1;
        g.tag = tag;
        h->graphs.push_back(g);
        return body();
    }
    if (!ent->exec) {
        cudaGraph_t graph = nullptr;
        SP_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
        c.capturing = true;
        c.captured = 0;
        int rc = body();
        c.capturing = false;
        cudaError_t e = cudaStreamEndCapture(c.stream, &graph);
        if (rc != SPARSH_OK) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        SP_CUDA(e);
        ent->kernels = c.captured;
        SP_CUDA(cudaGraphInstantiate(&ent->exec, graph, 0));
        SP_CUDA(cudaGraphDestroy(graph));
    }
    SP_CUDA(cudaGraphLaunch(ent->exec, c.stream));
    c.launches += ent->kernels;
    return SPARSH_OK;
}

}  // namespace

extern "C" {

int sparsh_dist_get_unique_id(char *id128) {
    static_assert(sizeof(ncclUniqueId) <= SPARSH_NCCL_ID_BYTES, "ncclUniqueId larger than the ABI slot");
    ncclUniqueId id;
    SP_NCCL(ncclGetUniqueId(&id));
    std::memset(id128, 0, SPARSH_NCCL_ID_BYTES);
    std::memcpy(id128, &id, sizeof id);
    return SPARSH_OK;
}

int sparsh_dist_init(const char *id128, int nranks, int rank) {
    SP_TRY(ensure_init());
    Comm &m = comm();
    if (m.comm) return SPARSH_OK;
    SP_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / world size");
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    SP_NCCL(ncclCommInitRank(&m.comm, nranks, id, rank));
    m.nranks = nranks;
    m.rank = rank;
    SP_CUDA(cudaStreamCreateWithFlags(&m.comm_stream, cudaStreamNonBlocking));
    SP_CUDA(cudaEventCreateWithFlags(&m.ev_ready, cudaEventDisableTiming));
    SP_CUDA(cudaEventCreateWithFlags(&m.ev_done, cudaEventDisableTiming));
    return SPARSH_OK;
}

int sparsh_dist_finalize(void) {
    Comm &m = comm();
    if (!m.comm) return SPARSH_OK;
    cudaStreamSynchronize(ctx().stream);
    cudaStreamSynchronize(m.comm_stream);
    ncclCommDestroy(m.comm);
    cudaStreamDestroy(m.comm_stream);
    cudaEventDestroy(m.ev_ready);
    cudaEventDestroy(m.ev_done);
    m = Comm();
    return SPARSH_OK;
}

int sparsh_dist_info(int *nranks, int *rank) {
    if (nranks) *nranks = comm().nranks;
    if (rank) *rank = comm().rank;
    return SPARSH_OK;
}

int sparsh_dist_hierarchy_create(int nd, const sparsh_dist_level_desc *lev, int ntail, const sparsh_level_desc *tail,
                                 const int *tail_counts, const int *tail_rows, const sparsh_params *params,
                                 sparsh_dist_t *out) {
    SP_TRY(ensure_init());
    Comm &m = comm();
    SP_REQUIRE(m.comm != nullptr || m.nranks == 1, "sparsh_dist_init has not been called");
    if (!m.comm_stream) {  // single-rank use without NCCL (tests): still needs the stream/event plumbing
        SP_CUDA(cudaStreamCreateWithFlags(&m.comm_stream, cudaStreamNonBlocking));
        SP_CUDA(cudaEventCreateWithFlags(&m.ev_ready, cudaEventDisableTiming));
        SP_CUDA(cudaEventCreateWithFlags(&m.ev_done, cudaEventDisableTiming));
    }
    SP_REQUIRE(nd >= 1 && ntail >= 1 && lev && tail && tail_counts && tail_rows && out, "bad distributed hierarchy description");
    sparsh_dist_s *h = new sparsh_dist_s();
    if (params)
        h->prm = *params;
    else
        sparsh_params_default(&h->prm);
    SP_REQUIRE(h->prm.smoother == 0, "the distributed path implements the Jacobi smoother only");
    h->lev.resize(nd);
    for (int l = 0; l < nd; l++) {
        DistLevel &L = h->lev[l];
        SP_TRY(make_op(lev[l].A, L.A));
        SP_TRY(make_op(lev[l].P, L.P));
        SP_TRY(make_op(lev[l].R, L.R));
        L.n = lev[l].A.nrow;
        L.n_next = lev[l].R.nrow;
        int halo = lev[l].A.nhalo;
        if (l > 0) halo = std::max(halo, lev[l - 1].P.nhalo);
        L.xcap = (size_t)L.n + (size_t)halo + 2;
        SP_CUDA(cudaMalloc(&L.tbuf, sizeof(double) * L.xcap));
        SP_CUDA(cudaMalloc(&L.rbuf, sizeof(double) * ((size_t)L.n + lev[l].R.nhalo + 2)));
        if (l > 0) {
            SP_CUDA(cudaMalloc(&L.xbuf, sizeof(double) * L.xcap));
            SP_CUDA(cudaMalloc(&L.bbuf, sizeof(double) * ((size_t)L.n + 2)));
        }
    }
    // replicated tail
    sparsh_params tp = h->prm;
    tp.use_graph = 0;  // its launches are captured as part of the enclosing distributed graph
    SP_TRY(sparsh_hierarchy_create(ntail, tail, &tp, &h->tail));
    h->n_tail0 = tail[0].nrow;
    int maxc = 0, displ = 0, my_displ = 0;
    for (int r = 0; r < m.nranks; r++) {
        maxc = std::max(maxc, tail_counts[r]);
        if (r == m.rank) my_displ = displ;
        displ += tail_counts[r];
    }
    SP_REQUIRE(displ == h->n_tail0, "tail_counts do not add up to the rows of the first replicated level");
    h->tail_maxc = std::max(maxc, 1);
    h->n_own_tail0 = tail_counts[m.rank];
    SP_REQUIRE(h->n_own_tail0 == h->lev[nd - 1].n_next, "owned rows of the first replicated level disagree with R");
    std::vector<int> map((size_t)h->tail_maxc * m.nranks, -1);
    displ = 0;
    for (int r = 0; r < m.nranks; r++) {
        for (int k = 0; k < tail_counts[r]; k++) map[(size_t)r * h->tail_maxc + k] = tail_rows[displ + k];
        displ += tail_counts[r];
    }
    if (m.nranks == 1) map.assign(tail_rows, tail_rows + h->n_tail0);
    SP_CUDA(cudaMalloc(&h->d_tail_map, sizeof(int) * map.size()));
    SP_CUDA(cudaMemcpy(h->d_tail_map, map.data(), sizeof(int) * map.size(), cudaMemcpyHostToDevice));
    SP_CUDA(cudaMalloc(&h->d_tail_rows, sizeof(int) * (size_t)std::max(h->n_own_tail0, 1)));
    SP_CUDA(cudaMemcpy(h->d_tail_rows, tail_rows + my_displ, sizeof(int) * (size_t)h->n_own_tail0, cudaMemcpyHostToDevice));
    SP_CUDA(cudaMalloc(&h->tail_send, sizeof(double) * (size_t)h->tail_maxc));
    SP_CUDA(cudaMemset(h->tail_send, 0, sizeof(double) * (size_t)h->tail_maxc));
    SP_CUDA(cudaMalloc(&h->tail_recv, sizeof(double) * (size_t)h->tail_maxc * m.nranks));
    SP_CUDA(cudaMalloc(&h->tail_b, sizeof(double) * ((size_t)h->n_tail0 + 2)));
    SP_CUDA(cudaMalloc(&h->tail_x, sizeof(double) * ((size_t)h->n_tail0 + 2)));
    SP_CUDA(cudaMalloc(&h->xtail_local, sizeof(double) * ((size_t)h->n_own_tail0 + lev[nd - 1].P.nhalo + 2)));
    SP_CUDA(cudaMalloc(&h->btail_local, sizeof(double) * ((size_t)h->n_own_tail0 + 2)));
    SP_CUDA(cudaMalloc(&h->d_sc, sizeof(double) * 16));
    SP_CUDA(cudaMallocHost(&h->h_sc, sizeof(double) * 16));
    *out = h;
    return SPARSH_OK;
}

int sparsh_dist_hierarchy_destroy(sparsh_dist_t h) {
    if (!h) return SPARSH_OK;
    if (ctx().ready) cudaStreamSynchronize(ctx().stream);
    for (auto &g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto &L : h->lev) {
        free_op(L.A);
        free_op(L.P);
        free_op(L.R);
        cudaFree(L.xbuf);
        cudaFree(L.tbuf);
        cudaFree(L.bbuf);
        cudaFree(L.rbuf);
    }
    sparsh_hierarchy_destroy(h->tail);
    cudaFree(h->tail_send);
    cudaFree(h->tail_recv);
    cudaFree(h->tail_b);
    cudaFree(h->tail_x);
    cudaFree(h->d_tail_map);
    cudaFree(h->d_tail_rows);
    cudaFree(h->xtail_local);
    cudaFree(h->btail_local);
    for (int i = 0; i < 6; i++) cudaFree(h->kv[i]);
    cudaFree(h->d_sc);
    cudaFreeHost(h->h_sc);
    delete h;
    return SPARSH_OK;
}

int sparsh_dist_local_rows(sparsh_dist_t h, int level, int *nrow) {
    SP_REQUIRE(h != nullptr && level >= 0 && level <= (int)h->lev.size(), "bad level");
    *nrow = level < (int)h->lev.size() ? h->lev[level].n : h->n_own_tail0;
    return SPARSH_OK;
}

int sparsh_dist_spmv(sparsh_dist_t h, int level, const double *d_x_local, double *d_y_local) {
    SP_REQUIRE(h != nullptr && level >= 0 && level < (int)h->lev.size(), "bad level");
    DistLevel &L = h->lev[level];
    SP_CUDA(cudaMemcpyAsync(L.tbuf, d_x_local, sizeof(double) * (size_t)L.n, cudaMemcpyDeviceToDevice, ctx().stream));
    return apply(L.A, EPI_SPMV, L.tbuf, d_y_local, EpiArgs());
}

int sparsh_dist_vcycle(sparsh_dist_t h, const double *d_b_local, double *d_x_local, int cycles, int x_is_zero) {
    SP_REQUIRE(h != nullptr && cycles >= 0, "bad arguments");
    DistLevel &L0 = h->lev[0];
    if (!h->kv[1]) SP_CUDA(cudaMalloc(&h->kv[1], sizeof(double) * L0.xcap));
    double *z = h->kv[1];  // halo-capable staging for the caller's x
    SP_CUDA(cudaMemcpyAsync(z, d_x_local, sizeof(double) * (size_t)L0.n, cudaMemcpyDeviceToDevice, ctx().stream));
    for (int k = 0; k < cycles; k++) SP_TRY(enqueue_dist_vcycle(h, d_b_local, z, x_is_zero && k == 0));
    SP_CUDA(cudaMemcpyAsync(d_x_local, z, sizeof(double) * (size_t)L0.n, cudaMemcpyDeviceToDevice, ctx().stream));
    return SPARSH_OK;
}

// Same arithmetic as cg_impl(precond = true) in krylov.cu (reference src/AMG_main_solvers.cpp:107-167); the three
// reductions of an iteration are completed by in-place all-reduces of device scalars.
int sparsh_dist_pcg(sparsh_dist_t h, const double *b, double *x, double tol, int max_iter, double *hist, int *iters_out) {
    SP_REQUIRE(h != nullptr, "hierarchy is NULL");
    Context &c = ctx();
    DistLevel &L0 = h->lev[0];
    const size_t n = (size_t)L0.n;
    for (int i = 0; i < 5; i++)
        if (!h->kv[i]) SP_CUDA(cudaMalloc(&h->kv[i], sizeof(double) * L0.xcap));
    double *r = h->kv[0], *z = h->kv[1], *p = h->kv[2], *Ap = h->kv[3], *xs = h->kv[4];
    double *sc = h->d_sc;
    enum { S_PAP = 0, S_RZ = 1, S_RZNEW = 2, S_RR = 3 };
    auto read_rr = [&]() -> int {
        SP_CUDA(cudaMemcpyAsync(h->h_sc + S_RR, sc + S_RR, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        return SPARSH_OK;
    };
    SP_CUDA(cudaMemcpyAsync(xs, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, c.stream));
    EpiArgs a;
    a.b = b;
    SP_TRY(apply(L0.A, EPI_RESID, xs, r, a));
    SP_TRY(k_dot(n, r, r, sc + S_RR));
    SP_TRY(allreduce_sum(sc + S_RR, 1));
    SP_TRY(dist_run_graphed(h, r, z, 1, [&]() { return enqueue_dist_vcycle(h, r, z, true); }));
    SP_CUDA(cudaMemcpyAsync(p, z, sizeof(double) * n, cudaMemcpyDeviceToDevice, c.stream));
    SP_TRY(k_dot(n, r, z, sc + S_RZ));
    SP_TRY(allreduce_sum(sc + S_RZ, 1));
    SP_TRY(read_rr());
    SP_CUDA(cudaStreamSynchronize(c.stream));
    double r1 = std::sqrt(h->h_sc[S_RR]);
    if (hist) hist[0] = r1;

    auto body = [&]() -> int {
        // Ap = A p and the local part of p.Ap: the fused reduction needs the full row range, so the exchange is not
        // overlapped here (one of ~16 operator applications per level-0 visit)
        EpiArgs e;
        e.xi = p;
        e.red_out = sc + S_PAP;
        SP_TRY(apply(L0.A, EPI_SPMV_DOT, p, Ap, e));
        SP_TRY(allreduce_sum(sc + S_PAP, 1));
        SP_TRY(k_pcg_update_xr(n, p, Ap, xs, r, sc + S_RZ, sc + S_PAP, sc + S_RR));
        SP_TRY(allreduce_sum(sc + S_RR, 1));
        SP_TRY(enqueue_dist_vcycle(h, r, z, true));
        SP_TRY(k_dot(n, z, r, sc + S_RZNEW));
        SP_TRY(allreduce_sum(sc + S_RZNEW, 1));
        SP_TRY(k_pcg_update_p(n, z, p, sc + S_RZNEW, sc + S_RZ));
        SP_TRY(k_scalar_copy(sc + S_RZ, sc + S_RZNEW));
        SP_TRY(read_rr());
        return SPARSH_OK;
    };
    int count = 0;
    while (count < max_iter && r1 > tol) {
        count++;
        SP_TRY(dist_run_graphed(h, x, b, 10, body));
        SP_CUDA(cudaStreamSynchronize(c.stream));
        r1 = std::sqrt(h->h_sc[S_RR]);
        if (hist) hist[count] = r1;
        if (!std::isfinite(r1)) break;
    }
    SP_CUDA(cudaMemcpyAsync(x, xs, sizeof(double) * n, cudaMemcpyDeviceToDevice, c.stream));
    SP_CUDA(cudaStreamSynchronize(c.stream));
    if (iters_out) *iters_out = count;
    return r1 <= tol ? SPARSH_OK : SPARSH_ERR_NOT_CONVERGED;
}

}  // extern "C"
