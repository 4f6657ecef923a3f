// dist_plan.cpp — host-side partitioning of a hierarchy for the multi-GPU solve phase.
//
// The reference is single-process; this is new.  Every rank holds the whole host hierarchy (the setup phase is
// sequential host code anyway) and cuts out ITS part:
//   * ownership: level 0 is split into contiguous, balanced row blocks; a coarse row belongs to the rank that owns the
//     first fine row of its aggregate / its C point.  Inheriting ownership through P keeps restriction and prolongation
//     almost entirely local and makes the reversed numbering of odd HEM levels (backward sweep, reference
//     src/AMG_coarsening.cpp:56) a non-issue: blocks follow the grid, not the index.
//   * local operators: rows in ascending global order, columns relabelled to [owned | halo] positions WITHOUT reordering
//     the entries of a row, so device row sums are bit-identical to the single-GPU ones.
//   * exchange plans: who sends which owned entries to whom, where received entries land in the halo segment.
//   * levels with at most `tail_threshold` rows are not partitioned: they are replicated on every GPU.
// The same arrays feed the device (sparsh_dist_hierarchy_create) and the CPU tests (gloo, world size 2).
#include <omp.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <numeric>
#include <vector>

#include "../../include/sparsh_b200.h"
#include "sparsh_amg.hpp"

using sparsh::options;

namespace {

// std::vector whose resize() leaves trivially constructible elements uninitialised: the big arrays below are written in
// full by the threads right after (a value-initialising resize clears hundreds of MB on one thread first)
template <class T>
struct DefaultInit : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = DefaultInit<U>;
    };
    template <class U>
    void construct(U *p) noexcept {
        ::new (static_cast<void *>(p)) U;
    }
    template <class U, class... Args>
    void construct(U *p, Args &&...a) {
        ::new (static_cast<void *>(p)) U(std::forward<Args>(a)...);
    }
};

struct OpLocal {
    int nrow = 0, ncol_local = 0, nhalo = 0;
    std::vector<int> rp;
    std::vector<int, DefaultInit<int>> ci;
    std::vector<double, DefaultInit<double>> v, diag;
    std::vector<int> send_rank, send_ptr, send_idx, recv_rank, recv_ptr;
    std::vector<int> halo_global;  // global ids of the halo entries, in halo order (tests)
    int ib = 0, ie = 0;
    sparsh_dist_op_desc desc() const {
        sparsh_dist_op_desc d;
        std::memset(&d, 0, sizeof d);
        d.nrow = nrow;
        d.ncol_local = ncol_local;
        d.nhalo = nhalo;
        d.nnz = rp.empty() ? 0 : rp.back();
        d.rowptr = rp.data();
        d.colindex = ci.data();
        d.val = v.data();
        d.diag = diag.empty() ? nullptr : diag.data();
        d.n_send = (int)send_rank.size();
        d.send_rank = send_rank.data();
        d.send_ptr = send_ptr.data();
        d.send_idx = send_idx.data();
        d.n_recv = (int)recv_rank.size();
        d.recv_rank = recv_rank.data();
        d.recv_ptr = recv_ptr.data();
        d.interior_begin = ib;
        d.interior_end = ie;
        return d;
    }
};

struct Space {                 // distribution of one level's index space
    std::vector<int> owner;    // global id -> rank
    std::vector<int> loc;      // global id -> position inside its owner's ascending list
    std::vector<int> count;    // per rank
};

struct LevelLocal {
    OpLocal A, P, R;
    std::vector<int> rows;  // owned global ids of this level, ascending
};

struct DistPlan {
    int nranks = 1, rank = 0, nd = 0, nlevels = 0;
    std::vector<LevelLocal> lev;
    std::vector<int> tail_counts, tail_rows, tail_rows_mine;
    AMG_solver *S = nullptr;
    sparsh_dist_t device = nullptr;
};

void finish_space(Space &s, int nranks) {
    const int n = (int)s.owner.size();
    s.loc.resize(n);
    s.count.assign(nranks, 0);
    // loc[g] = number of ids below g with the same owner: per-chunk counts, offsets over the chunks, then the ids
    const int nt = std::max(1, options().threads);
    std::vector<int> chunk_count((size_t)(nt + 1) * nranks, 0);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num(), n_t = omp_get_num_threads();
        const int g0 = (int)((long long)n * t / n_t), g1 = (int)((long long)n * (t + 1) / n_t);
        int *mine = chunk_count.data() + (size_t)(t + 1) * nranks;
        for (int g = g0; g < g1; g++) mine[s.owner[g]]++;
#pragma omp barrier
#pragma omp single
        for (int k = 0; k < n_t; k++)
            for (int r = 0; r < nranks; r++) chunk_count[(size_t)(k + 1) * nranks + r] += chunk_count[(size_t)k * nranks + r];
        std::vector<int> run(chunk_count.begin() + (size_t)t * nranks, chunk_count.begin() + (size_t)(t + 1) * nranks);
        for (int g = g0; g < g1; g++) s.loc[g] = run[s.owner[g]]++;
#pragma omp barrier
#pragma omp single
        for (int r = 0; r < nranks; r++) s.count[r] = chunk_count[(size_t)n_t * nranks + r];
    }
}

// rows owned by `rank` of the global CSR (row space rs, column space cs) -> local operator + exchange plan
void build_op(int nrow_g, const int *rp, const int *ci, const double *v, const double *diag, const Space &rs,
              const Space &cs, int nranks, int rank, OpLocal &op) {
    std::vector<int> rows;
    rows.reserve(rs.count[rank]);
    for (int g = 0; g < nrow_g; g++)
        if (rs.owner[g] == rank) rows.push_back(g);
    op.nrow = (int)rows.size();
    op.ncol_local = cs.count[rank];
    // halo = referenced columns owned elsewhere, ordered by (owner, global id)
    std::vector<std::pair<int, int>> halo;
    op.rp.assign((size_t)op.nrow + 1, 0);
#pragma omp parallel num_threads(options().threads)
    {
        std::vector<std::pair<int, int>> mine;
#pragma omp for schedule(static) nowait
        for (int k = 0; k < op.nrow; k++) {
            const int g = rows[k];
            op.rp[k + 1] = rp[g + 1] - rp[g];
            for (int j = rp[g]; j < rp[g + 1]; j++)
                if (cs.owner[ci[j]] != rank) mine.emplace_back(cs.owner[ci[j]], ci[j]);
        }
        std::sort(mine.begin(), mine.end());
        mine.erase(std::unique(mine.begin(), mine.end()), mine.end());
#pragma omp critical
        halo.insert(halo.end(), mine.begin(), mine.end());
    }
    for (int k = 0; k < op.nrow; k++) op.rp[k + 1] += op.rp[k];
    std::sort(halo.begin(), halo.end());
    halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
    op.nhalo = (int)halo.size();
    op.halo_global.resize(halo.size());
    for (size_t h = 0; h < halo.size(); h++) op.halo_global[h] = halo[h].second;
    for (size_t h = 0; h < halo.size();) {
        size_t e = h;
        while (e < halo.size() && halo[e].first == halo[h].first) e++;
        op.recv_rank.push_back(halo[h].first);
        op.recv_ptr.push_back((int)h);
        h = e;
    }
    op.recv_ptr.push_back((int)halo.size());
    // local CSR: same entries in the same order, columns relabelled
    op.ci.resize((size_t)std::max(op.rp[op.nrow], 1));
    op.v.resize((size_t)std::max(op.rp[op.nrow], 1));
    if (diag) op.diag.resize((size_t)op.nrow);
    std::vector<char> boundary((size_t)op.nrow, 0);
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (int k = 0; k < op.nrow; k++) {
        const int g = rows[k];
        int o = op.rp[k];
        for (int j = rp[g]; j < rp[g + 1]; j++, o++) {
            const int c = ci[j];
            if (cs.owner[c] == rank) {
                op.ci[o] = cs.loc[c];
            } else {
                const auto it = std::lower_bound(halo.begin(), halo.end(), std::make_pair(cs.owner[c], c));
                op.ci[o] = op.ncol_local + (int)(it - halo.begin());
                boundary[k] = 1;
            }
            op.v[o] = v[j];
        }
        if (diag) op.diag[k] = diag[g];
    }
    // what the others need from me: columns I own that appear in rows owned by q != rank
    std::vector<std::pair<int, int>> need;
#pragma omp parallel num_threads(options().threads)
    {
        std::vector<std::pair<int, int>> mine;
#pragma omp for schedule(static) nowait
        for (int g = 0; g < nrow_g; g++) {
            const int q = rs.owner[g];
            if (q == rank) continue;
            for (int j = rp[g]; j < rp[g + 1]; j++)
                if (cs.owner[ci[j]] == rank) mine.emplace_back(q, ci[j]);
        }
        std::sort(mine.begin(), mine.end());
        mine.erase(std::unique(mine.begin(), mine.end()), mine.end());
#pragma omp critical
        need.insert(need.end(), mine.begin(), mine.end());
    }
    std::sort(need.begin(), need.end());
    need.erase(std::unique(need.begin(), need.end()), need.end());
    for (size_t h = 0; h < need.size();) {
        size_t e = h;
        while (e < need.size() && need[e].first == need[h].first) e++;
        op.send_rank.push_back(need[h].first);
        op.send_ptr.push_back((int)h);
        h = e;
    }
    op.send_ptr.push_back((int)need.size());
    op.send_idx.resize(need.size());
    for (size_t h = 0; h < need.size(); h++) op.send_idx[h] = cs.loc[need[h].second];
    (void)nranks;
    // square operators (A): a row whose value a neighbour needs also counts as boundary, so that the fused Jacobi
    // kernel that computes the boundary strips is the one that stores those values into the neighbours
    if (&rs == &cs)
        for (int k : op.send_idx) boundary[k] = 1;
    // longest run of rows that touch no halo entry (and feed none): computed while the exchange is in flight
    int best_b = 0, best_e = 0, run_b = 0;
    for (int k = 0; k <= op.nrow; k++) {
        if (k == op.nrow || boundary[k]) {
            if (k - run_b > best_e - best_b) {
                best_b = run_b;
                best_e = k;
            }
            run_b = k + 1;
        }
    }
    op.ib = best_b;
    op.ie = best_e;
}

// stable transpose (row c lists the source rows ascending); every thread owns a contiguous range of output rows and
// scans the entries in order; output arrays are not zero-filled first (see host/setup.cpp: transpose)
void transpose_csr(int nrow, int ncol, const int *rp, const int *ci, const double *v, std::vector<int> &trp,
                   std::unique_ptr<int[]> &tci, std::unique_ptr<double[]> &tv) {
    const int nnz = rp[nrow];
    trp.assign((size_t)ncol + 1, 0);
    tci.reset(new int[(size_t)std::max(nnz, 1)]);
    tv.reset(new double[(size_t)std::max(nnz, 1)]);
    for (int j = 0; j < nnz; j++) trp[ci[j] + 1]++;
    for (int c = 0; c < ncol; c++) trp[c + 1] += trp[c];
#pragma omp parallel num_threads(std::max(1, options().threads))
    {
        const int t = omp_get_thread_num(), n_t = omp_get_num_threads();
        const int c0 = (int)((long long)ncol * t / n_t), c1 = (int)((long long)ncol * (t + 1) / n_t);
        std::vector<int> cur(trp.begin() + c0, trp.begin() + c1);
        for (int i = 0; i < nrow; i++)
            for (int j = rp[i]; j < rp[i + 1]; j++) {
                const int c = ci[j];
                if (c >= c0 && c < c1) {
                    const int d = cur[c - c0]++;
                    tci[d] = i;
                    tv[d] = v[j];
                }
            }
    }
}

}  // namespace

extern "C" {

// tail_threshold: levels with at most this many rows are replicated (at least the coarsest always is, and level 0
// is always distributed)
void *sparsh_host_dist_plan(void *Sv, int nranks, int rank, int tail_threshold) {
    AMG_solver *S = (AMG_solver *)Sv;
    const double t_begin = omp_get_wtime();
    DistPlan *pl = new DistPlan();
    pl->nranks = nranks;
    pl->rank = rank;
    pl->S = S;
    pl->nlevels = S->l + 1;
    if (S->l == 0) {
        delete pl;
        return nullptr;  // a single-level hierarchy has nothing to distribute
    }
    int nd = 1;
    while (nd < S->l && S->Av[nd]->nrow > tail_threshold) nd++;
    pl->nd = nd;
    // index-space distributions of levels 0..nd
    std::vector<Space> sp((size_t)nd + 1);
    {
        const int n0 = S->Av[0]->nrow;
        sp[0].owner.resize(n0);
        for (int r = 0; r < nranks; r++) {
            const long b = (long)n0 * r / nranks, e = (long)n0 * (r + 1) / nranks;
            std::fill(sp[0].owner.begin() + b, sp[0].owner.begin() + e, r);
        }
        finish_space(sp[0], nranks);
    }
    for (int l = 0; l < nd; l++) {
        const sp_matrix_mg *P = S->Pv[l];
        std::vector<int> &own = sp[l + 1].owner;
        // the first (lowest) fine row of every coarse row decides: a minimum, so the threads may arrive in any order
        std::vector<int> first((size_t)P->ncol, INT_MAX);
#pragma omp parallel for num_threads(options().threads) schedule(static)
        for (int i = 0; i < P->nrow; i++)
            for (int j = P->rowptr[i]; j < P->rowptr[i + 1]; j++) {
                int *slot = &first[P->colindex[j]];
                int seen = __atomic_load_n(slot, __ATOMIC_RELAXED);
                while (i < seen && !__atomic_compare_exchange_n(slot, &seen, i, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                }
            }
        own.resize((size_t)P->ncol);
#pragma omp parallel for num_threads(options().threads) schedule(static)
        for (int c = 0; c < P->ncol; c++)  // (a coarse row nobody interpolates from cannot happen with HEM/Beck: rank 0)
            own[c] = first[c] == INT_MAX ? 0 : sp[l].owner[first[c]];
        finish_space(sp[l + 1], nranks);
    }
    const bool tm = getenv("SPARSH_SETUP_TIMING") != nullptr;
    const double ts = omp_get_wtime();
    if (tm) std::printf("  plan: index spaces %.3f s\n", ts - t_begin);
    pl->lev.resize(nd);
    for (int l = 0; l < nd; l++) {
        const sp_matrix_mg *A = S->Av[l], *P = S->Pv[l];
        LevelLocal &L = pl->lev[l];
        const double t0 = omp_get_wtime();
        for (int g = 0; g < A->nrow; g++)
            if (sp[l].owner[g] == rank) L.rows.push_back(g);
        build_op(A->nrow, A->rowptr, A->colindex, A->val, A->diagonal, sp[l], sp[l], nranks, rank, L.A);
        const double t1 = omp_get_wtime();
        build_op(P->nrow, P->rowptr, P->colindex, P->val, nullptr, sp[l], sp[l + 1], nranks, rank, L.P);
        const double t2 = omp_get_wtime();
        std::vector<int> trp;
        std::unique_ptr<int[]> tci;
        std::unique_ptr<double[]> tv;
        transpose_csr(P->nrow, P->ncol, P->rowptr, P->colindex, P->val, trp, tci, tv);
        const double t3 = omp_get_wtime();
        build_op(P->ncol, trp.data(), tci.get(), tv.get(), nullptr, sp[l + 1], sp[l], nranks, rank, L.R);
        if (tm)
            std::printf("  plan: level %d A %.3f P %.3f transpose %.3f R %.3f s\n", l, t1 - t0, t2 - t1, t3 - t2,
                        omp_get_wtime() - t3);
    }
    pl->tail_counts = sp[nd].count;
    pl->tail_rows.clear();
    for (int r = 0; r < nranks; r++)
        for (int g = 0; g < (int)sp[nd].owner.size(); g++)
            if (sp[nd].owner[g] == r) pl->tail_rows.push_back(g);
    return pl;
}

void sparsh_host_dist_plan_free(void *plv) {
    DistPlan *pl = (DistPlan *)plv;
    if (!pl) return;
    if (pl->device) sparsh_dist_hierarchy_destroy(pl->device);
    delete pl;
}

int sparsh_host_dist_plan_levels(void *plv, int *nd, int *nlevels) {
    DistPlan *pl = (DistPlan *)plv;
    *nd = pl->nd;
    *nlevels = pl->nlevels;
    return 0;
}

// owned global row ids of distributed level l (l == nd: the first replicated level)
int sparsh_host_dist_plan_rows(void *plv, int l, const int **rows) {
    DistPlan *pl = (DistPlan *)plv;
    if (l < pl->nd) {
        *rows = pl->lev[l].rows.data();
        return (int)pl->lev[l].rows.size();
    }
    int displ = 0;
    for (int r = 0; r < pl->rank; r++) displ += pl->tail_counts[r];
    *rows = pl->tail_rows.data() + displ;
    return pl->tail_counts[pl->rank];
}

// which: 0 = A, 1 = P, 2 = R of distributed level l (arrays stay owned by the plan)
int sparsh_host_dist_plan_op(void *plv, int l, int which, sparsh_dist_op_desc *out, const int **halo_global) {
    DistPlan *pl = (DistPlan *)plv;
    const OpLocal &op = which == 0 ? pl->lev[l].A : which == 1 ? pl->lev[l].P : pl->lev[l].R;
    *out = op.desc();
    if (halo_global) *halo_global = op.halo_global.data();
    return 0;
}

// upload: distributed levels + replicated tail -> sparsh_dist_t (sparsh_dist_init must have been called)
void *sparsh_host_dist_upload(void *plv) {
    DistPlan *pl = (DistPlan *)plv;
    if (pl->device) return pl->device;
    const sparsh::Options &o = options();
    std::vector<sparsh_dist_level_desc> d((size_t)pl->nd);
    for (int l = 0; l < pl->nd; l++) {
        d[l].A = pl->lev[l].A.desc();
        d[l].P = pl->lev[l].P.desc();
        d[l].R = pl->lev[l].R.desc();
    }
    AMG_solver *S = pl->S;
    const int ntail = pl->nlevels - pl->nd;
    std::vector<sparsh_level_desc> t((size_t)ntail);
    for (int k = 0; k < ntail; k++) {
        const int l = pl->nd + k;
        std::memset(&t[k], 0, sizeof(sparsh_level_desc));
        sp_matrix_mg *A = S->Av[l];
        t[k].nrow = A->nrow;
        t[k].nnz = A->rowptr[A->nrow];
        t[k].rowptr = A->rowptr;
        t[k].colindex = A->colindex;
        t[k].val = A->val;
        t[k].diag = A->diagonal;
        if (l < S->l) {
            sp_matrix_mg *P = S->Pv[l];
            t[k].p_ncol = P->ncol;
            t[k].p_nnz = P->rowptr[P->nrow];
            t[k].p_rowptr = P->rowptr;
            t[k].p_colindex = P->colindex;
            t[k].p_val = P->val;
        }
    }
    sparsh_params prm;
    sparsh_params_default(&prm);
    prm.omega = o.relax;
    prm.use_graph = o.use_graph;
    prm.pre_sweeps = prm.post_sweeps = o.sweeps;
    prm.halo_mode = o.halo_mode;
    int rc = sparsh_dist_hierarchy_create(pl->nd, d.data(), ntail, t.data(), pl->tail_counts.data(), pl->tail_rows.data(),
                                          &prm, &pl->device);
    if (rc != SPARSH_OK) {
        std::fprintf(stderr, "sparsh_amg: distributed upload failed (%d): %s\n", rc, sparsh_last_error());
        return nullptr;
    }
    return pl->device;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Multi-GPU entry points in the style of the reference's AMG.hpp solvers (the reference itself is single-GPU:
// src/AMG_gpu_phases_2.cu).  Every rank — one process per GPU — calls with the SAME global A, b, x it would hand to
// Solver_PCG_1, plus its rank, the number of ranks and the 128-byte id that rank 0 obtained from
// sparsh_dist_get_unique_id and passed on (MPI_Bcast, a file, ...).  The hierarchy is built by the host setup on every
// rank, the rank's row blocks are uploaded, the solve runs on the N GPUs, and x returns the GLOBAL solution on every rank.
// ---------------------------------------------------------------------------------------------------------
namespace {

enum DistMethod { D_AMG = 0, D_PCG = 1, D_PBICG = 2 };

void dist_driver(sp_matrix_mg &A, double *b, double *x, DistMethod m, int nranks, int rank, const char *id128, const char *label) {
    const sparsh::Options &o = options();
    auto check = [&](int rc, const char *what) {
        if (rc != SPARSH_OK) {
            std::fprintf(stderr, "sparsh_amg: %s (%s) failed (%d): %s\n", label, what, rc, sparsh_last_error());
            std::exit(1);  // as the single-GPU entry points: loudly, never by falling back to the CPU
        }
    };
    check(sparsh_init(o.device >= 0 ? o.device : rank), "sparsh_init");
    check(sparsh_dist_init(id128, nranks, rank), "sparsh_dist_init");
    const double t1 = omp_get_wtime();
    AMG_GPU1_solver *S = new AMG_GPU1_solver();
    S->AMG_solver_setup_jacobi(A);
    void *plan = sparsh_host_dist_plan(S, nranks, rank, o.tail_threshold);
    if (!plan) {
        std::fprintf(stderr, "sparsh_amg: %s: single-level hierarchy, nothing to distribute (use the single-GPU entry point)\n", label);
        std::exit(1);
    }
    sparsh_dist_t dh = (sparsh_dist_t)sparsh_host_dist_upload(plan);
    if (!dh) std::exit(1);
    const int *rows = nullptr;
    const int nl = sparsh_host_dist_plan_rows(plan, 0, &rows);
    const int n = A.nrow;
    std::vector<double> bl((size_t)nl), xl((size_t)nl);
    for (int i = 0; i < nl; i++) {
        bl[i] = b[rows[i]];
        xl[i] = x[rows[i]];
    }
    double *db = nullptr, *dx = nullptr;
    check(sparsh_malloc(sizeof(double) * ((size_t)nl + 2), (void **)&db), "malloc");
    check(sparsh_malloc(sizeof(double) * ((size_t)nl + 2), (void **)&dx), "malloc");
    const double t2 = omp_get_wtime();
    check(sparsh_memcpy_h2d(db, bl.data(), sizeof(double) * (size_t)nl), "h2d");
    check(sparsh_memcpy_h2d(dx, xl.data(), sizeof(double) * (size_t)nl), "h2d");
    double tol = o.tol;
    if (o.tol_mode != sparsh::TOL_ABSOLUTE) {  // ||b|| of the GLOBAL right-hand side: every rank holds it
        double s = 0.0;
        for (int i = 0; i < n; i++) s += b[i] * b[i];
        tol = o.tol * std::sqrt(s);
    }
    std::vector<double> hist((size_t)o.max_iter + 2, 0.0);
    int it = 0;
    int rc = m == D_AMG   ? sparsh_dist_amg_solve(dh, db, dx, tol, o.max_iter, hist.data(), &it)
             : m == D_PCG ? sparsh_dist_pcg(dh, db, dx, tol, o.max_iter, hist.data(), &it)
                          : sparsh_dist_pbicgstab(dh, db, dx, tol, o.max_iter, hist.data(), &it);
    if (rc != SPARSH_OK && rc != SPARSH_ERR_NOT_CONVERGED) check(rc, "solve");
    check(sparsh_memcpy_d2h(xl.data(), dx, sizeof(double) * (size_t)nl), "d2h");
    const double t3 = omp_get_wtime();
    check(sparsh_dist_allgather_rows(xl.data(), rows, nl, x, n), "gather");
    sparsh::Report &r = sparsh::last_report();
    r.iterations = it;
    r.converged = rc == SPARSH_OK;
    r.history.assign(hist.begin(), hist.begin() + it + 1);
    r.solve_seconds = t3 - t2;
    if (o.print_solve && rank == 0) {
        for (int k = 1; k <= it; k++) std::cout << k << "\t" << hist[k] << "\n";
        std::cout << label << " Setup Phase Time\t" << t2 - t1 << "\n";
        std::cout << label << " Solve Phase Time\t" << t3 - t2 << "\n";
    }
    sparsh_free(db);
    sparsh_free(dx);
    sparsh_host_dist_plan_free(plan);  // destroys the distributed device hierarchy too
    delete S;
}

}  // namespace

void AMG_Solver_MG(sp_matrix_mg &A, double *&b, double *&x, int nranks, int rank, const char *id128) {
    dist_driver(A, b, x, D_AMG, nranks, rank, id128, "AMG (multi-GPU)");
}
void Solver_PCG_MG(sp_matrix_mg &A, double *&b, double *&x, int nranks, int rank, const char *id128) {
    dist_driver(A, b, x, D_PCG, nranks, rank, id128, "PCG (multi-GPU)");
}
void Solver_PBiCG_MG(sp_matrix_mg &A, double *&b, double *&x, int nranks, int rank, const char *id128) {
    dist_driver(A, b, x, D_PBICG, nranks, rank, id128, "PBiCGStab (multi-GPU)");
}

