// solvers.cpp — the reference's solver objects and entry points, running on the B200 through the C-ABI.
//
//   AMG_solver::AMG_solve_jacobi / AMG_solve_SOR          reference src/AMG_phases.cpp:151-306
//   AMG_GPU1_solver ("MI")                                 reference src/AMG_gpu_phases_2.cu:13-338
//   AMG_GPU_solver  ("CI")                                 reference src/AMG_gpu_phases.cu:13-639
//   sp_matrix_gpu                                          reference src/AMG_gpu_matrix.cu:26-142
//   AMG_Solver_* / Solver_CG_* / Solver_PCG_* / Solver_PBiCG_*   reference src/AMG_main_solvers.cpp:14-580,
//                                                          src/AMG_main_solvers.cu:35-763
// Host code only orchestrates: every vector operation, SpMV, smoother sweep, transfer and the coarse solve run in
// libsparsh_b200.so.  There is no CPU arithmetic path here; a missing GPU makes every entry point fail loudly.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "../../include/sparsh_b200.h"
#include "sparsh_amg.hpp"

using sparsh::last_report;
using sparsh::options;

namespace {

[[noreturn]] void die(const char *where, int rc) {
    // the reference's only fatal path is exit(1) from the coarse solver (src/AMG_coarse_level_solver.cpp:56,74);
    // a missing GPU / failed CUDA call is treated the same way: loudly, never by falling back to the CPU
    std::fprintf(stderr, "sparsh_amg: %s failed (%d): %s\n", where, rc, sparsh_last_error());
    std::exit(1);
}
#define CK(call)                                   \
    do {                                           \
        int rc__ = (call);                         \
        if (rc__ != SPARSH_OK) die(#call, rc__);   \
    } while (0)

double effective_tol(const double *b, int n) {
    const sparsh::Options &o = options();
    if (o.tol_mode == sparsh::TOL_ABSOLUTE) return o.tol;
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) num_threads(o.threads)
    for (int i = 0; i < n; i++) s += b[i] * b[i];
    return o.tol * std::sqrt(s);
}

void record(int rc, int iters, const std::vector<double> &hist, double seconds, const char *label, int first_index) {
    sparsh::Report &r = last_report();
    r.iterations = iters;
    r.converged = rc == SPARSH_OK;
    r.history.assign(hist.begin(), hist.begin() + iters + 1);
    r.solve_seconds = seconds;
    if (options().print_solve)
        for (int k = 1; k <= iters; k++) std::cout << (k - 1 + first_index) << "\t" << hist[k] << "\n";
    (void)label;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// AMG_solver
// ---------------------------------------------------------------------------------------------------------
void AMG_solver::upload() {
    if (device) return;
    const double t0 = omp_get_wtime();
    const sparsh::Options &o = options();
    std::vector<sparsh_level_desc> d((size_t)l + 1);
    const bool sor = Av[0]->color_count != nullptr && Av[0]->total_colors > 0;
    for (int k = 0; k <= l; k++) {
        std::memset(&d[k], 0, sizeof(sparsh_level_desc));
        sp_matrix_mg *A = Av[k];
        d[k].nrow = A->nrow;
        d[k].nnz = A->rowptr[A->nrow];
        d[k].rowptr = A->rowptr;
        d[k].colindex = A->colindex;
        d[k].val = A->val;
        d[k].diag = A->diagonal;
        if (sor) {
            d[k].total_colors = A->total_colors;
            d[k].color_count = A->color_count;
        }
        if (k < l) {
            sp_matrix_mg *P = Pv[k];
            d[k].p_ncol = P->ncol;
            d[k].p_nnz = P->rowptr[P->nrow];
            d[k].p_rowptr = P->rowptr;
            d[k].p_colindex = P->colindex;
            d[k].p_val = P->val;
        }
    }
    sparsh_params prm;
    sparsh_params_default(&prm);
    prm.omega = o.relax;
    prm.use_graph = o.use_graph;
    if (sor) {
        prm.smoother = 1;
        prm.pre_sweeps = prm.post_sweeps = 6;  // literal 6 in the reference (src/AMG_phases.cpp:251,265,281,295)
    } else {
        prm.pre_sweeps = prm.post_sweeps = o.sweeps;
    }
    CK(sparsh_hierarchy_create(l + 1, d.data(), &prm, &device));
    last_report().upload_seconds = omp_get_wtime() - t0;
}

static void run_amg(AMG_solver &S, double *b, double *x, int iterations, bool on_device) {
    S.upload();
    const int n = S.Av[0]->nrow;
    const sparsh::Options &o = options();
    const double t0 = omp_get_wtime();
    if (iterations > 0) {
        // exactly `iterations` cycles (preconditioner use): reference src/AMG_phases.cpp:163-192
        if (on_device) {
            CK(sparsh_hierarchy_vcycle(S.device, b, x, iterations, 0));
        } else {
            double *db = nullptr, *dx = nullptr;
            CK(sparsh_malloc(sizeof(double) * (size_t)n, (void **)&db));
            CK(sparsh_malloc(sizeof(double) * (size_t)n, (void **)&dx));
            CK(sparsh_memcpy_h2d(db, b, sizeof(double) * (size_t)n));
            CK(sparsh_memcpy_h2d(dx, x, sizeof(double) * (size_t)n));
            CK(sparsh_hierarchy_vcycle(S.device, db, dx, iterations, 0));
            CK(sparsh_memcpy_d2h(x, dx, sizeof(double) * (size_t)n));
            sparsh_free(db);
            sparsh_free(dx);
        }
        last_report().solve_seconds = omp_get_wtime() - t0;
        return;
    }
    // iterations == -1: until ||r|| <= tol (reference src/AMG_phases.cpp:194-226)
    std::vector<double> hist((size_t)o.max_iter + 2, 0.0);
    int it = 0, rc;
    if (on_device) {
        std::vector<double> hb((size_t)n);
        double tol = o.tol;
        if (o.tol_mode == sparsh::TOL_RELATIVE) {
            CK(sparsh_memcpy_d2h(hb.data(), b, sizeof(double) * (size_t)n));
            tol = effective_tol(hb.data(), n);
        }
        rc = sparsh_hierarchy_amg_solve(S.device, b, x, tol, o.max_iter, hist.data(), &it);
    } else {
        rc = sparsh_hierarchy_solve_host(S.device, 0, b, x, effective_tol(b, n), o.max_iter, hist.data(), &it);
    }
    if (rc != SPARSH_OK && rc != SPARSH_ERR_NOT_CONVERGED) die("AMG solve", rc);
    record(rc, it, hist, omp_get_wtime() - t0, "amg", 1);  // the CPU path prints 1-based cycles (:218-223)
}

void AMG_solver::AMG_solve_jacobi(double *&b, double *&x, int iterations) { run_amg(*this, b, x, iterations, false); }
void AMG_solver::AMG_solve_SOR(double *&b, double *&x, int iterations) { run_amg(*this, b, x, iterations, false); }

AMG_solver::~AMG_solver() {
    if (torn_down) return;  // explicit destructor call followed by a real one must be harmless
    torn_down = true;
    if (device) sparsh_hierarchy_destroy(device);
    device = nullptr;
    if (shared_mapping) {
        // arrays belong to read-only file mappings shared between ranks: release the mappings, never delete[] into them
        extern void sparsh_release_shared_hierarchy(AMG_solver *);
        sparsh_release_shared_hierarchy(this);
    } else if (Av) {
        for (int q = l; q > 0; q--) {
            if (Av[q]) {
                delete[] Av[q]->rowptr;
                delete[] Av[q]->colindex;
                delete[] Av[q]->val;
                delete Av[q];
            }
            if (Pv[q - 1]) {
                delete[] Pv[q - 1]->rowptr;
                delete[] Pv[q - 1]->colindex;
                delete[] Pv[q - 1]->val;
                delete Pv[q - 1];
            }
        }
    }
    delete[] Av;
    delete[] Pv;
    delete[] Xv;
    delete[] Bv;
    delete[] Rv;
    Av = Pv = nullptr;
    Xv = Bv = Rv = nullptr;
}

// ---------------------------------------------------------------------------------------------------------
// AMG_GPU1_solver ("MI") and AMG_GPU_solver ("CI"): one engine, hierarchy resident in HBM
// ---------------------------------------------------------------------------------------------------------
void AMG_GPU1_solver::GPU_Allocations() { upload(); }
void AMG_GPU1_solver::helper(double *b, double *x, int iterations) { run_amg(*this, b, x, iterations, false); }
void AMG_GPU1_solver::AMG_Solve(double *b, double *x, int iterations) { run_amg(*this, b, x, iterations, true); }
AMG_GPU1_solver::~AMG_GPU1_solver() {}

void AMG_GPU_solver::GPU_Allocations() { upload(); }
void AMG_GPU_solver::AMG_GPU_solve(double *b, double *x, int iterations) { run_amg(*this, b, x, iterations, false); }
void AMG_GPU_solver::AMG_GPU_solve_1(double *b, double *x, int iterations) { run_amg(*this, b, x, iterations, true); }
AMG_GPU_solver::~AMG_GPU_solver() {}

// ---------------------------------------------------------------------------------------------------------
// sp_matrix_gpu
// ---------------------------------------------------------------------------------------------------------
sp_matrix_gpu::sp_matrix_gpu(sp_matrix_mg &A) : nrow(A.nrow), ncol(A.ncol), nnz(A.rowptr[A.nrow]) {}
void sp_matrix_gpu::matrix_transfer_gpu(sp_matrix_mg &A, void *stream) {
    if (stream) CK(sparsh_set_stream(stream));
    if (!handle) CK(sparsh_matrix_create(A.nrow, A.ncol, A.rowptr[A.nrow], A.rowptr, A.colindex, A.val, A.diagonal, &handle));
}
void sp_matrix_gpu::smooth_jacobi(double *bgpu, double *xgpu, double *hgpu, void *stream, int steps) {
    if (stream) CK(sparsh_set_stream(stream));
    CK(sparsh_jacobi(handle, bgpu, xgpu, hgpu, options().relax, steps));
}
sp_matrix_gpu::~sp_matrix_gpu() {
    if (handle) sparsh_matrix_destroy(handle);
    handle = nullptr;
}

// ---------------------------------------------------------------------------------------------------------
// entry points
// ---------------------------------------------------------------------------------------------------------
namespace {

enum Method { M_AMG = 0, M_PCG = 1, M_PBICG = 2, M_PGMRES = 3 };

// setup + upload + solve with host b/x, timing printed like the reference's wrappers
void amg_driver(sp_matrix_mg &A, double *b, double *x, Method m, bool sor, const char *label) {
    const sparsh::Options &o = options();
    AMG_GPU1_solver *S = new AMG_GPU1_solver();
    const double t1 = omp_get_wtime();
    if (sor) {
        S->AMG_solver_setup_SOR(A);
        parallel::reorder_rhs(A, b);  // reference src/AMG_main_solvers.cpp:35
        parallel::reorder_rhs(A, x);
    } else {
        S->AMG_solver_setup_jacobi(A);
    }
    S->GPU_Allocations();
    const double t2 = omp_get_wtime();
    std::vector<double> hist((size_t)o.max_iter + 2, 0.0);
    int it = 0;
    const int n = A.nrow;
    int rc = sparsh_hierarchy_solve_host(S->device, (int)m, b, x, effective_tol(b, n), o.max_iter, hist.data(), &it);
    if (rc != SPARSH_OK && rc != SPARSH_ERR_NOT_CONVERGED) die(label, rc);
    const double t3 = omp_get_wtime();
    record(rc, it, hist, t3 - t2, label, m == M_PBICG ? 0 : 1);
    if (sor) {
        // undo the colour permutation: the reference applies the forward permutation again (src/AMG_main_solvers.cpp:38-39,
        // a defect: SURVEY Appendix B); the inverse is what returns x and b to the caller's ordering
        std::vector<double> t((size_t)n);
        for (int i = 0; i < n; i++) t[A.color[i]] = x[i];
        std::copy(t.begin(), t.end(), x);
        for (int i = 0; i < n; i++) t[A.color[i]] = b[i];
        std::copy(t.begin(), t.end(), b);
    }
    if (o.print_solve) {
        std::cout << label << " Setup Phase Time\t" << t2 - t1 << "\n";
        std::cout << label << " Solve Phase Time\t" << t3 - t2 << "\n";
        std::cout << label << " Total Time\t      " << t3 - t1 << "\n";
    }
    delete S;
}

enum Plain { K_CG = 0, K_BICG = 1, K_GMRES = 2 };
void krylov_driver(sp_matrix_mg &A, double *b, double *x, Plain which, const char *label) {
    const bool cg = which == K_CG;
    const sparsh::Options &o = options();
    const int n = A.nrow;
    sparsh_matrix_t dA = nullptr;
    double *db = nullptr, *dx = nullptr;
    if (!A.diagonal) A.sp_matrix_fill_diagonal();
    CK(sparsh_matrix_create(n, n, A.rowptr[n], A.rowptr, A.colindex, A.val, A.diagonal, &dA));
    CK(sparsh_malloc(sizeof(double) * (size_t)n, (void **)&db));
    CK(sparsh_malloc(sizeof(double) * (size_t)n, (void **)&dx));
    const double t0 = omp_get_wtime();
    CK(sparsh_memcpy_h2d(db, b, sizeof(double) * (size_t)n));
    CK(sparsh_memcpy_h2d(dx, x, sizeof(double) * (size_t)n));
    std::vector<double> hist((size_t)o.max_iter + 2, 0.0);
    int it = 0;
    const double tol = effective_tol(b, n);
    int rc = cg                 ? sparsh_cg(dA, db, dx, tol, o.max_iter, hist.data(), &it)
             : which == K_BICG ? sparsh_bicgstab(dA, db, dx, tol, o.max_iter, hist.data(), &it)
                               : sparsh_gmres(dA, db, dx, tol, o.gmres_restart, o.max_iter, hist.data(), &it);
    if (rc != SPARSH_OK && rc != SPARSH_ERR_NOT_CONVERGED) die(label, rc);
    CK(sparsh_memcpy_d2h(x, dx, sizeof(double) * (size_t)n));
    record(rc, it, hist, omp_get_wtime() - t0, label, cg ? 1 : 0);
    sparsh_free(db);
    sparsh_free(dx);
    sparsh_matrix_destroy(dA);
}

}  // namespace

// AMG as solver.  The *_CPU_baseline / _1 names keep their place in the API; in this library they run the same device
// engine (there is no CPU solve path) with the CPU reference's semantics — options().sweeps = 7 by default.
void AMG_Solver_CPU_baseline(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_AMG, false, "AMG"); }
void AMG_Solver_1(sp_matrix_mg &A, double *&b, double *&x) { AMG_Solver_CPU_baseline(A, b, x); }
void AMG_Solver_2(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_AMG, true, "AMG-SOR"); }
void AMG_Solver_CPU_GPU_CI(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_AMG, false, "Time AMG Hybrid AMG 1"); }
void AMG_Solver_CPU_GPU_MI(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_AMG, false, "Time AMG Hybrid AMG 2"); }

void Solver_CG_1(sp_matrix_mg &A, double *&b, double *&x) { krylov_driver(A, b, x, K_CG, "CG"); }
void Solver_CG_2(sp_matrix_mg &A, double *&b, double *&x) { krylov_driver(A, b, x, K_CG, "CG"); }
void Solver_BiCG_1(sp_matrix_mg &A, double *&b, double *&x) { krylov_driver(A, b, x, K_BICG, "BiCGStab"); }
// additions (the reference's README advertises GMRES, its sources have none: SURVEY F3, §8f.2)
void Solver_GMRES_1(sp_matrix_mg &A, double *&b, double *&x) { krylov_driver(A, b, x, K_GMRES, "GMRES"); }
void Solver_PGMRES_1(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_PGMRES, false, "PGMRES"); }

void Solver_PCG_1(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_PCG, false, "PCG"); }
void Solver_PCG_2(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_PCG, false, "PCG-2"); }
void Solver_PCG_3(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_PCG, false, "PCG-3"); }
void Solver_PCG_4(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_PCG, false, "PCG-4"); }

// all four names run Solver_PBiCG_1's arithmetic: the reference's _2/_3/_4 variants are defective
void Solver_PBiCG_1(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_PBICG, false, "PBiCGStab"); }
void Solver_PBiCG_2(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_PBICG, false, "PBiCGStab-2"); }
void Solver_PBiCG_3(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_PBICG, false, "PBiCGStab-3"); }
void Solver_PBiCG_4(sp_matrix_mg &A, double *&b, double *&x) { amg_driver(A, b, x, M_PBICG, false, "PBiCGStab-4"); }

// reference src/AMG_main_solvers.cpp:566-580: colour A, then 100 single multicolour-SOR sweeps printing the residual
void coarsening_2(sp_matrix_mg &A, double *&b, double *&x) {
    A.color_matrix_and_reorder();
    const int n = A.nrow;
    sparsh_matrix_t dA = nullptr;
    double *db = nullptr, *dx = nullptr;
    CK(sparsh_matrix_create(n, n, A.rowptr[n], A.rowptr, A.colindex, A.val, A.diagonal, &dA));
    CK(sparsh_malloc(sizeof(double) * (size_t)n, (void **)&db));
    CK(sparsh_malloc(sizeof(double) * (size_t)n, (void **)&dx));
    CK(sparsh_memcpy_h2d(db, b, sizeof(double) * (size_t)n));
    CK(sparsh_memcpy_h2d(dx, x, sizeof(double) * (size_t)n));
    std::vector<double> hist(101, 0.0);
    for (int count = 1; count <= 100; count++) {
        CK(sparsh_mc_sor(dA, A.color_count, A.total_colors, db, dx, options().relax, 1));
        CK(sparsh_residual_norm(dA, db, dx, &hist[count]));
        if (options().print_solve) std::cout << count << "\t" << hist[count] << std::endl;
    }
    CK(sparsh_memcpy_d2h(x, dx, sizeof(double) * (size_t)n));
    sparsh::Report &r = last_report();
    r.iterations = 100;
    r.history = hist;
    sparsh_free(db);
    sparsh_free(dx);
    sparsh_matrix_destroy(dA);
}
