// ref_harness.cpp — extern "C" driver around the UNMODIFIED reference host sources (TEST INFRASTRUCTURE).
//
// Built by oracle/Makefile into oracle/_ref/libsparsh_ref.so together with /root/reference/src/*.cpp
// (compiled where they lie, never copied) and the MKL shim.  This file is our own code: it only calls the
// reference's public classes/functions and copies their results out so that Python tests can compare the
// C restatement (oracle/sparsh_oracle.c) and the CUDA path against "the reference run here".
//
// Header handling: oracle/Makefile generates _ref/include/AMG.hpp from the reference's AMG.hpp with sed,
// replacing three macro values so they can be chosen at run time / are large enough:
//   th     -> (sparsh_ref_num_threads())     tol1 -> (sparsh_ref_tol())     level1 -> 32
// Nothing else is touched.
#include <omp.h>

#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <limits>
#include <new>
#include <sstream>
#include <string>
#include <vector>

#include "AMG.hpp"
#include "AMG_coarsening.hpp"
#include "AMG_cycle_utilities.hpp"
#include "AMG_phases.hpp"
#include "AMG_smoothers.hpp"

// The reference hands an uninitialised z0 to its first preconditioner call (src/AMG_main_solvers.cpp:112,132;
// SURVEY Appendix B).  We cannot edit the source, so every array the reference allocates inside this library is
// zero-filled instead (the library is linked -Bsymbolic, the override stays private to it).
void *operator new[](std::size_t sz) {
    void *p = std::calloc(sz ? sz : 1, 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete[](void *p) noexcept { std::free(p); }
void operator delete[](void *p, std::size_t) noexcept { std::free(p); }

static int g_threads = 1;
static double g_tol = 1e-8;
static long g_tol_calls = 0;
static long g_tol_call_limit = -1;  // after this many evaluations tol1 reads +inf (bounded CPU samples)

extern "C" int sparsh_ref_num_threads(void) { return g_threads; }
extern "C" double sparsh_ref_tol(void) {
    if (g_tol_call_limit >= 0 && g_tol_calls++ >= g_tol_call_limit) return std::numeric_limits<double>::infinity();
    return g_tol;
}

// Stubs for the GPU classes AMG_main_solvers.cpp refers to (:193,242,258,497).  The reference's .cu files
// cannot be built with CUDA >= 11; nothing in the oracle ever calls these.
#include "AMG_gpu_phases.hpp"
#include "AMG_gpu_phases_2.hpp"
static void no_gpu() {
    std::cerr << "sparsh_ref: the reference GPU path is not buildable (cusparseDcsrmv removed in CUDA 11)\n";
    std::abort();
}
void AMG_GPU_solver::GPU_Allocations() { no_gpu(); }
void AMG_GPU_solver::AMG_GPU_solve(double *, double *, int) { no_gpu(); }
AMG_GPU_solver::~AMG_GPU_solver() {}
void AMG_GPU1_solver::GPU_Allocations() { no_gpu(); }
void AMG_GPU1_solver::helper(double *, double *, int) { no_gpu(); }
AMG_GPU1_solver::~AMG_GPU1_solver() {}

namespace {

struct CoutCapture {
    std::ostringstream ss;
    std::streambuf *old;
    std::streamsize oldprec;
    CoutCapture() {
        old = std::cout.rdbuf(ss.rdbuf());
        oldprec = std::cout.precision(17);
    }
    ~CoutCapture() {
        std::cout.rdbuf(old);
        std::cout.precision(oldprec);
    }
};

// lines of the form "<int><sep><double>" are residual-history lines (src/AMG_phases.cpp:223,
// src/AMG_main_solvers.cpp:156,441)
int parse_history(const std::string &text, double *hist, int maxhist) {
    std::istringstream in(text);
    std::string line;
    int k = 0;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string a, b, c;
        if (!(ls >> a >> b) || (ls >> c)) continue;
        if (a.find_first_not_of("0123456789") != std::string::npos) continue;
        char *end = nullptr;
        double val = std::strtod(b.c_str(), &end);
        if (end == b.c_str() || *end != '\0') continue;
        if (k < maxhist) hist[k] = val;
        k++;
    }
    return k;
}

sp_matrix_mg *make_matrix(int n, int nnz, const int *rp, const int *ci, const double *v) {
    sp_matrix_mg *A = new sp_matrix_mg(n, n, nnz);
    std::memcpy(A->rowptr, rp, sizeof(int) * ((size_t)n + 1));
    std::memcpy(A->colindex, ci, sizeof(int) * (size_t)nnz);
    std::memcpy(A->val, v, sizeof(double) * (size_t)nnz);
    A->sp_matrix_fill();           // main.cpp:21
    A->sp_matrix_fill_diagonal();  // main.cpp:22
    return A;
}

struct RefAmg {
    sp_matrix_mg *A = nullptr;
    AMG_solver *S = nullptr;
    double setup_seconds = 0.0;
};

}  // namespace

extern "C" {

void ref_set_threads(int nt) { g_threads = nt > 0 ? nt : 1; }
void ref_set_tol(double tol) { g_tol = tol; }
void ref_set_tol_call_limit(long limit) {
    g_tol_call_limit = limit;
    g_tol_calls = 0;
}
int ref_level1(void) { return level1; }

// coarsening 0: AMG_solver::AMG_solver_setup_jacobi exactly as shipped (HEM, src/AMG_phases.cpp:35-90).
// coarsening 1: the same loop with the commented alternative sequential::beck_prolongator (:63) switched in;
//               the loop is replayed here because the choice is made by editing the source.
void *ref_amg_setup(int n, int nnz, const int *rp, const int *ci, const double *v, int coarsening) {
    CoutCapture cap;
    RefAmg *h = new RefAmg();
    h->A = make_matrix(n, nnz, rp, ci, v);
    h->S = new AMG_solver();
    double t0 = omp_get_wtime();
    if (coarsening == 0) {
        h->S->AMG_solver_setup_jacobi(*h->A);
    } else {
        AMG_solver *S = h->S;
        int l = 0;
        S->Av[0] = h->A;
        S->Xv[0] = new double[n];
        S->Bv[0] = new double[n];
        S->Rv[0] = new double[n];
        while (S->Av[l]->nrow > limit_upper && l < (level1 - 1)) {
            sequential::beck_prolongator(*S->Av[l], S->Pv[l]);
            parallel::coarsen_matrix(*S->Av[l], S->Av[l + 1], *S->Pv[l]);
            l = l + 1;
            S->Xv[l] = new double[S->Av[l]->nrow]();
            S->Bv[l] = new double[S->Av[l]->nrow]();
            S->Rv[l] = new double[S->Av[l]->nrow]();
            if (S->Av[l]->nrow < limit_lower) break;
        }
        S->l = l;
        S->Directsolve = new Direct_Solver_Pardiso(*S->Av[l]);
    }
    h->setup_seconds = omp_get_wtime() - t0;
    return h;
}

// Adopt an externally built hierarchy ("same hierarchy, built once, shared with the GPU path") so that
// the CPU baseline can time the reference's own V-cycle / PCG loop at full size without repeating its sequential setup.
// Arrays are copied into reference objects; the coarsest level gets the reference's Direct_Solver_Pardiso.
void *ref_amg_from_levels(int nlevels, const int *nrow, const int *const *rp, const int *const *ci,
                          const double *const *v, const int *pncol, const int *const *prp, const int *const *pci,
                          const double *const *pv) {
    CoutCapture cap;
    RefAmg *h = new RefAmg();
    AMG_solver *S = h->S = new AMG_solver();
    for (int k = 0; k < nlevels; k++) {
        S->Av[k] = make_matrix(nrow[k], rp[k][nrow[k]], rp[k], ci[k], v[k]);
        S->Xv[k] = new double[nrow[k]]();
        S->Bv[k] = new double[nrow[k]]();
        S->Rv[k] = new double[nrow[k]]();
        if (k < nlevels - 1) {
            const int pn = prp[k][nrow[k]];
            sp_matrix_mg *P = new sp_matrix_mg(nrow[k], pncol[k], pn);
            std::memcpy(P->rowptr, prp[k], sizeof(int) * ((size_t)nrow[k] + 1));
            std::memcpy(P->colindex, pci[k], sizeof(int) * (size_t)pn);
            std::memcpy(P->val, pv[k], sizeof(double) * (size_t)pn);
            P->sp_matrix_fill();
            S->Pv[k] = P;
        }
    }
    h->A = S->Av[0];
    S->l = nlevels - 1;
    S->Directsolve = new Direct_Solver_Pardiso(*S->Av[S->l]);
    return h;
}

double ref_amg_setup_seconds(void *hv) { return ((RefAmg *)hv)->setup_seconds; }
int ref_amg_nlevels(void *hv) { return ((RefAmg *)hv)->S->l + 1; }

void ref_amg_level_dims(void *hv, int lvl, int *nrow, int *nnz, int *p_ncol, int *p_nnz) {
    AMG_solver *S = ((RefAmg *)hv)->S;
    *nrow = S->Av[lvl]->nrow;
    *nnz = S->Av[lvl]->rowptr[S->Av[lvl]->nrow];
    if (lvl < S->l) {
        *p_ncol = S->Pv[lvl]->ncol;
        *p_nnz = S->Pv[lvl]->rowptr[S->Pv[lvl]->nrow];
    } else {
        *p_ncol = 0;
        *p_nnz = 0;
    }
}

void ref_amg_level_copy(void *hv, int lvl, int *rp, int *ci, double *v, double *diag, int *prp, int *pci,
                        double *pv) {
    AMG_solver *S = ((RefAmg *)hv)->S;
    sp_matrix_mg *A = S->Av[lvl];
    int nnz = A->rowptr[A->nrow];
    std::memcpy(rp, A->rowptr, sizeof(int) * ((size_t)A->nrow + 1));
    std::memcpy(ci, A->colindex, sizeof(int) * (size_t)nnz);
    std::memcpy(v, A->val, sizeof(double) * (size_t)nnz);
    std::memcpy(diag, A->diagonal, sizeof(double) * (size_t)A->nrow);
    if (lvl < S->l) {
        sp_matrix_mg *P = S->Pv[lvl];
        int pnnz = P->rowptr[P->nrow];
        std::memcpy(prp, P->rowptr, sizeof(int) * ((size_t)P->nrow + 1));
        std::memcpy(pci, P->colindex, sizeof(int) * (size_t)pnnz);
        std::memcpy(pv, P->val, sizeof(double) * (size_t)pnnz);
    }
}

// AMG_solver::AMG_solve_jacobi(b, x, cycles) with cycles > 0 (src/AMG_phases.cpp:163-192)
void ref_amg_vcycle(void *hv, const double *b, double *x, int cycles) {
    CoutCapture cap;
    AMG_solver *S = ((RefAmg *)hv)->S;
    double *bb = const_cast<double *>(b);
    S->AMG_solve_jacobi(bb, x, cycles);
}

// AMG_solve_jacobi(b, x, -1): loop until ||r|| <= tol1 (src/AMG_phases.cpp:194-226); history parsed from the
// reference's own prints at 17 significant digits.  Returns the number of cycles.
int ref_amg_solve(void *hv, const double *b, double *x, double *hist, int maxhist) {
    CoutCapture cap;
    AMG_solver *S = ((RefAmg *)hv)->S;
    double *bb = const_cast<double *>(b);
    S->AMG_solve_jacobi(bb, x, -1);
    return parse_history(cap.ss.str(), hist, maxhist);
}

double ref_residual(void *hv, int lvl, const double *b, const double *x) {
    AMG_solver *S = ((RefAmg *)hv)->S;
    double *bb = const_cast<double *>(b), *xx = const_cast<double *>(x);
    return parallel::residual(*S->Av[lvl], bb, xx);
}
void ref_jacobi(void *hv, int lvl, const double *b, double *x, int iteration) {
    AMG_solver *S = ((RefAmg *)hv)->S;
    double *bb = const_cast<double *>(b);
    parallel::jacobi_smoother(*S->Av[lvl], bb, x, iteration);
}
void ref_store_residual(void *hv, int lvl, const double *b, const double *x, double *r) {
    AMG_solver *S = ((RefAmg *)hv)->S;
    double *bb = const_cast<double *>(b), *xx = const_cast<double *>(x);
    parallel::store_residual(*S->Av[lvl], bb, xx, r);
}
void ref_transfer_residual(void *hv, int lvl, const double *r, double *bc) {
    AMG_solver *S = ((RefAmg *)hv)->S;
    double *rr = const_cast<double *>(r);
    parallel::transfer_residual(*S->Pv[lvl], rr, bc);
}
void ref_transfer_solution(void *hv, int lvl, const double *xc, double *xf) {
    AMG_solver *S = ((RefAmg *)hv)->S;
    double *xx = const_cast<double *>(xc);
    parallel::transfer_solution(*S->Pv[lvl], xx, xf);
}
void ref_coarse_solve(void *hv, const double *b, double *x) {
    AMG_solver *S = ((RefAmg *)hv)->S;
    double *bb = const_cast<double *>(b);
    S->Directsolve->Direct_Solver_Pardiso_solve(bb, x);
}
void ref_spmv(void *hv, int lvl, const double *x, double *y) {
    AMG_solver *S = ((RefAmg *)hv)->S;
    sp_matrix_mg *A = S->Av[lvl];
    mkl_set_num_threads(g_threads);
    mkl_sparse_d_mv(SPARSE_OPERATION_NON_TRANSPOSE, 1.0, A->A1, A->des, x, 0.0, y);
}

// m iterations of the Solver_PCG_1 loop body (src/AMG_main_solvers.cpp:136-152) on an existing hierarchy:
// the bounded CPU sample of bench.py.  The preconditioner is the reference's own AMG_solve_jacobi(r,z,1);
// SpMV/BLAS-1 go through the same shim entry points the reference calls.  Returns seconds for the m
// iterations (the initial residual + first preconditioner call, :124-134, are outside the timed region).
// tol < 0: exactly m iterations (bounded sample).  tol >= 0: the reference's own stopping rule, `count++ < nrow && r1 > tol`
// (:136), capped at m; *iters receives the number of iterations done.
static double ref_pcg_run(void *hv, const double *b, double *x, int m, double tol, double *hist, int *iters) {
    CoutCapture cap;
    RefAmg *h = (RefAmg *)hv;
    sp_matrix_mg &A = *h->A;
    int n = A.nrow;
    double *Ap = new double[n], *p = new double[n], *z0 = new double[n], *r0 = new double[n];
    mkl_set_num_threads(g_threads);
    mkl_sparse_d_mv(SPARSE_OPERATION_NON_TRANSPOSE, 1.0, A.A1, A.des, x, 0.0, r0);
    cblas_daxpby(n, 1.0, b, 1, -1.0, r0, 1);
    double r1 = cblas_dnrm2(n, r0, 1);
    if (hist) hist[0] = r1;
    std::fill(z0, z0 + n, 0);  // the reference leaves z0 uninitialised here
    h->S->AMG_solve_jacobi(r0, z0, 1);
    std::copy(z0, z0 + n, p);
    double t0 = omp_get_wtime();
    int it = 0;
    while (it < m && (tol < 0.0 || r1 > tol)) {
        it++;
        mkl_sparse_d_mv(SPARSE_OPERATION_NON_TRANSPOSE, 1.0, A.A1, A.des, p, 0.0, Ap);
        double alpha = cblas_ddot(n, p, 1, Ap, 1);
        double s = cblas_ddot(n, r0, 1, z0, 1);
        alpha = s / alpha;
        cblas_daxpy(n, alpha, p, 1, x, 1);
        cblas_daxpy(n, -alpha, Ap, 1, r0, 1);
        std::fill(z0, z0 + n, 0);
        h->S->AMG_solve_jacobi(r0, z0, 1);
        double beta = cblas_ddot(n, z0, 1, r0, 1) / s;
        cblas_daxpby(n, 1.0, z0, 1, beta, p, 1);
        r1 = cblas_dnrm2(n, r0, 1);
        if (hist) hist[it] = r1;
    }
    double t = omp_get_wtime() - t0;
    if (iters) *iters = it;
    delete[] Ap;
    delete[] p;
    delete[] z0;
    delete[] r0;
    return t;
}
double ref_pcg_sample(void *hv, const double *b, double *x, int m, double *hist) {
    return ref_pcg_run(hv, b, x, m, -1.0, hist, nullptr);
}
// the same loop run to convergence (||r|| <= tol, at most max_iter iterations): bench.py's converged reference solve
double ref_pcg_solve(void *hv, const double *b, double *x, double tol, int max_iter, double *hist, int *iters) {
    return ref_pcg_run(hv, b, x, max_iter, tol, hist, iters);
}

// Whole reference solvers by name, history captured from their own prints.  x is in/out.
// name: AMG_Solver_CPU_baseline | AMG_Solver_2 | Solver_CG_1 | Solver_PCG_1 | Solver_BiCG_1 | Solver_PBiCG_1
int ref_solve(const char *name, int n, int nnz, const int *rp, const int *ci, const double *v, const double *b,
              double *x, double *hist, int maxhist) {
    CoutCapture cap;
    sp_matrix_mg *A = make_matrix(n, nnz, rp, ci, v);
    double *bb = new double[n];
    std::memcpy(bb, b, sizeof(double) * (size_t)n);
    double *xx = x;
    std::string s(name);
    if (s == "AMG_Solver_CPU_baseline")
        AMG_Solver_CPU_baseline(*A, bb, xx);
    else if (s == "AMG_Solver_2")
        AMG_Solver_2(*A, bb, xx);
    else if (s == "Solver_CG_1")
        Solver_CG_1(*A, bb, xx);
    else if (s == "Solver_PCG_1")
        Solver_PCG_1(*A, bb, xx);
    else if (s == "Solver_BiCG_1")
        Solver_BiCG_1(*A, bb, xx);
    else if (s == "Solver_PBiCG_1")
        Solver_PBiCG_1(*A, bb, xx);
    else
        return -1;
    delete[] bb;
    return parse_history(cap.ss.str(), hist, maxhist);
}

// sp_matrix_mg::color_matrix_and_reorder (src/AMG_cpu_matrix.cpp:81-199).  Outputs: perm[new] = old (the
// reference's reused `color` array), color_count[0..total_colors] prefix offsets, and the permuted CSR.
int ref_color_reorder(int n, int nnz, const int *rp, const int *ci, const double *v, int *perm, int *color_count,
                      int *qrp, int *qci, double *qv, double *qdiag) {
    CoutCapture cap;
    sp_matrix_mg *A = make_matrix(n, nnz, rp, ci, v);
    A->color_matrix_and_reorder();
    std::memcpy(perm, A->color, sizeof(int) * (size_t)n);
    std::memcpy(color_count, A->color_count, sizeof(int) * ((size_t)A->total_colors + 1));
    std::memcpy(qrp, A->rowptr, sizeof(int) * ((size_t)n + 1));
    std::memcpy(qci, A->colindex, sizeof(int) * (size_t)A->rowptr[n]);
    std::memcpy(qv, A->val, sizeof(double) * (size_t)A->rowptr[n]);
    std::memcpy(qdiag, A->diagonal, sizeof(double) * (size_t)n);
    return A->total_colors;
}

// parallel::sor_smoother on an already colour-permuted matrix (src/AMG_smoothers.cpp:78-102)
void ref_sor(int n, int nnz, const int *rp, const int *ci, const double *v, const int *color_count, int total_colors,
             const double *b, double *x, int iteration) {
    sp_matrix_mg *A = make_matrix(n, nnz, rp, ci, v);
    A->color_count = new int[total_colors + 1];
    std::memcpy(A->color_count, color_count, sizeof(int) * ((size_t)total_colors + 1));
    A->total_colors = total_colors;
    double *bb = const_cast<double *>(b);
    parallel::sor_smoother(*A, bb, x, iteration);
}

// readcoo (src/AMG_file_read.cpp:39-72): used once, here in the build container, by
// tests/golden/make_golden.py to turn the bundled fixture into a committed .npz.
int ref_readcoo_dims(const char *matrixfile, const char *rhsfile, int *n, int *nnz) {
    sp_matrix_mg *A = new sp_matrix_mg();
    double *b = nullptr;
    readcoo(const_cast<char *>(matrixfile), const_cast<char *>(rhsfile), A, b);
    *n = A->nrow;
    *nnz = A->nnz;
    return 0;
}
int ref_readcoo(const char *matrixfile, const char *rhsfile, int *rp, int *ci, double *v, double *bout) {
    sp_matrix_mg *A = new sp_matrix_mg();
    double *b = nullptr;
    readcoo(const_cast<char *>(matrixfile), const_cast<char *>(rhsfile), A, b);
    std::memcpy(rp, A->rowptr, sizeof(int) * ((size_t)A->nrow + 1));
    std::memcpy(ci, A->colindex, sizeof(int) * (size_t)A->nnz);
    std::memcpy(v, A->val, sizeof(double) * (size_t)A->nnz);
    std::memcpy(bout, b, sizeof(double) * (size_t)A->nrow);
    return 0;
}

}  // extern "C"
