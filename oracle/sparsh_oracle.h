/*
 * sparsh_oracle.h — CPU restatement of the SParSH-AMG solve phase (and the
 * host setup that feeds it).  TEST INFRASTRUCTURE ONLY: nothing outside
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product path never does.
 *
 * Parity pinning: the reference ships no golden outputs.  This
 * restatement is pinned against the reference's own host sources compiled
 * unmodified here against an OpenMP MKL shim (oracle/_ref, see oracle/Makefile)
 * on the bundled fixture and on synthetic Poisson matrices; the resulting
 * histories are frozen in tests/golden/.  Real Intel MKL is absent from this
 * image, so "reference" always means "reference sources + our MKL shim".
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose arithmetic and evaluation order it follows.
 */
#ifndef SPARSH_ORACLE_H_
#define SPARSH_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- threads (reference: macro th, include/AMG.hpp:15) ---- */
void so_set_threads(int nt);
int so_get_threads(void);

/* ---- cycle primitives ---- */
/* y = A x                      (mkl_sparse_d_mv N, alpha=1, beta=0) */
void so_spmv(int nrow, const int *rp, const int *ci, const double *v, const double *x, double *y);
/* y(ncol) = A^T x              (mkl_sparse_d_mv T; src/AMG_cycle_utilities.cpp:102) */
void so_spmv_t(int nrow, int ncol, const int *rp, const int *ci, const double *v, const double *x, double *y);
/* parallel::jacobi_smoother    (src/AMG_smoothers.cpp:53-76): iteration+1 sweeps */
void so_jacobi(int n, const int *rp, const int *ci, const double *v, const double *diag, const double *b, double *x,
               double *helper, double omega, int iteration);
/* parallel::residual           (src/AMG_cycle_utilities.cpp:83-94): ||A x - b||_2 */
double so_residual(int n, const int *rp, const int *ci, const double *v, const double *b, const double *x,
                   double *helper);
/* parallel::store_residual     (src/AMG_cycle_utilities.cpp:115-123): r = b - A x */
void so_store_residual(int n, const int *rp, const int *ci, const double *v, const double *b, const double *x,
                       double *r);
/* parallel::transfer_residual  (src/AMG_cycle_utilities.cpp:97-104): bc = P^T r */
void so_transfer_residual(int nf, int nc, const int *prp, const int *pci, const double *pv, const double *r,
                          double *bc);
/* parallel::transfer_solution  (src/AMG_cycle_utilities.cpp:107-112): xf += P xc */
void so_transfer_solution(int nf, const int *prp, const int *pci, const double *pv, const double *xc, double *xf);
/* parallel::sor_smoother       (src/AMG_smoothers.cpp:78-102): multicolour SOR on the colour-permuted system */
void so_sor_multicolor(int n, const int *rp, const int *ci, const double *v, const double *diag,
                       const int *color_count, int total_colors, const double *b, double *x, double *helper,
                       double omega, int iteration);
/* BLAS-1 with a fixed, thread-count independent summation tree */
double so_dot(int n, const double *x, const double *y);
double so_nrm2(int n, const double *x);

/* ---- setup restatement ---- */
/* sp_matrix_fill_diagonal      (src/AMG_cpu_matrix.cpp:35-51) */
void so_fill_diagonal(int n, const int *rp, const int *ci, const double *v, double *diag);
/* sequential::HEM_Prolongator  (src/AMG_coarsening.cpp:14-97): agg[i] = aggregate of row i; returns ncoarse */
int so_hem(int n, const int *rp, const int *ci, const double *v, int level, int *agg);
/* sequential::beck_prolongator (src/AMG_coarsening.cpp:269-339): returns malloc'ed P (n x *nc) */
void so_beck(int n, const int *rp, const int *ci, int *nc, int **prp, int **pci, double **pv);
/* parallel::coarsen_matrix     (src/AMG_cycle_utilities.cpp:126-146): Ac = P^T (A P), columns sorted */
void so_rap(int n, const int *rp, const int *ci, const double *v, int nc, const int *prp, const int *pci,
            const double *pv, int **crp, int **cci, double **cv);
/* color_matrix_and_reorder     (src/AMG_cpu_matrix.cpp:81-199): perm[new]=old, color_count[0..total], permuted
 * matrix returned malloc'ed; returns total_colors */
int so_color_reorder(int n, const int *rp, const int *ci, const double *v, int *perm, int *color_count,
                     int **qrp, int **qci, double **qv);
void so_free(void *p);

/* ---- coarse direct solve (stand-in for PARDISO phases 12/33, src/AMG_coarse_level_solver.cpp:9-76) ---- */
typedef struct so_lu so_lu;
so_lu *so_lu_factor(int n, const int *rp, const int *ci, const double *v);
void so_lu_solve(const so_lu *f, const double *b, double *x);
void so_lu_free(so_lu *f);
int so_lu_bandwidth(const so_lu *f);

/* ---- hierarchy + V-cycle + Krylov ---- */
typedef struct so_amg so_amg;
/* AMG_solver::AMG_solver_setup_jacobi (src/AMG_phases.cpp:35-90).  coarsening: 0 = HEM (shipped default),
 * 1 = Beck (the commented alternative, src/AMG_phases.cpp:63).  max_levels plays level1 (AMG.hpp:21). */
so_amg *so_amg_setup(int n, const int *rp, const int *ci, const double *v, int coarsening, int max_levels,
                     int limit_upper, int limit_lower);
/* adopt an externally built hierarchy (arrays are copied) */
so_amg *so_amg_from_levels(int nlevels, const int *nrow, const int *const *rp, const int *const *ci,
                           const double *const *v, const int *pncol, const int *const *prp, const int *const *pci,
                           const double *const *pv);
void so_amg_free(so_amg *h);
int so_amg_nlevels(const so_amg *h);
void so_amg_level_dims(const so_amg *h, int lvl, int *nrow, int *nnz, int *p_ncol, int *p_nnz);
void so_amg_level_get(const so_amg *h, int lvl, const int **rp, const int **ci, const double **v,
                      const double **diag, const int **prp, const int **pci, const double **pv);
void so_amg_set_smoother(so_amg *h, double omega, int smooth_iter);
/* AMG_solver::AMG_solve_jacobi(b,x,iterations>0) (src/AMG_phases.cpp:151-192): exactly `cycles` V-cycles */
void so_amg_vcycle(so_amg *h, const double *b, double *x, int cycles);
/* AMG_solve_jacobi(b,x,-1) (src/AMG_phases.cpp:194-226): until ||r|| <= tol (absolute) or max_cycles.
 * hist[0] = initial residual, hist[k] = residual after cycle k.  Returns cycles done. */
int so_amg_solve(so_amg *h, const double *b, double *x, double tol, int max_cycles, double *hist);
/* coarsest-level direct solve alone */
void so_amg_coarse_solve(so_amg *h, const double *b, double *x);
/* Solver_PCG_1 (src/AMG_main_solvers.cpp:107-167).  hist[0] = ||r0||, hist[k] = ||r|| after iteration k. */
int so_pcg(so_amg *h, const double *b, double *x, double tol, int max_iter, double *hist);
/* Solver_PBiCG_1 (src/AMG_main_solvers.cpp:358-458) */
int so_pbicgstab(so_amg *h, const double *b, double *x, double tol, int max_iter, double *hist);
/* Solver_CG_1 (src/AMG_main_solvers.cpp:47-103, assumes x0 = 0) / Solver_BiCG_1 (:271-355) */
int so_cg(int n, const int *rp, const int *ci, const double *v, const double *b, double *x, double tol,
          int max_iter, double *hist);
int so_bicgstab(int n, const int *rp, const int *ci, const double *v, const double *b, double *x, double tol,
                int max_iter, double *hist);

/* Restarted GMRES(m), right-preconditioned by one V-cycle of h (h == NULL: unpreconditioned, then the matrix is
 * (n, rp, ci, v)).  NOT in the reference (advertised by README.md:13, no code) — this states the algorithm
 * the CUDA path implements so that the two can be compared: Arnoldi with classical Gram-Schmidt applied twice (CGS2),
 * Givens rotations, x += M^-1 (V y) at the end of each cycle, true residual recomputed at every restart.
 * hist[0] = ||b - A x0||, hist[k] = Givens estimate of the residual norm after k inner iterations, except that the
 * entry of the last iteration of a cycle is overwritten by the true residual norm.  Returns inner iterations done. */
int so_gmres(so_amg *h, int n, const int *rp, const int *ci, const double *v, const double *b, double *x, double tol,
             int restart, int max_iter, double *hist);

/* AMG_solver_setup_SOR (src/AMG_phases.cpp:94-147): every level colour-permuted (color_matrix_and_reorder), P columns
 * re-labelled by the coarse permutation (reorder_prolongator, src/AMG_cycle_utilities.cpp:149-188). */
so_amg *so_amg_setup_sor(int n, const int *rp, const int *ci, const double *v, int max_levels, int limit_upper,
                         int limit_lower);
/* AMG_Solver_2 (src/AMG_main_solvers.cpp:30-43) = reorder_rhs + AMG_solve_SOR(b,x,-1) (src/AMG_phases.cpp:275-304):
 * V-cycles with 6 multicolour-SOR sweeps (the literal 6 of :281,:295) until ||r|| <= tol.  b and x are in the CALLER's
 * ordering (the reference never copies x back and permutes with the forward permutation twice — SURVEY Appendix B;
 * the oracle returns x properly).  hist as so_amg_solve.  Returns cycles. */
int so_amg_solve_sor(so_amg *h, const double *b, double *x, double tol, int max_cycles, double *hist);

/* ---- synthetic matrices of BASELINE.json's configs; arrays malloc'ed ---- */
void so_gen_poisson2d_5pt(int nx, int ny, int **rp, int **ci, double **v);
void so_gen_poisson3d_7pt(int nx, int ny, int nz, int **rp, int **ci, double **v);

#ifdef __cplusplus
}
#endif
#endif
