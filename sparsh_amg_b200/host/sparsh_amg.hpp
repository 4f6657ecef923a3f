// sparsh_amg.hpp — host-side mirror of the SParSH-AMG C++ interface for the B200 solve phase.
//
// A user of the reference includes "AMG.hpp" and calls free functions `void f(sp_matrix_mg& A, double*& b, double*& x)`
// (reference include/AMG.hpp:40-85).  This header keeps those names, argument meanings and conventions (0-based int32
// CSR, caller-owned A/b/x, x in/out, void return, explicit destructor calls tolerated) so callers link unchanged, and
// implements them on top of the C-ABI in include/sparsh_b200.h.  Nothing here needs MKL.
//
//   reference type / function                         here
//   sp_matrix            (include/AMG_matrix.hpp:6-32)        same public members
//   sp_matrix_mg         (include/AMG_cpu_matrix.hpp:12-51)   same public members minus the MKL handle (opaque slot kept)
//   AMG_solver           (include/AMG_phases.hpp:8-53)        same members; setup is native C++ (HEM/Beck + Galerkin),
//                                                     AMG_solve_jacobi runs the device V-cycle
//   sp_matrix_gpu        (include/AMG_gpu_matrix.hpp:10-48)   wraps sparsh_matrix_t
//   AMG_GPU1_solver "MI" (include/AMG_gpu_phases_2.hpp:9-40)  wraps sparsh_hierarchy_t (hierarchy resident in HBM)
//   AMG_GPU_solver  "CI" (include/AMG_gpu_phases.hpp:10-55)   same engine: with 180 GB of HBM nothing is streamed
//   16 solver entry points + readcoo/read_coo_new_format      same signatures; AMG_Solver_1 added (README alias)
//
// The reference's tunables are compile-time macros with collision-prone names (th, omega, tol1, level1 ...,
// include/AMG.hpp:15-27).  They remain available as defaults behind #ifndef guards when SPARSH_LEGACY_MACROS is
// defined; the library itself reads the run-time `sparsh::Options` below instead.
#ifndef SPARSH_AMG_HPP_
#define SPARSH_AMG_HPP_

#include <vector>

#ifdef SPARSH_LEGACY_MACROS
#ifndef th
#define th 2
#endif
#ifndef omega
#define omega 0.66667
#endif
#ifndef nsmooth
#define nsmooth 6
#endif
#ifndef tol1
#define tol1 1e-8
#endif
#ifndef limit_upper
#define limit_upper 4000
#endif
#ifndef limit_lower
#define limit_lower 2000
#endif
#ifndef level1
#define level1 6
#endif
#ifndef smooth_iter
#define smooth_iter 6
#endif
#ifndef print_setup_phase_details
#define print_setup_phase_details 1
#endif
#ifndef print_solve_phase_details
#define print_solve_phase_details 1
#endif
#ifndef thgpu
#define thgpu 1024
#endif
#endif  // SPARSH_LEGACY_MACROS

struct sparsh_matrix_s;
struct sparsh_hierarchy_s;

namespace sparsh {

enum Coarsening { COARSEN_HEM = 0, COARSEN_BECK = 1, COARSEN_SA = 2 };  // SA: smoothed aggregation (addition, F2)
enum ToleranceMode { TOL_ABSOLUTE = 0, TOL_RELATIVE = 1 };

// Run-time replacements of the reference's macros (same defaults), plus the guards the reference lacks.
struct Options {
    int threads = 2;               // th            host threads used by the setup phase
    double relax = 0.66667;        // omega
    double tol = 1e-8;             // tol1
    int tol_mode = TOL_ABSOLUTE;   // the reference stops on ||r|| <= tol1 (absolute); relative = tol*||b||
    int coarse_upper = 4000;       // limit_upper
    int coarse_lower = 2000;       // limit_lower
    int max_levels = 6;            // level1 (raise for large problems: SURVEY F10)
    int sweeps = 7;                // Jacobi sweeps per smoothing step: smooth_iter+1 = 7 is what the CPU reference runs,
                                   // 6 what its GPU path runs
    int print_setup = 1;           // print_setup_phase_details
    int print_solve = 1;           // print_solve_phase_details
    int coarsening = COARSEN_HEM;  // the shipped default (reference src/AMG_phases.cpp:61; Beck is :63)
    int max_iter = 10000;          // iteration cap (the reference's AMG and BiCGStab loops have none)
    int use_graph = 1;             // CUDA-graph the V-cycle / Krylov iteration
    int halo_mode = 1;             // multi-GPU halo exchange: 1 NVLink peer-memory pushes, 0 ncclSend/ncclRecv
    int gpu_rap = 0;               // 1: Galerkin products of the setup run on the device (sparsh_galerkin_rap; same result)
    int device = -1;               // multi-GPU entry points: CUDA device of this rank (-1: the rank itself)
    int tail_threshold = 1100000;  // multi-GPU: levels with at most this many rows are replicated on every GPU
    int gmres_restart = 30;        // Solver_GMRES_1 / Solver_PGMRES_1 (used by sparsh_gmres; the host-buffer path uses 30)
    double sa_theta = 0.08;        // COARSEN_SA: strength threshold (halved per level)
    double sa_relax = 4.0 / 3.0;   // COARSEN_SA: prolongator smoothing factor, omega = sa_relax / rho(D^-1 A)
};
Options &options();

// What the last solver call did (the reference only prints; tests and benches read this instead)
struct Report {
    int iterations = 0;
    int converged = 0;
    std::vector<double> history;  // history[0] = initial residual norm
    double setup_seconds = 0.0;   // host hierarchy construction
    double upload_seconds = 0.0;  // H2D of the hierarchy + device coarse factorisation
    double solve_seconds = 0.0;
};
Report &last_report();

}  // namespace sparsh

// ---------------------------------------------------------------------------------------------------------
// matrices
// ---------------------------------------------------------------------------------------------------------
class sp_matrix {
   public:
    int nrow = 0;
    int ncol = 0;
    int nnz = 0;
    int *rowptr = nullptr;
    int *colindex = nullptr;
    double *val = nullptr;

    sp_matrix(int r, int c, int n);  // allocates and zero-fills the three arrays
    sp_matrix();
    void check_sp_matrix();          // prints the matrix
};

class sp_matrix_mg : public sp_matrix {
   public:
    void *A1 = nullptr;  // the reference keeps an MKL handle here; unused (kept so member order is familiar)
    int sA = 0;
    double *diagonal = nullptr;
    double *helper = nullptr;
    double *entries = nullptr;
    int *color = nullptr;        // after color_matrix_and_reorder(): perm[new] = old
    int *color_count = nullptr;  // prefix offsets per colour
    int total_colors = 0;
    int max_color_row = 0;

    using sp_matrix::sp_matrix;
    void sp_matrix_fill();             // sorts the columns of every row in place (what mkl_sparse_order did)
    void sp_matrix_fill_diagonal();    // diagonal[] and helper[]
    void color_matrix_and_reorder();   // greedy first-fit colouring + symmetric permutation, in place
    void scale_system(double *&b);
    void normalize_matrix();
    ~sp_matrix_mg();
};

// ---------------------------------------------------------------------------------------------------------
// setup-phase building blocks (native re-implementations; integer outputs are bit-identical to the reference's)
// ---------------------------------------------------------------------------------------------------------
namespace sequential {
void HEM_Prolongator(sp_matrix_mg &A, sp_matrix_mg *&P, int l1);
void beck_prolongator(sp_matrix_mg &A, sp_matrix_mg *&P1);
void SA_Prolongator(sp_matrix_mg &A, sp_matrix_mg *&P, int level);  // addition: smoothed aggregation
}  // namespace sequential
namespace parallel {
void coarsen_matrix(sp_matrix_mg &A, sp_matrix_mg *&Ac, sp_matrix_mg &P1);  // Ac = P^T (A P), columns sorted
void reorder_prolongator(sp_matrix_mg &A, sp_matrix_mg *&P);
void reorder_rhs(sp_matrix_mg &A, double *&b);
}  // namespace parallel

// ---------------------------------------------------------------------------------------------------------
// device matrix
// ---------------------------------------------------------------------------------------------------------
class sp_matrix_gpu {
   public:
    int nrow = 0, ncol = 0, nnz = 0;
    sparsh_matrix_s *handle = nullptr;

    sp_matrix_gpu(sp_matrix_mg &A);                         // records the shape (reference: allocates)
    void matrix_transfer_gpu(sp_matrix_mg &A, void *stream = nullptr);  // uploads (once)
    // `steps` fused Jacobi sweeps on device vectors; hgpu is scratch (the reference's residual buffer)
    void smooth_jacobi(double *bgpu, double *xgpu, double *hgpu, void *stream, int steps);
    ~sp_matrix_gpu();
};

// ---------------------------------------------------------------------------------------------------------
// solver objects
// ---------------------------------------------------------------------------------------------------------
class AMG_solver {
   public:
    int l = 0;  // index of the coarsest level
    sp_matrix_mg **Av = nullptr;
    sp_matrix_mg **Pv = nullptr;
    double **Xv = nullptr;
    double **Bv = nullptr;
    double **Rv = nullptr;
    sparsh_hierarchy_s *device = nullptr;  // resident device hierarchy (created on first solve / GPU_Allocations)
    bool torn_down = false;
    void *shared_mapping = nullptr;  // set when the hierarchy's arrays live in files mapped read-only (host/share.cpp)
    int capacity = 0;                // slots of Av/Pv/Xv/Bv/Rv (the reference sizes them with the macro level1)

    AMG_solver();
    void reserve_levels(int nlevels);  // grow the per-level arrays (contents kept); setup and load call it
    void AMG_solver_setup_jacobi(sp_matrix_mg &A);
    void AMG_solver_setup_SOR(sp_matrix_mg &A);
    // host b, x (x in/out).  iterations > 0: exactly that many V-cycles; -1: until ||r|| <= tolerance.
    void AMG_solve_jacobi(double *&b, double *&x, int iterations);
    void AMG_solve_SOR(double *&b, double *&x, int iterations);
    void upload();  // idempotent: hierarchy -> HBM, R = P^T, device coarse factorisation
    ~AMG_solver();
};

class AMG_GPU1_solver : public AMG_solver {
   public:
    using AMG_solver::AMG_solver;
    void GPU_Allocations();
    void helper(double *b, double *x, int iterations);     // b, x on the host
    void AMG_Solve(double *b, double *x, int iterations);  // b, x on the device
    ~AMG_GPU1_solver();
};

class AMG_GPU_solver : public AMG_solver {
   public:
    using AMG_solver::AMG_solver;
    void GPU_Allocations();
    void AMG_GPU_solve(double *b, double *x, int iterations);    // b, x on the host
    void AMG_GPU_solve_1(double *b, double *x, int iterations);  // b, x on the device
    ~AMG_GPU_solver();
};

// ---------------------------------------------------------------------------------------------------------
// entry points (reference include/AMG.hpp:32-85)
// ---------------------------------------------------------------------------------------------------------
void readcoo(char *matrixfile, char *rhsfile, sp_matrix_mg *&A, double *&b);
void read_coo_new_format(char *matrixfile, sp_matrix_mg *&A, double *&b);
// addition: standard MatrixMarket coordinate files (1-based, unsorted, general | symmetric | pattern); nullptr on error
sp_matrix_mg *read_matrix_market(const char *path);
// addition: binary CSR (magic "SPRSHCSR", version, nrow, ncol, nnz, then the three arrays); 0 on success / nullptr on error
int write_binary_csr(const char *path, const sp_matrix_mg &A);
sp_matrix_mg *read_binary_csr(const char *path);

void AMG_Solver_CPU_baseline(sp_matrix_mg &A, double *&b, double *&x);
void AMG_Solver_1(sp_matrix_mg &A, double *&b, double *&x);  // README name of the above
void AMG_Solver_2(sp_matrix_mg &A, double *&b, double *&x);
void AMG_Solver_CPU_GPU_CI(sp_matrix_mg &A, double *&b, double *&x);
void AMG_Solver_CPU_GPU_MI(sp_matrix_mg &A, double *&b, double *&x);
void Solver_CG_1(sp_matrix_mg &A, double *&b, double *&x);
void Solver_CG_2(sp_matrix_mg &A, double *&b, double *&x);
void Solver_PCG_1(sp_matrix_mg &A, double *&b, double *&x);
void Solver_PCG_2(sp_matrix_mg &A, double *&b, double *&x);
void Solver_PCG_3(sp_matrix_mg &A, double *&b, double *&x);
void Solver_PCG_4(sp_matrix_mg &A, double *&b, double *&x);
void Solver_BiCG_1(sp_matrix_mg &A, double *&b, double *&x);
void Solver_PBiCG_1(sp_matrix_mg &A, double *&b, double *&x);
void Solver_PBiCG_2(sp_matrix_mg &A, double *&b, double *&x);
void Solver_PBiCG_3(sp_matrix_mg &A, double *&b, double *&x);
void Solver_PBiCG_4(sp_matrix_mg &A, double *&b, double *&x);
void coarsening_2(sp_matrix_mg &A, double *&b, double *&x);
// additions: restarted GMRES(options().gmres_restart), plain and V-cycle-preconditioned
void Solver_GMRES_1(sp_matrix_mg &A, double *&b, double *&x);
void Solver_PGMRES_1(sp_matrix_mg &A, double *&b, double *&x);
// additions: the same solvers row-sharded over N GPUs of one node, one process per GPU (host/dist_plan.cpp).  Every rank
// passes the same global A, b, x, its rank, the world size and the 128-byte id rank 0 got from sparsh_dist_get_unique_id;
// x returns the global solution on every rank.  options().device (-1: = rank) selects the GPU,
// options().tail_threshold the row count below which levels are replicated instead of partitioned.
void AMG_Solver_MG(sp_matrix_mg &A, double *&b, double *&x, int nranks, int rank, const char *id128);
void Solver_PCG_MG(sp_matrix_mg &A, double *&b, double *&x, int nranks, int rank, const char *id128);
void Solver_PBiCG_MG(sp_matrix_mg &A, double *&b, double *&x, int nranks, int rank, const char *id128);

#endif  // SPARSH_AMG_HPP_
