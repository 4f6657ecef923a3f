// additions_main.cpp — the pieces the reference's README advertises but its sources lack, used
// through the same header: smoothed-aggregation coarsening, GMRES(m), MatrixMarket / binary CSR input.
//
//   g++ -std=c++17 -I sparsh_amg_b200/host examples/additions_main.cpp -L sparsh_amg_b200/lib -lsparsh_amg -lsparsh_b200
//   ./a.out matrix.mtx            (b = A * ones)
#include <cstdio>
#include <string>
#include <vector>

#include "AMG.hpp"

int main(int argc, char *argv[]) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <matrix.mtx | matrix.csr>\n", argv[0]);
        return 2;
    }
    const std::string path(argv[1]);
    sp_matrix_mg *A = path.size() > 4 && path.substr(path.size() - 4) == ".csr" ? read_binary_csr(argv[1])
                                                                               : read_matrix_market(argv[1]);
    if (!A) {
        std::fprintf(stderr, "cannot read %s\n", argv[1]);
        return 1;
    }
    A->sp_matrix_fill();
    A->sp_matrix_fill_diagonal();
    const int n = A->nrow;
    std::vector<double> ones((size_t)n, 1.0), b((size_t)n, 0.0), x((size_t)n, 0.0);
    for (int i = 0; i < n; i++)
        for (int j = A->rowptr[i]; j < A->rowptr[i + 1]; j++) b[i] += A->val[j] * ones[A->colindex[j]];

    sparsh::Options &o = sparsh::options();
    o.print_setup = o.print_solve = 0;
    o.tol_mode = sparsh::TOL_RELATIVE;
    o.max_levels = 32;
    o.coarsening = sparsh::COARSEN_SA;  // smoothed aggregation instead of the shipped pairwise HEM
    o.gmres_restart = 30;

    double *bp = b.data(), *xp = x.data();
    Solver_PGMRES_1(*A, bp, xp);  // V-cycle-preconditioned GMRES(30)
    const sparsh::Report &r = sparsh::last_report();
    std::printf("REPORT Solver_PGMRES_1 iterations=%d converged=%d setup=%.3fs solve=%.3fs\n", r.iterations, r.converged,
                r.setup_seconds, r.solve_seconds);
    write_binary_csr((path + ".csr").c_str(), *A);  // next time: read_binary_csr
    A->~sp_matrix_mg();
    return r.converged ? 0 : 3;
}
