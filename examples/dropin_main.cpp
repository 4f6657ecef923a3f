// dropin_main.cpp — a caller written against the reference's public API only (AMG.hpp names), linked against this
// repository's libraries instead of the reference's.  Flow of the reference's main.cpp:13-43: read the two-file COO
// fixture, sp_matrix_fill, sp_matrix_fill_diagonal, call the solvers.  The only non-reference lines are the
// sparsh::last_report() prints the tests parse.
//
//   g++ -std=c++17 -I sparsh_amg_b200/host examples/dropin_main.cpp -L sparsh_amg_b200/lib -lsparsh_amg -lsparsh_b200
#include <algorithm>
#include <cstdio>
#include <iostream>

#include "AMG.hpp"

int main(int argc, char *argv[]) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s matrixfile rhsfile\n", argv[0]);
        return 2;
    }
    sparsh::options().print_solve = 0;
    sparsh::options().print_setup = 0;

    sp_matrix_mg *A = new sp_matrix_mg();
    double *b;
    readcoo(argv[1], argv[2], A, b);
    double *x = new double[A->nrow]();
    std::fill(x, x + A->nrow, 0);
    A->sp_matrix_fill();
    A->sp_matrix_fill_diagonal();
    std::cout << "Matrix Size\t" << A->nrow << std::endl;

    AMG_Solver_CPU_GPU_CI(*A, b, x);
    std::printf("REPORT AMG_Solver_CPU_GPU_CI iterations=%d converged=%d\n", sparsh::last_report().iterations,
                sparsh::last_report().converged);
    std::fill(x, x + A->nrow, 0);
    AMG_Solver_CPU_GPU_MI(*A, b, x);
    std::printf("REPORT AMG_Solver_CPU_GPU_MI iterations=%d converged=%d\n", sparsh::last_report().iterations,
                sparsh::last_report().converged);
    std::fill(x, x + A->nrow, 0);
    AMG_Solver_CPU_baseline(*A, b, x);
    std::printf("REPORT AMG_Solver_CPU_baseline iterations=%d converged=%d\n", sparsh::last_report().iterations,
                sparsh::last_report().converged);
    std::fill(x, x + A->nrow, 0);
    Solver_PCG_4(*A, b, x);
    std::printf("REPORT Solver_PCG_4 iterations=%d converged=%d\n", sparsh::last_report().iterations,
                sparsh::last_report().converged);
    std::fill(x, x + A->nrow, 0);
    Solver_PBiCG_4(*A, b, x);
    std::printf("REPORT Solver_PBiCG_4 iterations=%d converged=%d\n", sparsh::last_report().iterations,
                sparsh::last_report().converged);

    A->~sp_matrix_mg();  // the reference's teardown idiom (main.cpp:40): explicit destructor, no delete
    delete[] x;
    delete[] b;
    return 0;
}
