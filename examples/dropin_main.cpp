// dropin_main.cpp — a caller that uses nothing but the reference's public API (the names declared in its AMG.hpp),
// built against this repository's host/AMG.hpp and libraries instead of the reference's.
//
//   g++ -std=c++17 -I sparsh_amg_b200/host examples/dropin_main.cpp -L sparsh_amg_b200/lib -lsparsh_amg -lsparsh_b200
//   ./a.out matrix_poisson_P1_14401 matrix_poisson_P1rhs_14401
//
// It reads the two-file COO fixture, prepares the matrix the way the reference's example does (fill + diagonal), then
// runs a table of solver entry points from the same zero initial guess.  The REPORT lines (sparsh::last_report(), an
// addition of this library) are what tests/test_cpp_dropin.py parses.
#include <cstdio>
#include <vector>

#include "AMG.hpp"

typedef void (*solver_fn)(sp_matrix_mg &, double *&, double *&);

struct Entry {
    const char *name;
    solver_fn run;
};

int main(int argc, char *argv[]) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <matrix file> <rhs file>\n", argv[0]);
        return 2;
    }
    sparsh::options().print_solve = 0;
    sparsh::options().print_setup = 0;

    sp_matrix_mg *system_matrix = nullptr;
    double *rhs = nullptr;
    readcoo(argv[1], argv[2], system_matrix, rhs);
    system_matrix->sp_matrix_fill();
    system_matrix->sp_matrix_fill_diagonal();
    std::printf("Matrix Size\t%d\n", system_matrix->nrow);

    const Entry table[] = {{"AMG_Solver_CPU_GPU_CI", AMG_Solver_CPU_GPU_CI},
                           {"AMG_Solver_CPU_GPU_MI", AMG_Solver_CPU_GPU_MI},
                           {"AMG_Solver_CPU_baseline", AMG_Solver_CPU_baseline},
                           {"Solver_PCG_4", Solver_PCG_4},
                           {"Solver_PBiCG_4", Solver_PBiCG_4}};
    std::vector<double> guess((size_t)system_matrix->nrow);
    for (const Entry &e : table) {
        guess.assign(guess.size(), 0.0);
        double *x = guess.data();
        e.run(*system_matrix, rhs, x);
        const sparsh::Report &rep = sparsh::last_report();
        std::printf("REPORT %s iterations=%d converged=%d\n", e.name, rep.iterations, rep.converged);
    }
    system_matrix->~sp_matrix_mg();  // explicit destructor without delete: the reference's teardown idiom is tolerated
    delete[] rhs;
    return 0;
}
