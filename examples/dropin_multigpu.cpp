// dropin_multigpu.cpp — the reference's calling convention on N GPUs: one process per GPU, every process reads the same
// system and calls the *_MG twins of the reference's solver entry points (sparsh_amg.hpp) with its rank, the number
// of ranks and the 128-byte id of rank 0 (exchanged here through a file; an MPI application would MPI_Bcast it).
//
//   ./dropin_multigpu <matrix file> <rhs file> <nranks> <rank> <id file>        (start one per rank)
//
// REPORT lines are parsed by tests/test_cpp_dropin.py; the process leaves through sparsh_dist_finalize() and a normal
// return from main (no _exit): the communicator teardown is part of what is tested.
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "AMG.hpp"
#include "../include/sparsh_b200.h"

typedef void (*mg_fn)(sp_matrix_mg &, double *&, double *&, int, int, const char *);

int main(int argc, char *argv[]) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s <matrix file> <rhs file> <nranks> <rank> <id file>\n", argv[0]);
        return 2;
    }
    const int nranks = std::atoi(argv[3]), rank = std::atoi(argv[4]);
    sparsh::options().print_solve = 0;
    sparsh::options().print_setup = 0;
    sparsh::options().tail_threshold = 8000;  // the bundled system is small: distribute its finest level only
    if (sparsh_init(rank) != SPARSH_OK) {
        std::fprintf(stderr, "no GPU for rank %d: %s\n", rank, sparsh_last_error());
        return 3;
    }
    char id[SPARSH_NCCL_ID_BYTES];
    if (rank == 0) {
        if (sparsh_dist_get_unique_id(id) != SPARSH_OK) return 4;
        const std::string tmp = std::string(argv[5]) + ".tmp";
        FILE *f = std::fopen(tmp.c_str(), "wb");
        std::fwrite(id, 1, sizeof id, f);
        std::fclose(f);
        std::rename(tmp.c_str(), argv[5]);
    } else {
        FILE *f = nullptr;
        for (int tries = 0; tries < 600 && !(f = std::fopen(argv[5], "rb")); tries++) usleep(100000);
        if (!f || std::fread(id, 1, sizeof id, f) != sizeof id) return 5;
        std::fclose(f);
    }
    sp_matrix_mg *A = nullptr;
    double *rhs = nullptr;
    readcoo(argv[1], argv[2], A, rhs);
    A->sp_matrix_fill();
    A->sp_matrix_fill_diagonal();
    const struct {
        const char *name;
        mg_fn run;
    } table[] = {{"Solver_PCG_MG", Solver_PCG_MG}, {"AMG_Solver_MG", AMG_Solver_MG}, {"Solver_PBiCG_MG", Solver_PBiCG_MG}};
    std::vector<double> guess((size_t)A->nrow);
    for (const auto &e : table) {
        guess.assign(guess.size(), 0.0);
        double *x = guess.data();
        e.run(*A, rhs, x, nranks, rank, id);
        double r2 = 0.0;  // true residual of the GLOBAL solution every rank received
        for (int i = 0; i < A->nrow; i++) {
            double s = rhs[i];
            for (int j = A->rowptr[i]; j < A->rowptr[i + 1]; j++) s -= A->val[j] * x[A->colindex[j]];
            r2 += s * s;
        }
        const sparsh::Report &rep = sparsh::last_report();
        std::printf("REPORT rank=%d %s iterations=%d converged=%d residual=%.6e\n", rank, e.name, rep.iterations, rep.converged,
                    std::sqrt(r2));
    }
    if (sparsh_dist_finalize() != SPARSH_OK) return 6;
    std::printf("FINALIZED rank=%d\n", rank);
    delete[] rhs;
    return 0;
}
