// share.cpp — one host hierarchy for all ranks of a node.
//
// The setup phase is sequential host code; on N GPUs every rank used to repeat it and hold the whole hierarchy (25+ GB
// at 512^3, times 8 ranks).  Here rank 0 builds it once and writes the level arrays as raw binary files (ideally under
// /dev/shm); every rank maps the files read-only and cuts its part out (host/dist_plan.cpp).  The page cache holds a
// single copy.  No reference counterpart (the reference is single-process).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "sparsh_amg.hpp"

namespace {

struct Mapping {
    void *ptr = nullptr;
    size_t bytes = 0;
};
struct SharedState {
    std::vector<Mapping> maps;
};

bool write_file(const std::string &path, const void *data, size_t bytes) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const char *p = (const char *)data;
    size_t left = bytes;
    while (left > 0) {  // fwrite in 1 GiB pieces
        const size_t chunk = left < ((size_t)1 << 30) ? left : ((size_t)1 << 30);
        if (std::fwrite(p, 1, chunk, f) != chunk) {
            std::fclose(f);
            return false;
        }
        p += chunk;
        left -= chunk;
    }
    return std::fclose(f) == 0;
}

void *map_file(const std::string &path, size_t bytes, SharedState *st) {
    if (bytes == 0) return nullptr;
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return nullptr;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < bytes) {
        close(fd);
        return nullptr;
    }
    void *p = mmap(nullptr, bytes, PROT_READ, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return nullptr;
    st->maps.push_back(Mapping{p, bytes});
    return p;
}

std::string name(const std::string &dir, const char *what, int level, const char *ext) {
    return dir + "/" + what + std::to_string(level) + "." + ext;
}

}  // namespace

void sparsh_release_shared_hierarchy(AMG_solver *S) {
    SharedState *st = (SharedState *)S->shared_mapping;
    if (!st) return;
    for (int q = 0; q <= S->l; q++) {
        if (S->Av && S->Av[q]) {
            S->Av[q]->diagonal = nullptr;  // mapped, not owned: keep ~sp_matrix_mg from deleting them
            S->Av[q]->helper = nullptr;
            delete S->Av[q];
            S->Av[q] = nullptr;
        }
        if (S->Pv && q < S->l && S->Pv[q]) {
            delete S->Pv[q];
            S->Pv[q] = nullptr;
        }
    }
    for (auto &m : st->maps) munmap(m.ptr, m.bytes);
    delete st;
    S->shared_mapping = nullptr;
}

extern "C" {

// write every level of a built hierarchy under `dir` (which must exist); returns 0 on success
int sparsh_host_amg_save(void *Sv, const char *dir_c) {
    AMG_solver *S = (AMG_solver *)Sv;
    const std::string dir(dir_c);
    std::ofstream meta(dir + "/meta.txt");
    if (!meta) return -1;
    meta << (S->l + 1) << "\n";
    struct Job {
        std::string path;
        const void *ptr;
        size_t bytes;
    };
    std::vector<Job> jobs;
    for (int k = 0; k <= S->l; k++) {
        const sp_matrix_mg *A = S->Av[k];
        const int nnz = A->rowptr[A->nrow];
        const int pncol = k < S->l ? S->Pv[k]->ncol : 0, pnnz = k < S->l ? S->Pv[k]->rowptr[S->Pv[k]->nrow] : 0;
        meta << A->nrow << " " << nnz << " " << pncol << " " << pnnz << "\n";
        jobs.push_back({name(dir, "A", k, "rp"), A->rowptr, sizeof(int) * ((size_t)A->nrow + 1)});
        jobs.push_back({name(dir, "A", k, "ci"), A->colindex, sizeof(int) * (size_t)nnz});
        jobs.push_back({name(dir, "A", k, "v"), A->val, sizeof(double) * (size_t)nnz});
        jobs.push_back({name(dir, "A", k, "d"), A->diagonal, sizeof(double) * (size_t)A->nrow});
        if (k < S->l) {
            const sp_matrix_mg *P = S->Pv[k];
            jobs.push_back({name(dir, "P", k, "rp"), P->rowptr, sizeof(int) * ((size_t)P->nrow + 1)});
            jobs.push_back({name(dir, "P", k, "ci"), P->colindex, sizeof(int) * (size_t)pnnz});
            jobs.push_back({name(dir, "P", k, "v"), P->val, sizeof(double) * (size_t)pnnz});
        }
    }
    // the files are independent: several writers (the biggest arrays are listed first, dynamic schedule)
    int failed = 0;
#pragma omp parallel for num_threads(std::max(1, sparsh::options().threads)) schedule(dynamic, 1) reduction(+ : failed)
    for (int j = 0; j < (int)jobs.size(); j++)
        if (!write_file(jobs[j].path, jobs[j].ptr, jobs[j].bytes)) failed++;
    if (failed) return -2;
    meta.close();
    return meta ? 0 : -3;
}

// map a saved hierarchy read-only; returns an AMG_GPU1_solver* (nullptr on failure).  The object supports everything a
// built one does (upload, partition plans); its arrays must not be modified.
void *sparsh_host_amg_load(const char *dir_c) {
    const std::string dir(dir_c);
    std::ifstream meta(dir + "/meta.txt");
    int nlev = 0;
    if (!(meta >> nlev) || nlev < 1) return nullptr;
    AMG_GPU1_solver *S = new AMG_GPU1_solver();
    S->reserve_levels(nlev);  // sized from the file, the global option is left alone
    SharedState *st = new SharedState();
    S->shared_mapping = st;
    S->l = nlev - 1;
    bool ok = true;
    for (int k = 0; k < nlev && ok; k++) {
        int nrow = 0, nnz = 0, pncol = 0, pnnz = 0;
        if (!(meta >> nrow >> nnz >> pncol >> pnnz)) {
            ok = false;
            break;
        }
        sp_matrix_mg *A = new sp_matrix_mg();
        A->nrow = A->ncol = nrow;
        A->nnz = nnz;
        A->rowptr = (int *)map_file(name(dir, "A", k, "rp"), sizeof(int) * ((size_t)nrow + 1), st);
        A->colindex = (int *)map_file(name(dir, "A", k, "ci"), sizeof(int) * (size_t)nnz, st);
        A->val = (double *)map_file(name(dir, "A", k, "v"), sizeof(double) * (size_t)nnz, st);
        A->diagonal = (double *)map_file(name(dir, "A", k, "d"), sizeof(double) * (size_t)nrow, st);
        S->Av[k] = A;
        ok = A->rowptr && (nnz == 0 || (A->colindex && A->val)) && A->diagonal;
        if (ok && k < nlev - 1) {
            sp_matrix_mg *P = new sp_matrix_mg();
            P->nrow = nrow;
            P->ncol = pncol;
            P->nnz = pnnz;
            P->rowptr = (int *)map_file(name(dir, "P", k, "rp"), sizeof(int) * ((size_t)nrow + 1), st);
            P->colindex = (int *)map_file(name(dir, "P", k, "ci"), sizeof(int) * (size_t)pnnz, st);
            P->val = (double *)map_file(name(dir, "P", k, "v"), sizeof(double) * (size_t)pnnz, st);
            S->Pv[k] = P;
            ok = P->rowptr && (pnnz == 0 || (P->colindex && P->val));
        }
    }
    if (!ok) {
        delete S;
        return nullptr;
    }
    return S;
}

}  // extern "C"
