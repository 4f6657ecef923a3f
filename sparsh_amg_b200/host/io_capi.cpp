// io_capi.cpp — file readers, synthetic matrix generators and the extern "C" driver Python uses to reach the C++ API.
//
//   readcoo / read_coo_new_format     reference src/AMG_file_read.cpp:39-185 (formats kept: 0-based, row-sorted COO,
//                                     header "nrow ncol nnz"; the rhs file starts with its length)
//   generators                        SURVEY.md §8d synthetic configs (natural ordering, x fastest, Dirichlet by truncation)
//   sparsh_host_*                     plain-C handles over sp_matrix_mg / AMG_GPU1_solver for ctypes (tests, bench.py)
#include <omp.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/sparsh_b200.h"
#include "sparsh_amg.hpp"

using sparsh::options;

// ---------------------------------------------------------------------------------------------------------
// readers
// ---------------------------------------------------------------------------------------------------------
static void read_triplets(std::istream &in, sp_matrix_mg *&A, int nrow, int ncol, int nnz) {
    A = new sp_matrix_mg(nrow, ncol, nnz);
    for (int i = 0; i < nnz; i++) {
        int r = 0;
        in >> r >> A->colindex[i] >> A->val[i];
        A->rowptr[r + 1]++;  // rows are assumed sorted, as in the reference
    }
    for (int i = 0; i < nrow; i++) A->rowptr[i + 1] += A->rowptr[i];
}

void readcoo(char *matrixfile, char *rhsfile, sp_matrix_mg *&A, double *&b) {
    std::ifstream in(matrixfile);
    int nrow = 0, ncol = 0, nnz = 0;
    in >> nrow >> ncol >> nnz;
    read_triplets(in, A, nrow, ncol, nnz);
    in.close();
    b = new double[(size_t)nrow]();
    std::ifstream rhs(rhsfile);
    int count = 0;
    rhs >> count;
    for (int i = 0; i < nrow; i++) rhs >> b[i];
}

void read_coo_new_format(char *matrixfile, sp_matrix_mg *&A, double *&b) {
    std::ifstream in(matrixfile);
    std::string line;
    std::getline(in, line);  // banner: "%%MatrixMarket matrix coordinate real general|symmetric" — parsed and, like the
                             // reference (src/AMG_file_read.cpp:80-136), not acted upon; indices stay 0-based
    std::streampos pos = in.tellg();
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '%') {
            pos = in.tellg();
            continue;
        }
        break;
    }
    in.clear();
    in.seekg(pos);
    int nrow = 0, ncol = 0, nnz = 0;
    in >> nrow >> ncol >> nnz;
    read_triplets(in, A, nrow, ncol, nnz);
    b = new double[(size_t)nrow]();
    for (int i = 0; i < nrow; i++) in >> b[i];
}

// Binary CSR: the text readers spend minutes on 10^9 entries; this format is the three arrays as they
// lie in memory.  Layout (little endian): char magic[8] = "SPRSHCSR", int32 version = 1, int32 nrow, int32 ncol,
// int64 nnz, int32 rowptr[nrow+1], int32 colindex[nnz], float64 val[nnz].  The reference has no counterpart.
namespace {
const char kCsrMagic[8] = {'S', 'P', 'R', 'S', 'H', 'C', 'S', 'R'};
}
int write_binary_csr(const char *path, const sp_matrix_mg &A) {
    FILE *f = std::fopen(path, "wb");
    if (!f) return 1;
    const int32_t version = 1, nrow = A.nrow, ncol = A.ncol;
    const int64_t nnz = A.rowptr[A.nrow];
    bool ok = std::fwrite(kCsrMagic, 1, 8, f) == 8 && std::fwrite(&version, 4, 1, f) == 1 &&
              std::fwrite(&nrow, 4, 1, f) == 1 && std::fwrite(&ncol, 4, 1, f) == 1 && std::fwrite(&nnz, 8, 1, f) == 1 &&
              std::fwrite(A.rowptr, 4, (size_t)nrow + 1, f) == (size_t)nrow + 1 &&
              std::fwrite(A.colindex, 4, (size_t)nnz, f) == (size_t)nnz &&
              std::fwrite(A.val, 8, (size_t)nnz, f) == (size_t)nnz;
    ok = (std::fclose(f) == 0) && ok;
    return ok ? 0 : 2;
}
// nullptr on a missing / truncated / inconsistent file
sp_matrix_mg *read_binary_csr(const char *path) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return nullptr;
    char magic[8];
    int32_t version = 0, nrow = -1, ncol = -1;
    int64_t nnz = -1;
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, kCsrMagic, 8) == 0 &&
              std::fread(&version, 4, 1, f) == 1 && version == 1 && std::fread(&nrow, 4, 1, f) == 1 &&
              std::fread(&ncol, 4, 1, f) == 1 && std::fread(&nnz, 8, 1, f) == 1 && nrow >= 0 && ncol >= 0 && nnz >= 0 &&
              nnz <= INT32_MAX;
    sp_matrix_mg *A = nullptr;
    if (ok) {
        A = new sp_matrix_mg(nrow, ncol, (int)nnz);
        ok = std::fread(A->rowptr, 4, (size_t)nrow + 1, f) == (size_t)nrow + 1 &&
             std::fread(A->colindex, 4, (size_t)nnz, f) == (size_t)nnz &&
             std::fread(A->val, 8, (size_t)nnz, f) == (size_t)nnz && A->rowptr[0] == 0 && A->rowptr[nrow] == nnz;
        for (int i = 0; ok && i < nrow; i++) ok = A->rowptr[i] <= A->rowptr[i + 1];
        for (int64_t j = 0; ok && j < nnz; j++) ok = A->colindex[j] >= 0 && A->colindex[j] < ncol;
        if (!ok) {
            delete[] A->rowptr;
            delete[] A->colindex;
            delete[] A->val;
            A->rowptr = A->colindex = nullptr;
            A->val = nullptr;
            delete A;
            A = nullptr;
        }
    }
    std::fclose(f);
    return A;
}

// Real MatrixMarket coordinate files: 1-based indices, entries in any order, `symmetric` files store one
// triangle.  The reference's read_coo_new_format parses the banner but ignores `symmetric` and assumes 0-based sorted
// input (src/AMG_file_read.cpp:74-185); this reader does what the format says.  Duplicates are summed, columns sorted.
sp_matrix_mg *read_matrix_market(const char *path) {
    std::ifstream in(path);
    if (!in) return nullptr;
    std::string line;
    std::getline(in, line);
    std::string lower(line);
    std::transform(lower.begin(), lower.end(), lower.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (lower.find("%%matrixmarket") != 0 || lower.find("coordinate") == std::string::npos) return nullptr;
    const bool symmetric = lower.find("symmetric") != std::string::npos;
    const bool pattern = lower.find("pattern") != std::string::npos;
    while (std::getline(in, line))
        if (!line.empty() && line[0] != '%') break;
    long nrow = 0, ncol = 0, nent = 0;
    std::istringstream(line) >> nrow >> ncol >> nent;
    std::vector<int> ri, ci;
    std::vector<double> vv;
    ri.reserve((size_t)nent * (symmetric ? 2 : 1));
    ci.reserve(ri.capacity());
    vv.reserve(ri.capacity());
    for (long k = 0; k < nent; k++) {
        long r = 0, c = 0;
        double v = 1.0;
        in >> r >> c;
        if (!pattern) in >> v;
        if (!in || r < 1 || c < 1 || r > nrow || c > ncol) return nullptr;
        ri.push_back((int)r - 1);
        ci.push_back((int)c - 1);
        vv.push_back(v);
        if (symmetric && r != c) {
            ri.push_back((int)c - 1);
            ci.push_back((int)r - 1);
            vv.push_back(v);
        }
    }
    // COO -> CSR (stable counting sort by row), then sort each row by column and merge duplicates
    const size_t m = ri.size();
    std::vector<int> rp((size_t)nrow + 1, 0);
    for (size_t k = 0; k < m; k++) rp[ri[k] + 1]++;
    for (long i = 0; i < nrow; i++) rp[i + 1] += rp[i];
    std::vector<int> cur(rp.begin(), rp.end() - 1), cc(m);
    std::vector<double> cv(m);
    for (size_t k = 0; k < m; k++) {
        const int d = cur[ri[k]]++;
        cc[d] = ci[k];
        cv[d] = vv[k];
    }
    std::vector<int> orp((size_t)nrow + 1, 0), occ;
    std::vector<double> ocv;
    occ.reserve(m);
    ocv.reserve(m);
    std::vector<int> idx;
    for (long i = 0; i < nrow; i++) {
        idx.resize((size_t)(rp[i + 1] - rp[i]));
        for (size_t t = 0; t < idx.size(); t++) idx[t] = rp[i] + (int)t;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return cc[a] < cc[b]; });
        for (size_t t = 0; t < idx.size(); t++) {
            if (t > 0 && cc[idx[t]] == occ.back())
                ocv.back() += cv[idx[t]];
            else {
                occ.push_back(cc[idx[t]]);
                ocv.push_back(cv[idx[t]]);
            }
        }
        orp[i + 1] = (int)occ.size();
    }
    sp_matrix_mg *A = new sp_matrix_mg((int)nrow, (int)ncol, (int)occ.size());
    std::copy(orp.begin(), orp.end(), A->rowptr);
    std::copy(occ.begin(), occ.end(), A->colindex);
    std::copy(ocv.begin(), ocv.end(), A->val);
    return A;
}

// ---------------------------------------------------------------------------------------------------------
// generators
// ---------------------------------------------------------------------------------------------------------
static sp_matrix_mg *gen_7pt(int nx, int ny, int nz, double diag) {
    const long n = (long)nx * ny * nz;
    std::vector<int> rp((size_t)n + 1, 0);
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (long i = 0; i < n; i++) {
        const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((long)nx * ny));
        rp[i + 1] = 1 + (x > 0) + (x < nx - 1) + (y > 0) + (y < ny - 1) + (z > 0) + (z < nz - 1);
    }
    for (long i = 0; i < n; i++) rp[i + 1] += rp[i];
    // arrays allocated without the constructor's zero fill: the generating threads are the first to touch the pages
    sp_matrix_mg *A = new sp_matrix_mg();
    A->nrow = A->ncol = (int)n;
    A->nnz = rp[n];
    A->rowptr = new int[(size_t)n + 1];
    A->colindex = new int[(size_t)std::max(rp[n], 1)];
    A->val = new double[(size_t)std::max(rp[n], 1)];
    std::copy(rp.begin(), rp.end(), A->rowptr);
    const long plane = (long)nx * ny;
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (long i = 0; i < n; i++) {
        const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / plane);
        int o = rp[i];
        auto put = [&](long c, double v) {
            A->colindex[o] = (int)c;
            A->val[o++] = v;
        };
        if (z > 0) put(i - plane, -1.0);
        if (y > 0) put(i - nx, -1.0);
        if (x > 0) put(i - 1, -1.0);
        put(i, diag);
        if (x < nx - 1) put(i + 1, -1.0);
        if (y < ny - 1) put(i + nx, -1.0);
        if (z < nz - 1) put(i + plane, -1.0);
    }
    return A;
}

// 3D 27-point variable-coefficient anisotropic diffusion, trilinear FE on a uniform grid of (nx+1)(ny+1)(nz+1) cells
// with homogeneous Dirichlet boundary (interior nodes only).  Coefficients as proposed in SURVEY.md §8d: per-cell
// kappa log-uniform in [1,1e3] (seeded), axis anisotropy (1, 1e-2, 1e-3).  Nothing in the reference defines this
// config; it exists to exercise 27 nnz/row in the solve phase.
static double cell_kappa(long cx, long cy, long cz, unsigned seed) {
    unsigned long long h = (unsigned long long)seed * 0x9E3779B97F4A7C15ull;
    h ^= (unsigned long long)(cx + 1) * 0xBF58476D1CE4E5B9ull;
    h = (h ^ (h >> 31)) * 0x94D049BB133111EBull;
    h ^= (unsigned long long)(cy + 1) * 0xD6E8FEB86659FD93ull;
    h = (h ^ (h >> 29)) * 0xBF58476D1CE4E5B9ull;
    h ^= (unsigned long long)(cz + 1) * 0x9E3779B97F4A7C15ull;
    h = (h ^ (h >> 32)) * 0x94D049BB133111EBull;
    h ^= h >> 31;
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    return std::pow(10.0, 3.0 * u);
}

static sp_matrix_mg *gen_27pt_aniso(int nx, int ny, int nz, unsigned seed) {
    // 1-D element matrices on [0,1]: mass M = [[1/3,1/6],[1/6,1/3]], stiffness K = [[1,-1],[-1,1]]
    const double M1[2][2] = {{1.0 / 3, 1.0 / 6}, {1.0 / 6, 1.0 / 3}}, K1[2][2] = {{1, -1}, {-1, 1}};
    const double eps[3] = {1.0, 1e-2, 1e-3};
    const long n = (long)nx * ny * nz;
    std::vector<int> rp((size_t)n + 1, 0);
    for (long i = 0; i < n; i++) {
        const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((long)nx * ny));
        const int cx = 1 + (x > 0) + (x < nx - 1), cy = 1 + (y > 0) + (y < ny - 1), cz = 1 + (z > 0) + (z < nz - 1);
        rp[i + 1] = rp[i] + cx * cy * cz;
    }
    sp_matrix_mg *A = new sp_matrix_mg((int)n, (int)n, rp[n]);
    std::copy(rp.begin(), rp.end(), A->rowptr);
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (long i = 0; i < n; i++) {
        const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((long)nx * ny));
        double st[3][3][3] = {};
        // node (x,y,z) touches the 8 cells (x+ex-1 .. ) in the padded cell grid; local node index inside a cell = 1-e
        for (int ez = 0; ez < 2; ez++)
            for (int ey = 0; ey < 2; ey++)
                for (int ex = 0; ex < 2; ex++) {
                    const double k = cell_kappa(x + ex, y + ey, z + ez, seed);
                    const int ax = 1 - ex, ay = 1 - ey, az = 1 - ez;  // my local index in that cell
                    for (int bz = 0; bz < 2; bz++)
                        for (int by = 0; by < 2; by++)
                            for (int bx = 0; bx < 2; bx++) {
                                const double v = k * (eps[0] * K1[ax][bx] * M1[ay][by] * M1[az][bz] +
                                                      eps[1] * M1[ax][bx] * K1[ay][by] * M1[az][bz] +
                                                      eps[2] * M1[ax][bx] * M1[ay][by] * K1[az][bz]);
                                st[bz - az + 1][by - ay + 1][bx - ax + 1] += v;
                            }
                }
        int o = rp[i];
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const int xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
                    A->colindex[o] = (int)(((long)zz * ny + yy) * nx + xx);
                    A->val[o++] = st[dz + 1][dy + 1][dx + 1];
                }
    }
    return A;
}

// ---------------------------------------------------------------------------------------------------------
// extern "C" driver
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int sparsh_host_set_option(const char *name, double value) {
    sparsh::Options &o = options();
    const std::string s(name);
    if (s == "threads") o.threads = (int)value;
    else if (s == "relax") o.relax = value;
    else if (s == "tol") o.tol = value;
    else if (s == "tol_mode") o.tol_mode = (int)value;
    else if (s == "coarse_upper") o.coarse_upper = (int)value;
    else if (s == "coarse_lower") o.coarse_lower = (int)value;
    else if (s == "max_levels") o.max_levels = (int)value;
    else if (s == "sweeps") o.sweeps = (int)value;
    else if (s == "print_setup") o.print_setup = (int)value;
    else if (s == "print_solve") o.print_solve = (int)value;
    else if (s == "coarsening") o.coarsening = (int)value;
    else if (s == "max_iter") o.max_iter = (int)value;
    else if (s == "use_graph") o.use_graph = (int)value;
    else if (s == "halo_mode") o.halo_mode = (int)value;
    else if (s == "device") o.device = (int)value;
    else if (s == "gpu_rap") o.gpu_rap = (int)value;
    else if (s == "tail_threshold") o.tail_threshold = (int)value;
    else if (s == "gmres_restart") o.gmres_restart = (int)value;
    else if (s == "sa_theta") o.sa_theta = value;
    else if (s == "sa_relax") o.sa_relax = value;
    else return -1;
    return 0;
}

void *sparsh_host_matrix_poisson3d(int nx, int ny, int nz) { return gen_7pt(nx, ny, nz, 6.0); }
void *sparsh_host_matrix_poisson2d(int nx, int ny) { return gen_7pt(nx, ny, 1, 4.0); }
void *sparsh_host_matrix_diffusion27(int nx, int ny, int nz, unsigned seed) { return gen_27pt_aniso(nx, ny, nz, seed); }

void *sparsh_host_matrix_from_csr(int nrow, int ncol, int nnz, const int *rp, const int *ci, const double *v) {
    sp_matrix_mg *A = new sp_matrix_mg(nrow, ncol, nnz);
    std::memcpy(A->rowptr, rp, sizeof(int) * ((size_t)nrow + 1));
    std::memcpy(A->colindex, ci, sizeof(int) * (size_t)nnz);
    std::memcpy(A->val, v, sizeof(double) * (size_t)nnz);
    return A;
}
void *sparsh_host_matrix_read(const char *matrixfile, const char *rhsfile, double **b_out) {
    sp_matrix_mg *A = nullptr;
    double *b = nullptr;
    if (rhsfile && rhsfile[0])
        readcoo(const_cast<char *>(matrixfile), const_cast<char *>(rhsfile), A, b);
    else
        read_coo_new_format(const_cast<char *>(matrixfile), A, b);
    *b_out = b;
    return A;
}
void sparsh_host_free_array(double *p) { delete[] p; }
void *sparsh_host_matrix_read_mm(const char *path) { return read_matrix_market(path); }
void *sparsh_host_matrix_read_bin(const char *path) { return read_binary_csr(path); }
int sparsh_host_matrix_write_bin(void *Av, const char *path) { return write_binary_csr(path, *(sp_matrix_mg *)Av); }

void sparsh_host_matrix_prepare(void *Av) {  // what main.cpp:21-22 does before any solver call
    sp_matrix_mg *A = (sp_matrix_mg *)Av;
    A->sp_matrix_fill();
    A->sp_matrix_fill_diagonal();
}
void sparsh_host_matrix_free(void *Av) {
    sp_matrix_mg *A = (sp_matrix_mg *)Av;
    if (!A) return;
    delete[] A->rowptr;
    delete[] A->colindex;
    delete[] A->val;
    delete A;
}
void sparsh_host_matrix_dims(void *Av, int *nrow, int *ncol, int *nnz) {
    sp_matrix_mg *A = (sp_matrix_mg *)Av;
    *nrow = A->nrow;
    *ncol = A->ncol;
    *nnz = A->rowptr[A->nrow];
}
void sparsh_host_matrix_arrays(void *Av, int **rp, int **ci, double **v, double **diag) {
    sp_matrix_mg *A = (sp_matrix_mg *)Av;
    *rp = A->rowptr;
    *ci = A->colindex;
    *v = A->val;
    *diag = A->diagonal;
}
// y = A x on the host (used only to manufacture right-hand sides b = A x*; not part of the solve path)
void sparsh_host_matrix_times(void *Av, const double *x, double *y) {
    sp_matrix_mg *A = (sp_matrix_mg *)Av;
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (int i = 0; i < A->nrow; i++) {
        double s = 0.0;
        for (int j = A->rowptr[i]; j < A->rowptr[i + 1]; j++) s += A->val[j] * x[A->colindex[j]];
        y[i] = s;
    }
}
int sparsh_host_matrix_color(void *Av, int *perm, int *color_count) {
    sp_matrix_mg *A = (sp_matrix_mg *)Av;
    A->color_matrix_and_reorder();
    std::memcpy(perm, A->color, sizeof(int) * (size_t)A->nrow);
    std::memcpy(color_count, A->color_count, sizeof(int) * ((size_t)A->total_colors + 1));
    return A->total_colors;
}

// hierarchy: AMG_GPU1_solver built by the native setup; sor != 0 uses AMG_solver_setup_SOR
void *sparsh_host_amg_setup(void *Av, int sor) {
    AMG_GPU1_solver *S = new AMG_GPU1_solver();
    if (sor)
        S->AMG_solver_setup_SOR(*(sp_matrix_mg *)Av);
    else
        S->AMG_solver_setup_jacobi(*(sp_matrix_mg *)Av);
    return S;
}
void sparsh_host_amg_free(void *Sv) { delete (AMG_GPU1_solver *)Sv; }
int sparsh_host_amg_nlevels(void *Sv) { return ((AMG_GPU1_solver *)Sv)->l + 1; }
void sparsh_host_amg_level_dims(void *Sv, int k, int *nrow, int *nnz, int *p_ncol, int *p_nnz) {
    AMG_GPU1_solver *S = (AMG_GPU1_solver *)Sv;
    *nrow = S->Av[k]->nrow;
    *nnz = S->Av[k]->rowptr[S->Av[k]->nrow];
    *p_ncol = k < S->l ? S->Pv[k]->ncol : 0;
    *p_nnz = k < S->l ? S->Pv[k]->rowptr[S->Pv[k]->nrow] : 0;
}
void sparsh_host_amg_level_arrays(void *Sv, int k, int **rp, int **ci, double **v, double **diag, int **prp, int **pci,
                                  double **pv) {
    AMG_GPU1_solver *S = (AMG_GPU1_solver *)Sv;
    *rp = S->Av[k]->rowptr;
    *ci = S->Av[k]->colindex;
    *v = S->Av[k]->val;
    *diag = S->Av[k]->diagonal;
    *prp = k < S->l ? S->Pv[k]->rowptr : nullptr;
    *pci = k < S->l ? S->Pv[k]->colindex : nullptr;
    *pv = k < S->l ? S->Pv[k]->val : nullptr;
}
void sparsh_host_amg_upload(void *Sv) { ((AMG_GPU1_solver *)Sv)->GPU_Allocations(); }
void *sparsh_host_amg_device(void *Sv) { return ((AMG_GPU1_solver *)Sv)->device; }  // sparsh_hierarchy_t

// the reference-named entry points, by name; b is copied (AMG_Solver_2 permutes it), x is in/out
int sparsh_host_call(const char *name, void *Av, const double *b, double *x) {
    sp_matrix_mg &A = *(sp_matrix_mg *)Av;
    std::vector<double> bc(b, b + A.nrow);
    double *bp = bc.data(), *xp = x;
    const std::string s(name);
    typedef void (*fn_t)(sp_matrix_mg &, double *&, double *&);
    struct Entry {
        const char *n;
        fn_t f;
    };
    static const Entry table[] = {{"AMG_Solver_CPU_baseline", AMG_Solver_CPU_baseline},
                                  {"AMG_Solver_1", AMG_Solver_1},
                                  {"AMG_Solver_2", AMG_Solver_2},
                                  {"AMG_Solver_CPU_GPU_CI", AMG_Solver_CPU_GPU_CI},
                                  {"AMG_Solver_CPU_GPU_MI", AMG_Solver_CPU_GPU_MI},
                                  {"Solver_CG_1", Solver_CG_1},
                                  {"Solver_CG_2", Solver_CG_2},
                                  {"Solver_PCG_1", Solver_PCG_1},
                                  {"Solver_PCG_2", Solver_PCG_2},
                                  {"Solver_PCG_3", Solver_PCG_3},
                                  {"Solver_PCG_4", Solver_PCG_4},
                                  {"Solver_BiCG_1", Solver_BiCG_1},
                                  {"Solver_PBiCG_1", Solver_PBiCG_1},
                                  {"Solver_PBiCG_2", Solver_PBiCG_2},
                                  {"Solver_PBiCG_3", Solver_PBiCG_3},
                                  {"Solver_PBiCG_4", Solver_PBiCG_4},
                                  {"Solver_GMRES_1", Solver_GMRES_1},
                                  {"Solver_PGMRES_1", Solver_PGMRES_1},
                                  {"coarsening_2", coarsening_2}};
    for (const Entry &e : table)
        if (s == e.n) {
            e.f(A, bp, xp);
            return 0;
        }
    return -1;
}

int sparsh_host_report(int *iterations, int *converged, double *setup_s, double *upload_s, double *solve_s,
                       double *hist, int maxhist) {
    const sparsh::Report &r = sparsh::last_report();
    *iterations = r.iterations;
    *converged = r.converged;
    *setup_s = r.setup_seconds;
    *upload_s = r.upload_seconds;
    *solve_s = r.solve_seconds;
    const int k = (int)std::min<size_t>(r.history.size(), (size_t)maxhist);
    for (int i = 0; i < k; i++) hist[i] = r.history[i];
    return (int)r.history.size();
}

}  // extern "C"
