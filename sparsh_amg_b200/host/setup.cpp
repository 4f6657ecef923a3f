// setup.cpp — matrices and the host setup phase (coarsening + Galerkin products) behind the reference's names.
//
// The solve phase is the product of this repository; the setup phase only has to hand it the same hierarchy the
// reference would build.  These are native re-implementations (no MKL): integer outputs (aggregates, C/F splitting,
// colour permutation, sparsity patterns) are bit-identical to the reference's, coarse values agree to rounding.
//   sp_matrix / sp_matrix_mg            reference src/AMG_matrix.cpp:15-67, src/AMG_cpu_matrix.cpp:17-237
//   sequential::HEM_Prolongator         reference src/AMG_coarsening.cpp:14-97
//   sequential::beck_prolongator        reference src/AMG_coarsening.cpp:269-339
//   parallel::coarsen_matrix            reference src/AMG_cycle_utilities.cpp:126-146
//   AMG_solver::AMG_solver_setup_*      reference src/AMG_phases.cpp:35-147
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <numeric>
#include <cstdio>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../../include/sparsh_b200.h"
#include "sparsh_amg.hpp"

namespace sparsh {
Options &options() {
    static Options o;
    return o;
}
Report &last_report() {
    static Report r;
    return r;
}
}  // namespace sparsh

using sparsh::options;

// ---------------------------------------------------------------------------------------------------------
// sp_matrix / sp_matrix_mg
// ---------------------------------------------------------------------------------------------------------
sp_matrix::sp_matrix(int r, int c, int n) : nrow(r), ncol(c), nnz(n) {
    rowptr = new int[(size_t)r + 1]();
    colindex = new int[(size_t)(n > 0 ? n : 1)]();
    val = new double[(size_t)(n > 0 ? n : 1)]();
}
sp_matrix::sp_matrix() {}

void sp_matrix::check_sp_matrix() {
    std::cout << "\n Number of Rows: " << nrow << "\n Number of columns " << ncol << "\n Number of non-zeros " << nnz
              << std::endl;
    for (int i = 0; i < nrow; i++) {
        std::cout << i << std::endl;
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++) std::cout << colindex[j] << "\t" << val[j] << "\t" << std::endl;
        std::cout << std::endl << std::endl;
    }
}

static void sort_row(int len, int *c, double *v) {
    if (len <= 32) {
        for (int a = 1; a < len; a++) {
            const int cc = c[a];
            const double vv = v[a];
            int b = a - 1;
            while (b >= 0 && c[b] > cc) {
                c[b + 1] = c[b];
                v[b + 1] = v[b];
                b--;
            }
            c[b + 1] = cc;
            v[b + 1] = vv;
        }
        return;
    }
    std::vector<int> idx(len);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int p, int q) { return c[p] < c[q]; });
    std::vector<int> cc(c, c + len);
    std::vector<double> vv(v, v + len);
    for (int k = 0; k < len; k++) {
        c[k] = cc[idx[k]];
        v[k] = vv[idx[k]];
    }
}

static void sort_columns(int nrow, const int *rp, int *ci, double *v) {
#pragma omp parallel for num_threads(options().threads) schedule(dynamic, 4096)
    for (int i = 0; i < nrow; i++) sort_row(rp[i + 1] - rp[i], ci + rp[i], v + rp[i]);
}

// The reference wraps the arrays in an MKL handle and lets mkl_sparse_order sort each row's columns in place
// (src/AMG_cpu_matrix.cpp:22-29); only the sort has an observable effect.
void sp_matrix_mg::sp_matrix_fill() {
    nnz = rowptr[nrow];
    sort_columns(nrow, rowptr, colindex, val);
}

// src/AMG_cpu_matrix.cpp:35-51: first stored entry whose column equals the row
void sp_matrix_mg::sp_matrix_fill_diagonal() {
    delete[] diagonal;
    delete[] helper;
    diagonal = new double[(size_t)std::max(nrow, 1)];  // zeroed by the threads below: they touch the pages first
    helper = new double[(size_t)std::max(nrow, 1)];
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (int i = 0; i < nrow; i++) {
        double d = 0.0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
            if (colindex[j] == i) {
                d = val[j];
                break;
            }
        diagonal[i] = d;
        helper[i] = 0.0;
    }
}

sp_matrix_mg::~sp_matrix_mg() {
    // tolerate the reference's idiom of explicit destructor calls followed by nothing (main.cpp:40)
    delete[] diagonal;
    delete[] helper;
    delete[] entries;
    delete[] color;
    delete[] color_count;
    diagonal = helper = entries = nullptr;
    color = color_count = nullptr;
}

// src/AMG_cpu_matrix.cpp:81-199: greedy first-fit colouring in natural order over the stored columns; `color` then
// becomes perm[new] = old (grouped by colour, ascending inside a colour), color_count the prefix offsets, and the
// matrix is permuted symmetrically in place.
void sp_matrix_mg::color_matrix_and_reorder() {
    const int n = nrow;
    int max_count = 0;
    for (int i = 0; i < n; i++) max_count = std::max(max_count, rowptr[i + 1] - rowptr[i]);
    std::vector<int> col_of(n, 0), forbidden(max_count + 1, -1), cc(max_count + 1, 0);
    total_colors = 0;
    for (int i = 0; i < n; i++) {
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
            if (col_of[colindex[j]] != 0) forbidden[col_of[colindex[j]]] = i;
        int c = 0x7fffffff;
        for (int k = 1; k < max_count + 1; k++)
            if (forbidden[k] != i) {
                c = k;
                break;
            }
        col_of[i] = c;
        cc[c]++;
        total_colors = std::max(total_colors, c);
    }
    for (int k = 0; k < total_colors; k++) cc[k + 1] += cc[k];
    delete[] color;
    delete[] color_count;
    color = new int[(size_t)n];
    color_count = new int[(size_t)max_count + 1]();
    std::vector<int> cur(total_colors + 1, 0);
    for (int k = 1; k <= total_colors; k++) cur[k] = cc[k - 1];
    for (int i = 0; i < n; i++) color[cur[col_of[i]]++] = i;
    for (int k = 0; k <= total_colors; k++) color_count[k] = cc[k];

    std::vector<int> inv(n);
    for (int i = 0; i < n; i++) inv[color[i]] = i;
    int *qrp = new int[(size_t)n + 1];
    qrp[0] = 0;
    for (int i = 0; i < n; i++) qrp[i + 1] = qrp[i] + (rowptr[color[i] + 1] - rowptr[color[i]]);
    int *qci = new int[(size_t)std::max(qrp[n], 1)];
    double *qv = new double[(size_t)std::max(qrp[n], 1)];
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (int i = 0; i < n; i++) {
        int o = qrp[i];
        for (int j = rowptr[color[i]]; j < rowptr[color[i] + 1]; j++, o++) {
            qci[o] = inv[colindex[j]];
            qv[o] = val[j];
        }
    }
    // the reference re-points rowptr/colindex/val at the product's arrays (and leaks the old ones); we own ours
    delete[] rowptr;
    delete[] colindex;
    delete[] val;
    rowptr = qrp;
    colindex = qci;
    val = qv;
    sp_matrix_fill();
    sp_matrix_fill_diagonal();
}

// src/AMG_cpu_matrix.cpp:203-219
void sp_matrix_mg::normalize_matrix() {
    std::vector<double> norm1(ncol, 0.0);
    for (int i = 0; i < nrow; i++)
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++) norm1[colindex[j]] += val[j] * val[j];
    for (int i = 0; i < nrow; i++)
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++) val[j] = val[j] / norm1[colindex[j]];
}

// src/AMG_cpu_matrix.cpp:223-237
void sp_matrix_mg::scale_system(double *&b) {
    for (int i = 0; i < nrow; i++) {
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++) val[j] = val[j] / diagonal[i];
        b[i] = b[i] / std::sqrt(diagonal[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// coarsening
// ---------------------------------------------------------------------------------------------------------
namespace sequential {

constexpr int PAT_SA_MAXROW = 512;  // distinct aggregates one fine row may touch (strong neighbours + itself)

// Heavy-edge matching.  Forward sweep on even levels, backward on odd ones; a free row pairs with its free neighbour of
// strictly largest |a_ij| (ties: first in column order; zeros never); leftovers become singletons numbered last.
// P is n x n_coarse with a single 1.0 per row.  The greedy sweep is inherently sequential, O(nnz).
void HEM_Prolongator(sp_matrix_mg &A, sp_matrix_mg *&P, int l1) {
    const int n = A.nrow;
    P = new sp_matrix_mg();
    P->nrow = n;
    P->ncol = 1;
    P->nnz = n;
    P->rowptr = new int[(size_t)n + 1];
    P->colindex = new int[(size_t)std::max(n, 1)];
    P->val = new double[(size_t)std::max(n, 1)];
    int *agg = P->colindex;
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (int i = 0; i < n; i++) {  // the arrays are written here for the first time, by all threads
        agg[i] = -1;
        P->val[i] = 1.0;
        P->rowptr[i] = i;
    }
    P->rowptr[n] = n;
    int next_id = 0;
    const int first = (l1 % 2 == 0) ? 0 : n - 1, step = (l1 % 2 == 0) ? 1 : -1;
    const int *__restrict rp = A.rowptr, *__restrict ci = A.colindex;
    const double *__restrict av = A.val;
    for (int t = 0, i = first; t < n; t++, i += step) {
        if (agg[i] != -1) continue;
        int mate = -1;
        double heaviest = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++) {  // selects instead of branches: the outcome is data, not pattern
            const int c = ci[j];
            const double w = std::fabs(av[j]);
            const bool take = (agg[c] == -1) & (w > heaviest) & (c != i);
            heaviest = take ? w : heaviest;
            mate = take ? c : mate;
        }
        if (mate != -1) {
            agg[i] = next_id;
            agg[mate] = next_id;
            next_id++;
        }
    }
    for (int i = 0; i < n; i++)
        if (agg[i] == -1) agg[i] = next_id++;
    P->ncol = next_id;
    P->nnz = n;
}

// Beck's classical coarsening: first-fit C-point selection in natural order (a row still at 0 becomes a C point and
// decrements every stored neighbour); C rows are identity rows, an F row averages its C neighbours with weight
// 1/|c_f[i]|.  Columns sorted per row.
void beck_prolongator(sp_matrix_mg &A, sp_matrix_mg *&P1) {
    const int n = A.nrow;
    std::vector<int> cf(n, 0);
    int ncoarse = 0;
    for (int i = 0; i < n; i++) {
        if (cf[i] != 0) continue;
        for (int j = A.rowptr[i]; j < A.rowptr[i + 1]; j++) cf[A.colindex[j]] -= 1;
        cf[i] = ++ncoarse;
    }
    std::vector<int> rp(n + 1, 0);
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (int i = 0; i < n; i++) {
        int cnt = 0;
        if (cf[i] > 0)
            cnt = 1;
        else if (cf[i] < 0)
            for (int j = A.rowptr[i]; j < A.rowptr[i + 1]; j++) cnt += cf[A.colindex[j]] > 0;
        rp[i + 1] = cnt;
    }
    for (int i = 0; i < n; i++) rp[i + 1] += rp[i];
    P1 = new sp_matrix_mg(n, ncoarse, rp[n]);
    std::copy(rp.begin(), rp.end(), P1->rowptr);
#pragma omp parallel for num_threads(options().threads) schedule(static)
    for (int i = 0; i < n; i++) {
        int o = rp[i];
        if (cf[i] > 0) {
            P1->colindex[o] = cf[i] - 1;
            P1->val[o] = 1.0;
        } else if (cf[i] < 0) {
            const double w = 1 / std::fabs((double)cf[i]);
            for (int j = A.rowptr[i]; j < A.rowptr[i + 1]; j++) {
                const int k = A.colindex[j];
                if (cf[k] > 0) {
                    P1->colindex[o] = cf[k] - 1;
                    P1->val[o] = w;
                    o++;
                }
            }
        }
    }
    P1->sp_matrix_fill();
}

// Smoothed aggregation (Vanek, Mandel, Brezina 1996) — SURVEY §8f.2: advertised by the reference's README
// (README.md:10,13) but absent from its sources (F2), so there is nothing to restate; this is the textbook algorithm,
// made deterministic (natural order everywhere, no random vectors):
//   strength   j is a strong neighbour of i  iff  j != i and |a_ij| >= theta_l * sqrt(|a_ii| |a_jj|),
//              theta_l = sa_theta * 0.5^level
//   pass 1     a free row whose strong neighbours are all free roots a new aggregate with them
//   pass 2     a row still free joins the pass-1 aggregate of its strong neighbour of largest |a_ij| (ties: first)
//   pass 3     leftovers root aggregates with their free strong neighbours (rows without any: singletons)
//   tentative  T: one 1.0 per row (the constant near-null vector, columns left unnormalised like HEM's P)
//   smoothing  P = (I - omega D_F^-1 A_F) T,  A_F = A with the weak off-diagonals lumped onto the diagonal,
//              omega = sa_relax / rho,  rho = max_i sum_j |a_F,ij| / |a_F,ii|  (bound on rho(D_F^-1 A_F), = 2 for Poisson)
void SA_Prolongator(sp_matrix_mg &A, sp_matrix_mg *&P, int level) {
    const sparsh::Options &o = options();
    const int n = A.nrow;
    const int *rp = A.rowptr, *ci = A.colindex;
    const double *v = A.val;
    std::vector<double> dabs((size_t)n, 0.0);
    for (int i = 0; i < n; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (ci[j] == i) {
                dabs[i] = std::fabs(v[j]);
                break;
            }
    const double theta = o.sa_theta * std::pow(0.5, level);
    std::vector<char> strong((size_t)std::max(rp[n], 1), 0);
#pragma omp parallel for num_threads(o.threads) schedule(static)
    for (int i = 0; i < n; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            const int c = ci[j];
            strong[j] = c != i && v[j] != 0.0 && std::fabs(v[j]) >= theta * std::sqrt(dabs[i] * dabs[c]);
        }
    std::vector<int> agg((size_t)n, -1);
    int nagg = 0;
    for (int i = 0; i < n; i++) {  // pass 1
        if (agg[i] != -1) continue;
        bool any = false, all_free = true;
        for (int j = rp[i]; j < rp[i + 1] && all_free; j++)
            if (strong[j]) {
                any = true;
                all_free = agg[ci[j]] == -1;
            }
        if (!any || !all_free) continue;
        agg[i] = nagg;
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (strong[j]) agg[ci[j]] = nagg;
        nagg++;
    }
    {  // pass 2 (against the snapshot left by pass 1, so the result does not depend on the sweep direction)
        const std::vector<int> snap(agg);
        for (int i = 0; i < n; i++) {
            if (snap[i] != -1) continue;
            double best = 0.0;
            int to = -1;
            for (int j = rp[i]; j < rp[i + 1]; j++)
                if (strong[j] && snap[ci[j]] != -1 && std::fabs(v[j]) > best) {
                    best = std::fabs(v[j]);
                    to = snap[ci[j]];
                }
            if (to != -1) agg[i] = to;
        }
    }
    for (int i = 0; i < n; i++) {  // pass 3
        if (agg[i] != -1) continue;
        agg[i] = nagg;
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (strong[j] && agg[ci[j]] == -1) agg[ci[j]] = nagg;
        nagg++;
    }
    // filtered diagonal and the bound on rho(D_F^-1 A_F)
    std::vector<double> dF((size_t)n, 0.0);
    double rho = 0.0;
#pragma omp parallel for num_threads(o.threads) schedule(static) reduction(max : rho)
    for (int i = 0; i < n; i++) {
        double d = 0.0, off = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            if (ci[j] == i || !strong[j])
                d += v[j];  // weak couplings are lumped
            else
                off += std::fabs(v[j]);
        }
        dF[i] = d;
        if (d != 0.0) rho = std::max(rho, (std::fabs(d) + off) / std::fabs(d));
    }
    const double omega = rho > 0.0 ? o.sa_relax / rho : 0.0;
    // P row i = e_agg[i] - (omega / dF_i) * sum_j aF_ij e_agg[j]; entries merged per aggregate in column order of A
    std::vector<int> cnt((size_t)n + 1, 0);
#pragma omp parallel for num_threads(o.threads) schedule(static)
    for (int i = 0; i < n; i++) {
        int ids[PAT_SA_MAXROW];
        int m = 0;
        auto touch = [&](int a) {
            for (int t = 0; t < m; t++)
                if (ids[t] == a) return;
            if (m < PAT_SA_MAXROW) ids[m++] = a;
        };
        touch(agg[i]);
        if (dF[i] != 0.0)
            for (int j = rp[i]; j < rp[i + 1]; j++)
                if (strong[j]) touch(agg[ci[j]]);
        cnt[i + 1] = m;
    }
    for (int i = 0; i < n; i++) cnt[i + 1] += cnt[i];
    P = new sp_matrix_mg(n, nagg, cnt[n]);
    std::copy(cnt.begin(), cnt.end(), P->rowptr);
#pragma omp parallel for num_threads(o.threads) schedule(static)
    for (int i = 0; i < n; i++) {
        int *pc = P->colindex + cnt[i];
        double *pv = P->val + cnt[i];
        const int cap = cnt[i + 1] - cnt[i];
        int m = 0;
        auto add = [&](int a, double w) {
            for (int t = 0; t < m; t++)
                if (pc[t] == a) {
                    pv[t] += w;
                    return;
                }
            if (m < cap) {
                pc[m] = a;
                pv[m] = w;
                m++;
            }
        };
        add(agg[i], 1.0);
        if (dF[i] != 0.0) {
            const double s = omega / dF[i];
            add(agg[i], -s * dF[i]);  // the (filtered) diagonal term
            for (int j = rp[i]; j < rp[i + 1]; j++)
                if (strong[j]) add(agg[ci[j]], -s * v[j]);
        }
    }
    P->sp_matrix_fill();
}

}  // namespace sequential

// ---------------------------------------------------------------------------------------------------------
// Galerkin product
// ---------------------------------------------------------------------------------------------------------
namespace {

// intermediate CSR whose index/value arrays are allocated WITHOUT being zero-filled: the threads that compute the
// entries are the first to touch the pages (a std::vector would fault in and clear gigabytes on one thread first)
struct Csr {
    int nrow = 0, ncol = 0;
    std::vector<int> rp;
    std::unique_ptr<int[]> ci;
    std::unique_ptr<double[]> v;
    void alloc(size_t nnz) {
        ci.reset(new int[std::max<size_t>(nnz, 1)]);
        v.reset(new double[std::max<size_t>(nnz, 1)]);
    }
};

// rp[0] = 0, rp[i + 1] holds a count on entry and the inclusive running sum on return (block sums, then offsets)
void counts_to_offsets(std::vector<int> &rp) {
    const long long n = (long long)rp.size() - 1;
    const int nt = std::max(1, options().threads);
    if (n < (1 << 16) || nt == 1) {
        for (long long i = 0; i < n; i++) rp[(size_t)i + 1] += rp[(size_t)i];
        return;
    }
    std::vector<long long> block((size_t)nt + 1, 0);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num(), n_t = omp_get_num_threads();
        const long long lo = n * t / n_t, hi = n * (t + 1) / n_t;
        long long sum = 0;
        for (long long i = lo; i < hi; i++) sum += rp[(size_t)i + 1];
        block[(size_t)t + 1] = sum;
#pragma omp barrier
#pragma omp single
        for (int k = 0; k < n_t; k++) block[(size_t)k + 1] += block[k];
        long long run = block[t];
        for (long long i = lo; i < hi; i++) {
            run += rp[(size_t)i + 1];
            rp[(size_t)i + 1] = (int)run;
        }
    }
}

// row-wise (Gustavson) C = A * B.  The entries of a C row appear in first-touch order and each is accumulated in
// traversal order, which fixes the rounding independently of the thread count.  Rows of the Galerkin products are
// short (tens of entries), so the accumulator of a row is a small list searched linearly — it stays in L1, where a
// dense position table of B's width per thread (tens of MB at 256^3) misses on every touch; rows that outgrow the
// list fall back to such a table, allocated by the thread on first need.  Same order either way.
void spgemm(int arow, const int *arp, const int *aci, const double *av, int bcol, const int *brp, const int *bci,
            const double *bv, Csr &C) {
    constexpr int SMALL = 96;
    C.nrow = arow;
    C.ncol = bcol;
    C.rp.assign((size_t)arow + 1, 0);
    const int nt = options().threads;
#pragma omp parallel num_threads(nt)
    {
        std::vector<int> mark;  // dense fallback
        int list[SMALL];
#pragma omp for schedule(dynamic, 4096)
        for (int i = 0; i < arow; i++) {
            int cnt = 0;
            bool dense = false;
            for (int ja = arp[i]; ja < arp[i + 1] && !dense; ja++) {
                const int k = aci[ja];
                for (int jb = brp[k]; jb < brp[k + 1]; jb++) {
                    const int c = bci[jb];
                    int t = 0;
                    while (t < cnt && list[t] != c) t++;
                    if (t == cnt) {
                        if (cnt == SMALL) {
                            dense = true;
                            break;
                        }
                        list[cnt++] = c;
                    }
                }
            }
            if (dense) {
                if (mark.empty()) mark.assign((size_t)std::max(bcol, 1), -1);
                cnt = 0;
                for (int ja = arp[i]; ja < arp[i + 1]; ja++) {
                    const int k = aci[ja];
                    for (int jb = brp[k]; jb < brp[k + 1]; jb++)
                        if (mark[bci[jb]] != i) {
                            mark[bci[jb]] = i;
                            cnt++;
                        }
                }
            }
            C.rp[i + 1] = cnt;
        }
    }
    counts_to_offsets(C.rp);
    C.alloc((size_t)C.rp[arow]);
#pragma omp parallel num_threads(nt)
    {
        std::vector<int> pos;  // dense fallback: position of a column inside the current row, valid when >= base
#pragma omp for schedule(dynamic, 4096)
        for (int i = 0; i < arow; i++) {
            const int base = C.rp[i], len = C.rp[i + 1] - base;
            int *cc = C.ci.get() + base;
            double *cv = C.v.get() + base;
            int o = 0;
            if (len <= SMALL) {
                for (int ja = arp[i]; ja < arp[i + 1]; ja++) {
                    const int k = aci[ja];
                    const double a = av[ja];
                    for (int jb = brp[k]; jb < brp[k + 1]; jb++) {
                        const int c = bci[jb];
                        int t = 0;
                        while (t < o && cc[t] != c) t++;
                        if (t == o) {
                            cc[o] = c;
                            cv[o] = a * bv[jb];
                            o++;
                        } else {
                            cv[t] += a * bv[jb];
                        }
                    }
                }
            } else {
                if (pos.empty()) pos.assign((size_t)std::max(bcol, 1), -1);
                for (int ja = arp[i]; ja < arp[i + 1]; ja++) {
                    const int k = aci[ja];
                    const double a = av[ja];
                    for (int jb = brp[k]; jb < brp[k + 1]; jb++) {
                        const int c = bci[jb];
                        if (pos[c] < base) {
                            pos[c] = base + o;
                            cc[o] = c;
                            cv[o] = a * bv[jb];
                            o++;
                        } else {
                            cv[pos[c] - base] += a * bv[jb];
                        }
                    }
                }
            }
        }
    }
}

// T = M^T, stable: row c of T lists the rows of M in ascending order.  Every thread owns a contiguous range of T's rows
// and scans all entries of M in order, keeping those that fall into its range — nt reads of the (small) index array
// instead of one sequential counting sort.
void transpose(int nrow, int ncol, const int *rp, const int *ci, const double *v, Csr &T) {
    const int nnz = rp[nrow];
    T.nrow = ncol;
    T.ncol = nrow;
    T.rp.assign((size_t)ncol + 1, 0);
    T.alloc((size_t)nnz);
    for (int j = 0; j < nnz; j++) T.rp[ci[j] + 1]++;
    for (int c = 0; c < ncol; c++) T.rp[c + 1] += T.rp[c];
    const int nt = std::max(1, options().threads);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num(), n_t = omp_get_num_threads();
        const int c0 = (int)((long long)ncol * t / n_t), c1 = (int)((long long)ncol * (t + 1) / n_t);
        std::vector<int> cur(T.rp.begin() + c0, T.rp.begin() + c1);
        for (int i = 0; i < nrow; i++)
            for (int j = rp[i]; j < rp[i + 1]; j++) {
                const int c = ci[j];
                if (c >= c0 && c < c1) {
                    const int d = cur[c - c0]++;
                    T.ci[d] = i;
                    T.v[d] = v[j];
                }
            }
    }
}


// C = P^T (A P) when P has exactly ONE entry per row (aggregation: HEM).  The two products collapse into one sweep over
// the aggregates without A P or P^T ever being stored, yet every number is produced by the operations, in the order, of
// the two general products above: per fine row i the sums s_i[col] = sum_j a_ij * p_j over the entries of the row that
// fall into the same aggregate (first-touch order, accumulated in traversal order — row i of A P), then per aggregate c
// its fine rows in ascending order (row c of the stable P^T) add p_i * s_i[col] into the coarse row, again first touch
// first.  Same bits, half the memory traffic.  false: not applicable (P is not an aggregation, or a row outgrows the
// small lists) — the caller runs the general products.
bool aggregation_rap(int n, const int *arp, const int *aci, const double *av, int nc, const int *prp, const int *agg,
                     const double *pv, Csr &C) {
    constexpr int SMALL = 96, ROWMAX = 128;
    const int nt = std::max(1, options().threads);
    if (n == 0 || nc == 0 || prp[n] != n) return false;
    int bad = 0;
#pragma omp parallel for num_threads(nt) schedule(static) reduction(| : bad)
    for (int i = 0; i < n; i++) bad |= (prp[i] != i) | (arp[i + 1] - arp[i] > ROWMAX) | (agg[i] < 0) | (agg[i] >= nc);
    if (bad) return false;
    // the fine rows of every aggregate, ascending
    const double tt0 = omp_get_wtime();
    std::vector<int> mrp((size_t)nc + 1, 0);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int i = 0; i < n; i++) {
#pragma omp atomic
        mrp[(size_t)agg[i] + 1]++;
    }
    counts_to_offsets(mrp);
    std::unique_ptr<int[]> mem(new int[(size_t)n]);
    {
        std::unique_ptr<int[]> cur(new int[(size_t)nc]);
#pragma omp parallel num_threads(nt)
        {
#pragma omp for schedule(static)
            for (int c = 0; c < nc; c++) cur[c] = mrp[c];
#pragma omp for schedule(static)
            for (int i = 0; i < n; i++) {
                int d;
#pragma omp atomic capture
                d = cur[agg[i]]++;
                mem[(size_t)d] = i;
            }
#pragma omp for schedule(static)
            for (int c = 0; c < nc; c++)  // threads arrive in any order: ascending inside an aggregate (they are small)
                for (int a = mrp[c] + 1; a < mrp[c + 1]; a++) {
                    const int x = mem[a];
                    int b = a - 1;
                    while (b >= mrp[c] && mem[b] > x) {
                        mem[b + 1] = mem[b];
                        b--;
                    }
                    mem[b + 1] = x;
                }
        }
    }
    const double tt1 = omp_get_wtime();
    C.nrow = nc;
    C.ncol = nc;
    C.rp.assign((size_t)nc + 1, 0);
    int overflow = 0;
#pragma omp parallel num_threads(nt) reduction(| : overflow)
    {
        int list[SMALL];
#pragma omp for schedule(dynamic, 2048)
        for (int c = 0; c < nc; c++) {
            int cnt = 0;
            for (int m = mrp[c]; m < mrp[c + 1] && cnt >= 0; m++) {
                const int i = mem[m];
                for (int j = arp[i]; j < arp[i + 1]; j++) {
                    const int col = agg[aci[j]];
                    int t = 0;
                    while (t < cnt && list[t] != col) t++;
                    if (t == cnt) {
                        if (cnt == SMALL) {
                            cnt = -1;
                            break;
                        }
                        list[cnt++] = col;
                    }
                }
            }
            if (cnt < 0) {
                overflow = 1;
                cnt = 0;
            }
            C.rp[(size_t)c + 1] = cnt;
        }
    }
    if (overflow) return false;
    const double tt2 = omp_get_wtime();
    counts_to_offsets(C.rp);
    C.alloc((size_t)C.rp[nc]);
    const double tt3 = omp_get_wtime();
#pragma omp parallel num_threads(nt)
    {
        int col1[ROWMAX];
        double s1[ROWMAX];
#pragma omp for schedule(dynamic, 2048)
        for (int c = 0; c < nc; c++) {
            int *cc = C.ci.get() + C.rp[c];
            double *cv = C.v.get() + C.rp[c];
            int o = 0;
            for (int m = mrp[c]; m < mrp[c + 1]; m++) {
                const int i = mem[m];
                int l1 = 0;  // row i of A P
                for (int j = arp[i]; j < arp[i + 1]; j++) {
                    const int k = aci[j], col = agg[k];
                    const double prod = av[j] * pv[k];
                    int t = 0;
                    while (t < l1 && col1[t] != col) t++;
                    if (t == l1) {
                        col1[l1] = col;
                        s1[l1] = prod;
                        l1++;
                    } else {
                        s1[t] += prod;
                    }
                }
                const double r = pv[i];  // the entry of P^T
                for (int q = 0; q < l1; q++) {
                    const int col = col1[q];
                    const double x = r * s1[q];
                    int t = 0;
                    while (t < o && cc[t] != col) t++;
                    if (t == o) {
                        cc[o] = col;
                        cv[o] = x;
                        o++;
                    } else {
                        cv[t] += x;
                    }
                }
            }
        }
    }
    if (getenv("SPARSH_SETUP_TIMING"))
        std::cout << "    fused: members " << tt1 - tt0 << " count " << tt2 - tt1 << " scan " << tt3 - tt2 << " fill " << omp_get_wtime() - tt3 << std::endl;
    return true;
}

}  // namespace

// Products that the device computed stay there until the next level has used them as its fine matrix (options().gpu_rap):
// keyed by the host matrix they were fetched into.  A matrix that is modified afterwards (the colouring reorders it) must
// be forgotten first.
namespace {
std::unordered_map<const sp_matrix_mg *, sparsh_rap_t> g_device_products;
sparsh_rap_t device_product_of(const sp_matrix_mg *A) {
    auto it = g_device_products.find(A);
    return it == g_device_products.end() ? nullptr : it->second;
}
void remember_device_product(const sp_matrix_mg *A, sparsh_rap_t h) { g_device_products[A] = h; }
void forget_device_product(const sp_matrix_mg *A) {
    auto it = g_device_products.find(A);
    if (it == g_device_products.end()) return;
    sparsh_rap_destroy(it->second);
    g_device_products.erase(it);
}
void forget_all_device_products() {
    for (auto &kv : g_device_products) sparsh_rap_destroy(kv.second);
    g_device_products.clear();
}
}  // namespace

namespace parallel {

// Ac = P^T (A P), columns sorted, diagonal extracted
void coarsen_matrix(sp_matrix_mg &A, sp_matrix_mg *&Ac, sp_matrix_mg &P1) {
    Csr AP, R, C;
    const bool tm = getenv("SPARSH_SETUP_TIMING") != nullptr;
    double t0 = omp_get_wtime();
    if (options().gpu_rap) {  // same product, computed by the device (csrc/rap.cu): identical integers and values
        sparsh_rap_t rap = nullptr;
        int cnnz = -1;
        // the fine matrix of this level is the product of the previous one when that is still on the device
        sparsh_rap_t fine = device_product_of(&A);
        const int rc = fine ? sparsh_galerkin_rap_next(fine, P1.ncol, P1.rowptr, P1.colindex, P1.val, &rap, &cnnz)
                            : sparsh_galerkin_rap(A.nrow, A.rowptr, A.colindex, A.val, P1.ncol, P1.rowptr, P1.colindex,
                                                  P1.val, &rap, &cnnz);
        forget_device_product(&A);
        if (rc != SPARSH_OK) {
            std::fprintf(stderr, "sparsh_amg: device Galerkin product failed (%d): %s\n", rc, sparsh_last_error());
            std::exit(1);
        }
        if (rap) {
            Ac = new sp_matrix_mg();
            Ac->nrow = Ac->ncol = P1.ncol;
            Ac->nnz = cnnz;
            Ac->rowptr = new int[(size_t)P1.ncol + 1];
            Ac->colindex = new int[(size_t)std::max(cnnz, 1)];
            Ac->val = new double[(size_t)std::max(cnnz, 1)];
            if (sparsh_rap_fetch(rap, Ac->rowptr, Ac->colindex, Ac->val) != SPARSH_OK) {
                std::fprintf(stderr, "sparsh_amg: device Galerkin product: fetch failed: %s\n", sparsh_last_error());
                std::exit(1);
            }
            remember_device_product(Ac, rap);  // released by the next level's product or at the end of the setup
            Ac->sp_matrix_fill_diagonal();
            if (tm) std::cout << "  RAP on the device: " << omp_get_wtime() - t0 << std::endl;
            return;
        }  // else: a product row is too long for the device kernel — the host product below
    }
    // aggregation prolongators (one entry per row: HEM) take the fused sweep, everything else the two general products;
    // SPARSH_RAP_GENERAL=1 forces the latter (tests compare the two bit for bit)
    const char *force_general = getenv("SPARSH_RAP_GENERAL");
    const bool fused = !(force_general && atoi(force_general) != 0) &&
                       aggregation_rap(A.nrow, A.rowptr, A.colindex, A.val, P1.ncol, P1.rowptr, P1.colindex, P1.val, C);
    double t1 = omp_get_wtime(), t2 = t1;
    if (!fused) {
        spgemm(A.nrow, A.rowptr, A.colindex, A.val, P1.ncol, P1.rowptr, P1.colindex, P1.val, AP);
        t1 = omp_get_wtime();
        transpose(P1.nrow, P1.ncol, P1.rowptr, P1.colindex, P1.val, R);
        t2 = omp_get_wtime();
        spgemm(R.nrow, R.rp.data(), R.ci.get(), R.v.get(), P1.ncol, AP.rp.data(), AP.ci.get(), AP.v.get(), C);
    }
    double t3 = omp_get_wtime();
    const int nc = P1.ncol, cnnz = C.rp[nc];
    Ac = new sp_matrix_mg();  // adopts the product's arrays (both sides use new[] / delete[])
    Ac->nrow = nc;
    Ac->ncol = nc;
    Ac->nnz = cnnz;
    Ac->rowptr = new int[(size_t)nc + 1];
    std::copy(C.rp.begin(), C.rp.end(), Ac->rowptr);
    Ac->colindex = C.ci.release();
    Ac->val = C.v.release();
    double t4 = omp_get_wtime();
    Ac->sp_matrix_fill();
    double t5 = omp_get_wtime();
    Ac->sp_matrix_fill_diagonal();
    if (tm)
        std::cout << (fused ? "  RAP (fused aggregation sweep " : "  RAP (general: A*P ") << t1 - t0 << " transpose " << t2 - t1 << " R*(AP) " << t3 - t2 << ") alloc+copy " << t4 - t3
                  << " sort " << t5 - t4 << " diag " << omp_get_wtime() - t5 << std::endl;
}

// reference src/AMG_cycle_utilities.cpp:149-188: the columns of P follow the coarse matrix' colour permutation,
// P <- P * Pt^T with Pt[new, perm[new]] = 1, i.e. column c becomes inv[c]
void reorder_prolongator(sp_matrix_mg &A, sp_matrix_mg *&P) {
    std::vector<int> inv(A.nrow);
    for (int i = 0; i < A.nrow; i++) inv[A.color[i]] = i;
    for (int j = 0; j < P->rowptr[P->nrow]; j++) P->colindex[j] = inv[P->colindex[j]];
    P->sp_matrix_fill();
}

// reference src/AMG_cycle_utilities.cpp:191-223: b <- Pt b, b_new[i] = b[perm[i]]
void reorder_rhs(sp_matrix_mg &A, double *&b) {
    std::vector<double> t(A.nrow);
    for (int i = 0; i < A.nrow; i++) t[i] = b[A.color[i]];
    std::copy(t.begin(), t.end(), b);
}

}  // namespace parallel

// ---------------------------------------------------------------------------------------------------------
// AMG_solver: hierarchy construction
// ---------------------------------------------------------------------------------------------------------
AMG_solver::AMG_solver() { reserve_levels(std::max(options().max_levels, 1)); }

// The per-level arrays belong to the object, not to the option that happened to be set when it was constructed:
// options().max_levels may be raised between construction and setup (the reference's level1 is a compile-time macro).
void AMG_solver::reserve_levels(int nlevels) {
    if (nlevels <= capacity) return;
    auto grow = [&](auto *&arr) {
        using T = std::remove_reference_t<decltype(arr[0])>;
        T *bigger = new T[nlevels]();
        for (int k = 0; k < capacity; k++) bigger[k] = arr[k];
        delete[] arr;
        arr = bigger;
    };
    grow(Av);
    grow(Pv);
    grow(Xv);
    grow(Bv);
    grow(Rv);
    capacity = nlevels;
}

static void build_hierarchy(AMG_solver &S, sp_matrix_mg &A, bool colour) {
    const sparsh::Options &o = options();
    const double t0 = omp_get_wtime();
    const bool timing = getenv("SPARSH_SETUP_TIMING") != nullptr;  // developer switch: where does the setup go?
    S.l = 0;
    S.reserve_levels(std::max(o.max_levels, 1));
    S.Av[0] = &A;  // borrowed, as in the reference (src/AMG_phases.cpp:40)
    if (o.print_setup) std::cout << "AMG Setup Phase Details " << (colour ? "SOR Smoother" : "Jacobi smoother") << std::endl;
    if (colour) S.Av[0]->color_matrix_and_reorder();
    int l = 0;
    while (S.Av[l]->nrow > o.coarse_upper && l < o.max_levels - 1 && l < S.capacity - 1) {
        if (o.print_setup) std::cout << "Level " << l << ":\t" << S.Av[l]->nrow << std::endl;
        const double tp = omp_get_wtime();
        if (o.coarsening == sparsh::COARSEN_BECK)
            sequential::beck_prolongator(*S.Av[l], S.Pv[l]);
        else if (o.coarsening == sparsh::COARSEN_SA)
            sequential::SA_Prolongator(*S.Av[l], S.Pv[l], l);
        else
            sequential::HEM_Prolongator(*S.Av[l], S.Pv[l], l);
        if (timing) std::cout << "  prolongator (" << S.Av[l]->nrow << " rows): " << omp_get_wtime() - tp << std::endl;
        parallel::coarsen_matrix(*S.Av[l], S.Av[l + 1], *S.Pv[l]);
        if (colour) {
            forget_device_product(S.Av[l + 1]);  // the reordering below makes the device copy stale
            S.Av[l + 1]->color_matrix_and_reorder();
            parallel::reorder_prolongator(*S.Av[l + 1], S.Pv[l]);
        }
        l++;
        if (S.Av[l]->nrow < o.coarse_lower) break;
    }
    if (o.print_setup) std::cout << "Level " << l << ":\t" << S.Av[l]->nrow << std::endl;
    forget_all_device_products();
    S.l = l;
    sparsh::last_report().setup_seconds = omp_get_wtime() - t0;
}

void AMG_solver::AMG_solver_setup_jacobi(sp_matrix_mg &A) { build_hierarchy(*this, A, false); }
void AMG_solver::AMG_solver_setup_SOR(sp_matrix_mg &A) { build_hierarchy(*this, A, true); }
