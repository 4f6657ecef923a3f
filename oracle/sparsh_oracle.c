/*
 * sparsh_oracle.c — CPU restatement of the SParSH-AMG solve phase.
 * TEST INFRASTRUCTURE ONLY (see sparsh_oracle.h for the rules and for how parity is pinned).
 * Plain C + OpenMP.  All file:line citations are relative to /root/reference.
 */
#include "sparsh_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

static int g_threads = 1;

void so_set_threads(int nt) { g_threads = nt > 0 ? nt : 1; }
int so_get_threads(void) { return g_threads; }
void so_free(void *p) { free(p); }

static void *xmalloc(size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p) abort();
    return p;
}
static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) abort();
    return p;
}

/* ------------------------------------------------------------------------------------------------
 * BLAS-1.  The reference calls MKL's cblas_ddot / cblas_dnrm2 (e.g. src/AMG_main_solvers.cpp:140-152);
 * MKL's internal summation order is unknown, so the oracle fixes one that does not depend on the
 * thread count: 1024-element blocks summed left to right, block sums added left to right.
 * ------------------------------------------------------------------------------------------------ */
#define SO_BLK 1024
double so_dot(int n, const double *x, const double *y) {
    int nb = (n + SO_BLK - 1) / SO_BLK;
    double *part = (double *)xmalloc(sizeof(double) * (size_t)nb);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int b = 0; b < nb; b++) {
        int lo = b * SO_BLK, hi = lo + SO_BLK < n ? lo + SO_BLK : n;
        double s = 0.0;
        for (int i = lo; i < hi; i++) s += x[i] * y[i];
        part[b] = s;
    }
    double s = 0.0;
    for (int b = 0; b < nb; b++) s += part[b];
    free(part);
    return s;
}
double so_nrm2(int n, const double *x) { return sqrt(so_dot(n, x, x)); }

/* ------------------------------------------------------------------------------------------------
 * SpMV.  mkl_sparse_d_mv(NON_TRANSPOSE, 1.0, A, des, x, 0.0, y): row sums left to right over the
 * column-sorted row (sp_matrix_fill sorts columns in place, src/AMG_cpu_matrix.cpp:29).
 * ------------------------------------------------------------------------------------------------ */
void so_spmv(int nrow, const int *rp, const int *ci, const double *v, const double *x, double *y) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < nrow; i++) {
        double s = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++) s += v[j] * x[ci[j]];
        y[i] = s;
    }
}

/* mkl_sparse_d_mv(TRANSPOSE, 1.0, P, des, r, 0.0, b): y = 0, then scatter in row order. */
void so_spmv_t(int nrow, int ncol, const int *rp, const int *ci, const double *v, const double *x, double *y) {
    for (int j = 0; j < ncol; j++) y[j] = 0.0;
    for (int i = 0; i < nrow; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++) y[ci[j]] += v[j] * x[i];
}

/* src/AMG_smoothers.cpp:53-76.  `while(count++ <= iteration)` runs iteration+1 sweeps:
 *   :62  helper = A x
 *   :63  helper = 1.0*b + (-1.0)*helper
 *   :71  x[i] += omega*helper[i]/diagonal[i]        evaluated as (omega*h)/d                        */
void so_jacobi(int n, const int *rp, const int *ci, const double *v, const double *diag, const double *b, double *x,
               double *helper, double omega, int iteration) {
    int count = 0;
    while (count++ <= iteration) {
        so_spmv(n, rp, ci, v, x, helper);
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) helper[i] = b[i] - helper[i];
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) x[i] += omega * helper[i] / diag[i];
    }
}

/* src/AMG_cycle_utilities.cpp:83-94: helper = A x; helper += (-1.0)*b; return nrm2(helper). */
double so_residual(int n, const int *rp, const int *ci, const double *v, const double *b, const double *x,
                   double *helper) {
    so_spmv(n, rp, ci, v, x, helper);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < n; i++) helper[i] = helper[i] - b[i];
    return so_nrm2(n, helper);
}

/* src/AMG_cycle_utilities.cpp:115-123: r = A x; r = 1.0*b + (-1.0)*r. */
void so_store_residual(int n, const int *rp, const int *ci, const double *v, const double *b, const double *x,
                       double *r) {
    so_spmv(n, rp, ci, v, x, r);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < n; i++) r[i] = b[i] - r[i];
}

/* src/AMG_cycle_utilities.cpp:97-104 */
void so_transfer_residual(int nf, int nc, const int *prp, const int *pci, const double *pv, const double *r,
                          double *bc) {
    so_spmv_t(nf, nc, prp, pci, pv, r, bc);
}

/* src/AMG_cycle_utilities.cpp:107-112: x1 = 1.0*(P x) + 1.0*x1 */
void so_transfer_solution(int nf, const int *prp, const int *pci, const double *pv, const double *xc, double *xf) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < nf; i++) {
        double s = 0.0;
        for (int j = prp[i]; j < prp[i + 1]; j++) s += pv[j] * xc[pci[j]];
        xf[i] = s + xf[i];
    }
}

/* src/AMG_smoothers.cpp:78-102: `while(count++ < iteration)` = `iteration` sweeps; per colour k the rows
 * color_count[k]..color_count[k+1] are independent:
 *   :90-95 helper = sum a_lj x_j (from 0.0, left to right)   :97 helper -= b   :98 x -= omega*helper/diag  */
void so_sor_multicolor(int n, const int *rp, const int *ci, const double *v, const double *diag,
                       const int *color_count, int total_colors, const double *b, double *x, double *helper,
                       double omega, int iteration) {
    (void)n;
    int count = 0;
    while (count++ < iteration) {
        for (int k = 0; k < total_colors; k++) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
            for (int l = color_count[k]; l < color_count[k + 1]; l++) {
                double h = 0.0;
                for (int lj = rp[l]; lj < rp[l + 1]; lj++) h += v[lj] * x[ci[lj]];
                h -= b[l];
                helper[l] = h;
                x[l] -= omega * h / diag[l];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Setup restatement
 * ------------------------------------------------------------------------------------------------ */
/* src/AMG_cpu_matrix.cpp:35-51: first stored entry with column == row.  (The reference leaves the slot
 * uninitialised when a row stores no diagonal; the oracle writes 0.0 there.) */
void so_fill_diagonal(int n, const int *rp, const int *ci, const double *v, double *diag) {
    for (int i = 0; i < n; i++) {
        diag[i] = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (ci[j] == i) {
                diag[i] = v[j];
                break;
            }
    }
}

static void sort_row(int len, int *c, double *v) {
    for (int a = 1; a < len; a++) { /* insertion sort: rows are short */
        int cc = c[a];
        double vv = v[a];
        int b = a - 1;
        while (b >= 0 && c[b] > cc) {
            c[b + 1] = c[b];
            v[b + 1] = v[b];
            b--;
        }
        c[b + 1] = cc;
        v[b + 1] = vv;
    }
}
static void sort_columns(int nrow, const int *rp, int *ci, double *v) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1024)
    for (int i = 0; i < nrow; i++) sort_row(rp[i + 1] - rp[i], ci + rp[i], v + rp[i]);
}

/* src/AMG_coarsening.cpp:14-97.  Sweep forward on even levels, backward on odd ones (:25,:56); an
 * unaggregated row pairs with its unaggregated neighbour of strictly largest |a_ij|, j != i (:38,:67);
 * leftovers become singletons numbered last in ascending row order (:84-91). */
int so_hem(int n, const int *rp, const int *ci, const double *v, int level, int *agg) {
    int newnum = 0;
    for (int i = 0; i < n; i++) agg[i] = -1;
    int start = (level % 2 == 0) ? 0 : n - 1;
    int step = (level % 2 == 0) ? 1 : -1;
    for (int t = 0, i = start; t < n; t++, i += step) {
        if (agg[i] != -1) continue;
        int id = -1;
        double max1 = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            if (agg[ci[j]] == -1 && fabs(v[j]) > max1 && ci[j] != i) {
                max1 = fabs(v[j]);
                id = ci[j];
            }
        }
        if (id != -1) {
            agg[i] = newnum;
            agg[id] = newnum;
            newnum++;
        }
    }
    for (int i = 0; i < n; i++)
        if (agg[i] == -1) agg[i] = newnum++;
    return newnum;
}

/* src/AMG_coarsening.cpp:269-339.  First-fit C-point selection in natural order (:274-285); C rows are
 * identity rows (:293-299); an F row interpolates from every C neighbour with weight 1/|c_f[i]| (:301-315);
 * sp_matrix_fill() then sorts the columns of each row (:328 -> src/AMG_cpu_matrix.cpp:29). */
void so_beck(int n, const int *rp, const int *ci, int *nc, int **prp_o, int **pci_o, double **pv_o) {
    int *c_f = (int *)xcalloc((size_t)n, sizeof(int));
    int c_count = 0;
    for (int i = 0; i < n; i++) {
        if (c_f[i] == 0) {
            for (int j = rp[i]; j < rp[i + 1]; j++) c_f[ci[j]] -= 1;
            c_f[i] = c_count + 1;
            c_count++;
        }
    }
    int *prp = (int *)xcalloc((size_t)n + 1, sizeof(int));
    for (int i = 0; i < n; i++) {
        int cnt = 0;
        if (c_f[i] > 0)
            cnt = 1;
        else if (c_f[i] < 0)
            for (int j = rp[i]; j < rp[i + 1]; j++)
                if (c_f[ci[j]] > 0) cnt++;
        prp[i + 1] = prp[i] + cnt;
    }
    int pnnz = prp[n];
    int *pci = (int *)xmalloc(sizeof(int) * (size_t)pnnz);
    double *pv = (double *)xmalloc(sizeof(double) * (size_t)pnnz);
    for (int i = 0; i < n; i++) {
        int o = prp[i];
        if (c_f[i] > 0) {
            pci[o] = c_f[i] - 1;
            pv[o] = 1.0;
        } else if (c_f[i] < 0) {
            double p1 = 1 / fabs((double)c_f[i]);
            for (int j = rp[i]; j < rp[i + 1]; j++) {
                int k = ci[j];
                if (c_f[k] > 0) {
                    pci[o] = c_f[k] - 1;
                    pv[o] = p1;
                    o++;
                }
            }
        }
    }
    sort_columns(n, prp, pci, pv);
    free(c_f);
    *nc = c_count;
    *prp_o = prp;
    *pci_o = pci;
    *pv_o = pv;
}

/* stable transpose of an nrow x ncol CSR (rows of the result list source rows in ascending order) */
static void csr_transpose(int nrow, int ncol, const int *rp, const int *ci, const double *v, int **trp_o,
                          int **tci_o, double **tv_o) {
    int nnz = rp[nrow];
    int *trp = (int *)xcalloc((size_t)ncol + 1, sizeof(int));
    int *tci = (int *)xmalloc(sizeof(int) * (size_t)nnz);
    double *tv = (double *)xmalloc(sizeof(double) * (size_t)nnz);
    for (int j = 0; j < nnz; j++) trp[ci[j] + 1]++;
    for (int c = 0; c < ncol; c++) trp[c + 1] += trp[c];
    int *cur = (int *)xmalloc(sizeof(int) * ((size_t)ncol + 1));
    memcpy(cur, trp, sizeof(int) * ((size_t)ncol + 1));
    for (int i = 0; i < nrow; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            int d = cur[ci[j]]++;
            tci[d] = i;
            tv[d] = v[j];
        }
    free(cur);
    *trp_o = trp;
    *tci_o = tci;
    *tv_o = tv;
}

/* Gustavson row-wise C = A B; entries of a C row appear in first-touch order, each value accumulated in
 * traversal order (k ascending over A's row, then B's row left to right). */
static void csr_spgemm(int arow, const int *arp, const int *aci, const double *av, int bcol, const int *brp,
                       const int *bci, const double *bv, int **crp_o, int **cci_o, double **cv_o) {
    int *crp = (int *)xcalloc((size_t)arow + 1, sizeof(int));
#pragma omp parallel num_threads(g_threads)
    {
        int *mark = (int *)xmalloc(sizeof(int) * (size_t)(bcol > 0 ? bcol : 1));
        for (int c = 0; c < bcol; c++) mark[c] = -1;
#pragma omp for schedule(dynamic, 2048)
        for (int i = 0; i < arow; i++) {
            int cnt = 0;
            for (int ja = arp[i]; ja < arp[i + 1]; ja++) {
                int k = aci[ja];
                for (int jb = brp[k]; jb < brp[k + 1]; jb++) {
                    int c = bci[jb];
                    if (mark[c] != i) {
                        mark[c] = i;
                        cnt++;
                    }
                }
            }
            crp[i + 1] = cnt;
        }
        free(mark);
    }
    for (int i = 0; i < arow; i++) crp[i + 1] += crp[i];
    int cnnz = crp[arow];
    int *cci = (int *)xmalloc(sizeof(int) * (size_t)cnnz);
    double *cv = (double *)xmalloc(sizeof(double) * (size_t)cnnz);
#pragma omp parallel num_threads(g_threads)
    {
        int *pos = (int *)xmalloc(sizeof(int) * (size_t)(bcol > 0 ? bcol : 1));
        for (int c = 0; c < bcol; c++) pos[c] = -1;
#pragma omp for schedule(dynamic, 2048)
        for (int i = 0; i < arow; i++) {
            int base = crp[i], o = base;
            for (int ja = arp[i]; ja < arp[i + 1]; ja++) {
                int k = aci[ja];
                double a = av[ja];
                for (int jb = brp[k]; jb < brp[k + 1]; jb++) {
                    int c = bci[jb];
                    if (pos[c] < base) {
                        pos[c] = o;
                        cci[o] = c;
                        cv[o] = a * bv[jb];
                        o++;
                    } else {
                        cv[pos[c]] += a * bv[jb];
                    }
                }
            }
        }
        free(pos);
    }
    *crp_o = crp;
    *cci_o = cci;
    *cv_o = cv;
}

/* src/AMG_cycle_utilities.cpp:126-146: C1 = A P (:134); Ac = P^T C1 (:135); mkl_sparse_order (:137). */
void so_rap(int n, const int *rp, const int *ci, const double *v, int nc, const int *prp, const int *pci,
            const double *pv, int **crp, int **cci, double **cv) {
    int *t_rp, *t_ci, *r_rp, *r_ci;
    double *t_v, *r_v;
    csr_spgemm(n, rp, ci, v, nc, prp, pci, pv, &t_rp, &t_ci, &t_v);
    csr_transpose(n, nc, prp, pci, pv, &r_rp, &r_ci, &r_v);
    csr_spgemm(nc, r_rp, r_ci, r_v, nc, t_rp, t_ci, t_v, crp, cci, cv);
    sort_columns(nc, *crp, *cci, *cv);
    free(t_rp);
    free(t_ci);
    free(t_v);
    free(r_rp);
    free(r_ci);
    free(r_v);
}

/* src/AMG_cpu_matrix.cpp:81-199.  Greedy first-fit colouring in natural order over the stored columns
 * (:102-127); colours start at 1; `color` is then reused as perm[new] = old, grouped by colour, ascending
 * old index inside a colour (:129-141); color_count becomes prefix offsets.  The matrix is permuted
 * symmetrically, B = Pt A Pt^T with Pt[new, perm[new]] = 1 (:161-187), columns sorted. */
int so_color_reorder(int n, const int *rp, const int *ci, const double *v, int *perm, int *color_count,
                     int **qrp_o, int **qci_o, double **qv_o) {
    int max_count = 0;
    for (int i = 0; i < n; i++)
        if (rp[i + 1] - rp[i] > max_count) max_count = rp[i + 1] - rp[i];
    int *color = (int *)xcalloc((size_t)n, sizeof(int));
    int *forbidden = (int *)xmalloc(sizeof(int) * ((size_t)max_count + 1));
    int *cc = (int *)xcalloc((size_t)max_count + 1, sizeof(int));
    int total_colors = 0;
    for (int k = 0; k < max_count + 1; k++) forbidden[k] = -1;
    for (int i = 0; i < n; i++) {
        /* the reference refills forbidden[] with -1 each row; marking with the row index is equivalent */
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (color[ci[j]] != 0) forbidden[color[ci[j]]] = i;
        int c = 0x7fffffff;
        for (int k = 1; k < max_count + 1; k++)
            if (forbidden[k] != i) {
                c = k;
                break;
            }
        color[i] = c;
        cc[c]++;
        if (c > total_colors) total_colors = c;
    }
    /* prefix offsets, then group rows by colour */
    for (int k = 0; k < total_colors; k++) cc[k + 1] += cc[k];
    int *cur = (int *)xmalloc(sizeof(int) * ((size_t)total_colors + 1));
    for (int k = 0; k <= total_colors; k++) cur[k] = k == 0 ? 0 : cc[k - 1];
    for (int i = 0; i < n; i++) perm[cur[color[i]]++] = i;
    for (int k = 0; k <= total_colors; k++) color_count[k] = cc[k];
    free(cur);
    free(forbidden);
    free(cc);
    free(color);

    int *inv = (int *)xmalloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) inv[perm[i]] = i;
    int *qrp = (int *)xmalloc(sizeof(int) * ((size_t)n + 1));
    qrp[0] = 0;
    for (int i = 0; i < n; i++) qrp[i + 1] = qrp[i] + (rp[perm[i] + 1] - rp[perm[i]]);
    int *qci = (int *)xmalloc(sizeof(int) * (size_t)qrp[n]);
    double *qv = (double *)xmalloc(sizeof(double) * (size_t)qrp[n]);
    for (int i = 0; i < n; i++) {
        int o = qrp[i];
        for (int j = rp[perm[i]]; j < rp[perm[i] + 1]; j++, o++) {
            qci[o] = inv[ci[j]];
            qv[o] = v[j];
        }
    }
    sort_columns(n, qrp, qci, qv);
    free(inv);
    *qrp_o = qrp;
    *qci_o = qci;
    *qv_o = qv;
    return total_colors;
}

/* ------------------------------------------------------------------------------------------------
 * Coarse direct solve.  The reference factors with PARDISO (mtype 11, src/AMG_coarse_level_solver.cpp:
 * 23-52) and solves with phase 33 (:66-70).  PARDISO is closed source; the oracle uses the textbook
 * equivalent: reverse Cuthill-McKee ordering + banded LU with partial pivoting (LAPACK dgbtf2/dgbtrs
 * algorithm).  Both are backward-stable direct solves; results agree to O(cond * eps).
 * ------------------------------------------------------------------------------------------------ */
struct so_lu {
    int n, kl, ku, ldab;
    int *perm;   /* perm[new] = old */
    int *ipiv;
    double *ab;  /* column-major band, ldab x n */
};

static void rcm_order(int n, const int *rp, const int *ci, int *perm) {
    /* symmetrised adjacency */
    int *deg = (int *)xcalloc((size_t)n + 1, sizeof(int));
    for (int i = 0; i < n; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (ci[j] != i) {
                deg[i + 1]++;
                deg[ci[j] + 1]++;
            }
    for (int i = 0; i < n; i++) deg[i + 1] += deg[i];
    int *adj = (int *)xmalloc(sizeof(int) * (size_t)(deg[n] > 0 ? deg[n] : 1));
    int *cur = (int *)xmalloc(sizeof(int) * ((size_t)n + 1));
    memcpy(cur, deg, sizeof(int) * ((size_t)n + 1));
    for (int i = 0; i < n; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (ci[j] != i) {
                adj[cur[i]++] = ci[j];
                adj[cur[ci[j]]++] = i;
            }
    free(cur);
    char *seen = (char *)xcalloc((size_t)n, 1);
    int *order = (int *)xmalloc(sizeof(int) * (size_t)n);
    int *lvl = (int *)xmalloc(sizeof(int) * (size_t)n);
    int filled = 0;
    for (int s0 = 0; s0 < n; s0++) {
        if (seen[s0]) continue;
        /* pseudo-peripheral start: repeat BFS from the last-level minimum-degree node a few times */
        int start = s0;
        for (int rep = 0; rep < 4; rep++) {
            int head = 0, tail = 0;
            int *q = order + filled;
            q[tail++] = start;
            lvl[start] = 0;
            seen[start] = 2;
            while (head < tail) {
                int u = q[head++];
                for (int e = deg[u]; e < deg[u + 1]; e++) {
                    int w = adj[e];
                    if (!seen[w]) {
                        seen[w] = 2;
                        lvl[w] = lvl[u] + 1;
                        q[tail++] = w;
                    }
                }
            }
            int last = q[tail - 1], maxl = lvl[last], best = last;
            for (int t = tail - 1; t >= 0 && lvl[q[t]] == maxl; t--)
                if (deg[q[t] + 1] - deg[q[t]] < deg[best + 1] - deg[best]) best = q[t];
            for (int t = 0; t < tail; t++) seen[q[t]] = 0;
            if (best == start) break;
            start = best;
        }
        /* Cuthill-McKee BFS with neighbours in ascending degree */
        int head = filled, tail = filled;
        order[tail++] = start;
        seen[start] = 1;
        while (head < tail) {
            int u = order[head++];
            int first = tail;
            for (int e = deg[u]; e < deg[u + 1]; e++) {
                int w = adj[e];
                if (!seen[w]) {
                    seen[w] = 1;
                    order[tail++] = w;
                }
            }
            for (int a = first + 1; a < tail; a++) { /* insertion sort by degree, stable */
                int w = order[a], dw = deg[w + 1] - deg[w], b = a - 1;
                while (b >= first && deg[order[b] + 1] - deg[order[b]] > dw) {
                    order[b + 1] = order[b];
                    b--;
                }
                order[b + 1] = w;
            }
        }
        filled = tail;
    }
    for (int i = 0; i < n; i++) perm[i] = order[n - 1 - i]; /* reverse */
    free(order);
    free(lvl);
    free(seen);
    free(adj);
    free(deg);
}

so_lu *so_lu_factor(int n, const int *rp, const int *ci, const double *v) {
    so_lu *f = (so_lu *)xcalloc(1, sizeof(so_lu));
    f->n = n;
    f->perm = (int *)xmalloc(sizeof(int) * (size_t)n);
    f->ipiv = (int *)xmalloc(sizeof(int) * (size_t)n);
    rcm_order(n, rp, ci, f->perm);
    int *inv = (int *)xmalloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) inv[f->perm[i]] = i;
    int kl = 0, ku = 0;
    for (int i = 0; i < n; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            int r = inv[i], c = inv[ci[j]];
            if (r - c > kl) kl = r - c;
            if (c - r > ku) ku = c - r;
        }
    f->kl = kl;
    f->ku = ku;
    int kv = kl + ku, ldab = 2 * kl + ku + 1;
    f->ldab = ldab;
    double *ab = (double *)xcalloc((size_t)ldab * (size_t)n, sizeof(double));
    f->ab = ab;
#define AB(i, j) ab[(size_t)(j) * ldab + (kv + (i) - (j))]
    for (int i = 0; i < n; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++) AB(inv[i], inv[ci[j]]) += v[j];
    free(inv);
    int ju = 0;
    for (int j = 0; j < n; j++) {
        int km = kl < n - 1 - j ? kl : n - 1 - j;
        int jp = 0;
        double best = fabs(AB(j, j));
        for (int p = 1; p <= km; p++)
            if (fabs(AB(j + p, j)) > best) {
                best = fabs(AB(j + p, j));
                jp = p;
            }
        f->ipiv[j] = j + jp;
        if (AB(j + jp, j) != 0.0) {
            int cand = j + ku + jp < n - 1 ? j + ku + jp : n - 1;
            if (cand > ju) ju = cand;
            if (jp != 0)
                for (int c = j; c <= ju; c++) {
                    double t = AB(j + jp, c);
                    AB(j + jp, c) = AB(j, c);
                    AB(j, c) = t;
                }
            double piv = AB(j, j);
            for (int p = 1; p <= km; p++) AB(j + p, j) /= piv;
#pragma omp parallel for num_threads(g_threads) schedule(static) if (ju - j > 64)
            for (int c = j + 1; c <= ju; c++) {
                double u = AB(j, c);
                if (u != 0.0)
                    for (int p = 1; p <= km; p++) AB(j + p, c) -= AB(j + p, j) * u;
            }
        }
    }
    return f;
}

void so_lu_solve(const so_lu *f, const double *b, double *x) {
    int n = f->n, kl = f->kl, ku = f->ku, ldab = f->ldab, kv = kl + ku;
    const double *ab = f->ab;
    double *y = (double *)xmalloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++) y[i] = b[f->perm[i]];
    for (int j = 0; j < n - 1; j++) {
        int lm = kl < n - 1 - j ? kl : n - 1 - j;
        int l = f->ipiv[j];
        if (l != j) {
            double t = y[l];
            y[l] = y[j];
            y[j] = t;
        }
        double yj = y[j];
        if (yj != 0.0)
            for (int p = 1; p <= lm; p++) y[j + p] -= AB(j + p, j) * yj;
    }
    for (int j = n - 1; j >= 0; j--) {
        y[j] /= AB(j, j);
        double yj = y[j];
        int lo = j - kv > 0 ? j - kv : 0;
        for (int i = lo; i < j; i++) y[i] -= AB(i, j) * yj;
    }
    for (int i = 0; i < n; i++) x[f->perm[i]] = y[i];
    free(y);
#undef AB
}

void so_lu_free(so_lu *f) {
    if (!f) return;
    free(f->perm);
    free(f->ipiv);
    free(f->ab);
    free(f);
}
int so_lu_bandwidth(const so_lu *f) { return f->kl > f->ku ? f->kl : f->ku; }

/* ------------------------------------------------------------------------------------------------
 * Hierarchy (class AMG_solver, include/AMG_phases.hpp:8-53)
 * ------------------------------------------------------------------------------------------------ */
#define SO_MAXLEV 64
typedef struct {
    int n, nnz;
    int *rp, *ci;
    double *v, *diag, *helper;
    int pncol, pnnz;
    int *prp, *pci;
    double *pv;
    double *X, *B, *R;
    int total_colors;  /* SOR hierarchies only */
    int *color_count, *perm;
} so_level;

struct so_amg {
    int l; /* index of the coarsest level, as AMG_solver::l */
    so_level lev[SO_MAXLEV];
    so_lu *lu;
    double omega;
    int smooth_iter;
};

static void level_alloc_vectors(so_level *L) {
    L->X = (double *)xcalloc((size_t)L->n, sizeof(double));
    L->B = (double *)xcalloc((size_t)L->n, sizeof(double));
    L->R = (double *)xcalloc((size_t)L->n, sizeof(double));
    L->helper = (double *)xcalloc((size_t)L->n, sizeof(double));
}

static void level_set_matrix(so_level *L, int n, const int *rp, const int *ci, const double *v, int copy) {
    L->n = n;
    L->nnz = rp[n];
    if (copy) {
        L->rp = (int *)xmalloc(sizeof(int) * ((size_t)n + 1));
        L->ci = (int *)xmalloc(sizeof(int) * (size_t)L->nnz);
        L->v = (double *)xmalloc(sizeof(double) * (size_t)L->nnz);
        memcpy(L->rp, rp, sizeof(int) * ((size_t)n + 1));
        memcpy(L->ci, ci, sizeof(int) * (size_t)L->nnz);
        memcpy(L->v, v, sizeof(double) * (size_t)L->nnz);
    } else {
        L->rp = (int *)rp;
        L->ci = (int *)ci;
        L->v = (double *)v;
    }
    L->diag = (double *)xmalloc(sizeof(double) * (size_t)n);
    so_fill_diagonal(n, L->rp, L->ci, L->v, L->diag);
    level_alloc_vectors(L);
}

/* src/AMG_phases.cpp:35-90 */
so_amg *so_amg_setup(int n, const int *rp, const int *ci, const double *v, int coarsening, int max_levels,
                     int limit_upper, int limit_lower) {
    so_amg *h = (so_amg *)xcalloc(1, sizeof(so_amg));
    h->omega = 0.66667; /* include/AMG.hpp:16 */
    h->smooth_iter = 6; /* include/AMG.hpp:22 */
    if (max_levels > SO_MAXLEV) max_levels = SO_MAXLEV;
    int l = 0;
    level_set_matrix(&h->lev[0], n, rp, ci, v, 1);
    /* sp_matrix_fill(): columns sorted in place (src/AMG_cpu_matrix.cpp:29) */
    sort_columns(n, h->lev[0].rp, h->lev[0].ci, h->lev[0].v);
    so_fill_diagonal(n, h->lev[0].rp, h->lev[0].ci, h->lev[0].v, h->lev[0].diag);
    while (h->lev[l].n > limit_upper && l < max_levels - 1) { /* :51 */
        so_level *F = &h->lev[l];
        if (coarsening == 1) {
            so_beck(F->n, F->rp, F->ci, &F->pncol, &F->prp, &F->pci, &F->pv);
        } else {
            int *agg = (int *)xmalloc(sizeof(int) * (size_t)F->n);
            F->pncol = so_hem(F->n, F->rp, F->ci, F->v, l, agg); /* :61 */
            F->prp = (int *)xmalloc(sizeof(int) * ((size_t)F->n + 1));
            F->pv = (double *)xmalloc(sizeof(double) * (size_t)F->n);
            for (int i = 0; i <= F->n; i++) F->prp[i] = i;
            for (int i = 0; i < F->n; i++) F->pv[i] = 1.0;
            F->pci = agg;
        }
        F->pnnz = F->prp[F->n];
        int *crp, *cci;
        double *cv;
        so_rap(F->n, F->rp, F->ci, F->v, F->pncol, F->prp, F->pci, F->pv, &crp, &cci, &cv); /* :68 */
        l = l + 1;
        level_set_matrix(&h->lev[l], F->pncol, crp, cci, cv, 0);
        if (h->lev[l].n < limit_lower) break; /* :77 */
    }
    h->l = l;
    h->lu = so_lu_factor(h->lev[l].n, h->lev[l].rp, h->lev[l].ci, h->lev[l].v); /* :89 */
    return h;
}

/* colour-permute level L in place (src/AMG_cpu_matrix.cpp:81-199), keeping perm and the colour offsets */
static void level_color(so_level *L) {
    int maxdeg = 0;
    for (int i = 0; i < L->n; i++)
        if (L->rp[i + 1] - L->rp[i] > maxdeg) maxdeg = L->rp[i + 1] - L->rp[i];
    L->perm = (int *)xmalloc(sizeof(int) * (size_t)L->n);
    L->color_count = (int *)xcalloc((size_t)maxdeg + 2, sizeof(int));
    int *qrp, *qci;
    double *qv;
    L->total_colors = so_color_reorder(L->n, L->rp, L->ci, L->v, L->perm, L->color_count, &qrp, &qci, &qv);
    free(L->rp);
    free(L->ci);
    free(L->v);
    L->rp = qrp;
    L->ci = qci;
    L->v = qv;
    so_fill_diagonal(L->n, L->rp, L->ci, L->v, L->diag);
}

/* src/AMG_phases.cpp:94-147 */
so_amg *so_amg_setup_sor(int n, const int *rp, const int *ci, const double *v, int max_levels, int limit_upper,
                         int limit_lower) {
    so_amg *h = (so_amg *)xcalloc(1, sizeof(so_amg));
    h->omega = 0.66667;
    h->smooth_iter = 6;
    if (max_levels > SO_MAXLEV) max_levels = SO_MAXLEV;
    int l = 0;
    level_set_matrix(&h->lev[0], n, rp, ci, v, 1);
    sort_columns(n, h->lev[0].rp, h->lev[0].ci, h->lev[0].v);
    level_color(&h->lev[0]); /* :108 */
    while (h->lev[l].n > limit_upper && l < max_levels - 1) { /* :109 */
        so_level *F = &h->lev[l];
        int *agg = (int *)xmalloc(sizeof(int) * (size_t)F->n);
        F->pncol = so_hem(F->n, F->rp, F->ci, F->v, l, agg); /* :118, on the permuted matrix */
        F->prp = (int *)xmalloc(sizeof(int) * ((size_t)F->n + 1));
        F->pv = (double *)xmalloc(sizeof(double) * (size_t)F->n);
        for (int i = 0; i <= F->n; i++) F->prp[i] = i;
        for (int i = 0; i < F->n; i++) F->pv[i] = 1.0;
        F->pci = agg;
        F->pnnz = F->n;
        int *crp, *cci;
        double *cv;
        so_rap(F->n, F->rp, F->ci, F->v, F->pncol, F->prp, F->pci, F->pv, &crp, &cci, &cv); /* :124 */
        l = l + 1;
        level_set_matrix(&h->lev[l], F->pncol, crp, cci, cv, 0);
        level_color(&h->lev[l]); /* :125-126 */
        /* reorder_prolongator (:128): column c of P becomes inv[c] of the coarse permutation */
        {
            so_level *C = &h->lev[l];
            int *inv = (int *)xmalloc(sizeof(int) * (size_t)C->n);
            for (int i = 0; i < C->n; i++) inv[C->perm[i]] = i;
            for (int j = 0; j < F->pnnz; j++) F->pci[j] = inv[F->pci[j]];
            free(inv);
            sort_columns(F->n, F->prp, F->pci, F->pv);
        }
        if (h->lev[l].n < limit_lower) break; /* :135 */
    }
    h->l = l;
    h->lu = so_lu_factor(h->lev[l].n, h->lev[l].rp, h->lev[l].ci, h->lev[l].v); /* :146 */
    return h;
}

static double level0_residual(so_amg *h);

/* src/AMG_phases.cpp:279-297, one cycle */
static void one_cycle_sor(so_amg *h) {
    int l = h->l;
    for (int l1 = 0; l1 < l; l1++) {
        so_level *F = &h->lev[l1], *C = &h->lev[l1 + 1];
        so_sor_multicolor(F->n, F->rp, F->ci, F->v, F->diag, F->color_count, F->total_colors, F->B, F->X, F->helper,
                          h->omega, 6);                                                          /* :281 */
        so_store_residual(F->n, F->rp, F->ci, F->v, F->B, F->X, F->R);                           /* :282 */
        so_transfer_residual(F->n, F->pncol, F->prp, F->pci, F->pv, F->R, C->B);                 /* :283 */
        memset(C->X, 0, sizeof(double) * (size_t)C->n);                                          /* :284 */
    }
    so_lu_solve(h->lu, h->lev[l].B, h->lev[l].X); /* :289 */
    for (int l1 = l; l1 > 0; l1--) {
        so_level *F = &h->lev[l1 - 1], *C = &h->lev[l1];
        so_transfer_solution(F->n, F->prp, F->pci, F->pv, C->X, F->X);                           /* :294 */
        so_sor_multicolor(F->n, F->rp, F->ci, F->v, F->diag, F->color_count, F->total_colors, F->B, F->X, F->helper,
                          h->omega, 6);                                                          /* :295 */
    }
}

int so_amg_solve_sor(so_amg *h, const double *b, double *x, double tol, int max_cycles, double *hist) {
    so_level *L = &h->lev[0];
    for (int i = 0; i < L->n; i++) { /* reorder_rhs: b <- b[perm] (src/AMG_main_solvers.cpp:35) */
        L->B[i] = b[L->perm[i]];
        L->X[i] = x[L->perm[i]];
    }
    double r1 = level0_residual(h); /* :242 */
    int cycles = 0;
    if (hist) hist[0] = r1;
    while (r1 > tol && cycles < max_cycles) { /* :277 */
        one_cycle_sor(h);
        cycles++;
        r1 = level0_residual(h); /* :300 */
        if (hist) hist[cycles] = r1;
    }
    for (int i = 0; i < L->n; i++) x[L->perm[i]] = L->X[i];
    return cycles;
}

so_amg *so_amg_from_levels(int nlevels, const int *nrow, const int *const *rp, const int *const *ci,
                           const double *const *v, const int *pncol, const int *const *prp, const int *const *pci,
                           const double *const *pv) {
    so_amg *h = (so_amg *)xcalloc(1, sizeof(so_amg));
    h->omega = 0.66667;
    h->smooth_iter = 6;
    for (int k = 0; k < nlevels; k++) {
        level_set_matrix(&h->lev[k], nrow[k], rp[k], ci[k], v[k], 1);
        if (k < nlevels - 1) {
            so_level *F = &h->lev[k];
            int pn = prp[k][nrow[k]];
            F->pncol = pncol[k];
            F->pnnz = pn;
            F->prp = (int *)xmalloc(sizeof(int) * ((size_t)nrow[k] + 1));
            F->pci = (int *)xmalloc(sizeof(int) * (size_t)pn);
            F->pv = (double *)xmalloc(sizeof(double) * (size_t)pn);
            memcpy(F->prp, prp[k], sizeof(int) * ((size_t)nrow[k] + 1));
            memcpy(F->pci, pci[k], sizeof(int) * (size_t)pn);
            memcpy(F->pv, pv[k], sizeof(double) * (size_t)pn);
        }
    }
    h->l = nlevels - 1;
    h->lu = so_lu_factor(h->lev[h->l].n, h->lev[h->l].rp, h->lev[h->l].ci, h->lev[h->l].v);
    return h;
}

void so_amg_free(so_amg *h) {
    if (!h) return;
    for (int k = 0; k <= h->l; k++) {
        so_level *L = &h->lev[k];
        free(L->rp);
        free(L->ci);
        free(L->v);
        free(L->diag);
        free(L->helper);
        free(L->prp);
        free(L->pci);
        free(L->pv);
        free(L->X);
        free(L->B);
        free(L->R);
        free(L->color_count);
        free(L->perm);
    }
    so_lu_free(h->lu);
    free(h);
}

int so_amg_nlevels(const so_amg *h) { return h->l + 1; }
void so_amg_level_dims(const so_amg *h, int lvl, int *nrow, int *nnz, int *p_ncol, int *p_nnz) {
    const so_level *L = &h->lev[lvl];
    *nrow = L->n;
    *nnz = L->nnz;
    *p_ncol = lvl < h->l ? L->pncol : 0;
    *p_nnz = lvl < h->l ? L->pnnz : 0;
}
void so_amg_level_get(const so_amg *h, int lvl, const int **rp, const int **ci, const double **v,
                      const double **diag, const int **prp, const int **pci, const double **pv) {
    const so_level *L = &h->lev[lvl];
    *rp = L->rp;
    *ci = L->ci;
    *v = L->v;
    *diag = L->diag;
    *prp = L->prp;
    *pci = L->pci;
    *pv = L->pv;
}
void so_amg_set_smoother(so_amg *h, double omega, int smooth_iter) {
    h->omega = omega;
    h->smooth_iter = smooth_iter;
}
void so_amg_coarse_solve(so_amg *h, const double *b, double *x) { so_lu_solve(h->lu, b, x); }

static double level0_residual(so_amg *h) {
    so_level *L = &h->lev[0];
    return so_residual(L->n, L->rp, L->ci, L->v, L->B, L->X, L->helper);
}

/* one V-cycle on Xv/Bv, src/AMG_phases.cpp:167-186 (== :198-216) */
static void one_cycle(so_amg *h) {
    int l = h->l;
    for (int l1 = 0; l1 < l; l1++) {
        so_level *F = &h->lev[l1], *C = &h->lev[l1 + 1];
        so_jacobi(F->n, F->rp, F->ci, F->v, F->diag, F->B, F->X, F->helper, h->omega, h->smooth_iter); /* :169 */
        so_store_residual(F->n, F->rp, F->ci, F->v, F->B, F->X, F->R);                                 /* :170 */
        so_transfer_residual(F->n, F->pncol, F->prp, F->pci, F->pv, F->R, C->B);                       /* :171 */
        memset(C->X, 0, sizeof(double) * (size_t)C->n);                                                /* :172 */
    }
    so_lu_solve(h->lu, h->lev[l].B, h->lev[l].X); /* :177 */
    for (int l1 = l; l1 > 0; l1--) {
        so_level *F = &h->lev[l1 - 1], *C = &h->lev[l1];
        so_transfer_solution(F->n, F->prp, F->pci, F->pv, C->X, F->X);                                 /* :183 */
        so_jacobi(F->n, F->rp, F->ci, F->v, F->diag, F->B, F->X, F->helper, h->omega, h->smooth_iter); /* :184 */
    }
}

void so_amg_vcycle(so_amg *h, const double *b, double *x, int cycles) {
    so_level *L = &h->lev[0];
    memcpy(L->B, b, sizeof(double) * (size_t)L->n); /* :156 */
    memcpy(L->X, x, sizeof(double) * (size_t)L->n); /* :157 */
    for (int c = 0; c < cycles; c++) one_cycle(h);
    memcpy(x, L->X, sizeof(double) * (size_t)L->n); /* :228 */
}

int so_amg_solve(so_amg *h, const double *b, double *x, double tol, int max_cycles, double *hist) {
    so_level *L = &h->lev[0];
    memcpy(L->B, b, sizeof(double) * (size_t)L->n);
    memcpy(L->X, x, sizeof(double) * (size_t)L->n);
    double r1 = level0_residual(h); /* :159 */
    int cycles = 0;
    if (hist) hist[0] = r1;
    while (r1 > tol && cycles < max_cycles) { /* :196 (the reference has no cap) */
        one_cycle(h);
        cycles++;
        r1 = level0_residual(h); /* :219 */
        if (hist) hist[cycles] = r1;
    }
    memcpy(x, L->X, sizeof(double) * (size_t)L->n);
    return cycles;
}

/* src/AMG_main_solvers.cpp:107-167.  z0 is zero-initialised here; the reference passes it uninitialised to
 * the first preconditioner call (:112,:132). */
int so_pcg(so_amg *h, const double *b, double *x, double tol, int max_iter, double *hist) {
    so_level *A = &h->lev[0];
    int n = A->n;
    double *Ap = (double *)xmalloc(sizeof(double) * (size_t)n);
    double *p = (double *)xmalloc(sizeof(double) * (size_t)n);
    double *z0 = (double *)xcalloc((size_t)n, sizeof(double));
    double *r0 = (double *)xmalloc(sizeof(double) * (size_t)n);
    so_store_residual(n, A->rp, A->ci, A->v, b, x, r0); /* :124-125 */
    double r1 = so_nrm2(n, r0);                        /* :127 */
    if (hist) hist[0] = r1;
    so_amg_vcycle(h, r0, z0, 1);                       /* :132 */
    memcpy(p, z0, sizeof(double) * (size_t)n);         /* :134 */
    int count = 0;
    while (count < n && count < max_iter && r1 > tol) { /* :136 */
        count++;
        so_spmv(n, A->rp, A->ci, A->v, p, Ap);         /* :138 */
        double alpha = so_dot(n, p, Ap);               /* :140 */
        double s = so_dot(n, r0, z0);                  /* :141 */
        alpha = s / alpha;                             /* :142 */
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) x[i] += alpha * p[i]; /* :144 */
        double nalpha = -alpha;
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) r0[i] += nalpha * Ap[i]; /* :145 */
        memset(z0, 0, sizeof(double) * (size_t)n);     /* :146 */
        so_amg_vcycle(h, r0, z0, 1);                   /* :147 */
        double beta = so_dot(n, z0, r0) / s;           /* :149 */
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) p[i] = z0[i] + beta * p[i]; /* :150 */
        r1 = so_nrm2(n, r0);                           /* :152 */
        if (hist) hist[count] = r1;
    }
    free(Ap);
    free(p);
    free(z0);
    free(r0);
    return count;
}

/* Restarted right-preconditioned GMRES(m) with CGS2 — see sparsh_oracle.h.  Not part of the reference. */
int so_gmres(so_amg *h, int n, const int *rp, const int *ci, const double *v, const double *b, double *x, double tol,
             int restart, int max_iter, double *hist) {
    if (h) {
        n = h->lev[0].n;
        rp = h->lev[0].rp;
        ci = h->lev[0].ci;
        v = h->lev[0].v;
    }
    const int m = restart > 0 ? restart : 30;
    size_t bytes = sizeof(double) * (size_t)n;
    double *V = (double *)xmalloc(bytes * (size_t)(m + 1));
    double *z = (double *)xmalloc(bytes), *w = (double *)xmalloc(bytes), *u = (double *)xmalloc(bytes);
    double *H = (double *)xcalloc((size_t)(m + 1) * (size_t)m, sizeof(double)); /* column j at H + j*(m+1) */
    double *cs = (double *)xmalloc(sizeof(double) * (size_t)m), *sn = (double *)xmalloc(sizeof(double) * (size_t)m);
    double *g = (double *)xmalloc(sizeof(double) * (size_t)(m + 1)), *y = (double *)xmalloc(sizeof(double) * (size_t)m);
    double *hc = (double *)xmalloc(sizeof(double) * (size_t)(m + 1));
    so_store_residual(n, rp, ci, v, b, x, w);
    double beta = so_nrm2(n, w);
    if (hist) hist[0] = beta;
    int it = 0;
    while (beta > tol && it < max_iter) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) V[i] = w[i] / beta;
        g[0] = beta;
        int j = 0;
        double est = beta;
        while (j < m && it < max_iter) {
            double *vj = V + (size_t)j * (size_t)n, *vn = V + (size_t)(j + 1) * (size_t)n, *Hj = H + (size_t)j * (size_t)(m + 1);
            if (h) {
                memset(z, 0, bytes);
                so_amg_vcycle(h, vj, z, 1);
                so_spmv(n, rp, ci, v, z, w);
            } else {
                so_spmv(n, rp, ci, v, vj, w);
            }
            for (int pass = 0; pass < 2; pass++) { /* classical Gram-Schmidt, twice */
                for (int i = 0; i <= j; i++) hc[i] = so_dot(n, V + (size_t)i * (size_t)n, w);
#pragma omp parallel for num_threads(g_threads) schedule(static)
                for (int r = 0; r < n; r++) {
                    double t = w[r];
                    for (int i = 0; i <= j; i++) t -= hc[i] * V[(size_t)i * (size_t)n + (size_t)r];
                    w[r] = t;
                }
                for (int i = 0; i <= j; i++) Hj[i] = pass == 0 ? hc[i] : Hj[i] + hc[i];
            }
            double hn = so_nrm2(n, w);
            Hj[j + 1] = hn;
            if (hn > 0.0) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
                for (int r = 0; r < n; r++) vn[r] = w[r] / hn;
            }
            for (int i = 0; i < j; i++) { /* earlier rotations */
                double t = cs[i] * Hj[i] + sn[i] * Hj[i + 1];
                Hj[i + 1] = -sn[i] * Hj[i] + cs[i] * Hj[i + 1];
                Hj[i] = t;
            }
            double den = sqrt(Hj[j] * Hj[j] + Hj[j + 1] * Hj[j + 1]);
            cs[j] = den > 0.0 ? Hj[j] / den : 1.0;
            sn[j] = den > 0.0 ? Hj[j + 1] / den : 0.0;
            Hj[j] = den;
            Hj[j + 1] = 0.0;
            g[j + 1] = -sn[j] * g[j];
            g[j] = cs[j] * g[j];
            est = fabs(g[j + 1]);
            j++;
            it++;
            if (hist) hist[it] = est;
            if (est <= tol || hn == 0.0) break;
        }
        for (int i = j - 1; i >= 0; i--) { /* back substitution */
            double t = g[i];
            for (int k = i + 1; k < j; k++) t -= H[(size_t)k * (size_t)(m + 1) + (size_t)i] * y[k];
            y[i] = t / H[(size_t)i * (size_t)(m + 1) + (size_t)i];
        }
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int r = 0; r < n; r++) {
            double t = 0.0;
            for (int i = 0; i < j; i++) t += y[i] * V[(size_t)i * (size_t)n + (size_t)r];
            u[r] = t;
        }
        if (h) {
            memset(z, 0, bytes);
            so_amg_vcycle(h, u, z, 1);
        } else {
            memcpy(z, u, bytes);
        }
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int r = 0; r < n; r++) x[r] += z[r];
        so_store_residual(n, rp, ci, v, b, x, w);
        beta = so_nrm2(n, w);
        if (hist) hist[it] = beta;
        if (j == 0) break;
    }
    free(V);
    free(z);
    free(w);
    free(u);
    free(H);
    free(cs);
    free(sn);
    free(g);
    free(y);
    free(hc);
    return it;
}

/* src/AMG_main_solvers.cpp:358-458 */
int so_pbicgstab(so_amg *h, const double *b, double *x, double tol, int max_iter, double *hist) {
    so_level *A = &h->lev[0];
    int n = A->n;
    size_t bytes = sizeof(double) * (size_t)n;
    double *r0 = (double *)xmalloc(bytes), *r = (double *)xmalloc(bytes), *p = (double *)xmalloc(bytes);
    double *Ap = (double *)xmalloc(bytes), *s = (double *)xmalloc(bytes), *As = (double *)xmalloc(bytes);
    double *p1 = (double *)xmalloc(bytes), *s1 = (double *)xmalloc(bytes);
    so_store_residual(n, A->rp, A->ci, A->v, b, x, r0); /* :383-384 */
    memcpy(r, r0, bytes);                              /* :387 */
    memcpy(p, r0, bytes);                              /* :388 */
    double res = so_nrm2(n, r0);                       /* :390 */
    if (hist) hist[0] = res;
    int count = 0;
    while (res > tol && count < max_iter) {            /* :397 (no cap in the reference) */
        memset(p1, 0, bytes);                          /* :399 */
        so_amg_vcycle(h, p, p1, 1);                    /* :400 */
        double alpha1 = so_dot(n, r, r0);              /* :402 */
        so_spmv(n, A->rp, A->ci, A->v, p1, Ap);        /* :403 */
        double alpha = so_dot(n, Ap, r0);              /* :404 */
        alpha = alpha1 / alpha;                        /* :406 */
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) s[i] = r[i] - alpha * Ap[i]; /* :411 */
        memset(s1, 0, bytes);                          /* :414 */
        so_amg_vcycle(h, s, s1, 1);                    /* :415 */
        so_spmv(n, A->rp, A->ci, A->v, s1, As);        /* :416 */
        double omega1 = so_dot(n, As, s);              /* :418 */
        omega1 /= so_dot(n, As, As);                   /* :419 */
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) {
            x[i] = x[i] + alpha * p1[i] + omega1 * s1[i]; /* :424 */
            r[i] = s[i] - omega1 * As[i];                 /* :425 */
        }
        double beta = so_dot(n, r, r0) / alpha1;       /* :428 */
        beta = beta * (alpha / omega1);                /* :429 */
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) p[i] = r[i] + beta * (p[i] - omega1 * Ap[i]); /* :434 */
        res = so_nrm2(n, r);                           /* :437 */
        count++;
        if (hist) hist[count] = res;
    }
    free(r0);
    free(r);
    free(p);
    free(Ap);
    free(s);
    free(As);
    free(p1);
    free(s1);
    return count;
}

/* src/AMG_main_solvers.cpp:47-103.  The reference overwrites its computed residual with b (:64), i.e. it
 * assumes x0 = 0; the restatement keeps that. */
int so_cg(int n, const int *rp, const int *ci, const double *v, const double *b, double *x, double tol,
          int max_iter, double *hist) {
    size_t bytes = sizeof(double) * (size_t)n;
    double *Ap = (double *)xmalloc(bytes), *p = (double *)xmalloc(bytes), *r = (double *)xmalloc(bytes);
    memcpy(r, b, bytes); /* :64 */
    memcpy(p, r, bytes); /* :65 */
    double r1 = so_nrm2(n, r);
    if (hist) hist[0] = r1;
    int count = 0;
    while (count < n && count < max_iter && r1 > tol) { /* :70 */
        count++;
        so_spmv(n, rp, ci, v, p, Ap);    /* :73 */
        double alpha = so_dot(n, p, Ap); /* :75 */
        double s = so_dot(n, r, r);      /* :76 */
        alpha = s / alpha;
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) x[i] += alpha * p[i];
        double nalpha = -alpha;
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) r[i] += nalpha * Ap[i];
        double beta = so_dot(n, r, r) / s; /* :82 */
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) p[i] = r[i] + beta * p[i]; /* :83 */
        r1 = sqrt(s * beta);             /* :85 */
        if (hist) hist[count] = r1;
    }
    free(Ap);
    free(p);
    free(r);
    return count;
}

/* src/AMG_main_solvers.cpp:271-355 */
int so_bicgstab(int n, const int *rp, const int *ci, const double *v, const double *b, double *x, double tol,
                int max_iter, double *hist) {
    size_t bytes = sizeof(double) * (size_t)n;
    double *r0 = (double *)xmalloc(bytes), *r = (double *)xmalloc(bytes), *p = (double *)xmalloc(bytes);
    double *Ap = (double *)xmalloc(bytes), *s = (double *)xmalloc(bytes), *As = (double *)xmalloc(bytes);
    so_store_residual(n, rp, ci, v, b, x, r0);
    memcpy(r, r0, bytes);
    memcpy(p, r0, bytes);
    double res = so_nrm2(n, r0);
    if (hist) hist[0] = res;
    int count = 0;
    while (res > tol && count < max_iter) {
        double alpha1 = so_dot(n, r, r0);
        so_spmv(n, rp, ci, v, p, Ap);
        double alpha = so_dot(n, Ap, r0);
        alpha = alpha1 / alpha;
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) s[i] = r[i] - alpha * Ap[i];
        so_spmv(n, rp, ci, v, s, As);
        double omega1 = so_dot(n, As, s);
        omega1 /= so_dot(n, As, As);
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) {
            x[i] = x[i] + alpha * p[i] + omega1 * s[i];
            r[i] = s[i] - omega1 * As[i];
        }
        double beta = so_dot(n, r, r0) / alpha1;
        beta = beta * (alpha / omega1);
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++) p[i] = r[i] + beta * (p[i] - omega1 * Ap[i]);
        res = so_nrm2(n, r);
        count++;
        if (hist) hist[count] = res;
    }
    free(r0);
    free(r);
    free(p);
    free(Ap);
    free(s);
    free(As);
    return count;
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic matrices: natural ordering, x fastest, Dirichlet by truncation.
 * ------------------------------------------------------------------------------------------------ */
void so_gen_poisson2d_5pt(int nx, int ny, int **rp_o, int **ci_o, double **v_o) {
    so_gen_poisson3d_7pt(nx, ny, 1, rp_o, ci_o, v_o);
    /* a 1-plane 7-point stencil has diagonal 6; the 2D operator has 4 */
    int n = nx * ny;
    int *rp = *rp_o, *ci = *ci_o;
    double *v = *v_o;
    for (int i = 0; i < n; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (ci[j] == i) v[j] = 4.0;
}

void so_gen_poisson3d_7pt(int nx, int ny, int nz, int **rp_o, int **ci_o, double **v_o) {
    long n = (long)nx * ny * nz;
    int *rp = (int *)xmalloc(sizeof(int) * ((size_t)n + 1));
    rp[0] = 0;
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                long i = ((long)z * ny + y) * nx + x;
                int c = 1 + (x > 0) + (x < nx - 1) + (y > 0) + (y < ny - 1) + (z > 0) + (z < nz - 1);
                rp[i + 1] = rp[i] + c;
            }
    int *ci = (int *)xmalloc(sizeof(int) * (size_t)rp[n]);
    double *v = (double *)xmalloc(sizeof(double) * (size_t)rp[n]);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                long i = ((long)z * ny + y) * nx + x;
                int o = rp[i];
                if (z > 0) { ci[o] = (int)(i - (long)nx * ny); v[o++] = -1.0; }
                if (y > 0) { ci[o] = (int)(i - nx); v[o++] = -1.0; }
                if (x > 0) { ci[o] = (int)(i - 1); v[o++] = -1.0; }
                ci[o] = (int)i; v[o++] = 6.0;
                if (x < nx - 1) { ci[o] = (int)(i + 1); v[o++] = -1.0; }
                if (y < ny - 1) { ci[o] = (int)(i + nx); v[o++] = -1.0; }
                if (z < nz - 1) { ci[o] = (int)(i + (long)nx * ny); v[o++] = -1.0; }
            }
    *rp_o = rp;
    *ci_o = ci;
    *v_o = v;
}
