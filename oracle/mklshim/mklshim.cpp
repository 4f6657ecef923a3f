// mklshim.cpp — implementation of the MKL surface declared in mkl.h (TEST INFRASTRUCTURE).
//
// Lets the reference's host sources run unmodified without Intel MKL.  Semantics follow Intel's public
// interface documentation for each routine; where MKL's internal evaluation order is unspecified the shim
// picks the obvious sequential one (row sums left to right, transposed products scattered in row order) so
// that it agrees bit for bit with oracle/sparsh_oracle.c.
//
// Traps in the reference's usage that the shim must honour:
//  * create_csr wraps the caller's arrays (no copy); mkl_sparse_order sorts columns IN PLACE in them
//    (src/AMG_cpu_matrix.cpp:22-29).
//  * export_csr is called with rows_end == &rowptr + 1, which aliases &colindex
//    (src/AMG_cycle_utilities.cpp:138): *rows_end must be stored BEFORE *col_indx.
//  * cblas_daxpy(n,-1.0,b,1.0,h,1) passes a double as incx (src/AMG_cycle_utilities.cpp:90).
#include "mkl.h"

#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../sparsh_oracle.h"

struct sparsh_shim_csr {
    int rows = 0, cols = 0;
    int *rs = nullptr;  // rows_start
    int *re = nullptr;  // rows_end
    int *ci = nullptr;
    double *v = nullptr;
    bool owned = false;  // arrays allocated by the shim (spmm / sp2m results)
};

static int g_nt = 1;

extern "C" {

void mkl_set_num_threads(int nt) {
    g_nt = nt > 0 ? nt : 1;
    so_set_threads(g_nt);
}
void mkl_set_dynamic(int) {}
void kmp_set_warnings_off(void) {}

sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t *A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols,
                                        MKL_INT *rows_start, MKL_INT *rows_end, MKL_INT *col_indx, double *values) {
    if (indexing != SPARSE_INDEX_BASE_ZERO) return SPARSE_STATUS_NOT_SUPPORTED;
    sparsh_shim_csr *h = new sparsh_shim_csr();
    h->rows = rows;
    h->cols = cols;
    h->rs = rows_start;
    h->re = rows_end;
    h->ci = col_indx;
    h->v = values;
    *A = h;
    return SPARSE_STATUS_SUCCESS;
}

sparse_status_t mkl_sparse_destroy(sparse_matrix_t A) {
    if (!A) return SPARSE_STATUS_NOT_INITIALIZED;
    if (A->owned) {
        free(A->rs);
        free(A->ci);
        free(A->v);
    }
    delete A;
    return SPARSE_STATUS_SUCCESS;
}

sparse_status_t mkl_sparse_order(sparse_matrix_t A) {
    if (!A) return SPARSE_STATUS_NOT_INITIALIZED;
#pragma omp parallel for num_threads(g_nt) schedule(dynamic, 1024)
    for (int i = 0; i < A->rows; i++) {
        int lo = A->rs[i], len = A->re[i] - lo;
        int *c = A->ci + lo;
        double *v = A->v + lo;
        for (int a = 1; a < len; a++) {
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
    return SPARSE_STATUS_SUCCESS;
}

sparse_status_t mkl_sparse_d_mv(sparse_operation_t op, double alpha, const sparse_matrix_t A, struct matrix_descr,
                                const double *x, double beta, double *y) {
    if (!A) return SPARSE_STATUS_NOT_INITIALIZED;
    if (op == SPARSE_OPERATION_NON_TRANSPOSE) {
#pragma omp parallel for num_threads(g_nt) schedule(static)
        for (int i = 0; i < A->rows; i++) {
            double s = 0.0;
            for (int j = A->rs[i]; j < A->re[i]; j++) s += A->v[j] * x[A->ci[j]];
            if (alpha != 1.0) s = alpha * s;
            y[i] = beta == 0.0 ? s : s + beta * y[i];
        }
    } else {
        for (int c = 0; c < A->cols; c++) y[c] = beta == 0.0 ? 0.0 : beta * y[c];
        for (int i = 0; i < A->rows; i++) {
            double xi = alpha * x[i];
            for (int j = A->rs[i]; j < A->re[i]; j++) y[A->ci[j]] += A->v[j] * xi;
        }
    }
    return SPARSE_STATUS_SUCCESS;
}

static sparsh_shim_csr *transpose(const sparsh_shim_csr *A) {
    sparsh_shim_csr *T = new sparsh_shim_csr();
    T->rows = A->cols;
    T->cols = A->rows;
    T->owned = true;
    size_t nnz = 0;
    for (int i = 0; i < A->rows; i++) nnz += (size_t)(A->re[i] - A->rs[i]);
    T->rs = (int *)calloc((size_t)T->rows + 1, sizeof(int));
    T->re = T->rs + 1;
    T->ci = (int *)malloc(sizeof(int) * (nnz ? nnz : 1));
    T->v = (double *)malloc(sizeof(double) * (nnz ? nnz : 1));
    for (int i = 0; i < A->rows; i++)
        for (int j = A->rs[i]; j < A->re[i]; j++) T->rs[A->ci[j] + 1]++;
    for (int c = 0; c < T->rows; c++) T->rs[c + 1] += T->rs[c];
    std::vector<int> cur(T->rs, T->rs + T->rows);
    for (int i = 0; i < A->rows; i++)
        for (int j = A->rs[i]; j < A->re[i]; j++) {
            int d = cur[A->ci[j]]++;
            T->ci[d] = i;
            T->v[d] = A->v[j];
        }
    return T;
}

// Gustavson C = A B.  Entry order inside a row: first touch; each value accumulated in traversal order.
static sparsh_shim_csr *multiply(const sparsh_shim_csr *A, const sparsh_shim_csr *B) {
    sparsh_shim_csr *C = new sparsh_shim_csr();
    C->rows = A->rows;
    C->cols = B->cols;
    C->owned = true;
    C->rs = (int *)calloc((size_t)C->rows + 1, sizeof(int));
    C->re = C->rs + 1;
    int ncol = B->cols > 0 ? B->cols : 1;
#pragma omp parallel num_threads(g_nt)
    {
        std::vector<int> mark((size_t)ncol, -1);
#pragma omp for schedule(dynamic, 2048)
        for (int i = 0; i < A->rows; i++) {
            int cnt = 0;
            for (int ja = A->rs[i]; ja < A->re[i]; ja++) {
                int k = A->ci[ja];
                for (int jb = B->rs[k]; jb < B->re[k]; jb++) {
                    int c = B->ci[jb];
                    if (mark[c] != i) {
                        mark[c] = i;
                        cnt++;
                    }
                }
            }
            C->rs[i + 1] = cnt;
        }
    }
    for (int i = 0; i < C->rows; i++) C->rs[i + 1] += C->rs[i];
    size_t nnz = (size_t)C->rs[C->rows];
    C->ci = (int *)malloc(sizeof(int) * (nnz ? nnz : 1));
    C->v = (double *)malloc(sizeof(double) * (nnz ? nnz : 1));
#pragma omp parallel num_threads(g_nt)
    {
        std::vector<int> pos((size_t)ncol, -1);
#pragma omp for schedule(dynamic, 2048)
        for (int i = 0; i < A->rows; i++) {
            int base = C->rs[i], o = base;
            for (int ja = A->rs[i]; ja < A->re[i]; ja++) {
                int k = A->ci[ja];
                double a = A->v[ja];
                for (int jb = B->rs[k]; jb < B->re[k]; jb++) {
                    int c = B->ci[jb];
                    if (pos[c] < base) {
                        pos[c] = o;
                        C->ci[o] = c;
                        C->v[o] = a * B->v[jb];
                        o++;
                    } else {
                        C->v[pos[c]] += a * B->v[jb];
                    }
                }
            }
        }
    }
    return C;
}

sparse_status_t mkl_sparse_sp2m(sparse_operation_t transA, struct matrix_descr, const sparse_matrix_t A,
                                sparse_operation_t transB, struct matrix_descr, const sparse_matrix_t B,
                                sparse_request_t request, sparse_matrix_t *C) {
    if (!A || !B) return SPARSE_STATUS_NOT_INITIALIZED;
    if (request != SPARSE_STAGE_FULL_MULT) return SPARSE_STATUS_NOT_SUPPORTED;
    sparsh_shim_csr *At = transA == SPARSE_OPERATION_NON_TRANSPOSE ? nullptr : transpose(A);
    sparsh_shim_csr *Bt = transB == SPARSE_OPERATION_NON_TRANSPOSE ? nullptr : transpose(B);
    *C = multiply(At ? At : A, Bt ? Bt : B);
    if (At) mkl_sparse_destroy(At);
    if (Bt) mkl_sparse_destroy(Bt);
    return SPARSE_STATUS_SUCCESS;
}

sparse_status_t mkl_sparse_spmm(sparse_operation_t operation, const sparse_matrix_t A, const sparse_matrix_t B,
                                sparse_matrix_t *C) {
    matrix_descr d;
    d.type = SPARSE_MATRIX_TYPE_GENERAL;
    d.mode = SPARSE_FILL_MODE_FULL;
    d.diag = SPARSE_DIAG_NON_UNIT;
    return mkl_sparse_sp2m(operation, d, A, SPARSE_OPERATION_NON_TRANSPOSE, d, B, SPARSE_STAGE_FULL_MULT, C);
}

sparse_status_t mkl_sparse_d_export_csr(const sparse_matrix_t A, sparse_index_base_t *indexing, MKL_INT *rows,
                                        MKL_INT *cols, MKL_INT **rows_start, MKL_INT **rows_end, MKL_INT **col_indx,
                                        double **values) {
    if (!A) return SPARSE_STATUS_NOT_INITIALIZED;
    int *rs = A->rs, *re = A->re, *ci = A->ci;
    double *v = A->v;
    *indexing = SPARSE_INDEX_BASE_ZERO;
    *rows = A->rows;
    *cols = A->cols;
    *rows_start = rs;
    *rows_end = re;   // may alias the caller's colindex slot: written first on purpose ...
    *col_indx = ci;   // ... and overwritten here with the right pointer
    *values = v;
    return SPARSE_STATUS_SUCCESS;
}

// ---- CBLAS level 1 (unit strides only: that is all the reference uses) ----
void cblas_daxpy(MKL_INT n, double a, const double *x, MKL_INT, double *y, MKL_INT) {
#pragma omp parallel for num_threads(g_nt) schedule(static)
    for (int i = 0; i < n; i++) y[i] += a * x[i];
}
void cblas_daxpby(MKL_INT n, double a, const double *x, MKL_INT, double b, double *y, MKL_INT) {
#pragma omp parallel for num_threads(g_nt) schedule(static)
    for (int i = 0; i < n; i++) y[i] = a * x[i] + b * y[i];
}
double cblas_ddot(MKL_INT n, const double *x, MKL_INT, const double *y, MKL_INT) {
    so_set_threads(g_nt);
    return so_dot(n, x, y);
}
double cblas_dnrm2(MKL_INT n, const double *x, MKL_INT) {
    so_set_threads(g_nt);
    return so_nrm2(n, x);
}

// ---- PARDISO: phase 12 = analyse + factor, phase 33 = solve (src/AMG_coarse_level_solver.cpp:51-52,66-70) ----
void PARDISO(void *pt, const MKL_INT *, const MKL_INT *, const MKL_INT *mtype, const MKL_INT *phase, const MKL_INT *n,
             const void *a, const MKL_INT *ia, const MKL_INT *ja, MKL_INT *, const MKL_INT *nrhs, MKL_INT *iparm,
             const MKL_INT *, void *b, void *x, MKL_INT *error) {
    void **slot = (void **)pt;
    *error = 0;
    if (*mtype != 11 || *nrhs != 1 || iparm[34] != 1) {
        *error = -1;
        return;
    }
    so_set_threads(g_nt);
    if (*phase == 12) {
        if (slot[0]) so_lu_free((so_lu *)slot[0]);
        slot[0] = so_lu_factor(*n, ia, ja, (const double *)a);
    } else if (*phase == 33) {
        if (!slot[0]) {
            *error = -2;
            return;
        }
        so_lu_solve((const so_lu *)slot[0], (const double *)b, (double *)x);
    } else if (*phase == -1) {
        if (slot[0]) so_lu_free((so_lu *)slot[0]);
        slot[0] = nullptr;
    } else {
        *error = -3;
    }
}

}  // extern "C"
