/*
 * mkl.h — MKL-compat shim (TEST INFRASTRUCTURE, not product code).
 *
 * Intel MKL is closed source, unpinned by the reference (README.md:26 only
 * "recommends" MKL 2020; CMakeLists.txt:42 links mkl_rt) and not installable
 * offline in this image.  This header declares exactly the MKL surface the
 * reference's host sources use so that they compile UNMODIFIED:
 *   include/AMG_cpu_matrix.hpp:6-8        (the three #includes)
 *   src/AMG_cpu_matrix.cpp:22,29,58,167-187
 *   src/AMG_cycle_utilities.cpp:23-72,88-138,174-219
 *   src/AMG_smoothers.cpp:11-25,55-63
 *   src/AMG_coarse_level_solver.cpp:49-52,67-70
 *   src/AMG_main_solvers.cpp (mkl_sparse_d_mv / cblas_* call sites)
 * The implementation is oracle/mklshim/mklshim.cpp (C++/OpenMP, our own code,
 * following Intel's published interface semantics).
 */
#ifndef SPARSH_MKLSHIM_MKL_H_
#define SPARSH_MKLSHIM_MKL_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef int MKL_INT;

/* ---- inspector-executor sparse BLAS ---- */
struct sparsh_shim_csr;
typedef struct sparsh_shim_csr *sparse_matrix_t;

typedef enum {
    SPARSE_STATUS_SUCCESS = 0,
    SPARSE_STATUS_NOT_INITIALIZED = 1,
    SPARSE_STATUS_ALLOC_FAILED = 2,
    SPARSE_STATUS_INVALID_VALUE = 3,
    SPARSE_STATUS_EXECUTION_FAILED = 4,
    SPARSE_STATUS_INTERNAL_ERROR = 5,
    SPARSE_STATUS_NOT_SUPPORTED = 6
} sparse_status_t;

typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;

typedef enum {
    SPARSE_OPERATION_NON_TRANSPOSE = 10,
    SPARSE_OPERATION_TRANSPOSE = 11,
    SPARSE_OPERATION_CONJUGATE_TRANSPOSE = 12
} sparse_operation_t;

typedef enum {
    SPARSE_MATRIX_TYPE_GENERAL = 20,
    SPARSE_MATRIX_TYPE_SYMMETRIC = 21,
    SPARSE_MATRIX_TYPE_HERMITIAN = 22,
    SPARSE_MATRIX_TYPE_TRIANGULAR = 23,
    SPARSE_MATRIX_TYPE_DIAGONAL = 24
} sparse_matrix_type_t;

typedef enum { SPARSE_FILL_MODE_LOWER = 40, SPARSE_FILL_MODE_UPPER = 41, SPARSE_FILL_MODE_FULL = 42 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50, SPARSE_DIAG_UNIT = 51 } sparse_diag_type_t;

struct matrix_descr {
    sparse_matrix_type_t type;
    sparse_fill_mode_t mode;
    sparse_diag_type_t diag;
};

typedef enum {
    SPARSE_STAGE_FULL_MULT = 90,
    SPARSE_STAGE_NNZ_COUNT = 91,
    SPARSE_STAGE_FINALIZE_MULT = 92,
    SPARSE_STAGE_FULL_MULT_NO_VAL = 93,
    SPARSE_STAGE_FINALIZE_MULT_NO_VAL = 94
} sparse_request_t;

sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t *A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols,
                                        MKL_INT *rows_start, MKL_INT *rows_end, MKL_INT *col_indx, double *values);
sparse_status_t mkl_sparse_destroy(sparse_matrix_t A);
sparse_status_t mkl_sparse_order(sparse_matrix_t A);
sparse_status_t mkl_sparse_d_mv(sparse_operation_t operation, double alpha, const sparse_matrix_t A,
                                struct matrix_descr descr, const double *x, double beta, double *y);
sparse_status_t mkl_sparse_spmm(sparse_operation_t operation, const sparse_matrix_t A, const sparse_matrix_t B,
                                sparse_matrix_t *C);
sparse_status_t mkl_sparse_sp2m(sparse_operation_t transA, struct matrix_descr descrA, const sparse_matrix_t A,
                                sparse_operation_t transB, struct matrix_descr descrB, const sparse_matrix_t B,
                                sparse_request_t request, sparse_matrix_t *C);
sparse_status_t mkl_sparse_d_export_csr(const sparse_matrix_t source, sparse_index_base_t *indexing, MKL_INT *rows,
                                        MKL_INT *cols, MKL_INT **rows_start, MKL_INT **rows_end, MKL_INT **col_indx,
                                        double **values);

/* ---- CBLAS level 1 ---- */
void cblas_daxpy(MKL_INT n, double a, const double *x, MKL_INT incx, double *y, MKL_INT incy);
void cblas_daxpby(MKL_INT n, double a, const double *x, MKL_INT incx, double b, double *y, MKL_INT incy);
double cblas_ddot(MKL_INT n, const double *x, MKL_INT incx, const double *y, MKL_INT incy);
double cblas_dnrm2(MKL_INT n, const double *x, MKL_INT incx);

/* ---- service ---- */
void mkl_set_num_threads(int nt);
void mkl_set_dynamic(int flag);
void kmp_set_warnings_off(void);

/* ---- PARDISO (phases 12 and 33, mtype 11, nrhs 1 only) ---- */
void PARDISO(void *pt, const MKL_INT *maxfct, const MKL_INT *mnum, const MKL_INT *mtype, const MKL_INT *phase,
             const MKL_INT *n, const void *a, const MKL_INT *ia, const MKL_INT *ja, MKL_INT *perm,
             const MKL_INT *nrhs, MKL_INT *iparm, const MKL_INT *msglvl, void *b, void *x, MKL_INT *error);

#ifdef __cplusplus
}
#endif
#endif
