// rap.cu — Galerkin product A_c = P^T (A P) on the device (replaces, for the setup, what the reference does
// with two mkl_sparse_spmm calls on the host: parallel::coarsen_matrix, src/AMG_cycle_utilities.cpp:126-146).
//
// The product is a pair of row-wise (Gustavson) sparse products, one thread per output row, written so that its result
// is the one host/setup.cpp computes BIT FOR BIT: the entries of a row appear in first-touch order and each is
// accumulated in traversal order with unfused multiply/add, then the columns are sorted (sp_matrix_fill).  Integer
// output (row pointers, sorted column indices) equals the reference's; the coarsening decisions that follow on the host
// (HEM / Beck are sequential greedy sweeps) therefore see identical matrices.
//   transpose   R = P^T, stable (a coarse row lists its fine rows ascending): count, scan, scatter, per-row sort
//   spgemm      two passes (count distinct columns per row -> scan -> fill), rows longer than RAP_MAX_ROW entries make the
//               call report "not applicable" and the host product runs instead (never a silent difference)
// No cuSPARSE / thrust: the scans and sorts are the few kernels below.
#include <vector>

#include "internal.cuh"

namespace sparsh {

namespace {

constexpr int RAP_MAX_ROW = 192;  // distinct columns per product row handled by the per-thread list
constexpr int SCAN_T = 1024, SCAN_ITEMS = 4;

struct DevCsr {
    int nrow = 0, ncol = 0, nnz = 0;
    int *rp = nullptr, *ci = nullptr;
    double *v = nullptr;
    void release() {
        cudaFree(rp);
        cudaFree(ci);
        cudaFree(v);
        rp = ci = nullptr;
        v = nullptr;
    }
};

// ---- exclusive scan: out[0] = 0, out[i + 1] = in[0] + ... + in[i] ------------------------------------------------------
__global__ void __launch_bounds__(SCAN_T) scan_block_kernel(const int *__restrict__ in, int n, int *out, int *block_sum) {
    __shared__ int warp_tot[SCAN_T / 32];
    const int base = blockIdx.x * SCAN_T * SCAN_ITEMS + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], run = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        run += v[k];
    }
    int incl = run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += t;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    int excl = incl - run + (warp > 0 ? warp_tot[warp - 1] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        excl += v[k];
        if (base + k < n) out[base + k + 1] = excl;  // inclusive value at position +1: exclusive scan shifted by one
    }
    if (threadIdx.x == SCAN_T - 1) block_sum[blockIdx.x] = excl;
}
__global__ void __launch_bounds__(SCAN_T) scan_sums_kernel(int *block_sum, int nblocks) {  // one CTA: exclusive scan in place
    __shared__ int carry_s;
    __shared__ int warp_tot[SCAN_T / 32];
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int start = 0; start < nblocks; start += SCAN_T) {
        const int i = start + threadIdx.x;
        const int x = i < nblocks ? block_sum[i] : 0;
        int incl = x;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + incl - x + (warp > 0 ? warp_tot[warp - 1] : 0);
        if (i < nblocks) block_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == SCAN_T - 1) carry_s = excl + x;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SCAN_T) scan_add_kernel(int *out, int n, const int *__restrict__ block_off) {
    const int base = blockIdx.x * SCAN_T * SCAN_ITEMS + threadIdx.x * SCAN_ITEMS;
    const int off = block_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k + 1] += off;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}
int exclusive_scan(const int *in, int n, int *out, cudaStream_t st) {
    if (n == 0) {
        SP_CUDA(cudaMemsetAsync(out, 0, sizeof(int), st));
        return SPARSH_OK;
    }
    const int nblocks = (n + SCAN_T * SCAN_ITEMS - 1) / (SCAN_T * SCAN_ITEMS);
    int *sums = nullptr;
    SP_CUDA(cudaMalloc(&sums, sizeof(int) * (size_t)nblocks));
    scan_block_kernel<<<nblocks, SCAN_T, 0, st>>>(in, n, out, sums);
    scan_sums_kernel<<<1, SCAN_T, 0, st>>>(sums, nblocks);
    scan_add_kernel<<<nblocks, SCAN_T, 0, st>>>(out, n, sums);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(sums);
    SP_CUDA(e);
    SP_CUDA(cudaGetLastError());
    return SPARSH_OK;
}

// ---- per-row insertion sort by column (rows are short) ----------------------------------------------------------------
__global__ void __launch_bounds__(256) sort_rows_kernel(int nrow, const int *__restrict__ rp, int *ci, double *v) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nrow) return;
    const int lo = rp[i], hi = rp[i + 1];
    for (int a = lo + 1; a < hi; a++) {
        const int c = ci[a];
        const double x = v[a];
        int b = a - 1;
        while (b >= lo && ci[b] > c) {
            ci[b + 1] = ci[b];
            v[b + 1] = v[b];
            b--;
        }
        ci[b + 1] = c;
        v[b + 1] = x;
    }
}

// ---- transpose ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) count_cols_kernel(int nnz, const int *__restrict__ ci, int *count) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j < nnz) atomicAdd(&count[ci[j]], 1);
}
__global__ void __launch_bounds__(256)
    scatter_T_kernel(int nrow, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ v,
                     const int *__restrict__ trp, int *cursor, int *tci, double *tv) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nrow) return;
    for (int j = rp[i]; j < rp[i + 1]; j++) {
        const int c = ci[j];
        const int d = trp[c] + atomicAdd(&cursor[c], 1);
        tci[d] = i;  // the order inside a row of T is fixed afterwards by sort_rows_kernel (ascending source row)
        tv[d] = v[j];
    }
}
int transpose(const DevCsr &M, DevCsr *T, cudaStream_t st) {
    T->nrow = M.ncol;
    T->ncol = M.nrow;
    T->nnz = M.nnz;
    int *count = nullptr;
    SP_CUDA(cudaMalloc(&T->rp, sizeof(int) * ((size_t)M.ncol + 1)));
    SP_CUDA(cudaMalloc(&T->ci, sizeof(int) * (size_t)std::max(M.nnz, 1)));
    SP_CUDA(cudaMalloc(&T->v, sizeof(double) * (size_t)std::max(M.nnz, 1)));
    SP_CUDA(cudaMalloc(&count, sizeof(int) * (size_t)std::max(M.ncol, 1)));
    SP_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)std::max(M.ncol, 1), st));
    if (M.nnz > 0) count_cols_kernel<<<(M.nnz + 255) / 256, 256, 0, st>>>(M.nnz, M.ci, count);
    int rc = exclusive_scan(count, M.ncol, T->rp, st);
    if (rc == SPARSH_OK) {
        cudaMemsetAsync(count, 0, sizeof(int) * (size_t)std::max(M.ncol, 1), st);
        if (M.nrow > 0) scatter_T_kernel<<<(M.nrow + 255) / 256, 256, 0, st>>>(M.nrow, M.rp, M.ci, M.v, T->rp, count, T->ci, T->v);
        if (T->nrow > 0) sort_rows_kernel<<<(T->nrow + 255) / 256, 256, 0, st>>>(T->nrow, T->rp, T->ci, T->v);
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = SPARSH_ERR_CUDA;
    }
    cudaFree(count);
    return rc;
}

// ---- C = A * B, Gustavson, one thread per row ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    spgemm_count_kernel(int arow, const int *__restrict__ arp, const int *__restrict__ aci, const int *__restrict__ brp,
                        const int *__restrict__ bci, int *count, int *too_long) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= arow) return;
    int list[RAP_MAX_ROW];
    int cnt = 0;
    for (int ja = arp[i]; ja < arp[i + 1]; ja++) {
        const int k = aci[ja];
        for (int jb = brp[k]; jb < brp[k + 1]; jb++) {
            const int c = bci[jb];
            int t = 0;
            while (t < cnt && list[t] != c) t++;
            if (t == cnt) {
                if (cnt == RAP_MAX_ROW) {
                    atomicExch(too_long, 1);
                    count[i] = 0;
                    return;
                }
                list[cnt++] = c;
            }
        }
    }
    count[i] = cnt;
}
// entries of row i of C in first-touch order, each accumulated in traversal order with separate multiply and add:
// the host product's arithmetic (host/setup.cpp: spgemm), hence its bits
__global__ void __launch_bounds__(128)
    spgemm_fill_kernel(int arow, const int *__restrict__ arp, const int *__restrict__ aci, const double *__restrict__ av,
                       const int *__restrict__ brp, const int *__restrict__ bci, const double *__restrict__ bv,
                       const int *__restrict__ crp, int *cci, double *cv) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= arow) return;
    const int base = crp[i];
    int *cc = cci + base;
    double *cx = cv + base;
    int o = 0;
    for (int ja = arp[i]; ja < arp[i + 1]; ja++) {
        const int k = aci[ja];
        const double a = av[ja];
        for (int jb = brp[k]; jb < brp[k + 1]; jb++) {
            const int c = bci[jb];
            const double prod = __dmul_rn(a, bv[jb]);
            int t = 0;
            while (t < o && cc[t] != c) t++;
            if (t == o) {
                cc[o] = c;
                cx[o] = prod;
                o++;
            } else {
                cx[t] = __dadd_rn(cx[t], prod);
            }
        }
    }
}
// returns SPARSH_OK with *applicable = false when a row of the product has more than RAP_MAX_ROW entries
int spgemm(const DevCsr &A, const DevCsr &B, DevCsr *C, bool *applicable, cudaStream_t st) {
    *applicable = true;
    C->nrow = A.nrow;
    C->ncol = B.ncol;
    int *count = nullptr, *flag = nullptr;
    SP_CUDA(cudaMalloc(&count, sizeof(int) * (size_t)std::max(A.nrow, 1)));
    SP_CUDA(cudaMalloc(&flag, sizeof(int)));
    SP_CUDA(cudaMalloc(&C->rp, sizeof(int) * ((size_t)A.nrow + 1)));
    cudaMemsetAsync(flag, 0, sizeof(int), st);
    const int grid = (A.nrow + 127) / 128;
    if (A.nrow > 0) spgemm_count_kernel<<<grid, 128, 0, st>>>(A.nrow, A.rp, A.ci, B.rp, B.ci, count, flag);
    int rc = exclusive_scan(count, A.nrow, C->rp, st);
    int h_flag = 0, nnz = 0;
    if (rc == SPARSH_OK) {
        cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&nnz, C->rp + A.nrow, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = SPARSH_ERR_CUDA;
    }
    cudaFree(count);
    cudaFree(flag);
    if (rc != SPARSH_OK) return rc;
    if (h_flag) {
        *applicable = false;
        return SPARSH_OK;
    }
    C->nnz = nnz;
    SP_CUDA(cudaMalloc(&C->ci, sizeof(int) * (size_t)std::max(nnz, 1)));
    SP_CUDA(cudaMalloc(&C->v, sizeof(double) * (size_t)std::max(nnz, 1)));
    if (A.nrow > 0) spgemm_fill_kernel<<<grid, 128, 0, st>>>(A.nrow, A.rp, A.ci, A.v, B.rp, B.ci, B.v, C->rp, C->ci, C->v);
    SP_CUDA(cudaGetLastError());
    return SPARSH_OK;
}

int upload_csr(int nrow, int ncol, const int *rp, const int *ci, const double *v, DevCsr *M, cudaStream_t st) {
    M->nrow = nrow;
    M->ncol = ncol;
    M->nnz = rp[nrow];
    SP_CUDA(cudaMalloc(&M->rp, sizeof(int) * ((size_t)nrow + 1)));
    SP_CUDA(cudaMalloc(&M->ci, sizeof(int) * (size_t)std::max(M->nnz, 1)));
    SP_CUDA(cudaMalloc(&M->v, sizeof(double) * (size_t)std::max(M->nnz, 1)));
    SP_CUDA(cudaMemcpyAsync(M->rp, rp, sizeof(int) * ((size_t)nrow + 1), cudaMemcpyHostToDevice, st));
    if (M->nnz > 0) {
        SP_CUDA(cudaMemcpyAsync(M->ci, ci, sizeof(int) * (size_t)M->nnz, cudaMemcpyHostToDevice, st));
        SP_CUDA(cudaMemcpyAsync(M->v, v, sizeof(double) * (size_t)M->nnz, cudaMemcpyHostToDevice, st));
    }
    return SPARSH_OK;
}

}  // namespace

}  // namespace sparsh

using namespace sparsh;

struct sparsh_rap_s {
    DevCsr C;  // the Galerkin product on the device, columns sorted
};

extern "C" {

int sparsh_galerkin_rap(int nrow, const int *h_rowptr, const int *h_colindex, const double *h_val, int ncoarse,
                        const int *h_p_rowptr, const int *h_p_colindex, const double *h_p_val, sparsh_rap_t *out,
                        int *nnz_coarse) {
    SP_TRY(ensure_init());
    SP_REQUIRE(nrow >= 0 && ncoarse >= 0 && h_rowptr && h_p_rowptr && out && nnz_coarse, "bad arguments");
    NvtxRange nvtx("sparsh:galerkin-rap");
    cudaStream_t st = ctx().stream;
    *out = nullptr;
    *nnz_coarse = -1;
    DevCsr A, P, R, AP;
    sparsh_rap_s *h = new sparsh_rap_s();
    bool ok1 = true, ok2 = true;
    int rc = upload_csr(nrow, nrow, h_rowptr, h_colindex, h_val, &A, st);
    if (rc == SPARSH_OK) rc = upload_csr(nrow, ncoarse, h_p_rowptr, h_p_colindex, h_p_val, &P, st);
    if (rc == SPARSH_OK) rc = spgemm(A, P, &AP, &ok1, st);                 // A P
    if (rc == SPARSH_OK && ok1) rc = transpose(P, &R, st);                 // R = P^T
    if (rc == SPARSH_OK && ok1) rc = spgemm(R, AP, &h->C, &ok2, st);       // R (A P)
    if (rc == SPARSH_OK && ok1 && ok2) {
        if (h->C.nrow > 0) sort_rows_kernel<<<(h->C.nrow + 255) / 256, 256, 0, st>>>(h->C.nrow, h->C.rp, h->C.ci, h->C.v);
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            set_error("device Galerkin product failed");
            rc = SPARSH_ERR_CUDA;
        }
    }
    A.release();
    P.release();
    R.release();
    AP.release();
    if (rc != SPARSH_OK || !ok1 || !ok2) {
        h->C.release();
        delete h;
        return rc;  // SPARSH_OK with *out == NULL: a product row exceeds the per-thread list; the caller uses its host product
    }
    *nnz_coarse = h->C.nnz;
    *out = h;
    return SPARSH_OK;
}

int sparsh_rap_fetch(sparsh_rap_t h, int *h_rowptr, int *h_colindex, double *h_val) {
    SP_REQUIRE(h != nullptr && h_rowptr != nullptr, "bad arguments");
    cudaStream_t st = ctx().stream;
    SP_CUDA(cudaMemcpyAsync(h_rowptr, h->C.rp, sizeof(int) * ((size_t)h->C.nrow + 1), cudaMemcpyDeviceToHost, st));
    if (h->C.nnz > 0) {
        SP_CUDA(cudaMemcpyAsync(h_colindex, h->C.ci, sizeof(int) * (size_t)h->C.nnz, cudaMemcpyDeviceToHost, st));
        SP_CUDA(cudaMemcpyAsync(h_val, h->C.v, sizeof(double) * (size_t)h->C.nnz, cudaMemcpyDeviceToHost, st));
    }
    SP_CUDA(cudaStreamSynchronize(st));
    return SPARSH_OK;
}

int sparsh_rap_destroy(sparsh_rap_t h) {
    if (!h) return SPARSH_OK;
    h->C.release();
    delete h;
    return SPARSH_OK;
}

}  // extern "C"
