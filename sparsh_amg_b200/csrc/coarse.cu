// coarse.cu — K7 of SURVEY §8a: the coarsest-level solve, entirely on the device.
//
// The reference factors the coarsest matrix with host PARDISO (src/AMG_coarse_level_solver.cpp:9-62) and, in both GPU
// variants, ships B down and X up over PCIe every cycle (src/AMG_gpu_phases_2.cu:131-141,192-203).  Here the coarsest
// level (<= limit_upper = 4000 rows, include/AMG.hpp:19) is inverted ONCE at setup by in-place Gauss-Jordan with
// partial pivoting on the device (fp64), and every solve is one dense GEMV  x = A^{-1} b  — 8 n^2 bytes, HBM/L2-bound,
// no host round trip, deterministic.
#include <vector>

#include "internal.cuh"

namespace sparsh {

constexpr int GJ_T = 1024;

// Step k, phase 1 (one CTA): pivot search in column k (largest |a_ik|, i >= k, ties -> smallest i), row swap,
// then extraction of the pivot row and of column k.
__global__ void __launch_bounds__(GJ_T) gj_pivot_kernel(double *M, int n, int k, int *piv, double *prow, double *fcol) {
    __shared__ double sv[GJ_T / 32];
    __shared__ int si[GJ_T / 32];
    __shared__ int s_p;
    double best = -1.0;
    int bi = n;
    for (int i = k + threadIdx.x; i < n; i += GJ_T) {
        double a = fabs(M[(size_t)i * n + k]);
        if (a > best) {  // strictly greater: keeps the smallest index per thread
            best = a;
            bi = i;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, off);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) {
            best = ob;
            bi = oi;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sv[warp] = best;
        si[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
        best = sv[lane];
        bi = si[lane];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, off);
            int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob > best || (ob == best && oi < bi)) {
                best = ob;
                bi = oi;
            }
        }
        if (lane == 0) {
            s_p = bi;
            piv[k] = bi;
        }
    }
    __syncthreads();
    const int p = s_p;
    double *rk = M + (size_t)k * n, *rp = M + (size_t)p * n;
    for (int j = threadIdx.x; j < n; j += GJ_T) {
        double a = rk[j], b = rp[j];
        if (p != k) {
            rk[j] = b;
            rp[j] = a;
        }
        prow[j] = (p != k) ? b : a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += GJ_T) fcol[i] = M[(size_t)i * n + k];
}

// Step k, phase 2 (whole grid): in-place Gauss-Jordan update
//   row k   <- prow/pivot with a_kk = 1/pivot
//   row i!=k <- row i - f_i * (row k), with the k-th column treated as the unit vector
__global__ void __launch_bounds__(256) gj_update_kernel(double *M, int n, int k, const double *prow, const double *fcol) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int i0 = blockIdx.y * 8;
    if (j >= n) return;
    const double pinv = 1.0 / prow[k];
    const double pj = (j == k ? 1.0 : prow[j]) * pinv;
#pragma unroll
    for (int ii = 0; ii < 8; ii++) {
        const int i = i0 + ii;
        if (i >= n) break;
        double *m = M + (size_t)i * n + j;
        if (i == k) {
            *m = pj;
        } else {
            const double f = fcol[i];
            const double a = (j == k) ? 0.0 : *m;
            *m = a - f * pj;
        }
    }
}

// undo the row interchanges: columns swapped in reverse pivot order; one CTA per row, row staged in shared memory
__global__ void __launch_bounds__(256) gj_unscramble_kernel(double *M, int n, const int *piv) {
    extern __shared__ double srow[];
    double *row = M + (size_t)blockIdx.x * n;
    for (int j = threadIdx.x; j < n; j += 256) srow[j] = row[j];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = n - 1; k >= 0; k--) {
            const int p = piv[k];
            if (p != k) {
                double t = srow[k];
                srow[k] = srow[p];
                srow[p] = t;
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += 256) row[j] = srow[j];
}

// x = Ainv b: one warp per row, fixed lane-strided accumulation + shuffle tree (deterministic)
__global__ void __launch_bounds__(256) dense_gemv_kernel(const double *__restrict__ Minv, int n, const double *__restrict__ b, double *x) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *m = Minv + (size_t)row * n;
    double s0 = 0.0, s1 = 0.0;
    int j = lane * 2;
    if ((n & 1) == 0) {
        for (; j + 1 < n; j += 64) {
            const double2 a = *reinterpret_cast<const double2 *>(m + j);
            const double2 v = *reinterpret_cast<const double2 *>(b + j);
            s0 = fma(a.x, v.x, s0);
            s1 = fma(a.y, v.y, s1);
        }
    } else {
        for (j = lane; j < n; j += 32) s0 = fma(m[j], b[j], s0);
    }
    double s = s0 + s1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) x[row] = s;
}

int coarse_build_inverse(int n, const int *rp, const int *ci, const double *v, CoarseInverse *out) {
    Context &c = ctx();
    if (n > 16384) {
        set_error("coarsest level too large for the dense device solve (n > 16384): coarsen further (raise level1)");
        return SPARSH_ERR_INVALID;
    }
    out->n = n;
    if (n == 0) return SPARSH_OK;
    std::vector<double> dense((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++) dense[(size_t)i * n + ci[j]] += v[j];
    double *M = nullptr, *prow = nullptr, *fcol = nullptr;
    int *piv = nullptr;
    SP_CUDA(cudaMalloc(&M, sizeof(double) * (size_t)n * n));
    SP_CUDA(cudaMalloc(&prow, sizeof(double) * (size_t)n));
    SP_CUDA(cudaMalloc(&fcol, sizeof(double) * (size_t)n));
    SP_CUDA(cudaMalloc(&piv, sizeof(int) * (size_t)n));
    SP_CUDA(cudaMemcpyAsync(M, dense.data(), sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, c.stream));
    dim3 ugrid((n + 255) / 256, (n + 7) / 8);
    for (int k = 0; k < n; k++) {
        gj_pivot_kernel<<<1, GJ_T, 0, c.stream>>>(M, n, k, piv, prow, fcol);
        gj_update_kernel<<<ugrid, 256, 0, c.stream>>>(M, n, k, prow, fcol);
    }
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(gj_unscramble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        attr = true;
    }
    gj_unscramble_kernel<<<n, 256, sizeof(double) * (size_t)n, c.stream>>>(M, n, piv);
    SP_CUDA(cudaGetLastError());
    SP_CUDA(cudaStreamSynchronize(c.stream));
    cudaFree(prow);
    cudaFree(fcol);
    cudaFree(piv);
    out->inv = M;
    return SPARSH_OK;
}

int coarse_apply(const CoarseInverse &ci, const double *b, double *x) {
    if (ci.n == 0) return SPARSH_OK;
    dense_gemv_kernel<<<(ci.n + 7) / 8, 256, 0, ctx().stream>>>(ci.inv, ci.n, b, x);
    count_launch();
    SP_CUDA(cudaGetLastError());
    return SPARSH_OK;
}

void coarse_free(CoarseInverse *ci) {
    if (ci->inv) cudaFree(ci->inv);
    ci->inv = nullptr;
    ci->n = 0;
}

}  // namespace sparsh
