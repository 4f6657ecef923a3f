// tail.cu — the small levels of a V-cycle as ONE kernel (cooperative grid, or one thread-block cluster).
//
// Below ~10^5 rows a level's operator applications are a few microseconds of work each, and the gap between two graph
// nodes (drain, launch, CTA scheduling: ~3 us) costs as much as the work: on the 14-level hierarchy of 3D Poisson 256^3
// the levels with at most 131 072 rows are ~120 of the ~250 kernels of a cycle and ~0.7 ms of a 4.8 ms PCG iteration
// on one GPU — and exactly the same 0.7 ms on 8 GPUs, where these levels are replicated and the iteration is 2 ms.
// Here that whole bottom of the cycle — pre-smoothing, residual, restriction down to the coarsest level, the dense
// coarse solve, prolongation-correction and post-smoothing back up — is a list of operations executed by one grid of
// co-resident CTAs (cooperative launch, one CTA per SM) with a grid barrier between two operations instead of a kernel
// boundary.  Every operation is the thread-per-row (dense solve: warp-per-row) evaluation of the kernels it replaces
// with the same entry order and the same unfused arithmetic, so the cycle's result is bit-identical with or without
// it.  Vectors written by one operation and read by the next are loaded through L2 (ld.global.cg): L1 is not coherent
// across the CTAs of a running grid.
#include <cooperative_groups.h>

#include <cstdlib>
#include <vector>

#include "hierarchy.cuh"

namespace sparsh {

namespace {

enum TailKind { T_JACOBI_ZERO = 0, T_JACOBI = 1, T_RESID = 2, T_SPMV = 3, T_PROLONG = 4, T_GEMV = 5, T_FILL0 = 6 };

struct TailOp {
    int kind, n;
    const int *rp, *ci;
    const double *val, *diag;  // CSR (T_GEMV: val = dense inverse, n x n row-major)
    const double *x, *b;       // gathered vector / right-hand side
    double *y;                 // result (T_PROLONG: read-modify-write)
    double omega;
};

constexpr int TAIL_T = 1024;

__device__ __forceinline__ double row_sum(const TailOp &op, int row) {
    double s = 0.0;
    for (int k = op.rp[row]; k < op.rp[row + 1]; k++) s = __dadd_rn(s, __dmul_rn(__ldg(op.val + k), __ldcg(op.x + __ldg(op.ci + k))));
    return s;
}

// CLUSTER = false: cooperative launch, one CTA per SM, grid barrier between operations.
// CLUSTER = true : ONE thread-block cluster of 16 CTAs (non-portable size) and the hardware cluster barrier
//                  (barrier.cluster, ~0.2 us) between operations — for the levels small enough for 16 SMs.
template <bool CLUSTER>
__global__ void __launch_bounds__(TAIL_T) tail_cycle_kernel(const TailOp *__restrict__ ops, int nops) {
    const int tid = blockIdx.x * TAIL_T + threadIdx.x, nthreads = gridDim.x * TAIL_T;
    for (int o = 0; o < nops; o++) {
        const TailOp op = ops[o];
        if (op.kind == T_GEMV) {
            // x = Ainv b, warp per row: the arithmetic of dense_gemv_kernel (coarse.cu), lane for lane
            const int lane = threadIdx.x & 31, warp = tid >> 5, nwarps = nthreads >> 5;
            const bool vec = (op.n & 1) == 0 && (reinterpret_cast<uintptr_t>(op.b) & 15) == 0;
            for (int row = warp; row < op.n; row += nwarps) {
                const double *m = op.val + (size_t)row * op.n;
                double s0 = 0.0, s1 = 0.0;
                if (vec) {
                    for (int j = lane * 2; j + 1 < op.n; j += 64) {
                        const double2 a = *reinterpret_cast<const double2 *>(m + j);
                        const double v0 = __ldcg(op.b + j), v1 = __ldcg(op.b + j + 1);
                        s0 = fma(a.x, v0, s0);
                        s1 = fma(a.y, v1, s1);
                    }
                } else {
                    for (int j = lane; j < op.n; j += 32) s0 = fma(m[j], __ldcg(op.b + j), s0);
                }
                double s = s0 + s1;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (lane == 0) op.y[row] = s;
            }
        } else {
            for (int row = tid; row < op.n; row += nthreads) {
                switch (op.kind) {
                    case T_JACOBI_ZERO:  // first sweep from a zero guess: x = (omega*b)/d   (jacobi_zero_kernel)
                        op.y[row] = __ddiv_rn(__dmul_rn(op.omega, __ldcg(op.b + row)), op.diag[row]);
                        break;
                    case T_JACOBI: {     // x' = x + (omega*(b - A x))/d                     (EPI_JACOBI)
                        const double h = __dsub_rn(__ldcg(op.b + row), row_sum(op, row));
                        op.y[row] = __dadd_rn(__ldcg(op.x + row), __ddiv_rn(__dmul_rn(op.omega, h), op.diag[row]));
                        break;
                    }
                    case T_RESID:        // r = b - A x                                        (EPI_RESID)
                        op.y[row] = __dsub_rn(__ldcg(op.b + row), row_sum(op, row));
                        break;
                    case T_SPMV:         // b_c = R r                                          (EPI_SPMV)
                        op.y[row] = row_sum(op, row);
                        break;
                    case T_PROLONG:      // x_f = (P x_c) + x_f                                (EPI_PROLONG)
                        op.y[row] = __dadd_rn(row_sum(op, row), __ldcg(op.y + row));
                        break;
                    default:             // T_FILL0
                        op.y[row] = 0.0;
                }
            }
        }
        if (CLUSTER)
            cooperative_groups::this_cluster().sync();  // release/acquire at cluster scope: the CTAs' global writes are ordered
        else
            cooperative_groups::this_grid().sync();
    }
}

}  // namespace

// SPARSH_TAIL_MODE: 0 (default) every level launches its own kernels; 1 the levels with at most SPARSH_TAIL_ROWS rows
// (default 131072) run as one cooperative kernel; 2 the levels with at most SPARSH_TAIL_ROWS rows (default 32768) run as one
// 16-CTA cluster.  Read when a hierarchy takes its decision (once per hierarchy), so tests can compare the variants.
// MEASURED (B200, 256^3, profiles/r02o_*): mode 1 removes 45 % of the launches of a solve and is bit-identical, but is no
// faster — 0.1457 s against 0.1443 s on one GPU, 0.0982 s against 0.0963 s on two: inside a CUDA graph a kernel boundary
// costs about what a grid barrier over 128 CTAs costs.  Hence off by default.
static int tail_mode() {
    const char *e = getenv("SPARSH_TAIL_MODE");
    return e ? atoi(e) : 0;
}
static int tail_rows() {
    const char *e = getenv("SPARSH_TAIL_ROWS");
    if (e) return atoi(e);
    return tail_mode() == 2 ? 32768 : 131072;
}

// first level of the fused bottom (>= 1: level 0 works on the caller's vectors), or -1
int tail_level(sparsh_hierarchy_s *h) {
    if (h->tail_state == 2) return -1;  // tried and refused (no cooperative launch, capture failure): classical launches
    if (h->tail_state == 1) return h->tail_first;
    h->tail_state = 2;
    const int L = (int)h->lev.size() - 1;
    if (tail_mode() <= 0 || tail_rows() <= 0 || h->prm.smoother != 0 || L < 1 || h->coarse.n == 0) return -1;
    int first = L;  // the coarsest level alone is not worth it: need at least one smoothed level
    while (first > 1 && h->lev[first - 1].n <= tail_rows()) first--;
    if (first >= L) return -1;
    // only CSR kernels that evaluate a row left to right are reproduced bit for bit: no vector-family matrices below
    for (int l = first; l < L; l++)
        if (h->lev[l].A->kind == KIND_VECTOR || h->lev[l].P->kind == KIND_VECTOR || h->lev[l].R->kind == KIND_VECTOR) return -1;
    const bool cluster = tail_mode() == 2;
    int coop = 0;
    if (!cluster && (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx().device) != cudaSuccess || !coop)) return -1;
    if (cluster && cudaFuncSetAttribute(tail_cycle_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    // the program: the operations enqueue_vcycle would launch for levels first..L, with its ping-pong bookkeeping
    std::vector<TailOp> prog;
    auto mat = [](TailOp &op, const sparsh_matrix_s *M) {
        op.n = M->nrow;
        op.rp = M->rowptr;
        op.ci = M->col;
        op.val = M->val;
    };
    std::vector<double *> X(L + 1), T(L + 1);
    for (int l = first; l <= L; l++) {
        X[l] = h->lev[l].xbuf;
        T[l] = h->lev[l].tbuf;
    }
    const double omega = h->prm.omega;
    auto smooth = [&](int l, int sweeps, bool zero) {
        Level &F = h->lev[l];
        if (sweeps == 0) {
            if (zero) {
                TailOp op = {};
                op.kind = T_FILL0;
                op.n = F.n;
                op.y = X[l];
                prog.push_back(op);
            }
            return;
        }
        for (int s = 0; s < sweeps; s++) {
            TailOp op = {};
            mat(op, F.A);
            op.diag = F.A->diag;
            op.b = F.bbuf;
            op.omega = omega;
            op.y = T[l];
            if (s == 0 && zero) {
                op.kind = T_JACOBI_ZERO;
            } else {
                op.kind = T_JACOBI;
                op.x = X[l];
            }
            prog.push_back(op);
            std::swap(X[l], T[l]);
        }
    };
    for (int l = first; l < L; l++) {
        Level &F = h->lev[l];
        smooth(l, h->prm.pre_sweeps, true);
        TailOp r = {};
        mat(r, F.A);
        r.kind = T_RESID;
        r.x = X[l];
        r.b = F.bbuf;
        r.y = F.rbuf;
        prog.push_back(r);
        TailOp t = {};
        mat(t, F.R);
        t.kind = T_SPMV;
        t.x = F.rbuf;
        t.y = h->lev[l + 1].bbuf;
        prog.push_back(t);
    }
    {
        TailOp g = {};
        g.kind = T_GEMV;
        g.n = h->coarse.n;
        g.val = h->coarse.inv;
        g.b = h->lev[L].bbuf;
        g.y = X[L];
        prog.push_back(g);
    }
    for (int l = L; l > first; l--) {
        Level &F = h->lev[l - 1];
        TailOp p = {};
        mat(p, F.P);
        p.kind = T_PROLONG;
        p.x = X[l];
        p.y = X[l - 1];
        prog.push_back(p);
        smooth(l - 1, h->prm.post_sweeps, false);
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tail_cycle_kernel<false>, TAIL_T, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return -1;
    }
    // enough CTAs for a row per thread on the largest fused level, never more than one per SM (all co-resident);
    // cluster mode: exactly one cluster of 16 CTAs
    const int want = (h->lev[first].n + TAIL_T - 1) / TAIL_T;
    h->tail_grid = cluster ? -16 : std::max(1, std::min(want, ctx().sm_count));
    if (cudaMalloc(&h->tail_prog, sizeof(TailOp) * prog.size()) != cudaSuccess ||
        cudaMemcpy(h->tail_prog, prog.data(), sizeof(TailOp) * prog.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    h->tail_nops = (int)prog.size();
    h->tail_first = first;
    h->tail_x = X[first];  // where the correction of level `first` ends up after its post-smoothing
    h->tail_state = 1;
    return first;
}

// enqueue the fused bottom of the cycle (levels tail_first..L); B[tail_first] = lev[tail_first].bbuf must be in place
int enqueue_tail(sparsh_hierarchy_s *h) {
    Context &c = ctx();
    const TailOp *prog = static_cast<const TailOp *>(h->tail_prog);
    int nops = h->tail_nops;
    if (h->tail_grid < 0) {  // one cluster of -tail_grid CTAs: an ordinary (capturable) launch with a cluster dimension
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(-h->tail_grid);
        cfg.blockDim = dim3(TAIL_T);
        cfg.stream = c.stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)(-h->tail_grid);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SP_CUDA(cudaLaunchKernelEx(&cfg, tail_cycle_kernel<true>, prog, nops));
    } else {
        void *args[] = {(void *)&prog, (void *)&nops};
        SP_CUDA(cudaLaunchCooperativeKernel((const void *)tail_cycle_kernel<false>, dim3(h->tail_grid), dim3(TAIL_T), args, 0, c.stream));
    }
    count_launch();
    return SPARSH_OK;
}

void tail_free(sparsh_hierarchy_s *h) {
    cudaFree(h->tail_prog);
    h->tail_prog = nullptr;
    h->tail_state = 0;
}

}  // namespace sparsh
