// spmv_pattern.cu — csr-pattern8, lean variant: the row-sum kernel for matrices whose rows almost all repeat ONE pattern
// (the interior stencil row: 97.7 % of the rows of 3D Poisson 256^3 and of every level of its HEM hierarchy).
//
// What the first pattern kernel (spmv.cu: csr_pattern_kernel) pays for and this one does not:
//  * two dependent memory round trips per tile (pattern byte -> table look-up -> gathers).  Pattern 0 travels BY VALUE
//    in the kernel parameters (Pat0, internal.cuh), so its offsets are constant-bank operands known before anything is
//    loaded: every thread issues the pattern byte, b, x_i AND all gathers of pattern 0 for all its rows at once —
//    one round trip — and only looks at the pattern byte when the data is there.  Rows that turn out to carry another
//    pattern (boundary rows, a few per cent) are redone from the shared-memory table, escapes from the CSR arrays;
//  * ~250 instructions per row of predicated 8-wide steps, selects and shared-memory look-ups (ncu on the dict kernel:
//    issue slots 65 % busy, ALU the top pipe).  The fast path is straight-line code: per entry one integer add, one
//    address computation, one LDG, one DMUL and one DADD whose second operands come from the constant bank;
//  * the table copy into shared memory: done only by CTAs that meet a row with another pattern.
// Same entries, same order, same unfused arithmetic as the CSR kernels: results are bit-identical.
#include "spmv_common.cuh"

namespace sparsh {

namespace {

template <bool COHERENT, bool DIST, bool SAFE>
__device__ __forceinline__ double gather(const double *x, int c, int ncol, int halo_begin) {
    if (!SAFE) c = min(max(c, 0), ncol - 1);  // speculative address of a row that may not carry pattern 0
    return load_xd<COHERENT, DIST>(x, c, halo_begin);
}

template <int THREADS, int RPT, int LEN0, int EPI, bool DIST>
__global__ void __launch_bounds__(THREADS)
    csr_pat2_kernel(CsrView A, PatView P, const __grid_constant__ Pat0 Z, const double *x, double *y, EpiArgs args,
                    RowRange rr, double *partials, HaloSync hs) {
    constexpr bool NEEDS_D = (EPI == EPI_JACOBI || EPI == EPI_SOR);
    constexpr bool COH = EpiTraits<EPI>::coherent_x;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sval = reinterpret_cast<double *>(smem_raw);   // n_ent values       } only filled by CTAs that own a row
    double *sdiag = sval + P.n_ent;                        // n_pat              } with another pattern than 0
    int *soff = reinterpret_cast<int *>(sdiag + P.n_pat);  // n_ent offsets
    int *sstart = soff + P.n_ent;                          // n_pat + 1

    const int tid = threadIdx.x;
    int r0, row_end;
    block_rows(rr, THREADS * RPT, r0, row_end);
    const int nrows = min(THREADS * RPT, row_end - r0);
    HaloTurn hs_turn;
    if (DIST) hs_turn = halo_wait(hs);  // multi-GPU: the halo slices have landed before any (speculative) gather
    const bool strip_cta = !DIST || (int)blockIdx.x < hs.nstrip;
    // do the pattern-0 gathers of EVERY row of this tile stay inside the vector?  (CTA-uniform; false only at the ends)
    const bool safe = r0 + Z.lo >= 0 && r0 + THREADS * RPT - 1 + Z.hi < A.ncol;
    const bool table_d = P.use_pdiag != 0;

    int pid[RPT];
    EpiRegs e[RPT];
    double g[RPT][LEN0];
#pragma unroll
    for (int s = 0; s < RPT; s++) {
        pid[s] = -1;
        if (s * THREADS + tid < nrows) {
            const int row = r0 + s * THREADS + tid;
            pid[s] = __ldg(P.pat + row);
            e[s] = epi_load<EPI, false>(args, y, row);
            if (NEEDS_D && !table_d) e[s].d = args.d[row];
            if (safe) {
#pragma unroll
                for (int k = 0; k < LEN0; k++) g[s][k] = gather<COH, DIST, true>(x, row + Z.off[k], A.ncol, hs.halo_begin);
            } else {
#pragma unroll
                for (int k = 0; k < LEN0; k++) g[s][k] = gather<COH, DIST, false>(x, row + Z.off[k], A.ncol, hs.halo_begin);
            }
        }
    }

    double contrib = 0.0;
    bool other = false;
#pragma unroll
    for (int s = 0; s < RPT; s++) {
        if (pid[s] == 0) {
            const int row = r0 + s * THREADS + tid;
            double sum = 0.0;
#pragma unroll
            for (int k = 0; k < LEN0; k++) sum = __dadd_rn(sum, __dmul_rn(Z.val[k], g[s][k]));
            if (NEEDS_D && table_d) e[s].d = Z.diag;
            // reduction contributions: pattern-0 rows in row order, then the others in row order — a fixed tree
            contrib = __dadd_rn(contrib, epi_store<EPI, DIST>(args, e[s], sum, y, row, strip_cta));
        } else if (pid[s] > 0) {
            other = true;
        }
    }

    if (__syncthreads_or(other)) {  // somebody in this CTA met another pattern: bring the table in
        for (int i = tid; i < P.n_ent; i += THREADS) {
            const int4 q = __ldg(reinterpret_cast<const int4 *>(P.ent) + i);
            sval[i] = __hiloint2double(q.y, q.x);
            soff[i] = q.z;
        }
        for (int i = tid; i < P.n_pat; i += THREADS) sdiag[i] = __ldg(P.pdiag + i);
        for (int i = tid; i <= P.n_pat; i += THREADS) sstart[i] = __ldg(P.start + i);
        __syncthreads();
#pragma unroll
        for (int s = 0; s < RPT; s++) {
            if (pid[s] > 0) {
                const int row = r0 + s * THREADS + tid;
                double sum = 0.0;
                if (pid[s] != PAT_ESCAPE) {
                    const int st = sstart[pid[s]], en = sstart[pid[s] + 1];
#pragma unroll 1  // rare path: small code, the instruction cache belongs to the fast path
                    for (int k = st; k < en; k++)
                        sum = __dadd_rn(sum, __dmul_rn(sval[k], load_xd<COH, DIST>(x, row + soff[k], hs.halo_begin)));
                    if (NEEDS_D && table_d) e[s].d = sdiag[pid[s]];
                } else {
                    const int lo = A.rowptr[row], hi = A.rowptr[row + 1];
#pragma unroll 1
                    for (int k = lo; k < hi; k++)
                        sum = __dadd_rn(sum, __dmul_rn(__ldg(A.val + k), load_xd<COH, DIST>(x, __ldg(A.col + k), hs.halo_begin)));
                    if (NEEDS_D) e[s].d = args.d[row];
                }
                contrib = __dadd_rn(contrib, epi_store<EPI, DIST>(args, e[s], sum, y, row, strip_cta));
            }
        }
    }
    if (pushes<EPI, DIST>(args, strip_cta)) {
#pragma unroll 1
        for (int s = 0; s < RPT; s++)
            if (pid[s] >= 0) fused_push<EPI, DIST>(args, y, r0 + s * THREADS + tid);
    }
    if (EpiTraits<EPI>::reduces) block_partial<THREADS>(contrib, partials);
    if (DIST) halo_done(hs, hs_turn);
}

template <int THREADS, int RPT, int LEN0, int EPI>
int launch_cfg(const sparsh_matrix_s *A, const double *x, double *y, const EpiArgs &args, LaunchDesc d) {
    Context &c = ctx();
    const int grid = grid_for(d, THREADS * RPT);
    if (grid > RED_MAX_BLOCKS) {
        set_error("matrix too large for the reduction workspace");
        return SPARSH_ERR_INVALID;
    }
    const PatView P = A->pattern(args.d != nullptr && args.d == A->diag);
    const size_t smem = (size_t)A->n_pent * 12 + (size_t)A->n_pat * 8 + (size_t)(A->n_pat + 1) * 4;  // <= 27 KB
    if (d.dist)
        csr_pat2_kernel<THREADS, RPT, LEN0, EPI, true><<<grid, THREADS, smem, c.stream>>>(A->view(), P, A->pat0, x, y, args, d.rr, c.partials, d.hs);
    else
        csr_pat2_kernel<THREADS, RPT, LEN0, EPI, false><<<grid, THREADS, smem, c.stream>>>(A->view(), P, A->pat0, x, y, args, d.rr, c.partials, d.hs);
    count_launch();
    SP_CUDA(cudaGetLastError());
    if (EpiTraits<EPI>::reduces) return launch_finalize_partials(grid, args.red_out);
    return SPARSH_OK;
}

// rows per thread (SPARSH_PAT2_RPT = 1 | 2 | 4 overrides the default for experiments)
int lean_rpt() {
    static const int v = [] {
        const char *e = getenv("SPARSH_PAT2_RPT");
        const int r = e ? atoi(e) : 2;
        return (r == 1 || r == 2 || r == 4) ? r : 2;
    }();
    return v;
}

template <int LEN0, int EPI>
int launch_len(const sparsh_matrix_s *A, const double *x, double *y, const EpiArgs &args, const LaunchDesc &d) {
    const int rpt = lean_rpt();
    if (A->threads == 128) {
        if (rpt == 1) return launch_cfg<128, 1, LEN0, EPI>(A, x, y, args, d);
        if (rpt == 4) return launch_cfg<128, 4, LEN0, EPI>(A, x, y, args, d);
        return launch_cfg<128, 2, LEN0, EPI>(A, x, y, args, d);
    }
    if (rpt == 1) return launch_cfg<256, 1, LEN0, EPI>(A, x, y, args, d);
    if (rpt == 4) return launch_cfg<256, 4, LEN0, EPI>(A, x, y, args, d);
    return launch_cfg<256, 2, LEN0, EPI>(A, x, y, args, d);
}

template <int EPI>
int launch_epi(const sparsh_matrix_s *A, const double *x, double *y, const EpiArgs &args, const LaunchDesc &d) {
    switch (A->pat0.len) {
        case 5:
            return launch_len<5, EPI>(A, x, y, args, d);
        default:
            return launch_len<7, EPI>(A, x, y, args, d);
    }
}

}  // namespace

// stencils whose interior row has 5 or 7 entries (2D/3D Poisson-like operators and their semi-coarsened Galerkin
// levels), as long as that row is the majority; SPARSH_PAT2=0 keeps the first pattern kernel (A/B measurements)
bool pattern_lean_applies(const sparsh_matrix_s *A) {
    static const bool enabled = [] {
        const char *e = getenv("SPARSH_PAT2");
        return !(e && atoi(e) == 0);
    }();
    const int len = A->pat0.len;
    return enabled && A->has_pat && (len == 5 || len == 7) && A->pat0.cover >= 0.5;
}

int launch_pattern_lean(const sparsh_matrix_s *A, int epi, const double *x, double *y, const EpiArgs &args, LaunchDesc d) {
    switch (epi) {
        case EPI_SPMV:
            return launch_epi<EPI_SPMV>(A, x, y, args, d);
        case EPI_RESID:
            return launch_epi<EPI_RESID>(A, x, y, args, d);
        case EPI_JACOBI:
            return launch_epi<EPI_JACOBI>(A, x, y, args, d);
        case EPI_PROLONG:
            return launch_epi<EPI_PROLONG>(A, x, y, args, d);
        case EPI_SOR:
            return launch_epi<EPI_SOR>(A, x, y, args, d);
        case EPI_SPMV_DOT:
            return launch_epi<EPI_SPMV_DOT>(A, x, y, args, d);
        case EPI_RESNORM:
            return launch_epi<EPI_RESNORM>(A, x, y, args, d);
    }
    set_error("unknown epilogue");
    return SPARSH_ERR_INVALID;
}

}  // namespace sparsh
