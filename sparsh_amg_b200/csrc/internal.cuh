// internal.cuh — shared declarations of libsparsh_b200 (never includes the reference's AMG.hpp: its macros
// `th`, `omega`, ... break CUDA headers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include <nvtx3/nvToolsExt.h>  // header-only (dlopen's the tools library when a profiler is attached; a no-op otherwise)

#include "../../include/sparsh_b200.h"

namespace sparsh {

// ---- error plumbing -------------------------------------------------------------------------------
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SP_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) return ::sparsh::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define SP_TRY(call)                      \
    do {                                  \
        int rc__ = (call);                \
        if (rc__ != SPARSH_OK) return rc__; \
    } while (0)
#define SP_REQUIRE(cond, msg)             \
    do {                                  \
        if (!(cond)) {                    \
            ::sparsh::set_error(msg);     \
            return SPARSH_ERR_INVALID;    \
        }                                 \
    } while (0)

// ---- NVTX ranges (the reference links nvToolsExt without using it, CMakeLists.txt:42): phases of the solve show up by
// name in ncu / nsys timelines: sparsh:upload, sparsh:coarse-inverse, sparsh:pcg, sparsh:bicgstab, sparsh:amg-solve, ...
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// An iteration is ~250 dependent kernels, most of them a few microseconds long on the small levels.  Launched with the
// programmatic-stream-serialisation attribute, kernel k+1 is scheduled as soon as every CTA of kernel k has passed its
// griddepcontrol.launch_dependents (the first thing our kernels do after their own griddepcontrol.wait; SASS: ACQBULK /
// PREEXIT) and then blocks in griddepcontrol.wait until kernel k has completed and its memory operations are visible.
// A kernel launched WITHOUT the attribute passes the wait at once, and a predecessor that never triggers releases its
// dependents when it exits, so mixing is safe; CUDA graphs capture the attribute as programmatic dependency edges.
// MEASURED (B200, 256^3, profiles/r02n_*): no gain — 0.1442 s against 0.1446 s on one GPU, 0.0991 s against 0.0961 s on
// two (inside a CUDA graph the node-to-node gap is already small, and a dependent grid that is resident while it waits
// takes SM slots from the strips of the other stream).  Hence OFF by default; SPARSH_PDL=1 turns it on.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();
void pdl_disable();  // a launch with the attribute was refused (old driver): everything is launched classically from then on
template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
        if (e != cudaErrorNotSupported) return e;
        cudaGetLastError();
        pdl_disable();
        cfg.attrs = nullptr;
        cfg.numAttrs = 0;
    }
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- runtime context --------------------------------------------------------------------------------
struct Context {
    bool ready = false;
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    long long launches = 0;  // kernels launched (direct) + kernel nodes replayed (graphs)
    bool capturing = false;  // launches issued while capturing are counted per replay instead
    long long captured = 0;
    // reduction workspace (two-stage deterministic reductions)
    double *partials = nullptr;  // RED_MAX_BLOCKS * RED_MAX_VALUES doubles
    unsigned int *ticket = nullptr;
    double *d_scalar = nullptr;  // small device scratch for scalar results (16 doubles)
    double *h_scalar = nullptr;  // pinned mirror
};
Context &ctx();
int ensure_init();
inline void count_launch() {
    Context &c = ctx();
    if (c.capturing)
        c.captured++;
    else
        c.launches++;
}

constexpr int RED_MAX_BLOCKS = 1 << 20;  // partial sums per reduced value (2^28 rows at 256 rows per block)
constexpr int RED_MAX_VALUES = 4;

// ---- device CSR ---------------------------------------------------------------------------------------
enum KernelKind { KIND_SCALAR = 0, KIND_STREAM = 1, KIND_VECTOR = 2, KIND_DICT = 3, KIND_PATTERN = 4 };

// csr-dict16: lossless re-encoding of a CSR whose entries take at most 256 distinct values and 256 distinct column
// offsets (col - row) — stencil-like matrices and their Galerkin coarsenings.  One 16-bit code per entry
// (value id << 8 | offset id) replaces the 8-byte value and the 4-byte column: 2 B/nnz instead of 12.
struct DictView {
    const unsigned short *__restrict__ code;
    const double *__restrict__ val;  // n_val distinct values
    const int *__restrict__ off;     // n_off distinct col-row offsets
    int n_val, n_off;
};

// csr-pattern8: one byte per ROW.  Rows of a stencil matrix (and of its Galerkin coarsenings) repeat a handful of
// patterns — the ordered list of (col - row offset, value) pairs; the 7-point Poisson hierarchy has 27 per level.
// pat[i] < 255 selects a pattern of the table (entries start[p]..start[p+1]); pat[i] == 255 is the escape: that row is
// evaluated from the plain CSR arrays, which stay resident.  Same entries, same order, same unfused arithmetic as the
// CSR kernels: results are bit-identical.
struct PatEntry {
    double v;
    int off;
    int pad;
};
struct PatView {
    const unsigned char *__restrict__ pat;  // nrow
    const PatEntry *__restrict__ ent;       // n_ent (+8 zero entries)
    const int *__restrict__ start;          // n_pat + 1
    const double *__restrict__ pdiag;       // n_pat: the row's diagonal as sp_matrix_fill_diagonal extracts it
    int n_pat, n_ent;
    int use_pdiag;  // the epilogue's d[] is this matrix' own diagonal: take it from the table
    int far_off;    // largest positive offset of the table (0: none): the only x line of a row that rows swept earlier
                    // have not pulled into L2 yet — prefetched while the pattern byte is still on its way
};
// x windows of the TMA-staged pattern kernel: a tile of PAT_TILE consecutive rows gathers x only from a few contiguous
// index ranges [r0 + lo_w, r0 + lo_w + len_w) — one per group of table offsets that lie within a tile length of each
// other — which one thread fetches with bulk copies.  win[k] tells which window serves table entry k.
constexpr int PAT_TILE = 512;
constexpr int PAT_MAX_WIN = 8;
constexpr int PAT_WIN_DOUBLES = 3072;  // sum of the window lengths that still leaves >= 6 tiles resident per SM
struct PatWindows {
    int nwin;
    int lo[PAT_MAX_WIN], len[PAT_MAX_WIN];  // len is even
    int total;                              // sum of (len + 2): doubles of shared memory
    int w0;                                 // window that contains offset 0 (-1: none)
    const unsigned char *__restrict__ win;  // n_ent
};
// Pattern 0 (the most frequent row: patterns are numbered by decreasing row count) travels BY VALUE in the kernel
// parameters, so the lean kernel reads its offsets and values as constant-bank operands: no shared-memory look-up, no
// register, and — because the offsets do not depend on anything loaded — every gather of a row can be issued before
// the row's pattern byte has arrived (csr_pat2_kernel in spmv.cu).
constexpr int PAT0_MAX = 32;
struct Pat0 {
    int len = 0;      // entries of pattern 0 (0: the lean kernel does not apply)
    int kdiag = -1;   // index of its offset-0 entry (-1: none)
    int lo = 0, hi = 0;  // smallest / largest offset
    double diag = 0.0;   // the tabulated diagonal (pdiag[0])
    double cover = 0.0;  // fraction of the rows that carry pattern 0
    int off[PAT0_MAX] = {};
    double val[PAT0_MAX] = {};
};
constexpr int PAT_ESCAPE = 255;    // pattern id of an escape row
constexpr int PAT_MAX_ENT = 2048;  // table entries over all patterns (32 KB)
constexpr int PAT_MAX_ROW = 64;    // longer rows are never tabulated

struct CsrView {
    int nrow, ncol, nnz;
    const int *__restrict__ rowptr;
    const int *__restrict__ col;
    const double *__restrict__ val;
};

}  // namespace sparsh

struct sparsh_matrix_s {
    int nrow = 0, ncol = 0, nnz = 0;
    int *rowptr = nullptr;  // nrow+1 (+ padding)
    int *col = nullptr;     // nnz padded to a multiple of 4, +8
    double *val = nullptr;
    double *diag = nullptr;  // nrow, or null for rectangular operators
    // kernel selection (made at upload from host-side row statistics)
    int kind = sparsh::KIND_VECTOR;
    int threads = 256;  // stream/scalar: rows per CTA == threads per CTA
    bool threads_forced = false;  // set by sparsh_matrix_force_kernel: the caller's launch shape is honoured as given
    int lanes = 8;      // vector: lanes per row
    int max_row = 0;
    double mean_row = 0.0;
    int win128 = 0, win256 = 0;  // max nnz over any window of 128 / 256 consecutive rows
    int smem_bytes = 0;          // dynamic shared memory of the stream kernel
    // csr-dict16 twin (present when the dictionaries fit; see DictView)
    unsigned short *code = nullptr;
    double *dict_val = nullptr;
    int *dict_off = nullptr;
    int n_dval = 0, n_doff = 0;
    bool has_dict = false;
    // csr-pattern8 twin (see PatView)
    unsigned char *pat = nullptr;
    sparsh::PatEntry *pat_ent = nullptr;
    int *pat_start = nullptr;
    double *pat_diag = nullptr;
    int n_pat = 0, n_pent = 0, n_escape = 0, pat_far = 0;
    bool has_pat = false;
    sparsh::Pat0 pat0;                    // len == 0: the lean (speculative-gather) variant does not apply
    sparsh::PatWindows pat_windows = {};  // nwin == 0: the TMA-staged variant does not apply
    unsigned char *pat_win = nullptr;
    sparsh::PatView pattern(bool use_pdiag) const {
        return sparsh::PatView{pat, pat_ent, pat_start, pat_diag, n_pat, n_pent, use_pdiag ? 1 : 0, pat_far};
    }
    sparsh::CsrView view() const { return sparsh::CsrView{nrow, ncol, nnz, rowptr, col, val}; }
    sparsh::DictView dict() const { return sparsh::DictView{code, dict_val, dict_off, n_dval, n_doff}; }
};

namespace sparsh {

// ---- SpMV family (spmv.cu) ----------------------------------------------------------------------------
// y_i = epilogue(sum_j a_ij x_j) over rows [row_begin,row_end)
enum EpiKind {
    EPI_SPMV = 0,      // y = s
    EPI_RESID = 1,     // y = b - s
    EPI_JACOBI = 2,    // y = xi + (omega*(b - s))/d
    EPI_PROLONG = 3,   // y = s + y
    EPI_SOR = 4,       // y = y - (omega*(s - b))/d              (in place, one colour)
    EPI_SPMV_DOT = 5,  // y = s, reduce x_i*s
    EPI_RESNORM = 6    // reduce (s - b)^2, nothing stored
};
struct EpiArgs {
    const double *b = nullptr;
    const double *xi = nullptr;  // the "own" x entry stream (same array as the gathered x for Jacobi)
    const double *d = nullptr;
    double omega = 0.0;
    double *red_out = nullptr;  // device scalar receiving the reduction (EPI_SPMV_DOT / EPI_RESNORM)
    // multi-GPU, EPI_JACOBI only: rows whose new value a neighbour needs store it ALSO into that neighbour's halo
    // segment over NVLink (compute fused with the halo exchange of the next sweep).  CSR over the local rows:
    // entries pm_ptr[row]..pm_ptr[row+1] give (neighbour slot, position inside the slice sent to it).
    const int *pm_ptr = nullptr, *pm_nbr = nullptr, *pm_off = nullptr;
    double *const *pm_dst = nullptr;  // device table: per neighbour slot, the peer address of my slice (output vector)
};
int launch_csr(const sparsh_matrix_s *A, int epi, const double *x, double *y, const EpiArgs &args, int row_begin,
               int row_end);
// rows [b1,e1), [b2,e2) and [b3,e3) in ONE launch: the first nblk1 CTAs own range 1, the next nblk2 range 2, the rest
// range 3.  Multi-GPU: ranges 1 and 2 are the two boundary strips of a row block (their CTAs have the lowest block
// indices, so they are scheduled first and run the halo handshake), range 3 the interior rows.
struct RowRange {
    int b1, e1, b2, e2, b3, e3, nblk1, nblk2;
};
// multi-GPU: flag handshake fused into the consuming kernel (dist.cu).  Every CTA waits (acquire, system scope) until
// each neighbour's halo slice has landed before it gathers x; the last CTA to finish advances the sequence counter and
// stores the acks that allow the neighbours to overwrite the slices.
typedef unsigned long long sparsh_u64;
struct HaloSync {
    // consumer side: this kernel reads halo slices
    int nnbr = 0;
    const sparsh_u64 *flag_local[8];
    sparsh_u64 *ack_dst[8];
    sparsh_u64 *expect = nullptr;
    // producer side (fused Jacobi): this kernel also writes the NEXT slices into its neighbours (EpiArgs::pm_*); it
    // first makes sure they have consumed the previous ones, and the last CTA raises their flags
    int nsend = 0;
    sparsh_u64 *flag_dst[8];
    const sparsh_u64 *ack_local[8];
    sparsh_u64 *seq = nullptr;
    unsigned int *ticket = nullptr;
    int *err = nullptr;
    // CTAs [0, nstrip) take part in the handshake and in the fused push (the boundary strips); the CTAs behind them
    // own interior rows, which reference no halo entry and feed no neighbour: they neither wait nor signal.  Set by
    // the launcher.
    int nstrip = 0;
    // columns >= halo_begin of the gathered vector are halo entries, written by peers over NVLink while this grid may
    // already be resident: they are loaded through L2 (ld.global.cg), never through the non-coherent path
    int halo_begin = 0x7fffffff;
    long long timeout_ns = 10000000000ll;  // a wait that lasts longer raises *err instead of hanging the GPU
};
// rows [b1,e1) + [b2,e2) (strips: handshake CTAs when hs is given) + [b3,e3) (interior) in one launch
int launch_csr3(const sparsh_matrix_s *A, int epi, const double *x, double *y, const EpiArgs &args, int b1, int e1,
                int b2, int e2, int b3, int e3, const HaloSync *hs);
inline int launch_csr2(const sparsh_matrix_s *A, int epi, const double *x, double *y, const EpiArgs &args, int b1, int e1,
                       int b2, int e2, const HaloSync *hs) {
    return launch_csr3(A, epi, x, y, args, b1, e1, b2, e2, 0, 0, hs);
}

// ---- BLAS-1 (blas1.cu) -------------------------------------------------------------------------------------
int k_fill(double *x, size_t n, double v);
int k_axpy(size_t n, double a, const double *x, double *y);
int k_axpby(size_t n, double a, const double *x, double b, double *y);
int k_axpbypcz(size_t n, double a, const double *x, double b, const double *y, double c, double *z);
int k_dot(size_t n, const double *x, const double *y, double *d_out);        // d_out[0] = x.y
int k_jacobi_zero(size_t n, const double *b, const double *d, double omega, double *x);  // x = (omega*b)/d
// Krylov fused updates; scalars live in device memory (s[] indices documented in krylov.cu)
int k_pcg_update_xr(size_t n, const double *p, const double *Ap, double *x, double *r, const double *rz,
                    const double *pAp, double *rr_out);
int k_pcg_update_p(size_t n, const double *z, double *p, const double *rz_new, const double *rz_old);
int k_scalar_copy(double *dst, const double *src);
int k_cg_update_p(size_t n, const double *r, double *p, const double *rr_new, const double *rr_old);
int k_bicg_s(size_t n, const double *r, const double *Ap, double *s, const double *alpha1, const double *apr0);
int k_bicg_xr(size_t n, double *x, const double *ph, const double *sh, const double *s, const double *As, double *r,
              const double *alpha1, const double *apr0, const double *ass, const double *asas, const double *r0,
              double *out2);
int k_bicg_p(size_t n, const double *r, double *p, const double *Ap, const double *sc);
int k_dot2(size_t n, const double *a, const double *b, const double *c, double *d_out2);  // out[0]=a.b out[1]=a.c
// GMRES(m): V holds basis vectors ld apart.  d_out[q] = V_q.w for q < k (d_out needs k rounded up to 4 slots);
// w -= sum h[q] V_q;  out = sum y[q] V_q;  out = in / sqrt(*d_nrm2)
int k_mdot(size_t n, const double *V, size_t ld, int k, const double *w, double *d_out);
int k_maxpy_sub(size_t n, const double *V, size_t ld, int k, const double *d_h, double *w);
int k_lincomb(size_t n, const double *V, size_t ld, int k, const double *d_y, double *out);
int k_scale_inv_sqrt(size_t n, const double *in, const double *d_nrm2, double *out);

// matrix uploads issued by the calling host thread go to `s` (nullptr: the library's stream) — hierarchy.cu builds the
// levels of a hierarchy with several host threads, each on a stream of its own
void set_upload_stream(cudaStream_t s);
// pageable host memory <-> device through the calling thread's pinned staging buffers (matrix.cu)
int copy_h2d_staged(void *dst, const void *src, size_t bytes, cudaStream_t st);
int copy_d2h_staged(void *dst, const void *src, size_t bytes, cudaStream_t st);  // returns when the data has arrived
void release_upload_stage();  // hands the calling thread's pinned staging buffers back to the pool (matrix.cu)

// ---- device memory of a hierarchy (runtime.cu) ------------------------------------------------------------------------------
// A device allocation or free is a call into the kernel driver: 0.6-3 ms each on the shared GPU hosts, with 0.1-1 s
// outliers (tools/alloc_probe.py) — the ~450 of them behind a 13-level hierarchy cost more than moving its 3.7 GB.  So a
// hierarchy takes ONE slab; the arrays that live as long as it does are carved from the slab by the threads bound to it
// (bump pointer, 256-byte aligned; dev_free of such a pointer is a no-op, the slab is freed as a whole), and short-lived
// temporaries come from a grow-only per-thread scratch buffer.  Without a bound slab, or once it is exhausted, dev_alloc
// is cudaMalloc and dev_free is cudaFree.
struct DevSlab;
DevSlab *slab_create(size_t bytes);  // nullptr when the allocation fails: the caller goes on without
void slab_bind(DevSlab *s);          // the calling thread's dev_alloc calls carve from s; nullptr unbinds
bool slab_bound();                   // is the calling thread bound to a slab?
void slab_destroy(DevSlab *s);       // everything carved from the slab dies with it
cudaError_t dev_alloc_bytes(void **p, size_t bytes);
template <typename T>
inline cudaError_t dev_alloc(T **p, size_t bytes) {
    return dev_alloc_bytes(reinterpret_cast<void **>(p), bytes);
}
void dev_free(void *p);
// at least `bytes` of device memory owned by the calling thread, valid until its next thread_scratch call; the caller
// synchronises its stream before it lets go of the buffer.  nullptr when the allocation fails.
void *thread_scratch(size_t bytes);
void release_thread_scratch();

// ---- dense coarse solve (coarse.cu) ----------------------------------------------------------------------
struct CoarseInverse {
    int n = 0;
    double *inv = nullptr;  // n x n row-major
};
int coarse_build_inverse(int n, const int *h_rowptr, const int *h_col, const double *h_val, CoarseInverse *out);
int coarse_apply(const CoarseInverse &ci, const double *b, double *x);
void coarse_free(CoarseInverse *ci);

}  // namespace sparsh
