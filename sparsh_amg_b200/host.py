"""ctypes binding of lib/libsparsh_amg.so — the host-side mirror of the reference's C++ API (host/sparsh_amg.hpp).

`HostMatrix` is an sp_matrix_mg, `HostAmg` an AMG_GPU1_solver (native HEM/Beck + Galerkin setup on the host, then
uploaded once through the C-ABI), `call_solver` invokes the reference-named entry points (Solver_PCG_4, ...).
"""
import ctypes as C
import os

import numpy as np

from . import capi
from .device import DeviceHierarchy  # noqa: F401  (re-export convenience)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsparsh_amg.so")
_lib = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)

SOLVER_NAMES = ["AMG_Solver_CPU_baseline", "AMG_Solver_1", "AMG_Solver_2", "AMG_Solver_CPU_GPU_CI",
                "AMG_Solver_CPU_GPU_MI", "Solver_CG_1", "Solver_CG_2", "Solver_PCG_1", "Solver_PCG_2", "Solver_PCG_3",
                "Solver_PCG_4", "Solver_BiCG_1", "Solver_PBiCG_1", "Solver_PBiCG_2", "Solver_PBiCG_3", "Solver_PBiCG_4",
                "coarsening_2"]


def load():
    global _lib
    if _lib is None:
        capi.load()  # libsparsh_b200.so first (the host library links against it)
        if not os.path.exists(LIB_PATH):
            raise capi.SparshError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        lib = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        lib.sparsh_host_set_option.argtypes = [C.c_char_p, C.c_double]
        for f in ("sparsh_host_matrix_poisson3d", "sparsh_host_matrix_poisson2d", "sparsh_host_matrix_diffusion27",
                  "sparsh_host_matrix_from_csr", "sparsh_host_matrix_read", "sparsh_host_matrix_read_mm",
                  "sparsh_host_matrix_read_bin",
                  "sparsh_host_amg_setup",
                  "sparsh_host_amg_device"):
            getattr(lib, f).restype = vp
        lib.sparsh_host_matrix_poisson3d.argtypes = [C.c_int] * 3
        lib.sparsh_host_matrix_poisson2d.argtypes = [C.c_int] * 2
        lib.sparsh_host_matrix_diffusion27.argtypes = [C.c_int] * 3 + [C.c_uint]
        lib.sparsh_host_matrix_from_csr.argtypes = [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p]
        lib.sparsh_host_matrix_read.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(c_dbl_p)]
        lib.sparsh_host_matrix_read_mm.argtypes = [C.c_char_p]
        lib.sparsh_host_matrix_read_bin.argtypes = [C.c_char_p]
        lib.sparsh_host_matrix_write_bin.argtypes = [vp, C.c_char_p]
        lib.sparsh_host_free_array.argtypes = [c_dbl_p]
        lib.sparsh_host_matrix_prepare.argtypes = [vp]
        lib.sparsh_host_matrix_free.argtypes = [vp]
        lib.sparsh_host_matrix_dims.argtypes = [vp, c_int_p, c_int_p, c_int_p]
        lib.sparsh_host_matrix_arrays.argtypes = [vp, C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(c_dbl_p),
                                                  C.POINTER(c_dbl_p)]
        lib.sparsh_host_matrix_times.argtypes = [vp, c_dbl_p, c_dbl_p]
        lib.sparsh_host_matrix_color.argtypes = [vp, c_int_p, c_int_p]
        lib.sparsh_host_amg_setup.argtypes = [vp, C.c_int]
        lib.sparsh_host_amg_free.argtypes = [vp]
        lib.sparsh_host_amg_nlevels.argtypes = [vp]
        lib.sparsh_host_amg_level_dims.argtypes = [vp, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]
        lib.sparsh_host_amg_level_arrays.argtypes = [vp, C.c_int] + [C.POINTER(c_int_p), C.POINTER(c_int_p),
                                                                     C.POINTER(c_dbl_p), C.POINTER(c_dbl_p),
                                                                     C.POINTER(c_int_p), C.POINTER(c_int_p),
                                                                     C.POINTER(c_dbl_p)]
        lib.sparsh_host_amg_save.argtypes = [vp, C.c_char_p]
        lib.sparsh_host_amg_load.restype = vp
        lib.sparsh_host_amg_load.argtypes = [C.c_char_p]
        lib.sparsh_host_amg_upload.argtypes = [vp]
        lib.sparsh_host_amg_device.argtypes = [vp]
        lib.sparsh_host_call.argtypes = [C.c_char_p, vp, c_dbl_p, c_dbl_p]
        lib.sparsh_host_report.argtypes = [c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_int]
        _lib = lib
    return _lib


def set_options(**kw):
    """threads, relax, tol, tol_mode (0 abs / 1 rel), coarse_upper, coarse_lower, max_levels, sweeps, print_setup,
    print_solve, coarsening (0 HEM / 1 Beck / 2 smoothed aggregation), max_iter, use_graph, sa_theta, sa_relax — the
    run-time twins of the reference's macros (plus the knobs of the additions)."""
    lib = load()
    for k, v in kw.items():
        if lib.sparsh_host_set_option(k.encode(), float(v)) != 0:
            raise KeyError(k)


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,))


class HostMatrix:
    """an sp_matrix_mg living in the host library (arrays are exposed as zero-copy numpy views)"""

    def __init__(self, handle, prepare=True):
        self.lib = load()
        self.h = handle
        if prepare:
            self.lib.sparsh_host_matrix_prepare(self.h)  # sp_matrix_fill + sp_matrix_fill_diagonal (main.cpp:21-22)
        self._refresh()

    def _refresh(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.sparsh_host_matrix_dims(self.h, C.byref(a), C.byref(b), C.byref(c))
        self.nrow, self.ncol, self.nnz = a.value, b.value, c.value
        rp, ci, v, d = c_int_p(), c_int_p(), c_dbl_p(), c_dbl_p()
        self.lib.sparsh_host_matrix_arrays(self.h, C.byref(rp), C.byref(ci), C.byref(v), C.byref(d))
        self.rowptr = _view(rp, self.nrow + 1, np.int32)
        self.colindex = _view(ci, self.nnz, np.int32)
        self.val = _view(v, self.nnz, np.float64)
        self.diag = _view(d, self.nrow, np.float64) if d else None

    @classmethod
    def poisson3d(cls, nx, ny, nz):
        return cls(load().sparsh_host_matrix_poisson3d(nx, ny, nz))

    @classmethod
    def poisson2d(cls, nx, ny):
        return cls(load().sparsh_host_matrix_poisson2d(nx, ny))

    @classmethod
    def diffusion27(cls, nx, ny, nz, seed=1234):
        return cls(load().sparsh_host_matrix_diffusion27(nx, ny, nz, seed))

    @classmethod
    def from_csr(cls, A):
        rp = np.ascontiguousarray(A.rowptr, dtype=np.int32)
        ci = np.ascontiguousarray(A.colindex, dtype=np.int32)
        v = np.ascontiguousarray(A.val, dtype=np.float64)
        return cls(load().sparsh_host_matrix_from_csr(A.nrow, A.ncol, int(rp[-1]), capi.ip(rp), capi.ip(ci), capi.dp(v)))

    @classmethod
    def read(cls, matrixfile, rhsfile=""):
        lib = load()
        bp = c_dbl_p()
        h = lib.sparsh_host_matrix_read(os.fsencode(matrixfile), os.fsencode(rhsfile), C.byref(bp))
        M = cls(h)
        b = np.ctypeslib.as_array(bp, shape=(M.nrow,)).copy()
        lib.sparsh_host_free_array(bp)
        return M, b

    @classmethod
    def read_matrix_market(cls, path):
        """standard MatrixMarket coordinate file (1-based; general, symmetric or pattern)"""
        h = load().sparsh_host_matrix_read_mm(os.fsencode(path))
        if not h:
            raise capi.SparshError(f"{path}: not a readable MatrixMarket coordinate file")
        return cls(h)

    @classmethod
    def read_binary(cls, path):
        """binary CSR written by write_binary (magic SPRSHCSR; the arrays as they lie in memory)"""
        h = load().sparsh_host_matrix_read_bin(os.fsencode(path))
        if not h:
            raise capi.SparshError(f"{path}: not a readable binary CSR file")
        return cls(h)

    def write_binary(self, path):
        if self.lib.sparsh_host_matrix_write_bin(self.h, os.fsencode(path)) != 0:
            raise capi.SparshError(f"could not write {path}")

    def times(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.nrow)
        self.lib.sparsh_host_matrix_times(self.h, capi.dp(x), capi.dp(y))
        return y

    def color_reorder(self):
        perm = np.empty(self.nrow, dtype=np.int32)
        cc = np.zeros(int(np.max(np.diff(self.rowptr))) + 2, dtype=np.int32)
        nc = self.lib.sparsh_host_matrix_color(self.h, capi.ip(perm), capi.ip(cc))
        self._refresh()
        return nc, perm, cc[: nc + 1].copy()

    def free(self):
        if getattr(self, "h", None):
            self.lib.sparsh_host_matrix_free(self.h)
            self.h = None


class _Level:
    pass


class HostAmg:
    """AMG_GPU1_solver: native host setup (reference src/AMG_phases.cpp:35-147), hierarchy uploaded once"""

    def __init__(self, A, sor=False):
        self.lib = load()
        self.A = A
        self.h = self.lib.sparsh_host_amg_setup(A.h, int(sor))
        if sor:
            A._refresh()
        self.nlevels = self.lib.sparsh_host_amg_nlevels(self.h)

    def save(self, directory):
        """write every level as raw binary files under `directory` (ideally /dev/shm/...) for HostAmg.load"""
        os.makedirs(directory, exist_ok=True)
        rc = self.lib.sparsh_host_amg_save(self.h, os.fsencode(directory))
        if rc != 0:
            raise capi.SparshError(f"could not save the hierarchy to {directory} ({rc})")

    @classmethod
    def load(cls, directory):
        """map a saved hierarchy read-only: N ranks of a node share ONE copy in the page cache (host/share.cpp)"""
        self = cls.__new__(cls)
        self.lib = load()
        self.A = None
        self.h = self.lib.sparsh_host_amg_load(os.fsencode(directory))
        if not self.h:
            raise capi.SparshError(f"could not map a hierarchy from {directory}")
        self.nlevels = self.lib.sparsh_host_amg_nlevels(self.h)
        return self

    def level_dims(self, k):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.lib.sparsh_host_amg_level_dims(self.h, k, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, c.value, d.value

    def levels(self):
        """list of dict(A=..., diag=..., P=...) with zero-copy numpy views (same shape DeviceHierarchy takes)"""
        out = []
        for k in range(self.nlevels):
            nrow, nnz, pncol, pnnz = self.level_dims(k)
            rp, ci, v, d, prp, pci, pv = c_int_p(), c_int_p(), c_dbl_p(), c_dbl_p(), c_int_p(), c_int_p(), c_dbl_p()
            self.lib.sparsh_host_amg_level_arrays(self.h, k, *[C.byref(p) for p in (rp, ci, v, d, prp, pci, pv)])
            A = _Level()
            A.nrow, A.ncol, A.nnz = nrow, nrow, nnz
            A.rowptr, A.colindex, A.val = _view(rp, nrow + 1, np.int32), _view(ci, nnz, np.int32), _view(v, nnz, np.float64)
            P = None
            if k < self.nlevels - 1:
                P = _Level()
                P.nrow, P.ncol, P.nnz = nrow, pncol, pnnz
                P.rowptr, P.colindex, P.val = (_view(prp, nrow + 1, np.int32), _view(pci, pnnz, np.int32),
                                               _view(pv, pnnz, np.float64))
            out.append(dict(A=A, diag=_view(d, nrow, np.float64), P=P))
        return out

    def upload(self):
        self.lib.sparsh_host_amg_upload(self.h)  # GPU_Allocations()
        return self.device()

    def device(self):
        """borrowed sparsh_hierarchy_t wrapped for direct C-ABI calls"""
        d = DeviceHierarchy.__new__(DeviceHierarchy)
        d.lib = capi.load()
        d.h = self.lib.sparsh_host_amg_device(self.h)
        d.n = self.level_dims(0)[0]
        d.nlevels = self.nlevels
        d.free = lambda: None  # owned by the AMG_GPU1_solver
        return d

    def free(self):
        if getattr(self, "h", None):
            self.lib.sparsh_host_amg_free(self.h)
            self.h = None


def report(maxhist=100000):
    lib = load()
    it, cv = C.c_int(), C.c_int()
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    hist = np.zeros(maxhist)
    n = lib.sparsh_host_report(C.byref(it), C.byref(cv), C.byref(a), C.byref(b), C.byref(c), capi.dp(hist), maxhist)
    return dict(iterations=it.value, converged=bool(cv.value), setup_seconds=a.value, upload_seconds=b.value,
                solve_seconds=c.value, history=hist[: min(n, maxhist)].copy())


def call_solver(name, A, b, x):
    """name in SOLVER_NAMES; b, x numpy (x in/out), exactly the reference's `void f(sp_matrix_mg&, double*&, double*&)`"""
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert x.dtype == np.float64 and x.flags["C_CONTIGUOUS"]
    if load().sparsh_host_call(name.encode(), A.h, capi.dp(b), capi.dp(x)) != 0:
        raise KeyError(name)
    return report()
