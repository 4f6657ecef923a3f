"""Synthetic matrices of BASELINE.json's configs, vectorised numpy.

Natural ordering, x fastest, Dirichlet boundary by truncation, columns sorted — the layout the reference's readers
produce (0-based int32 CSR).  The host library (host/generators.cpp) has the fast multi-threaded twins used by
bench.py; these numpy versions serve small tests and tools.
"""
import numpy as np


class HostCSR:
    def __init__(self, nrow, ncol, rowptr, colindex, val):
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.colindex = np.ascontiguousarray(colindex, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.nnz = int(self.rowptr[-1])


def poisson_7pt(nx, ny, nz, diag=6.0):
    n = nx * ny * nz
    i = np.arange(n, dtype=np.int64)
    x, y, z = i % nx, (i // nx) % ny, i // (nx * ny)
    offs = [(-nx * ny, z > 0), (-nx, y > 0), (-1, x > 0), (0, np.ones(n, bool)), (1, x < nx - 1), (nx, y < ny - 1),
            (nx * ny, z < nz - 1)]
    counts = np.zeros(n, dtype=np.int64)
    for _, m in offs:
        counts += m
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    col = np.empty(rowptr[-1], dtype=np.int32)
    val = np.empty(rowptr[-1], dtype=np.float64)
    pos = rowptr[:-1].copy()
    for off, m in offs:
        idx = pos[m]
        col[idx] = (i[m] + off).astype(np.int32)
        val[idx] = diag if off == 0 else -1.0
        pos += m
    return HostCSR(n, n, rowptr, col, val)


def poisson_5pt(nx, ny):
    return poisson_7pt(nx, ny, 1, diag=4.0)
