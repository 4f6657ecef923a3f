"""The drop-in claim, literally: a C++ caller written against the reference's API (examples/dropin_main.cpp, same
flow as the reference's main.cpp) compiles and links against host/AMG.hpp + our two libraries (CPU test), and on a GPU
reproduces the reference's iteration counts on the bundled fixture."""
import os
import re
import subprocess

import numpy as np
import pytest
from conftest import ROOT

EXE = os.path.join(ROOT, "examples", "dropin_main")


def build(src="dropin_main.cpp", exe=EXE):
    lib = os.path.join(ROOT, "sparsh_amg_b200", "lib")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "sparsh_amg_b200", "host"),
           os.path.join(ROOT, "examples", src), "-o", exe, "-L", lib, "-lsparsh_amg", "-lsparsh_b200",
           f"-Wl,-rpath,{lib}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


def write_fixture(tmp_path, A, b):
    """the reference's two-file format (src/AMG_file_read.cpp:39-72)"""
    mf, rf = tmp_path / "matrix.txt", tmp_path / "rhs.txt"
    rows = np.repeat(np.arange(A.nrow), np.diff(A.rowptr))
    with open(mf, "w") as f:
        f.write(f"{A.nrow} {A.ncol} {A.nnz}\n")
        f.write("".join(f"{r}\t{c}\t{v!r}\t\n" for r, c, v in zip(rows.tolist(), A.colindex.tolist(), A.val.tolist())))
    with open(rf, "w") as f:
        f.write(f"{A.nrow}\n" + "\n".join(repr(float(v)) for v in b) + "\n")
    return mf, rf


def test_multigpu_caller_compiles_and_links(tmp_path):
    build("dropin_multigpu.cpp", str(tmp_path / "dropin_multigpu"))


def test_reference_style_caller_compiles_and_links():
    build()
    assert os.path.exists(EXE)


def test_additions_example_compiles_and_links(tmp_path):
    """smoothed aggregation, GMRES and the extra readers are reachable from the same header"""
    lib = os.path.join(ROOT, "sparsh_amg_b200", "lib")
    exe = str(tmp_path / "additions_main")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "sparsh_amg_b200", "host"),
           os.path.join(ROOT, "examples", "additions_main.cpp"), "-o", exe, "-L", lib, "-lsparsh_amg", "-lsparsh_b200",
           f"-Wl,-rpath,{lib}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    assert subprocess.run([exe], capture_output=True).returncode == 2  # usage message without a GPU or arguments


@pytest.mark.gpu
def test_reference_style_caller_runs(tmp_path, fixture_system, golden):
    build()
    A, b = fixture_system
    mf, rf = write_fixture(tmp_path, A, b)
    out = subprocess.run([EXE, str(mf), str(rf)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rep = dict(re.findall(r"REPORT (\w+) iterations=(\d+)", out.stdout))
    assert f"Matrix Size\t{A.nrow}" in out.stdout
    want_amg = len(golden["fixture"]["AMG_Solver_CPU_baseline"]["hist"])  # 30
    assert int(rep["AMG_Solver_CPU_GPU_CI"]) == want_amg
    assert int(rep["AMG_Solver_CPU_GPU_MI"]) == want_amg
    assert int(rep["AMG_Solver_CPU_baseline"]) == want_amg
    assert int(rep["Solver_PCG_4"]) == len(golden["fixture"]["Solver_PCG_1"]["hist"])      # 13
    assert int(rep["Solver_PBiCG_4"]) == len(golden["fixture"]["Solver_PBiCG_1"]["hist"])  # 7


@pytest.mark.gpu
def test_multigpu_entry_points_from_cpp(tmp_path, fixture_system, golden):
    """Solver_PCG_MG / AMG_Solver_MG / Solver_PBiCG_MG called from a plain C++ program, one process per GPU (no Python, no
    torchrun: the 128-byte id travels through a file): the reference's iteration counts on the bundled system, the global
    solution on every rank, and a clean sparsh_dist_finalize() + return from main."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = str(tmp_path / "dropin_multigpu")
    build("dropin_multigpu.cpp", exe)
    A, b = fixture_system
    mf, rf = write_fixture(tmp_path, A, b)
    idf = str(tmp_path / "nccl_id.bin")
    procs = [subprocess.Popen([exe, str(mf), str(rf), "2", str(r), idf], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = []
    try:
        for p in procs:
            o, e = p.communicate(timeout=300)
            outs.append((p.returncode, o, e))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()  # the exact processes this test started
    want = {"Solver_PCG_MG": len(golden["fixture"]["Solver_PCG_1"]["hist"]),            # 13
            "AMG_Solver_MG": len(golden["fixture"]["AMG_Solver_CPU_baseline"]["hist"]),  # 30
            "Solver_PBiCG_MG": len(golden["fixture"]["Solver_PBiCG_1"]["hist"])}         # 7
    for r, (rc, o, e) in enumerate(outs):
        assert rc == 0, o[-2000:] + e[-2000:]
        assert f"FINALIZED rank={r}" in o
        rep = {m.group(1): (int(m.group(2)), int(m.group(3)), float(m.group(4)))
               for m in re.finditer(r"REPORT rank=\d+ (\w+) iterations=(\d+) converged=(\d) residual=(\S+)", o)}
        for name, count in want.items():
            it, conv, res = rep[name]
            assert conv == 1 and abs(it - count) <= 1, (name, it, count)
            assert res <= 2e-8, (name, res)
