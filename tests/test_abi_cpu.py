"""CPU tests of the drop-in boundary: libmtg_cuda.so builds for sm_100a, loads,
exports every entry point include/mtg_cuda.h declares, refuses to run without
a device (no CPU fallback), and its host-side constant tables satisfy the
time-scaling identity against the oracle's restatement of the reference.
No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import mav_tube_trajectory_generation_b200 as m

    m._build.build()
    return m.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mtg_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mtg_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert "mtg_solve_batch" in names and "mtg_create" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mtg_cuda.h but not exported"
    assert lib.mtg_abi_version() == 2


def test_library_contains_sm100a_code_only():
    import mav_tube_trajectory_generation_b200 as m

    out = subprocess.run(["cuobjdump", "-lelf", m._build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import mav_tube_trajectory_generation_b200 as m

    with pytest.raises(m.MtgError, match="no CUDA device"):
        m.Context(0)


def test_product_does_not_link_the_oracle():
    import mav_tube_trajectory_generation_b200 as m

    out = subprocess.run(["ldd", m._build.LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out and "rpoly" not in out
    pkg = os.path.join(ROOT, "mav_tube_trajectory_generation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text, f


@pytest.mark.parametrize("N,d", [(10, 4), (10, 3), (10, 2), (10, 0), (8, 3), (6, 2), (12, 5), (4, 1)])
def test_time_scaling_identity(po, N, d):
    """H(T) = T^(1-2d) S H1 S and A(T)^-1 = diag(T^-j) A(1)^-1 S versus the reference
    formulation (LIN_I:101-111, 132-169, 557-573) as restated by the oracle."""
    import mav_tube_trajectory_generation_b200 as m

    H1, Ai1 = m.get_tables(N, d)
    assert np.array_equal(H1, H1.T)
    h = N // 2
    alpha = np.array(list(range(h)) * 2, dtype=float)
    for T in (0.37, 1.0, 4.3, 9.6):
        A = po.setup_mapping_matrix(N, T)
        Ainv = po.invert_mapping_matrix(A)
        Q = po.quadratic_cost_jacobian(N, d, T)
        H = Ainv.T @ Q @ Ainv
        S = T ** alpha
        H_closed = T ** (1 - 2 * d) * (S[:, None] * H1 * S[None, :])
        # the fp64 reference order loses ~cond(A(T)) * eps: 1e-10 at N = 10, 1e-8 at N = 12
        tol = 2e-9 if N <= 10 else 2e-7
        assert np.abs(H - H_closed).max() <= tol * np.abs(H_closed).max()
        Ainv_closed = (T ** -np.arange(N))[:, None] * Ai1 * S[None, :]
        assert np.abs(Ainv - Ainv_closed).max() <= tol * np.abs(Ainv_closed).max()
    with pytest.raises(m.MtgError):
        m.get_tables(10, 5)


def test_tables_against_exact_rationals():
    """H1 and A(1)^-1 are rationals: check against a Fraction computation."""
    from fractions import Fraction as F

    import mav_tube_trajectory_generation_b200 as m

    N, d, h = 10, 4, 5

    def ff(n, i):
        r = F(1)
        for k in range(i - n + 1, i + 1):
            r *= k
        return r if i >= n else F(0)

    A = [[F(0)] * N for _ in range(N)]
    for r in range(h):
        A[r][r] = ff(r, r)
        for j in range(r, N):
            A[r + h][j] = ff(r, j)
    # Gauss-Jordan in exact arithmetic
    M = [row[:] + [F(int(i == j)) for j in range(N)] for i, row in enumerate(A)]
    for c in range(N):
        p = next(r for r in range(c, N) if M[r][c] != 0)
        M[c], M[p] = M[p], M[c]
        inv = 1 / M[c][c]
        M[c] = [x * inv for x in M[c]]
        for r in range(N):
            if r != c and M[r][c] != 0:
                f = M[r][c]
                M[r] = [x - f * y for x, y in zip(M[r], M[c])]
    Ai = [row[N:] for row in M]
    Q = [[ff(d, a) * ff(d, b) * 2 / F(a + b - 2 * d + 1) if a >= d and b >= d else F(0)
          for b in range(N)] for a in range(N)]
    QA = [[sum(Q[i][k] * Ai[k][j] for k in range(N)) for j in range(N)] for i in range(N)]
    H = [[sum(Ai[k][i] * QA[k][j] for k in range(N)) for j in range(N)] for i in range(N)]
    H1, Ai1 = m.get_tables(N, d)
    for i in range(N):
        for j in range(N):
            assert H1[i, j] == float(H[i][j])
            assert Ai1[i, j] == float(Ai[i][j])


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference runs without a GPU (it times the oracle port on the host cores) and
    puts exactly one JSON object on stdout, with the keys the driver reads."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "trajectories/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("configs[1]")


def test_bench_workload_generator_matches_the_reference_recipe():
    """make_workload: vertices in the +-10 m box, consecutive vertices further than 0.2 m apart
    (vertex.cpp:65-72), Nfabian times (vertex.cpp:252-269) equal to the oracle's."""
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    from oracle import pyoracle as po

    pos, times = bench.make_workload(512, seed=1)
    assert pos.shape == (11, 3, 512) and times.shape == (10, 512)
    assert np.abs(pos).max() <= 10.0
    d = np.sqrt((np.diff(pos, axis=0) ** 2).sum(axis=1))
    assert d.min() > 0.2
    for b in (0, 17, 511):
        assert np.allclose(times[:, b], po.estimate_segment_times_nfabian(pos[:, :, b], 3.0, 5.0), rtol=1e-14)
