"""Generates tests/golden/*.npz (committed). Run from the repo root in the
authoring container:  python tests/golden/make_golden.py

Contents:
  two_vertices_setup.npz  the reference's only golden vector (TEST_OPT:707-751,
                          Matlab literals at :741-744) with its inputs.
  convolution.npz         TEST_POLY:68-79.
  reference_params.npz    for the parameter sets of TEST_OPT:754-839 with
                          K <= 10 plus ConstraintPacking seeds 12345..12354
                          (+-50 m box): inputs from the restated mt19937
                          generator (VTX_C:27-82), oracle (dense-QR) outputs and
                          the 60-digit mpmath solution of the same equations.
  rpoly_kat.npz           roots returned by the UNMODIFIED reference rpoly
                          (oracle/_ref) for fixed polynomials, so the GPU box can
                          check extrema without /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as po  # noqa: E402
from exact_solver import exact_solve  # noqa: E402
from conftest import REFERENCE_PARAMS  # noqa: E402


def main():
    # --- TwoVerticesSetup
    matlab = np.array([-0.000000000000004, 0.000000000000004, -0.000000000000006,
                       0.000000000000003, -0.000000000000001, 0.201600000000015,
                       -0.134400000000012, 0.034560000000004, -0.004032000000000,
                       0.000179200000000])
    mask = np.ones((2, 5), dtype=np.uint8)
    values = np.zeros((2, 5, 1))
    values[1, 0, 0] = 5.0
    np.savez(os.path.join(HERE, "two_vertices_setup.npz"), matlab_coeffs=matlab, mask=mask,
             values=values, times=np.array([5.0]), N=10, derivative=4)

    # --- Convolution
    np.savez(os.path.join(HERE, "convolution.npz"), data=np.array([1.0, 2.0]),
             kernel=np.array([-1.0, 3.0]), expected=np.array([-1.0, 1.0, 6.0]))

    # --- reference parameter sets
    out = {}
    names = []
    cases = []
    for name, (D, der, K, seed, box, v, a) in REFERENCE_PARAMS.items():
        if K <= 10:
            cases.append((name, D, der, K, seed, box, v, a))
    for i in range(10):
        cases.append((f"packing_{12345 + i}", 3, 4, 10, 12345 + i, 50.0, 3.0, 5.0))
    for name, D, der, K, seed, box, v, a in cases:
        mask, values = po.create_random_vertices(4, K, [-box] * D, [box] * D, seed)
        times = po.estimate_segment_times_nfabian(values[:, 0, :], v, a)
        s = po.solve(10, der, times, mask, values, solver=0)
        ce, cost_e, dp_e = exact_solve(10, der, times, mask, values)
        names.append(name)
        out[name + "/mask"] = mask
        out[name + "/values"] = values
        out[name + "/times"] = times
        out[name + "/derivative"] = der
        out[name + "/oracle_coeffs"] = s.coeffs
        out[name + "/oracle_cost"] = s.cost
        out[name + "/oracle_d_p"] = s.d_p
        out[name + "/exact_coeffs"] = ce
        out[name + "/exact_cost"] = cost_e
        out[name + "/exact_d_p"] = dp_e
        if der == 4 or True:
            for dd in (1, 2):
                r = po.opt_max_magnitude(s.coeffs, times, dd)
                out[name + f"/oracle_max_mag_{dd}"] = np.array(r)
        print(name, "done")
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "reference_params.npz"), **out)

    # --- rpoly known answers from the reference core
    rng = np.random.RandomState(1234567)
    polys, roots = [], []
    for i in range(40):
        n = rng.randint(3, 19)
        c = rng.uniform(-100, 100, size=n)
        ok, r = po.find_roots_jenkins_traub(c)
        assert ok
        p = np.zeros(22)
        p[:n] = c
        rr = np.full(21, np.nan, dtype=np.complex128)
        rr[: r.size] = r
        polys.append(p)
        roots.append(rr)
    np.savez(os.path.join(HERE, "rpoly_kat.npz"), polys=np.array(polys), roots=np.array(roots))


if __name__ == "__main__":
    main()
