import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


# Reference parameter sets, TEST_OPT:754-839 (D, derivative, K, seed, box, v_max, a_max)
REFERENCE_PARAMS = {
    "segment_1_dim_1": (1, 4, 1, 100, 10.0, 3.0, 5.0),
    "segment_10_dim_1": (1, 4, 10, 102, 10.0, 3.0, 5.0),
    "segment_50_dim_1": (1, 4, 50, 103, 10.0, 3.0, 5.0),
    "segment_1_dim_3": (3, 4, 1, 104, 10.0, 3.0, 5.0),
    "segment_10_dim_3": (3, 4, 10, 105, 10.0, 3.0, 5.0),
    "segment_50_dim_3": (3, 4, 50, 106, 10.0, 3.0, 5.0),
    "segment_75_dim_3": (3, 4, 75, 106, 10.0, 3.0, 5.0),
    "accel_5_dim_1": (1, 2, 5, 107, 10.0, 3.0, 5.0),
    "accel_1_dim_3": (3, 2, 1, 108, 10.0, 1.0, 2.0),
    "accel_5_dim_3": (3, 2, 5, 109, 10.0, 3.0, 5.0),
    "jerk_5_dim_3": (3, 3, 5, 110, 10.0, 3.0, 5.0),
}


def make_reference_problem(name):
    """Fixture of TEST_OPT:62-78: createRandomVertices(4, K, +-box, seed) + Nfabian times."""
    from oracle import pyoracle as po

    D, derivative, K, seed, box, v_max, a_max = REFERENCE_PARAMS[name]
    mask, values = po.create_random_vertices(4, K, [-box] * D, [box] * D, seed)
    times = po.estimate_segment_times_nfabian(values[:, 0, :], v_max, a_max)
    return dict(D=D, derivative=derivative, K=K, seed=seed, mask=mask, values=values,
                times=times, v_max=v_max, a_max=a_max, N=10)


def normwise_error(a, b):
    """max over polynomials of ||a-b||_inf / ||b||_inf (last axis = coefficients)."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.abs(b).max(axis=-1)
    num = np.abs(a - b).max(axis=-1)
    den = np.where(den == 0.0, 1.0, den)
    return float((num / den).max())


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle

    pyoracle.lib()
    return pyoracle
