"""ctypes binding of libmtg_cuda.so (include/mtg_cuda.h) — the only way Python
reaches the product path. There is no CPU fallback: if the library is missing or
no CUDA device is present this module raises.

Tensors are batches of records in SoA (batch innermost) or AoS (record-contiguous)
layout (see mtg_cuda.h). A
``torch`` CUDA tensor selects MTG_MEM_DEVICE (nothing is copied; work is
enqueued on torch's current stream); a numpy array or CPU torch tensor selects
MTG_MEM_HOST (the library stages H2D/D2H itself).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

MTG_MEM_DEVICE = 0
MTG_MEM_HOST = 1
LAYOUT_SOA = 0
LAYOUT_AOS = 1

ST_BAD_TIME = 1
ST_NOT_SPD = 2
ST_OUT_OF_RANGE = 4
ST_TRUNCATED = 8
ST_NO_CONVERGENCE = 16


class MtgError(RuntimeError):
    pass


class ProblemDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("D", C.c_int32), ("N", C.c_int32),
                ("derivative_to_optimize", C.c_int32), ("memory", C.c_int32), ("layout", C.c_int32)]


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MTG_CUDA_LIB", _build.LIB)   # override: A/B measurements of two builds in one run
    if not os.path.exists(path):
        raise MtgError(f"{path} is missing: run __graft_entry__.build() "
                       "(python -m mav_tube_trajectory_generation_b200._build). "
                       "There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, dp, u32p = C.c_void_p, C.c_void_p, C.c_void_p
    lib.mtg_abi_version.restype = C.c_int
    lib.mtg_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.mtg_destroy.argtypes = [vp]
    lib.mtg_destroy.restype = None
    lib.mtg_last_error.argtypes = [vp]
    lib.mtg_last_error.restype = C.c_char_p
    lib.mtg_launch_count.argtypes = [vp]
    lib.mtg_launch_count.restype = C.c_uint64
    lib.mtg_sync.argtypes = [vp, vp]
    lib.mtg_get_tables.argtypes = [C.c_int, C.c_int, dp, dp]
    lib.mtg_solve_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, dp, dp, u32p, vp]
    lib.mtg_cost_time_fd_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, C.c_double, C.c_int,
                                           dp, dp, dp, dp, u32p, vp]
    lib.mtg_extrema_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, C.c_int, dp, dp, vp, dp, dp, vp, dp, dp,
                                      u32p, vp]
    lib.mtg_extrema_candidates_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, C.c_int, C.c_int,
                                                 C.c_int, dp, dp, vp, u32p, vp]
    lib.mtg_poly_real_roots_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_int, dp, vp,
                                              u32p, vp]
    lib.mtg_soft_constraint_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, C.c_int, vp, dp, C.c_double,
                                              C.c_double, dp, dp, u32p, vp]
    lib.mtg_control_points_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp,
                                             vp, u32p, vp]
    lib.mtg_vertex_at_time_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, C.c_int, dp, vp, u32p, vp]
    lib.mtg_pick_dimensions_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, C.c_int, dp, C.c_int, vp, dp, vp]
    lib.mtg_concat_segments_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                              dp, dp, vp]
    lib.mtg_compute_cost_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, u32p, vp]
    lib.mtg_sample_dump_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, C.c_double, C.c_int, dp, vp, u32p, vp]
    lib.mtg_cost_derivative_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, dp, dp, dp, u32p, vp]
    lib.mtg_soft_constraint_gradient_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, C.c_int, vp, dp, C.c_double,
                                                       C.c_double, C.c_double, C.c_int, dp, dp, u32p, vp]
    lib.mtg_nl_descent_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, C.c_int, vp, dp, C.c_double,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                         dp, dp, u32p, vp]
    lib.mtg_collision_cost_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, vp, vp, C.c_double, vp, vp, C.c_double,
                                             C.c_double, C.c_double, C.c_double, dp, dp, vp, vp, u32p, vp]
    lib.mtg_generate_candidates_batch.argtypes = [vp, C.POINTER(ProblemDesc), C.c_uint64, C.c_int64, vp, vp, C.c_double,
                                                  C.c_double, C.c_double, dp, dp, vp]
    lib.mtg_argmin_batch.argtypes = [vp, dp, u32p, C.c_int64, C.c_int64, C.c_int, vp, vp]
    lib.mtg_set_solve_overlap.argtypes = [vp, C.c_int]
    lib.mtg_solve_argmin_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, dp, dp, u32p, C.c_int64, C.c_int,
                                           vp, vp]
    lib.mtg_nccl_unique_id.argtypes = [vp, C.c_char_p]
    lib.mtg_nccl_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
    lib.mtg_argmin_allgather.argtypes = [vp, dp, u32p, C.c_int64, C.c_int64, C.POINTER(C.c_double),
                                         C.POINTER(C.c_int64), vp]
    lib.mtg_best_allgather.argtypes = [vp, vp, vp, vp]
    lib.mtg_probe_fp64_fma.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), vp]
    lib.mtg_set_free_constraints_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, dp, dp, u32p, vp]
    lib.mtg_solve_generic_batch.argtypes = [vp, C.POINTER(ProblemDesc), vp, dp, dp, dp, dp, dp, u32p, vp]
    lib.mtg_coeffs_from_derivatives_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, u32p, vp]
    lib.mtg_max_time_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, vp]
    lib.mtg_eval_range_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, dp, C.c_int, C.c_int,
                                         dp, dp, vp, vp, u32p, vp]
    lib.mtg_eval_at_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, C.c_int, C.c_int, dp, vp,
                                      u32p, vp]
    lib.mtg_feasibility_batch.argtypes = [vp, C.POINTER(ProblemDesc), dp, dp, dp, dp, C.c_double,
                                          C.c_double, dp, dp, dp, C.c_int, dp, vp, dp, dp, vp, vp, u32p, vp]
    _lib = lib
    return lib


def get_tables(N: int, derivative: int):
    """H1 and A(1)^-1 (host computation, no device needed)."""
    lib = load()
    H1 = np.zeros((N, N))
    Ai = np.zeros((N, N))
    rc = lib.mtg_get_tables(N, derivative, H1.ctypes.data, Ai.ctypes.data)
    if rc:
        raise MtgError("invalid (N, derivative)")
    return H1, Ai


def _is_torch(x) -> bool:
    return hasattr(x, "data_ptr")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


class Context:
    """One mtg_ctx on one CUDA device."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.mtg_create(device, C.byref(h))
        if rc == -3:
            raise MtgError("no CUDA device: libmtg_cuda.so has no CPU path")
        if rc:
            raise MtgError(f"mtg_create failed rc={rc}")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mtg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ utils
    def _check(self, rc: int, what: str):
        if rc:
            raise MtgError(f"{what} failed rc={rc}: {self._lib.mtg_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self._lib.mtg_launch_count(self._h))

    def _mode(self, x) -> int:
        return MTG_MEM_DEVICE if (_is_torch(x) and x.is_cuda) else MTG_MEM_HOST

    def _stream(self, mode, stream):
        if stream is not None:
            return C.c_void_p(int(stream))
        if mode == MTG_MEM_DEVICE:
            import torch

            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)

    def _empty(self, like, shape, dtype="f8"):
        if _is_torch(like):
            import torch

            tdt = {"f8": torch.float64, "u4": torch.int32, "u1": torch.uint8, "i4": torch.int32}[dtype]
            pin = (not like.is_cuda) and like.is_pinned()
            return torch.empty(shape, dtype=tdt, device=like.device, pin_memory=pin)
        return np.empty(shape, dtype={"f8": np.float64, "u4": np.uint32, "u1": np.uint8,
                                      "i4": np.int32}[dtype])

    @staticmethod
    def _contig(x, name):
        ok = x.is_contiguous() if _is_torch(x) else x.flags["C_CONTIGUOUS"]
        if not ok:
            raise MtgError(f"{name} must be contiguous (SoA, batch innermost)")

    # ------------------------------------------------------------------ solve
    def solve_batch(self, positions, seg_times, end_derivatives=None, N: int = 10,
                    derivative: int = 4, layout: str = "soa", want_cost=True, want_free=False,
                    want_status=True, want_coeffs=True, out=None, stream=None):
        """mtg_solve_batch.
        layout "soa": positions [K+1,D,B], seg_times [K,B], end_derivatives [2,N/2-1,D,B];
                      returns coeffs [K,D,N,B], free [D,K-1,N/2-1,B]
        layout "aos": positions [B,K+1,D], seg_times [B,K], end_derivatives [B,2,N/2-1,D];
                      returns coeffs [B,K,D,N], free [B,D,K-1,N/2-1]
        plus cost [B], status [B]."""
        aos = layout == "aos"
        if aos:
            B, Kp1, D = positions.shape
        else:
            Kp1, D, B = positions.shape
        K = Kp1 - 1
        assert tuple(seg_times.shape) == ((B, K) if aos else (K, B))
        for nm, x in (("positions", positions), ("seg_times", seg_times)):
            self._contig(x, nm)
        mode = self._mode(positions)
        desc = ProblemDesc(B, K, D, N, derivative, mode, LAYOUT_AOS if aos else LAYOUT_SOA)
        out = out or {}
        nf = N // 2 - 1
        coeffs = out.get("coeffs") if want_coeffs else None
        if want_coeffs and coeffs is None:
            coeffs = self._empty(positions, (B, K, D, N) if aos else (K, D, N, B))
        cost = out.get("cost") if want_cost else None
        if want_cost and cost is None:
            cost = self._empty(positions, (B,))
        free = out.get("free") if want_free else None
        if want_free and free is None:
            free = self._empty(positions, (B, D, max(K - 1, 0), nf) if aos else (D, max(K - 1, 0), nf, B))
        status = out.get("status") if want_status else None
        if want_status and status is None:
            status = self._empty(positions, (B,), "u4")
        rc = self._lib.mtg_solve_batch(self._h, C.byref(desc), _ptr(positions), _ptr(end_derivatives),
                                       _ptr(seg_times), _ptr(coeffs), _ptr(cost), _ptr(free),
                                       _ptr(status), self._stream(mode, stream))
        self._check(rc, "mtg_solve_batch")
        return dict(coeffs=coeffs, cost=cost, free=free, status=status)

    def solve_generic_batch(self, mask, values, seg_times, N: int = 10, derivative: int = 4, layout: str = "soa",
                            want_free=True, stream=None):
        """mtg_solve_generic_batch. mask [K+1, N/2] uint8 numpy (shared by the batch); values soa
        [K+1, N/2, D, B] / aos [B, K+1, N/2, D]; seg_times soa [K, B] / aos [B, K]."""
        aos = layout == "aos"
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        h = N // 2
        if aos:
            B, Kp1, hh, D = values.shape
        else:
            Kp1, hh, D, B = values.shape
        assert hh == h and mask.shape == (Kp1, h)
        K = Kp1 - 1
        n_free = int((mask == 0).sum())
        mode = self._mode(values)
        desc = ProblemDesc(B, K, D, N, derivative, mode, LAYOUT_AOS if aos else LAYOUT_SOA)
        coeffs = self._empty(values, (B, K, D, N) if aos else (K, D, N, B))
        cost = self._empty(values, (B,))
        status = self._empty(values, (B,), "u4")
        free = self._empty(values, (B, D, n_free) if aos else (D, n_free, B)) if (want_free and n_free) else None
        rc = self._lib.mtg_solve_generic_batch(self._h, C.byref(desc), mask.ctypes.data, _ptr(values), _ptr(seg_times),
                                               _ptr(coeffs), _ptr(cost), _ptr(free), _ptr(status),
                                               self._stream(mode, stream))
        self._check(rc, "mtg_solve_generic_batch")
        return dict(coeffs=coeffs, cost=cost, free=free, status=status)

    def coeffs_from_derivatives_batch(self, derivatives, seg_times, N: int = 10, derivative: int = 4,
                                      layout: str = "soa", stream=None):
        """mtg_coeffs_from_derivatives_batch. derivatives soa [K+1, N/2, D, B] / aos [B, K+1, N/2, D]."""
        aos = layout == "aos"
        if aos:
            B, Kp1, h, D = derivatives.shape
        else:
            Kp1, h, D, B = derivatives.shape
        K = Kp1 - 1
        mode = self._mode(derivatives)
        desc = ProblemDesc(B, K, D, N, derivative, mode, LAYOUT_AOS if aos else LAYOUT_SOA)
        coeffs = self._empty(derivatives, (B, K, D, N) if aos else (K, D, N, B))
        cost = self._empty(derivatives, (B,))
        status = self._empty(derivatives, (B,), "u4")
        rc = self._lib.mtg_coeffs_from_derivatives_batch(self._h, C.byref(desc), _ptr(derivatives), _ptr(seg_times),
                                                         _ptr(coeffs), _ptr(cost), _ptr(status),
                                                         self._stream(mode, stream))
        self._check(rc, "mtg_coeffs_from_derivatives_batch")
        return dict(coeffs=coeffs, cost=cost, status=status)

    def set_free_constraints_batch(self, positions, seg_times, free, end_derivatives=None, N: int = 10,
                                   derivative: int = 4, layout: str = "soa", stream=None):
        """mtg_set_free_constraints_batch: coefficients + cost from given free derivatives d_p."""
        aos = layout == "aos"
        if aos:
            B, Kp1, D = positions.shape
        else:
            Kp1, D, B = positions.shape
        K = Kp1 - 1
        mode = self._mode(positions)
        desc = ProblemDesc(B, K, D, N, derivative, mode, LAYOUT_AOS if aos else LAYOUT_SOA)
        coeffs = self._empty(positions, (B, K, D, N) if aos else (K, D, N, B))
        cost = self._empty(positions, (B,))
        status = self._empty(positions, (B,), "u4")
        rc = self._lib.mtg_set_free_constraints_batch(self._h, C.byref(desc), _ptr(positions),
                                                      _ptr(end_derivatives), _ptr(seg_times), _ptr(free),
                                                      _ptr(coeffs), _ptr(cost), _ptr(status),
                                                      self._stream(mode, stream))
        self._check(rc, "mtg_set_free_constraints_batch")
        return dict(coeffs=coeffs, cost=cost, status=status)

    def cost_time_fd_batch(self, positions, seg_times, free, increment_time, central=True,
                           end_derivatives=None, N: int = 10, derivative: int = 4, layout: str = "soa",
                           stream=None):
        """mtg_cost_time_fd_batch. Returns J [B], J_plus/J_minus/grad ([K,B] soa, [B,K] aos), status."""
        aos = layout == "aos"
        if aos:
            B, Kp1, D = positions.shape
        else:
            Kp1, D, B = positions.shape
        K = Kp1 - 1
        mode = self._mode(positions)
        desc = ProblemDesc(B, K, D, N, derivative, mode, LAYOUT_AOS if aos else LAYOUT_SOA)
        shape = (B, K) if aos else (K, B)
        J = self._empty(positions, (B,))
        Jp, grad = self._empty(positions, shape), self._empty(positions, shape)
        Jm = self._empty(positions, shape) if central else None
        status = self._empty(positions, (B,), "u4")
        rc = self._lib.mtg_cost_time_fd_batch(self._h, C.byref(desc), _ptr(positions), _ptr(end_derivatives),
                                              _ptr(seg_times), _ptr(free), float(increment_time),
                                              1 if central else 0, _ptr(J), _ptr(Jp), _ptr(Jm), _ptr(grad),
                                              _ptr(status), self._stream(mode, stream))
        self._check(rc, "mtg_cost_time_fd_batch")
        return dict(J=J, J_plus=Jp, J_minus=Jm, grad=grad, status=status)

    def generate_candidates_batch(self, B, K, D, seed, first_index=0, pos_min=-10.0, pos_max=10.0, v_max=3.0, a_max=5.0,
                                  magic=6.5, layout="soa", device=None, stream=None):
        """mtg_generate_candidates_batch: (positions, seg_times) CUDA tensors, soa [K+1,D,B] / [K,B] or aos."""
        import torch

        aos = layout == "aos"
        dev = torch.device("cuda", self.device if device is None else device)
        pos = torch.empty((B, K + 1, D) if aos else (K + 1, D, B), dtype=torch.float64, device=dev)
        times = torch.empty((B, K) if aos else (K, B), dtype=torch.float64, device=dev)
        lo = (C.c_double * D)(*np.broadcast_to(np.asarray(pos_min, dtype=np.float64), (D,)))
        hi = (C.c_double * D)(*np.broadcast_to(np.asarray(pos_max, dtype=np.float64), (D,)))
        desc = ProblemDesc(B, K, D, 10, 4, MTG_MEM_DEVICE, LAYOUT_AOS if aos else LAYOUT_SOA)
        rc = self._lib.mtg_generate_candidates_batch(self._h, C.byref(desc), int(seed), int(first_index), lo, hi,
                                                     float(v_max), float(a_max), float(magic), _ptr(pos), _ptr(times),
                                                     self._stream(MTG_MEM_DEVICE, stream))
        self._check(rc, "mtg_generate_candidates_batch")
        return pos, times

    # ------------------------------------------------------------ sweep argmin
    def argmin_batch(self, cost, status=None, global_offset: int = 0, best=None, accumulate=False, stream=None):
        """mtg_argmin_batch on CUDA tensors. `best` is a 2-element int64 CUDA tensor holding the device
        struct {double cost; int64 idx}; returns it (decode with `decode_best`)."""
        import torch

        if not (_is_torch(cost) and cost.is_cuda):
            raise MtgError("argmin_batch takes CUDA tensors")
        if best is None:
            best = torch.zeros(2, dtype=torch.int64, device=cost.device)
            accumulate = False
        rc = self._lib.mtg_argmin_batch(self._h, _ptr(cost), _ptr(status), cost.numel(), int(global_offset),
                                        1 if accumulate else 0, _ptr(best), self._stream(MTG_MEM_DEVICE, stream))
        self._check(rc, "mtg_argmin_batch")
        return best

    def prepare_solve_argmin(self, positions, seg_times, end_derivatives=None, N: int = 10, derivative: int = 4,
                             layout: str = "soa", best=None, out=None, stream=None):
        """A prepared mtg_solve_argmin_batch call on fixed CUDA tensors: returns run(global_offset, accumulate),
        which is ONE ctypes call (the descriptor and the pointers are built here, once) — a launch loop driven
        from Python then costs a few microseconds of host time per step, like a C++ host."""
        aos = layout == "aos"
        if aos:
            B, Kp1, D = positions.shape
        else:
            Kp1, D, B = positions.shape
        if not (_is_torch(positions) and positions.is_cuda) or best is None:
            raise MtgError("prepare_solve_argmin takes CUDA tensors and a `best` pair")
        for nm, x in (("positions", positions), ("seg_times", seg_times)):
            self._contig(x, nm)
        out = out or {}
        desc = ProblemDesc(B, Kp1 - 1, D, N, derivative, MTG_MEM_DEVICE, LAYOUT_AOS if aos else LAYOUT_SOA)
        keep = (positions, seg_times, end_derivatives, best, dict(out))   # the tensors outlive the closure's calls
        fixed = (self._h, C.byref(desc), _ptr(positions), _ptr(end_derivatives), _ptr(seg_times),
                 _ptr(out.get("coeffs")), _ptr(out.get("cost")), _ptr(out.get("free")), _ptr(out.get("status")))
        tail = (_ptr(best), self._stream(MTG_MEM_DEVICE, stream))
        fn = self._lib.mtg_solve_argmin_batch

        def run(global_offset: int, accumulate: bool, _keep=keep, _desc=desc):
            rc = fn(*fixed, global_offset, 1 if accumulate else 0, *tail)
            if rc:
                self._check(rc, "mtg_solve_argmin_batch")

        return run

    def set_solve_overlap(self, enabled: bool):
        """mtg_set_solve_overlap: consecutive device-memory solves of a stream may overlap (see include/mtg_cuda.h
        for the contract on their inputs)."""
        self._check(self._lib.mtg_set_solve_overlap(self._h, 1 if enabled else 0), "mtg_set_solve_overlap")

    def solve_argmin_batch(self, positions, seg_times, end_derivatives=None, N: int = 10, derivative: int = 4,
                           layout: str = "soa", global_offset: int = 0, best=None, accumulate=False, out=None,
                           stream=None):
        """mtg_solve_argmin_batch on CUDA tensors: the solve with the argmin of its costs fused into the kernel.
        `out` may hold any of coeffs / cost / free / status tensors to be filled (none is required). Returns
        `best`, the 2-element int64 CUDA tensor holding the device pair {double cost; int64 idx}."""
        import torch

        aos = layout == "aos"
        if aos:
            B, Kp1, D = positions.shape
        else:
            Kp1, D, B = positions.shape
        K = Kp1 - 1
        if not (_is_torch(positions) and positions.is_cuda):
            raise MtgError("solve_argmin_batch takes CUDA tensors")
        for nm, x in (("positions", positions), ("seg_times", seg_times)):
            self._contig(x, nm)
        desc = ProblemDesc(B, K, D, N, derivative, MTG_MEM_DEVICE, LAYOUT_AOS if aos else LAYOUT_SOA)
        out = out or {}
        if best is None:
            best = torch.zeros(2, dtype=torch.int64, device=positions.device)
            accumulate = False
        rc = self._lib.mtg_solve_argmin_batch(self._h, C.byref(desc), _ptr(positions), _ptr(end_derivatives),
                                              _ptr(seg_times), _ptr(out.get("coeffs")), _ptr(out.get("cost")),
                                              _ptr(out.get("free")), _ptr(out.get("status")), int(global_offset),
                                              1 if accumulate else 0, _ptr(best),
                                              self._stream(MTG_MEM_DEVICE, stream))
        self._check(rc, "mtg_solve_argmin_batch")
        return best

    @staticmethod
    def decode_best(best):
        """(cost, idx) from the device pair (synchronises)."""
        import torch

        h = best.cpu()
        return float(h[:1].view(torch.float64)[0]), int(h[1])

    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self._lib.mtg_nccl_unique_id(self._h, buf), "mtg_nccl_unique_id")
        return buf.raw

    def nccl_init(self, unique_id: bytes, rank: int, world: int):
        self._check(self._lib.mtg_nccl_init(self._h, C.create_string_buffer(unique_id, 128), rank, world),
                    "mtg_nccl_init")

    def argmin_allgather(self, cost, status=None, global_offset: int = 0, stream=None):
        """mtg_argmin_allgather: (best_cost, best_global_idx), identical on every rank."""
        bc, bi = C.c_double(0.0), C.c_int64(-1)
        rc = self._lib.mtg_argmin_allgather(self._h, _ptr(cost), _ptr(status), cost.numel(), int(global_offset),
                                            C.byref(bc), C.byref(bi), self._stream(MTG_MEM_DEVICE, stream))
        self._check(rc, "mtg_argmin_allgather")
        return bc.value, bi.value

    def best_allgather(self, best, out=None, stream=None):
        """mtg_best_allgather: all-gather of the device pair `best` (as argmin_batch keeps it) over the
        context's NCCL communicator and the final selection on the device, no host synchronisation;
        returns the 2-element int64 CUDA tensor holding the global pair (decode with `decode_best`)."""
        import torch

        if out is None:
            out = torch.zeros(2, dtype=torch.int64, device=best.device)
        rc = self._lib.mtg_best_allgather(self._h, _ptr(best), _ptr(out), self._stream(MTG_MEM_DEVICE, stream))
        self._check(rc, "mtg_best_allgather")
        return out

    def probe_fp64_fma(self, reps: int = 5, stream=None):
        """Measured fp64 FMA throughput of this device in TFLOP/s (DFMA micro-benchmark)."""
        tf, ms = C.c_double(0.0), C.c_double(0.0)
        rc = self._lib.mtg_probe_fp64_fma(self._h, int(reps), C.byref(tf), C.byref(ms),
                                          self._stream(MTG_MEM_DEVICE, stream))
        self._check(rc, "mtg_probe_fp64_fma")
        return tf.value

    # ------------------------------------------------------------- evaluation
    def _shape_kdn(self, coeffs, aos):
        if aos:
            B, K, D, N = coeffs.shape
        else:
            K, D, N, B = coeffs.shape
        return B, K, D, N

    def _desc_for(self, coeffs, layout):
        aos = layout == "aos"
        B, K, D, N = self._shape_kdn(coeffs, aos)
        mode = self._mode(coeffs)
        # derivative_to_optimize only selects the constant tables; evaluation uses the base table
        return aos, B, K, D, N, mode, ProblemDesc(B, K, D, N, N // 2 - 1, mode,
                                                  LAYOUT_AOS if aos else LAYOUT_SOA)

    def _bvec(self, like, x, B):
        """[B] float64 vector in the memory space of `like` from a scalar or array."""
        if _is_torch(like):
            import torch

            if _is_torch(x):
                return x.to(device=like.device, dtype=torch.float64).contiguous()
            return torch.as_tensor(np.broadcast_to(np.asarray(x, dtype=np.float64), (B,)).copy(),
                                   device=like.device)
        return np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (B,)))

    def max_time_batch(self, seg_times, layout="soa", stream=None):
        aos = layout == "aos"
        B, K = seg_times.shape if aos else seg_times.shape[::-1]
        mode = self._mode(seg_times)
        desc = ProblemDesc(B, K, 1, 10, 4, mode, LAYOUT_AOS if aos else LAYOUT_SOA)
        out = self._empty(seg_times, (B,))
        self._check(self._lib.mtg_max_time_batch(self._h, C.byref(desc), _ptr(seg_times), _ptr(out),
                                                 self._stream(mode, stream)), "mtg_max_time_batch")
        return out

    def eval_range_batch(self, coeffs, seg_times, t_start, t_end, dt, derivative=0, max_samples=1024,
                         layout="soa", want_samples=True, want_times=False, want_segments=False,
                         out=None, stream=None):
        """mtg_eval_range_batch. soa: samples [S,D,B], times/segments [S,B]; aos: [B,S,D], [B,S]."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, layout)
        S = max_samples
        out = out or {}
        t0, t1, dtv = (self._bvec(coeffs, x, B) for x in (t_start, t_end, dt))
        samples = out.get("samples")
        if want_samples and samples is None:
            samples = self._empty(coeffs, (B, S, D) if aos else (S, D, B))
        times = self._empty(coeffs, (B, S) if aos else (S, B)) if want_times else None
        segs = self._empty(coeffs, (B, S) if aos else (S, B), "i4") if want_segments else None
        n = out.get("n_samples")
        if n is None:
            n = self._empty(coeffs, (B,), "i4")
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_eval_range_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times), _ptr(t0),
                                            _ptr(t1), _ptr(dtv), derivative, S, _ptr(samples), _ptr(times),
                                            _ptr(segs), _ptr(n), _ptr(status), self._stream(mode, stream))
        self._check(rc, "mtg_eval_range_batch")
        return dict(samples=samples, sampling_times=times, segment_idx=segs, n_samples=n, status=status)

    def eval_at_batch(self, coeffs, seg_times, t, derivative=0, layout="soa", stream=None):
        """mtg_eval_at_batch. t: soa [M,B] / aos [B,M]. Returns out soa [M,D,B] / aos [B,M,D]."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, layout)
        M = t.shape[1] if aos else t.shape[0]
        out = self._empty(coeffs, (B, M, D) if aos else (M, D, B))
        segs = self._empty(coeffs, (B, M) if aos else (M, B), "i4")
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_eval_at_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times), _ptr(t), M,
                                         derivative, _ptr(out), _ptr(segs), _ptr(status),
                                         self._stream(mode, stream))
        self._check(rc, "mtg_eval_at_batch")
        return dict(out=out, segment_idx=segs, status=status)

    def extrema_batch(self, coeffs, seg_times, derivative, layout="soa", want_segments=False, stream=None):
        """mtg_extrema_batch: min/max of |p^(derivative)| per trajectory (value, segment-relative time,
        segment) and optionally per segment (seg_max_value/seg_max_time: soa [K,B], aos [B,K])."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, layout)
        out = {k: self._empty(coeffs, (B,)) for k in ("min_value", "min_time", "max_value", "max_time")}
        out["min_seg"] = self._empty(coeffs, (B,), "i4")
        out["max_seg"] = self._empty(coeffs, (B,), "i4")
        out["status"] = self._empty(coeffs, (B,), "u4")
        sv = st = None
        if want_segments:
            sv = self._empty(coeffs, (B, K) if aos else (K, B))
            st = self._empty(coeffs, (B, K) if aos else (K, B))
        rc = self._lib.mtg_extrema_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times), derivative,
                                         _ptr(out["min_value"]), _ptr(out["min_time"]), _ptr(out["min_seg"]),
                                         _ptr(out["max_value"]), _ptr(out["max_time"]), _ptr(out["max_seg"]),
                                         _ptr(sv), _ptr(st), _ptr(out["status"]), self._stream(mode, stream))
        self._check(rc, "mtg_extrema_batch")
        out["seg_max_value"], out["seg_max_time"] = sv, st
        return out

    def extrema_candidates_batch(self, coeffs, seg_times, derivative, t_start=None, t_end=None, dim_mask=0,
                                 max_candidates=24, layout="soa", stream=None):
        """mtg_extrema_candidates_batch: cand_time / cand_value soa [K,MC,B] / aos [B,K,MC], n_candidates
        soa [K,B] / aos [B,K]. t_start / t_end: per-segment records shaped like seg_times, or None."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, layout)
        MC = max_candidates
        ct = self._empty(coeffs, (B, K, MC) if aos else (K, MC, B))
        cv = self._empty(coeffs, (B, K, MC) if aos else (K, MC, B))
        nc = self._empty(coeffs, (B, K) if aos else (K, B), "i4")
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_extrema_candidates_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times),
                                                    _ptr(t_start), _ptr(t_end), derivative, int(dim_mask), MC,
                                                    _ptr(ct), _ptr(cv), _ptr(nc), _ptr(status),
                                                    self._stream(mode, stream))
        self._check(rc, "mtg_extrema_candidates_batch")
        return dict(cand_time=ct, cand_value=cv, n_candidates=nc, status=status)

    def poly_real_roots_batch(self, coeffs, t_lo, t_hi, max_roots=24, layout="aos", stream=None):
        """mtg_poly_real_roots_batch: coeffs aos [B,n] / soa [n,B] (increasing powers); roots aos [B,MR] /
        soa [MR,B] ascending, n_roots [B]."""
        aos = layout == "aos"
        B, n = coeffs.shape if aos else coeffs.shape[::-1]
        mode = self._mode(coeffs)
        lo, hi = self._bvec(coeffs, t_lo, B), self._bvec(coeffs, t_hi, B)
        roots = self._empty(coeffs, (B, max_roots) if aos else (max_roots, B))
        nr = self._empty(coeffs, (B,), "i4")
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_poly_real_roots_batch(self._h, B, n, mode, LAYOUT_AOS if aos else LAYOUT_SOA, _ptr(coeffs),
                                                 _ptr(lo), _ptr(hi), max_roots, _ptr(roots), _ptr(nr), _ptr(status),
                                                 self._stream(mode, stream))
        self._check(rc, "mtg_poly_real_roots_batch")
        return dict(roots=roots, n_roots=nr, status=status)

    def soft_constraint_batch(self, coeffs, seg_times, derivatives, limits, weight, maximum_cost, layout="soa",
                              want_violations=True, stream=None):
        """mtg_soft_constraint_batch (CUDA tensors): cost [B], violations [n,B]."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, layout)
        n = len(derivatives)
        der = (C.c_int32 * max(n, 1))(*[int(x) for x in derivatives])
        lim = (C.c_double * max(n, 1))(*[float(x) for x in limits])
        cost = self._empty(coeffs, (B,))
        viol = self._empty(coeffs, (n, B)) if want_violations else None
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_soft_constraint_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times), n, der, lim,
                                                 float(weight), float(maximum_cost), _ptr(cost), _ptr(viol),
                                                 _ptr(status), self._stream(mode, stream))
        self._check(rc, "mtg_soft_constraint_batch")
        return dict(cost=cost, violations=viol, status=status)

    def control_points_batch(self, seg_times, coeffs=None, derivatives=None, positions=None, radii=None, N=10,
                             layout="soa", stream=None):
        """mtg_control_points_batch. coeffs soa [K,D,N,B] / aos [B,K,D,N] or derivatives soa [K+1,h,D,B] /
        aos [B,K+1,h,D]. Returns control_points soa [K,N,D,B] / aos [B,K,N,D] and, with positions + radii
        (D = 3), tube / cap_start / cap_end soa [K,N-2,B] / aos [B,K,N-2], sphere [K,B] / [B,K], max_value,
        feasible [B]."""
        aos = layout == "aos"
        like = derivatives if derivatives is not None else coeffs
        B, K = seg_times.shape if aos else seg_times.shape[::-1]
        D = like.shape[-1] if aos and derivatives is not None else (like.shape[2] if derivatives is not None
                                                                    else (like.shape[2] if aos else like.shape[1]))
        mode = self._mode(like)
        desc = ProblemDesc(B, K, D, N, N // 2 - 1, mode, LAYOUT_AOS if aos else LAYOUT_SOA)
        cps = self._empty(like, (B, K, N, D) if aos else (K, N, D, B))
        con = positions is not None and radii is not None
        out = dict(control_points=cps, status=self._empty(like, (B,), "u4"))
        tube = cs = ce = sph = mx = fe = None
        if con:
            tube, cs, ce = (self._empty(like, (B, K, N - 2) if aos else (K, N - 2, B)) for _ in range(3))
            sph = self._empty(like, (B, K) if aos else (K, B))
            mx = self._empty(like, (B,))
            fe = self._empty(like, (B,), "u1")
        rc = self._lib.mtg_control_points_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(derivatives),
                                                _ptr(seg_times), _ptr(positions), _ptr(radii), _ptr(cps), _ptr(tube),
                                                _ptr(cs), _ptr(ce), _ptr(sph), _ptr(mx), _ptr(fe),
                                                _ptr(out["status"]), self._stream(mode, stream))
        self._check(rc, "mtg_control_points_batch")
        out.update(tube=tube, cap_start=cs, cap_end=ce, sphere=sph, max_value=mx, feasible=fe)
        return out

    # ------------------------------------------------------------- non-linear objective (N1), CUDA tensors
    def cost_derivative_batch(self, positions, seg_times, free, end_derivatives=None, N=10, derivative=4,
                              layout="soa", want_diag=False, stream=None):
        """mtg_cost_derivative_batch: J_d [B], grad shaped like `free`, diag soa [K-1,NF,B] / aos [B,K-1,NF]."""
        aos = layout == "aos"
        if aos:
            B, Kp1, D = positions.shape
        else:
            Kp1, D, B = positions.shape
        K = Kp1 - 1
        desc = ProblemDesc(B, K, D, N, derivative, MTG_MEM_DEVICE, LAYOUT_AOS if aos else LAYOUT_SOA)
        J = self._empty(positions, (B,))
        grad = self._empty(positions, tuple(free.shape))
        NF = N // 2 - 1
        diag = self._empty(positions, (B, K - 1, NF) if aos else (K - 1, NF, B)) if want_diag else None
        status = self._empty(positions, (B,), "u4")
        rc = self._lib.mtg_cost_derivative_batch(self._h, C.byref(desc), _ptr(positions), _ptr(end_derivatives),
                                                 _ptr(seg_times), _ptr(free), _ptr(J), _ptr(grad), _ptr(diag),
                                                 _ptr(status), self._stream(MTG_MEM_DEVICE, stream))
        self._check(rc, "mtg_cost_derivative_batch")
        return dict(J_d=J, grad=grad, diag=diag, status=status)

    def soft_constraint_gradient_batch(self, coeffs, seg_times, derivatives, limits, weight, maximum_cost, increment,
                                       central=True, want_grad=True, stream=None):
        """mtg_soft_constraint_gradient_batch (AoS): J_sc [B], grad [B, D, K-1, NF]."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, "aos")
        n = len(derivatives)
        der = (C.c_int32 * n)(*[int(x) for x in derivatives])
        lim = (C.c_double * n)(*[float(x) for x in limits])
        J = self._empty(coeffs, (B,))
        grad = self._empty(coeffs, (B, D, K - 1, N // 2 - 1)) if want_grad else None
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_soft_constraint_gradient_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times), n, der,
                                                          lim, float(weight), float(maximum_cost), float(increment),
                                                          1 if central else 0, _ptr(J), _ptr(grad), _ptr(status),
                                                          self._stream(mode, stream))
        self._check(rc, "mtg_soft_constraint_gradient_batch")
        return dict(J_sc=J, grad=grad, status=status)

    def nl_descent_batch(self, positions, seg_times, free, derivatives=(), limits=(), w_d=1.0, w_sc=1.0,
                         soft_weight=100.0, maximum_cost=1e12, increment=0.05, step=0.5, precondition=True,
                         iterations=10, end_derivatives=None, N=10, derivative=4, want_history=True, stream=None):
        """mtg_nl_descent_batch (AoS): updates `free` in place; returns coeffs [B,K,D,N], history [it+2,3,B]
        (J_d, J_sc, accepted of every trial point, then of the returned point)."""
        B, Kp1, D = positions.shape
        K = Kp1 - 1
        desc = ProblemDesc(B, K, D, N, derivative, MTG_MEM_DEVICE, LAYOUT_AOS)
        n = len(derivatives)
        der = (C.c_int32 * max(n, 1))(*[int(x) for x in derivatives])
        lim = (C.c_double * max(n, 1))(*[float(x) for x in limits])
        coeffs = self._empty(positions, (B, K, D, N))
        hist = self._empty(positions, (iterations + 2, 3, B)) if want_history else None
        status = self._empty(positions, (B,), "u4")
        rc = self._lib.mtg_nl_descent_batch(self._h, C.byref(desc), _ptr(positions), _ptr(end_derivatives),
                                            _ptr(seg_times), _ptr(free), n, der, lim, float(w_d), float(w_sc),
                                            float(soft_weight), float(maximum_cost), float(increment), float(step),
                                            1 if precondition else 0, int(iterations), _ptr(coeffs), _ptr(hist),
                                            _ptr(status), self._stream(MTG_MEM_DEVICE, stream))
        self._check(rc, "mtg_nl_descent_batch")
        return dict(coeffs=coeffs, history=hist, free=free, status=status)

    def collision_cost_batch(self, coeffs, seg_times, grid, origin, res, min_bound, max_bound, dt, epsilon=0.5,
                             robot_radius=0.5, multiplier=1.0, layout="soa", want_grad=True, stream=None):
        """mtg_collision_cost_batch (CUDA tensors). grid: [nx,ny,nz] float64 CUDA tensor of distances (m).
        Returns J_c [B], grad soa [3,K-1,NF,B] / aos [B,3,K-1,NF], in_collision, n_checks [B]."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, layout)
        NF = N // 2 - 1
        size = (C.c_int32 * 3)(*[int(x) for x in grid.shape])
        org = (C.c_int32 * 3)(*[int(x) for x in origin])
        lo = (C.c_double * 3)(*[float(x) for x in min_bound])
        hi = (C.c_double * 3)(*[float(x) for x in max_bound])
        J = self._empty(coeffs, (B,))
        grad = self._empty(coeffs, (B, 3, K - 1, NF) if aos else (3, K - 1, NF, B)) if want_grad else None
        col = self._empty(coeffs, (B,), "u1")
        chk = self._empty(coeffs, (B,), "i4")
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_collision_cost_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times), _ptr(grid), size,
                                                org, float(res), lo, hi, float(dt), float(epsilon), float(robot_radius),
                                                float(multiplier), _ptr(J), _ptr(grad), _ptr(col), _ptr(chk),
                                                _ptr(status), self._stream(mode, stream))
        self._check(rc, "mtg_collision_cost_batch")
        return dict(J_c=J, grad=grad, in_collision=col, n_checks=chk, status=status)

    # ------------------------------------------------------------- composition / I/O (N3)
    def vertex_at_time_batch(self, coeffs, seg_times, t, max_derivative_order, layout="soa", stream=None):
        """mtg_vertex_at_time_batch: out soa [M+1,D,B] / aos [B,M+1,D], segment_idx [B]."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, layout)
        M = int(max_derivative_order)
        tv = self._bvec(coeffs, t, B)
        out = self._empty(coeffs, (B, M + 1, D) if aos else (M + 1, D, B))
        seg = self._empty(coeffs, (B,), "i4")
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_vertex_at_time_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times), _ptr(tv), M,
                                                _ptr(out), _ptr(seg), _ptr(status), self._stream(mode, stream))
        self._check(rc, "mtg_vertex_at_time_batch")
        return dict(out=out, segment_idx=seg, status=status)

    def pick_dimensions_batch(self, coeffs_a, pick, coeffs_b=None, layout="soa", stream=None):
        """mtg_pick_dimensions_batch: output dimension q = dimension pick[q] of a (< D_a) or pick[q] - D_a of b."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs_a, layout)
        Db = 0 if coeffs_b is None else (coeffs_b.shape[2] if aos else coeffs_b.shape[1])
        n_out = len(pick)
        arr = (C.c_int32 * n_out)(*[int(x) for x in pick])
        out = self._empty(coeffs_a, (B, K, n_out, N) if aos else (K, n_out, N, B))
        rc = self._lib.mtg_pick_dimensions_batch(self._h, C.byref(desc), _ptr(coeffs_a), Db, _ptr(coeffs_b), n_out, arr,
                                                 _ptr(out), self._stream(mode, stream))
        self._check(rc, "mtg_pick_dimensions_batch")
        return out

    def concat_segments_batch(self, coeffs_list, times_list, layout="soa", stream=None):
        """mtg_concat_segments_batch: addTrajectories over whole batches -> (coeffs, seg_times)."""
        aos = layout == "aos"
        first = coeffs_list[0]
        if aos:
            B, _, D, N = first.shape
            Ks = [c.shape[1] for c in coeffs_list]
        else:
            _, D, N, B = first.shape
            Ks = [c.shape[0] for c in coeffs_list]
        mode = self._mode(first)
        Kt = sum(Ks)
        oc = self._empty(first, (B, Kt, D, N) if aos else (Kt, D, N, B))
        ot = self._empty(first, (B, Kt) if aos else (Kt, B))
        n = len(coeffs_list)
        karr = (C.c_int32 * n)(*Ks)
        carr = (C.c_void_p * n)(*[_ptr(c) for c in coeffs_list])
        tarr = (C.c_void_p * n)(*[_ptr(t) for t in times_list])
        rc = self._lib.mtg_concat_segments_batch(self._h, B, D, N, mode, LAYOUT_AOS if aos else LAYOUT_SOA, n, karr,
                                                 carr, tarr, _ptr(oc), _ptr(ot), self._stream(mode, stream))
        self._check(rc, "mtg_concat_segments_batch")
        return oc, ot

    def compute_cost_batch(self, coeffs, seg_times, derivative=4, layout="soa", stream=None):
        """mtg_compute_cost_batch: computeCost (LIN_I:113-130) of given coefficients."""
        aos, B, K, D, N, mode, _ = self._desc_for(coeffs, layout)
        desc = ProblemDesc(B, K, D, N, derivative, mode, LAYOUT_AOS if aos else LAYOUT_SOA)
        cost = self._empty(coeffs, (B,))
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_compute_cost_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times), _ptr(cost),
                                              _ptr(status), self._stream(mode, stream))
        self._check(rc, "mtg_compute_cost_batch")
        return dict(cost=cost, status=status)

    def sample_dump_batch(self, coeffs, seg_times, dt, max_rows, layout="soa", stream=None):
        """mtg_sample_dump_batch: rows [B, max_rows, 5 D + 2], n_rows [B]."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, layout)
        rows = self._empty(coeffs, (B, max_rows, 5 * D + 2))
        n = self._empty(coeffs, (B,), "i4")
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_sample_dump_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times), float(dt),
                                             int(max_rows), _ptr(rows), _ptr(n), _ptr(status),
                                             self._stream(mode, stream))
        self._check(rc, "mtg_sample_dump_batch")
        return dict(rows=rows, n_rows=n, status=status)

    def feasibility_batch(self, coeffs, seg_times, t_start, t_end, dt, v_max, a_max, positions=None,
                          radii=None, max_samples=1024, layout="soa", want_samples=False, want_flags=True,
                          out=None, stream=None):
        """mtg_feasibility_batch. positions soa [K+1,3,B] / aos [B,K+1,3]; radii soa [K,2,B] / aos [B,K,2]."""
        aos, B, K, D, N, mode, desc = self._desc_for(coeffs, layout)
        S = max_samples
        out = out or {}
        t0, t1, dtv = (self._bvec(coeffs, x, B) for x in (t_start, t_end, dt))
        samples = out.get("samples")
        if want_samples and samples is None:
            samples = self._empty(coeffs, (B, S, D) if aos else (S, D, B))
        flags = out.get("flags")
        if want_flags and flags is None:
            flags = self._empty(coeffs, (B, S) if aos else (S, B), "u1")
        max_v, max_a = self._empty(coeffs, (B,)), self._empty(coeffs, (B,))
        feasible = self._empty(coeffs, (B,), "u1")
        n = self._empty(coeffs, (B,), "i4")
        status = self._empty(coeffs, (B,), "u4")
        rc = self._lib.mtg_feasibility_batch(self._h, C.byref(desc), _ptr(coeffs), _ptr(seg_times),
                                             _ptr(positions), _ptr(radii), float(v_max), float(a_max),
                                             _ptr(t0), _ptr(t1), _ptr(dtv), S, _ptr(samples), _ptr(flags),
                                             _ptr(max_v), _ptr(max_a), _ptr(feasible), _ptr(n), _ptr(status),
                                             self._stream(mode, stream))
        self._check(rc, "mtg_feasibility_batch")
        return dict(samples=samples, flags=flags, max_v=max_v, max_a=max_a, feasible=feasible, n_samples=n,
                    status=status)
