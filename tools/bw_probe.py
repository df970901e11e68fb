#!/usr/bin/env python
"""HBM bandwidth probe: write-only (fill), read-only (sum) and copy streams over a 4 GiB buffer.
The sweep kernels are write streams; MEASURED_PEAKS.json holds the COPY figure (read + write bytes)."""
import json
import torch


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


n = 1 << 29  # doubles: 4 GiB
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")
nbytes = n * 8
out = {}
out["fill_write_only_GBs"] = nbytes / timeit(lambda: a.fill_(1.0)) / 1e9
out["memset_write_only_GBs"] = nbytes / timeit(lambda: a.zero_()) / 1e9
out["sum_read_only_GBs"] = nbytes / timeit(lambda: a.sum()) / 1e9
out["copy_read_plus_write_GBs"] = 2 * nbytes / timeit(lambda: b.copy_(a)) / 1e9
print(json.dumps(out))
