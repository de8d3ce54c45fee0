"""Five device-evaluated predicates over one int32 column of 10M rows (ncu target for where_kernel and the
passing-row list kernels; profiles/README.md)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import DeviceShard  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
s = DeviceShard(16, "l2", capacity=rows)
s.add_synthetic(1, 0, rows, False)
s.set_column(0, np.random.default_rng(1).integers(0, 100, rows).astype(np.int32))
s.set_column(1, np.random.default_rng(2).integers(0, 100, rows).astype(np.int32))
for cut in (1, 10, 50, 10, 10):
    t0 = time.perf_counter()
    f = s.where([(0, "<", cut), (1, ">=", 5)])
    dt = time.perf_counter() - t0
    print(f"bucket < {cut} and other >= 5: passing {f.passing}, {dt * 1e3:.3f} ms host wall")
    f.close()
s.close()
