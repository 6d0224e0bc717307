"""setuppartition at C3 scale (N = 1M, levels = 13): host mirror vs the level-stepped device path; checks they agree bit for bit."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import synth

N, levels = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 13
X = synth.uniform_points(25, N, [-5.0, -10.0], [5.0, 10.0])
h = P.Handle(0)
P.setuppartition_device(X[:5000], 4, handle=h)      # warm-up (context, CUB temp storage)
t0 = time.perf_counter()
rd, _, id_ = P.setuppartition_device(X, levels, handle=h)
t1 = time.perf_counter()
rh, _, ih = P.setuppartition(X, levels)
t2 = time.perf_counter()
same = np.array_equal(rd.hps_v, rh.hps_v) and np.array_equal(rd.hps_c, rh.hps_c) and all(np.array_equal(a, b) for a, b in zip(id_, ih))
print(f"setuppartition N={N} levels={levels}: device path {1e3 * (t1 - t0):.1f} ms (incl. H2D of X, {levels - 1} host svd rounds, D2H of the ids), "
      f"host mirror {1e3 * (t2 - t1):.1f} ms, bit-identical: {same}, launches {h.launch_count()}")
assert same
