"""Standalone Gram kernel (constructkernelmatrix) timing: python tools/gram_time.py [n] ; PMK_GRAM_NO_MIRROR=1 for the every-tile-evaluates mode."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
X = synth.uniform_points(25, n, [-5.0, -10.0], [5.0, 10.0])
th = P.GaussianKernel1DType(408.0)
kp = th.params
h = _lib.Handle(0)
K = np.empty((n, n), order="F")
ts = []
for _ in range(4):
    h.check(_lib.lib().pmk_gram(h.raw, 2, n, _lib.ptr(X), th.kernel_id, _lib.ptr(kp), 1, 0.0, _lib.ptr(K)))
    ts.append(float(h.timings()[_lib.T_GRAM]))
t = min(ts[1:])
print(f"k_gram n={n} mirror={'off' if os.environ.get('PMK_GRAM_NO_MIRROR') else 'on'}: {t:.4f} ms = {8.0 * n * n / t / 1e6:.0f} GB/s written, symmetric: {np.array_equal(K, K.T)}")
