"""One GEMM shape for ncu: python -m tests.tools.gemm_one B Tin K N taps"""
import sys
import numpy as np
from indextts_b200.engine import Engine
B, Tin, K, N, taps = (int(x) for x in sys.argv[1:6])
e = Engine(0)
rng = np.random.default_rng(0)
A = rng.standard_normal((B, Tin, K)).astype(np.float32)
wk = (rng.standard_normal((N, taps * K)) / np.sqrt(taps * K)).astype(np.float32)
for _ in range(3):
    e.debug_conv_gemm(A, wk, taps, 1, (taps - 1) // 2, backend=2)
e.close()
