"""IDX_GEMM_REPS=20 python -m tests.tools.gemm_quick : three representative shapes, TC back end only."""
import numpy as np
from indextts_b200.engine import Engine
e = Engine(0)
rng = np.random.default_rng(0)
for (B, Tin, K, N, taps) in [(2, 1741, 512, 1536, 1), (2, 1741, 1536, 512, 1), (1, 14080, 384, 384, 7)]:
    A = rng.standard_normal((B, Tin, K)).astype(np.float32)
    wk = (rng.standard_normal((N, taps * K)) / np.sqrt(taps * K)).astype(np.float32)
    e.debug_conv_gemm(A, wk, taps, 1, (taps - 1) // 2, backend=2)
e.close()
