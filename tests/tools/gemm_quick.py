"""IDX_GEMM_REPS=20 python -m tests.tools.gemm_quick : representative CFM/BigVGAN shapes, TC back end only."""
import numpy as np
from indextts_b200.engine import Engine
e = Engine(0)
rng = np.random.default_rng(0)
for (B, Tin, K, N, taps, res) in [(2, 1741, 512, 1536, 1, 0), (2, 1741, 1536, 512, 1, 0), (2, 1741, 1536, 512, 1, 1),
                                  (2, 1741, 512, 512, 1, 1), (2, 1745, 512, 1024, 5, 0), (1, 14080, 384, 384, 7, 1)]:
    A = rng.standard_normal((B, Tin, K)).astype(np.float32)
    wk = (rng.standard_normal((N, taps * K)) / np.sqrt(taps * K)).astype(np.float32)
    M = Tin - (taps - 1) if taps == 5 else Tin
    pad = 0 if taps == 5 else (taps - 1) // 2
    r = np.zeros((B, M, N), np.float32) if res else None
    print("res" if res else "   ", end=" ", flush=True)
    e.debug_conv_gemm(A, wk, taps, 1, pad, M=M, res=r, backend=2)
e.close()
