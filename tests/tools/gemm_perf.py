"""GPU-box diagnostic: per-launch time of the GEMM back ends on the pipeline's shapes.
    IDX_GEMM_REPS=20 python -m tests.tools.gemm_perf"""
import numpy as np
from indextts_b200.engine import Engine

SHAPES = [  # B, Tin, K, N, taps, dil, pad
    (2, 1741, 512, 1536, 1, 1, 0),    # DiT wqkv
    (2, 1741, 512, 3072, 1, 1, 0),    # DiT w1|w3
    (2, 1741, 1536, 512, 1, 1, 0),    # DiT w2
    (2, 1741, 512, 1024, 5, 1, 2),    # WN in_layer k5
    (16, 1741, 64, 1741, 1, 1, 0),    # attention S = QK^T (weights per batch not modelled: same cost)
    (16, 1741, 1744, 64, 1, 1, 0),    # attention O = PV
    (1, 3520, 768, 768, 11, 1, 5),    # BigVGAN stage 0 conv k11
    (1, 14080, 384, 384, 7, 3, 9),    # BigVGAN stage 1 conv k7 d3
    (1, 56320, 96, 96, 11, 5, 25),    # BigVGAN stage 3
    (1, 225280, 24, 24, 11, 1, 5),    # BigVGAN stage 5
]
e = Engine(0)
rng = np.random.default_rng(0)
for (B, Tin, K, N, taps, dil, pad) in SHAPES:
    A = rng.standard_normal((B, Tin, K)).astype(np.float32)
    wk = (rng.standard_normal((N, taps * K)) / np.sqrt(taps * K)).astype(np.float32)
    for backend in (2, 1):
        e.debug_conv_gemm(A, wk, taps, dil, pad, backend=backend)
e.close()
