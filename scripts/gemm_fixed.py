"""Diagnostics: fixed cost of one 2-CTA GEMM launch.  `empty` = barrier hand-offs only (DLV3P_GEMM_DBG=7), 20 launches in one
CUDA graph per timing; M = 256 is one pair tile on one pair, M = 16384 is 2.59 waves of tiles on 74 pairs."""
import os
import sys

os.environ["DLV3P_GEMM_DBG_ENABLE"] = "1"
os.environ.setdefault("DLV3P_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             "deeplabv3plus_keras_b200", "libdlv3p_diag.so"))   # DLV3P_DIAG=1 bash csrc/build.sh
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev, bf = "cuda", torch.bfloat16
REP = 20


def timeit(fn, reps=5):
    """REP launches captured in ONE CUDA graph (no host launch cost in the measurement), replayed `reps` times."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(REP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / REP)
    ts.sort()
    return ts[len(ts) // 2] * 1e3


N = K = 736
for M in (256, 512, 4096, 9472, 16384, 18944, 37888):
    a = torch.randn((M, K), device=dev).to(bf)
    b = torch.randn((N, K), device=dev).to(bf)
    c = torch.empty((M, N), device=dev, dtype=bf)
    row = f"M={M:6d} pair-tiles={-(-M // 256) * 3:4d}"
    for mode, name in [(7, "empty"), (3, "MMA only"), (6, "loads only"), (0, "full")]:
        os.environ["DLV3P_GEMM_DBG"] = str(mode)
        row += f" | {name} {timeit(lambda: ops.gemm_bf16(a, b, M, N, K, c)):6.2f} us"
    os.environ["DLV3P_GEMM_DBG"] = "0"
    print(row, flush=True)
a = torch.zeros(1 << 20, device=dev)
print(f"torch elementwise add on 4 MB (launch-bound reference): {timeit(lambda: a.add_(1.0)):6.2f} us")
