import os, sys
os.environ["DLV3P_GEMM_DBG_ENABLE"] = "1"
sys.path.insert(0, "/root/repo")
import torch
from deeplabv3plus_keras_b200 import ops
dev, bf = "cuda", torch.bfloat16
REP = 20
def timeit(fn, reps=5):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(REP): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / REP)
    ts.sort(); return ts[len(ts)//2]*1e3
M, N, K = 16384, 736, 736
a = torch.randn((M, K), device=dev).to(bf); dy = torch.randn((M, N), device=dev).to(bf)
dw = torch.zeros((K, N), device=dev)
b = torch.randn((N, K), device=dev).to(bf); c = torch.empty((M, N), device=dev, dtype=bf); stats = torch.zeros(2*N, device=dev)
for mode, name in [(0,"full"),(8,"no C stores"),(2,"no epilogue"),(3,"MMA only"),(6,"loads only"),(7,"empty")]:
    os.environ["DLV3P_GEMM_DBG"] = str(mode)
    print(f"{name:12s} wgrad {timeit(lambda: ops.gemm_wgrad_bf16(a, dy, dw, M, K, N)):6.2f} | gemm {timeit(lambda: ops.gemm_bf16(a, b, M, N, K, c)):6.2f} | gemm+stats {timeit(lambda: ops.gemm_bf16(a, b, M, N, K, c, col_stats=stats)):6.2f} us", flush=True)
os.environ["DLV3P_GEMM_DBG"] = "0"
# other hot kernels of the middle flow, in-graph isolated
x = torch.randn(16,32,32,736, device=dev).to(bf); y = torch.empty_like(x); g = torch.randn(16,32,32,736, device=dev).to(bf)
w = torch.randn(3,3,736, device=dev); dwg = torch.zeros(3,3,736, device=dev)
sc, sh, mu, isd = (torch.rand(736, device=dev)+0.5 for _ in range(4)); red = torch.zeros(2*736, device=dev)
sums = torch.zeros(2*736, device=dev); ops.bn_stats(x, M, 736, sums)
gam, bet, mm, mv = (torch.rand(736, device=dev)+0.5 for _ in range(4))
print("dw bn_fwd   ", round(timeit(lambda: ops.dwconv3x3_bn_fwd(x, w, sums, gam, bet, mm, mv, M, 1e-3, 0.99, 1, 1, sc, sh, mu, isd, out=y)),2))
print("dw dgrad_bnred", round(timeit(lambda: ops.dwconv3x3_dgrad_bnred(g, w, x.shape, x, sc, sh, 1, mu, isd, red, out=y)),2))
print("dw wgrad(aff)", round(timeit(lambda: ops.dwconv3x3_wgrad(x, g, dwg, 1, (1,1), in_scale=sc, in_shift=sh, in_act=1)),2))
print("bn_bwd_apply ", round(timeit(lambda: ops.bn_bwd_apply(g, x, sc, sh, mu, isd, 0, red, M, 736, y)),2))
print("bn_bwd_reduce", round(timeit(lambda: ops.bn_bwd_reduce(g, x, sc, sh, mu, isd, 0, M, 736, red)),2))
print("bn_train_apply+add", round(timeit(lambda: ops.bn_train_apply(x, M, 736, sums, gam, bet, mm, mv, M, 1e-3, 0.99, 1, 0, y, sc, sh, mu, isd, addend=g)),2))
