"""Cycle stamps of CTA 0 of one k6 linear launch (needs a build with GPNERF_NVCC_EXTRA=-DGPNERF_K6_TRACE)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import gpnerf_b200  # noqa
from gpnerf_b200 import train, _lib
dev = torch.device("cuda", 0)
k = train._Kernels(dev, train.PREC_TRAIN_TF32)
lib = _lib.load()
P = 262144
K, N = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 64
X = torch.randn(P, K, device=dev); W = torch.randn(N, K, device=dev) * 0.1; b = torch.randn(N, device=dev)
Y = torch.empty(P, N, device=dev)
fn = lib._lib.gpnerf_debug_k6_trace if hasattr(lib, "_lib") else C.CDLL(os.path.join(ROOT, "gp-nerf_b200", "libgpnerf_b200.so")).gpnerf_debug_k6_trace
fn.argtypes = [C.c_void_p, C.c_int]; fn.restype = C.c_int
for _ in range(300):
    k.linear(X, K, K, W, K, N, Y, N, P, bias=b, epi=1)
torch.cuda.synchronize()
fn(None, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); k.linear(X, K, K, W, K, N, Y, N, P, bias=b, epi=1); e1.record()
torch.cuda.synchronize()
print("K", K, "N", N, "us", e0.elapsed_time(e1) * 1e3, "GB/s", P * (K + N) * 4 / e0.elapsed_time(e1) / 1e6)
buf = (C.c_longlong * 2048)()
n = fn(buf, 1)
t = list(buf[:n])
names = ["top", "mma(i-1) done", "loads issued", "tile landed", "sync", "mma issued + epilogue"]
print("deltas", [t[i + 1] - t[i] for i in range(min(n - 1, 60))])
print("stamps", n, "first->last cycles", t[-1] - t[0], "tail deltas", [t[i + 1] - t[i] for i in range(max(0, n - 8), n - 1)])
lib2 = C.CDLL(os.path.join(ROOT, "gp-nerf_b200", "libgpnerf_b200.so"))
sp = (C.c_ulonglong * 1024)()
lib2.gpnerf_debug_k6_span(sp)
st = [sp[2 * i] for i in range(296)]; en = [sp[2 * i + 1] for i in range(296)]
t0 = min(st)
print("block starts (ns): min 0 max", max(st) - t0, " ends: min", min(en) - t0, "max", max(en) - t0, " durations min/max", min(e - s_ for s_, e in zip(st, en)), max(e - s_ for s_, e in zip(st, en)))
ref = torch.nn.functional.elu(X @ W.t() + b)
print("max err", float((Y - ref).abs().max()))
