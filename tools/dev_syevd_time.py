"""Time the full batched eigensolver (all eigenvectors) on device buffers.
Usage: python tools/dev_syevd_time.py N batch   (ARCB200_TWO_STAGE=0/1 to force a path)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arc_b200 import _lib
N = int(sys.argv[1]); nb = int(sys.argv[2])
rng = np.random.default_rng(0)
lib = _lib.load(); dev = torch.device("cuda", 0)
A = torch.randn((nb, N, N), dtype=torch.float64, device=dev); A = A + A.transpose(1, 2).contiguous()
w = torch.zeros((nb, N), dtype=torch.float64, device=dev); Z = torch.zeros((nb, N, N), dtype=torch.float64, device=dev)
wbytes = int(lib.arcb200_sweep_workspace_bytes(N, nb, N)); work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
def f(): _lib.check(lib.arcb200_syevd_batched(0, _lib.stream_ptr(0), N, nb, _lib.ptr(A), _lib.ptr(w), _lib.ptr(Z), _lib.ptr(work), wbytes, 0), "syevd")
f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
w0 = torch.linalg.eigvalsh(A[0]).cpu().numpy()
err = np.abs(w[0].cpu().numpy() - w0).max() / np.abs(w0).max()
print("syevd N=%d nb=%d two_stage=%s: %.1f ms (%.3f ms/matrix, %.0f solves/s) eig rel err %.1e" % (N, nb, os.environ.get("ARCB200_TWO_STAGE"), ms, ms / nb, nb / ms * 1e3, err))
