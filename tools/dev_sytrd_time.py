import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arc_b200 import sweep, _lib
N = int(sys.argv[1]); nb = int(sys.argv[2]); cl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rng = np.random.default_rng(0)
A = rng.standard_normal((nb, N, N)); A = A + A.transpose(0, 2, 1)
lib = _lib.load(); dev = torch.device("cuda", 0)
d_A = torch.as_tensor(A).to(dev)
outs = [torch.zeros((nb, N), dtype=torch.float64, device=dev) for _ in range(3)]
wbytes = int(lib.arcb200_sweep_workspace_bytes(N, nb, N)); work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
def sy():
    _lib.check(lib.arcb200_sytrd_batched(0, _lib.stream_ptr(0), N, nb, _lib.ptr(d_A), _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(work), wbytes, cl, 0), "sytrd")
sy(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sy(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("N=%d nb=%d cluster=%d: %.2f ms  (%.3f ms/matrix, %.0f GB/s algorithmic)" % (N, nb, cl, ms, ms / nb, 4 / 3 * N ** 3 * nb / ms / 1e6))
