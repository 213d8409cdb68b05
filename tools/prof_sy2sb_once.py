"""One band reduction of a batch for ncu (launch list / dram bytes): python tools/prof_sy2sb_once.py N batch"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arc_b200 import _lib
N, nb = int(sys.argv[1]), int(sys.argv[2])
lib = _lib.load()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(1)
A = rng.standard_normal((min(nb, 4), N, N)); A = A + A.transpose(0, 2, 1)
d_A = torch.as_tensor(A).to(dev).repeat((nb + len(A) - 1) // len(A), 1, 1)[:nb].contiguous()
wbytes = int(lib.arcb200_sweep_workspace_bytes(N, nb, N))
work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
band = torch.zeros((nb, N, 33), dtype=torch.float64, device=dev)
_lib.check(lib.arcb200_sy2sb_batched(0, _lib.stream_ptr(0), N, nb, _lib.ptr(d_A), _lib.ptr(band), _lib.ptr(work), wbytes, 0), "sy2sb")
torch.cuda.synchronize()
