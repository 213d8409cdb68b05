"""One-stage (ARCB200_TWO_STAGE=0) against two-stage (=1) full eigensolve (arcb200_syevd_batched, all
eigenvectors) over a range of N, to place the switch.   python tools/ab_crossover.py [batch]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arc_b200 import _lib
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
lib = _lib.load()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(3)
for N in (128, 192, 256, 320, 410, 512, 651, 800, 1000):
    A = rng.standard_normal((2, N, N)); A = A + A.transpose(0, 2, 1)
    wref = np.linalg.eigvalsh(A[0])
    d_A0 = torch.as_tensor(A).to(dev).repeat(nb // 2, 1, 1).contiguous()
    d_A = d_A0.clone()
    d_w = torch.empty((nb, N), dtype=torch.float64, device=dev)
    d_Z = torch.empty((nb, N, N), dtype=torch.float64, device=dev)
    wbytes = 0
    for env in ("0", "1"):
        os.environ["ARCB200_TWO_STAGE"] = env
        wbytes = max(wbytes, int(lib.arcb200_sweep_workspace_bytes(N, nb, N)))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    out = {"N": N, "batch": nb}
    for tag, env in (("one_stage", "0"), ("two_stage", "1")):
        os.environ["ARCB200_TWO_STAGE"] = env
        def run():
            d_A.copy_(d_A0)
            _lib.check(lib.arcb200_syevd_batched(0, _lib.stream_ptr(0), N, nb, _lib.ptr(d_A), _lib.ptr(d_w), _lib.ptr(d_Z), _lib.ptr(work), wbytes, 0), "syevd")
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        w = d_w[0].cpu().numpy(); Z = d_Z[0].cpu().numpy()          # Z[i] = eigenvector i
        res = float(np.abs(A[0] @ Z.T - Z.T * w[None, :]).max() / np.abs(wref).max())
        out[tag] = {"ms": round(ms, 3), "solves_per_s": round(nb / ms * 1e3, 1), "eig_err": float(np.abs(w - wref).max() / np.abs(wref).max()), "resid": res}
    print(json.dumps(out), flush=True)
