"""Developer micro-benchmark: C2 radial table (BASELINE.json configs[1]) on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import arc_b200
from arc_b200 import radial


def c2_workload(atom, nmin=20, nmax=120, lmax=20):
    states = [(n, l, j) for n in range(nmin, nmax + 1) for l in range(0, min(lmax, n - 1) + 1)
              for j in (l - .5, l + .5) if j > 0]
    idx = {s: i for i, s in enumerate(states)}
    E = np.array([atom.getEnergy(*s) for s in states])
    by_lj = {}
    for i, (n, l, j) in enumerate(states):
        by_lj.setdefault((l, j), []).append(i)
    pairs = []
    keys = sorted(by_lj)
    for (l1, j1) in keys:
        for (l2, j2) in keys:
            dl = l2 - l1
            if dl == 1 and abs(j1 - j2) < 1.1:
                power = 1
            elif (dl == 2 or (dl == 0 and (j2 > j1 or (j1 == j2)))) and abs(j1 - j2) < 2.1:
                power = 2
            else:
                continue
            A, B = np.array(by_lj[(l1, j1)]), np.array(by_lj[(l2, j2)])
            a, b = np.meshgrid(A, B, indexing="ij")
            a, b = a.ravel(), b.ravel()
            if (l1, j1) == (l2, j2):
                keep = a <= b
                a, b = a[keep], b[keep]
            swap = E[a] > E[b]
            lo, hi = np.where(swap, b, a), np.where(swap, a, b)
            pairs.append(np.stack([lo, hi, np.full_like(lo, power)], axis=1))
    return states, np.concatenate(pairs).astype(np.int32)


if __name__ == "__main__":
    atom = arc_b200.Rubidium()
    nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    states, pairs = c2_workload(atom, nmax=nmax)
    print("states", len(states), "pairs", len(pairs), "dipole", int((pairs[:, 2] == 1).sum()),
          "quad", int((pairs[:, 2] == 2).sum()))
    recs = np.array([atom._state_record(*s) for s in states])
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        rb = radial.RadialBatch(recs)
        e1.record()
        out = rb.integrals_device(pairs)
        e2.record()
        torch.cuda.synchronize()
        npts = int(rb.lengths.sum())
        Lmin = np.minimum(rb.lengths[pairs[:, 0]], rb.lengths[pairs[:, 1]]).astype(np.int64).sum()
        t1, t2 = e0.elapsed_time(e1), e1.elapsed_time(e2)
        print("iter %d: numerov(+H2D) %.2f ms (%d pts, %.1f Gpts/s); integrals %.2f ms "
              "(%.3e pt-pairs, %.1f GB/s at 24 B/pt); elements/s %.3e"
              % (it, t1, npts, npts / t1 / 1e6, t2, Lmin, Lmin * 24 / t2 / 1e6,
                 len(pairs) / ((t1 + t2) * 1e-3)))
    print("sample", out[:3].cpu().numpy())
