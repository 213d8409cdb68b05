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
    pairs = np.concatenate(pairs).astype(np.int32)
    # the table sorted by (power, lower-energy state, partner): consecutive pairs share psi_a
    return states, pairs[np.lexsort((pairs[:, 1], pairs[:, 0], pairs[:, 2]))]


def c2_blocks(atom, nmin=20, nmax=120, lmax=20):
    """Same table as c2_workload, as Gram blocks ((l1,j1) x (l2,j2), power).  Returns
    (states, blocks, n_elements) with n_elements the number of DISTINCT allowed pairs."""
    states = [(n, l, j) for n in range(nmin, nmax + 1) for l in range(0, min(lmax, n - 1) + 1)
              for j in (l - .5, l + .5) if j > 0]
    by_lj = {}
    for i, (n, l, j) in enumerate(states):
        by_lj.setdefault((l, j), []).append(i)
    blocks, nel = [], 0
    keys = sorted(by_lj)
    for (l1, j1) in keys:
        for (l2, j2) in keys:
            dl = l2 - l1
            if dl == 1 and abs(j1 - j2) < 1.1:
                power = 1
            elif (dl == 2 or (dl == 0 and j2 >= j1)) and abs(j1 - j2) < 2.1:
                power = 2
            else:
                continue
            A, B = by_lj[(l1, j1)], by_lj[(l2, j2)]
            blocks.append((A, B, power))
            nel += len(A) * (len(A) + 1) // 2 if (l1, j1) == (l2, j2) else len(A) * len(B)
    return states, blocks, nel


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
    # tensor-core Gram path on the same table
    states2, blocks, nel = c2_blocks(atom, nmax=nmax)
    assert states2 == states
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g, offs = rb.gram_device(blocks)
        e1.record()
        torch.cuda.synchronize()
        print("gram iter %d: %.2f ms for %d blocks, %d distinct elements" % (it, e0.elapsed_time(e1), len(blocks), nel))
    # cross-check against the per-pair kernel on a sample
    gh = g.cpu().numpy(); oh = out.cpu().numpy()
    lookup = {}
    for b, (A, B, p) in enumerate(blocks):
        ia = {s_: q for q, s_ in enumerate(A)}; ib = {s_: q for q, s_ in enumerate(B)}
        lookup[b] = (ia, ib, len(B), p)
    rng = np.random.default_rng(1)
    worst = 0.0
    bl_of = {}
    for b, (A, B, p) in enumerate(blocks):
        bl_of[(states[A[0]][1], states[A[0]][2], states[B[0]][1], states[B[0]][2], p)] = b
    for q in rng.choice(len(pairs), 20000, replace=False):
        a, b_, p = pairs[q]
        ka = (states[a][1], states[a][2], states[b_][1], states[b_][2], int(p))
        kb = (states[b_][1], states[b_][2], states[a][1], states[a][2], int(p))
        if ka in bl_of:
            blk = bl_of[ka]; ia, ib, nc, _ = lookup[blk]; val = gh[offs[blk] + ia[a] * nc + ib[b_]]
        else:
            blk = bl_of[kb]; ia, ib, nc, _ = lookup[blk]; val = gh[offs[blk] + ia[b_] * nc + ib[a]]
        worst = max(worst, abs(val - oh[q]) / max(abs(oh[q]), 1e-300))
    print("gram vs per-pair kernel: max rel diff over 20000 sampled elements %.2e" % worst)
