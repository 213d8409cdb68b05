import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import arc_b200
from arc_b200 import radial
from oracle import oracle
from dev_radial_bench import c2_workload, c2_blocks
atom = arc_b200.Rubidium()
states, pairs = c2_workload(atom)
_, blocks, nel = c2_blocks(atom)
recs = np.array([atom._state_record(*s) for s in states])
rb = radial.RadialBatch(recs)
out = rb.integrals_device(pairs).cpu().numpy()
import time
for th in (1e-5, 1e-4, 3e-4, 1e-3):
    rb.gram_refined_device(blocks, thresh=th); torch.cuda.synchronize(); t0 = time.time()
    g, offs, nref = rb.gram_refined_device(blocks, thresh=th); torch.cuda.synchronize()
    dt = (time.time() - t0) * 1e3
    gh = g.cpu().numpy()
    bl_of, lookup = {}, {}
    for b, (A, B, p) in enumerate(blocks):
        bl_of[(states[A[0]][1], states[A[0]][2], states[B[0]][1], states[B[0]][2], p)] = b
        lookup[b] = ({s_: q for q, s_ in enumerate(A)}, {s_: q for q, s_ in enumerate(B)}, len(B))
    gv = np.zeros(len(pairs))
    for q in range(len(pairs)):
        a, b_, p = pairs[q]
        ka = (states[a][1], states[a][2], states[b_][1], states[b_][2], int(p))
        if ka in bl_of:
            blk = bl_of[ka]; ia, ib, nc = lookup[blk]; gv[q] = gh[offs[blk] + ia[a] * nc + ib[b_]]
        else:
            blk = bl_of[(ka[2], ka[3], ka[0], ka[1], ka[4])]; ia, ib, nc = lookup[blk]; gv[q] = gh[offs[blk] + ia[b_] * nc + ib[a]]
    rel = np.abs(gv - out) / np.maximum(np.abs(out), 1e-300)
    print("thresh %.0e: refined %d (%.1f%%) in %.1f ms total; rel>1e-8: %d  rel>3e-9: %d  max %.1e" % (th, nref, 100.0 * nref / len(pairs), dt, int((rel > 1e-8).sum()), int((rel > 3e-9).sum()), rel.max()), flush=True)
bl_of, lookup = {}, {}
for b, (A, B, p) in enumerate(blocks):
    bl_of[(states[A[0]][1], states[A[0]][2], states[B[0]][1], states[B[0]][2], p)] = b
    lookup[b] = ({s_: q for q, s_ in enumerate(A)}, {s_: q for q, s_ in enumerate(B)}, len(B))
gv = np.zeros(len(pairs))
for q in range(len(pairs)):
    a, b_, p = pairs[q]
    ka = (states[a][1], states[a][2], states[b_][1], states[b_][2], int(p))
    if ka in bl_of:
        blk = bl_of[ka]; ia, ib, nc = lookup[blk]; gv[q] = gh[offs[blk] + ia[a] * nc + ib[b_]]
    else:
        blk = bl_of[(ka[2], ka[3], ka[0], ka[1], ka[4])]; ia, ib, nc = lookup[blk]; gv[q] = gh[offs[blk] + ia[b_] * nc + ib[a]]
rel = np.abs(gv - out) / np.maximum(np.abs(out), 1e-300)
print("elements", len(pairs), "rel>1e-8:", int((rel > 1e-8).sum()), "rel>1e-9:", int((rel > 1e-9).sum()), "median %.1e" % np.median(rel))
def sp(i):
    n, l, j = states[i]
    d = dict(n=n, l=l, j=j, E_au=atom.getEnergy(n, l, j) / 27.211386245981, alphaC=atom.alphaC, alpha=atom.alpha, Z=atom.Z, mu=(atom.mass - 9.1093837139e-31) / atom.mass)
    for kk in ("a1", "a2", "a3", "a4", "rc"):
        d[kk] = getattr(atom, kk)[l] if l < 4 else 0.0
    return d
for q in np.argsort(-rel)[:6]:
    a, b_, p = pairs[q]
    ref = oracle.radial_integral(sp(a), sp(b_), int(p))
    r1, p1 = oracle.radial_wavefunction(sp(a)); r2, p2 = oracle.radial_wavefunction(sp(b_))
    up = min(len(r1), len(r2)); w = r1[:up] if p == 1 else r1[:up] ** 2
    S = oracle.trapezoid(np.abs(p1[:up] * p2[:up] * w), r1[:up])
    print(states[a], states[b_], int(p), "oracle %.12e perpair %.12e gram %.12e |S| %.3e  relerr perpair %.1e gram %.1e  abs/S gram %.1e" % (
        ref, out[q], gv[q], S, abs(out[q] - ref) / abs(ref), abs(gv[q] - ref) / abs(ref), abs(gv[q] - ref) / S))
