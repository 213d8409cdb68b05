"""Test helper: CPU stand-in for an atom's GPU radial batch, backed by the ORACLE.

Lets the `not gpu` tests exercise the host logic (basis enumeration, channel predicates,
angular algebra, matrix assembly) of StarkMap / PairStateInteractions on a box with no
GPU.  Test infrastructure only -- never imported by the product package.
"""
import os

from oracle import oracle


def _sp_from(params, n, l, j, E_au):
    d = dict(n=n, l=l, j=j, E_au=E_au, alphaC=params["alphaC"], alpha=params["alpha"],
             Z=params["Z"], mu=params["mu"])
    for k in ("a1", "a2", "a3", "a4", "rc"):
        d[k] = params[k][l] if l < 4 else 0.0
    return d


def _element_worker(job):
    params, a, b, power = job
    return oracle.radial_integral(_sp_from(params, *a), _sp_from(params, *b), power)


def patch_atom_with_oracle(atom, procs=1):
    def sp(n, l, j):
        d = dict(n=n, l=l, j=j, E_au=atom.getEnergy(n, l, j) / 27.211386245981,
                 alphaC=atom.alphaC, alpha=atom.alpha, Z=atom.Z,
                 mu=(atom.mass - 9.1093837139e-31) / atom.mass)
        for k in ("a1", "a2", "a3", "a4", "rc"):
            d[k] = getattr(atom, k)[l] if l < 4 else 0.0
        return d

    def precompute_radial(dipole_pairs=(), quadrupole_pairs=(), s=0.5, useLiterature=True):
        todo = []
        for pairs, memo, power, ok in ((dipole_pairs, atom._dipole, 1, atom._dipole_allowed),
                                       (quadrupole_pairs, atom._quadrupole, 2, atom._quadrupole_allowed)):
            for (n1, l1, j1, n2, l2, j2) in pairs:
                if not ok(l1, j1, l2, j2):
                    continue
                key = atom._ordered_key(n1, l1, j1, n2, l2, j2)
                if key in memo or (power == 1 and useLiterature and key in atom._lit):
                    continue
                memo[key] = None
                todo.append((memo, key, power))
        if procs <= 1 or len(todo) < 4 * procs:
            for memo, key, power in todo:
                a, b = sp(key[0], key[1], key[2] / 2.0), sp(key[3], key[4], key[5] / 2.0)
                memo[key] = oracle.radial_integral(a, b, power)
        else:
            import multiprocessing as mp
            params = dict(alphaC=atom.alphaC, alpha=atom.alpha, Z=atom.Z,
                          mu=(atom.mass - 9.1093837139e-31) / atom.mass,
                          a1=atom.a1, a2=atom.a2, a3=atom.a3, a4=atom.a4, rc=atom.rc)
            jobs = []
            for memo, key, power in todo:
                sa = (key[0], key[1], key[2] / 2.0)
                sb = (key[3], key[4], key[5] / 2.0)
                jobs.append((params, sa + (atom.getEnergy(*sa) / 27.211386245981,),
                             sb + (atom.getEnergy(*sb) / 27.211386245981,), power))
            oracle._lib()          # build liboracle.so once, before forking
            with mp.get_context("fork").Pool(procs) as pool:
                vals = pool.map(_element_worker, jobs, chunksize=max(1, len(jobs) // (8 * procs)))
            for (memo, key, power), v in zip(todo, vals):
                memo[key] = v
        return len(todo)
    atom.precompute_radial = precompute_radial
    return atom
