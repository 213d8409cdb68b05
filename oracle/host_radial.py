"""Test helper: CPU stand-in for an atom's GPU radial batch, backed by the ORACLE.

Lets the `not gpu` tests exercise the host logic (basis enumeration, channel predicates,
angular algebra, matrix assembly) of StarkMap / PairStateInteractions on a box with no
GPU.  Test infrastructure only -- never imported by the product package.
"""

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

    def compute_elements(keys, powers, s=0.5):
        """Stand-in for AlkaliAtom._compute_elements: memo keys (n, l, 2j, n', l', 2j'), lower-energy
        state first, and their powers -> values from the oracle."""
        todo = list(zip(keys, powers))
        if procs <= 1 or len(todo) < 4 * procs:
            return [oracle.radial_integral(sp(k[0], k[1], k[2] / 2.0), sp(k[3], k[4], k[5] / 2.0), int(p))
                    for k, p in todo]
        import multiprocessing as mp
        params = dict(alphaC=atom.alphaC, alpha=atom.alpha, Z=atom.Z,
                      mu=(atom.mass - 9.1093837139e-31) / atom.mass,
                      a1=atom.a1, a2=atom.a2, a3=atom.a3, a4=atom.a4, rc=atom.rc)
        jobs = []
        for key, power in todo:
            sa = (key[0], key[1], key[2] / 2.0)
            sb = (key[3], key[4], key[5] / 2.0)
            jobs.append((params, sa + (atom.getEnergy(*sa) / 27.211386245981,),
                         sb + (atom.getEnergy(*sb) / 27.211386245981,), int(power)))
        oracle._lib()          # build liboracle.so once, before forking
        with mp.get_context("fork").Pool(procs) as pool:
            return pool.map(_element_worker, jobs, chunksize=max(1, len(jobs) // (8 * procs)))
    atom._compute_elements = compute_elements
    return atom


def _semiclassical_worker(job):
    kind, nu, nu1, l1, l2 = job
    fn = oracle.semiclassical_dipole if kind == 1 else oracle.semiclassical_quadrupole
    return fn(nu, nu1, l1, l2)


def patch_divalent_with_oracle(atom, procs=1):
    """Same for a DivalentAtom: its compute hook (semiclassical elements from the GPU Anger kernel)
    is replaced by the oracle's mpmath restatement of alkali_atom_functions.py:2683-2802."""
    import numpy as np

    def compute_elements(keys, powers, s=None):
        powers = np.asarray(powers)
        jobs = [None] * len(keys)
        for kind in (1, 2):
            sel = np.flatnonzero(powers == kind)
            if len(sel):
                params = atom._semiclassical_params([keys[i] for i in sel], kind, s)
                for i, p in zip(sel, params):
                    jobs[i] = (kind, float(p[0]), float(p[1]), float(p[2]), float(p[3]))
        if procs <= 1 or len(jobs) < 4 * procs:
            return [_semiclassical_worker(j) for j in jobs]
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            return pool.map(_semiclassical_worker, jobs, chunksize=max(1, len(jobs) // (8 * procs)))
    atom._compute_elements = compute_elements
    return atom
