"""Test helper: CPU stand-in for an atom's GPU radial batch, backed by the ORACLE.

Lets the `not gpu` tests exercise the host logic (basis enumeration, channel predicates,
angular algebra, matrix assembly) of StarkMap / PairStateInteractions on a box with no
GPU.  Test infrastructure only -- never imported by the product package.
"""
from oracle import oracle


def patch_atom_with_oracle(atom):
    def sp(n, l, j):
        d = dict(n=n, l=l, j=j, E_au=atom.getEnergy(n, l, j) / 27.211386245981,
                 alphaC=atom.alphaC, alpha=atom.alpha, Z=atom.Z,
                 mu=(atom.mass - 9.1093837139e-31) / atom.mass)
        for k in ("a1", "a2", "a3", "a4", "rc"):
            d[k] = getattr(atom, k)[l] if l < 4 else 0.0
        return d

    def precompute_radial(dipole_pairs=(), quadrupole_pairs=(), s=0.5, useLiterature=True):
        n_new = 0
        for pairs, memo, power, ok in ((dipole_pairs, atom._dipole, 1, atom._dipole_allowed),
                                       (quadrupole_pairs, atom._quadrupole, 2, atom._quadrupole_allowed)):
            for (n1, l1, j1, n2, l2, j2) in pairs:
                if not ok(l1, j1, l2, j2):
                    continue
                key = atom._ordered_key(n1, l1, j1, n2, l2, j2)
                if key in memo or (power == 1 and useLiterature and key in atom._lit):
                    continue
                a, b = sp(key[0], key[1], key[2] / 2.0), sp(key[3], key[4], key[5] / 2.0)
                memo[key] = oracle.radial_integral(a, b, power)
                n_new += 1
        return n_new
    atom.precompute_radial = precompute_radial
    return atom
