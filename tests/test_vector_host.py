"""CPU tests of the vector host path (no GPU): `defineBasis` assembles its matrices from index
arrays and whole-table radial look-ups (atoms.radial_elements_indexed); here every entry is
rebuilt through the reference-shaped per-element call chain
(`_eFieldCouplingDivE` -> `getRadialMatrixElement`, `getRadialCoupling`,
calculations_atom_single.py:650-666, alkali_atom_functions.py:2260-2298, 4123-4129) and has to
match bit for bit.  The GPU compute hook `_compute_elements` is replaced by a deterministic
function of the memo key, so both routes see the same radial values."""
import numpy as np
import pytest
from scipy.constants import h as C_h, e as C_e, epsilon_0, physical_constants

import arc_b200


def _fake_value(key, power):
    return float(np.sin(sum((i + 1) * 1.37 * float(k) for i, k in enumerate(key[:6])) + power) * 100.0 + 3.0)


def _patch(atom):
    calls = []

    def compute(keys, powers, s=0.5):
        calls.append(len(keys))
        return [_fake_value(k, p) for k, p in zip(keys, powers)]
    atom._compute_elements = compute
    atom._compute_calls = calls
    return atom


STARK_CASES = [
    ("Rubidium", (28, 0, 0.5, 0.5, 26, 30, 8), {}),
    ("Rubidium", (6, 0, 0.5, 0.5, 4, 9, 5), {}),                       # reaches below groundStateN
    ("Caesium", (40, 2, 2.5, 1.5, 37, 43, 9), {"Bz": 1e-3}),
    ("Strontium88", (55, 0, 0, 0, 53, 57, 10), {"s": 0}),
    ("Strontium88", (55, 0, 1, 0, 53, 57, 10), {"s": 1}),
    ("Potassium", (5, 1, 1.5, 0.5, 3, 7, 3), {}),                       # literature dipole elements
]


@pytest.mark.parametrize("atomname,args,kw", STARK_CASES)
def test_stark_matrices_equal_per_element_path(atomname, args, kw):
    atom = _patch(getattr(arc_b200, atomname)())
    calc = arc_b200.StarkMap(atom)
    calc.defineBasis(*args, **kw)
    assert len(atom._compute_calls) <= 1                 # one batch for the whole table
    s = kw.get("s", 0.5)
    ref = _patch(getattr(arc_b200, atomname)())
    chk = arc_b200.StarkMap(ref)
    st = calc.basisStates
    N = len(st)
    mat2 = np.zeros((N, N))
    for i in range(N):
        for j in range(i + 1, N):
            c = chk._eFieldCouplingDivE(st[i][0], st[i][1], st[i][2], st[i][3],
                                        st[j][0], st[j][1], st[j][2], st[j][3], s=s) * 1.0e-9 / C_h
            mat2[i, j] = mat2[j, i] = c
    assert np.count_nonzero(mat2) > 0
    assert np.array_equal(calc.mat2, mat2)
    assert atom._dipole == ref._dipole
    for i in range(N):
        e = (ref.getEnergy(st[i][0], st[i][1], st[i][2], s=s) * C_e / C_h * 1e-9
             + ref.getZeemanEnergyShift(st[i][1], st[i][2], st[i][3], kw.get("Bz", 0), s=s) / C_h * 1.0e-9)
        assert calc.mat1[i, i] == e


def test_shirley_couplings_equal_per_element_path():
    atom = _patch(arc_b200.Rubidium85())
    calc = arc_b200.ShirleyMethod(atom)
    calc.defineBasis(40, 2, 2.5, 0.5, 0, 37, 43, 8)
    ref = _patch(arc_b200.Rubidium85())
    chk = arc_b200.ShirleyMethod(ref)
    st = calc.basisStates
    N = len(st)
    V = np.zeros((N, N))
    for i in range(N):
        for j in range(i + 1, N):
            # calculations_atom_single.py:3641-3660: both angular factors at the target's mj
            c = 0.5 * chk._eFieldCouplingDivE(st[i][0], st[i][1], st[i][2], calc.mj,
                                              st[j][0], st[j][1], st[j][2], calc.mj, s=0.5) / C_h
            V[i, j] = V[j, i] = c
    assert np.count_nonzero(V) > 0
    assert np.array_equal(calc.V, V)


PAIR_CASES = [
    (lambda: (arc_b200.Rubidium(), None), (45, 0, .5, 45, 0, .5, .5, .5), {}, (0., 0., 2, 3, 15e9), {}),
    (lambda: (arc_b200.Caesium(), None), (35, 2, 1.5, 36, 0, 0.5, 0.5, 0.5), {"interactionsUpTo": 2},
     (0.7, 0., 1, 3, 6e9), {"Bz": 2e-4}),
    (lambda: (arc_b200.Rubidium(), arc_b200.Caesium()), (45, 0, .5, 46, 1, 1.5, .5, .5), {}, (0., 0., 2, 3, 15e9), {}),
    (lambda: (arc_b200.Strontium88(), None), (45, 0, 1, 45, 0, 1, 1, 1), {"s": 1, "interactionsUpTo": 2},
     (0.3, 0., 1, 2, 10e9), {}),
]


@pytest.mark.parametrize("atoms,state,ckw,args,kw", PAIR_CASES)
def test_pair_channel_couplings_equal_per_element_path(atoms, state, ckw, args, kw):
    a1, a2 = atoms()
    _patch(a1)
    if a2 is not None:
        _patch(a2)
        ckw = dict(ckw, atom2=a2)
    calc = arc_b200.PairStateInteractions(a1, *state, **ckw)
    calc.defineBasis(*args, **kw)
    r1, r2 = atoms()
    _patch(r1)
    r2 = r1 if r2 is None else _patch(r2)
    a0 = physical_constants["Bohr radius"][0]
    nch, found = len(calc.channel), 0
    for o, cm in enumerate(calc.coupling):
        cm = cm.tocoo()
        assert len(cm.data) == 0 or (cm.row < cm.col).all()
        for i, j, v in zip(cm.row, cm.col, cm.data):
            ci, cj = calc.channel[i], calc.channel[j]
            x1 = r1.getRadialCoupling(ci[0], ci[1], ci[2], cj[0], cj[1], cj[2], s=calc.s1)
            x2 = r2.getRadialCoupling(ci[3], ci[4], ci[5], cj[3], cj[4], cj[5], s=calc.s2)
            want = (C_e ** 2 / (4.0 * np.pi * epsilon_0) * x1 * x2 * a0 ** (o + 2)) / C_h * 1.0e-9
            assert v == want
            found += 1
    assert found > 0 and nch > 1
    assert a1._dipole == r1._dipole and a1._quadrupole == r1._quadrupole


def test_radial_elements_vector_api():
    atom = _patch(arc_b200.Rubidium())
    rows = [(30, 0, .5, 30, 1, 1.5), (30, 1, 1.5, 30, 0, .5), (30, 0, .5, 31, 2, 1.5), (31, 1, .5, 30, 0, .5),
            (30, 0, .5, 30, 1, 1.5), (5, 0, .5, 5, 1, .5)]
    v = atom.radial_elements(rows, 1)
    assert atom._compute_calls == [2]                                  # two distinct new elements, one batch
    assert v[0] == v[1] == v[4] == atom.getRadialMatrixElement(30, 0, .5, 30, 1, 1.5)
    assert v[2] == 0                                                   # dl = 2: forbidden
    assert v[3] == atom.getRadialMatrixElement(30, 0, .5, 31, 1, .5)
    assert v[5] == atom._lit[(5, 0, 1, 5, 1, 1)][0]                    # literature value wins
    assert atom.radial_elements(rows, 1, useLiterature=False)[5] == atom._dipole[(5, 0, 1, 5, 1, 1)]
    q = atom.radial_elements(rows, 2)
    assert q[2] == atom.getQuadrupoleMatrixElement(30, 0, .5, 31, 2, 1.5) != 0
    assert atom.radial_elements([], 1).shape == (0,)
    assert atom.precompute_radial(dipole_pairs=rows) == 0              # everything memoised already
    assert atom.precompute_radial(dipole_pairs=np.array([(40, 0, .5, 40, 1, .5)]),
                                  quadrupole_pairs=[(40, 0, .5, 41, 2, 1.5)]) == 2
    sr = _patch(arc_b200.Strontium88())
    with pytest.raises(ValueError):
        sr.radial_elements([(50, 0, 0, 50, 1, 1)], 1)
    d = sr.radial_elements([(50, 0, 0, 50, 1, 1), (50, 1, 1, 50, 0, 0), (50, 0, 0, 50, 2, 2)], 1, s=0)
    assert d[0] == d[1] == sr.getRadialMatrixElement(50, 0, 0, 50, 1, 1, s=0) and d[2] == 0


@pytest.mark.parametrize("atomname,state,upto,theta,args,kw", [
    ("Rubidium", (45, 0, .5, 45, 0, .5, .5, .5), 1, 0.0, (2, 3, 15e9), {}),
    ("Caesium", (35, 2, 1.5, 36, 0, 0.5, 0.5, 0.5), 2, 0.7, (2, 3, 6e9), {"Bz": 2e-4})])
def test_matR_blocks_equal_the_closed_form_per_channel_pair(atomname, state, upto, theta, args, kw):
    """matR is expanded class pair by class pair; here it is rebuilt channel pair by channel pair
    from SURVEY a16': matR[o][rows of j, cols of i] = rho * Re[(D(j1') x D(j2')) d (D(j1) x D(j2))^+],
    entries with |angular| <= 1e-5 dropped, Hermitian mirror added (pair:3116-3217)."""
    from arc_b200 import WignerDmatrix
    atom = _patch(getattr(arc_b200, atomname)())
    calc = arc_b200.PairStateInteractions(atom, *state, interactionsUpTo=upto)
    calc.defineBasis(theta, 0., *args, **kw)
    N, index = len(calc.basisStates), calc.index.astype(np.int64)
    wgd = WignerDmatrix(theta, 0.)

    def prod_idx(ch_i):
        out = []
        for b in calc.basisStates[index[ch_i]:index[ch_i + 1]]:
            out.append(round(b[3] + b[2]) * round(2 * b[6] + 1) + round(b[7] + b[6]))
        return np.array(out, dtype=np.int64)
    seen = 0
    for o, cm in enumerate(calc.coupling):
        want = np.zeros((N, N))
        cm = cm.tocoo()
        for i, j, rho in zip(cm.row, cm.col, cm.data):
            chi, chj = calc.channel[i], calc.channel[j]
            if index[i] == index[i + 1] or index[j] == index[j + 1]:
                continue
            d = calc._getAngularMatrix_M(chi[1], chi[2], chi[4], chi[5], chj[1], chj[2], chj[4], chj[5])
            d = np.asarray(d.todense() if hasattr(d, "todense") else d)
            if theta != 0.0:
                d = (np.kron(wgd.get(chj[2]), wgd.get(chj[5])) @ d
                     @ np.kron(wgd.get(chi[2]), wgd.get(chi[5])).conj().T).real
            blk = d[np.ix_(prod_idx(j), prod_idx(i))]
            blk = np.where(np.abs(blk) > 1.0e-5, blk, 0.0) * rho
            want[index[j]:index[j + 1], index[i]:index[i + 1]] += blk
            want[index[i]:index[i + 1], index[j]:index[j + 1]] += blk.T
            seen += 1
        got = calc.matR[o].toarray()
        scale = max(np.abs(want).max(), 1e-300)
        assert np.abs(got - want).max() <= 1e-13 * scale
        # what is staged on the GPU (COO scatter) is the same matrix
        r, c, v = calc._matR_coo[o]
        dense = np.zeros((N, N))
        dense[r, c] = v
        assert np.array_equal(dense, got) and len(set(zip(r.tolist(), c.tolist()))) == len(r)
    assert seen >= 20


def test_unique_bounded_equals_numpy_unique():
    from arc_b200.atoms import _unique_bounded
    rng = np.random.default_rng(3)
    for hi, m in ((5, 40), (1000, 300), (1 << 23, 50), (3, 1), (10 ** 12, 200)):
        code = rng.integers(0, hi, size=m)
        uc, rep, inv = _unique_bounded(code)
        u2, _f2, i2 = np.unique(code, return_index=True, return_inverse=True)
        assert np.array_equal(uc, u2) and np.array_equal(inv, i2.ravel())
        assert np.array_equal(code[rep], uc)                      # a valid representative of every value
    for empty in _unique_bounded(np.zeros(0, dtype=np.int64)):
        assert len(empty) == 0


def test_index_states_round_trip():
    rng = np.random.default_rng(5)
    n = rng.integers(4, 80, size=(200, 2))
    l = np.minimum(rng.integers(0, 30, size=(200, 2)), n - 1)
    j = l + rng.choice([-0.5, 0.5], size=(200, 2))
    j = np.where(j < 0, 0.5, j)
    P = np.stack([n[:, 0], l[:, 0], j[:, 0], n[:, 1], l[:, 1], j[:, 1]], axis=1).astype(np.float64)
    S, ia, ib = arc_b200.AlkaliAtom._index_states(P)
    assert np.array_equal(S[ia], P[:, 0:3]) and np.array_equal(S[ib], P[:, 3:6])
    assert len(np.unique(S, axis=0)) == len(S)                    # every distinct state once
    S0, a0, b0 = arc_b200.AlkaliAtom._index_states(np.zeros((0, 6)))
    assert S0.shape == (0, 3) and len(a0) == len(b0) == 0


def test_forbidden_rows_never_touch_the_energies():
    """getRadialMatrixElement returns 0 for a forbidden pair before it looks at the states
    (aaf:1016-1019); the vector form must not trip over a state it would never have evaluated."""
    atom = _patch(arc_b200.Rubidium())
    rows = [(30, 0, .5, 30, 1, 1.5), (30, 0, .5, 5, 7, 7.5)]          # second row: dl = 7, and l >= n
    v = atom.radial_elements(rows, 1)
    assert v[0] != 0 and v[1] == 0
    with pytest.raises(ValueError):
        atom.getEnergy(5, 7, 7.5)


def test_greedy_assignment_fast_path_equals_the_scan():
    """PairStateInteractions._greedy_assignment: the one-pass assignment of fully dominant points gives
    what the reference's descending scan gives (calculations_atom_pairstate.py:3460-3480)."""
    from arc_b200.pairstate import PairStateInteractions as P

    def scan(S):
        k = S.shape[0]
        rows_, cols_ = np.unravel_index(np.argsort(S.ravel()), (k, k))
        rowPicked, colPicked = np.zeros(k, dtype=bool), np.zeros(k, dtype=bool)
        perm = np.zeros(k, dtype=np.int64)
        j, done = k * k - 1, 0
        while done < k:
            while rowPicked[rows_[j]] or colPicked[cols_[j]]:
                j -= 1
            rowPicked[rows_[j]] = colPicked[cols_[j]] = True
            perm[cols_[j]] = rows_[j]
            done += 1
        return perm
    rng = np.random.default_rng(11)
    fast = 0
    for k, nrot, thmax in ((40, 0, 0.0), (40, 30, 0.6), (40, 30, np.pi / 2), (7, 3, 0.5), (1, 0, 0.0), (64, 200, 0.3)):
        for _ in range(5):
            Q = np.eye(k)[rng.permutation(k)]
            for _r in range(nrot):
                a, b = rng.choice(k, 2, replace=False)
                th = rng.uniform(0, thmax)
                G = np.eye(k)
                G[a, a] = G[b, b] = np.cos(th)
                G[a, b], G[b, a] = -np.sin(th), np.sin(th)
                Q = G @ Q
            S = Q ** 2
            fast += int((S > 0.5).sum() == k)
            assert np.array_equal(P._greedy_assignment(S), scan(S))
    assert fast >= 10                                  # both branches were exercised
    S = rng.random((9, 9))                             # not overlaps of orthonormal sets: scan branch
    assert np.array_equal(P._greedy_assignment(S), scan(S))
