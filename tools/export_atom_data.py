"""Export the element INPUT DATA the hot path needs into a JSON the package ships.

The reference's element classes (alkali_atom_data.py / divalent_atom_data.py) are
out of scope to rewrite ("read as inputs"): this script reads
the published physical parameters off live reference atom objects -- model
potential a1..a4, rc, alphaC (Marinescu et al. PRA 49, 982), Rydberg-Ritz quantum
defects, masses, NIST level energies below minQuantumDefectN, literature dipole
matrix elements -- and writes them as plain data to
    arc-alkali-rydberg-calculator_b200/data/atoms.json
so that the package runs on a box where /root/reference does not exist.

Run here (container with /root/reference):   python tools/export_atom_data.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.refenv as refenv  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "arc-alkali-rydberg-calculator_b200", "data", "atoms.json")

ALKALI = ["Hydrogen", "Lithium6", "Lithium7", "Sodium", "Potassium39",
          "Potassium40", "Potassium41", "Rubidium85", "Rubidium87", "Caesium"]
DIVALENT = ["Strontium88", "Calcium40", "Ytterbium174"]
ALIASES = {"Rubidium": "Rubidium85", "Potassium": "Potassium39", "Cesium": "Caesium"}


def _f(x):
    return float(x)


def export_atom(arc, name, divalent):
    a = getattr(arc, name)()
    d = {
        "className": name,
        "elementName": a.elementName,
        "divalent": divalent,
        "Z": int(a.Z),
        "I": _f(a.I),
        "mass": _f(a.mass),
        "alpha": _f(a.alpha),
        "alphaC": _f(a.alphaC),
        "alpha_d_eff": _f(a.alpha_d_eff),
        "alpha_q_eff": _f(a.alpha_q_eff),
        "a1": [_f(v) for v in a.a1], "a2": [_f(v) for v in a.a2],
        "a3": [_f(v) for v in a.a3], "a4": [_f(v) for v in a.a4],
        "rc": [_f(v) for v in a.rc],
        "gS": _f(a.gS), "gL": _f(a.gL),
        "groundStateN": int(a.groundStateN),
        "extraLevels": [[_f(v) for v in lv] for lv in a.extraLevels],
        "minQuantumDefectN": int(a.minQuantumDefectN),
        "NISTdataLevels": int(a.NISTdataLevels),
        "scaledRydbergConstant": _f(a.scaledRydbergConstant),
        "ionisationEnergy": _f(a.ionisationEnergy),
        "quantumDefect": [[([_f(v) for v in row] + [0.0] * 6)[:6] for row in blk]
                          for blk in a.quantumDefect],
        "preferQuantumDefects": bool(a.preferQuantumDefects),
    }
    c = a.conn.cursor()
    if divalent:
        d["defectFittingRange"] = {k: [int(v[0]), int(v[1])]
                                   for k, v in a.defectFittingRange.items()}
        c.execute("SELECT n,l,j,s,energy FROM energyLevel")
        d["savedEnergy"] = [[int(r[0]), int(r[1]), _f(r[2]), _f(r[3]), _f(r[4])]
                            for r in c.fetchall()]
        c.execute("SELECT n1,l1,j1,n2,l2,j2,s,dme,errorEstimate FROM literatureDME")
        d["literatureDME"] = [[int(r[0]), int(r[1]), _f(r[2]), int(r[3]), int(r[4]),
                               _f(r[5]), _f(r[6]), _f(r[7]), _f(r[8])]
                              for r in c.fetchall()]
    else:
        se = np.asarray(a.sEnergy)
        rows = []
        if se.ndim == 2:
            nz = np.argwhere(np.abs(se) > 0)
            for i, j in nz:
                rows.append([int(i), int(j), _f(se[i, j])])
        d["sEnergy"] = rows          # sEnergy[n,l] = j=l-1/2 ; sEnergy[l,n] = j=l+1/2
        c.execute("SELECT n1,l1,j1_x2,n2,l2,j2_x2,dme,errorEstimate FROM literatureDME")
        d["literatureDME"] = [[int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]),
                               int(r[5]), _f(r[6]), _f(r[7])] for r in c.fetchall()]
    return d


def main():
    arc = refenv.load()
    out = {"_source": "exported from ARC %s element classes by tools/export_atom_data.py"
                      % getattr(arc, "__version__", "3.10.2"),
           "aliases": ALIASES, "atoms": {}}
    for n in ALKALI:
        out["atoms"][n] = export_atom(arc, n, False)
    for n in DIVALENT:
        try:
            out["atoms"][n] = export_atom(arc, n, True)
        except Exception as e:  # keep going; report
            print("skip", n, e)
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", list(out["atoms"]))


if __name__ == "__main__":
    main()
