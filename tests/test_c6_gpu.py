"""getC6perturbatively: the one pair-state function the reference's own
test-suite pins (test/test_pairstate.py:4-10, K 66S: -265 +- 1 GHz um^6)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def test_c6_perturbative_pins(gold_dir):
    import arc_b200
    g = json.load(open(os.path.join(gold_dir, "c6_golden.json")))
    for case in g:
        atom = getattr(arc_b200, case["atom"])()
        c = case["ctor"]
        calc = arc_b200.PairStateInteractions(atom, int(c[0]), int(c[1]), c[2], int(c[3]), int(c[4]), c[5], c[6], c[7])
        a = case["args"]
        v = calc.getC6perturbatively(a[0], a[1], int(a[2]), a[3])
        assert abs(v - case["value"]) <= 1e-7 * abs(case["value"]), (case, v)
    # the reference's own known-answer test
    k = arc_b200.Potassium()
    v = arc_b200.PairStateInteractions(k, 66, 0, .5, 66, 0, .5, .5, .5).getC6perturbatively(0, 0, 5, 30e9)
    assert abs(v - (-265)) < 1
