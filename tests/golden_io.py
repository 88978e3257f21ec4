"""Helpers shared by the oracle and GPU parity tests: load a golden fixture and compare a record
(`oracle.selfplay.play_game` schema) against it."""
import os

import numpy as np
import yaml

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# the SCS scenario files (re-emitted from the reference's Games/SCS/Game_configs by oracle/gen_golden_scs.py) ship with the package
SCS_CONFIGS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nuzero_b200", "configs", "scs")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["cfg"] = yaml.safe_load(str(g["cfg_yaml"]))
    g["training"] = bool(g["training"])
    g["salt"] = int(g["salt"])
    return g


def names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def assert_record_matches(rec, g, check_states=True, check_trees=True):
    """Bit-exact on every integer AND every float (W sums and priors are exact by construction of
    the dyadic stub; bias is a math.log result computed by the same libm on both sides)."""
    L = int(g["length"])
    assert rec["length"] == L
    assert rec["terminal_value"] == int(g["terminal_value"])
    assert list(rec["actions"]) == g["actions"].tolist()
    assert list(rec["root_N"]) == g["root_N"].tolist()
    assert list(rec["players"]) == g["players"].tolist()
    np.testing.assert_array_equal(np.asarray(rec["root_W"], dtype=np.float64), g["root_W"])
    np.testing.assert_array_equal(np.asarray(rec["bias"], dtype=np.float64), g["bias"])
    off = g["child_off"]
    for m in range(L):
        s = slice(off[m], off[m + 1])
        np.testing.assert_array_equal(rec["child_actions"][m], g["child_actions"][s], err_msg="move %d" % m)
        np.testing.assert_array_equal(rec["child_N"][m], g["child_N"][s], err_msg="move %d" % m)
        np.testing.assert_array_equal(rec["child_W"][m], g["child_W"][s], err_msg="move %d" % m)
        np.testing.assert_array_equal(rec["child_prior"][m], g["child_prior"][s], err_msg="move %d" % m)
        if check_states:
            np.testing.assert_array_equal(rec["states"][m], g["states"][m], err_msg="state, move %d" % m)
            np.testing.assert_array_equal(rec["masks"][m], g["masks"][m], err_msg="mask, move %d" % m)
    if check_trees:
        for m in g["tree_moves"].tolist():
            ti, tf = rec["trees"][m]
            np.testing.assert_array_equal(ti, g["tree%d_i" % m], err_msg="tree ints, move %d" % m)
            np.testing.assert_array_equal(tf, g["tree%d_f" % m], err_msg="tree floats, move %d" % m)
