"""The CPU arm of bench.py runs the reference from `oracle/_ref` — byte code that oracle/build_ref.py compiles from the
reference's own modules — wherever the source tree is absent (the GPU box).  This test replays two committed Gamer fixtures
(generated from the SOURCE tree by oracle/gen_golden_gamer.py) through the byte-code build in a fresh process: the real
`Training/Gamer.play_game` + `Training/ReplayBuffer.save_game` must return the same statistics and tuples bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_io

ROOT = os.path.dirname(golden_io.GOLDEN.rstrip("/")).rsplit("/tests", 1)[0]
REF = os.path.join(ROOT, "oracle", "_ref")

SCRIPT = r"""
import json, os, sys
import numpy as np, yaml
sys.path.insert(0, sys.argv[1])
from oracle import ref_harness as rh
from oracle.gen_golden_gamer import play
assert rh.is_bytecode(), rh.REFERENCE_ROOT
ns = rh.load()
z = np.load(sys.argv[2], allow_pickle=False)
cfg = yaml.safe_load(str(z["cfg_yaml"]))
desc = str(z["game"])
if desc == "ttt":
    cls, args = ns.tic_tac_toe, []
else:
    _, name, seed = desc.split(":")
    cls, args = ns.SCS_Game, [rh.scs_config_path(name)] + ([] if seed == "None" else [int(seed)])
tm = 200
g = np.zeros((tm, 128)); u = np.zeros((tm, 3))
g[: z["gamma_tape"].shape[0], : z["gamma_tape"].shape[1]] = z["gamma_tape"]
u[: z["unif_tape"].shape[0]] = z["unif_tape"]
stats, entries, moves = play(cls, args, cfg, int(z["salt"]), (g, u), int(z["game_index"]))
out = {"stats": [float(stats[k]) for k in [str(k) for k in z["stats_keys"]]],
       "states": np.concatenate([np.asarray(e[0], dtype=np.float32) for e in entries]).tobytes().hex(),
       "policy": np.array([e[1][1] for e in entries], dtype=np.float64).tobytes().hex(),
       "value": [float(e[1][0]) for e in entries], "gidx": [int(e[2]) for e in entries]}
print("RESULT" + json.dumps(out))
"""


@pytest.mark.parametrize("name", ["gamer_ttt_s60", "gamer_scs_mirrored5_s30"])
def test_bytecode_reference_reproduces_the_source_reference_fixtures(name):
    if not os.path.isfile(os.path.join(REF, "Search", "Explorer.refbc")):
        pytest.skip("oracle/_ref not built here (python -m oracle.build_ref needs the reference tree)")
    fixture = os.path.join(golden_io.GOLDEN, name + ".npz")
    env = dict(os.environ, NUZERO_REFERENCE=REF)
    res = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, fixture], env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT")][-1]
    got = json.loads(line[len("RESULT"):])
    z = np.load(fixture, allow_pickle=False)
    assert got["stats"] == [float(x) for x in z["stats"]]
    assert got["states"] == z["states"].astype(np.float32).tobytes().hex()
    assert got["policy"] == z["policy"].astype(np.float64).tobytes().hex()
    assert got["value"] == [float(x) for x in z["value"]] and got["gidx"] == [int(x) for x in z["gidx"]]
