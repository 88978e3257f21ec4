"""TEST INFRASTRUCTURE — builds `oracle/_ref/`: the reference's own implementation of the hot path, compiled from the
sources WHERE THEY LIE under /root/reference into sourceless byte code (pyc format, stored as `.refbc`: the gpurun snapshot
leaves `*.pyc` behind; oracle/ref_harness.py installs an import hook for the suffix), so that the unmodified reference travels to
the GPU box (which has no /root/reference) the way a compiled `.so` would.  No reference source is copied into the repo:
`oracle/_ref/` holds compiler output only, is git-ignored, and is rebuilt by `__graft_entry__.build()` whenever
/root/reference is present.

    python -m oracle.build_ref

The module list is the import closure of Training/Gamer.py, Training/ReplayBuffer.py, Search/Explorer.py, the two games and
Neural_Networks/Network_Manager.py (found with sys.modules after importing them through oracle/ref_harness.py).  The
reference tests `os.path.isfile("Games/SCS/Images/p<player>_<unit>.jpg")` when it creates a unit (SCS_Game.py:1803-1804) and
never opens the file: empty placeholders with those names are created for the scenarios this repo ships.
"""
import os
import py_compile
import sys

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE_ROOT = os.environ.get("NUZERO_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SUFFIX = ".refbc"

MODULES = [
    "Games/__init__.py", "Games/Game.py",
    "Games/Tic_Tac_Toe/__init__.py", "Games/Tic_Tac_Toe/tic_tac_toe.py",
    "Games/SCS/__init__.py", "Games/SCS/SCS_Game.py", "Games/SCS/SCS_Renderer.py", "Games/SCS/Terrain.py", "Games/SCS/Tile.py",
    "Games/SCS/Unit.py",
    "Neural_Networks/__init__.py", "Neural_Networks/Network_Manager.py",
    "Search/Explorer.py", "Search/Node.py",
    "Training/Gamer.py", "Training/ReplayBuffer.py",
    "Testing/__init__.py", "Testing/Agents/__init__.py", "Testing/Agents/Agent.py", "Testing/Agents/Generic/__init__.py",
    "Testing/Agents/Generic/RandomAgent.py", "Testing/Agents/Generic/PolicyAgent.py", "Testing/Agents/Generic/MctsAgent.py",
    "Utils/__init__.py", "Utils/Caches/Cache.py", "Utils/Caches/DictCache.py", "Utils/Caches/KeylessCache.py",
    "Utils/Functions/__init__.py", "Utils/Functions/general_utils.py", "Utils/Functions/loading_utlis.py",
    "Utils/Functions/ray_utils.py", "Utils/Functions/stats_utils.py", "Utils/Functions/yaml_utils.py",
    "Utils/Progress_Bars/PrintBar.py",
]


def build(verbose=False):
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "Search", "Explorer.py")):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import warnings

    import shutil

    shutil.rmtree(OUT, ignore_errors=True)
    n = 0
    for rel in MODULES:
        src = os.path.join(REFERENCE_ROOT, rel)
        if not os.path.isfile(src):
            if rel.endswith("__init__.py"):
                continue  # a namespace package in the reference
            raise RuntimeError("reference module missing: %s" % rel)
        dst = os.path.join(OUT, rel[:-3] + SUFFIX)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # SyntaxWarning: invalid escape sequence in the reference's ASCII art
            py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        n += 1
    # Scenario files for the byte-code reference: this repo's own scenario data (nuzero_b200/configs/scs, rendering keys
    # stripped) with the `image_path` key the reference's Terrain constructor insists on (SCS_Game.py:1674; never opened),
    # and empty placeholders for the unit counters its loader looks for.
    img = os.path.join(OUT, "Games", "SCS", "Images")
    cfg_out = os.path.join(OUT, "Games", "SCS", "Game_configs")
    os.makedirs(img, exist_ok=True)
    os.makedirs(cfg_out, exist_ok=True)
    cfg_dir = os.path.join(ROOT, "nuzero_b200", "configs", "scs")
    for f in sorted(os.listdir(cfg_dir)):
        data = yaml.safe_load(open(os.path.join(cfg_dir, f)))
        for name in (data.get("Units") or {}):
            unit = data["Units"][name]
            for player in (0, 1):
                open(os.path.join(img, "p%d_%s.jpg" % (player, unit.get("name", name))), "a").close()
        for props in (data.get("Terrain") or {}).values():
            props.setdefault("image_path", "")
        with open(os.path.join(cfg_out, f), "w") as fh:
            yaml.safe_dump(data, fh, sort_keys=False, default_flow_style=None)
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as fh:
        fh.write("byte code of %d reference modules compiled by oracle/build_ref.py with python %s\n" % (n, sys.version.split()[0]))
    if verbose:
        print("oracle/_ref: %d modules" % n)
    return OUT


if __name__ == "__main__":
    build(verbose=True)
