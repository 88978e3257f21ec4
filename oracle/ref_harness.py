"""TEST INFRASTRUCTURE — imports the *unmodified* reference (read-only, /root/reference) so that
golden vectors can be generated from it in the build container.

Nothing in the product (`nuzero_b200/`) may import this module.  It only works where
`/root/reference` exists (the build container), never on the GPU box; the fixtures it produces are
committed under `tests/golden/` by `oracle/gen_golden.py`.

What is patched, and why:
  * five import stubs (`oracle/stubs/`) for UI / RPC packages that are not installed
    (termcolor, gymnasium, pettingzoo, pygame, ray) — none of them carries arithmetic;
  * shim I1: `tic_tac_toe.generate_network_input = tic_tac_toe.generate_state_image`
    (reference `Search/Explorer.py:145` calls a method `Games/Tic_Tac_Toe/tic_tac_toe.py:135`
    names differently);
  * parity protocol: `Search.Explorer.softmax` -> identity for the *network output* (the stub net
    emits dyadic probabilities directly) and `Search.Explorer.np.random` -> an RNG tape, so that
    both sides consume the same pre-drawn numbers.  `softmax_action` (visit-count softmax,
    `Explorer.py:187-199`) keeps the real scipy softmax.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")
_SUFFIX = ".refbc"  # pyc-format files under another name (oracle/build_ref.py)
_BYTECODE = os.path.join(_HERE, "_ref")  # oracle/build_ref.py: the reference compiled to sourceless byte code (travels to the GPU box)


def _find_root():
    env = os.environ.get("NUZERO_REFERENCE")
    for root in ([env] if env else []) + ["/root/reference", _BYTECODE]:
        if os.path.isfile(os.path.join(root, "Search", "Explorer.py")) or os.path.isfile(os.path.join(root, "Search", "Explorer" + _SUFFIX)):
            return root
    return env or "/root/reference"


REFERENCE_ROOT = _find_root()


def available():
    return (os.path.isfile(os.path.join(REFERENCE_ROOT, "Search", "Explorer.py")) or
            os.path.isfile(os.path.join(REFERENCE_ROOT, "Search", "Explorer" + _SUFFIX)))


def is_bytecode():
    """True when the reference in use is oracle/_ref (byte code of the unmodified sources), not the source tree."""
    return not os.path.isfile(os.path.join(REFERENCE_ROOT, "Search", "Explorer.py"))


_loaded = {}


def load():
    """Import the reference hot-path modules; returns a namespace of the classes."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
    if is_bytecode():
        # directories of the byte-code build are searched for `<module>.refbc` with python's own sourceless loader
        import importlib.machinery as mach

        root = os.path.realpath(REFERENCE_ROOT)

        def hook(path):
            if not os.path.realpath(path).startswith(root):
                raise ImportError
            return mach.FileFinder(path, (mach.SourcelessFileLoader, [_SUFFIX]))

        sys.path_hooks.insert(0, hook)
        sys.path_importer_cache.clear()
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _STUBS)
    # unit image paths in SCS_Game.create_unit are relative to the reference root
    os.chdir(REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from Search.Node import Node
        from Search import Explorer as explorer_mod
        from Games.Tic_Tac_Toe.tic_tac_toe import tic_tac_toe
        from Games.SCS.SCS_Game import SCS_Game

    tic_tac_toe.generate_network_input = tic_tac_toe.generate_state_image  # shim I1
    ns = types.SimpleNamespace(
        Node=Node,
        explorer_mod=explorer_mod,
        Explorer=explorer_mod.Explorer,
        tic_tac_toe=tic_tac_toe,
        SCS_Game=SCS_Game,
        scipy_softmax=explorer_mod.softmax,
    )
    _loaded["ns"] = ns
    return ns


def scs_config_path(name):
    # source tree: the reference's own files; byte-code build: the files oracle/build_ref.py derives from this repo's
    # scenario data (nuzero_b200/configs/scs)
    return os.path.join(REFERENCE_ROOT, "Games", "SCS", "Game_configs", name)


def make_scs(name, seed=None):
    ns = load()
    with contextlib.redirect_stdout(io.StringIO()):
        return ns.SCS_Game(scs_config_path(name), seed)


# --------------------------------------------------------------------------------------------
# parity protocol plumbing
# --------------------------------------------------------------------------------------------
class TapeRandom:
    """Stands in for `np.random` inside Search/Explorer.py.  Draw order per move follows
    `Explorer.run_mcts` (`Explorer.py:45-46` gamma) then `select_action` (`:74-89`)."""

    def __init__(self, gamma_tape, unif_tape):
        self.gamma_tape = gamma_tape  # [M, Kmax] f64
        self.unif_tape = unif_tape  # [M, 3]    f64: eps_softmax, eps_random, choice-uniform
        self.move = -1
        self._eps = 0

    def gamma(self, alpha, beta, n):
        self.move += 1
        self._eps = 0
        return np.array(self.gamma_tape[self.move, :n], dtype=np.float64)

    def random(self):
        v = float(self.unif_tape[self.move, self._eps])
        self._eps += 1
        return v

    def choice(self, a, p=None):
        # numpy legacy RandomState.choice(a, p=p): cdf = cumsum(p); cdf /= cdf[-1];
        # idx = searchsorted(cdf, random_sample(), side="right")
        p = np.asarray(p, dtype=np.float64)
        cdf = p.cumsum()
        cdf /= cdf[-1]
        u = float(self.unif_tape[self.move, 2])
        idx = int(cdf.searchsorted(u, side="right"))
        if isinstance(a, (int, np.integer)):
            return idx
        return a[idx]


class _NpProxy:
    def __init__(self, tape):
        self.random = tape

    def __getattr__(self, name):
        return getattr(np, name)


@contextlib.contextmanager
def parity_patches(tape=None, identity_softmax=True):
    """Temporarily rebinds names *inside the imported module object* (never the files)."""
    ns = load()
    mod = ns.explorer_mod
    saved_softmax, saved_np = mod.softmax, mod.np
    real_softmax = saved_softmax

    def softmax_passthrough(x, *a, **k):
        # network output (a torch tensor / ndarray of probabilities) passes through untouched;
        # the visit-count softmax of `softmax_action` receives a python list -> real softmax
        if isinstance(x, list):
            return real_softmax(x, *a, **k)
        return np.asarray(x)

    try:
        if identity_softmax:
            mod.softmax = softmax_passthrough
        if tape is not None:
            mod.np = _NpProxy(tape)
        yield
    finally:
        mod.softmax, mod.np = saved_softmax, saved_np
