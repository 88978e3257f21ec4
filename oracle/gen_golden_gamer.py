"""TEST INFRASTRUCTURE — fixtures from the UNMODIFIED reference `Training/Gamer.py:39-97` (imported read-only from
/root/reference through oracle/ref_harness.py; `ray` is the pass-through stub in oracle/stubs, buffer and shared storage
are tiny in-process objects with a `.remote` attribute, as SURVEY.md Appendix B describes).

    python -m oracle.gen_golden_gamer

Each fixture `tests/golden/gamer_*.npz` holds what one `Gamer.play_game()` call returns and ships: the six statistics
(Gamer.py:42-50,81-92) and the tuples the reference `Training/ReplayBuffer.save_game` (ReplayBuffer.py:24-36, the real class)
appended for the game: state [1, C, R, Cc] f32, (value target, policy target over all actions), game index.
Protocol: the parity protocol of gen_golden.py (dyadic stub network, identity softmax for its output, RNG tape) —
`Gamer` always searches with training=True (Gamer.py:33).
"""
import contextlib
import importlib
import io
import os

import numpy as np
import yaml

from . import ref_harness as rh
from .gen_golden import GOLDEN, make_tape, search_config
from .stubnet_np import StubNetwork


class _Remote:
    """An actor handle of the ray stub: `handle.method.remote(*args)` is a local call."""

    def __init__(self, obj):
        self._obj = obj

    def __getattr__(self, name):
        fn = getattr(self._obj, name)

        class M:
            remote = staticmethod(fn)

        return M


class _Storage:
    def __init__(self, net):
        self.net = net

    def get(self):
        return self.net


def play(game_class, game_args, cfg, salt, tape_arrays, game_index):
    rh.load()
    Gamer = importlib.import_module("Training.Gamer").Gamer
    RB = importlib.import_module("Training.ReplayBuffer").ReplayBuffer
    with contextlib.redirect_stdout(io.StringIO()):
        probe = game_class(*game_args)
    net = StubNetwork(probe.get_action_space_shape(), salt)
    buf = RB(1000, 8)
    gamer = Gamer(_Remote(buf), _Remote(_Storage(net)), game_class, game_args, game_index, cfg, 2, "disabled")
    tape = rh.TapeRandom(*tape_arrays)
    with rh.parity_patches(tape=tape, identity_softmax=True), contextlib.redirect_stdout(io.StringIO()):
        stats, cache = gamer.play_game()
    assert cache is None
    entries = buf.get_buffer()
    return stats, entries, tape.move + 1


def save(name, stats, entries, cfg, salt, tape_arrays, moves, game_desc, game_index):
    kmax = 1
    states = np.concatenate([np.asarray(e[0], dtype=np.float32) for e in entries])
    policy = np.array([e[1][1] for e in entries], dtype=np.float64)
    value = np.array([e[1][0] for e in entries], dtype=np.float64)
    gidx = np.array([e[2] for e in entries], dtype=np.int64)
    kmax = max(kmax, int((policy > 0).sum(1).max()))
    keys = ["number_of_moves", "average_children", "average_tree_size", "final_tree_size", "average_bias_value", "final_bias_value"]
    assert sorted(stats) == sorted(keys)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(
        path, cfg_yaml=np.array(yaml.safe_dump(cfg)), salt=np.int64(salt), game=np.array(game_desc), game_index=np.int64(game_index),
        stats_keys=np.array(keys), stats=np.array([float(stats[k]) for k in keys], dtype=np.float64),
        states=states, policy=policy, value=value, gidx=gidx,
        gamma_tape=tape_arrays[0][:moves + 1, :max(kmax, 16)], unif_tape=tape_arrays[1][:moves + 1])
    print("%-32s moves=%3d value=%+d entries=%d bytes=%d" % (name, int(stats["number_of_moves"]), int(value[0]), len(entries),
                                                           os.path.getsize(path)))


def main():
    if not rh.available():
        raise SystemExit("reference tree not found; goldens can only be generated in the build container")
    ns = rh.load()
    cases = [
        ("gamer_ttt_s60", ns.tic_tac_toe, [], "ttt", 60, 11, dict(epsilon_random_exploration=0.3)),
        ("gamer_ttt_s200_soft2", ns.tic_tac_toe, [], "ttt", 200, 12, dict(number_of_softmax_moves=2)),
        ("gamer_scs_solo5_s40", ns.SCS_Game, [rh.scs_config_path("solo_soldier_config_5.yml"), 1], "scs:solo_soldier_config_5.yml:1", 40, 13, {}),
        ("gamer_scs_mirrored5_s30", ns.SCS_Game, [rh.scs_config_path("mirrored_config_5.yml")], "scs:mirrored_config_5.yml:None", 30, 14,
         dict(epsilon_softmax_exploration=0.2)),
    ]
    for i, (name, cls, args, desc, sims, salt, over) in enumerate(cases):
        cfg = search_config(sims, **over)
        tape = make_tape(2000 + salt, cfg, 160, 128)
        stats, entries, moves = play(cls, args, cfg, salt, tape, game_index=3 + i)
        save(name, stats, entries, cfg, salt, tape, moves, desc, 3 + i)


if __name__ == "__main__":
    main()
