"""The REAL reference `Training/Gamer.play_game` (fixtures from oracle/gen_golden_gamer.py: the unmodified class with the
ray stub, the reference ReplayBuffer as sink) against (a) the oracle port of the loop — CPU — and (b) this repo's
`nuzero_b200.gamer.Gamer` on the GPU: the six statistics (Gamer.py:42-50,81-92) and every tuple that reached
`ReplayBuffer.save_game` (ReplayBuffer.py:24-36): state planes, value target, policy target over all actions, game index."""
import os

import numpy as np
import pytest
import yaml

import golden_io

NAMES = golden_io.names("gamer_")
SCS_CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nuzero_b200", "configs", "scs")


def _load(name):
    z = np.load(os.path.join(golden_io.GOLDEN, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["cfg"] = yaml.safe_load(str(g["cfg_yaml"]))
    g["stats"] = dict(zip([str(k) for k in g["stats_keys"]], g["stats"].tolist()))
    return g


def _game_of(g):
    parts = str(g["game"]).split(":")
    if parts[0] == "ttt":
        return "ttt", None, None
    return "scs", parts[1], (None if parts[2] == "None" else int(parts[2]))


def test_fixtures_exist():
    assert len(NAMES) >= 4, "run python -m oracle.gen_golden_gamer in the build container"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_loop_matches_the_real_gamer(name):
    from oracle import mcts, selfplay
    from oracle import scs as oscs
    from oracle.stubnet_np import stub_forward
    from oracle.ttt import TicTacToe

    g = _load(name)
    kind, cfg_name, seed = _game_of(g)
    if kind == "ttt":
        game, A = TicTacToe(), 9
    else:
        sc = oscs.load_scenario(os.path.join(SCS_CFG, cfg_name), seed)
        game, A = oscs.SCS(sc), sc.A
    salt = int(g["salt"])
    rec = selfplay.play_game(game, lambda s: stub_forward(s, A, salt), g["cfg"], True, True, mcts.ReplayTape(g["gamma_tape"], g["unif_tape"]))
    st = selfplay.stats(rec)
    assert st == g["stats"]  # ints and floats, bit for bit
    np.testing.assert_array_equal(np.stack(rec["states"]).astype(np.float32), g["states"].reshape((len(rec["states"]),) + rec["states"][0].shape))
    np.testing.assert_array_equal(selfplay.policy_targets(rec, A), g["policy"])
    assert (g["value"] == rec["terminal_value"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gamer_dropin_matches_the_real_gamer(name):
    import torch

    from nuzero_b200.gamer import Gamer
    from nuzero_b200.games.device_game import SCS_Game, tic_tac_toe
    from nuzero_b200.replay import ReplayBuffer
    from nuzero_b200.stubnet import StubNetworkManager

    g = _load(name)
    kind, cfg_name, seed = _game_of(g)
    if kind == "ttt":
        cls, args, shape = tic_tac_toe, [], (1, 3, 3)
    else:
        cls, args = SCS_Game, [os.path.join(SCS_CFG, cfg_name), seed]
        shape = cls(*args).get_action_space_shape()
    net = StubNetworkManager(shape, salt=int(g["salt"]), uid_mul=0)

    class Storage:
        def get(self):
            return net

    buf = ReplayBuffer(1000, 8)
    gamer = Gamer(buf, Storage(), cls, args, int(g["game_index"]), g["cfg"], 2, "disabled", pool_nodes=60000,
                  rng_tape=(g["gamma_tape"], g["unif_tape"]))
    stats, cache = gamer.play_game()
    assert cache is None
    assert stats == g["stats"], (stats, g["stats"])
    entries = buf.get_buffer()
    assert len(entries) == g["states"].shape[0] == int(g["stats"]["number_of_moves"])
    for i, (state, (value, policy), gidx) in enumerate(entries):
        assert isinstance(state, torch.Tensor) and state.dtype == torch.float32 and tuple(state.shape) == (1,) + tuple(g["states"].shape[1:])
        np.testing.assert_array_equal(state[0].numpy(), g["states"][i], err_msg="state %d" % i)
        assert value == g["value"][i] and gidx == int(g["gidx"][i])
        np.testing.assert_array_equal(np.asarray(policy, dtype=np.float64), g["policy"][i], err_msg="policy %d" % i)


@pytest.mark.gpu
def test_consecutive_play_game_calls_draw_fresh_random_streams():
    """a fresh engine per call restarted the device generator at game id 0, so play_forever() with unchanged
    weights replayed the same game over and over."""
    from nuzero_b200.gamer import Gamer
    from nuzero_b200.games.device_game import tic_tac_toe
    from nuzero_b200.replay import ReplayBuffer
    from nuzero_b200.stubnet import StubNetworkManager

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 30
    cfg["Exploration"]["epsilon_random_exploration"] = 0.3
    net = StubNetworkManager((1, 3, 3), salt=5, uid_mul=0)  # the same network for every game: only the random draws differ

    class Storage:
        def get(self):
            return net

    gamer = Gamer(ReplayBuffer(1000, 8), Storage(), tic_tac_toe, [], 0, cfg, 2, "disabled", pool_nodes=20000)
    first = [tuple(g.action_history) for g in gamer.play_games(8)[1]]
    second = [tuple(g.action_history) for g in gamer.play_games(8)[1]]
    assert first != second
    singles = {tuple(gamer.play_games(1)[1][0].action_history) for _ in range(6)}
    assert len(singles) > 1
