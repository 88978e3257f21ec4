"""TEST INFRASTRUCTURE — SCS fixtures generated from the UNMODIFIED reference SCS_Game/Explorer
(see oracle/gen_golden.py for the protocol).  Also writes normalised copies of the scenario YAMLs
the fixtures use (input data of the reference, `Games/SCS/Game_configs/*.yml`) so that the tests can
run where the reference tree is absent.
"""
import contextlib
import io
import os

import numpy as np
import yaml

from . import ref_harness as rh
from .gen_golden import GOLDEN, make_tape, play_reference, save, search_config

CFG_DIR = os.path.join(os.path.dirname(GOLDEN), os.pardir, "nuzero_b200", "configs", "scs")  # shipped with the package


def export_config(name):
    """Re-emit the scenario with key order preserved (the loader's RNG draw order depends on it)."""
    os.makedirs(CFG_DIR, exist_ok=True)
    with open(rh.scs_config_path(name)) as f:
        data = yaml.safe_load(f)
    for section in ("Units", "Terrain"):
        for props in data.get(section, {}).values():
            props.pop("image_path", None)  # rendering only
    with open(os.path.join(CFG_DIR, name), "w") as f:
        f.write("# scenario data of the reference (Games/SCS/Game_configs/%s), re-emitted by oracle/gen_golden_scs.py\n" % name)
        yaml.safe_dump(data, f, sort_keys=False, default_flow_style=None)


def scenario_arrays(game):
    """What the reference loader produced for this game instance (terrain ids, VPs)."""
    terr = [[game.terrain_types.index(game.board[i][j].terrain) for j in range(game.columns)] for i in range(game.rows)]
    return dict(
        sc_terrain=np.array(terr, dtype=np.int32),
        sc_terrain_types=np.array([[t.attack_modifier, t.defense_modifier, t.cost] for t in game.terrain_types],
                                  dtype=np.float64),
        sc_vp0=np.array(game.victory_points[0], dtype=np.int32).reshape(-1, 2),
        sc_vp1=np.array(game.victory_points[1], dtype=np.int32).reshape(-1, 2),
    )


def random_playout(name, seed, play_seed, max_steps=2000):
    game = rh.make_scs(name, seed)
    rng = np.random.Generator(np.random.Philox(play_seed))
    rec = dict(actions=[], masks=[], states=[], players=[], stages=[], turns=[], lengths=[])
    with contextlib.redirect_stdout(io.StringIO()):
        while not game.is_terminal() and len(rec["actions"]) < max_steps:
            mask = np.asarray(game.possible_actions()).reshape(-1)
            rec["masks"].append(np.packbits(mask != 0))
            rec["states"].append(game.generate_network_input()[0].numpy().copy())
            rec["players"].append(game.get_current_player())
            rec["stages"].append(game.current_stage)
            rec["turns"].append(game.current_turn)
            rec["lengths"].append(game.get_length())
            valid = np.flatnonzero(mask)
            a = int(valid[rng.integers(len(valid))])
            rec["actions"].append(a)
            game.step(game.get_action_coords(a))
    out = dict(
        game=np.array("scs:%s" % name), seed=np.int64(seed or 0),
        actions=np.array(rec["actions"], dtype=np.int32), masks=np.stack(rec["masks"]),
        states=np.stack(rec["states"]).astype(np.float32), players=np.array(rec["players"], dtype=np.int8),
        stages=np.array(rec["stages"], dtype=np.int8), turns=np.array(rec["turns"], dtype=np.int16),
        lengths=np.array(rec["lengths"], dtype=np.int32),
        terminal=np.int64(game.is_terminal()), terminal_value=np.int64(game.get_terminal_value()),
        winner=np.int64(game.get_winner()), final_length=np.int64(game.get_length()),
        final_state=game.generate_network_input()[0].numpy().copy(),
        final_player=np.int64(game.get_current_player()), final_stage=np.int64(game.current_stage),
        action_shape=np.array(game.get_action_space_shape(), dtype=np.int32),
        state_shape=np.array(game.get_state_shape(), dtype=np.int32),
    )
    out.update(scenario_arrays(game))
    return out


ENV_CASES = [
    ("solo_soldier_config_5.yml", 1, 11), ("mirrored_config_5.yml", None, 12), ("mirrored_config_5.yml", None, 13),
    ("unbalanced_config_5.yml", None, 14), ("randomized_config_5.yml", 3, 15), ("randomized_config_5.yml", 4, 16),
    ("r_unbalanced_config_6.yml", 2, 17), ("test_config.yml", None, 18), ("solo_soldier_config_15.yml", 1, 19),
    ("randomized_config_10.yml", 5, 20), ("mirrored_config_super_soldiers.yml", None, 21),
    ("randomized_config_7.yml", 6, 22), ("unbalanced_config_8.yml", None, 23),
    ("solo_soldier_config_30.yml", 2, 24),
]

MCTS_CASES = [
    # name, config, seed, sims, training, salt, overrides
    ("scs_p0_solo5", "solo_soldier_config_5.yml", 1, 60, False, 11, {}),
    ("scs_p0_mirrored5", "mirrored_config_5.yml", None, 50, False, 12, {}),
    ("scs_p0_randomized5", "randomized_config_5.yml", 3, 30, False, 13, {}),
    ("scs_p0_unbalanced5_vf", "unbalanced_config_5.yml", None, 40, False, 14, dict(value_factor=0.5, pb_c_init=1.4)),
    ("scs_p0_test_s3", "test_config.yml", None, 16, False, 15, {}),
    ("scs_p0_solo15", "solo_soldier_config_15.yml", 1, 40, False, 16, {}),
    ("scs_p1_mirrored5", "mirrored_config_5.yml", None, 40, True, 17, {}),
    ("scs_p1_unbalanced5_eps", "unbalanced_config_5.yml", None, 30, True, 18,
     dict(epsilon_softmax_exploration=0.3, epsilon_random_exploration=0.4, number_of_softmax_moves=4)),
    ("scs_p1_randomized5", "randomized_config_5.yml", 4, 30, True, 19, dict(root_exploration_fraction=0.25)),
    ("scs_p0_solo30", "solo_soldier_config_30.yml", 2, 12, False, 20, {}),
]


def main(only=""):
    names = sorted({c[0] for c in ENV_CASES} | {c[1] for c in MCTS_CASES})
    for n in names:
        export_config(n)
    for cfg_name, seed, play_seed in ENV_CASES:
        name = "scsenv_%s_s%s_p%d" % (cfg_name.replace("_config", "").replace(".yml", ""), seed or 0, play_seed)
        if only and not name.startswith(only):
            continue
        out = random_playout(cfg_name, seed, play_seed)
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-44s steps=%4d tv=%+d bytes=%d" % (name, len(out["actions"]), int(out["terminal_value"]),
                                                  os.path.getsize(path)))
    for name, cfg_name, seed, sims, training, salt, over in MCTS_CASES:
        if only and not name.startswith(only):
            continue
        cfg = search_config(sims, **over)
        game = rh.make_scs(cfg_name, seed)
        extra = scenario_arrays(game)
        tape = make_tape(2000 + salt, cfg, 400, 128)
        with contextlib.redirect_stdout(io.StringIO()):
            rec = play_reference(game, cfg, training, salt, tape, tree_dump_moves=(0, 3, 10), max_moves=400)
        save(name, rec, cfg, training, salt, tape, "scs:%s:%s" % (cfg_name, seed or 0), extra=extra)


if __name__ == "__main__":
    import sys

    main(sys.argv[1] if len(sys.argv) > 1 else "")
