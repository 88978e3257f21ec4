"""Batched self-play driver: the loop of Training/Gamer.py:64-79 for thousands of games at once.

Each iteration is two launches on one stream — the search kernel (`nz_advance`) and the network
forward — with no host synchronisation in between; the loop can be captured in a CUDA graph.
"""
import numpy as np
import torch

from . import _ffi


def run_until_idle(engine, net, max_launches=1_000_000, check_every=64):
    """Advance until every slot has played its quota (auto mode with games_per_slot > 0)."""
    for it in range(max_launches):
        engine.advance()
        net()
        if (it + 1) % check_every == 0:
            ph = engine.phases()
            if bool(((ph == _ffi.PHASE_IDLE) | (ph == _ffi.PHASE_ERROR)).all()):
                break
    engine.raise_on_error()
    ph = engine.phases()
    if not bool((ph == _ffi.PHASE_IDLE).all()):
        raise _ffi.NzError("self-play did not finish within %d launches" % max_launches)


def group_games(records):
    """records (engine.drain_records) -> {uid: [move records in order]} for finished games only."""
    games = {}
    for r in records:
        games.setdefault(r["uid"], []).append(r)
    done = {}
    for uid, moves in games.items():
        moves.sort(key=lambda r: r["move"])
        if moves[-1]["game_end"] and [m["move"] for m in moves] == list(range(len(moves))):
            done[uid] = moves
    return done


def game_record(moves, env=None, map_id=None):
    """One finished game in the schema of oracle/selfplay.py (for parity tests and for building
    replay tuples).  With `env` (EnvOps) the root states are re-encoded to the reference's float32
    network input and legal masks."""
    rec = dict(
        actions=[m["action"] for m in moves],
        root_N=[m["root_N"] for m in moves],
        root_W=[m["root_W"] for m in moves],
        bias=[m["bias"] for m in moves],
        n_children=[m["n_children"] for m in moves],
        child_actions=[m["child_actions"] for m in moves],
        child_N=[m["child_N"] for m in moves],
        players=[m["player"] for m in moves],
        terminal_value=moves[-1]["terminal_value"],
        length=moves[-1]["length"],
        slot=moves[0]["slot"],
        uid=moves[0]["uid"],
        trees={},
    )
    if "child_W" in moves[0]:
        rec["child_W"] = [m["child_W"] for m in moves]
        rec["child_prior"] = [m["child_prior"] for m in moves]
    if env is not None:
        st = torch.from_numpy(np.stack([m["state"] for m in moves]).astype(np.int64)).to(torch.int32).to(env.e.device)
        maps = None if map_id is None else [map_id] * len(moves)
        rec["states"] = list(env.encode(st, maps).cpu().numpy())
        rec["masks"] = [np.packbits(m != 0) for m in env.mask(st, maps).cpu().numpy()]
    return rec
