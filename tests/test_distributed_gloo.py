"""CPU, world_size 2, gloo: the trajectory all-gather (variable-length record words) and the record
codec it carries."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nuzero_b200 import _ffi
from nuzero_b200.engine import parse_records


def encode_record(uid, move, action, player, root_n, root_w, slot, bias, length, child, state, child_actions, child_n,
                  game_end=False, tv=0):
    """Inverse of parse_records (csrc/mcts.cuh write_record) for synthetic test data."""
    K = len(child_actions)
    ln = _ffi.REC_HDR + len(state) + 2 * K
    w = np.zeros(ln, dtype=np.uint32)
    rw = np.array([root_w, bias], dtype=np.float64).view(np.uint32)
    w[0], w[1], w[2] = ln, uid, move | (K << 16)
    w[3] = action | (player << 16) | (((2 if game_end else 0) | ((tv + 1) << 2)) << 24)
    w[4], w[5], w[6], w[7], w[8], w[9], w[10], w[11] = root_n, rw[0], rw[1], slot, rw[2], rw[3], length, child
    w[_ffi.REC_HDR:_ffi.REC_HDR + len(state)] = state
    w[_ffi.REC_HDR + len(state)::2] = child_actions
    w[_ffi.REC_HDR + len(state) + 1::2] = child_n
    return w


def _rank_words(rank):
    rng = np.random.default_rng(rank)
    out = []
    for g in range(3 + 2 * rank):  # ranks hold different amounts of data
        moves = 4 + g
        for m in range(moves):
            K = int(rng.integers(1, 9))
            out.append(encode_record(uid=100 * rank + g, move=m, action=int(rng.integers(9)), player=1 + m % 2,
                                     root_n=50 + m, root_w=0.25 * m, slot=g, bias=1.15 + 0.001 * m, length=m + 1,
                                     child=0, state=[int(rng.integers(1 << 20))], child_actions=np.arange(K),
                                     child_n=rng.integers(0, 30, K), game_end=(m == moves - 1), tv=(g % 3) - 1))
    return np.concatenate(out)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nuzero_b200.distributed import all_gather_records
    from nuzero_b200.selfplay import group_games

    mine = torch.from_numpy(_rank_words(rank).astype(np.int64)).to(torch.int32)
    parts = all_gather_records(mine)
    ok = len(parts) == world
    total_games = 0
    for r, p in enumerate(parts):
        ok &= np.array_equal(p.numpy().view(np.uint32), _rank_words(r))
        games = group_games(parse_records(p.numpy().view(np.uint32), 1))
        ok &= len(games) == 3 + 2 * r
        total_games += len(games)
        for uid, moves in games.items():
            ok &= moves[-1]["game_end"] and moves[-1]["terminal_value"] == ((uid % 100) % 3) - 1
            ok &= [m["move"] for m in moves] == list(range(len(moves)))
    # the indexed form (what SelfPlayRunner.collect feeds to DeviceReplayBuffer.ingest_words): offsets travel with the words
    from nuzero_b200.distributed import all_gather_indexed

    def offsets_of(words):
        out, pos = [], 0
        while pos < len(words):
            out.append(pos)
            pos += int(words[pos])
        return np.array(out, dtype=np.int64)

    parts2 = all_gather_indexed(mine, torch.from_numpy(offsets_of(_rank_words(rank))))
    for r, (w, o) in enumerate(parts2):
        ref = _rank_words(r)
        ok &= np.array_equal(w.numpy().view(np.uint32), ref) and np.array_equal(o.numpy(), offsets_of(ref))
        ok &= o.dtype == torch.int64
    # sharded window (SelfPlayRunner(gather_to=None)): every rank keeps the games (uid + source rank) % world == rank of the
    # union; the shards are disjoint, cover every record and keep whole records in arena order
    from nuzero_b200.replay import DeviceReplayBuffer

    n_owned = 0
    for r, (w, o) in enumerate(parts2):
        ow, oo = DeviceReplayBuffer._owned_records(w, o, r, (rank, world))
        recs = parse_records(ow.numpy().view(np.uint32), 1)
        ok &= all((x["uid"] + r) % world == rank for x in recs)
        ok &= [int(v) for v in oo] == offsets_of(ow.numpy().view(np.uint32)).tolist()
        ref = [x for x in parse_records(_rank_words(r), 1) if (x["uid"] + r) % world == rank]
        ok &= len(ref) == len(recs) and all(a["uid"] == b["uid"] and a["move"] == b["move"] and a["child_N"].tolist() == b["child_N"].tolist()
                                            for a, b in zip(ref, recs))
        n_owned += len(recs)
    t = torch.tensor([n_owned])
    dist.all_reduce(t)
    ok &= int(t) == sum(len(offsets_of(_rank_words(r))) for r in range(world))
    # a step in which one rank recorded nothing (every step of a quiet phase does this to some rank)
    empty = rank == 1
    parts3 = all_gather_indexed(mine[:0] if empty else mine,
                                torch.zeros(0, dtype=torch.int64) if empty else torch.from_numpy(offsets_of(_rank_words(rank))))
    ok &= int(parts3[1][0].numel()) == 0 and int(parts3[1][1].numel()) == 0
    ok &= np.array_equal(parts3[0][0].numpy().view(np.uint32), _rank_words(0))
    q.put((rank, bool(ok), total_games))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_records_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, 8), (1, True, 8)]


def test_record_codec_roundtrip():
    w = encode_record(7, 3, 5, 2, 801, -1.5, 2, 1.234, 4, 1, [123456], [1, 5, 8], [10, 700, 90], True, -1)
    (r,) = parse_records(w, 1)
    assert (r["uid"], r["move"], r["action"], r["player"], r["root_N"], r["slot"], r["length"], r["child"]) == (7, 3, 5, 2, 801, 2, 4, 1)
    assert r["root_W"] == -1.5 and r["bias"] == 1.234 and r["game_end"] and r["terminal_value"] == -1
    assert r["child_actions"].tolist() == [1, 5, 8] and r["child_N"].tolist() == [10, 700, 90] and r["state"].tolist() == [123456]
