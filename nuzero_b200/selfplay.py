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


class SelfPlayRunner:
    """The batched counterpart of AlphaZero.run_selfplay's actor pool (Training/AlphaZero.py:503-594): thousands of games
    advance together on one GPU, finished games land in a DeviceReplayBuffer.

        runner = SelfPlayRunner(engine, net, replay, launches_per_step=256)
        positions = runner.step()        # one CUDA-graph replay of `launches_per_step` (search, network) pairs;
        ...                              # returns the positions the PREVIOUS step added to the replay window
        runner.flush()                   # everything played so far is in the window

    `net` is a callable that reads engine.leaf and writes engine.policy / engine.value on the current stream (a
    GraphedForward / FusedRecurrentForward / DyadicStubNet).  The record arena is append-only between resets, so the
    records of step i are read (on a side stream) while the search of step i+1 runs: the host-side grouping of moves
    into games overlaps with the GPU work.  With `world > 1` every rank's records are first merged with one all-gather
    (distributed.all_gather_indexed); then either rank `gather_to` ingests the union (one window, as the reference's single
    ReplayBuffer actor) or, with `gather_to=None`, the window is sharded: every rank passes its own `replay` and ingests the
    games (uid + source rank) % world == rank of the union, so no rank decodes more than it plays."""

    def __init__(self, engine, net, replay, launches_per_step=64, use_graph=True, rank=0, world=1, gather_to=0, collect_every=1):
        """collect_every: records are collected (and, with world > 1, all-gathered) on every k-th step only — fewer points at
        which the ranks wait for each other; the record arena must hold k steps of moves."""
        self.collect_every, self._step_i = max(1, int(collect_every)), 0
        self.e, self.net, self.replay = engine, net, replay
        self.launches_per_step, self.rank, self.world, self.gather_to = launches_per_step, rank, world, gather_to
        self.graph = None
        self.d2h_bytes = 0
        # replay-side work (record copies, grouping, decode, sampling) on a high-priority stream: its short kernels get the
        # SMs a search CTA frees ahead of the queued search work instead of waiting behind every launch of the step
        self.side = torch.cuda.Stream(engine.device, priority=-1)
        self._tops = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._k = 0
        self._snap = None
        self._read_words = self._read_recs = 0
        self._max_read_words = 0  # largest read cursor over all ranks (identical on every rank: reset decisions agree)
        if use_graph and getattr(net, "needs_host_sync", False):
            raise _ffi.NzError("this network callable synchronises with the host on every call (CachedForward reads its miss "
                               "count) and cannot be captured: use SelfPlayRunner(..., use_graph=False)")
        if use_graph:
            warm = torch.cuda.Stream(engine.device)
            warm.wait_stream(torch.cuda.current_stream(engine.device))
            with torch.cuda.stream(warm):  # warm-up outside the capture (lazy module loads, attribute calls)
                self._pairs(2)
            torch.cuda.current_stream(engine.device).wait_stream(warm)
            torch.cuda.synchronize(engine.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._pairs(launches_per_step)

    def _pairs(self, n):
        for _ in range(n):
            self.e.advance()
            self.net()

    def play(self):
        """The search part of a step only (no host work)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._pairs(self.launches_per_step)

    def _snapshot(self):
        buf = self._tops[self._k]
        self._k ^= 1
        buf.copy_(self.e.arena_top, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.e.device))
        return buf, ev

    def _collect(self, snap):
        """Records written up to the snapshot -> replay buffer (side stream).  Returns the positions that entered."""
        e = self.e
        buf, ev = snap
        ev.synchronize()
        used, dropped, n = min(int(buf[0]), e.c.arena_words), int(buf[1]), min(int(buf[2]), e.rec_index.numel())
        if dropped:
            raise _ffi.NzError("%d move records were dropped: the record arena is too small" % dropped)
        added = 0
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            words = e.arena[self._read_words:used].clone()
            offs = e.rec_index[self._read_recs:n].to(torch.int64) - self._read_words
            self.d2h_bytes += 16  # arena_top; the replay buffer counts the table it reads (DeviceReplayBuffer.d2h_bytes)
            if self.world > 1:
                from .distributed import all_gather_indexed

                parts = all_gather_indexed(words, offs)
                self._peer_words = [a + int(pw[0].numel()) for a, pw in zip(getattr(self, "_peer_words", [0] * self.world), parts)]
                self._max_read_words = max(self._peer_words)
                if self.gather_to is None:
                    added = self.replay.ingest_parts(parts, uid_mul=self.world, owner=(self.rank, self.world))
                elif self.rank == self.gather_to:
                    added = self.replay.ingest_parts(parts, uid_mul=self.world)
            else:
                added = self.replay.ingest_words(words, offs)
            # consumers of the replay window on OTHER streams wait for this event (wait_replay): the rows are written
            # on the side stream, which nothing else is ordered against
            self.replay_ready = torch.cuda.Event()
            self.replay_ready.record(self.side)
        self._read_words, self._read_recs = used, n
        if self.world == 1:
            self._max_read_words = used
        return added

    def wait_replay(self, stream=None):
        """Order `stream` (default: the current stream) after the last ingest into the replay window.  A trainer that samples
        the window with get_sample_tensors() / get_slice_tensors() on its own stream calls this first; step() does not do
        it implicitly because the ingest of step i is meant to overlap the search of step i + 1."""
        ev = getattr(self, "replay_ready", None)
        if ev is not None:
            (stream or torch.cuda.current_stream(self.e.device)).wait_event(ev)

    def collect(self):
        """Synchronous: everything the engine has recorded so far goes into the replay buffer."""
        added = self._collect(self._snapshot())
        self._snap = None
        self.side.synchronize()
        if self._max_read_words > self.e.c.arena_words // 2:
            self._reset_arena()
        return added

    def _reset_arena(self):
        torch.cuda.synchronize(self.e.device)
        self.e.arena_top.zero_()
        self._read_words = self._read_recs = self._max_read_words = 0
        self._peer_words = [0] * self.world
        self._snap = None

    def step(self):
        self._step_i += 1
        if self._step_i % self.collect_every:
            self.play()
            return 0
        prev = self._snap
        self.play()
        self._snap = self._snapshot()
        added = self._collect(prev) if prev is not None else 0
        if self._max_read_words > self.e.c.arena_words // 2:
            added += self.collect()  # rare: drain everything, then restart the append-only arena (all ranks together)
        return added

    def flush(self):
        return self.collect()
