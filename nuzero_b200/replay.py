"""Replay buffers with the interface of Training/ReplayBuffer.py:10-62 (the reference's Ray actor): a window of games,
entries `(state, (value_target, policy_target), game_index)`.

* `ReplayBuffer` — in-process restatement on Python lists (what the drop-in `Gamer.play_game` ships games to).
* `DeviceReplayBuffer` — the same window kept as dense tensors in HBM and filled straight from the search engine's move
  records: the compact records (a few dozen bytes per position) stay on the device, finished games are
  decoded by `nz_replay_decode` (root state -> float32 planes, visit counts -> policy row over all actions) directly
  into the rows of the window, and sampling / slicing return device tensors ready for the training step.  The host only
  reads one small table per ingest (positions, validity and result of each finished game); grouping moves into games,
  ordering and the compaction of the games still in play run on the device.
"""
import ctypes as C
import random

import numpy as np
import torch

from . import _ffi


def late_heavy_probs(num_positions, variation=0.5):
    """Sampling weights that favour recent positions, as AlphaZero.train_with_samples builds them when `late_heavy` is set
    (Training/AlphaZero.py:779-792): a ramp from (1 - variation) / 2 upwards in steps of variation / num_positions,
    normalised; same operations as the reference's loops, so the float64 values are identical.  Pass the result as
    `probs` to get_sample / get_sample_tensors."""
    if num_positions <= 0:
        return np.zeros(0, dtype=np.float64)
    offset = (1 - variation) / 2
    fraction = variation / num_positions
    steps = np.full(num_positions, fraction, dtype=np.float64)
    steps[0] = offset + fraction
    ramp = np.cumsum(steps)             # total += fraction, one position after the other
    total_sum = sum(ramp.tolist())      # the builtin, like the reference (compensated float summation since Python 3.12)
    return ramp / total_sum


def _has_probs(probs):
    # the reference tests `probs != []` (ReplayBuffer.py:41-48) and is only ever handed lists; an ndarray (what
    # late_heavy_probs returns) must not go through that comparison (it broadcasts, and raises on numpy >= 2)
    return probs is not None and len(probs) > 0


class ReplayBuffer:
    def __init__(self, window_size, batch_size):
        self.window_size, self.batch_size = window_size, batch_size
        self.buffer, self.n_games, self.full = [], 0, False

    def save_game(self, game, game_index):  # ReplayBuffer.py:24-36
        if self.n_games >= self.window_size:
            self.full = True
        else:
            self.full = False
            self.n_games += 1
        for i in range(len(game.state_history)):
            entry = (game.get_state_from_history(i), game.make_target(i), game_index)
            if self.full:
                self.buffer.pop(0)
            self.buffer.append(entry)

    def shuffle(self):
        random.shuffle(self.buffer)

    def get_slice(self, start_index, last_index):
        return self.buffer[start_index:last_index]

    def get_sample(self, batch_size, replace, probs):
        args = [len(self.buffer), batch_size, replace] + ([probs] if _has_probs(probs) else [])
        return [self.buffer[i] for i in np.random.choice(*args)]

    def get_buffer(self):
        return self.buffer

    def len(self):
        return len(self.buffer)

    def played_games(self):
        return self.n_games


class WindowRows:
    """Row bookkeeping of the reference's list semantics (ReplayBuffer.py:24-36) for a dense ring: while fewer than
    `window_size` games are held every position is appended; from then on every appended position evicts the oldest one
    (`buffer.pop(0)` per entry — the buffer keeps the LENGTH it had when the window filled, not a number of games).
    `place(n)` returns the physical rows for n new positions; logical index i lives in row `(start + i) % capacity`."""

    def __init__(self, window_size, capacity):
        self.window_size, self.capacity = int(window_size), int(capacity)
        self.start, self.count, self.n_games = 0, 0, 0

    def place(self, n):
        if self.n_games >= self.window_size:
            full = True
        else:
            full = False
            self.n_games += 1
        rows = np.empty(n, dtype=np.int64)
        for i in range(n):  # vectorised below; kept literal for the empty-buffer corner (pop(0) on an empty list raises)
            if full:
                if self.count == 0:
                    raise IndexError("pop from empty list")
                self.start = (self.start + 1) % self.capacity
                self.count -= 1
            if self.count >= self.capacity:
                raise _ffi.NzError("DeviceReplayBuffer capacity (%d positions) exceeded before the game window filled" % self.capacity)
            rows[i] = (self.start + self.count) % self.capacity
            self.count += 1
        return rows

    def place_many(self, counts):
        """Rows for several games appended one after the other (same result as place() per game).  The write position
        start + count advances by one per position whether or not the append evicts, so the rows are one arithmetic run;
        only the split into growing (count += n) and evicting (start += n) games depends on the window."""
        counts = np.asarray(counts, dtype=np.int64)
        total = int(counts.sum())
        rows = (self.start + self.count + np.arange(total, dtype=np.int64)) % self.capacity
        grow = int(max(0, min(len(counts), self.window_size - self.n_games)))
        n_grow = int(counts[:grow].sum())
        if self.count + n_grow > self.capacity:
            raise _ffi.NzError("DeviceReplayBuffer capacity (%d positions) exceeded before the game window filled" % self.capacity)
        if total - n_grow > 0 and self.count + n_grow == 0:
            raise IndexError("pop from empty list")
        self.n_games += grow
        self.count += n_grow
        self.start = (self.start + (total - n_grow)) % self.capacity
        return rows

    def place_many_first(self, counts):
        """place_many without materialising the rows: returns the first row; the rest follow as (first + i) % capacity."""
        first = (self.start + self.count) % self.capacity
        counts = np.asarray(counts, dtype=np.int64)
        total = int(counts.sum())
        grow = int(max(0, min(len(counts), self.window_size - self.n_games)))
        n_grow = int(counts[:grow].sum())
        if self.count + n_grow > self.capacity:
            raise _ffi.NzError("DeviceReplayBuffer capacity (%d positions) exceeded before the game window filled" % self.capacity)
        if total - n_grow > 0 and self.count + n_grow == 0:
            raise IndexError("pop from empty list")
        self.n_games += grow
        self.count += n_grow
        self.start = (self.start + (total - n_grow)) % self.capacity
        return first

    def logical_rows(self, start_index=0, last_index=None):
        lo, hi, _ = slice(start_index, last_index).indices(self.count)
        return (self.start + np.arange(lo, max(lo, hi), dtype=np.int64)) % self.capacity


class DeviceReplayBuffer:
    def __init__(self, engine, window_size, batch_size, capacity, game_index=0, drop_incomplete=False, host_mirror=False):
        """engine: the SearchEngine whose games are stored (gives shapes, the game's static tables and the decode kernel);
        capacity: positions the dense window can hold (>= the positions of `window_size` games);
        drop_incomplete: discard (and count in .games_dropped) finished games whose first moves were recorded before this
        buffer started listening, instead of raising;
        host_mirror: keep the window ALSO in pinned host memory — the reference's sink is a host-side list of
        `(state, (value, policy), game_index)` tuples (Training/ReplayBuffer.py:24-36): every ingest copies the rows it wrote
        device -> host (counted in .d2h_bytes), and `host_tuples` / `host_arrays` read them without touching the device."""
        self.drop_incomplete, self.games_dropped, self.h2d_bytes = drop_incomplete, 0, 0
        self.host_mirror = bool(host_mirror)
        self.e = engine
        self.window_size, self.batch_size = window_size, batch_size
        self.game_index = game_index
        dev = engine.device
        self.rows = WindowRows(window_size, capacity)
        self.states = torch.zeros((capacity,) + tuple(engine.state_shape), dtype=torch.float32, device=dev)
        self.policy = torch.zeros((capacity, engine.A), dtype=torch.float32, device=dev)
        self.value = torch.zeros(capacity, dtype=torch.float32, device=dev)
        self.gidx = torch.zeros(capacity, dtype=torch.int64, device=dev)
        self.uid = torch.zeros(capacity, dtype=torch.int64, device=dev)
        # records of games that are still being played: words on the device, headers on the host
        self.pend_words = torch.zeros(0, dtype=torch.int32, device=dev)
        self.pend_hdr = torch.zeros((0, 5), dtype=torch.int64, device=dev)  # offset, length, uid, move, flags
        self.d2h_bytes = 0
        self.positions_in = 0
        if self.host_mirror:
            self.h_states = torch.zeros(self.states.shape, dtype=torch.float32).pin_memory()
            self.h_policy = torch.zeros(self.policy.shape, dtype=torch.float32).pin_memory()
            self.h_value = torch.zeros(capacity, dtype=torch.float32).pin_memory()
            self.h_gidx = torch.zeros(capacity, dtype=torch.int64).pin_memory()
            self._mirror_event = None

    def _mirror_rows(self, first, n):
        """Device -> pinned host copy of the ring rows [first, first + n) (at most two contiguous runs), asynchronous on the
        current stream; `host_sync()` waits for the last one."""
        cap = self.rows.capacity
        n = min(n, cap)
        first %= cap
        for lo, hi in ((first, min(first + n, cap)), (0, max(0, first + n - cap))):
            if hi > lo:
                self.h_states[lo:hi].copy_(self.states[lo:hi], non_blocking=True)
                self.h_policy[lo:hi].copy_(self.policy[lo:hi], non_blocking=True)
                self.h_value[lo:hi].copy_(self.value[lo:hi], non_blocking=True)
                self.h_gidx[lo:hi].copy_(self.gidx[lo:hi], non_blocking=True)
                self.d2h_bytes += (hi - lo) * (self.states[0].numel() * 4 + self.policy.shape[1] * 4 + 4 + 8)
        self._mirror_event = torch.cuda.Event()
        self._mirror_event.record(torch.cuda.current_stream(self.states.device))

    def host_sync(self):
        if self.host_mirror and self._mirror_event is not None:
            self._mirror_event.synchronize()

    def host_arrays(self, start_index=0, last_index=None):
        """(states [n, C, R, Cc] f32, value targets [n], policy targets [n, A], game index [n]) of the logical entries
        [start_index:last_index] as numpy views / copies of the pinned host mirror."""
        self.host_sync()
        rows = self.rows.logical_rows(start_index, last_index)
        return (self.h_states.numpy()[rows], self.h_value.numpy()[rows], self.h_policy.numpy()[rows], self.h_gidx.numpy()[rows])

    def host_tuples(self, start_index=0, last_index=None):
        """The reference's list entries `(state [1, C, R, Cc], (value_target, policy_target), game_index)` from the host mirror."""
        st, v, p, g = self.host_arrays(start_index, last_index)
        return [(torch.from_numpy(st[i:i + 1]), (float(v[i]), p[i].tolist()), int(g[i])) for i in range(st.shape[0])]

    # -- filling ------------------------------------------------------------------------------------------------
    def ingest(self, engine=None, uid_mul=1, uid_add=0):
        """Take every record the engine has written since the last call (resets its arena).  Returns the number of
        positions that entered the window (positions of games that finished)."""
        e = engine or self.e
        top = e.arena_top.cpu()
        used, dropped, n = min(int(top[0]), e.c.arena_words), int(top[1]), int(top[2])
        if dropped:
            raise _ffi.NzError("%d move records were dropped: the record arena is too small" % dropped)
        n = min(n, e.rec_index.numel())
        words = e.arena[:used].clone()
        offs = e.rec_index[:n].to(torch.int64)
        e.arena_top.zero_()
        return self.ingest_words(words, offs, uid_mul, uid_add)

    def ingest_words(self, words, offsets, uid_mul=1, uid_add=0):
        """words: int32 device tensor of move records; offsets: int64 device tensor, first word of each record.
        uid_mul / uid_add make game ids unique across ranks (distributed.global_game_index).  Everything that scales with
        the number of records (grouping moves into games, ordering, compaction of the games still in play) runs on the
        device; the host reads one small table per call (positions, validity and result of each finished game)."""
        return self.ingest_parts([(words, offsets)], uid_mul, [uid_add])

    @staticmethod
    def _owned_records(words, offsets, add, owner):
        """The records of `words` whose game belongs to this rank's shard of the window: game `uid` of source rank `add`
        goes to rank (uid + add) % world, so every rank's shard holds 1 / world of EVERY rank's games."""
        rank, world = owner
        uid = words[offsets + 1].to(torch.int64) & 0xFFFFFFFF
        keep = torch.nonzero((uid + add) % world == rank)[:, 0]
        if int(keep.numel()) == int(offsets.numel()):
            return words, offsets
        src = offsets[keep]
        lens = words[src].to(torch.int64)
        new_off = torch.cumsum(lens, 0) - lens
        total = int(lens.sum())
        idx = torch.repeat_interleave(src - new_off, lens, output_size=total) + torch.arange(total, device=words.device)
        return words[idx], new_off

    def ingest_parts(self, parts, uid_mul=1, uid_adds=None, owner=None):
        """Several (words, offsets) pairs at once — the ranks' records after distributed.all_gather_indexed — with ONE pass
        over the pending table: part r's game ids become uid * uid_mul + uid_adds[r].  owner = (rank, world) keeps only this
        rank's shard of the games (the window is then sharded over the ranks: each holds and decodes 1 / world of the union)."""
        uid_adds = list(range(len(parts))) if uid_adds is None else uid_adds
        if owner is not None and owner[1] > 1:
            parts = [self._owned_records(w, o.to(torch.int64), a, owner) if int(o.numel()) else (w, o) for (w, o), a in zip(parts, uid_adds)]
        live = [(w, o, a) for (w, o), a in zip(parts, uid_adds) if int(o.numel())]
        if live:
            dev = live[0][0].device
            # all parts in one pass (a handful of launches whatever the number of ranks): part r's records sit behind the
            # pending words and the parts before it
            n_words = [int(w.numel()) for w, _, _ in live]
            n_recs = [int(o.numel()) for _, o, _ in live]
            bases = np.concatenate([[0], np.cumsum(n_words)[:-1]]) + int(self.pend_words.numel())
            per_part = torch.tensor(np.stack([bases, np.array([a for _, _, a in live])], 1), dtype=torch.int64, device=dev)
            per_rec = torch.repeat_interleave(per_part, torch.tensor(n_recs, device=dev), dim=0, output_size=sum(n_recs))
            self.pend_words = torch.cat([self.pend_words] + [w for w, _, _ in live])
            # the index is filled by a second atomic counter, so its order can differ from the arena order by a few
            # records; games enter the window in the order their last record sits in the arena (rank by rank)
            offsets, perm = torch.sort(torch.cat([o for _, o, _ in live]) + per_rec[:, 0])
            add = per_rec[perm, 1]
            hdr = self.pend_words[offsets[:, None] + torch.arange(4, device=dev)[None, :]].to(torch.int64) & 0xFFFFFFFF
            new_hdr = torch.stack([offsets, hdr[:, 0], hdr[:, 1] * uid_mul + add, hdr[:, 2] & 0xFFFF, hdr[:, 3] >> 24], 1)
            self.pend_hdr = torch.cat([self.pend_hdr, new_hdr])
        return self._commit_finished()

    def _commit_finished(self):
        h = self.pend_hdr  # [n, 5] int64 on the device: offset, length, uid, move, flags (bit 1 = game end, bits 2-3 = tv + 1)
        if h.shape[0] == 0:
            return 0
        dev = h.device
        ends = torch.nonzero((h[:, 4] & 2) != 0)[:, 0]  # end records, in the order the games finished
        n_end = int(ends.numel())
        if n_end == 0:
            return 0
        # positions of finished games, ordered by (finish order, move): the order save_game appends them in
        sorted_uid, by_uid = torch.sort(h[ends, 2], stable=True)
        pos = torch.searchsorted(sorted_uid, h[:, 2].contiguous()).clamp_(max=n_end - 1)
        fin_mask = sorted_uid[pos] == h[:, 2]
        fin = torch.nonzero(fin_mask)[:, 0]
        rank = by_uid[pos[fin]]                       # finish order of the game each position belongs to
        order = torch.argsort(rank * 65536 + h[fin, 3], stable=True)
        fin, rank = fin[order], rank[order]
        counts = torch.bincount(rank, minlength=n_end)
        good = counts == h[ends, 3] + 1
        tv = ((h[ends, 4] >> 2) & 3) - 1
        meta = torch.stack([counts, good.to(torch.int64), tv], 1).cpu().numpy()  # the one table the host reads
        self.d2h_bytes += meta.nbytes
        counts_h, good_h = meta[:, 0], meta[:, 1].astype(bool)
        if not good_h.all():
            if not self.drop_incomplete:
                raise _ffi.NzError("move records of a finished game are missing")
            self.games_dropped += int((~good_h).sum())
            sel = good[rank]
            fin, rank = fin[sel], rank[sel]
            counts_h = counts_h[good_h]
        n_in = int(counts_h.sum())
        if n_in:
            # rows of the window, game by game (the reference decides per GAME whether the window is full): one run
            first = self.rows.place_many_first(counts_h)
            skip = max(0, n_in - self.rows.capacity)  # one ingest may wrap the ring: a row keeps its LAST writer
            rows = (first + torch.arange(skip, n_in, device=dev, dtype=torch.int64)) % self.rows.capacity
            fin_w, rank_w = fin[skip:], rank[skip:]
            off_t = h[fin_w, 0].contiguous()
            _ffi.check(_ffi.lib().nz_replay_decode(self.e.h, C.c_void_p(self.pend_words.data_ptr()), C.c_void_p(off_t.data_ptr()),
                                                   C.c_void_p(rows.data_ptr()), C.c_void_p(self.states.data_ptr()),
                                                   C.c_void_p(self.policy.data_ptr()), int(rows.numel()), self.e._stream()))
            self.value[rows] = tv[rank_w].to(torch.float32)
            self.gidx[rows] = self.game_index
            self.uid[rows] = h[fin_w, 2]
            self.positions_in += n_in
            if self.host_mirror:
                self._mirror_rows(first + skip, n_in - skip)
        # keep the records of the games still in play, compacted
        keep = torch.nonzero(~fin_mask)[:, 0]
        if keep.numel() == 0:
            self.pend_words = self.pend_words[:0]
            self.pend_hdr = h[:0]
        else:
            kept = h[keep]
            lens = kept[:, 1]
            new_off = torch.cumsum(lens, 0) - lens
            total = int(lens.sum())
            idx = torch.repeat_interleave(kept[:, 0] - new_off, lens, output_size=total) + torch.arange(total, device=dev)
            self.pend_words = self.pend_words[idx]
            kept = kept.clone()
            kept[:, 0] = new_off
            self.pend_hdr = kept
        return n_in

    def save_game(self, game, game_index):
        """Compatibility path (ReplayBuffer.py:24-36) for a game object that carries float tensors on the host."""
        n = len(game.state_history)
        dst = torch.from_numpy(self.rows.place(n)).to(self.states.device)
        st = torch.cat([torch.as_tensor(game.get_state_from_history(i)).reshape((1,) + tuple(self.e.state_shape)) for i in range(n)])
        targets = [game.make_target(i) for i in range(n)]
        self.states[dst] = st.to(self.states.device, torch.float32)
        self.policy[dst] = torch.tensor([t[1] for t in targets], dtype=torch.float32).to(self.policy.device)
        self.value[dst] = torch.tensor([t[0] for t in targets], dtype=torch.float32).to(self.value.device)
        self.gidx[dst] = game_index
        self.uid[dst] = -1
        self.positions_in += n

    # -- reading (device tensors) -------------------------------------------------------------------------------
    def _rows(self, start_index=0, last_index=None):
        """Physical rows of the logical entries [start_index:last_index] (python slice semantics), built on the device."""
        lo, hi, _ = slice(start_index, last_index).indices(self.rows.count)
        idx = torch.arange(lo, max(lo, hi), device=self.states.device, dtype=torch.int64)
        return (idx + self.rows.start) % self.rows.capacity

    def tensors(self, rows):
        """(states [n, C, R, Cc], value targets [n], policy targets [n, A], game index [n]) of the given physical rows."""
        return self.states[rows], self.value[rows], self.policy[rows], self.gidx[rows]

    def get_slice_tensors(self, start_index, last_index):
        return self.tensors(self._rows(start_index, last_index))

    def get_sample_tensors(self, batch_size, replace=True, probs=None):
        """np.random.choice(len, batch_size, replace[, probs]) like ReplayBuffer.py:41-48, drawn on the device."""
        n = self.rows.count
        dev = self.states.device
        if _has_probs(probs):
            pick = torch.multinomial(torch.as_tensor(probs, dtype=torch.float64, device=dev), batch_size, replacement=bool(replace))
        elif replace:
            pick = torch.randint(n, (batch_size,), device=dev)
        else:
            pick = torch.randperm(n, device=dev)[:batch_size]
        return self.tensors((pick + self.rows.start) % self.rows.capacity)

    def shuffle(self):
        """random.shuffle of the logical order (ReplayBuffer.py:38-39): one gather per tensor on the device."""
        rows = self._rows()
        perm = rows[torch.randperm(rows.numel(), device=rows.device)]
        for t in (self.states, self.policy, self.value, self.gidx, self.uid):
            t[rows] = t[perm]

    # -- the reference's list-shaped views (slow; for drop-in callers and tests) ----------------------------------
    def _tuples(self, rows):
        st, v, p, g = (x.cpu() for x in self.tensors(rows))
        return [(st[i:i + 1], (float(v[i]), p[i].tolist()), int(g[i])) for i in range(st.shape[0])]

    def get_slice(self, start_index, last_index):
        return self._tuples(self._rows(start_index, last_index))

    def get_sample(self, batch_size, replace, probs):
        n = self.rows.count
        args = [n, batch_size, replace] + ([probs] if _has_probs(probs) else [])
        pick = np.random.choice(*args)
        return self._tuples(self._rows()[torch.from_numpy(np.asarray(pick, dtype=np.int64)).to(self.states.device)])

    def get_buffer(self):
        return self._tuples(self._rows())

    def len(self):
        return self.rows.count

    def played_games(self):
        return self.rows.n_games
