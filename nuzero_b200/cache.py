"""Device inference cache: the reference's `Utils/Caches` — `DictCache` maps a state to the network's
(policy, value) and `Explorer.evaluate` consults it before every inference (Explorer.py:146-155) — for the whole leaf batch.

    net = CachedForward(engine, lambda view: FusedRecurrentForward(view, model, iters), capacity_log2=22)
    engine.advance(); net()          # hits are served from HBM, the network runs on the missed rows only

Keys are the compact leaf states (+ scenario map), compared word for word: a hit returns exactly what the network returned
for that state, and every network in this package computes a row independently of the rest of the batch, so the search is
bit-identical with and without the cache (tests/test_gpu_cache.py).  The network runs on the smallest prepared batch size
that holds the missed rows (each size has its own CUDA graph); the number of misses is the one value the host reads per step.
"""
import ctypes as C

import torch

from . import _ffi
from ._ffi import check, lib


class _RowsView:
    """What a batched forward needs from an engine, over the first `rows` rows of the dense staging batch."""

    def __init__(self, engine, rows, leaf, policy, value):
        self.device, self.state_shape, self.A, self.rows, self.G, self.V = engine.device, engine.state_shape, engine.A, rows, rows, 1
        self.c = engine.c
        self._engine = engine
        self.leaf, self.policy, self.value = leaf[:rows], policy[:rows], value[:rows]

    def _stream(self):
        return self._engine._stream()


def batch_ladder(rows, min_rows, in_kernel):
    """Prepared network batch sizes (each one CUDA graph over a prefix of the rows), ascending, always ending with `rows`.
    in_kernel: a fine ladder (x 1.25, multiples of 64) — the missed rows of a launch land a little above any target, and the
    batch the network runs on stays within 25 % of the rows that need it; otherwise the coarse x 4 ladder of the two-kernel form."""
    if in_kernel:
        sizes, r = [rows], float(min_rows)
        while r < rows:
            sizes.append((int(r) + 63) & ~63)
            r *= 1.25
    else:
        sizes, r = [], rows
        while r > min_rows:
            sizes.append(r)
            r = (r + 3) // 4
        sizes.append(min(max(r, 1), rows) if rows < min_rows else max(r, min_rows))
    return sorted(set(min(n, rows) for n in sizes))


class CachedForward:
    needs_host_sync = True  # the host reads the miss count every call: not capturable in a CUDA graph (SelfPlayRunner checks)

    def __init__(self, engine, make_forward, capacity_log2=20, min_rows=256, in_kernel=False, miss_target=0, batch_sizes=None, park_target=0, pipeline=False, publish_width=64):
        """make_forward(view) -> callable that reads view.leaf and writes view.policy / view.value (e.g.
        `lambda v: FusedRecurrentForward(v, model, iters)`); capacity_log2: table slots = 2 ** capacity_log2.
        in_kernel: the SEARCH KERNEL consults the table (nz_engine_attach_cache) — the reference's order, Explorer.evaluate
        asks the cache before the inference (Explorer.py:146-155): a leaf that was evaluated before is expanded inside the
        launch and the game runs on (up to the engine's max_sims_per_launch simulations per launch), only the missed leaves
        wait for the network, in rows 0..n-1 of engine.leaf (dense rows).  Same results, several times fewer launches.
        miss_target (in_kernel): the launch ends once that many leaves wait for the network (a full batch) — games start no
        further simulation then — instead of only after max_sims_per_launch simulations per game; park_target: likewise
        once that many games wait for the network (on a row of their own, or on the row of another game with the same state).
        batch_sizes: the prepared network batch sizes (each one CUDA graph over a prefix of the rows); default: a ladder.
        pipeline (in_kernel): two lanes of leaf / policy / value tensors — launch k searches in lane k & 1 while the network
        evaluates the rows of launch k - 1 on a second stream; a game that parked in launch k continues in launch k + 2.  The
        calling convention stays `engine.advance(); net()`; call `drain()` before reading the tensors from the host.
        publish_width (in_kernel): expansions of up to this many children are published next to their table entry — the first
        game that expands a cached state writes its (action, prior) list, all later expansions of the state copy it instead of
        recomputing legal mask and soft-max (0 = off; games with more children than this per node skip it)."""
        e = self.e = engine
        dev = e.device
        self.in_kernel = bool(in_kernel)
        self.cap_log2 = int(capacity_log2)
        cap = 1 << self.cap_log2
        self.kw = e.state_words + 1
        self.keys = torch.zeros((cap, self.kw), dtype=torch.int32, device=dev)
        self.meta = torch.zeros(cap, dtype=torch.int32, device=dev)
        self.pol = torch.zeros((cap, e.A), dtype=e.policy.dtype, device=dev)
        self.val = torch.zeros(cap, dtype=torch.float32, device=dev)
        self.miss_rows = torch.zeros(e.rows, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(2, dtype=torch.int32, device=dev)
        self._host = torch.zeros(2, dtype=torch.int32).pin_memory()
        if batch_sizes is not None:
            sizes = sorted(set(min(int(n), e.rows) for n in list(batch_sizes) + [e.rows]))
        else:
            sizes = batch_ladder(e.rows, min_rows, in_kernel)
        # one dense staging batch; every prepared batch size is a prefix of it (the look-up kernel writes missed row i's planes
        # to row i, the insert kernel reads the outputs of row i: no gather / scatter launches in between)
        if self.in_kernel:
            # the search kernel itself writes the missed leaves to rows 0..n-1 of the engine's tensors and reads the
            # network's answer from the same row
            self.stage_leaf, self.stage_policy, self.stage_value = e.leaf, e.policy, e.value
            self.row = torch.zeros(cap, dtype=torch.int32, device=dev)
            check(lib().nz_engine_attach_cache(e.h, C.c_void_p(self.keys.data_ptr()), C.c_void_p(self.meta.data_ptr()),
                                               C.c_void_p(self.row.data_ptr()),
                                               C.c_void_p(self.pol.data_ptr()), C.c_void_p(self.val.data_ptr()), self.cap_log2,
                                               int(miss_target), int(park_target)))
            self._hits0 = 0
            if publish_width > 0 and e.c.max_children <= publish_width:
                w = int(e.c.max_children)
                self.exp_meta = torch.zeros(cap, dtype=torch.int32, device=dev)
                self.exp_act = torch.zeros((cap, w), dtype=torch.int16, device=dev)
                self.exp_prior = torch.zeros((cap, w), dtype=torch.float64, device=dev)
                check(lib().nz_engine_attach_expansions(e.h, C.c_void_p(self.exp_meta.data_ptr()), C.c_void_p(self.exp_act.data_ptr()),
                                                        C.c_void_p(self.exp_prior.data_ptr()), w))
        else:
            self.stage_leaf = torch.zeros((e.rows,) + tuple(e.state_shape), dtype=e.leaf.dtype, device=dev)
            self.stage_policy = torch.zeros((e.rows, e.A), dtype=e.policy.dtype, device=dev)
            self.stage_value = torch.zeros((e.rows,), dtype=torch.float32, device=dev)
        self.views = [_RowsView(e, n, self.stage_leaf, self.stage_policy, self.stage_value) for n in sizes]
        self.forwards = [make_forward(v) for v in self.views]
        self.hits = self.misses = self.calls = 0
        self.pipeline = bool(pipeline) and self.in_kernel
        if self.pipeline:
            lane1 = (torch.zeros_like(e.leaf), torch.zeros_like(e.policy), torch.zeros_like(e.value))
            self._lane_tensors = [(e.leaf, e.policy, e.value), lane1]
            views1 = [_RowsView(e, v.rows, *lane1) for v in self.views]
            self._lane_forwards = [self.forwards, [make_forward(v) for v in views1]]
            self._fwd_stream = torch.cuda.Stream(dev, priority=-1)
            self._ev_adv, self._ev_fwd = [None, None], [None, None]
            self._count_host = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(2)]
            self._k = self._done = 0
            e._pre_advance = self._pre_advance

    def _args(self):
        return (self.e.h, C.c_void_p(self.keys.data_ptr()), C.c_void_p(self.meta.data_ptr()), C.c_void_p(self.pol.data_ptr()),
                C.c_void_p(self.val.data_ptr()), self.cap_log2)

    # -- two-lane pipeline: search of launch k overlaps the network call of launch k - 1 ---------------------------------
    def _pre_advance(self):
        e = self.e
        lane = self._k & 1
        if self._ev_fwd[lane] is not None:  # this lane's last network call + insert (launch k - 2) must have finished
            torch.cuda.current_stream(e.device).wait_event(self._ev_fwd[lane])
        e.leaf, e.policy, e.value = self._lane_tensors[lane]
        check(lib().nz_engine_set_lane(e.h, lane))

    def _finish(self, j):
        """Network call for launch j (its search has been enqueued; waits for it on the host to learn the row count)."""
        if j < self._done:
            return
        self._done = j + 1
        e, lane = self.e, j & 1
        self._ev_adv[lane].synchronize()
        n_miss = int(self._count_host[lane][0])
        self.calls += 1
        self.misses += n_miss
        self._ev_fwd[lane] = None
        if n_miss == 0:
            return
        leaf, policy, value = self._lane_tensors[lane]
        k = next(i for i, v in enumerate(self.views) if v.rows >= n_miss)
        with torch.cuda.stream(self._fwd_stream):
            self._fwd_stream.wait_event(self._ev_adv[lane])
            self._lane_forwards[lane][k]()
            check(lib().nz_cache_insert_dense(*self._args(), C.c_void_p(policy.data_ptr()), C.c_void_p(value.data_ptr()), n_miss, lane,
                                              C.c_void_p(self._fwd_stream.cuda_stream)))
            ev = torch.cuda.Event()
            ev.record(self._fwd_stream)
        self._ev_fwd[lane] = ev

    def _call_pipelined(self):
        e = self.e
        lane = self._k & 1
        s = torch.cuda.current_stream(e.device)
        self._count_host[lane].copy_(e.dense_count[4 * lane: 4 * lane + 4], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(s)
        self._ev_adv[lane] = ev
        if self._k >= 1:
            self._finish(self._k - 1)
        self._k += 1

    def drain(self):
        """pipeline: run the network call of the last launch and wait for everything (before the host reads engine state that
        depends on it; the next advance() continues normally)."""
        if self.pipeline and self._k >= 1:
            self._finish(self._k - 1)
            self._fwd_stream.synchronize()

    def _call_in_kernel(self):
        if self.pipeline:
            return self._call_pipelined()
        e = self.e
        self._host[:1].copy_(e.dense_count[:1], non_blocking=True)
        torch.cuda.current_stream(e.device).synchronize()
        n_miss = int(self._host[0])
        self.calls += 1
        self.misses += n_miss
        if n_miss == 0:
            return
        k = next(i for i, v in enumerate(self.views) if v.rows >= n_miss)
        self.forwards[k]()  # rows n_miss.. of the prefix hold older planes: computed and ignored
        check(lib().nz_cache_insert_dense(*self._args(), C.c_void_p(e.policy.data_ptr()), C.c_void_p(e.value.data_ptr()), n_miss, 0, e._stream()))

    def __call__(self):
        if self.in_kernel:
            return self._call_in_kernel()
        e = self.e
        self.counters.zero_()
        check(lib().nz_cache_lookup(*self._args(), C.c_void_p(e.policy.data_ptr()), C.c_void_p(e.value.data_ptr()),
                                    C.c_void_p(self.miss_rows.data_ptr()), C.c_void_p(self.counters.data_ptr()),
                                    C.c_void_p(e.leaf.data_ptr()), C.c_void_p(self.stage_leaf.data_ptr()), e._stream()))
        self._host.copy_(self.counters, non_blocking=True)
        torch.cuda.current_stream(e.device).synchronize()
        n_miss, n_hit = int(self._host[0]), int(self._host[1])
        self.calls += 1
        self.hits += n_hit
        self.misses += n_miss
        if n_miss == 0:
            return
        k = next(i for i, v in enumerate(self.views) if v.rows >= n_miss)
        self.forwards[k]()  # rows n_miss.. of the prefix hold older planes: computed and ignored
        check(lib().nz_cache_insert(*self._args(), C.c_void_p(self.stage_policy.data_ptr()), C.c_void_p(self.stage_value.data_ptr()),
                                    C.c_void_p(self.miss_rows.data_ptr()), n_miss, C.c_void_p(e.policy.data_ptr()),
                                    C.c_void_p(e.value.data_ptr()), e._stream()))

    def hit_rate(self):
        if self.in_kernel:
            c = self.e.counters()
            self.hits = c["cache_hits"] - self._hits0
            self.shared = c["cache_shared"]  # leaves that waited for another game's row of the same launch
            return self.hits / max(1, self.hits + self.misses + self.shared)
        return self.hits / max(1, self.hits + self.misses)

    def detach(self):
        """in_kernel: give the engine back its one-row-per-game leaf tensor."""
        if self.in_kernel:
            check(lib().nz_engine_attach_cache(self.e.h, None, None, None, None, None, 0, 0, 0))

    def clear(self):
        """Forget everything (Network_Manager weights changed: MctsAgent.set_network clears its cache, MctsAgent.py:57-59)."""
        self.meta.zero_()
        if getattr(self, "exp_meta", None) is not None:
            self.exp_meta.zero_()
        self.hits = self.misses = self.calls = 0
        if self.in_kernel:
            self._hits0 = self.e.counters()["cache_hits"]
