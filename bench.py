#!/usr/bin/env python
"""Headline benchmark: MCTS simulations/s of batched Tic-Tac-Toe self-play
(BASELINE.json configs[1]: 800 sims/move, 16384 concurrent games per B200, deterministic stub net).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA engine
    python bench.py --impl reference [...]                      # CPU arm: the oracle port of the
                                                                 # reference Explorer/Gamer on host cores

One "step" = one CUDA-graph replay of `--inner` (search launch + network forward) pairs over all
game slots.  `value` counts every simulation all ranks completed inside the timed region divided by
the max-over-ranks device time (CUDA events).  Games restart when they end (steady state); every
move appends its trajectory record to the device arena, which is drained to the host between steps
in the e2e leg.
"""
import argparse
import json
import math
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mcts_sims_per_sec"
UNIT = "sims/s"


def load_cfg(sims):
    import yaml

    cfg = yaml.safe_load(open(os.path.join(ROOT, "nuzero_b200", "configs", "a1_search_config.yaml")))
    cfg["Simulation"]["mcts_simulations"] = sims
    return cfg


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML in-process (two cheap queries per sample, no power
    read-out): the `nvidia-smi -lms` loop of the profiling recipe cost 8 % of a 100 ms timed region on this host; it
    remains the fallback when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.02):
        self.rows, self.proc, self.idx, self.period = [], None, gpu_index, period_s
        self.nvml, self.stop_flag, self.samples, self.reasons, self.mx = None, False, [], set(), None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remapping: take the handle by the UUID / PCI bus id torch reports
            import torch

            bus = torch.cuda.get_device_properties(self.idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.idx), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(hh).bus) == int(bus):
                        h = hh
                        break
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nvml, self.h = pynvml, h
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", 0x8), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", 0x20), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", 0x4))
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                bits = int(get_reasons(self.h))
                for name, _sym, mask in names:
                    if bits & mask:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = sorted(self.samples)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the REFERENCE ITSELF on the host cores.  oracle/ref_harness.py imports the unmodified tree — the sources under
# $NUZERO_REFERENCE or /root/reference where they exist, else oracle/_ref, the same sources compiled to sourceless byte
# code by oracle/build_ref.py (that is what travels to the GPU box) — and P pinned processes loop the reference's own
# `Training/Gamer.play_game` (Gamer.py:39-97: Explorer.run_mcts, game.step, ReplayBuffer.save_game), which is what the
# reference's Ray actors run (AlphaZero.py:525-578) minus the pickling.  Ray is the pass-through stub of oracle/stubs.
# Only when no reference is present at all does the arm fall back to the oracle port (kind "port").
# ------------------------------------------------------------------------------------------------
def _reference_present():
    try:
        from oracle import ref_harness as rh

        return rh.available(), rh.is_bytecode()
    except Exception:
        return False, False


def _pin(core):
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass


def _ref_player(kind, idx, core, spec, shared):
    """One CPU Gamer: plays games forever, publishing completed moves / games; the parent stops it."""
    _pin(core)
    import importlib

    import numpy as np
    import torch

    torch.set_num_threads(1)
    torch.cuda.is_available = lambda: False  # Network_Manager.check_devices would move the model to the GPU (this is the CPU arm)
    sims_done, moves, games, t0 = shared
    seed = 1000 * (idx + 1)
    np.random.seed(seed)
    torch.manual_seed(0)
    cfg = load_cfg(spec["sims"])
    if kind == "port":
        from oracle import mcts, selfplay
        from oracle.stubnet_np import stub_forward
        from oracle.ttt import TicTacToe

        t0[idx] = time.time()
        while True:
            rec = selfplay.play_game(TicTacToe(), lambda s, sl=seed + int(games[idx]): stub_forward(s, 9, sl), cfg, True, True, keep_states=True)
            moves[idx] += rec["length"]
            sims_done[idx] += rec["length"] * spec["sims"]
            games[idx] += 1
    from oracle import ref_harness as rh
    from oracle.gen_golden_gamer import _Remote, _Storage
    from oracle.stubnet_np import StubNetwork

    ns = rh.load()
    Gamer = importlib.import_module("Training.Gamer").Gamer
    ReplayBuffer = importlib.import_module("Training.ReplayBuffer").ReplayBuffer
    if spec["game"] == "ttt":
        game_class, game_args, iters = ns.tic_tac_toe, [], 2
        storage = _Storage(None)
        storage.get = lambda: StubNetwork((1, 3, 3), seed + int(games[idx]))  # a different deterministic network per game
        identity_softmax = True   # the stub emits probabilities (parity protocol)
    else:
        from nuzero_b200.nets import RecurrentNet, initialize_parameters

        Network_Manager = importlib.import_module("Neural_Networks.Network_Manager").Network_Manager
        game_class, game_args, iters = ns.SCS_Game, [rh.scs_config_path(spec["config"])], spec["iters"]
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            probe = game_class(*game_args)
        # hexagdly is not installed anywhere here: the architecture is this repo's restatement of the reference's
        # RecurrentNet (nuzero_b200/nets.py), fp32, batch 1, inside the reference's own Network_Manager.inference
        model = RecurrentNet(probe.get_state_shape()[0], probe.get_action_space_shape()[0], spec["filters"], 2, recall=True,
                             policy_head="conv", value_head="reduce", value_activation="relu", hex=True)
        initialize_parameters(model)
        storage = _Storage(Network_Manager(model))
        identity_softmax = False  # logits -> the reference's scipy softmax
    buf = ReplayBuffer(200, 8)
    gamer = Gamer(_Remote(buf), _Remote(storage), game_class, game_args, idx, cfg, iters, "disabled")
    run = gamer.explorer.run_mcts

    def counted(*a, **k):
        out = run(*a, **k)
        moves[idx] += 1
        return out

    gamer.explorer.run_mcts = counted
    evaluate = gamer.explorer.evaluate  # called exactly once per simulation (Explorer.py:49-62)

    def counted_sim(*a, **k):
        sims_done[idx] += 1
        return evaluate(*a, **k)

    gamer.explorer.evaluate = counted_sim
    import contextlib
    import io

    with rh.parity_patches(tape=None, identity_softmax=identity_softmax), contextlib.redirect_stdout(io.StringIO()):
        t0[idx] = time.time()
        while True:
            gamer.play_game()
            games[idx] += 1


SCS5_MEAN_MOVES_PER_GAME = 76.35  # mirrored_config_5 under the seed-0 RecurrentNet-256x6: 312 748 positions / 4 096 games (CUDA arm)


def cpu_arm(spec, seconds, procs=None, windows=1, warm_windows=0):
    """spec: {"game": "ttt" | "scs", "sims": ..., ("config", "filters", "iters")}.  Starts P pinned Gamer processes ONCE, lets
    them play, and reads their counters over `warm_windows` untimed and `windows` timed windows of `seconds` each: sims/s
    (Explorer.evaluate calls: one per simulation), moves/s (Explorer.run_mcts calls) and finished games."""
    procs = procs or os.cpu_count() or 1
    present, bytecode = _reference_present()
    kind = "reference" if present else "port"
    if kind == "port" and spec["game"] != "ttt":
        return {"value": None, "unit": UNIT, "cores": procs, "kind": "port", "sample": "no reference present; the SCS port arm is not built"}
    ctx = mp.get_context("fork")
    sims_done, moves, games = (ctx.Array("q", procs, lock=False) for _ in range(3))
    t0 = ctx.Array("d", procs, lock=False)
    ncore = os.cpu_count() or 1
    ps = [ctx.Process(target=_ref_player, args=(kind, i, i % ncore, spec, (sims_done, moves, games, t0)), daemon=True) for i in range(procs)]
    wall0 = time.time()
    for p_ in ps:
        p_.start()
    deadline = time.time() + 180.0
    while time.time() < deadline and any(t0[i] == 0.0 and ps[i].is_alive() for i in range(procs)):
        time.sleep(0.05)
    startup = time.time() - wall0
    per_window = []
    for w in range(warm_windows + windows):
        a_t, a = time.time(), (sum(sims_done), sum(moves), sum(games))
        time.sleep(seconds)
        b_t, b = time.time(), (sum(sims_done), sum(moves), sum(games))
        if w >= warm_windows:
            per_window.append([(y - x) / (b_t - a_t) for x, y in zip(a, b)] + [b_t - a_t, b[0] - a[0], b[1] - a[1], b[2] - a[2]])
    starts = list(t0)
    dead = [i for i in range(procs) if not ps[i].is_alive()]
    for p_ in ps:
        p_.terminate()
    for p_ in ps:
        p_.join(timeout=5)
    if dead or any(st == 0.0 for st in starts):
        raise RuntimeError("CPU arm: %d of %d reference Gamer processes died or never started" % (len(dead) + sum(st == 0.0 for st in starts), procs))
    n = len(per_window)
    rate_sims, rate_moves, rate_games = (sum(w[k] for w in per_window) / n for k in range(3))
    tot_s, tot_sims, tot_moves, tot_games = (sum(w[k] for w in per_window) for k in range(3, 7))
    what = ("the UNMODIFIED reference (%s): Training/Gamer.play_game -> Search/Explorer.run_mcts -> Training/ReplayBuffer.save_game"
            % ("byte code of its sources, oracle/_ref" if bytecode else "source tree")) if kind == "reference" else \
        "the oracle PORT of the reference loop (no reference tree or oracle/_ref present)"
    net = "dyadic stub network" if spec["game"] == "ttt" else \
        "RecurrentNet-%d x%d (this repo's restatement of the hexagdly architecture) fp32 batch 1 through the reference's Network_Manager" % (spec["filters"], spec["iters"])
    return {"value": rate_sims, "unit": UNIT, "cores": procs, "kind": kind, "per_window": [w[0] for w in per_window],
            "moves_per_sec": rate_moves, "games_finished": int(tot_games), "games_per_sec_finished": rate_games,
            "sample": "%d pinned processes x %s, %s, %d sims/move, training=True, %d window(s) of %.1f s of continuous play (%d simulations = "
                      "Explorer.evaluate calls, %d moves, %d finished games; %.1f s process start-up and %d warm-up window(s) excluded; wall %.1f s)"
                      % (procs, what, net, spec["sims"], n, seconds, tot_sims, tot_moves, tot_games, startup, warm_windows, time.time() - wall0)}


def cpu_baseline(sims, seconds, procs=None, **kw):
    return cpu_arm({"game": "ttt", "sims": sims}, seconds, procs, **kw)


def cpu_baseline_scs(config, sims, filters, iters, seconds, procs=None, mean_moves_per_game=None):
    """SCS games take minutes per core on the CPU: the arm measures simulations/s and moves/s of whole reference moves and
    derives games/s = sims/s / (sims per move x mean moves per game), the game length being the CUDA arm's measurement of the
    same scenario and network (a full generation) when available."""
    out = cpu_arm({"game": "scs", "sims": sims, "config": config, "filters": filters, "iters": iters}, seconds, procs)
    mm = mean_moves_per_game or SCS5_MEAN_MOVES_PER_GAME
    if out.get("value"):
        out["games_per_sec_derived"] = out["value"] / (sims * mm)
        out["mean_moves_per_game_assumed"] = mm
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(1.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    t0 = time.perf_counter()
    last = cpu_baseline(args.sims, per_step, windows=max(1, args.steps), warm_windows=args.warmup)
    dt = time.perf_counter() - t0
    v = last["value"]
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "wall_seconds": dt,
           "config": workload_config(args, max(1, args.gpus)), "cpu_baseline": last,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    if not args.no_secondary:
        # the SCS half of BASELINE.json's metric on the same host cores: whole moves of the reference Gamer with the real network
        sec = cpu_baseline_scs(args.scs_config, args.scs_sims, args.filters, args.iters, args.cpu_seconds)
        out["secondary"] = {"metric": "selfplay_games_per_sec", "unit": "games/s", "value": sec.get("games_per_sec_derived"),
                            "sims_per_sec": sec["value"], "moves_per_sec": sec.get("moves_per_sec"), "config": scs_config(args, max(1, args.gpus)),
                            "cpu_baseline": sec}
    print(json.dumps(out))


def workload_config(args, world):
    return {"workload": "tic_tac_toe_selfplay_800sims_16384games_stubnet" if (args.sims, args.games) == (800, 16384)
            else "tic_tac_toe_selfplay_%dsims_%dgames_stubnet" % (args.sims, args.games),
            "game": "Tic_Tac_Toe", "sims_per_move": args.sims, "concurrent_games_per_gpu": args.games,
            "network": "deterministic dyadic stub (CUDA kernel)", "training": True, "keep_subtree": True,
            "search_config": "a1_search_config (pb_c_base 10000, pb_c_init 1.15, noise 0.2/0.15)",
            "launch_pairs_per_step": args.inner * args.reps, "launch_pairs_per_graph": args.inner, "max_sims_per_launch": args.budget,
            "leaves_in_flight_per_game": max(1, args.virtual_loss),
            "l2_policy": "node pools (%.1f GB/GPU) exceed the 126 MB L2; no flush" % (args.games * args.pool * 32 / 1e9),
            "parallelism": "independent game batches per GPU, no collective on the search path (x%d)" % world}


def scs_config(args, world):
    return {"workload": "scs_%s_%dsims_%dgames_recurrentnet%d_x%d" % (args.scs_config.replace(".yml", ""), args.scs_sims, args.scs_games,
                                                                     args.filters, args.iters),
            "game": "SCS", "scenario": args.scs_config, "sims_per_move": args.scs_sims, "concurrent_games_per_gpu": args.scs_games,
            "network": "RecurrentNet(C, planes, %d filters, 2 blocks, recall, hex) x %d iterations, random init (xavier, seed 0)" % (args.filters, args.iters),
            "training": True, "keep_subtree": True,
            "parallelism": "independent game batches per GPU, no collective on the search path; trajectories all-gathered into the "
                           "rank-sharded replay window (x%d)" % world}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
SURVEY_BYTES_PER_SIM_TTT = 520.0  # algorithmic HBM bytes per simulation of this workload (0.52 KB)


def algorithmic_bytes(d, leaf_elem_bytes, A, leaf_elems):
    """HBM traffic of the search data structure for the COUNTED work, without any per-launch term:
    select reads 28 B per scanned child (prior 8, W 8, N 4, child range + action 8) + the root header (12 B); backup
    reads+writes N and W of every path node (24 B); expand writes 28 B per created child, reads the policy row and value,
    writes the leaf's child range, saves/restores the path; the encoder writes one leaf row."""
    sims, levels, scanned = d["sims"], d["levels"], d["scanned"]
    exp, created = d["expansions"], d["created"]
    b = scanned * 28 + sims * 12
    b += (levels + sims) * 24
    b += created * 28 + exp * (A * 4 + 4 + 16 + leaf_elems * leaf_elem_bytes + 8)
    b += (levels + exp) * 8  # path save + restore for simulations that wait on the network
    return b


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.stubnet import DyadicStubNet

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = load_cfg(args.sims)
    G = args.games
    e = SearchEngine(tic_tac_toe_spec(), cfg, G, True, device=dev, pool_nodes=args.pool, policy_is_prob=True,
                     leaf_dtype=_ffi.BF16, policy_dtype=_ffi.F32, auto_advance=True, games_per_slot=0,
                     max_sims_per_launch=args.budget, seed=1234 + rank, arena_words=args.arena_words,
                     virtual_loss=args.virtual_loss)
    # the per-slot salts of the stub network are the step's host input: uploaded from pinned memory in the e2e leg
    salts_host = torch.arange(G, dtype=torch.int32).mul_(7919).remainder_(65521).pin_memory()
    net = DyadicStubNet(e, salt=salts_host, uid_mul=1 if args.virtual_loss <= 1 else 0)

    def pair():
        e.advance()
        net()

    # de-synchronise the slots (games at all stages) before capturing / timing
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        # all slots start their first game together and every launch is exactly one simulation per game, so the batch
        # moves through the game phases in lock-step (cheap late-game trees, expensive openings) until the differing game
        # lengths have spread the slots: ~10 games per slot before anything is timed
        for i in range(args.presteps):
            pair()
            if i % 4096 == 4095:
                e.arena_top.zero_()  # nobody reads the warm-up trajectories
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    e.raise_on_error()
    e.arena_top.zero_()

    # public API of the batched path: SelfPlayRunner = CUDA-graph replay of `inner` (search, network) launch pairs, then the
    # finished games' records are decoded on the device into the replay window (DeviceReplayBuffer) and mirrored into
    # pinned host memory: the reference's sink is a host-side list of tuples (Training/ReplayBuffer.py:24-36)
    from nuzero_b200.replay import DeviceReplayBuffer
    from nuzero_b200.selfplay import SelfPlayRunner

    # under torchrun every rank's records are all-gathered (NCCL) and the window is sharded over the ranks: each rank ingests
    # and mirrors the games (uid + source rank) % world == rank of the union, so no rank decodes more than it plays
    ingest = True
    drb = DeviceReplayBuffer(e, window_size=args.window_games, batch_size=2048, capacity=args.window_games * 9 + 9,
                             drop_incomplete=True, host_mirror=ingest)
    runner = SelfPlayRunner(e, net, drb, launches_per_step=args.inner, use_graph=True, rank=rank, world=world, gather_to=None)
    reps = args.reps
    kernels_per_step = 2 * args.inner * reps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def play_step():
        # device-resident leg: the move records are written (arena + index) and dropped at the end of the step with one
        # 16-byte memset on the same stream; the e2e leg below is the one that consumes them
        for _ in range(reps):
            runner.play()
        e.arena_top.zero_()

    for _ in range(max(args.warmup, 3)):
        play_step()
    runner._reset_arena()
    barrier()

    # ---- device-resident leg: K steps of `reps` graph replays, CUDA events --------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    c0 = e.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        play_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    c1 = e.counters()
    e.raise_on_error()
    d = {k: c1[k] - c0[k] for k in c1}
    runner.collect()

    # ---- the dominant kernel INSIDE the step: a second graph holds `inner` launches of the network stand-in alone (it is
    # idempotent on an unchanged leaf tensor); search time per launch = (graph of pairs - graph of stand-ins) / inner, which
    # includes the search launch's own scheduling gap and cannot exceed the step's time per pair
    stub_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(stub_graph):
        for _ in range(args.inner):
            net()
    for _ in range(3):
        stub_graph.replay()
    n_k = 8
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    c2 = e.counters()
    torch.cuda.synchronize(dev)
    evs[0].record()
    for _ in range(n_k):
        runner.play()
    evs[1].record()
    for _ in range(n_k):
        stub_graph.replay()
    evs[2].record()
    torch.cuda.synchronize(dev)
    c3 = e.counters()
    dk = {k: c3[k] - c2[k] for k in c3}
    launches_k = n_k * args.inner
    pair_us = evs[0].elapsed_time(evs[1]) * 1000.0 / launches_k
    stub_us = evs[1].elapsed_time(evs[2]) * 1000.0 / launches_k
    k_avg_s = max(pair_us - stub_us, 1e-3) * 1e-6
    counted = algorithmic_bytes(dk, 2, e.A, 18) / launches_k
    runner.collect()

    # nvidia-smi polling visibly perturbs the host-driven e2e leg: the clocks are sampled over the device-resident leg and
    # the kernel-only leg, which run at the same load
    clocks = sampler.stop() if rank == 0 else None
    # ---- end-to-end leg: the same steps through the public API with HOST buffers on both sides.  Per step: the salts of the
    # network stand-in go host -> device from pinned memory (the step's input), `reps` graph replays play, the finished games'
    # records are grouped, decoded to float32 training tuples on the device and copied device -> pinned host memory (the
    # replay window the reference keeps as a host list), and a batch of value targets is read from that host window.
    def host_readback():
        if ingest and drb.len() > 0:
            drb.host_sync()
            n = drb.len()
            return float(drb.h_value.numpy()[drb.rows.logical_rows(max(0, n - 256), n)].sum())
        return 0.0

    def e2e_step():
        net.salt.copy_(salts_host, non_blocking=True)
        drb.h2d_bytes += salts_host.numel() * 4
        added = 0
        for _ in range(reps):
            added += runner.step()
        return added

    for _ in range(max(args.warmup, 3)):  # the W untimed warm-up steps of THIS leg: same calls as the timed ones
        e2e_step()
        host_readback()
    runner.flush()
    barrier()
    c4 = e.counters()
    runner.d2h_bytes, drb.h2d_bytes, drb.d2h_bytes = 0, 0, 0
    pos0 = drb.positions_in
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(args.steps):
        e2e_step()
        checksum += host_readback()
    runner.flush()
    drb.host_sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    c5 = e.counters()
    e.raise_on_error()
    de = {k: c5[k] - c4[k] for k in c5}
    d2h = runner.d2h_bytes + drb.d2h_bytes
    h2d = drb.h2d_bytes
    positions = drb.positions_in - pos0

    dropped = int(e.arena_top[1])
    t = torch.tensor([ms / 1000.0, e2e_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([d["sims"], de["sims"], d["moves"], d2h if ingest else 0, dropped, d["games"], h2d, positions], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        sims_per_launch = dk["sims"] / launches_k
        achieved = SURVEY_BYTES_PER_SIM_TTT * sims_per_launch / k_avg_s / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "advance_traffic.json"))).get("bytes_per_launch")
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": float(tot[0]) / float(t[0]), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": float(t[0]) * 1000.0 / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "timed_seconds": float(t[0]),
            "e2e": {"value": float(tot[1]) / float(t[1]), "unit": UNIT, "h2d_bytes_per_step": float(tot[6]) / args.steps,
                    "d2h_bytes_per_step": float(tot[3]) / args.steps, "seconds": float(t[1]),
                    "records_dropped": int(tot[4]), "positions_into_host_replay_window_per_step": float(tot[7]) / args.steps,
                    "api": "SelfPlayRunner.step() -> DeviceReplayBuffer(host_mirror=True): (state, (value, policy), game_index) rows "
                           "of Training/ReplayBuffer.py:24-36 in pinned host memory (window of %d games)" % args.window_games,
                    "note": "per step the host uploads the network stand-in's per-slot salts from pinned memory (the step's "
                            "input) and receives every finished game's float32 state planes, policy targets, value targets and "
                            "game indices in pinned host memory; moves are grouped into games and decoded on the device; the "
                            "replay-side work of step i overlaps the search of step i+1; the final flush and the wait for the "
                            "last device->host copy are inside the timed region"},
            "gpu_launches": kernels_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": "advance_kernel<TTT>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "avg_launch_us": k_avg_s * 1e6, "pair_us_in_graph": pair_us, "network_standin_us_in_graph": stub_us,
                         "algorithmic_bytes_per_sim": SURVEY_BYTES_PER_SIM_TTT, "sims_per_launch": sims_per_launch,
                         "algorithmic_bytes_per_launch": SURVEY_BYTES_PER_SIM_TTT * sims_per_launch,
                         "counted_bytes_per_launch": counted,
                         "method": "achieved = SURVEY 8(d) bytes per simulation x simulations per launch / in-graph launch time; "
                                   "launch time = (CUDA-graph replay of search+stand-in pairs - replay of the stand-ins alone) / "
                                   "launches, CUDA events on the launching stream; counted_bytes_per_launch evaluates DESIGN 3's "
                                   "per-item bytes on the engine's own work counters (no per-launch term)"},
            "moves_per_sec": float(tot[2]) / float(t[0]),
            "games_per_sec": float(tot[5]) / float(t[0]),
            "work": {"sims": d["sims"], "levels": d["levels"], "children_scanned": d["scanned"],
                     "expansions": d["expansions"], "children_created": d["created"], "moves": d["moves"],
                     "terminal_leaves": d["terminal_leaves"]},
        }
    # free the headline workload's 17 GB before the secondary one allocates
    del runner, drb, stub_graph, net
    e.close()
    del e
    torch.cuda.empty_cache()
    if not args.no_secondary:
        sec = scs_secondary(args, dev, rank, world)
        if rank == 0:
            out["secondary"] = sec
    if rank == 0:
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args.sims, args.cpu_seconds)
            if not args.no_secondary:
                out["secondary"]["cpu_baseline"] = cpu_baseline_scs(args.scs_config, args.scs_sims, args.filters, args.iters, args.cpu_seconds,
                                                                    mean_moves_per_game=out["secondary"].get("mean_moves_per_game"))
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _cpu_worker_ttt_net(args):
    sims, filters, iters, seconds, seed, core = args
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    import numpy as np
    import torch

    from nuzero_b200.nets import RecurrentNet, initialize_parameters
    from oracle import selfplay
    from oracle.ttt import TicTacToe

    torch.set_num_threads(1)
    np.random.seed(seed)
    torch.manual_seed(0)
    model = RecurrentNet(2, 1, filters, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=False)
    initialize_parameters(model)
    model.eval()

    def net(state):  # Network_Manager.inference (Network_Manager.py:46-64): batch 1, fp32, CPU
        with torch.no_grad():
            (p, v), _ = model(torch.from_numpy(np.asarray(state, dtype=np.float32)).reshape(1, 2, 3, 3), iters)
        return p.reshape(-1).numpy(), float(v.reshape(-1)[0])

    cfg = load_cfg(sims)
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        rec = selfplay.play_game(TicTacToe(), net, cfg, True, False, keep_states=False)
        done += rec["length"] * sims
    return done, 0, time.perf_counter() - t0


def run_gpu_ttt_net(args):
    """BASELINE.json configs[0] (the reference's own CPU-runnable case): Tic-Tac-Toe with a real
    recurrent network — 64 filters, orthogonal 3x3 convolutions, 2 iterations (the README's intent for preset 0; the shipped
    best_ttt_config model has these shapes) — 100 simulations per move as in that model's search config."""
    import torch

    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.fastnet import FusedRecurrentForward
    from nuzero_b200.nets import RecurrentNet, initialize_parameters

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    sims, G, filters, iters = 100, args.games, 64, 2
    cfg = load_cfg(sims)
    e = SearchEngine(tic_tac_toe_spec(), cfg, G, True, device=dev, pool_nodes=8192, policy_is_prob=False, leaf_dtype=_ffi.BF16,
                     policy_dtype=_ffi.BF16, auto_advance=True, games_per_slot=0, max_sims_per_launch=8 if args.cache else 1, seed=5,
                     arena_words=1 << 25)
    torch.manual_seed(0)
    model = RecurrentNet(2, 1, filters, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=False)
    initialize_parameters(model)
    cache = None
    if args.cache:
        from nuzero_b200.cache import CachedForward

        net = cache = CachedForward(e, lambda view: FusedRecurrentForward(view, model, iters, use_graph=True), capacity_log2=16,
                                    min_rows=128, in_kernel=True, miss_target=2048)
    else:
        net = FusedRecurrentForward(e, model, iters, use_graph=True)
    for i in range(3000):
        e.advance()
        net()
        if i % 1024 == 1023:
            e.arena_top.zero_()
    torch.cuda.synchronize(dev)
    e.raise_on_error()
    e.arena_top.zero_()
    c0 = e.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_launch = args.steps * 64
    ev0.record()
    for _ in range(n_launch):
        e.advance()
        net()
    ev1.record()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1)
    c1 = e.counters()
    e.raise_on_error()
    d = {k: c1[k] - c0[k] for k in c1}
    out = {"metric": METRIC, "value": d["sims"] / (ms / 1000.0), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": 3,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16 network / f64 search", "data": "synthetic",
           "config": {"workload": "tic_tac_toe_selfplay_100sims_%dgames_recurrentnet64_x2" % G, "network": "RecurrentNet(2, 1, 64 filters, "
                      "2 blocks, recall, orthogonal 3x3) x 2 iterations, random init, fused tcgen05 forward under a CUDA graph",
                      "launch_pairs_per_step": 64},
           "games_per_sec": d["games"] / (ms / 1000.0), "moves_per_sec": d["moves"] / (ms / 1000.0), "gpu_launches": n_launch * 2, "work": d}
    if cache is not None:
        out["config"]["inference_cache"] = "device table, 2^16 slots, exact keys, consulted inside the search kernel (nuzero_b200.cache.CachedForward(in_kernel=True)), up to 8 simulations per game and launch"
        out["cache_hit_rate"] = cache.hit_rate()
    if not args.no_cpu:
        procs = os.cpu_count() or 1
        ctx = mp.get_context("fork")
        with ctx.Pool(procs) as pool:
            res = pool.map(_cpu_worker_ttt_net, [(sims, filters, iters, args.cpu_seconds, 1000 * (i + 1), i % procs) for i in range(procs)])
        out["cpu_baseline"] = {"value": sum(r[0] for r in res) / max(r[2] for r in res), "unit": UNIT, "cores": procs, "kind": "port",
                               "sample": "%d processes x whole TTT self-play games (oracle Explorer port, the same network fp32 batch 1 on one "
                                         "CPU thread, 100 sims/move) for %.0f s" % (procs, args.cpu_seconds)}
    print(json.dumps(out))


def _scs_flops_per_leaf(args, scn):
    """7-tap hex-conv FLOPs of one RecurrentNet forward on one leaf (projection, `iters` recurrent passes of the recall
    convolution + two residual blocks, the policy and value heads)."""
    cells = scn.rows * scn.cols
    F_, Cin, P_ = args.filters, scn.C, scn.planes
    conv = lambda ci, co: 2 * 7 * ci * co  # FLOPs per cell
    per_cell = conv(Cin, F_) + args.iters * (conv(F_ + Cin, F_) + 4 * conv(F_, F_))
    mid_p = int(F_ + (P_ - F_) / 2)
    per_cell += conv(F_, mid_p) + conv(mid_p, P_)
    w = [F_ + (1 - F_) * k / 4 for k in range(5)]
    per_cell += sum(conv(int(w[k]), int(w[k + 1])) for k in range(4))
    return per_cell * cells


def scs_measure(args, dev, rank, world, cache, steady, full):
    """SCS self-play (BASELINE.json configs[2] / [4]): `--scs-games` concurrent games per GPU, `--scs-sims` simulations per
    move, RecurrentNet(C, planes, `--filters`, 2 blocks, recall, hex) x `--iters` in bf16 (hand-written tcgen05 kernels under a
    CUDA graph), optionally behind the device inference cache.
      steady: games restart when they end; K steps of `--scs-inner` (search, network) launch pairs timed with CUDA events,
              then the two launches timed separately -> simulations/s and both rooflines;
      full:   ONE GENERATION through the public API — a fresh batch of games played to the end by SelfPlayRunner.step(),
              every rank's finished trajectories all-gathered (NCCL) and decoded into the replay window, which is sharded over
              the ranks and mirrored to pinned HOST memory (the reference's sink, Training/ReplayBuffer.py:24-36) -> games/s.
    Returns a dict (whole-job figures on every rank)."""
    import torch
    import torch.distributed as dist

    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine
    from nuzero_b200.fastnet import FastRecurrentForward, FusedRecurrentForward
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.nets import RecurrentNet, initialize_parameters
    from nuzero_b200.network import GraphedForward

    cfg = load_cfg(args.scs_sims)
    seeds = list(range(1, 65)) if "randomized" in args.scs_config else [None]
    scn = ScsScenario(os.path.join(ROOT, "nuzero_b200", "configs", "scs", args.scs_config), seeds)
    G = args.scs_games
    torch.manual_seed(0)
    model = RecurrentNet(scn.C, scn.planes, args.filters, 2, recall=True, policy_head="conv", value_head="reduce",
                         value_activation="relu", hex=True)
    initialize_parameters(model)
    net_cls = {"module": GraphedForward, "fast": FastRecurrentForward, "fused": FusedRecurrentForward}[args.net_path]
    flops_leaf = _scs_flops_per_leaf(args, scn)
    cells = scn.rows * scn.cols

    def make(games_per_slot, seed, vl):
        in_kernel = cache and args.scs_cache_budget > 0
        e = SearchEngine(scn.spec(), cfg, G, True, device=dev, pool_nodes=args.scs_pool, policy_is_prob=False,
                         leaf_dtype=_ffi.BF16, policy_dtype=_ffi.BF16, auto_advance=True, games_per_slot=games_per_slot,
                         max_sims_per_launch=args.scs_cache_budget if in_kernel else args.budget, seed=seed + rank, arena_words=1 << 24,
                         max_depth=256, max_levels_per_launch=args.scs_levels, virtual_loss=1 if in_kernel else vl)
        e.set_maps([i % len(seeds) for i in range(G)])
        e.reset()
        if cache:
            from nuzero_b200.cache import CachedForward

            # in_kernel: the search kernel consults the table itself and runs up to --scs-cache-budget simulations per game
            # and launch; only the missed leaves go to the network, as one dense batch
            # table size: 2^22 entries for the 5x5 boards (A = 525: 4.4 GB of policy rows), fewer where the action space is
            # large (30x30, S = 2: A = 18 900) so that the stored policy rows stay below ~8 GB
            cap_log2 = max(12, min(22, int(math.log2(8e9 / (e.A * 2)))))
            net = CachedForward(e, lambda view: net_cls(view, model, args.iters, use_graph=True), capacity_log2=cap_log2,
                                min_rows=args.scs_min_rows if in_kernel else 512, in_kernel=in_kernel,
                                miss_target=args.scs_miss_target if in_kernel else 0, park_target=args.scs_park_target if in_kernel else 0,
                                pipeline=in_kernel and args.scs_pipeline)
        else:
            net = net_cls(e, model, args.iters, use_graph=True)
        return e, net

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    out = {"inference_cache": ("in the search kernel, up to %d simulations per game and launch" % args.scs_cache_budget
                               if args.scs_cache_budget > 0 else "separate look-up kernel after every search launch") if cache else False}
    kernels_per_pair = 1 + (41 if args.net_path == "fused" else 0) + (2 if cache else 0)
    if steady:
        e, net = make(0, 99, args.virtual_loss)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

        def pair():
            e.advance()
            net()

        for _ in range(args.scs_presteps):
            pair()
        torch.cuda.synchronize(dev)
        e.raise_on_error()
        e.arena_top.zero_()
        for _ in range(max(3, args.warmup)):
            pair()
        torch.cuda.synchronize(dev)
        c0 = e.counters()
        if cache:
            m0 = net.misses
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ev[0].record()
        for _ in range(args.scs_steps * args.scs_inner):
            pair()
        if hasattr(net, "drain"):
            net.drain()  # pipeline: the last launch's network call (its own stream) belongs to the timed region
            torch.cuda.current_stream(dev).wait_stream(net._fwd_stream) if getattr(net, "pipeline", False) else None
        ev[1].record()
        torch.cuda.synchronize(dev)
        ms = ev[0].elapsed_time(ev[1])
        c1 = e.counters()
        e.raise_on_error()
        d = {k: c1[k] - c0[k] for k in c1}
        n_pairs = args.scs_steps * args.scs_inner
        rows_fwd = (net.misses - m0) / n_pairs if cache else e.rows
        t_adv = t_net = 0.0
        n_probe = 20
        for _ in range(n_probe):
            ev[0].record(); e.advance(); ev[1].record(); net(); ev[2].record()
            torch.cuda.synchronize(dev)
            t_adv += ev[0].elapsed_time(ev[1]); t_net += ev[1].elapsed_time(ev[2])
        t_adv, t_net = t_adv / n_probe, t_net / n_probe
        if getattr(net, "pipeline", False):
            t_net = float("nan")  # the network call of a launch runs beside the next launch on its own stream: no serial share
        tt = torch.tensor([ms / 1000.0], dtype=torch.float64, device=dev)
        tot = torch.tensor([d["sims"], d["games"], d["moves"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        # search side (SURVEY.md §8d / DESIGN §3): per simulation `levels` x (children scanned x 28 B) + root header + backup +
        # expansion (children created, policy row, value, leaf link, one encoded leaf row) + path save/restore; and per launch
        # and slot the compact root and leaf state moved between HBM and shared memory (2 x state_words x 4 B)
        sbytes = (d["scanned"] * 28 + d["sims"] * 12 + (d["levels"] + d["sims"]) * 24 + d["created"] * 28 +
                  d["expansions"] * (e.A * 2 + 4 + 16 + scn.C * cells * 2 + 8) + (d["levels"] + d["expansions"]) * 8 +
                  n_pairs * G * 2 * e.state_words * 4 +
                  (d["sims"] + d["expansions"]) * ((e.state_words + 3) & ~3) * 4)  # node_state_cache: one state row read per simulation, one written per expansion
        out["steady"] = {
            "value": float(tot[0]) / float(tt[0]), "unit": UNIT, "seconds": float(tt[0]), "launch_pairs": n_pairs,
            "games_per_sec": float(tot[1]) / float(tt[0]), "moves_per_sec": float(tot[2]) / float(tt[0]),
            "split_us": {"advance_kernel": t_adv * 1000, "network_forward": None if t_net != t_net else t_net * 1000},
            "gpu_launches": n_pairs * kernels_per_pair, "work": d}
        if cache:
            out["steady"]["cache_hit_rate"] = net.hit_rate()
            out["steady"]["forward_rows_per_call"] = rows_fwd
        else:
            ach = flops_leaf * e.rows / (t_net / 1000) / 1e12
            out["roofline"] = {"bound": "tensor", "kernel": "network forward (%s, bf16, CUDA graph of 41 launches)" %
                               {"fused": "hexconv_kernel: tcgen05 gather+GEMM", "fast": "im2col kernel + cuBLAS GEMM", "module": "nn.Module / cuDNN"}[args.net_path],
                               "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "traffic": None,
                               "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                               "algorithmic_flops_per_leaf": flops_leaf, "leaf_rows": e.rows, "forward_us": t_net * 1000}
        sach = sbytes / n_pairs / (t_adv / 1000) / 1e9
        out["roofline_search"] = {"bound": "hbm", "kernel": "advance_kernel<SCS>", "achieved": sach, "peak": hbm, "unit": "GB/s",
                                  "frac": sach / hbm, "avg_launch_us": t_adv * 1000, "algorithmic_bytes_per_launch": sbytes / n_pairs,
                                  "levels_per_sim": d["levels"] / max(1, d["sims"]),
                                  "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"}
        del net
        e.close()
        del e
        torch.cuda.empty_cache()
    if full:
        from nuzero_b200.replay import DeviceReplayBuffer
        from nuzero_b200.selfplay import SelfPlayRunner

        e2, net2 = make(1, 7, 1)
        drb = DeviceReplayBuffer(e2, window_size=G, batch_size=2048, capacity=G * args.scs_positions_per_game, host_mirror=True)
        runner = SelfPlayRunner(e2, net2, drb, launches_per_step=args.scs_inner, use_graph=False, rank=rank, world=world, gather_to=None,
                                collect_every=args.scs_collect_every)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        steps_full, checksum = 0, 0.0
        while True:
            runner.step()
            steps_full += 1
            if steps_full % max(8, args.scs_collect_every) == 0:
                e2.raise_on_error()  # a faulted slot never goes idle
                if steps_full > 200000:
                    raise RuntimeError("the SCS generation did not finish within 200000 steps")
                done = torch.tensor([1 if bool((e2.phases() == _ffi.PHASE_IDLE).all()) else 0], dtype=torch.int32, device=dev)
                if world > 1:  # every step holds an all-gather: the ranks leave the loop together
                    dist.all_reduce(done, op=dist.ReduceOp.MIN)
                if int(done.item()):
                    break
        if hasattr(net2, "drain"):
            net2.drain()
        runner.flush()
        drb.host_sync()
        n = drb.len()
        if n:  # the training side's view: the value targets of the newest positions, from HOST memory
            checksum = float(drb.h_value.numpy()[drb.rows.logical_rows(max(0, n - 256), n)].sum())
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        full_s = time.perf_counter() - t0
        e2.raise_on_error()
        cf = e2.counters()
        agg = torch.tensor([cf["sims"], cf["games"], cf["moves"], drb.len(), drb.h2d_bytes, runner.d2h_bytes + drb.d2h_bytes, drb.rows.n_games],
                           dtype=torch.float64, device=dev)
        tmax = torch.tensor([full_s], dtype=torch.float64, device=dev)
        if world > 1:  # whole-job figures: work summed over the ranks, the slowest rank's time
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        full_s = float(tmax[0])
        out["generation"] = {
            "value": float(agg[0]) / full_s, "unit": UNIT, "games": int(agg[1]), "games_per_sec": float(agg[1]) / full_s,
            "moves_per_sec": float(agg[2]) / full_s, "seconds": full_s, "steps": steps_full,
            "positions_in_host_replay_window": int(agg[3]), "games_in_replay_window": int(agg[6]),
            "h2d_bytes_per_step": float(agg[4]) / steps_full, "d2h_bytes_per_step": float(agg[5]) / steps_full,
            "d2h_bytes_total": float(agg[5]), "gpu_launches": steps_full * args.scs_inner * kernels_per_pair,
            "host_checksum": checksum,
            "api": "SelfPlayRunner.step() -> DeviceReplayBuffer(host_mirror=True): one generation of %d games per GPU played to the "
                   "end, (state, (value, policy), game_index) rows in pinned host memory%s"
                   % (G, "; every rank's records all-gathered over NCCL, the window sharded over the ranks" if world > 1 else "")}
        if cache:
            out["generation"]["cache_hit_rate"] = net2.hit_rate()
        del runner, drb, net2
        e2.close()
        del e2
        torch.cuda.empty_cache()
    return out


def scs_secondary(args, dev, rank, world):
    """The SCS half of BASELINE.json's metric, on the same JSON line under `secondary`: configs[2] (and configs[4] under
    torchrun) without and with the inference cache."""
    off = scs_measure(args, dev, rank, world, cache=False, steady=True, full=not args.scs_skip_uncached_generation)
    on = scs_measure(args, dev, rank, world, cache=True, steady=True, full=True)
    head = on["generation"]
    sec = {"metric": "selfplay_games_per_sec", "unit": "games/s", "value": head["games_per_sec"], "n_gpus": world,
           "higher_is_better": True, "scaling": "weak", "dtype": "bf16 network / f32-f64 search", "data": "synthetic",
           "config": scs_config(args, world), "sims_per_sec": head["value"],
           "headline": "one generation of games played to the end with the inference cache (result-identical to the run without it)",
           "roofline": off.get("roofline"), "roofline_search": off.get("roofline_search"),
           "cache_off": {k: off[k] for k in ("steady", "generation") if k in off},
           "cache_on": {k: on[k] for k in ("steady", "generation", "roofline_search") if k in on},
           "e2e": {"value": head["games_per_sec"], "unit": "games/s", "h2d_bytes_per_step": head["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": head["d2h_bytes_per_step"], "seconds": head["seconds"], "api": head["api"]},
           "gpu_launches": head["gpu_launches"],
           "mean_moves_per_game": head["moves_per_sec"] / max(head["games_per_sec"], 1e-9)}
    return sec


def run_gpu_scs(args):
    """`--workload scs5`: the secondary workload alone (one cache setting), for profiling."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = scs_measure(args, dev, rank, world, cache=args.cache, steady=True, full=args.scs_full_games)
    if rank == 0:
        st = m["steady"]
        out = {"metric": METRIC, "value": st["value"], "unit": UNIT, "n_gpus": world, "steps": args.scs_steps, "warmup": max(3, args.warmup),
               "ms_per_step": st["seconds"] * 1000 / args.scs_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "bf16 network / f32-f64 search", "data": "synthetic", "config": scs_config(args, world),
               "games_per_sec": st["games_per_sec"], "moves_per_sec": st["moves_per_sec"], "gpu_launches": st["gpu_launches"]}
        out.update({k: m[k] for k in ("roofline", "roofline_search", "steady", "generation") if k in m})
        if "generation" in m:
            g = m["generation"]
            out["e2e"] = {"value": g["value"], "unit": UNIT, "games_per_sec": g["games_per_sec"], "h2d_bytes_per_step": g["h2d_bytes_per_step"],
                          "d2h_bytes_per_step": g["d2h_bytes_per_step"], "seconds": g["seconds"], "api": g["api"]}
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline_scs(args.scs_config, args.scs_sims, args.filters, args.iters, args.cpu_seconds)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _keep_stdout_for_the_json_line():
    """The contract is ONE line on stdout.  Libraries write there too (NCCL prints its version banner on file descriptor 1
    when NCCL_DEBUG is set in the environment): descriptor 1 is pointed at stderr for the whole run and python's own
    sys.stdout — which only the final print(json.dumps(...)) uses — keeps the real one."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    _keep_stdout_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--games", type=int, default=16384)
    ap.add_argument("--pool", type=int, default=32768)
    ap.add_argument("--inner", type=int, default=256, help="(search launch + net forward) pairs per step")
    ap.add_argument("--budget", type=int, default=1, help="max simulations per game per launch")
    ap.add_argument("--virtual-loss", type=int, default=1, help="leaves one game may have waiting at the network; > 1 is the "
                    "throughput mode (not bit-identical to the reference), use with --budget >= that width")
    ap.add_argument("--presteps", type=int, default=60000, help="untimed launch pairs that de-synchronise the game slots")
    ap.add_argument("--arena-words", type=int, default=1 << 24)
    ap.add_argument("--window-games", type=int, default=400000, help="replay window of the e2e leg, in games")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="ttt", choices=["ttt", "scs5", "ttt_net"])
    ap.add_argument("--cache", action="store_true", help="ttt_net / scs5: serve repeated leaf states from the device inference "
                    "cache (result-identical; the network runs on the missed rows only)")
    ap.add_argument("--scs-config", default="mirrored_config_5.yml")
    ap.add_argument("--scs-games", type=int, default=4096)
    ap.add_argument("--scs-sims", type=int, default=200)
    ap.add_argument("--scs-pool", type=int, default=131072)
    ap.add_argument("--scs-inner", type=int, default=16)
    ap.add_argument("--scs-presteps", type=int, default=300)
    ap.add_argument("--scs-levels", type=int, default=0, help="tree levels per game per launch (0 = unlimited)")
    ap.add_argument("--reps", type=int, default=24, help="ttt: CUDA-graph replays per step (a step = reps x inner launch pairs)")
    ap.add_argument("--no-secondary", action="store_true", help="ttt: skip the SCS workload that the default run reports under `secondary`")
    ap.add_argument("--scs-steps", type=int, default=8, help="scs: timed steps of --scs-inner launch pairs in the steady-state leg")
    ap.add_argument("--scs-positions-per-game", type=int, default=128, help="scs: replay-window rows reserved per game")
    ap.add_argument("--scs-skip-uncached-generation", action="store_true", help="secondary: play the full generation with the cache only")
    ap.add_argument("--scs-collect-every", type=int, default=8, help="scs generation: collect (and all-gather) the finished games' records "
                    "on every k-th step")
    ap.add_argument("--scs-miss-target", type=int, default=512, help="scs with the in-kernel cache: the search launch ends once this many "
                    "leaves wait for the network (0 = only the per-game budget ends it)")
    ap.add_argument("--scs-park-target", type=int, default=0, help="scs with the in-kernel cache: the search launch ends once this many "
                    "games wait for the network, on their own row or a shared one (0 = off)")
    ap.add_argument("--scs-pipeline", action="store_true", help="scs with the in-kernel cache: two-lane pipeline (network call of launch k "
                    "beside the search of launch k + 1).  Measured slower on one B200: the convolution CTAs (200 KB of shared memory) "
                    "cannot become resident beside the search CTAs, so nothing overlaps and every game plays in every other launch")
    ap.add_argument("--scs-min-rows", type=int, default=128, help="scs with the in-kernel cache: smallest prepared network batch")
    ap.add_argument("--scs-cache-budget", type=int, default=8, help="scs with the inference cache: simulations one game may run per "
                    "launch while its leaves hit the cache inside the search kernel (0 = look the cache up with a kernel of its own)")
    ap.add_argument("--scs-full-games", action="store_true", help="scs5: also play one generation of games to the end "
                    "through SelfPlayRunner (games/s, e2e); takes about a minute")
    ap.add_argument("--filters", type=int, default=256)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--net-path", default="fused", choices=["fused", "fast", "module"],
                    help="fused: hand-written tcgen05 gather+GEMM kernel per conv; fast: im2col kernel + cuBLAS GEMM "
                         "per conv; module: the nn.Module (cuDNN convs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "scs5":
        run_gpu_scs(args)
    elif args.workload == "ttt_net":
        run_gpu_ttt_net(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
