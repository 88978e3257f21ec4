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
# CPU arm: the oracle port (reference Explorer + Gamer loop restated) on host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    sims, seconds, seed, core = args
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    import numpy as np

    from oracle import selfplay
    from oracle.stubnet_np import stub_forward
    from oracle.ttt import TicTacToe

    np.random.seed(seed)
    cfg = load_cfg(sims)
    done_sims, games, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        rec = selfplay.play_game(TicTacToe(), lambda s, sl=seed + games: stub_forward(s, 9, sl), cfg, True, True,
                                 keep_states=True)
        done_sims += rec["length"] * sims
        games += 1
    return done_sims, games, time.perf_counter() - t0


def cpu_baseline(sims, seconds, procs=None):
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(sims, seconds, 1000 * (i + 1), i % (os.cpu_count() or 1)) for i in range(procs)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    span = max(r[2] for r in res)
    return {"value": total / span, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": "%d processes x whole TTT self-play games (%d sims/move, dyadic stub net, training=True) "
                      "for %.0f s each; %d games, %d sims; wall %.1f s"
                      % (procs, sims, seconds, sum(r[1] for r in res), total, wall)}


def _cpu_worker_scs(args):
    config, sims, filters, iters, seconds, seed, core = args
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    import numpy as np
    import torch

    from nuzero_b200.nets import RecurrentNet, initialize_parameters
    from oracle import selfplay
    from oracle.scs import SCS, load_scenario

    torch.set_num_threads(1)
    np.random.seed(seed)
    torch.manual_seed(0)
    sc = load_scenario(os.path.join(ROOT, "nuzero_b200", "configs", "scs", config), seed=None)
    model = RecurrentNet(sc.C, sc.planes, filters, 2, recall=True, policy_head="conv", value_head="reduce",
                         value_activation="relu", hex=True)
    initialize_parameters(model)
    model.eval()
    calls = [0]

    def net(state):  # Network_Manager.inference (Network_Manager.py:46-64): batch 1, fp32, CPU
        calls[0] += 1
        with torch.no_grad():
            (p, v), _ = model(torch.from_numpy(np.asarray(state, dtype=np.float32)).reshape((1,) + tuple(sc_shape)), iters)
        return p.reshape(-1).numpy(), float(v.reshape(-1)[0])

    game = SCS(sc)
    sc_shape = game.state_shape
    cfg = load_cfg(sims)
    from oracle import mcts

    # a bounded sample: simulations of the opening moves of one game (whole games take minutes per core on CPU)
    root = mcts.Node(0)
    cfg1 = {k: dict(v) if isinstance(v, dict) else v for k, v in cfg.items()}
    cfg1["Simulation"]["mcts_simulations"] = 8
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds and not game.is_terminal():
        action, child, _ = mcts.run_mcts(cfg1, game, net, root, True, False, mcts.LiveTape(), 0)
        done += 8
        if root.N >= sims:
            game.step(action)
            root = child
    return done, 0, time.perf_counter() - t0


def cpu_baseline_scs(config, sims, filters, iters, seconds, procs=None):
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker_scs, [(config, sims, filters, iters, seconds, 1000 * (i + 1), i % (os.cpu_count() or 1))
                                         for i in range(procs)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    span = max(r[2] for r in res)
    return {"value": total / span, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": "%d processes x opening-move simulations of one SCS game each (oracle Explorer port, RecurrentNet-%d x%d "
                      "fp32 batch-1 forward on one CPU thread, training=True) for %.0f s; %d sims; wall %.1f s"
                      % (procs, filters, iters, seconds, total, wall)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_baseline(args.sims, min(per_step, 2.0))
    vals, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_baseline(args.sims, per_step)
        vals.append(last["value"])
    dt = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    last["value"] = v
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(1, args.steps), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, 1), "cpu_baseline": last,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


def workload_config(args, world):
    return {"workload": "tic_tac_toe_selfplay_800sims_16384games_stubnet" if (args.sims, args.games) == (800, 16384)
            else "tic_tac_toe_selfplay_%dsims_%dgames_stubnet" % (args.sims, args.games),
            "game": "Tic_Tac_Toe", "sims_per_move": args.sims, "concurrent_games_per_gpu": args.games,
            "network": "deterministic dyadic stub (CUDA kernel)", "training": True, "keep_subtree": True,
            "search_config": "a1_search_config (pb_c_base 10000, pb_c_init 1.15, noise 0.2/0.15)",
            "inner_launch_pairs_per_step": args.inner, "max_sims_per_launch": args.budget,
            "leaves_in_flight_per_game": max(1, args.virtual_loss),
            "l2_policy": "node pools (%.1f GB/GPU) exceed the 126 MB L2; no flush" % (args.games * args.pool * 32 / 1e9),
            "parallelism": "independent game batches per GPU, no collective on the search path (x%d)" % world}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(d, leaf_elem_bytes, A, leaf_elems, G, launches):
    """Minimal HBM traffic of the search data structure for the counted work:
    select reads 28 B per scanned child (prior 8, W 8, N 4, child range + action 8) + the root header
    (12 B); backup reads+writes N and W of every path node (24 B); expand writes 28 B per created
    child, reads the policy row and value, rewrites the leaf's link, saves/restores the path; the
    encoder writes one leaf row; every launch reads and writes each slot's 128-byte control block."""
    sims, levels, scanned = d["sims"], d["levels"], d["scanned"]
    exp, created = d["expansions"], d["created"]
    b = scanned * 28 + sims * 12
    b += (levels + sims) * 24
    b += created * 28 + exp * (A * 4 + 4 + 16 + leaf_elems * leaf_elem_bytes + 8)
    b += (levels + exp) * 8  # path save + restore for simulations that wait on the network
    b += launches * G * 256
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
    net = DyadicStubNet(e, uid_mul=1 if args.virtual_loss <= 1 else 0)

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
    # finished games' records are decoded on the device into the replay window (DeviceReplayBuffer)
    from nuzero_b200.replay import DeviceReplayBuffer
    from nuzero_b200.selfplay import SelfPlayRunner

    drb = DeviceReplayBuffer(e, window_size=args.window_games, batch_size=2048, capacity=args.window_games * 9 + 9,
                             drop_incomplete=True)
    runner = SelfPlayRunner(e, net, drb, launches_per_step=args.inner, use_graph=True, rank=rank, world=world)
    kernels_per_step = 2 * args.inner

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        runner.play()
    runner.collect()
    barrier()

    # ---- device-resident leg: K graph replays, CUDA events ----------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    c0 = e.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        runner.play()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    c1 = e.counters()
    e.raise_on_error()
    d = {k: c1[k] - c0[k] for k in c1}
    runner.collect()

    # ---- dominant kernel alone: CUDA events around search launches only ------------------------
    n_k = 200
    k_ms = []
    c2 = e.counters()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_k)]
    for a, b in evs:
        a.record()
        e.advance()
        b.record()
        net()
    torch.cuda.synchronize(dev)
    k_ms = [a.elapsed_time(b) for a, b in evs]
    c3 = e.counters()
    dk = {k: c3[k] - c2[k] for k in c3}
    kbytes = algorithmic_bytes(dk, 2, e.A, 18, G, n_k)
    k_avg_s = sum(k_ms) / len(k_ms) / 1000.0
    runner.collect()

    # nvidia-smi polling visibly perturbs the host-driven e2e leg (4.4e8 vs 6.0e8 sims/s with / without it): the clocks are
    # sampled over the device-resident leg and the kernel-only leg, which run at the same load
    clocks = sampler.stop() if rank == 0 else None
    # ---- end-to-end leg through the public API: every step = graph replay + records -> replay window.  Per step the host
    # reads the record headers (D2H), groups moves into games, uploads the row assignment (H2D) and nz_replay_decode writes
    # float32 planes + policy rows into the device-resident window; a sample batch is read back at the end of every step.
    def sample_readback():
        if (rank == 0 or world == 1) and drb.len() > 0:
            with torch.cuda.stream(runner.side):
                st_b, v_b, p_b, _g = drb.get_sample_tensors(256, True)
                return float(v_b.sum().cpu())  # D2H read of a training batch's value targets
        return 0.0

    for _ in range(max(args.warmup, 3)):  # the W untimed warm-up steps of THIS leg: same calls as the timed ones
        runner.step()
        sample_readback()
    runner.flush()
    barrier()
    c4 = e.counters()
    runner.d2h_bytes, drb.h2d_bytes, drb.d2h_bytes = 0, 0, 0
    pos0 = drb.positions_in
    t0 = time.perf_counter()
    checksum = 0.0
    step_ts = []
    for _ in range(args.steps):
        runner.step()
        checksum += sample_readback()
        if rank == 0 or world == 1:
            runner.d2h_bytes += 4
        step_ts.append(time.perf_counter() - t0)
    runner.flush()
    step_ts.append(time.perf_counter() - t0)
    barrier()
    e2e_s = time.perf_counter() - t0
    if os.environ.get("NZ_BENCH_TRACE") and rank == 0:
        print("e2e host timeline (ms, last = after flush):", " ".join("%.1f" % (1e3 * x) for x in step_ts), file=sys.stderr)
    c5 = e.counters()
    e.raise_on_error()
    de = {k: c5[k] - c4[k] for k in c5}
    d2h = runner.d2h_bytes + drb.d2h_bytes
    h2d = drb.h2d_bytes
    positions = drb.positions_in - pos0

    dropped = int(e.arena_top[1])
    t = torch.tensor([ms / 1000.0, e2e_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([d["sims"], de["sims"], d["moves"], d2h if rank == 0 or world == 1 else 0, dropped, d["games"]],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = kbytes / n_k / k_avg_s / 1e9
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
            "clocks": clocks,
            "e2e": {"value": float(tot[1]) / float(t[1]), "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": float(tot[3]) / args.steps,
                    "records_dropped": int(tot[4]), "positions_into_replay_window_per_step": positions / args.steps,
                    "api": "nuzero_b200.selfplay.SelfPlayRunner.step() -> DeviceReplayBuffer (window of %d games)" % args.window_games,
                    "note": "self-play has no per-step host input tensor (H2D is launch arguments only): per step the host "
                            "reads the arena counters and one row per finished game (positions, validity, result); moves "
                            "are grouped into games, ordered and decoded to float32 training tuples on the device and stay "
                            "in HBM; one sampled value-target batch is read back per step.  The replay-side work of step i "
                            "runs on a side stream while the search of step i+1 runs; the final flush is inside the timed "
                            "region"},
            "gpu_launches": kernels_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": "advance_kernel<TTT>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "avg_launch_us": k_avg_s * 1e6, "algorithmic_bytes_per_launch": kbytes / n_k,
                         "sims_per_launch": dk["sims"] / n_k},
            "moves_per_sec": float(tot[2]) / float(t[0]),
            "games_per_sec": float(tot[5]) / float(t[0]),
            "work": {"sims": d["sims"], "levels": d["levels"], "children_scanned": d["scanned"],
                     "expansions": d["expansions"], "children_created": d["created"], "moves": d["moves"],
                     "terminal_leaves": d["terminal_leaves"]},
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args.sims, args.cpu_seconds)
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
                     policy_dtype=_ffi.BF16, auto_advance=True, games_per_slot=0, max_sims_per_launch=1, seed=5, arena_words=1 << 25)
    torch.manual_seed(0)
    model = RecurrentNet(2, 1, filters, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=False)
    initialize_parameters(model)
    cache = None
    if args.cache:
        from nuzero_b200.cache import CachedForward

        net = cache = CachedForward(e, lambda view: FusedRecurrentForward(view, model, iters, use_graph=True), capacity_log2=16)
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
        out["config"]["inference_cache"] = "device table, 2^16 slots, exact keys (nuzero_b200.cache.CachedForward)"
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


def run_gpu_scs(args):
    """Secondary workload (BASELINE.json configs[2]): SCS 5x5 self-play, 200 sims/move, 4096 concurrent
    games, RecurrentNet(86, 21, 256 filters, 2 blocks, recall, hex) x 6 iterations in bf16 under a CUDA graph."""
    import torch
    import torch.distributed as dist

    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.nets import RecurrentNet, initialize_parameters
    from nuzero_b200.fastnet import FastRecurrentForward, FusedRecurrentForward
    from nuzero_b200.network import GraphedForward

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = load_cfg(args.scs_sims)
    seeds = list(range(1, 65)) if "randomized" in args.scs_config else [None]
    scn = ScsScenario(os.path.join(ROOT, "nuzero_b200", "configs", "scs", args.scs_config), seeds)
    G = args.scs_games
    e = SearchEngine(scn.spec(), cfg, G, True, device=dev, pool_nodes=args.scs_pool, policy_is_prob=False,
                     leaf_dtype=_ffi.BF16, policy_dtype=_ffi.BF16, auto_advance=True, games_per_slot=0,
                     max_sims_per_launch=args.budget, seed=99 + rank, arena_words=1 << 24, max_depth=256,
                     max_levels_per_launch=args.scs_levels, virtual_loss=args.virtual_loss)
    e.set_maps([i % len(seeds) for i in range(G)])
    e.reset()
    torch.manual_seed(0)
    model = RecurrentNet(scn.C, scn.planes, args.filters, 2, recall=True, policy_head="conv", value_head="reduce",
                         value_activation="relu", hex=True)
    initialize_parameters(model)
    net_cls = {"module": GraphedForward, "fast": FastRecurrentForward, "fused": FusedRecurrentForward}[args.net_path]
    if args.cache:
        from nuzero_b200.cache import CachedForward

        net = CachedForward(e, lambda view: net_cls(view, model, args.iters, use_graph=True), capacity_log2=22, min_rows=512)
    else:
        net = net_cls(e, model, args.iters, use_graph=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

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
    steps = args.steps
    t_adv = t_net = 0.0
    ev[0].record()
    for _ in range(steps):
        for _ in range(args.scs_inner):
            pair()
    ev[1].record()
    torch.cuda.synchronize(dev)
    ms = ev[0].elapsed_time(ev[1])
    c1 = e.counters()
    e.raise_on_error()
    d = {k: c1[k] - c0[k] for k in c1}
    n_probe = 20
    for _ in range(n_probe):
        ev[0].record(); e.advance(); ev[1].record(); net(); ev[2].record()
        torch.cuda.synchronize(dev)
        t_adv += ev[0].elapsed_time(ev[1]); t_net += ev[1].elapsed_time(ev[2])
    t_adv, t_net = t_adv / n_probe, t_net / n_probe
    cells = scn.rows * scn.cols
    F_, Cin, P_ = args.filters, scn.C, scn.planes
    conv = lambda ci, co: 2 * 7 * ci * co  # 7-tap hex conv, FLOPs per cell
    per_cell = conv(Cin, F_) + args.iters * (conv(F_ + Cin, F_) + 4 * conv(F_, F_))
    mid_p = int(F_ + (P_ - F_) / 2)
    per_cell += conv(F_, mid_p) + conv(mid_p, P_)
    w = [F_ + (1 - F_) * k / 4 for k in range(5)]
    per_cell += sum(conv(int(w[k]), int(w[k + 1])) for k in range(4))
    flops = per_cell * cells * e.rows  # the forward runs on every leaf row (G x leaves in flight per game)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    # ---- end-to-end leg through the public API: a fresh batch of G games played to the end (one generation),
    # SelfPlayRunner.step() = CUDA-graph-free launch pairs + pipelined record collection into a DeviceReplayBuffer
    e2e = None
    if args.scs_full_games:
        from nuzero_b200.replay import DeviceReplayBuffer
        from nuzero_b200.selfplay import SelfPlayRunner

        e2 = SearchEngine(scn.spec(), cfg, G, True, device=dev, pool_nodes=args.scs_pool, policy_is_prob=False,
                          leaf_dtype=_ffi.BF16, policy_dtype=_ffi.BF16, auto_advance=True, games_per_slot=1,
                          max_sims_per_launch=args.budget, seed=7 + rank, arena_words=1 << 24, max_depth=256,
                          max_levels_per_launch=args.scs_levels)
        e2.set_maps([i % len(seeds) for i in range(G)])
        e2.reset()
        if args.cache:
            net2 = CachedForward(e2, lambda view: net_cls(view, model, args.iters, use_graph=True), capacity_log2=22, min_rows=512)
        else:
            net2 = net_cls(e2, model, args.iters, use_graph=True)
        drb = DeviceReplayBuffer(e2, window_size=G, batch_size=2048, capacity=G * 160)
        runner = SelfPlayRunner(e2, net2, drb, launches_per_step=args.scs_inner, use_graph=False, rank=rank, world=world)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        steps_full = 0
        while True:
            runner.step()
            steps_full += 1
            if steps_full % 8 == 0:
                done = torch.tensor([1 if bool((e2.phases() == _ffi.PHASE_IDLE).all()) else 0], dtype=torch.int32, device=dev)
                if world > 1:  # every step holds an all-gather: the ranks leave the loop together
                    dist.all_reduce(done, op=dist.ReduceOp.MIN)
                if int(done.item()):
                    break
        runner.flush()
        torch.cuda.synchronize(dev)
        full_s = time.perf_counter() - t0
        e2.raise_on_error()
        cf = e2.counters()
        if world > 1:  # whole-job figures: work summed over the ranks, the slowest rank's time
            agg = torch.tensor([cf["sims"], cf["games"], cf["moves"]], dtype=torch.float64, device=dev)
            tmax = torch.tensor([full_s], dtype=torch.float64, device=dev)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            cf = dict(cf, sims=float(agg[0]), games=float(agg[1]), moves=float(agg[2]))
            full_s = float(tmax[0])
        e2e = {"value": cf["sims"] / full_s, "unit": UNIT, "h2d_bytes_per_step": drb.h2d_bytes / steps_full,
               "d2h_bytes_per_step": (runner.d2h_bytes + drb.d2h_bytes) / steps_full, "games": cf["games"], "games_per_sec": cf["games"] / full_s,
               "moves_per_sec": cf["moves"] / full_s, "positions_in_replay_window": drb.len(), "seconds": full_s,
               "api": "SelfPlayRunner.step() -> DeviceReplayBuffer, one generation of %d games per GPU played to the end%s"
                      % (G, "; all ranks' records all-gathered into rank 0's window" if world > 1 else "")}
        if args.cache:
            e2e["cache_hit_rate"] = net2.hit_rate()
    tt = torch.tensor([ms / 1000.0], dtype=torch.float64, device=dev)
    tot = torch.tensor([d["sims"], d["games"], d["moves"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        # search side: per simulation d levels x (K children x 28 B + header) + backup + expansion + encode
        sbytes = (d["scanned"] * 28 + d["sims"] * 12 + (d["levels"] + d["sims"]) * 24 + d["created"] * 28 +
                  d["expansions"] * (e.A * 2 + 4 + 16 + scn.C * cells * 2 + 8) + (d["levels"] + d["expansions"]) * 8 +
                  steps * args.scs_inner * G * (256 + 2 * e.state_words * 4))
        out = {
            "metric": METRIC, "value": float(tot[0]) / float(tt[0]), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(3, args.warmup), "ms_per_step": float(tt[0]) * 1000 / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 network / f32-f64 search", "data": "synthetic",
            "config": {"workload": "scs_%s_%dsims_%dgames_recurrentnet%d_x%d" % (args.scs_config.replace(".yml", ""), args.scs_sims, G, args.filters, args.iters),
                       "inner_launch_pairs_per_step": args.scs_inner, "max_sims_per_launch": args.budget,
                       "leaves_in_flight_per_game": max(1, args.virtual_loss),
                       "l2_policy": "activations of one forward (52 MB per layer) and the node pools exceed L2 across a step; no flush"},
            "games_per_sec": float(tot[1]) / float(tt[0]), "moves_per_sec": float(tot[2]) / float(tt[0]),
            "gpu_launches": steps * args.scs_inner * (1 + (41 if args.net_path == "fused" else 0)),
            "split_us": {"advance_kernel": t_adv * 1000, "network_forward": t_net * 1000},
            "roofline": {"bound": "tensor", "kernel": "network forward (%s, bf16, CUDA graph)" % {"fused": "tcgen05 gather+GEMM kernel", "fast": "im2col kernel + cuBLAS GEMM", "module": "nn.Module / cuDNN"}[args.net_path],
                         "achieved": flops / (t_net / 1000) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                         "frac": flops / (t_net / 1000) / 1e12 / tpeak, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                         "algorithmic_flops_per_leaf": per_cell * cells},
            "roofline_search": {"bound": "hbm", "kernel": "advance_kernel<SCS>", "achieved": sbytes / (steps * args.scs_inner) / (t_adv / 1000) / 1e9,
                                "peak": hbm, "unit": "GB/s", "frac": sbytes / (steps * args.scs_inner) / (t_adv / 1000) / 1e9 / hbm,
                                "avg_launch_us": t_adv * 1000, "levels_per_sim": d["levels"] / max(1, d["sims"])},
            "work": d}
        if args.cache:
            out["config"]["inference_cache"] = "device table, 2^22 slots, exact keys (nuzero_b200.cache.CachedForward)"
            out["cache_hit_rate"] = net.hit_rate()
        if e2e is not None:
            out["e2e"] = e2e
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
