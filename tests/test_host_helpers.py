"""CPU: small host-side rules that the GPU paths rest on."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_batch_ladders_cover_every_row_count():
    from nuzero_b200.cache import batch_ladder

    for rows, min_rows in ((4096, 128), (4096, 512), (96, 32), (24, 8), (100, 256), (1, 1)):
        for in_kernel in (False, True):
            sizes = batch_ladder(rows, min_rows, in_kernel)
            assert sizes == sorted(set(sizes)) and sizes[-1] == rows and all(0 < n <= rows for n in sizes)
            if in_kernel:  # a batch is never more than 25 % (+ one 64-row step) larger than the rows that need it
                for need in range(max(1, min(min_rows, rows)), rows + 1, max(1, rows // 97)):
                    got = next(n for n in sizes if n >= need)
                    assert got <= max(need * 1.25 + 64, min_rows + 63), (rows, min_rows, need, got)
    assert batch_ladder(4096, 512, True)[:4] == [512, 640, 832, 1024]
    assert batch_ladder(4096, 512, False) == [512, 1024, 4096]


def test_reference_arm_of_the_bench_prints_the_contract_line():
    """`bench.py --impl reference` on the host cores: one JSON line, the driver's keys, the reference itself as the CPU arm
    wherever the tree (or its byte code, oracle/_ref) is present."""
    from oracle import ref_harness as rh

    if not rh.available():
        pytest.skip("no reference tree / oracle/_ref here")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "60", "--warmup", "0",
                          "--no-secondary"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "mcts_sims_per_sec" and d["unit"] == "sims/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0} and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "tic_tac_toe_selfplay_800sims_16384games_stubnet"
