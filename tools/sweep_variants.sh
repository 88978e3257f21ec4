#!/bin/bash
# runs a quick parity check and the headline bench (device leg, no CPU arm) once per variant library in build_variants/
mkdir -p gpurun_out
for lib in build_variants/libnz_*.so; do
  name=$(basename $lib .so)
  ok=$(NZ_ENGINE_LIB=$PWD/$lib python tools/debug_ttt_auto.py 60 2>&1 | grep -c " ok$")
  NZ_ENGINE_LIB=$PWD/$lib python bench.py --steps 20 --warmup 3 --no-cpu "$@" > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sweep_$name.json"))
    print("$name", "parity_ok_lines $ok/24", "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "launch_us %.2f" % d["roofline"]["avg_launch_us"])
except Exception as ex:
    print("$name", "parity_ok_lines $ok/24", "FAILED", ex)
PY
done
