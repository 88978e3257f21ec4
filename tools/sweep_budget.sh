#!/bin/bash
mkdir -p gpurun_out
for b in 1 2 3 4 8; do
  python bench.py --steps 20 --warmup 3 --no-cpu --budget $b > gpurun_out/budget_$b.json 2> gpurun_out/budget_$b.err
  python -c "
import json
d=json.load(open('gpurun_out/budget_$b.json'))
print('budget $b value %.4g ms/step %.3f e2e %.4g launch_us %.2f sims/launch %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['sims_per_launch']))"
done
