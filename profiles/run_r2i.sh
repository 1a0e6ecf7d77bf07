#!/bin/bash
# probe: footing at 50^3 with several displacement magnitudes / autoinc, then the 100^3 runs
set -u
mkdir -p gpurun_out
t0=$SECONDS
for uz in -0.01 -0.004 -0.002; do
  timeout 120 python profiles/run_full_configs.py --config 3 --scale 2 --uz $uz --out gpurun_out/footing_probe.jsonl 2>&1 | head -1 | cut -c1-700; echo "probe $uz: $((SECONDS-t0)) s"
done
timeout 200 python profiles/run_full_configs.py --config 3 --scale 2 --uz -0.01 --autoinc --out gpurun_out/footing_probe.jsonl 2>&1 | head -1 | cut -c1-700; echo "probe autoinc: $((SECONDS-t0)) s"
