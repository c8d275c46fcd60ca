#!/bin/bash
# One gpurun call that answers the round-2 question (≈ 5 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash experiments/dfma/run_round2.sh'
# 1. probe: parity of the FP64 field / curve code against the integer code, and the pipe-overlap measurement
# 2. the product test-suite and bench.py on the library built with the FP64 accumulation path, for a few run-split ratios
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
make -C experiments/dfma > gpurun_out/dfma_build.log 2>&1 || { tail -5 gpurun_out/dfma_build.log; exit 1; }
timeout 120 experiments/dfma/dfma_probe 20 2000 | tee gpurun_out/dfma_probe.jsonl
make -C circom_compat_b200/csrc fp64 >> gpurun_out/dfma_build.log 2>&1 || { tail -5 gpurun_out/dfma_build.log; exit 1; }
export B2G_LIB=$PWD/circom_compat_b200/libb2groth_fp64.so
B2G_MSM_FP64_SHARE=50 timeout 400 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
for share in 0 30 45 60; do
  B2G_MSM_FP64_SHARE=$share timeout 150 python bench.py --no-cpu --steps 10 > gpurun_out/dfma_bench_$share.json 2> gpurun_out/dfma_bench_$share.log
  python - <<PY
import json
d = json.load(open("gpurun_out/dfma_bench_$share.json"))
print("share $share: value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "G1 acc ms", round(d["roofline"]["kernel_ms"], 3), "latency", round(d["single_proof_latency_ms"], 2), d["clocks"])
PY
done
