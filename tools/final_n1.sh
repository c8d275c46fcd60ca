#!/bin/bash
# One-GPU round-end check: full GPU suite, smoke, sanitizer on the NTT kernels, default bench line, ncu launch list.
tag=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log; tail -2 gpurun_out/${tag}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
{ echo "# compute-sanitizer (CUDA 12.9) on B200: racecheck + memcheck over the NTT pass kernels (radix-8 rounds and radix-2), plain and fused chains"
  for tool in racecheck memcheck; do
    echo "## $tool"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_ntt_pass_schedules and 13 or test_witness_map_pass_schedules or (test_ntt_plain and (14 or 5 or 9))" 2>&1 | grep -v "^$" | tail -8
    echo "$tool rc=${PIPESTATUS[0]}"
  done; } > gpurun_out/${tag}_sanitizer.txt 2>&1
tail -4 gpurun_out/${tag}_sanitizer.txt
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc=$?"; cut -c1-400 gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --skip-check --inflight 1 > gpurun_out/${tag}_launches.log 2>&1; echo "ncu rc=$?"
