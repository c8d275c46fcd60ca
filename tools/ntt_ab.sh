#!/bin/bash
# A/B of the NTT pass kernels on one B200: parity subset, then whole-proof bench lines per variant, then ncu pipe metrics.
# Usage (GPU box): bash tools/ntt_ab.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt or witness_map or synthetic_proof or libsnark" > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu --skip-check > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.log; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${tag}_${name}.json'))
    print('${name}', 'e2e', round(d['e2e']['value'], 2), 'value', round(d['value'], 2), 'witness_map_ms', round(d['phase_ms_one_proof_alone']['witness_map'], 3))
except Exception as e:
    print('${name}', 'failed', e)
PY
}
run radix2 B2G_NTT_RADIX2=1
run radix8 B2G_X=0
run radix8_tl9 B2G_NTT_TL=9
run radix8_tl11 B2G_NTT_TL=11
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
  --clock-control none -k regex:ntt_pass -c 9 --csv --log-file gpurun_out/${tag}_ncu_ntt.csv python bench.py --steps 1 --warmup 1 --no-cpu --skip-check --inflight 1 > gpurun_out/${tag}_ncu_ntt.log 2>&1
echo "ncu rc=$?"
