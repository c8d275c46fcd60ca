#!/bin/bash
# ncu: DRAM bytes and time of msm_accumulate_kernel for the load-width / L2-fetch-granularity variants (profiles/r2_load_width.md)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum
run() { # name, env...
  name=$1; shift
  env "$@" B2G_GRAPH=0 ncu --metrics $M --clock-control none -k regex:msm_accumulate -c 12 --csv --log-file gpurun_out/r2_lw_$name.csv python tools/prof_msm.py 20 1 > gpurun_out/r2_lw_$name.log 2>&1
}
run base X=1
run fetch32 B2G_L2_FETCH=32
run fetch64 B2G_L2_FETCH=64
run fetch128 B2G_L2_FETCH=128
