#!/bin/bash
# final-code ncu captures of the replay step (launch list + full set of the step's kernels)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-graphs --fast-setup > gpurun_out/r02_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_tail|k_stage0|k_stage1|k_bwd1|wgrad|k_wsplit|adam' -s 60 -c 12 -f -o gpurun_out/r02_staged \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-graphs --fast-setup > gpurun_out/r02_under_ncu2.log 2>&1; echo "ncu full rc=$?"
