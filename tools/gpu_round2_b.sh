#!/bin/bash
# ncu captures for profiles/r02_*: the replay step's kernels (full set, with source) and the environment step's kernels
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-graphs --fast-setup > gpurun_out/r02_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_tail|k_stage0|k_stage1|k_bwd1|wgrad|k_wsplit|adam' -s 60 -c 24 -f -o gpurun_out/r02_staged \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-graphs --fast-setup > gpurun_out/r02_under_ncu2.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_smooth' -s 2 -c 1 -f -o gpurun_out/r02_smooth \
    python tools/smooth_bench.py > gpurun_out/r02_under_ncu3.log 2>&1; echo "ncu smooth rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r02_launches_env.csv \
    python tools/env_profile.py > gpurun_out/r02_under_ncu4.log 2>&1; echo "ncu env list rc=$?"
ls -la gpurun_out | grep r02_
