#!/bin/bash
# one-box validation + profile capture for round 2 (run through gpurun from the repo root)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02_bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-graphs --fast-setup > gpurun_out/r02_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tail|k_stage0|k_stage1|k_bwd1' -s 40 -c 16 -f -o gpurun_out/r02_staged \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-graphs --fast-setup > gpurun_out/r02_under_ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
