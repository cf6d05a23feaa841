#!/bin/bash
# everything profiles/ holds for this round, in one GPU call (run under gpurun from the repo root)
set -x
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/ref.err
python bench.py > gpurun_out/r01_bench_c4.json 2> gpurun_out/bench.err
python bench.py --steps 5 --warmup 3 --workload c5 --md-steps 200 --no-cpu > gpurun_out/r01_bench_c5.json 2>> gpurun_out/bench.err
python tools/bench_remd.py > gpurun_out/r01_remd_c3.json 2>> gpurun_out/bench.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
tools/ncu_launches.sh gpurun_out/r01_launches_c4.csv python bench.py --steps 2 --warmup 1 --md-steps 300 --no-cpu > gpurun_out/r01_launches_c4_summary.txt 2>&1
# full capture of the step kernel: first launch after fresh lists, 40 MD steps (no rebuild inside)
ncu --set full --import-source on --clock-control none -k regex:k_md_fast -c 1 -o gpurun_out/r01_k_md_fast_c4 python bench.py --steps 1 --warmup 1 --md-steps 40 --no-cpu > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_nl_rows -c 1 -o gpurun_out/r01_k_nl_rows_c4 python bench.py --steps 1 --warmup 1 --md-steps 200 --no-cpu > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_md_fused -c 1 -o gpurun_out/r01_k_md_fused_c5 python bench.py --steps 1 --warmup 1 --md-steps 20 --workload c5 --no-cpu > /dev/null 2>&1
ls -la gpurun_out/
tail -2 gpurun_out/bench.err gpurun_out/ref.err
