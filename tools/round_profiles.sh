#!/bin/bash
# everything profiles/ holds for round 2, in one 1-GPU call (run under gpurun from the repo root)
set -x
mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c4.json 2> gpurun_out/bench.err
python bench.py --steps 5 --warmup 3 --workload c5 --md-steps 200 --no-extra > gpurun_out/r02_bench_c5.json 2>> gpurun_out/bench.err
python bench.py --steps 5 --warmup 2 --workload c3 --md-steps 2000 > gpurun_out/r02_bench_c3.json 2>> gpurun_out/bench.err
python bench.py --steps 5 --warmup 2 --workload c3 --md-steps 2000 --opt remd_inkernel=0 > gpurun_out/r02_bench_c3_launch_chain.json 2>> gpurun_out/bench.err
python bench.py --steps 5 --warmup 2 --workload c3 --md-steps 2000 --opt rep_kernel=0 > gpurun_out/r02_bench_c3_k_md_fast.json 2>> gpurun_out/bench.err
for n in 16 148 296 592; do python bench.py --steps 5 --warmup 2 --workload c3 --nrep $n --md-steps 2000 > gpurun_out/r02_bench_c3_nrep$n.json 2>> gpurun_out/bench.err; done
(for a in "--nrep 64 --nex 10" "--nrep 64 --nex 0" "--nrep 592 --nex 0" "--nrep 64 --nex 0 --mc 1" "--nrep 64 --nex 0 --mc 1 --opt rep_kernel=0"; do python tools/rep_profile.py $a; done) > gpurun_out/r02_k_md_rep_phase_clocks.txt 2>&1
(python tools/bench_remd_sweep.py --nrep 8; python tools/bench_remd_sweep.py --nrep 64; python tools/bench_remd_sweep.py --nrep 256) > gpurun_out/r02_remd_sweep.txt 2>&1
python tools/c1_speed.py > gpurun_out/r02_c1_speed.txt 2>&1
python tools/small_sys.py > gpurun_out/r02_small_sys.txt 2>&1
# launch lists of the bench commands (cold-cache, serialised: shares, not absolutes)
tools/ncu_launches.sh gpurun_out/r02_launches_c4.csv python bench.py --steps 2 --warmup 1 --md-steps 300 --no-cpu --no-extra > gpurun_out/r02_launches_c4_summary.txt 2>&1
tools/ncu_launches.sh gpurun_out/r02_launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 1 --md-steps 100 --no-cpu --no-extra > gpurun_out/r02_launches_c5_summary.txt 2>&1
NCU_COUNT=1500 tools/ncu_launches.sh gpurun_out/r02_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --md-steps 500 > gpurun_out/r02_launches_c3_summary.txt 2>&1
# full captures of the step kernels: C4 first launch after fresh lists, 40 MD steps (no rebuild inside); C5 one MD step
ncu --set full --import-source on --clock-control none -k regex:k_md_fast -c 1 -o gpurun_out/r02_k_md_fast_c4 python bench.py --steps 1 --warmup 1 --md-steps 40 --no-cpu --no-extra > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_md_big -s 60 -c 1 -o gpurun_out/r02_k_md_big_c5 python bench.py --workload c5 --steps 1 --warmup 1 --md-steps 100 --no-cpu --no-extra > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_md_rep -s 1 -c 1 -o gpurun_out/r02_k_md_rep_c3 python bench.py --workload c3 --steps 1 --warmup 1 --md-steps 200 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_md_rep -s 1 -c 1 -o gpurun_out/r02_k_md_rep_c3_592 python bench.py --workload c3 --nrep 592 --steps 1 --warmup 1 --md-steps 100 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_nl_rows -c 1 -o gpurun_out/r02_k_nl_rows_c5 python bench.py --workload c5 --steps 1 --warmup 1 --md-steps 100 --no-cpu --no-extra > /dev/null 2>&1
# summaries here (ncu reads its own reports without a GPU too); the reports themselves are too large to bring back together
for k in k_md_rep_c3 k_md_rep_c3_592 k_md_fast_c4 k_md_big_c5 k_nl_rows_c5; do
    python tools/ncu_summary.py gpurun_out/r02_$k.ncu-rep --top 14 > gpurun_out/r02_${k}_ncu.txt 2>/dev/null
    python tools/ncu_lines.py gpurun_out/r02_$k.ncu-rep >> gpurun_out/r02_${k}_ncu.txt 2>/dev/null
done
cp profiles/r02_kernel_metrics.json gpurun_out/ 2>/dev/null
python tools/ncu_traffic.py gpurun_out/r02_k_md_fast_c4.ncu-rep c4 40
python tools/ncu_traffic.py gpurun_out/r02_k_md_big_c5.ncu-rep c5 1
python tools/ncu_traffic.py gpurun_out/r02_k_md_rep_c3.ncu-rep c3 200
python tools/ncu_traffic.py gpurun_out/r02_k_md_rep_c3_592.ncu-rep c3_592 100
cp profiles/r02_kernel_metrics.json gpurun_out/r02_kernel_metrics.json
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/ | tail -40
tail -n 2 gpurun_out/bench.err gpurun_out/ref.err
