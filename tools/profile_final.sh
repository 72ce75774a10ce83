#!/bin/bash
# Final evidence of the round in one GPU call: the bench line, the ncu --set full summary of the forward kernel and the launch list
# of the bench command (every command first runs WITHOUT ncu and must exit 0). Outputs in gpurun_out/prof/.
set -u
O=gpurun_out/prof; mkdir -p $O
python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final.err || { echo "bench failed"; exit 1; }
python bench.py --steps 2 --warmup 3 --no-extras > $O/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_tc -s 3 -c 1 -f -o $O/r02_tc_solve python bench.py --steps 2 --warmup 3 --no-extras > $O/tc.ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $O/r02_tc_solve.ncu-rep "-k regex:solve_tc -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-extras   (BASELINE config 2: 4096 columns x 1152 steps x 12 RHS evaluations; 147 CTAs of 28 columns)" > $O/r02_tc_solve_ncu_summary.txt 2>/dev/null
rm -f $O/r02_tc_solve.ncu-rep
python bench.py --steps 2 --warmup 3 > $O/r02_bench_for_launches.json 2> $O/r02_bench_for_launches.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 > $O/launches.ncu.log 2>&1; echo "launches rc=$?"
ls -la $O
