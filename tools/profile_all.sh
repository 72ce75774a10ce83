#!/bin/bash
# One GPU call that captures the round's ncu evidence into gpurun_out/prof/ (every command first runs WITHOUT ncu and must exit 0).
set -u
O=gpurun_out/prof; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
run() { name=$1; shift; kre=$1; shift; skip=$1; shift; cnt=$1; shift
  "$@" > $O/$name.plain.log 2>&1 || { echo "$name: plain run failed"; return; }
  timeout 900 $NCU -k "regex:$kre" -s $skip -c $cnt -f -o $O/$name "$@" > $O/$name.ncu.log 2>&1; echo "$name rc=$?"; }
run r02_tc_solve 'solve_tc' 3 1 python bench.py --steps 2 --warmup 3 --no-extras
run r02_adjoint_tc 'adjoint_tc_kernel|wgrad_tc_kernel|solve_tc_kernel' 0 3 python tools/profile_solve.py --mode grad --ncol 4736 --steps 18 --reps 1
run r02_fc1 'fc1_train' 1 1 python tools/profile_fc1.py --ncol 1 --steps 18
run r02_closure_uvt 'closure_uvt' 2 1 python tools/profile_closure_uvt.py
python bench.py --steps 2 --warmup 3 > $O/r02_bench_for_launches.json 2> $O/r02_bench_for_launches.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 > $O/launches.ncu.log 2>&1; echo "launches rc=$?"
# summaries are made here, on the box: the reports themselves (60+ MB) would exceed what gpurun copies back
S() { python tools/ncu_summary.py $O/$1.ncu-rep "$2" > $O/$1_ncu_summary.txt 2>/dev/null; }
S r02_tc_solve "-k regex:solve_tc -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-extras   (BASELINE config 2: 4096 columns x 1152 steps x 12 RHS evaluations; 147 CTAs of 28 columns)"
S r02_adjoint_tc "-k regex:'adjoint_tc_kernel|wgrad_tc_kernel|solve_tc_kernel' -c 3 python tools/profile_solve.py --mode grad --ncol 4736 --steps 18 --reps 1   (148 tiles of 32 columns, 18 steps x 2 sub-steps x 6 stages; the three kernels of the tensor-core adjoint: forward pass with stage records, reverse sweep, weight gradient)"
S r02_fc1 "-k regex:fc1_train -s 1 -c 1 python tools/profile_fc1.py --ncol 1 --steps 18   (ONE T-only column, CA + mPP base, 18 steps x 15 sub-steps x 6 stages, forward + adjoint on one CTA)"
S r02_closure_uvt "-k regex:closure_uvt -s 2 -c 1 python tools/profile_closure_uvt.py   (512 x 64 x 32 slab, 1024 tiles of 32 columns on 148 persistent CTAs)"
for k in r02_tc_solve r02_adjoint_tc r02_closure_uvt; do rm -f $O/$k.ncu-rep; done
ls -la $O
