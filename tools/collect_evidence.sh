#!/bin/bash
# Round evidence on one B200 (run under gpurun): tests, bench lines, launch list, ncu captures, stand-alone tools.
#   bash tools/collect_evidence.sh TAG      -> gpurun_out/TAG_*
T="${1:-rXX}"; O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${T}_gputests.log
python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --steps 200 --warmup 5 --no-matrix > $O/${T}_bench_200steps.json 2>> $O/${T}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
python tools/timeline.py > $O/${T}_timeline.log 2>&1
python tools/solver_bench.py > $O/${T}_solver_bench.log 2>&1
python tools/lk_bench.py > $O/${T}_lk_bench.log 2>&1
python tools/lk_bench_bgr.py > $O/${T}_lk_bench_bgr.log 2>&1
python tools/orb_bench.py 2>&1 | tail -6 > $O/${T}_orb_bench.log
python tools/sgbm_bench.py 2>&1 | tail -12 > $O/${T}_sgbm_bench.log
# profiler passes (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${T}_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-matrix > $O/${T}_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_orb_launches.csv python tools/orb_bench.py > $O/${T}_ncu_orb.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^lk_kernel$" -s 5 -c 1 -f -o $O/${T}_prof_lk python tools/lk_bench.py > $O/${T}_ncu_lk.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:pnp_solve_kernel|pnp_refine_kernel|fmat_solve_kernel|pnp_score_kernel|mask_compact_kernel" -s 10 -c 6 -f -o $O/${T}_prof_solvers python tools/solver_bench.py > $O/${T}_ncu_solvers.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:orb_" -s 40 -c 20 -f -o $O/${T}_prof_orb python tools/orb_bench.py > $O/${T}_ncu_orb2.log 2>&1
echo done
