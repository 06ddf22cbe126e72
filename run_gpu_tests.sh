#!/bin/bash
# Runs on the B200 box under gpurun: every GPU test file separately (own timeout, own log), then smoke, the bench
# (ours + reference arm) and, with PROFILE=1, the ncu launch list / full captures / BA phase timing that are
# summarised under profiles/.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
TESTS="${TESTS:-test_chain_gpu test_pgraph_gpu test_gemm_gpu test_update_gpu test_graph_gpu test_lie_gpu test_corr_gpu test_ba_gpu test_parity_ref_gpu test_projective_gpu test_step_gpu test_dropin_gpu}"
for t in $TESTS; do
  if [ -f tests/$t.py ]; then
    timeout -s KILL 600 python -m pytest tests/$t.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -s > gpurun_out/$t.log 2>&1
    echo "== $t: exit $? : $(tail -1 gpurun_out/$t.log)"
  fi
done
if [ -z "$SKIP_BENCH" ]; then
  timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
  timeout -s KILL 600 python bench.py > gpurun_out/bench_ours.log 2>&1; echo "== bench exit $?"; tail -1 gpurun_out/bench_ours.log | cut -c1-3000
  timeout -s KILL 600 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "== bench reference exit $?"; tail -1 gpurun_out/bench_reference.log | cut -c1-600
fi
if [ -n "$PROFILE" ]; then
  timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1; echo "== ncu launches exit $?"
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"corr_fwd_tc|ba_solve|ba_reduce" -s 4 -c 5 -o gpurun_out/prof_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_final.log 2>&1; echo "== ncu full exit $?"
  DPVO_B200_BA_TIMING=1 timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph 2>&1 | grep "ba_solve phases" | tail -2 > gpurun_out/ba_phases.txt; cat gpurun_out/ba_phases.txt
  for e in 0 3 4; do DPVO_B200_GEMM_TIMING=1 timeout -s KILL 100 python tools/one_gemm.py $e 2>&1 | grep -A7 "linear_f16 CTA 0" | tail -8; done > gpurun_out/gemm_tiles.txt 2>&1
fi
