#!/bin/bash
# Runs on the B200 box under gpurun: every GPU test file separately (own timeout, own log), then smoke, the bench
# (ours + reference arm) and, with PROFILE=1, the artefacts summarised under profiles/: ncu launch list of one update,
# ncu --set full of the main kernels, the fast.yaml bench line, the scaled-E bundle-adjustment measurement and the
# stand-alone chain-kernel timings.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
TESTS="${TESTS:-test_chain_gpu test_pgraph_gpu test_gemm_gpu test_update_gpu test_graph_gpu test_lie_gpu test_corr_gpu test_ba_gpu test_parity_ref_gpu test_projective_gpu test_step_gpu test_dropin_gpu test_train_gpu}"
for t in $TESTS; do
  if [ -f tests/$t.py ]; then
    timeout -s KILL 300 python -m pytest tests/$t.py -m gpu -q --tb=short -p no:cacheprovider --timeout 150 -s > gpurun_out/$t.log 2>&1
    echo "== $t: exit $? : $(tail -1 gpurun_out/$t.log)"
  fi
done
if [ -z "$SKIP_BENCH" ]; then
  timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
  timeout -s KILL 600 python bench.py > gpurun_out/bench_ours.log 2>&1; echo "== bench exit $?"; tail -1 gpurun_out/bench_ours.log | cut -c1-1200
  timeout -s KILL 600 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "== bench reference exit $?"; tail -1 gpurun_out/bench_reference.log | cut -c1-400
fi
if [ -n "$PROFILE" ]; then
  timeout -s KILL 400 python bench.py --config fast --no-cpu-baseline > gpurun_out/bench_fast.log 2>&1; echo "== bench fast exit $?"; tail -1 gpurun_out/bench_fast.log | cut -c1-300
  timeout -s KILL 300 python tools/bench_chain.py > gpurun_out/bench_chain.log 2>&1; cat gpurun_out/bench_chain.log
  timeout -s KILL 300 python tools/bench_ba_scaled.py > gpurun_out/ba_scaled.json 2>gpurun_out/ba_scaled.err; echo "== ba scaled exit $?"; tail -1 gpurun_out/ba_scaled.json
  timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-reference-cuda > gpurun_out/ncu_launches.log 2>&1; echo "== ncu launches exit $?"
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|linear_f16|corr_fwd_tc|ba_reduce|ba_solve|softagg|group_edges" -s 48 -c 16 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-reference-cuda > gpurun_out/ncu_final.log 2>&1; echo "== ncu full exit $?"
fi
