#!/bin/bash
# Runs on the B200 box under gpurun: every GPU test file separately (own timeout, own log), then
# smoke, bench and the ncu launch list.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nvidia-smi --help-query-gpu 2>/dev/null | grep -i -E "reasons|clocks_event" | head -40 >> gpurun_out/gpu.txt
TESTS="${TESTS:-test_graph_gpu test_lie_gpu test_corr_gpu test_ba_gpu test_gemm_gpu test_update_gpu test_parity_ref_gpu}"
for t in $TESTS; do
  if [ -f tests/$t.py ]; then
    timeout -s KILL 600 python -m pytest tests/$t.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$t.log 2>&1
    echo "== $t: exit $? : $(tail -1 gpurun_out/$t.log)"
  fi
done
if [ -z "$SKIP_BENCH" ]; then
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
for g in ${GEMMS:-tcgen05 cublas}; do
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --gemm $g --no-cpu-baseline > gpurun_out/bench_$g.log 2>&1; echo "== bench $g exit $?"; tail -1 gpurun_out/bench_$g.log | cut -c1-1500
done
fi
if [ -n "$PROFILE" ]; then
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${PROFILE}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --gemm ${PROFILE} > gpurun_out/ncu_launches.log 2>&1; echo "== ncu launches exit $?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:corr_fwd_mma -s 3 -c 1 -o gpurun_out/prof_corr python bench.py --steps 2 --warmup 1 --no-cpu-baseline --gemm ${PROFILE} > gpurun_out/ncu_corr.log 2>&1; echo "== ncu corr exit $?"
fi
