#!/bin/bash
# Runs on the B200 box under gpurun: every GPU test file separately (own timeout, own log), then
# smoke, bench and the ncu launch list.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_graph_gpu test_lie_gpu test_corr_gpu test_ba_gpu test_update_gpu test_parity_ref_gpu; do
  if [ -f tests/$t.py ]; then
    DPVO_GOLDEN_OUT=gpurun_out/golden timeout -s KILL 600 python -m pytest tests/$t.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$t.log 2>&1
    echo "== $t: exit $? : $(tail -1 gpurun_out/$t.log)"
  fi
done
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "== bench exit $?"; tail -2 gpurun_out/bench.log
