#!/bin/bash
# perf-attribution run (library built with DPVO_B200_PERF_EXPERIMENTS=1): GEMM tile timelines, micro-benchmarks, launch list
mkdir -p gpurun_out
for e in 0 1 5; do DPVO_B200_GEMM_TIMING=1 timeout -s KILL 100 python tools/one_gemm.py $e 2>&1 | grep -A7 "linear_f16 CTA 0" | tail -8; done > gpurun_out/gemm_tiles.txt 2>&1
cat gpurun_out/gemm_tiles.txt
timeout -s KILL 200 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cat gpurun_out/bench_gemm.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02a.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-reference-cuda > gpurun_out/ncu_launches.log 2>&1; echo "== ncu launches exit $?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_r02a.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); idc=h.index('ID')
recs=[(int(r[idc]), r[kn], float(r[mv].replace(',',''))) for r in rows[hdr+1:] if len(r)>mv]
# last update = last 40 kernels up to the last ba_solve
last=max(i for i,(n,k,v) in enumerate(recs) if 'ba_solve' in k)
start=max(i for i,(n,k,v) in enumerate(recs[:last-3]) if 'reproject' in k)
tot=0
for n,k,v in recs[start:last+1]:
    print("%8.1f  %s" % (v/1000.0, k[:90])); tot+=v
print("sum %.1f us" % (tot/1000.0))
PY
