#!/bin/bash
# ncu --set full captures of the scan kernels of every bench workload (one steady-state sub-chunk each) -> gpurun_out/r2_ncu_<workload>.ncu-rep
# then tools/make_traffic.py condenses them into profiles/traffic.json and profiles/r2_ncu_<workload>.txt (run here, no GPU needed).
mkdir -p gpurun_out
for w in ${WORKLOADS:-delade_cls delade_cls_ref bm25 bm25_ref dense}; do
  # skip the launches of the two warm-up searches and the first chunks of the measured one; capture one K2 + K1t pair (or 2 x K2)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lex_tile_kernel|dense_tile_ts' -s ${SKIP:-60} -c 4 \
      -f -o gpurun_out/r2_ncu_$w python tools/k1t_bench.py --workload $w --rows 606208 --reps 1 --no-check > gpurun_out/r2_ncu_$w.log 2>&1
  echo "$w rc=$?"
done
