#!/bin/bash
# ncu --set full captures of the scan kernels of every bench workload at FULL size (8,841,823 rows, one batch of 256 queries):
# steady-state sub-chunk launches of the last chunk (tight admission threshold, like most launches of the benchmark)
# -> gpurun_out/r2_ncu_<workload>.ncu-rep; tools/make_traffic.py condenses them into profiles/traffic.json and
# profiles/r2_ncu_<workload>.txt (run here, no GPU needed).
mkdir -p gpurun_out
run() {  # workload, launches to skip, launches to capture
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'lex_tile_kernel|dense_tile_ts' -s $2 -c $3 \
      -f -o gpurun_out/r2_ncu_$1 python bench.py --workload $1 --queries 256 --steps 1 --warmup 1 --no-verify --no-cpu-baseline > gpurun_out/r2_ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
for w in ${WORKLOADS:-delade_cls delade_cls_ref delade_cls_zipf}; do run $w 600 4; done
for w in ${WORKLOADS_LEX:-bm25 bm25_ref}; do run $w 300 2; done
run dense 7 1
