#!/bin/bash
# compute-sanitizer passes over the small-shape GPU tests: memcheck, racecheck, synccheck, initcheck.
# Each tool runs in its own process under `timeout`; summaries land in gpurun_out/sanitizer_<tool>.txt.
mkdir -p gpurun_out
SEL='golden_grid_bit_exact or golden_dense_only or tile_path_options_agree or rerank_and_merge or search_keys_equals_search or merge_keys or pipelined_searcher_single or rowmajor'
for tool in memcheck racecheck synccheck initcheck; do
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_api.py -m gpu -q -x -k "$SEL" \
      > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitizer_$tool.txt | tail -4
done
