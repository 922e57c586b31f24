#!/bin/bash
# First-light on a GPU box: each stage in its own process under `timeout`, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
for v in 0 1; do
timeout 180 python - <<PY > gpurun_out/smoke_v$v.log 2>&1
import sys, numpy as np
sys.path.insert(0, '.')
from dhr_b200 import GipIndex, synth
from oracle import c_oracle, gip_oracle
cv, ci = synth.corpus_numpy('delade_cls', 0, 20000)
qv, qi = synth.queries_numpy('delade_cls', 8)
with GipIndex.from_arrays(cv, ci, n_slices=128, group=6) as ix:
    ix.set_option('scan_variant', $v)
    for qb in (1, 2, 4, 8):
        ix.set_option('query_block', qb)
        s, r, c = ix.search(qv, qi, 100)
        bad = 0
        for i in range(8):
            ex = c_oracle.scores(cv, ci, qv[i].astype(np.float32), qi[i], 128, 6)
            m = gip_oracle.check_topk_against_exact(r[i], s[i], ex, 100)
            if m: bad += 1; print('variant $v qb', qb, 'query', i, m)
        print('variant $v qb', qb, 'bad', bad, ix.stats())
PY
echo "smoke variant $v rc=$?" >> gpurun_out/stages.txt
done
cat gpurun_out/stages.txt gpurun_out/smoke_v0.log gpurun_out/smoke_v1.log | tail -40
