timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for w in delade_lex delade_ref_lex delade_cls; do
timeout 120 python tools/k1t_bench.py --workload $w 2>&1 | tail -1 | cut -c1-200
done
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_filt.json 2>gpurun_out/r2_bench_filt.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_filt.json')); print('n1', d['value'], d['ms_per_step'], d['verified']['ok'])"
