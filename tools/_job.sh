timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for l in 1 2; do
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --option lanes=$l > gpurun_out/r2_bench_lanes$l.json 2>gpurun_out/r2_bench_lanes$l.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_lanes$l.json')); print('lanes $l', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['verified']['ok'], d['verified']['max_abs_score_err'], d['verified']['missed_rows'], d['config']['index_bytes'])"
tail -2 gpurun_out/r2_bench_lanes$l.err
done
timeout 200 python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_dense_l2.json 2>gpurun_out/r2_bench_dense_l2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_dense_l2.json')); print('dense', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['verified']['ok'])"
