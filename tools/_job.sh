timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x -k "dense or cta_pair or ip_retrieval" 2>&1 | tail -3
for v in 1 2; do
timeout 200 python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline --dense-variant $v > gpurun_out/r2_bench_dense_v$v.json 2>gpurun_out/r2_bench_dense_v$v.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_dense_v$v.json')); print('variant $v', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['verified']['ok'], d['verified']['max_abs_score_err'], d['verified']['missed_rows'])"
tail -2 gpurun_out/r2_bench_dense_v$v.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_dense_launches_c.csv python bench.py --workload dense --queries 256 --steps 1 --warmup 1 --no-verify --no-cpu-baseline > gpurun_out/r2_dense_ncu_c.log 2>&1
grep dense_tile gpurun_out/r2_dense_launches_c.csv | tail -8 | awk -F'","' '{print $NF}'
