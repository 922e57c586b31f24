timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_sharded.py -m gpu -q -x -k "dense or cta_pair or ip_retrieval or unmasked or adversarial or overflow or golden_dense or edge or k_range" 2>&1 | tail -3
timeout 200 python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_dense_seg.json 2>gpurun_out/r2_bench_dense_seg.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_dense_seg.json')); print('dense', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['verified']['ok'], d['verified']['missed_rows'], d['fallback_queries'])"
tail -2 gpurun_out/r2_bench_dense_seg.err
