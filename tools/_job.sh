timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_sel.json 2>gpurun_out/r2_bench_sel.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_sel.json')); print('n1', d['value'], d['ms_per_step'], d['verified']['ok'], d['roofline']['select_stream_ms_per_step'])"
