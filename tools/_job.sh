timeout 300 python bench.py --unmasked --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_unmasked_ip.json 2>gpurun_out/r2_bench_n1_unmasked_ip.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_unmasked_ip.json')); print('unmasked', d['value'], d['ms_per_step'], d['roofline']['bound'], d['roofline']['achieved'], d['roofline']['frac'], d['verified']['ok'], d['verified']['max_abs_score_err'], d['config']['scan_variant'], d['config']['index_bytes'])"
tail -2 gpurun_out/r2_bench_n1_unmasked_ip.err
