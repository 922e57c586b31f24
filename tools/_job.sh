timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu -k "dense or cta_pair or unmasked or tile_path_options or golden or overflow or staging" 2>&1 | tail -4
timeout 200 python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/dense_new.json 2>gpurun_out/dense_new.err
python -c "
import json; d=json.load(open('gpurun_out/dense_new.json')); print('dense', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['verified']['ok'])"
tail -2 gpurun_out/dense_new.err
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -DDHR_K2_TRACE -o gpurun_out/k2_micro tools/k2_micro.cu -lcuda > gpurun_out/k2_probe_c.txt 2>&1
for tau in 1e30 0.06; do
  echo "== k2_micro 221045 256 768 1 0 1 $tau" >> gpurun_out/k2_probe_c.txt
  timeout 120 ./gpurun_out/k2_micro 221045 256 768 1 0 1 $tau >> gpurun_out/k2_probe_c.txt 2>&1
done
grep -E "==|mode 0|rror|dbg" gpurun_out/k2_probe_c.txt
