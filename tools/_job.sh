timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu -k "dense or cta_pair or unmasked or tile_path_options or golden or overflow or staging" 2>&1 | tail -3
timeout 200 python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/dense_new.json 2>gpurun_out/dense_new.err
python -c "
import json; d=json.load(open('gpurun_out/dense_new.json')); print('dense', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['verified']['ok'])"
tail -2 gpurun_out/dense_new.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/def_new.json 2>gpurun_out/def_new.err
python -c "
import json; d=json.load(open('gpurun_out/def_new.json')); print('default', d['value'], d['ms_per_step'], d['verified']['ok'])"
tail -2 gpurun_out/def_new.err
