timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "n8 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n8.json')); print('n8', d['value'], d['ms_per_step'], d['verified']['ok'], d['e2e']['value'], d.get('breakdown'))"
tail -2 gpurun_out/r2_bench_n8.err
