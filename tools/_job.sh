for pf in 0 1; do
timeout 200 python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline --option dense_prefetch=$pf > gpurun_out/dense_pf$pf.json 2>gpurun_out/dense_pf$pf.err
python -c "
import json; d=json.load(open('gpurun_out/dense_pf$pf.json')); print('dense pf$pf', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['verified']['ok'])"
tail -2 gpurun_out/dense_pf$pf.err
done
for pf in 0 1; do
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --option dense_prefetch=$pf > gpurun_out/def_pf$pf.json 2>gpurun_out/def_pf$pf.err
python -c "
import json; d=json.load(open('gpurun_out/def_pf$pf.json')); print('default pf$pf', d['value'], d['ms_per_step'], d['verified']['ok'])"
tail -2 gpurun_out/def_pf$pf.err
done
