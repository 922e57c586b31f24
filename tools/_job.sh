timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err; echo "default rc=$?"
timeout 300 python bench.py --workload delade_cls_zipf --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_delade_cls_zipf.json 2> gpurun_out/r2_bench_n1_delade_cls_zipf.err; echo "zipf rc=$?"
timeout 300 python bench.py --workload bm25 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_bm25.json 2> gpurun_out/r2_bench_n1_bm25.err; echo "bm25 rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_default_n1','r2_bench_n1_delade_cls_zipf','r2_bench_n1_bm25'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value'],1), round(d['e2e']['value'],1), d['verified']['ok'], d['roofline']['frac'], d['clocks'].get('reasons'))
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
