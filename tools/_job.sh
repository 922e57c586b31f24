timeout 300 python bench.py --rows 1105228 --steps 3 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/shard8.json 2>gpurun_out/shard8.err
python -c "
import json; d=json.load(open('gpurun_out/shard8.json')); print('shard8 proxy', d['value'], d['ms_per_step'], d['roofline'].get('scan_stream_ms_per_step'), d['roofline'].get('select_stream_ms_per_step'))"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dense_tile|select|init_slots|lex_tile_kernel|merge' -s 200 -c 80 --csv --log-file gpurun_out/shard8_launches.csv python bench.py --rows 1105228 --queries 512 --steps 1 --warmup 1 --no-cpu-baseline --no-verify --option lanes=1 > gpurun_out/shard8_ll.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/shard8_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:40], r[vi], r[hdr.index('Grid Size')])
PY
