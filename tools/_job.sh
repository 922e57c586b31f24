for w in delade_lex delade_ref_lex delade_cls; do
timeout 120 python tools/k1t_bench.py --workload $w 2>&1 | tail -1 | cut -c1-200
done
