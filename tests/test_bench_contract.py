"""The committed bench lines (profiles/) carry every key of the bench.py contract; guards the JSON shape without a GPU."""
import json
import os

from conftest import ROOT


def _load(name):
    with open(os.path.join(ROOT, 'profiles', name)) as f:
        lines = [l for l in f.read().splitlines() if l.startswith('{')]
    assert len(lines) == 1, 'one JSON line per bench run'
    return json.loads(lines[0])


def test_own_arm_line_has_the_contract_keys():
    d = _load('r2_bench_default_n1.json')
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'roofline', 'e2e', 'gpu_launches', 'clocks', 'cpu_baseline', 'verified'):
        assert key in d, key
    assert d['metric'] == 'queries/sec' and d['higher_is_better'] is True and d['vs_baseline'] is None and d['data'] == 'synthetic'
    assert d['warmup'] >= 3 and d['n_gpus'] == 1 and 'workload' in d['config'] and 'model' not in d['config']
    r = d['roofline']
    for key in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert key in r, key
    assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    e = d['e2e']
    assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and 0 < e['value'] <= d['value'] * 1.05
    c = d['cpu_baseline']
    assert c['kind'] in ('reference', 'port') and c['cores'] >= 1 and c['value'] > 0 and c['sample']
    assert d['gpu_launches'] > 0
    v = d['verified']
    assert v['ok'] is True and v['missed_rows'] == 0 and v['rows_checked'] == v['rows_expected'] and v['max_abs_score_err'] <= 1e-3
    assert c['kind'] == 'reference' and 'one_thread' in c and 'linearity' in c and 'index_bytes' in d['config']
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}


def test_reference_arm_line():
    d = _load('r2_bench_reference_arm.json')
    own = _load('r2_bench_default_n1.json')
    assert d['impl'] == 'reference' and d['metric'] == own['metric'] and d['unit'] == own['unit']
    assert d['config']['workload'] == own['config']['workload']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['cpu_baseline']['value'] == d['value'] and d['cpu_baseline']['kind'] in ('reference', 'port')


def test_scaling_lines_are_single_json_lines():
    for n in (2, 4, 8):
        d = _load('r2_bench_n%d.json' % n)
        assert d['n_gpus'] == n and d['scaling'] == 'strong' and d['value'] > 0
        assert d['verified']['ok'] is True and 'breakdown' in d                    # merged NCCL answer checked at the benchmarked size


def test_other_workload_lines_are_verified():
    for w in ('delade_cls_ref', 'bm25', 'bm25_ref', 'dense', 'delade_cls_zipf'):
        d = _load('r2_bench_n1_%s.json' % w)
        assert d['verified']['ok'] is True and d['roofline']['bound'] == ('tensor' if w == 'dense' else 'hbm')
