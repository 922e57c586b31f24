#!/usr/bin/env python
"""bench.py -- queries/sec of the GIP retrieval hot path at MS MARCO scale (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                          # reference algorithm on the host cores

A step = one search of the whole query set (Q queries, top-k) over the resident corpus.
`value`   queries/sec with queries and results resident in HBM (device pointers through the API),
`e2e`     the same through the public API with HOST (pinned) query buffers and host result buffers,
          i.e. with the H2D / D2H copies inside the timed region,
`roofline` achieved HBM bandwidth of the dominant kernel (the fused scan K1): logical corpus passes x
          N x row_bytes / summed CUDA-event duration of the scan launches, vs the measured HBM peak,
`cpu_baseline` the torch-op port of the reference loop (oracle/gip_oracle.py) on the host cores, on a
          bounded sample (rank 0, N=1 only).
Synthetic encoded vectors of MS MARCO shape (dhr_b200/synth.py); inputs are far larger than L2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='delade_cls')
    ap.add_argument('--rows', type=int, default=None, help='corpus rows (default 8,841,823)')
    ap.add_argument('--queries', type=int, default=None, help='queries per step (default 6808)')
    ap.add_argument('--topk', type=int, default=1000)
    ap.add_argument('--query-block', type=int, default=None)
    ap.add_argument('--query-groups', type=int, default=None)
    ap.add_argument('--scan-variant', type=int, default=None)
    ap.add_argument('--overlap', type=int, default=None, help='hybrid tile path: K2 on a second stream (1, default) or in line (0)')
    ap.add_argument('--dense-multicast', type=int, default=None, help='K2: cluster of two query groups sharing corpus tiles by TMA multicast (1, default) or not (0)')
    ap.add_argument('--dense-variant', type=int, default=None, help='K2: 1 = queries in TMEM (default), 0 = both operands in shared memory')
    ap.add_argument('--cpu-rows', type=int, default=400000, help='rows of the bounded CPU-baseline sample')
    ap.add_argument('--cpu-queries', type=int, default=24)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines, self.t = gpu_index, None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
        self.t.start()

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


class CpuPort:
    """Torch-op port of the reference loop (gip_retrieval.py:110-126 / :70-79) on a bounded sample of `rows` rows; the
    algorithm is exactly linear in N, so q/s at n_total rows = measured q/s * rows / n_total."""

    def __init__(self, workload, rows, n_queries, topk, threads):
        import torch
        from dhr_b200 import synth
        from oracle import gip_oracle as go
        self.go, self.torch = go, torch
        self.cfg = cfg = synth.CONFIGS[workload]
        torch.set_num_threads(threads)
        cv, ci = synth.corpus_numpy(workload, 0, rows)
        qv, qi = synth.queries_numpy(workload, n_queries)
        G = cfg['G']
        self.c = torch.from_numpy(cv.astype(np.float32))              # :313 the CPU path works on fp32 copies
        self.q = torch.from_numpy(qv.astype(np.float32))
        self.qids = list(range(n_queries))
        self.k = min(topk, rows)
        if cfg['S'] > 0:                                              # G > 1: one idx per value column
            self.cidx = torch.from_numpy(np.repeat(ci, G, axis=1).astype(np.int16))
            self.qidx = torch.from_numpy(np.repeat(qi, G, axis=1).astype(np.int16))

    def run(self, n=None):
        """time the port on the first n (default all) sample queries"""
        go, cfg = self.go, self.cfg
        n = len(self.qids) if n is None else min(n, len(self.qids))
        t0 = time.perf_counter()
        if cfg['S'] > 0:
            go.GIP_retrieval_port(self.qids[:n], self.q[:n], self.qidx[:n], self.c, self.cidx,
                                  go.make_args(emb_dim=cfg['S'] * cfg['G'], topk=self.k, brute_force=True))
        else:
            go.IP_retrieval_port(self.qids[:n], self.q[:n], self.c, go.make_args(topk=self.k))
        return time.perf_counter() - t0


def workload_string(workload, n_total, n_q, k):
    from dhr_b200 import synth
    cfg = synth.CONFIGS[workload]
    return '%s: %d passages, %d queries, S=%d x G=%d lexical (%s idx) + %d dense fp16, top-%d' % (
        workload, n_total, n_q, cfg['S'], cfg['G'], cfg['idx'], cfg['C'], k)


def run_reference(args):
    """--impl reference: the reference algorithm (torch-op port, kind "port": the reference is pure Python and
    /root/reference does not exist on the GPU box) on the host cores, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from dhr_b200 import synth
    import torch
    n_total = args.rows or synth.N_MSMARCO
    threads = os.cpu_count() or 1
    rows = min(args.cpu_rows, n_total)
    n_q = args.queries or synth.Q_MSMARCO
    port = CpuPort(args.workload, rows, args.cpu_queries, args.topk, threads)
    times = []
    for i in range(args.warmup + args.steps):
        dt = port.run()
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    qps = args.cpu_queries / (ms / 1e3) * rows / n_total
    sample = '%d queries x %d rows per step, torch %s CPU, %d threads, scaled linearly to %d rows' % (
        args.cpu_queries, rows, torch.__version__, threads, n_total)
    line = {
        'impl': 'reference', 'metric': 'queries/sec', 'value': qps, 'unit': 'queries/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_string(args.workload, n_total, n_q, args.topk)},
        'cpu_baseline': {'value': qps, 'unit': 'queries/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': qps, 'unit': 'queries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dhr_b200 import GipIndex, synth, topk_merge
    from dhr_b200.gip_retrieval import shard_bounds

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL prints its version banner on fd 1 at the first collective; stdout must carry the JSON line only, so fd 1 points
        # to stderr until the result is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)

    cfg = synth.CONFIGS[args.workload]
    n_total = args.rows or synth.N_MSMARCO
    n_q = args.queries or synth.Q_MSMARCO
    k = args.topk
    lo, hi = shard_bounds(n_total, world, rank)

    # ---- build the resident shard (not timed: index load is outside the metric) ----
    t_build = time.perf_counter()
    ix = GipIndex(cfg['S'], cfg['C'], cfg['G'], capacity=hi - lo, idx_dtype=np.dtype(cfg['idx']), device=local_rank, row_offset=lo)
    for vals, idx in synth.corpus_torch_segments(args.workload, lo, hi, dev):
        ix.append(vals, idx)
    ix.finalize()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    if args.query_block is not None:
        ix.set_option('query_block', args.query_block)
    if args.query_groups is not None:
        ix.set_option('query_groups', args.query_groups)
    if args.scan_variant is not None:
        ix.set_option('scan_variant', args.scan_variant)
    if args.overlap is not None:
        ix.set_option('overlap', args.overlap)
    if args.dense_variant is not None:
        ix.set_option('dense_variant', args.dense_variant)
    if args.dense_multicast is not None:
        ix.set_option('dense_multicast', args.dense_multicast)
    ix.set_option('profile', 1)

    qv_dev, qi_dev = synth.queries_torch(args.workload, n_q, dev)
    qv_host = qv_dev.cpu().pin_memory()
    qi_host = qi_dev.cpu().pin_memory() if qi_dev is not None else None
    out_dev = (torch.empty((n_q, k), dtype=torch.float32, device=dev), torch.empty((n_q, k), dtype=torch.int64, device=dev),
               torch.empty((n_q,), dtype=torch.int32, device=dev))
    out_host = (torch.empty((n_q, k), dtype=torch.float32).pin_memory(), torch.empty((n_q, k), dtype=torch.int64).pin_memory(),
                torch.empty((n_q,), dtype=torch.int32).pin_memory())
    if world > 1:
        gat_s = torch.empty((world, n_q, k), dtype=torch.float32, device=dev)
        gat_r = torch.empty((world, n_q, k), dtype=torch.int64, device=dev)

    from dhr_b200.distributed import sharded_search

    def step(host_io):
        """one search of all queries over the (sharded) corpus; returns the final [Q,k] (scores, rows)"""
        qv, qi = (qv_host, qi_host) if host_io else (qv_dev, qi_dev)
        if world == 1:
            res = ix.search(qv, qi, k, out=out_host if host_io else out_dev)[:2]
            st = ix.stats()
        else:
            # per-shard search, then the exchange step: NCCL all-gather of the [Q,k] lists + merge kernel
            res = sharded_search(ix, qv, qi, k, local_out=out_dev, gather_out=(gat_s, gat_r))
            st = ix.stats()
            st['n_kernel_launches'] += 1
            if host_io and rank == 0:
                out_host[0].copy_(res[0], non_blocking=True)
                out_host[1].copy_(res[1], non_blocking=True)
        return res, st

    def timed(n_steps, host_io):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats = []
        e0.record()
        for _ in range(n_steps):
            _, st = step(host_io)
            stats.append(st)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), stats

    for _ in range(max(3, args.warmup)):
        step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, stats = timed(args.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    step(True)
    ms_e2e, _ = timed(args.steps, True)

    # HBM-bound operating point of the scan (K1, one query per corpus pass, one group per launch): bounded sample
    qb1 = None
    if cfg['S'] > 0 or True:
        n_s = min(8, n_q)
        ix.set_option('tile_mode', 0); ix.set_option('query_block', 1); ix.set_option('query_groups', 1); ix.set_option('scan_variant', 1)
        ix.search(qv_dev[:n_s], qi_dev[:n_s] if qi_dev is not None else None, k, out=tuple(o[:n_s] for o in out_dev))
        ix.search(qv_dev[:n_s], qi_dev[:n_s] if qi_dev is not None else None, k, out=tuple(o[:n_s] for o in out_dev))
        s1 = ix.stats()
        qb1 = {'scan_ms': s1['scan_ms'], 'passes': s1['corpus_passes'], 'queries': n_s, 'total_ms': s1['total_ms']}

    # ---- roofline of the dominant kernel (K1 scan), from CUDA events inside the library ----
    scan_ms = sum(s['scan_ms'] for s in stats)
    select_ms = sum(s['select_ms'] for s in stats)
    passes = sum(s['corpus_passes'] for s in stats)
    launches = sum(s['n_scan_launches'] for s in stats)
    bytes_per_pass = stats[0]['bytes_per_pass']
    peaks, peak_src = load_peaks()
    achieved = passes * bytes_per_pass / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(args.workload)
        except Exception:
            traffic = None

    if rank == 0:
        qps = n_q * args.steps / (ms_dev / 1e3)
        qps_e2e = n_q * args.steps / (ms_e2e / 1e3)
        W = cfg['S'] * cfg['G'] + cfg['C']
        h2d = int(qv_host.numel() * qv_host.element_size() + (qi_host.numel() * qi_host.element_size() if qi_host is not None else 0))
        d2h = int(n_q * k * 12 + (n_q * 4 if world == 1 else 0))
        line = {
            'metric': 'queries/sec', 'value': qps, 'unit': 'queries/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32',  # fp16 storage, exact fp16 x fp16 products accumulated in fp32 (FHFMA)
            'data': 'synthetic',
            'config': {
                'workload': workload_string(args.workload, n_total, n_q, k),
                'row_bytes': ix.row_bytes, 'corpus_bytes': ix.row_bytes * n_total, 'parallelism': 'range-shard x%d' % world,
                'query_block': stats[0]['query_block'], 'query_groups': stats[0]['query_groups'], 'scan_variant': stats[0]['scan_variant'],
                'l2': 'inputs (%.1f GB per GPU) far exceed the 126 MB L2' % (ix.row_bytes * (hi - lo) / 1e9),
                'index_build_s': t_build,
            },
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
                         'traffic': traffic, 'peak_source': peak_src,
                         'kernel': {1: 'gip_scan_tma (K1)', 0: 'gip_scan_direct (K1)', 2: 'dense_tile_ts (K2, tcgen05, queries in TMEM)',
                                    3: 'dense_tile_ts (K2, tcgen05, queries in TMEM) + lex_tile (K1t)'}.get(stats[0]['scan_variant'], '?'),
                         'queries_per_pass': stats[0]['query_block'],
                         'note': 'achieved = logical corpus passes (one per query tile of `queries_per_pass`) x N x row_bytes / kernel time; '
                                 'query tiles in flight share the pass through L2, so DRAM traffic is lower (see traffic)',
                         'bytes_per_launch': passes * bytes_per_pass / max(1, launches), 'launch_ms_avg': scan_ms / max(1, launches),
                         'corpus_passes_per_step': passes / args.steps, 'scan_share_of_step': scan_ms / ms_dev,
                         'select_share_of_step': select_ms / ms_dev},
            'e2e': {'value': qps_e2e, 'unit': 'queries/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(sum(s['n_kernel_launches'] for s in stats)),
            'clocks': clocks,
            'fallback_queries': int(sum(s['n_fallback_queries'] for s in stats)),
        }
        if qb1 and qb1['scan_ms'] > 0:
            a1 = qb1['passes'] * bytes_per_pass / (qb1['scan_ms'] / 1e3) / 1e9
            line['roofline_qb1'] = {'bound': 'hbm', 'achieved': a1, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': a1 / peaks['hbm_gbs'],
                                    'kernel': 'gip_scan_tma (K1), one query per corpus pass', 'queries': qb1['queries'],
                                    'queries_per_sec': qb1['queries'] / (qb1['total_ms'] / 1e3) if qb1['total_ms'] > 0 else None}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rows = min(args.cpu_rows, n_total)
            port = CpuPort(args.workload, rows, args.cpu_queries, k, threads)
            port.run(2)                       # warm-up on two queries
            dt = port.run()
            v = args.cpu_queries / dt * rows / n_total
            line['cpu_baseline'] = {
                'value': v, 'unit': 'queries/s', 'cores': threads, 'kind': 'port',
                'sample': '%d queries x %d rows (%.1f s), torch-op port of gip_retrieval.py:110-126, scaled linearly to %d rows' % (
                    args.cpu_queries, rows, dt, n_total)}
        if world > 1:
            import ctypes
            ctypes.CDLL(None).fflush(None)                  # C stdio may still buffer the banner
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    ix.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
