#!/usr/bin/env python
"""bench.py -- queries/sec of the GIP retrieval hot path at MS MARCO scale (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                          # the reference's own gip_retrieval.py on the host cores

A step = one search of the whole query set (Q queries, top-k) over the resident corpus.
`value`    queries/sec with queries and results resident in HBM (device pointers through the C ABI),
`e2e`      the same through the public API with HOST (pinned) query buffers and host result buffers,
           i.e. with the H2D / D2H copies inside the timed region,
`roofline` the scan kernels (K2 tcgen05 dense block + K1t lexical tile walk, or K2 alone for a dense-only index) against the
           ceiling that bounds them: algorithmic bytes (DESIGN.md units) / summed CUDA-event duration of the scan launches
           vs the measured HBM peak, and 2*Q*N*C / time vs the measured sustained tensor peak; next to them the
           physical DRAM rate (ncu `dram__bytes` per launch, profiles/traffic.json) and the CUDA-core lane-op ceiling,
`verified` the answer checked where it is benchmarked: sample queries spread over the super-batches are re-scored over the
           WHOLE corpus with plain torch fp32 ops (the reference's own operator sequence, gip_retrieval.py:119-120) on the
           same device and compared with the returned top-k (scores, completeness, order); at N > 1 the merged NCCL
           result is what is checked; at N = 1 the tile path is also compared with the independent row-scan kernel K1,
`cpu_baseline` the reference file itself (oracle/_ref, kind "reference"; the torch-op port when the copy is absent) on the
           host cores, on a bounded sample (rank 0, N=1 only).
Synthetic encoded vectors of MS MARCO shape (dhr_b200/synth.py); inputs are far larger than L2.
"""
from __future__ import annotations

import argparse
import contextlib
import io
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
    ap.add_argument('--option', action='append', default=[], help='extra index option name=value (repeatable)')
    ap.add_argument('--cpu-rows', type=int, default=400000, help='rows of the bounded CPU-baseline sample')
    ap.add_argument('--cpu-queries', type=int, default=12)
    ap.add_argument('--cpu-full', action='store_true', help='--impl reference: the full SURVEY 8(d) procedure (16 shards, 1 thread and all cores)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--lex-postings', action='store_true', help='experimental postings lexical layout (kernel K1p) instead of the tiled one (K1t)')
    ap.add_argument('--unmasked', action='store_true', help='time the --IP first stage (gip_retrieval.py:139): plain inner product over all columns')
    ap.add_argument('--no-verify', action='store_true')
    ap.add_argument('--verify-queries', type=int, default=16)
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


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


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own file (oracle/_ref) or, when the copy is absent, the torch-op port
# ---------------------------------------------------------------------------------------------------------------------
class CpuArm:
    """Times `GIP_retrieval` / `IP_retrieval` exactly as the reference's main() prepares its inputs (fp32 CPU tensors,
    gip_retrieval.py:275,313) on a bounded sample of `rows` rows; the algorithm is exactly linear in N (one pass per query,
    :115), so q/s at n_total rows = measured q/s * rows / n_total.  G > 1 workloads hand the reference one idx per value
    column (np.repeat) so that :119 computes the grouped sum."""

    def __init__(self, workload, rows, n_queries, topk):
        import torch
        from dhr_b200 import synth
        from oracle import gip_oracle as go
        from oracle import refshim
        self.torch, self.go = torch, go
        self.cfg = cfg = synth.CONFIGS[workload]
        if refshim.copy_available():
            self.kind, self.ref = 'reference', refshim.load_copy()
            self.what = 'castorini/dhr retrieval/gip_retrieval.py (verbatim copy under oracle/_ref), GIP_retrieval / IP_retrieval'
        else:
            self.kind, self.ref = 'port', None
            self.what = 'torch-op port of gip_retrieval.py:110-126 (oracle/_ref copy absent)'
        cv, ci = synth.corpus_numpy(workload, 0, rows)
        qv, qi = synth.queries_numpy(workload, n_queries)
        G = cfg['G']
        self.rows = rows
        self.c = torch.from_numpy(cv.astype(np.float32))              # :313 the CPU path works on fp32 copies
        self.q = torch.from_numpy(qv.astype(np.float32))
        self.qids = list(range(n_queries))
        self.k = min(topk, rows)
        if cfg['S'] > 0:
            self.cidx = torch.from_numpy(np.repeat(ci, G, axis=1).astype(np.int16))
            self.qidx = torch.from_numpy(np.repeat(qi, G, axis=1).astype(np.int16))

    def run(self, n=None, rows=None, threads=None):
        """seconds for the first n sample queries over the first `rows` sample rows with `threads` torch threads"""
        torch, go, cfg = self.torch, self.go, self.cfg
        n = len(self.qids) if n is None else min(n, len(self.qids))
        rows = self.rows if rows is None else min(rows, self.rows)
        if threads:
            torch.set_num_threads(threads)
        c = self.c[:rows]
        k = min(self.k, rows)
        sink = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            if cfg['S'] > 0:
                args = go.make_args(emb_dim=cfg['S'] * cfg['G'], topk=k, brute_force=True)
                fn = self.ref.GIP_retrieval if self.ref else go.GIP_retrieval_port
                fn(self.qids[:n], self.q[:n], self.qidx[:n], c, self.cidx[:rows], args)
            else:
                fn = self.ref.IP_retrieval if self.ref else go.IP_retrieval_port
                fn(self.qids[:n], self.q[:n], c, go.make_args(topk=k))
        return time.perf_counter() - t0


def workload_string(workload, n_total, n_q, k):
    from dhr_b200 import synth
    cfg = synth.CONFIGS[workload]
    return '%s: %d passages, %d queries, S=%d x G=%d lexical (%s idx) + %d dense fp16, top-%d' % (
        workload, n_total, n_q, cfg['S'], cfg['G'], cfg['idx'], cfg['C'], k)


def cpu_baseline_block(args, n_total, k, threads):
    """bounded sample (about 10-30 s of CPU work): all cores at two sizes (linearity), one thread as shipped (:259)"""
    arm = CpuArm(args.workload, min(args.cpu_rows, n_total), args.cpu_queries, k)
    rows = arm.rows
    arm.run(1, threads=threads)                                       # warm-up
    t_all = arm.run(threads=threads)
    t_half = arm.run(rows=rows // 2, threads=threads)
    n1 = max(1, min(2, args.cpu_queries))
    t_one = arm.run(n1, threads=1)
    nq = args.cpu_queries
    v = nq / t_all * rows / n_total
    return {
        'value': v, 'unit': 'queries/s', 'cores': threads, 'kind': arm.kind,
        'sample': '%d queries x %d rows (%.1f s), %s, torch.set_num_threads(%d), scaled linearly to %d rows' % (
            nq, rows, t_all, arm.what, threads, n_total),
        'one_thread': {'value': n1 / t_one * rows / n_total, 'unit': 'queries/s', 'cores': 1,
                       'sample': '%d queries x %d rows (%.1f s), as shipped: torch.set_num_threads(1) (gip_retrieval.py:259)' % (n1, rows, t_one)},
        'linearity': {'rows': [rows // 2, rows], 's_per_query': [t_half / nq, t_all / nq],
                      'ratio': (t_all / nq) / max(1e-12, t_half / nq)},
    }


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref copy of gip_retrieval.py; the
    torch-op port if the copy is absent) on the host cores.  Each step is a bounded sample of the workload; --cpu-full runs
    SURVEY 8(d)'s whole procedure once (T=16 shards through the reference's range-sharding rule, >= 20 queries per shard,
    one thread as shipped and all cores, plus a 1 M-row linearity point)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from dhr_b200 import synth
    import torch
    n_total = args.rows or synth.N_MSMARCO
    threads = os.cpu_count() or 1
    n_q = args.queries or synth.Q_MSMARCO
    extra = {}
    if args.cpu_full:
        per = n_total // 16
        nq = max(20, args.cpu_queries)
        arm = CpuArm(args.workload, max(per, min(1000000, n_total)), nq, args.topk)
        arm.run(1, rows=per, threads=threads)
        shard_all = [arm.run(rows=per, threads=threads) for _ in range(2)]
        shard_one = arm.run(4, rows=per, threads=1)
        lin = arm.run(4, rows=min(1000000, n_total), threads=threads)
        # every shard holds statistically identical rows, so 16 shard passes = 16 x one measured shard pass
        extra = {'procedure': 'SURVEY 8(d): T=16 shards of %d rows (gip_retrieval.py:292-306), %d queries per shard; per-query time = 16 x shard time' % (per, nq),
                 'all_cores': {'cores': threads, 's_per_query_mean': 16 * float(np.mean(shard_all)) / nq, 's_per_query_min': 16 * min(shard_all) / nq,
                               'queries_per_s': nq / (16 * float(np.mean(shard_all)))},
                 'one_thread': {'cores': 1, 's_per_query': 16 * shard_one / 4, 'queries_per_s': 4 / (16 * shard_one)},
                 'linearity_1m': {'rows': min(1000000, n_total), 's_per_query': lin / 4, 'shard_rows': per,
                                  's_per_query_shard': float(np.mean(shard_all)) / nq,
                                  'ratio_time': (lin / 4) / (float(np.mean(shard_all)) / nq), 'ratio_rows': min(1000000, n_total) / per}}
        rows, nqs = per, nq
    else:
        rows, nqs = min(args.cpu_rows, n_total), args.cpu_queries
        arm = CpuArm(args.workload, rows, nqs, args.topk)
    times = []
    for i in range(args.warmup + args.steps):
        dt = arm.run(rows=rows, threads=threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    qps = nqs / (ms / 1e3) * rows / n_total
    sample = '%d queries x %d rows per step, %s, torch %s CPU, %d threads, scaled linearly to %d rows' % (
        nqs, rows, arm.what, torch.__version__, threads, n_total)
    line = {
        'impl': 'reference', 'metric': 'queries/sec', 'value': qps, 'unit': 'queries/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_string(args.workload, n_total, n_q, args.topk)},
        'cpu_baseline': {'value': qps, 'unit': 'queries/s', 'cores': threads, 'kind': arm.kind, 'sample': sample},
        'e2e': {'value': qps, 'unit': 'queries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    if extra:
        line['cpu_full'] = extra
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# verification at the benchmarked size: plain torch fp32 re-scoring of sample queries over the whole (shard of the) corpus
# ---------------------------------------------------------------------------------------------------------------------
def torch_reference_scores(workload, lo, hi, qv, qi, dev, masked=True):
    """fp32 scores [n_sample, hi - lo] by the reference's operator sequence (gip_retrieval.py:119-120): eq-mask * values,
    row dot; corpus rows regenerated segment by segment from the same seeds as the index build."""
    import torch
    from dhr_b200 import synth
    cfg = synth.CONFIGS[workload]
    S, G, C = cfg['S'], cfg['G'], cfg['C']
    n = qv.shape[0]
    out = torch.empty((n, hi - lo), dtype=torch.float32, device=dev)
    q = qv.to(torch.float32)
    q_lex = q[:, :S * G].reshape(n, S, G) if S > 0 else None
    q_dns = q[:, S * G:] if C > 0 else None
    q_idx = qi.to(torch.int32) if S > 0 else None
    pos = 0
    for vals, idx in synth.corpus_torch_segments(workload, lo, hi, dev):
        m = vals.shape[0]
        v = vals.to(torch.float32)
        sc = torch.zeros((n, m), dtype=torch.float32, device=dev)
        if C > 0:
            sc += q_dns @ v[:, S * G:].T
        if S > 0:
            c_lex = v[:, :S * G].reshape(m, S, G)
            c_idx = (idx.view(torch.int16).to(torch.int32) & 0xFFFF) if idx.dtype == torch.uint16 else idx.to(torch.int32)
            for i in range(n):
                per_slice = (c_lex * q_lex[i]).sum(dim=2)                             # [m, S] grouped inner products
                sc[i] += (per_slice * (c_idx == q_idx[i])).sum(dim=1) if masked else per_slice.sum(dim=1)
        out[:, pos:pos + m] = sc
        pos += m
    return out


def verify_results(workload, lo, hi, sample, qv, qi, res_scores, res_rows, k, dev, world, dist, tol=1e-3, eps=2e-5, masked=True):
    """res_* [n_sample, k]: the benchmarked answer (global rows) for the sample queries.  Every rank checks the rows of its
    shard [lo, hi); counts are summed over ranks."""
    import torch
    ref = torch_reference_scores(workload, lo, hi, qv, qi, dev, masked)
    n = ref.shape[0]
    rows = res_rows.to(dev)
    scores = res_scores.to(dev)
    mine = (rows >= lo) & (rows < hi)
    local = (rows - lo).clamp(0, hi - lo - 1)
    ref_at = torch.gather(ref, 1, local)
    err = torch.where(mine, (ref_at - scores).abs(), torch.zeros_like(scores))
    max_err = float(err.max().item()) if err.numel() else 0.0
    # completeness: no row outside the answer may beat the k-th returned score by more than fp32 reorder noise (the torch
    # re-scoring sums in another order: the noise is what the returned rows themselves show, floor 2e-5)
    eps = max(eps, 2.0 * max_err)
    kth = scores[:, -1:].clone()
    covered = ref.clone()
    covered.scatter_(1, local, torch.where(mine, torch.full_like(scores, -float('inf')), torch.gather(ref, 1, local)))
    missed = int((covered > kth + eps).sum().item())
    n_mine = int(mine.sum().item())
    cnt = torch.tensor([n_mine, missed], dtype=torch.int64, device=dev)
    mx = torch.tensor([max_err], dtype=torch.float32, device=dev)
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    s = res_scores.cpu().numpy().astype(np.float64)
    r = res_rows.cpu().numpy()
    order_ok = bool(np.all((s[:, :-1] > s[:, 1:]) | ((s[:, :-1] == s[:, 1:]) & (r[:, :-1] < r[:, 1:]))))
    unique_ok = all(len(set(row.tolist())) == k for row in r)
    out = {'queries': n, 'sample': [int(x) for x in sample], 'rows_checked': int(cnt[0].item()), 'rows_expected': n * k,
           'max_abs_score_err': float(mx[0].item()), 'tolerance': tol, 'missed_rows': int(cnt[1].item()), 'near_tie_eps': eps,
           'order_score_desc_row_asc': order_ok, 'rows_unique': unique_ok,
           'against': 'torch fp32 eq-mask * values row-dot over all %d rows (gip_retrieval.py:119-120) on the same device' % (hi - lo)}
    out['ok'] = bool(out['rows_checked'] == n * k and out['max_abs_score_err'] <= tol and out['missed_rows'] == 0 and order_ok and unique_ok)
    return out


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dhr_b200 import GipIndex, synth
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
    torch.backends.cuda.matmul.allow_tf32 = False

    cfg = synth.CONFIGS[args.workload]
    n_total = args.rows or synth.N_MSMARCO
    n_q = args.queries or synth.Q_MSMARCO
    k = args.topk
    lo, hi = shard_bounds(n_total, world, rank)

    # ---- build the resident shard (not timed: index load is outside the metric) ----
    t_build = time.perf_counter()
    ix = GipIndex(cfg['S'], cfg['C'], cfg['G'], capacity=hi - lo, idx_dtype=np.dtype(cfg['idx']), device=local_rank, row_offset=lo, lex_postings=args.lex_postings)
    for vals, idx in synth.corpus_torch_segments(args.workload, lo, hi, dev):
        ix.append(vals, idx)
    ix.finalize()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    for name, val in (('query_block', args.query_block), ('query_groups', args.query_groups), ('scan_variant', args.scan_variant),
                      ('overlap', args.overlap), ('dense_variant', args.dense_variant), ('dense_multicast', args.dense_multicast)):
        if val is not None:
            ix.set_option(name, val)
    for kv in args.option:
        name, val = kv.split('=')
        ix.set_option(name, int(val))
    ix.set_option('profile', 1)

    qv_dev, qi_dev = synth.queries_torch(args.workload, n_q, dev)
    qv_host = qv_dev.cpu().pin_memory()
    qi_host = qi_dev.cpu().pin_memory() if qi_dev is not None else None
    out_dev = (torch.empty((n_q, k), dtype=torch.float32, device=dev), torch.empty((n_q, k), dtype=torch.int64, device=dev),
               torch.empty((n_q,), dtype=torch.int32, device=dev))
    out_host = (torch.empty((n_q, k), dtype=torch.float32).pin_memory(), torch.empty((n_q, k), dtype=torch.int64).pin_memory(),
                torch.empty((n_q,), dtype=torch.int32).pin_memory())
    searcher = None
    if world > 1:
        from dhr_b200.distributed import ShardedSearcher
        searcher = ShardedSearcher(ix, n_q, k)
        searcher.profile = True

    def step(host_io):
        """one search of all queries over the (sharded) corpus; returns the final [Q,k] (scores, rows)"""
        qv, qi = (qv_host, qi_host) if host_io else (qv_dev, qi_dev)
        if world == 1:
            res = ix.search(qv, qi, k, masked=not args.unmasked, out=out_host if host_io else out_dev)[:2]
            st = ix.stats()
        else:
            # per-shard search enqueued as one stream-ordered call; per batch of 256 queries the packed keys are all-gathered
            # (NCCL) and merged on a side stream while the next batch is scanned
            res = searcher.search(qv, qi, k, masked=not args.unmasked, out=out_dev[:2])
            st = ix.stats()
            st['n_kernel_launches'] += searcher.n_merge_launches
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
    index_bytes = ix.device_bytes                                     # resident during the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, stats = timed(args.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    exch = dict(searcher.breakdown) if searcher is not None else None
    step(True)
    ms_e2e, _ = timed(args.steps, True)

    # ---- the answer, checked where it is benchmarked ----
    verified = None
    if not args.no_verify:
        res, _ = step(False)
        torch.cuda.synchronize()
        nv = min(args.verify_queries, n_q)
        sample = np.unique(np.linspace(0, n_q - 1, nv).astype(np.int64))
        st = torch.from_numpy(sample).to(dev)
        qi_s = qi_dev.view(torch.int16)[st].to(torch.int32) & 0xFFFF if (qi_dev is not None and qi_dev.dtype == torch.uint16) else \
            (qi_dev[st] if qi_dev is not None else None)                  # torch cannot index uint16 tensors on the device
        verified = verify_results(args.workload, lo, hi, sample, qv_dev[st], qi_s, res[0][st], res[1][st], k, dev, world, dist,
                                  masked=not args.unmasked)
    tile_first = None
    if world == 1 and not args.no_verify and not args.unmasked:
        n_s = min(8, n_q)
        tile_first = (out_dev[0][:n_s].clone(), out_dev[1][:n_s].clone())

    # HBM-bound operating point of the scan (K1, one query per corpus pass, one group per launch): bounded sample.  K1 reads the
    # row-major arrays, which are rebuilt from the tiled copies for this leg and dropped again afterwards.
    qb1 = None
    if world == 1 and not args.unmasked:
        n_s = min(8, n_q)
        ix.set_option('tile_mode', 0); ix.set_option('query_block', 1); ix.set_option('query_groups', 1); ix.set_option('scan_variant', 1)
        k1_out = tuple(torch.empty_like(o[:n_s]) for o in out_dev)
        ix.search(qv_dev[:n_s], qi_dev[:n_s] if qi_dev is not None else None, k, out=k1_out)
        ix.search(qv_dev[:n_s], qi_dev[:n_s] if qi_dev is not None else None, k, out=k1_out)
        s1 = ix.stats()
        qb1 = {'scan_ms': s1['scan_ms'], 'passes': s1['corpus_passes'], 'queries': n_s, 'total_ms': s1['total_ms']}
        torch.cuda.synchronize()
        if tile_first is not None and verified is not None:
            same_rows = float((tile_first[1] == k1_out[1]).float().mean().item())
            d = float((tile_first[0] - k1_out[0]).abs().max().item())
            verified['cross_kernel'] = {'queries': n_s, 'rows_equal_frac': same_rows, 'max_abs_score_diff': d,
                                        'what': 'tile path (K2 + K1t) vs row-scan kernel K1 on the same queries at full size; positions may '
                                                'differ only inside fp32-reorder near-ties'}
            verified['ok'] = bool(verified['ok'] and same_rows >= 0.99 and d <= 1e-3)
        ix.set_option('tile_mode', 1); ix.set_option('rowmajor', 0)

    # ---- roofline of the scan kernels, from CUDA events inside the library ----
    scan_ms = sum(s['scan_ms'] for s in stats)
    select_ms = sum(s['select_ms'] for s in stats)
    passes = sum(s['corpus_passes'] for s in stats)
    launches = sum(s['n_scan_launches'] for s in stats)
    alg_bytes = sum(s['alg_bytes'] for s in stats)
    dense_flops = sum(s['dense_flops'] for s in stats)
    bytes_per_pass = stats[0]['bytes_per_pass']
    variant = stats[0]['scan_variant']
    peaks, peak_src = load_peaks()
    traffic_all = {}
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic_all = json.load(f)
        except Exception:
            traffic_all = {}
    tw = traffic_all.get(args.workload) if isinstance(traffic_all.get(args.workload), dict) else None

    if rank == 0:
        qps = n_q * args.steps / (ms_dev / 1e3)
        qps_e2e = n_q * args.steps / (ms_e2e / 1e3)
        h2d = int(qv_host.numel() * qv_host.element_size() + (qi_host.numel() * qi_host.element_size() if qi_host is not None else 0))
        d2h = int(n_q * k * 12 + (n_q * 4 if world == 1 else 0))
        # two batch lanes run concurrently, so the per-launch scan / select times of the two lanes overlap; the roofline divides by
        # the device time of the whole search (first to last launch, CUDA events), which also charges the selects to the scan kernels
        total_search_ms = sum(s_['total_ms'] for s_ in stats)
        scan_s = (total_search_ms if total_search_ms > 0 else scan_ms) / 1e3
        hbm_achieved = alg_bytes / scan_s / 1e9 if scan_s > 0 else 0.0
        tensor_achieved = dense_flops / scan_s / 1e12 if scan_s > 0 else 0.0
        tensor_peak = peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops'))
        kernel_names = {1: 'gip_scan_tma (K1)', 0: 'gip_scan_direct (K1)', 2: 'dense_tile_ts (K2, tcgen05, queries in TMEM)',
                        3: ('lex_post (K1p, postings walk)' if stats[0].get('lex_layout') else 'lex_tile (K1t)') +
                           (' + dense_tile_ts (K2, tcgen05, queries in TMEM)' if cfg['C'] > 0 else ''),
                        4: 'dense_tile_ts column passes (K2, tcgen05): unmasked --IP stage'}
        if variant in (2, 4):      # dense-only / unmasked: a GEMM: a GEMM -> tensor ceiling (sustained: timed inside a long step)
            roof = {'bound': 'tensor', 'achieved': tensor_achieved, 'peak': tensor_peak, 'unit': 'TFLOP/s',
                    'frac': tensor_achieved / tensor_peak, 'peak_kind': 'bf16 sustained (fp16 runs at the same rate)',
                    'hbm': {'achieved': hbm_achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': hbm_achieved / peaks['hbm_gbs'],
                            'what': 'N x C_pad x 2 bytes per group of 128 queries / scan time'}}
        else:
            roof = {'bound': 'hbm', 'achieved': hbm_achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': hbm_achieved / peaks['hbm_gbs'],
                    'what': 'algorithmic bytes in DESIGN.md units (K1t: lexical bytes per tile of 64 queries; K2: dense bytes per 128 queries; '
                            'K1: row bytes per group) / device time of the whole search (selects included)'}
            if dense_flops > 0:
                roof['tensor'] = {'achieved': tensor_achieved, 'peak': tensor_peak, 'unit': 'TFLOP/s', 'frac': tensor_achieved / tensor_peak,
                                  'what': 'K2 flops (2*Q*N*C) over the device time of the whole search (K2 shares the SMs with K1t)'}
            if cfg['S'] > 0:
                # SURVEY 8(d) lexical lane-op count of the compare-everything formulation vs the CUDA-core ceiling; the tile walk is
                # O(matches), so it may exceed that "ceiling"
                lane_ops = float(n_q) * args.steps * (hi - lo) * (cfg['S'] + cfg['S'] * cfg['G'])
                lane_peak = 148 * 128 * 1.965e9
                roof['alu'] = {'achieved': lane_ops / scan_s if scan_s > 0 else 0.0, 'peak': lane_peak, 'unit': 'lane-ops/s',
                               'frac': lane_ops / scan_s / lane_peak if scan_s > 0 else 0.0,
                               'what': 'Q*N*S compare+select + Q*N*D FMA lane-ops (SURVEY 8d) / scan time vs 148 SMs x 128 lanes x 1.965 GHz; '
                                       'K1t walks only the matching (query, passage, slice) triples, so > 1 is possible'}
        if tw:
            roof['traffic'] = tw.get('dram_bytes_per_launch')
            roof['traffic_note'] = '%s over %d rows x %d queries in flight (ncu --set full, %s)' % (
                tw.get('launch_kind', ''), tw.get('rows_per_launch', 0), tw.get('queries_in_flight', 0), tw.get('source', ''))
            roof['ncu'] = {key: tw[key] for key in ('issue_active_pct', 'lsu_shared_wavefront_pct', 'tensor_pipe_pct', 'dram_pct',
                                                    'pred_on_threads_per_inst', 'dominant') if key in tw}
            if tw.get('dram_bytes_per_launch') and tw.get('rows_per_launch'):
                batches = -(-n_q // tw.get('queries_in_flight', 256))
                per_step = tw['dram_bytes_per_launch'] / tw['rows_per_launch'] * (hi - lo) * batches
                phys = per_step * args.steps / scan_s / 1e9 if scan_s > 0 else 0.0
                roof['physical_dram'] = {'achieved': phys, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': phys / peaks['hbm_gbs'],
                                         'what': 'ncu dram__bytes per sub-chunk launch / rows per launch x rows x query batches / scan time'}
        else:
            roof['traffic'] = None
        roof.update({'peak_source': peak_src, 'kernel': kernel_names.get(variant, '?'), 'queries_per_pass': stats[0]['query_block'],
                     'bytes_per_launch': alg_bytes / max(1, launches), 'launch_ms_avg': scan_ms / max(1, launches),
                     'corpus_passes_per_step': passes / args.steps, 'logical_pass_bytes': bytes_per_pass,
                     'search_device_ms_per_step': total_search_ms / args.steps,
                     'scan_stream_ms_per_step': scan_ms / args.steps, 'select_stream_ms_per_step': select_ms / args.steps,
                     'stream_note': 'scan / select = summed per-launch stream times; with two batch lanes they overlap and may exceed the step'})
        line = {
            'metric': 'queries/sec', 'value': qps, 'unit': 'queries/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32',  # fp16 storage, exact fp16 x fp16 products accumulated in fp32 (FHFMA / tcgen05 kind::f16)
            'data': 'synthetic',
            'config': {
                'workload': workload_string(args.workload, n_total, n_q, k) + (' [unmasked --IP first stage]' if args.unmasked else ''),
                'row_bytes': ix.row_bytes, 'corpus_bytes': ix.row_bytes * n_total, 'index_bytes': index_bytes,
                'parallelism': 'range-shard x%d' % world,
                'query_block': stats[0]['query_block'], 'query_groups': stats[0]['query_groups'], 'scan_variant': variant,
                'l2': 'inputs (%.1f GB per GPU) far exceed the 126 MB L2' % (ix.row_bytes * (hi - lo) / 1e9),
                'index_build_s': t_build,
            },
            'roofline': roof,
            'e2e': {'value': qps_e2e, 'unit': 'queries/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(sum(s['n_kernel_launches'] for s in stats)),
            'clocks': clocks,
            'fallback_queries': int(sum(s['n_fallback_queries'] for s in stats)),
        }
        if verified is not None:
            line['verified'] = verified
        if world > 1:
            step_ms = ms_dev / args.steps
            line['breakdown'] = {'step_ms': step_ms, 'scan_ms': scan_ms / args.steps, 'select_ms': select_ms / args.steps,
                                 'exchange_ms_overlapped': (exch or {}).get('exchange_ms'), 'exchange_tail_ms': (exch or {}).get('tail_ms'),
                                 'search_device_ms': total_search_ms / args.steps,
                                 'other_ms': step_ms - total_search_ms / args.steps,
                                 'what': 'rank 0, last timed step for the exchange: scan/select = summed per-launch stream times (two batch lanes overlap); search_device = first to last launch of the shard search; exchange = '
                                         'NCCL all-gather + key merge per 256-query batch on the side stream (hidden behind the scan); tail = '
                                         'part of the exchange after the last scan launch; other = prep, launch gaps, rank skew'}
        if qb1 and qb1['scan_ms'] > 0:
            a1 = qb1['passes'] * bytes_per_pass / (qb1['scan_ms'] / 1e3) / 1e9
            line['roofline_qb1'] = {'bound': 'hbm', 'achieved': a1, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': a1 / peaks['hbm_gbs'],
                                    'kernel': 'gip_scan_tma (K1), one query per corpus pass', 'queries': qb1['queries'],
                                    'queries_per_sec': qb1['queries'] / (qb1['total_ms'] / 1e3) if qb1['total_ms'] > 0 else None}
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_block(args, n_total, k, os.cpu_count() or 1)
        if world > 1:
            import ctypes
            ctypes.CDLL(None).fflush(None)                  # C stdio may still buffer the banner
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    ok = verified is None or verified['ok']
    ix.close()
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.stderr.write('bench.py: VERIFICATION FAILED: %s\n' % json.dumps(verified))
        sys.exit(3)


if __name__ == '__main__':
    main()
