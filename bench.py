"""bench.py - headline benchmark of the B200-native DoppelSpeller hot path.

Workloads (`--workload`, BASELINE.json configs):
  c3          configs[2], the headline: synthetic 100,000 test x 500,000 truth titles with the example data set's word /
              trigram statistics (doppelspeller_b200/synthetic.py), nearest-n = 10 IDF-weighted Jaccard top-n
  c3-example  the same size grown from the REAL example titles (tests/golden/example_titles.npz tiled with typing edits)
  c5          configs[4]: synthetic 1,000,000 x 10,000,000, truth rows sharded over the ranks (`--truth-shards N`)
A "step" is one pass of the hot path over the whole query batch: every (query, truth) pair scored, the reference's
selection rule applied, [Q, top_n] candidate rows produced.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c3-example|c5]

`value`   whole-job titles/s with the query CSR already resident in HBM (device-timed, CUDA events, max over ranks);
`e2e`     the same through the public API with pinned HOST buffers (H2D of the queries and D2H of the candidate rows
          inside the timed region).
N > 1     (torchrun) 2-D layout: T truth shards (local scan with thresholds shared over NVLink peer memory -> all_gather ->
          merge, doppelspeller_b200/sharded.py) x N / T independent query groups; T = 1 for c3 (the index is 40 MB: every
          rank holds it, the queries are split), T = N for c5, --truth-shards picks T.  Total work is fixed -> "scaling":
          "strong".  Rank 0 checks a query sample of the gathered result against the CPU oracle at every N.
`extra`   (N = 1) the other BASELINE configs in the same record: c3-example, C4 (100M candidate pairs through the
          InDel-ratio and the 66-feature kernels), the candidate pairs of the step, C1 / C2 (example data through the
          drop-in classes, the reference's own numba path timed beside them).
--impl reference   times the reference's own CPU implementation of the path on the host cores: the staged, unmodified
          reference's numba kernels (oracle/_ref: match_maker.fast_jaccard + fast_arg_top_k driven like
          get_closest_matches, match_maker.py:192-203) when staged, else the C / OpenMP port (oracle/ds_oracle.c);
          each step a bounded query sample of the same workload; rank 0 only.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = 'titles/s'
WORKLOADS = {
    'c3': dict(queries=100000, truth=500000, source='synthetic', tag='C3 (BASELINE.json configs[2])'),
    'c3-example': dict(queries=100000, truth=500000, source='example', tag='C3 size, real example titles tiled'),
    'c5': dict(queries=1000000, truth=10000000, source='synthetic', tag='C5 (BASELINE.json configs[4])'),
}
SMEM_PEAK_BYTES_PER_CLK_PER_SM = 128     # shared-memory crossbar (B300_MICROARCH.md, LDS/STS)
N_SM = 148

_STDOUT = None


def emit(line):
    out = _STDOUT if _STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def log(*args):
    print(*args, file=sys.stderr, flush=True)


def metric_name(n_truth):
    if n_truth == 500000:
        return 'titles matched/sec vs 500k truth'
    return f'titles matched/sec vs {n_truth} truth'


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
def make_titles(source, n_queries, n_truth):
    """(truth titles, test titles): deterministic, identical on every rank."""
    from doppelspeller_b200 import synthetic
    if source == 'example':
        raw = np.load(os.path.join(ROOT, 'tests', 'golden', 'example_titles.npz'))
        truth = synthetic.tile_titles([str(t) for t in raw['truth_titles']], n_truth, synthetic.TRUTH_SEED)
        test = synthetic.tile_titles([str(t) for t in raw['test_titles']], n_queries, synthetic.TEST_SEED)
        return truth, test
    truth = synthetic.generate_truth_titles(n_truth)
    test, _ = synthetic.generate_test_titles(truth, n_queries)
    return truth, test


def encode_host(test, truth):
    from doppelspeller_b200 import encode
    return encode.encode_canonical(test, truth)


def oracle_index(enc, queries=None):
    """The CPU oracle's index of the workload (`queries`: only the posting lists those query rows touch - C5's 220M
    postings would take minutes to sort in numpy)."""
    from oracle import oracle
    return oracle.finish_index(dict(
        n_truth=int(enc['t_ptr'].shape[0]) - 1, w64=enc['idf64'], w32=enc['idf64'].astype(np.float32), t_ptr=enc['t_ptr'],
        t_cols=enc['t_cols'] if queries is not None else enc['t_cols'].astype(np.int32), q_ptr=enc['q_ptr'],
        q_cols=enc['q_cols'].astype(np.int32)), queries=queries)


def cpu_sample(n_queries, size):
    rng = np.random.default_rng(20240504)
    return np.sort(rng.choice(n_queries, size=min(size, n_queries), replace=False))


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_cpu_port(index, sample, k):
    """The CPU port of match_maker.py:192-203 (oracle/ds_oracle.c, OpenMP over queries) on `sample`."""
    from oracle import oracle
    t0 = time.perf_counter()
    rows, count, _ = oracle.topn(index, k, queries=sample, n_threads=host_threads())
    return time.perf_counter() - t0, rows, count


class ReferenceKernels:
    """The staged, unmodified reference's own numba kernels behind the data structures MatchMaker.__init__ builds
    (match_maker.py:111-133), filled from the same encoded index the GPU arm uses."""

    def __init__(self, index):
        import numba
        from oracle import ref_import
        ref = ref_import.import_match_maker()
        self.fast_jaccard = ref.match_maker.fast_jaccard
        self.fast_arg_top_k = ref.match_maker.fast_arg_top_k
        self.threads = min(host_threads(), numba.config.NUMBA_NUM_THREADS)
        numba.set_num_threads(self.threads)
        self.index = index
        self.n_truth = int(index['n_truth'])
        self.w64 = [float(x) for x in index['w64']]
        post_ptr, post_rows, w32 = index['post_ptr'], index['post_rows'], index['w32']
        columns = numba.typed.List()                                            # match_maker.py:122-133
        for v in range(post_ptr.shape[0] - 1):
            rows = np.ascontiguousarray(post_rows[post_ptr[v]:post_ptr[v + 1]], dtype=np.int32)
            columns.append((rows, np.full(rows.shape[0], w32[v], dtype=np.float32)))
        self.columns = columns
        self.sums = np.ascontiguousarray(index['sums'], dtype=np.float32)

    def closest_rows(self, q, k):
        """MatchMaker.get_closest_matches (match_maker.py:192-203) up to the row indexes (the pandas title-id lookup of
        :190 is left out on both arms)."""
        nz = np.ascontiguousarray(self.index['qs_cols'][self.index['qs_ptr'][q]:self.index['qs_ptr'][q + 1]], dtype=np.int32)
        mx = sum([self.w64[c] for c in nz])                                     # :197
        scores = self.fast_jaccard(self.n_truth, mx, nz, self.columns, self.sums)   # :199
        return self.fast_arg_top_k(scores, k)                                   # :187

    def run(self, queries, k):
        out = np.full((len(queries), k), -1, dtype=np.int64)
        t0 = time.perf_counter()
        for i, q in enumerate(queries):
            rows = self.closest_rows(int(q), k)
            out[i, :rows.shape[0]] = rows
        return time.perf_counter() - t0, out


def reference_staged():
    try:
        from oracle import ref_import
        import numba  # noqa: F401
        return ref_import.reference_available()
    except Exception:
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.device)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, sm_max, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
            except ValueError:
                continue
            for name, value in zip(names, parts[5:9]):
                if value.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(sm_max) if sm_max else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        peaks = json.load(open(path))
        return float(peaks['hbm_gbs']), float(peaks.get('sm_max_mhz', 1965.0)), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 1965.0, 'fallback (B200_PROFILING.md)'


def source_id(*names):
    digest = hashlib.sha256()
    for name in names:
        with open(os.path.join(ROOT, 'doppelspeller_b200', 'csrc', name), 'rb') as f:
            digest.update(f.read())
    return digest.hexdigest()[:16]


def committed_ncu(kernel):
    """profiles/r2_<kernel>_ncu.json: counters of the kernel from the committed `ncu --set full` capture (written by
    profiles/summarize.py together with the hash of the kernel sources it was taken from)."""
    path = os.path.join(ROOT, 'profiles', f'r2_{kernel}_ncu.json')
    if not os.path.exists(path):
        return None
    return json.load(open(path))


def measured_limiter(kernel, sources=('ds_pairs.cu', 'ds_common.cuh')):
    """Compact form of the committed ncu summary of a secondary kernel: its busiest unit and how busy it was."""
    ncu = committed_ncu(kernel)
    if ncu is None:
        return None
    units = {'shared-memory pipe': ncu.get('smem_wavefronts_pct'), 'issue slots': ncu.get('issue_active_pct'),
             'hbm': ncu.get('dram_throughput_pct')}
    bound = max((name for name in units if units[name] is not None), key=lambda name: units[name])
    return {'kernel': kernel, 'bound': bound, 'frac': units[bound] / 100.0, 'issue_active_pct': ncu.get('issue_active_pct'),
            'smem_wavefronts_pct': ncu.get('smem_wavefronts_pct'), 'dram_throughput_pct': ncu.get('dram_throughput_pct'),
            'source': ncu.get('source'), 'ncu_matches_current_sources': ncu.get('kernel_sources') == source_id(*sources)}


def workload_config(args, stats):
    spec = WORKLOADS.get(args.workload, {})
    tag = spec.get('tag', 'custom size')
    source = 'real example titles (tests/golden/example_titles.npz) tiled with typing edits' if args.source == 'example' else 'synthetic'
    cfg = {'workload': f'{tag}: {source} {args.queries} test x {args.truth} truth titles, nearest-n={args.top_n} '
                       f'IDF-weighted trigram Jaccard top-n',
           'queries': args.queries, 'truth': args.truth, 'top_n': args.top_n,
           'l2': 'flushed between timed steps (256 MiB memset, untimed)',
           'layout': f'{args.n_shards} truth shard(s) (contiguous row ranges; all_gather + merge inside each group) x '
                     f'{args.n_groups} query group(s)',
           'shared_thresholds': bool(args.n_shards > 1 and not getattr(args, 'no_share_thresholds', False))}
    cfg.update(stats)
    return cfg


def resolve_layout(args, world):
    # Truth rows are sharded no further than the data asks for: C3's index is 40 MB, so every rank holds all of it and the
    # ranks split the QUERIES (independent units, no data-path collective); C5 (BASELINE configs[4]: "truth sharded across
    # 2/4/8 B200 with NCCL top-n merge") shards the truth rows over all ranks.  --truth-shards T picks any T x (N / T) layout.
    args.n_shards = args.truth_shards if args.truth_shards > 0 else (world if args.workload == 'c5' else 1)
    if world % args.n_shards != 0:
        raise SystemExit(f'--truth-shards {args.n_shards} must divide the number of ranks {world}')
    args.n_groups = world // args.n_shards


# ---------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from doppelspeller_b200 import synthetic
    resolve_layout(args, max(1, args.gpus))
    truth, test = make_titles(args.source, args.queries, args.truth)
    enc = encode_host(test, truth)
    stats = synthetic.workload_statistics(enc)
    index = oracle_index(enc)
    k = args.top_n
    cores = host_threads()
    budget_s = 100.0 / max(1, args.steps + args.warmup)       # the whole run stays within a couple of minutes
    line = {'impl': 'reference', 'metric': metric_name(args.truth), 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic' if args.source == 'synthetic' else 'example titles tiled', 'config': workload_config(args, stats),
            'gpu_launches': 0}
    port_sample = cpu_sample(args.queries, args.cpu_sample)
    time_cpu_port(index, port_sample[:64], k)
    port_s, port_rows, _ = time_cpu_port(index, port_sample, k)
    port = {'value': len(port_sample) / port_s, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{len(port_sample)} sampled queries x all {args.truth} truth rows, top-{k}'}
    if reference_staged() and not args.port_only:
        kernels = ReferenceKernels(index)
        probe = cpu_sample(args.queries, 24)
        kernels.run(probe[:4], k)                                # JIT compile (excluded, like SURVEY.md 8(d) says)
        probe_s, _ = kernels.run(probe, k)
        n_sample = int(max(16, min(len(port_sample), budget_s / (probe_s / len(probe)))))
        sample = port_sample[:n_sample]
        for _ in range(args.warmup):
            kernels.run(sample[:max(4, n_sample // 8)], k)
        times, rows = [], None
        for _ in range(args.steps):
            seconds, rows = kernels.run(sample, k)
            times.append(seconds)
        per_step = float(np.mean(times))
        value = n_sample / per_step
        line['cpu_baseline'] = {'value': value, 'unit': UNIT, 'cores': kernels.threads, 'kind': 'reference',
                                'sample': f'{n_sample} sampled queries x all {args.truth} truth rows per step, top-{k}: the staged '
                                          f'reference\'s numba fast_jaccard + fast_arg_top_k called like get_closest_matches '
                                          f'(match_maker.py:192-203, JIT warm); titles/s = sample / time (linear in Q)',
                                'port': port}
        line['parity'] = {'reference_vs_port_checked_queries': n_sample,
                          'reference_vs_port_mismatching_queries': int((rows != port_rows[:n_sample]).any(axis=1).sum())}
    else:
        sample = port_sample
        for _ in range(args.warmup):
            time_cpu_port(index, sample[:max(8, len(sample) // 16)], k)
        times = [time_cpu_port(index, sample, k)[0] for _ in range(args.steps)]
        per_step = float(np.mean(times))
        value = len(sample) / per_step
        line['cpu_baseline'] = dict(port, value=value)
    line['value'] = value
    # one step of the workload = all Q queries: per-step time extrapolated linearly in Q from the sample
    line['ms_per_step'] = args.queries / value * 1e3
    line['ms_per_sampled_step'] = per_step * 1e3
    line['e2e'] = {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from doppelspeller_b200 import _native as nat
    from doppelspeller_b200 import encode, sharded, synthetic
    from doppelspeller_b200.index import TruthIndex

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a GPU: doppelspeller_b200 has no CPU fallback')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    k = args.top_n
    n_q_total, n_truth = args.queries, args.truth
    # 2-D layout: the truth rows are sharded over T ranks (exchange step: all_gather + merge inside each group of T
    # ranks) and the queries are split over G = world / T such groups (independent, no collective between groups).
    resolve_layout(args, world)
    n_shards, n_groups = args.n_shards, args.n_groups
    group_id, shard_id = rank // n_shards, rank % n_shards
    subgroup = None
    if world > 1:
        for g in range(n_groups):
            handle = dist.new_group(ranks=list(range(g * n_shards, (g + 1) * n_shards)))
            if g == group_id:
                subgroup = handle
    q_offs = sharded.shard_offsets(n_q_total, n_groups)
    q0, q1 = int(q_offs[group_id]), int(q_offs[group_id + 1])
    n_q = q1 - q0
    offs = sharded.shard_offsets(n_truth, n_shards)

    t0 = time.time()
    truth, test = make_titles(args.source, n_q_total, n_truth)
    t1 = time.time()
    device_encoded = world > 1 or args.device_encode or n_truth > 2000000
    if device_encoded:
        # index build entirely on the GPU (csrc/ds_encode.cu): trigram sets, canonical column ids, df, idf
        dev_enc = encode.encode_canonical_device(test, truth, device=local_rank)
        torch.cuda.synchronize()
        ptr, cols = sharded.slice_truth_csr(dev_enc['t_ptr'], dev_enc['t_cols'], int(offs[shard_id]), int(offs[shard_id + 1]))
        idf64 = dev_enc['idf64']
        enc = {'q_ptr': dev_enc['q_ptr'].cpu().numpy(), 'q_cols': dev_enc['q_cols'].cpu().numpy(), 'idf64': idf64.cpu().numpy()}
        if rank == 0:       # host copy of the truth CSR: statistics and the oracle's parity sample
            enc['t_ptr'], enc['t_cols'] = dev_enc['t_ptr'].cpu().numpy(), dev_enc['t_cols'].cpu().numpy()
        del dev_enc
    else:
        enc = encode_host(test, truth)
        ptr, cols = sharded.slice_truth_csr(enc['t_ptr'], enc['t_cols'], int(offs[shard_id]), int(offs[shard_id + 1]))
        idf64 = enc['idf64']
    log(f'[bench] rank {rank}: titles in {t1 - t0:.1f}s, encoded ({"GPU" if device_encoded else "host"}) in {time.time() - t1:.1f}s, '
        f'vocab {enc["idf64"].shape[0]}')
    stats = synthetic.workload_statistics(enc) if rank == 0 else {}
    t0 = time.time()
    index = TruthIndex(ptr, cols, idf64, device=local_rank, row_offset=int(offs[shard_id]), n_total=n_truth)
    torch.cuda.synchronize()
    log(f'[bench] rank {rank}: truth rows [{offs[shard_id]}, {offs[shard_id + 1]}) x queries [{q0}, {q1}), index built in '
        f'{time.time() - t0:.2f}s')
    del ptr, cols

    # this rank's query group
    my_q_ptr = np.ascontiguousarray(enc['q_ptr'][q0:q1 + 1] - enc['q_ptr'][q0])
    my_q_cols = np.ascontiguousarray(enc['q_cols'][enc['q_ptr'][q0]:enc['q_ptr'][q1]])
    d_q_ptr = torch.as_tensor(my_q_ptr).to(device)
    d_q_cols = torch.as_tensor(my_q_cols).to(device)
    h_q_ptr = torch.as_tensor(my_q_ptr).pin_memory()
    h_q_cols = torch.as_tensor(my_q_cols).pin_memory()
    h_rows = torch.empty((n_q, k), dtype=torch.int64).pin_memory()
    h_count = torch.empty((n_q,), dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    share = n_shards > 1 and not args.no_share_thresholds
    shard = sharded.GpuShard(index, my_q_ptr, my_q_cols, group=subgroup, share_thresholds=share) if n_shards > 1 else None
    phase_ms = {} if os.environ.get('DS_PHASE_TIMING') else None

    def step_device():
        if n_shards > 1:
            rows, count, _ = sharded.sharded_topn(shard, k, group=subgroup, timings=phase_ms)
            return rows, count
        return index.topn(d_q_ptr, d_q_cols, k)

    def step_e2e():
        if n_shards > 1:
            sh = sharded.GpuShard(index, h_q_ptr, h_q_cols, group=subgroup, share_thresholds=share)   # H2D of the queries
            rows, count, _ = sharded.sharded_topn(sh, k, group=subgroup)
            h_rows.copy_(rows, non_blocking=True)                                 # D2H of the result
            h_count.copy_(count, non_blocking=True)
            torch.cuda.synchronize()
            return h_rows, h_count
        return index.topn(h_q_ptr, h_q_cols, k, out_rows=h_rows, out_count=h_count)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        step_ms = []
        result = None
        for _ in range(steps):
            flush.zero_()
            barrier()
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            result = fn()
            stop.record()
            stop.synchronize()
            step_ms.append(start.elapsed_time(stop))
        barrier()
        # every step starts at a barrier, so a step lasts as long as its slowest rank: max over ranks PER STEP, then the sum
        # (the max of the ranks' sums would hide the steps in which different ranks were the slow one)
        t = torch.tensor(step_ms if step_ms else [0.0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.sum().item()), result

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches_before = nat.kernel_launches()
    nat.profile_begin()
    total_ms, (rows, count) = timed(step_device, args.steps)
    k1_profile = nat.profile_end_split()
    gpu_launches = nat.kernel_launches() - launches_before
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = n_q_total / (ms_per_step / 1e3)
    if phase_ms:
        log(f'[bench] rank {rank} phase ms per step: ' + ', '.join(f'{k_}={v / (args.steps + args.warmup):.2f}' for k_, v in phase_ms.items()))

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_ms, (e_rows, e_count) = timed(step_e2e, args.steps)
    e2e_value = n_q_total / (e2e_ms / args.steps / 1e3)
    # bytes copied per step summed over all ranks: every rank uploads its group's queries and reads back its rows
    h2d = int(n_shards * (enc['q_ptr'].nbytes + enc['q_cols'].nbytes))
    d2h = int(n_shards * (n_q_total * k * 8 + n_q_total * 4))

    # the result of every query group on rank 0 (parity sample below): one padded all_gather over all ranks, the first
    # rank of every group carries that group's rows
    rows_all = rows
    if world > 1:
        max_q = int(np.diff(q_offs).max())
        padded = torch.full((max_q, k), -1, dtype=torch.int64, device=device)
        padded[:n_q] = rows
        flat = torch.empty((world * max_q, k), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(flat, padded)
        if rank == 0:
            rows_all = torch.cat([flat[g * n_shards * max_q: g * n_shards * max_q + int(q_offs[g + 1] - q_offs[g])]
                                  for g in range(n_groups)], dim=0)
        del flat, padded
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, sm_mhz_max, peak_source = measured_peaks()
    mean_g = stats['mean_trigrams_per_truth_title']
    bytes_per_pair = 2.0 * mean_g + 8.0
    dominant = max(k1_profile, key=lambda name: k1_profile[name][0])
    scan_ms, scan_launches, scan_pairs = k1_profile[dominant]
    convention = scan_pairs * bytes_per_pair / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0
    roofline = build_roofline(dominant, scan_ms, scan_launches, total_ms, k1_profile, args.steps, convention, hbm_peak, peak_source,
                              bytes_per_pair, sm_mhz_max, clocks)

    line = {
        'metric': metric_name(n_truth), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic' if args.source == 'synthetic' else 'example titles tiled', 'config': workload_config(args, stats),
        'roofline': roofline,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_ms / args.steps},
        'gpu_launches': int(gpu_launches), 'clocks': clocks,
    }

    rows_np = rows_all.cpu().numpy() if hasattr(rows_all, 'cpu') else rows_all
    rows_mine = rows.cpu().numpy() if hasattr(rows, 'cpu') else rows
    e_rows_np = e_rows.numpy() if hasattr(e_rows, 'numpy') else e_rows
    line['parity'] = {'e2e_equals_device_path': bool(np.array_equal(rows_mine, e_rows_np))}
    if not args.no_cpu:
        # the parity gate of the run: a query sample of the (gathered, at N > 1) result against the CPU oracle
        big = n_truth > 2000000
        sample = cpu_sample(n_q_total, args.cpu_sample if not big else min(args.cpu_sample, 2000))
        index_cpu = oracle_index(enc, queries=sample if big else None)
        time_cpu_port(index_cpu, sample[:64], k)
        seconds, want_rows, want_count = time_cpu_port(index_cpu, sample, k)
        port = {'value': len(sample) / seconds, 'unit': UNIT, 'cores': host_threads(), 'kind': 'port',
                'sample': f'{len(sample)} sampled queries x all {n_truth} truth rows, top-{k}, {seconds:.1f}s; titles/s = sample / '
                          f'time (linear in Q)'}
        line['parity'].update({'checked_queries': int(len(sample)), 'checked_against': 'oracle/ds_oracle.c (CPU port)',
                               'mismatching_queries': int((rows_np[sample] != want_rows).any(axis=1).sum())})
        line['cpu_baseline'] = port
        if world == 1 and reference_staged() and not args.no_extra:
            try:
                kernels = ReferenceKernels(index_cpu)
                ref_sample = sample[:200 if n_truth <= 2000000 else 24]
                kernels.run(ref_sample[:4], k)
                ref_s, ref_rows = kernels.run(ref_sample, k)
                line['cpu_baseline'] = {
                    'value': len(ref_sample) / ref_s, 'unit': UNIT, 'cores': kernels.threads, 'kind': 'reference',
                    'sample': f'{len(ref_sample)} sampled queries x all {n_truth} truth rows, top-{k}, {ref_s:.1f}s: the staged '
                              f'reference\'s own numba fast_jaccard + fast_arg_top_k (match_maker.py:192-203, JIT warm)',
                    'port': port}
                line['parity']['reference_numba_checked_queries'] = int(len(ref_sample))
                line['parity']['reference_numba_mismatching_queries'] = int((rows_np[ref_sample] != ref_rows).any(axis=1).sum())
                del kernels
            except Exception as error:   # the port stays the baseline
                line['cpu_baseline']['reference_unavailable'] = repr(error)[:200]
        del index_cpu
    if world == 1 and not args.no_extra:
        extra = {}
        index.close()
        del index
        torch.cuda.empty_cache()
        nat.check(nat.lib.ds_trim(local_rank))
        for name, fn in (('candidate_pairs', lambda: pair_kernels(truth, test, rows_mine, device)),
                         ('c3_example', lambda: second_workload(args, device)),
                         ('c4', lambda: c4_pairs(device)),
                         ('c1_c2', lambda: example_dropin(device))):
            t0 = time.time()
            try:
                extra[name] = fn()
            except Exception as error:
                extra[name] = {'error': repr(error)[:300]}
            extra[name]['wall_s'] = round(time.time() - t0, 1) if isinstance(extra[name], dict) else None
            torch.cuda.empty_cache()
            nat.check(nat.lib.ds_trim(local_rank))
        line['extra'] = extra
    if world > 1:
        dist.destroy_process_group()
    emit(line)


def build_roofline(dominant, scan_ms, scan_launches, total_ms, k1_profile, steps, convention, hbm_peak, peak_source, bytes_per_pair,
                   sm_mhz_max, clocks):
    """The dominant kernel against its MEASURED limiter.  The K1 kernels walk an L2-resident index (0.02 % DRAM
    utilisation), so the limiter is taken from the committed `ncu --set full` capture of the current kernel sources
    (profiles/r2_<kernel>_ncu.json): the busiest of shared-memory pipe / issue slots / DRAM.  The SURVEY.md 8(d) byte
    convention (what a dense scan of the CSR rows would read) is kept as a labelled secondary figure."""
    ncu = committed_ncu(dominant)
    sm_mhz = (clocks or {}).get('sm_mhz') or sm_mhz_max
    smem_peak = N_SM * SMEM_PEAK_BYTES_PER_CLK_PER_SM * sm_mhz * 1e6 / 1e9        # GB/s at the clock seen under load
    out = {'kernel': dominant, 'launches': int(scan_launches), 'avg_launch_ms': scan_ms / max(1, scan_launches),
           'kernel_share_of_step': scan_ms / total_ms if total_ms > 0 else None,
           'k1_kernels': {name: {'ms_per_step': ms / steps, 'launches_per_step': n / steps,
                                 'pair_share': pairs / max(1.0, sum(v[2] for v in k1_profile.values()))}
                          for name, (ms, n, pairs) in k1_profile.items()},
           'hbm_convention': {'what': 'SURVEY.md 8(d): (query, truth) pairs x (2 * mean trigrams + 8) B over the kernel time - the '
                                      'bytes a dense scan of the truth CSR would read; the index is L2 resident and the kernel only '
                                      'touches the postings that hit a query, so this exceeds the HBM peak by construction',
                              'achieved_gbs': convention, 'peak_gbs': hbm_peak, 'frac': convention / hbm_peak,
                              'algorithmic_bytes_per_pair': bytes_per_pair, 'peak_source': peak_source}}
    if ncu is None:
        out.update({'bound': 'hbm', 'achieved': convention, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': convention / hbm_peak,
                    'traffic': None, 'note': 'no committed ncu capture of this kernel: byte convention only'})
        return out
    limiters = {'shared-memory pipe': ncu.get('smem_wavefronts_pct'), 'issue slots': ncu.get('issue_active_pct'),
                'hbm': ncu.get('dram_throughput_pct')}
    bound = max((name for name in limiters if limiters[name] is not None), key=lambda name: limiters[name])
    frac = limiters[bound] / 100.0
    current = source_id('ds_topn.cu', 'ds_common.cuh')
    if bound == 'shared-memory pipe':
        peak, unit = smem_peak, 'GB/s'
    elif bound == 'hbm':
        peak, unit = hbm_peak, 'GB/s'
    else:
        peak, unit = N_SM * 4 * sm_mhz * 1e6 / 1e9, 'G warp-instructions/s'
    out.update({'bound': bound, 'achieved': frac * peak, 'peak': peak, 'unit': unit, 'frac': frac,
                'traffic': ncu.get('dram_bytes_per_launch'),
                'ncu': {k_: ncu.get(k_) for k_ in ('smem_wavefronts_pct', 'issue_active_pct', 'dram_throughput_pct', 'warps_active_pct',
                                                   'lsu_pipe_pct', 'l2_bytes_per_launch', 'dram_bytes_per_launch', 'captured_launch_ms',
                                                   'registers', 'workload')},
                'ncu_source': ncu.get('source'), 'ncu_kernel_sources': ncu.get('kernel_sources'),
                'ncu_matches_current_sources': ncu.get('kernel_sources') == current,
                'note': 'frac = utilisation of the measured limiter (ncu, profiles/); peak = that unit\'s peak at the SM clock '
                        'sampled during the timed region; launch times and shares are measured live with CUDA events'})
    return out


# ---------------------------------------------------------------------------------------------------
# extras (N = 1): the other BASELINE configs in the same record
# ---------------------------------------------------------------------------------------------------
def pair_kernels(truth, test, rows, device):
    """K2 InDel ratio and K3 66-feature kernels on the candidate pairs of the step (compact title tables resident in
    HBM), and the whole hot path as one GPU-resident pass."""
    import torch
    from doppelspeller_b200 import _native as nat
    from doppelspeller_b200 import feature_engineering as fe
    from doppelspeller_b200 import pipeline as pl
    n_q, k = rows.shape
    codes_a, off_a = fe.encode_titles(test)
    codes_b, off_b = fe.encode_titles(truth)
    counts = pl.truth_word_counts(truth)
    dev = lambda x: torch.as_tensor(x).to(device)   # noqa: E731
    d = dict(a=dev(codes_a), oa=dev(off_a), b=dev(codes_b), ob=dev(off_b), c=dev(counts.view(np.int32)),
             ia=dev(np.repeat(np.arange(n_q, dtype=np.int32), k)), ib=dev(rows.reshape(-1).astype(np.int32)))
    n = n_q * k
    ratio = torch.empty(n, dtype=torch.uint8, device=device)
    feats = torch.empty((n, 66), dtype=torch.float32, device=device)

    def run_ratio():
        nat.check(nat.lib.ds_indel_ratio_pairs(nat.ptr(d['a']), nat.ptr(d['oa']), len(test), nat.ptr(d['b']), nat.ptr(d['ob']),
                                               len(truth), nat.ptr(d['ia']), nat.ptr(d['ib']), n, nat.ptr(ratio), None,
                                               nat.current_stream()))

    def run_feats():
        fe.construct_features_pairs((d['a'], d['oa']), (d['b'], d['ob']), d['c'], d['ia'], d['ib'], fe.SPACE_CODE, len(truth),
                                    response=feats)

    out = {}
    la = np.diff(off_a)[np.repeat(np.arange(n_q), k)]
    lb = np.diff(off_b)[rows.reshape(-1)]
    hbm_peak = measured_peaks()[0]
    for name, fn, bytes_per_pair in (('indel_ratio', run_ratio, float((la + lb).mean()) + 3.0),
                                     ('construct_features', run_feats, float((la + lb).mean()) + 2.0 + 60.0 + 264.0)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(3):
            fn()
        stop.record()
        stop.synchronize()
        ms = start.elapsed_time(stop) / 3
        out[name] = {'pairs_per_s': n / (ms / 1e3), 'pairs': n, 'ms': ms, 'algorithmic_bytes_per_pair': bytes_per_pair,
                     'hbm_frac': n * bytes_per_pair / (ms / 1e3) / 1e9 / hbm_peak,
                     'measured_limiter': measured_limiter('k_indel_groups' if name == 'indel_ratio' else 'k_feature_words')}
    # the whole hot path as one GPU-resident pass (north_star target: "100k test titles matched, nearest-n Jaccard +
    # Levenshtein features, against 500k truth titles"): host title strings in, candidate rows + features out
    pipeline = pl.CandidatePipeline(truth, device=device.index)
    for _ in range(2):
        pipeline.run(test, k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p_rows, _, p_feats = pipeline.run(test, k)
    torch.cuda.synchronize()
    seconds = time.perf_counter() - t0
    out['pipeline'] = {'what': 'title strings -> GPU index build (both sides) -> top-n candidates -> 66 features per candidate pair',
                       'titles_per_s': n_q / seconds, 'ms': seconds * 1e3, 'candidate_pairs': int(p_feats.shape[0]),
                       'rows_equal_step': bool(np.array_equal(p_rows.cpu().numpy(), rows))}
    return out


def second_workload(args, device):
    """The headline measurement repeated on the workload grown from the real example titles (density is the data
    set's own, not the generator's)."""
    import torch
    from doppelspeller_b200 import synthetic
    from doppelspeller_b200.index import TruthIndex
    source = 'example' if args.source == 'synthetic' else 'synthetic'
    n_q, n_truth, k = args.queries, args.truth, args.top_n
    truth, test = make_titles(source, n_q, n_truth)
    enc = encode_host(test, truth)
    stats = synthetic.workload_statistics(enc)
    index = TruthIndex(enc['t_ptr'], enc['t_cols'], enc['idf64'], device=device.index)
    d_q_ptr, d_q_cols = torch.as_tensor(enc['q_ptr']).to(device), torch.as_tensor(enc['q_cols']).to(device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    for _ in range(3):
        rows, _ = index.topn(d_q_ptr, d_q_cols, k)
    times = []
    for _ in range(5):
        flush.zero_()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        rows, _ = index.topn(d_q_ptr, d_q_cols, k)
        stop.record()
        stop.synchronize()
        times.append(start.elapsed_time(stop))
    ms = float(np.mean(times))
    index_cpu = oracle_index(enc)
    sample = cpu_sample(n_q, 5000)
    _, want_rows, _ = time_cpu_port(index_cpu, sample, k)
    got = rows.cpu().numpy()
    index.close()
    return {'workload': ('real example titles tiled with typing edits' if source == 'example' else 'synthetic') +
            f', {n_q} x {n_truth}, top-{k}', 'value': n_q / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms, 'steps': 5, 'statistics': stats,
            'parity': {'checked_queries': int(len(sample)), 'mismatching_queries': int((got[sample] != want_rows).any(axis=1).sum())}}


def c4_pairs(device):
    """BASELINE configs[3]: 100M synthetic candidate pairs, titles up to 128 characters (bench_pairs.py)."""
    import bench_pairs
    out = bench_pairs.run(pairs=100_000_000, titles=400_000, steps=2, sample=1_000_000, chunk=25_000_000, device=device)
    # the byte convention of SURVEY.md 8(d) next to what ncu measured: both kernels are bound by instruction issue / latency
    # (bit-parallel alignments in registers and shared memory), not by the bytes of the strings
    out['indel_ratio']['measured_limiter'] = measured_limiter('k_indel_pairs')
    out['construct_features']['measured_limiter'] = measured_limiter('k_feature_words')
    return out


def example_dropin(device):
    """BASELINE configs[0] / [1] on the example data: the drop-in classes driven like Prediction drives them, each stage
    timed beside the staged reference's own code on the host (oracle/_ref; numba JIT warm)."""
    import bench_example
    return bench_example.run(device)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=3)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    parser.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    parser.add_argument('--queries', type=int, default=0, help='override the workload\'s query count')
    parser.add_argument('--truth', type=int, default=0, help='override the workload\'s truth row count')
    parser.add_argument('--top-n', type=int, default=10)
    parser.add_argument('--cpu-sample', type=int, default=20000)
    parser.add_argument('--truth-shards', type=int, default=0,
                        help='ranks the truth rows are sharded over (0 = auto: 1 for c3, all ranks for c5); queries are split over N / T groups')
    parser.add_argument('--device-encode', action='store_true', help='build the index with the GPU encoder also at N=1')
    parser.add_argument('--no-share-thresholds', action='store_true', help='truth shards prune with their local thresholds only (A/B)')
    parser.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline / parity sample (profiling runs)')
    parser.add_argument('--no-extra', action='store_true', help='skip the secondary workloads of the N = 1 record')
    parser.add_argument('--port-only', action='store_true', help='--impl reference: time the C port even when the reference is staged')
    args = parser.parse_args()
    spec = WORKLOADS[args.workload]
    args.queries = args.queries or spec['queries']
    args.truth = args.truth or spec['truth']
    args.source = spec['source']
    if args.no_cpu:
        args.no_extra = True
    # stdout must carry exactly one JSON line: libraries that print to fd 1 (NCCL's version banner) are
    # diverted to stderr, the JSON goes to the original stdout
    global _STDOUT
    _STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    import __graft_entry__ as entry
    lib = os.path.join(ROOT, 'doppelspeller_b200', '_lib', 'libdoppelspeller_b200.so')
    if not os.path.exists(lib) or not os.path.exists(os.path.join(ROOT, 'oracle', '_build', 'libds_oracle.so')):
        entry.build()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
