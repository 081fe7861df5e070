"""bench_pairs.py - secondary benchmark: BASELINE.json configs[3] ("C4"), batched InDel ratio (K2) and the
66-feature construct_features kernel (K3) over synthetic candidate pairs, titles up to 128 characters.

    python bench_pairs.py [--pairs 100000000] [--steps 3] [--sample 200000]

Pairs (seed 20240503): 90 % of the titles follow the example length distribution, 10 % are
a stress slice with lengths uniform in [65, 128]; half of the pairs are candidate-like (a truth title and a
misspelling of it), half are random.  Titles live in compact (bytes, offsets) tables in HBM, pairs are
(title index, truth index) arrays - the B200 layout of the C ABI (`ds_*_pairs`).  Prints one JSON line with
pairs/s, the algorithmic-byte HBM fraction of each kernel and a parity check of a sample against the CPU
oracle.  Not the driver's headline bench (that is bench.py); kept for the "Levenshtein pairs/sec" metric.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def build_pairs(n_pairs, n_titles):
    from doppelspeller_b200 import synthetic
    rng = np.random.default_rng(synthetic.PAIRS_SEED)
    n_long = n_titles // 10
    truth = synthetic.generate_truth_titles(n_titles - n_long, seed=synthetic.PAIRS_SEED + 1)
    truth += synthetic.generate_long_titles(n_long, seed=synthetic.PAIRS_SEED + 2)
    test, source = synthetic.generate_test_titles(truth, n_titles, seed=synthetic.PAIRS_SEED + 3, matched=1.0)
    idx_a = rng.integers(0, n_titles, size=n_pairs).astype(np.int32)
    idx_b = np.where(rng.random(n_pairs) < 0.5, source[idx_a], rng.integers(0, n_titles, size=n_pairs)).astype(np.int32)
    return truth, test, idx_a, idx_b


def run(pairs=100_000_000, titles=400_000, steps=3, sample=200_000, chunk=25_000_000, device=None):
    import torch
    from doppelspeller_b200 import _native as nat
    from doppelspeller_b200 import feature_engineering as fe
    from doppelspeller_b200 import pipeline as pl
    from oracle import oracle
    device = torch.device('cuda', 0) if device is None else device
    t0 = time.time()
    truth, test, idx_a, idx_b = build_pairs(pairs, titles)
    counts = pl.truth_word_counts(truth)
    codes_a, off_a = fe.encode_titles(test)
    codes_b, off_b = fe.encode_titles(truth)
    print(f'[bench_pairs] {pairs} pairs over {len(test)} + {len(truth)} titles built in {time.time() - t0:.1f}s', file=sys.stderr)
    dev = lambda x: torch.as_tensor(x).to(device)   # noqa: E731
    d = dict(a=dev(codes_a), oa=dev(off_a), b=dev(codes_b), ob=dev(off_b), c=dev(counts.view(np.int32)), ia=dev(idx_a), ib=dev(idx_b))
    la = np.diff(off_a)[idx_a[:2_000_000]]
    lb = np.diff(off_b)[idx_b[:2_000_000]]
    mean_len = float((la + lb).mean())
    chunk = min(chunk, pairs)
    ratio = torch.empty(pairs, dtype=torch.uint8, device=device)
    feats = torch.empty((chunk, 66), dtype=torch.float32, device=device)

    def run_ratio():
        nat.check(nat.lib.ds_indel_ratio_pairs(nat.ptr(d['a']), nat.ptr(d['oa']), len(test), nat.ptr(d['b']), nat.ptr(d['ob']), len(truth),
                                               nat.ptr(d['ia']), nat.ptr(d['ib']), pairs, nat.ptr(ratio), None, nat.stream_for(ratio)))

    def run_feats():
        for c0 in range(0, pairs, chunk):
            c1 = min(pairs, c0 + chunk)
            fe.construct_features_pairs((d['a'], d['oa']), (d['b'], d['ob']), d['c'], d['ia'][c0:c1], d['ib'][c0:c1], fe.SPACE_CODE,
                                        len(truth), response=feats[:c1 - c0])

    peak = 6551.7
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        peak = float(json.load(open(path))['hbm_gbs'])
    line = {'metric': 'Levenshtein pairs/sec', 'unit': 'pairs/s', 'pairs': pairs, 'data': 'synthetic',
            'config': {'workload': f'C4 synthetic {pairs} candidate pairs, 10% titles in [65,128] chars (BASELINE.json configs[3])',
                       'mean_la_plus_lb': mean_len}}
    for name, fn, bpp in (('indel_ratio', run_ratio, mean_len + 3.0), ('construct_features', run_feats, mean_len + 2.0 + 60.0 + 264.0)):
        fn()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        stop.synchronize()
        ms = start.elapsed_time(stop) / steps
        achieved = pairs * bpp / (ms / 1e3) / 1e9
        line[name] = {'pairs_per_s': pairs / (ms / 1e3), 'ms': ms, 'algorithmic_bytes_per_pair': bpp, 'hbm_frac': achieved / peak,
                      'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                                   'traffic': None}}
    # parity of a sample against the oracle (padded layout on the CPU side)
    n = min(sample, pairs, chunk)
    sel = np.arange(n)
    la_s = np.diff(off_a)[idx_a[sel]].astype(np.uint8)
    lb_s = np.diff(off_b)[idx_b[sel]].astype(np.uint8)
    pa = _padded_rows(codes_a, off_a, idx_a[sel])
    pb = _padded_rows(codes_b, off_b, idx_b[sel])
    want_ratio = oracle.indel_ratio_u8_batch(pa, pb, la_s, lb_s)
    want_feats = oracle.construct_features(la_s, lb_s, pa, pb, counts[idx_b[sel]], fe.SPACE_CODE, len(truth))
    run_feats_first = fe.construct_features_pairs((d['a'], d['oa']), (d['b'], d['ob']), d['c'], d['ia'][:n], d['ib'][:n], fe.SPACE_CODE,
                                                  len(truth)).cpu().numpy()
    got_ratio = ratio[:n].cpu().numpy()
    exact = (run_feats_first[:, :36] == want_feats[:, :36]) | (np.isnan(run_feats_first[:, :36]) & np.isnan(want_feats[:, :36]))
    with np.errstate(all='ignore'):
        close = np.isclose(run_feats_first[:, 36:], want_feats[:, 36:], rtol=1e-6, atol=0, equal_nan=True)
    # CPU baseline: the C port of the reference (oracle/ds_oracle.c, OpenMP over pairs) on the same sample
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    oracle.indel_ratio_u8_batch(pa, pb, la_s, lb_s, n_threads=cores)
    t_ratio = time.perf_counter() - t0
    t0 = time.perf_counter()
    oracle.construct_features(la_s, lb_s, pa, pb, counts[idx_b[sel]], fe.SPACE_CODE, len(truth), n_threads=cores)
    t_feat = time.perf_counter() - t0
    line['cpu_baseline'] = {'kind': 'port', 'cores': cores, 'sample': f'the first {n} pairs of the batch', 'unit': 'pairs/s',
                            'value': n / t_feat, 'indel_ratio_pairs_per_s': n / t_ratio, 'construct_features_pairs_per_s': n / t_feat}
    line['parity'] = {'sampled_pairs': int(n), 'ratio_mismatches': int((got_ratio != want_ratio).sum()),
                      'integer_feature_mismatches': int((~exact).sum()), 'float_feature_mismatches': int((~close).sum())}
    return line


def _padded_rows(codes, offsets, index):
    """[n, 255] zero padded code rows of the titles `index` (the reference's layout, predict.py:199-204), vectorised."""
    lengths = (offsets[index + 1] - offsets[index]).astype(np.int64)
    out = np.zeros((index.shape[0], 255), dtype=np.uint8)
    row_of = np.repeat(np.arange(index.shape[0]), lengths)
    col_of = np.arange(int(lengths.sum())) - np.repeat(np.cumsum(lengths) - lengths, lengths)
    out[row_of, col_of] = codes[np.repeat(offsets[index], lengths) + col_of]
    return out


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--pairs', type=int, default=100_000_000)
    parser.add_argument('--titles', type=int, default=400_000)
    parser.add_argument('--steps', type=int, default=3)
    parser.add_argument('--sample', type=int, default=200_000)
    parser.add_argument('--chunk', type=int, default=25_000_000, help='pairs per kernel launch (bounds the 264 B/pair output)')
    args = parser.parse_args()
    print(json.dumps(run(args.pairs, args.titles, args.steps, args.sample, args.chunk)), flush=True)


if __name__ == '__main__':
    main()
