"""TEST INFRASTRUCTURE ONLY - run by tests/test_emulated_kernels.py in a subprocess against the HOST EMULATION of the
kernels (tests/emu/cuda_emu.h), never collected by the normal test run.

The `gpu` parity tests that need real CUDA tensors cannot run against the emulated device; this module drives the
same entry points through the C ABI directly, with host buffers and with buffers that live in (emulated) DEVICE memory
so that the in-place / no-staging branches, the device-side list lengths and the peer-memory threshold exchange of
ds_topn_local_shared run under AddressSanitizer as well.  Every result is compared with the CPU oracle or the host
forms, exactly like the tests it mirrors (tests/test_gpu_parity.py, tests/test_gpu_dataframe_api.py).
"""
import ctypes

import numpy as np
import pytest

from doppelspeller_b200 import _native as nat
from tests.conftest import oracle_index_from_encoded

pytestmark = pytest.mark.gpu

_vp = ctypes.c_void_p
nat.lib.ds_emu_malloc.restype = _vp
nat.lib.ds_emu_malloc.argtypes = [ctypes.c_size_t]
nat.lib.ds_emu_free.argtypes = [_vp]
nat.lib.ds_emu_reads_of_absent_lanes.restype = ctypes.c_uint64
nat.lib.ds_emu_live_device_blocks.restype = ctypes.c_uint64


class DeviceArray:
    """numpy view of a block of emulated device memory (freed with the object)."""

    def __init__(self, shape, dtype, fill=None):
        self.dtype = np.dtype(dtype)
        self.shape = (shape,) if np.isscalar(shape) else tuple(shape)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.address = nat.lib.ds_emu_malloc(max(1, self.nbytes))
        assert self.address
        buffer = (ctypes.c_char * max(1, self.nbytes)).from_address(self.address)
        self.array = np.frombuffer(buffer, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)
        if fill is not None:
            self.array[...] = fill

    @classmethod
    def of(cls, host):
        host = np.ascontiguousarray(host)
        return cls(host.shape, host.dtype, fill=host)

    def close(self):
        if self.address:
            self.array = None
            nat.lib.ds_emu_free(self.address)
            self.address = None

    def __del__(self):
        self.close()


def _encode_trigrams(t_bytes, t_off, q_bytes, q_off):
    """ds_encode_trigrams with the given (host numpy or DeviceArray.array) tables -> dict like encode.encode_canonical"""
    n_truth, n_q = len(t_off) - 1, len(q_off) - 1
    max_vocab = int(nat.lib.ds_encode_max_vocab())
    t_ptr = np.empty(n_truth + 1, np.int64)
    q_ptr = np.empty(n_q + 1, np.int64)
    t_cols = np.empty(max(1, int(t_off[-1])), np.uint16)
    q_cols = np.empty(max(1, int(q_off[-1])), np.uint16)
    idf64 = np.empty(max_vocab, np.float64)
    vocab = np.empty(max_vocab, np.int32)
    n_vocab, t_nnz, q_nnz = ctypes.c_int32(0), ctypes.c_int64(0), ctypes.c_int64(0)
    nat.check(nat.lib.ds_encode_trigrams(nat.ptr(t_bytes), nat.ptr(t_off), n_truth, nat.ptr(q_bytes), nat.ptr(q_off), n_q, nat.ptr(t_ptr),
                                         nat.ptr(t_cols), nat.ptr(q_ptr), nat.ptr(q_cols), nat.ptr(idf64), nat.ptr(vocab),
                                         ctypes.byref(n_vocab), ctypes.byref(t_nnz), ctypes.byref(q_nnz), 0, None))
    return dict(idf64=idf64[:n_vocab.value].copy(), t_ptr=t_ptr, t_cols=t_cols[:t_nnz.value].copy(), q_ptr=q_ptr,
                q_cols=q_cols[:q_nnz.value].copy(), vocab_codes=vocab[:n_vocab.value].copy())


def test_trigram_encoder_host_and_device_tables():
    """f1 (k_trigrams, k_remap, k_vocab): bit-identical to the host encoder, from host tables and from device tables."""
    from doppelspeller_b200 import encode, synthetic
    truth = synthetic.generate_truth_titles(6000, seed=21) + synthetic.generate_long_titles(60, seed=22) + ['abc', 'aaaaaaa', 'ab ab ab ab']
    test, _ = synthetic.generate_test_titles(truth, 500, seed=23)
    want = encode.encode_canonical(test, truth)
    t_bytes, t_off = encode.title_table(truth)
    q_bytes, q_off = encode.title_table(test)
    device = [DeviceArray.of(x) for x in (t_bytes, t_off, q_bytes, q_off)]
    for got in (_encode_trigrams(t_bytes, t_off, q_bytes, q_off), _encode_trigrams(*[d.array for d in device])):
        for key in ('t_ptr', 'q_ptr', 't_cols', 'q_cols', 'vocab_codes'):
            assert np.array_equal(got[key], want[key]), key
        assert np.array_equal(got['idf64'].view(np.uint64), want['idf64'].view(np.uint64))
    bad_bytes, bad_off = encode.title_table(['bad\ttitle'])
    with pytest.raises(Exception, match='outside'):
        _encode_trigrams(t_bytes, t_off, bad_bytes, bad_off)


def test_title_features_from_the_table(example_titles):
    """f3 (k_title_codes, k_words, k_word_lookup): encode_title codes and word document-frequency vectors."""
    from doppelspeller_b200 import encode, pipeline
    from doppelspeller_b200 import feature_engineering as fe
    titles = list(example_titles['truth_titles'][:3000])
    titles += ['a a a b', 'ltd ltd', 'x', 'w1 w2 w3 w4 w5 w6 w7 w8 w9 w10 w11 w12 w13 w14 w15 w16 w17 w1', 'zz top zz', '000']
    raw, offsets = encode.title_table(titles)
    want_codes, want_offsets = fe.encode_titles(titles)
    want_counts = pipeline.truth_word_counts(titles)
    assert np.array_equal(want_offsets, offsets)
    for source in ('host', 'device'):
        held = [DeviceArray.of(raw), DeviceArray.of(offsets)] if source == 'device' else None
        table = (held[0].array, held[1].array) if held else (raw, offsets)
        codes = np.empty(max(1, raw.shape[0]), np.uint8)
        counts = np.empty((len(titles), 15), np.uint32)
        nat.check(nat.lib.ds_title_features(nat.ptr(table[0]), nat.ptr(table[1]), len(titles), nat.ptr(codes), nat.ptr(counts), 0, None))
        assert np.array_equal(codes[:want_codes.shape[0]], want_codes)
        assert np.array_equal(counts, want_counts)
    bad = np.frombuffer(b'caf\xe9', dtype=np.uint8).copy()
    with pytest.raises(Exception):
        nat.check(nat.lib.ds_title_features(nat.ptr(bad), nat.ptr(np.array([0, 4], np.int64)), 1, nat.ptr(np.empty(4, np.uint8)), None, 0, None))


def test_prematch_cascade_and_selection(example_titles, golden_matchmaker):
    """f2 (k_prematch_filter, K2 mode 1 with device-side list lengths, k_prematch_again, k_select_close)."""
    from doppelspeller_b200 import predict
    from doppelspeller_b200.common import _string_table, _token_sort
    from oracle import oracle
    n_q, k = 120, 100
    truth, test = example_titles['truth_titles'], example_titles['test_titles'][:n_q]
    rows = golden_matchmaker['top100_rows'][:n_q]
    used = np.unique(rows)
    remap = np.full(len(truth), -1, np.int64)
    remap[used] = np.arange(len(used))
    truth_used = [truth[t] for t in used]
    idx_a = np.repeat(np.arange(n_q, dtype=np.int32), k)
    idx_b = remap[rows.reshape(-1)].astype(np.int32)
    tables = []
    for titles in (test, truth_used):
        tables += list(_string_table(titles)) + list(_string_table([_token_sort(t) for t in titles]))
    out = np.empty(len(idx_a), np.int32)
    for source in ('host', 'device'):
        held = [DeviceArray.of(x) for x in tables + [idx_a, idx_b]] if source == 'device' else None
        args = [h.array for h in held] if held else tables + [idx_a, idx_b]
        out[:] = -7
        nat.check(nat.lib.ds_prematch_pairs(nat.ptr(args[0]), nat.ptr(args[1]), nat.ptr(args[2]), nat.ptr(args[3]), len(test),
                                            nat.ptr(args[4]), nat.ptr(args[5]), nat.ptr(args[6]), nat.ptr(args[7]), len(truth_used),
                                            nat.ptr(args[8]), nat.ptr(args[9]), len(idx_a), 94, nat.ptr(out), None))
        want = np.array([oracle.prematch_ratio(test[a], truth_used[b]) for a, b in zip(idx_a, idx_b)])
        assert np.array_equal(out, want)
    assert (out > 94).sum() > 0 and (out == 0).sum() > 0
    test_index = np.repeat(np.arange(n_q), k)
    tied = out.copy()
    tied[k * 3 + 1] = tied[k * 3 + 2] = 99
    invalid = np.zeros(len(tied), dtype=np.uint8)
    invalid[k * 5:k * 6] = 1
    for ratios, mask in ((out, None), (tied, invalid)):
        want_pairs = np.full(n_q, -1, dtype=np.int64)
        kept = predict.select_close_matches(test_index, ratios if mask is None else np.where(mask != 0, 0, ratios))
        want_pairs[test_index[kept]] = kept
        assert np.array_equal(predict.select_close_matches_grouped(ratios, n_q, k, invalid=mask), want_pairs)


def test_topn_with_device_resident_queries(golden_matchmaker):
    """ds_topn with the query CSR and every output in device memory (no staging) equals the host-buffer call."""
    from doppelspeller_b200.index import TruthIndex
    g = golden_matchmaker
    n_q, k = 300, 10
    q_ptr = g['q_ptr'][:n_q + 1]
    q_cols = g['q_cols'][:int(q_ptr[-1])]
    index = TruthIndex(g['t_ptr'], g['t_cols'], g['w64'])
    want_rows, want_count = index.topn(q_ptr, q_cols, k)
    assert np.array_equal(want_rows, g['top10_rows'][:n_q].astype(np.int64))
    d_ptr, d_cols = DeviceArray.of(q_ptr), DeviceArray.of(q_cols)
    d_rows, d_count = DeviceArray((n_q, k), np.int64), DeviceArray(n_q, np.int32)
    index.topn(d_ptr.array, d_cols.array, k, out_rows=d_rows.array, out_count=d_count.array)
    assert np.array_equal(d_rows.array, want_rows) and np.array_equal(d_count.array, want_count)
    # the index itself from device-resident CSR arrays
    held = [DeviceArray.of(g[key]) for key in ('t_ptr', 't_cols', 'w64')]
    index2 = TruthIndex(*[h.array for h in held])
    rows2, _ = index2.topn(q_ptr, q_cols, k)
    assert np.array_equal(rows2, want_rows)
    index.close()
    index2.close()


def test_thresholds_shared_between_shards(golden_matchmaker):
    """ds_topn_local_shared: three shards publish / read their pruning thresholds through peer arrays (here: three blocks
    of emulated device memory, the shards scanned one after the other, so that later shards really prune with the bounds
    the earlier ones published) -> merge -> rescan: the single-index answer."""
    import torch
    from doppelspeller_b200 import sharded
    from doppelspeller_b200.index import TruthIndex, topn_merge
    g = golden_matchmaker
    n = int(g['t_ptr'].shape[0]) - 1
    n_q = 250
    q_ptr = g['q_ptr'][:n_q + 1]
    q_cols = g['q_cols'][:int(q_ptr[-1])]
    for k, key in ((10, 'top10_rows'), (100, 'top100_rows')):
        offs = sharded.shard_offsets(n, 3)
        shards = []
        for r in range(3):
            ptr, cols = sharded.slice_truth_csr(g['t_ptr'], g['t_cols'], int(offs[r]), int(offs[r + 1]))
            shards.append(TruthIndex(ptr, cols, g['w64'], row_offset=int(offs[r]), n_total=n))
        thetas = [DeviceArray(n_q, np.float64, fill=0.0) for _ in shards]
        local = []
        for r in (2, 0, 1):      # any order is valid: published bounds only ever prune rows that cannot qualify
            peers = [thetas[p].address for p in range(3) if p != r]
            local.append((r, shards[r].topn_local(q_ptr, q_cols, k, theta_own=thetas[r].array, theta_peers=peers)))
        assert all((t.array > 0).any() for t in thetas)          # every shard published bounds
        local = [l for _, l in sorted(local, key=lambda item: item[0])]
        all_score = np.stack([l[0] for l in local])
        all_row = np.stack([l[1] for l in local])
        rows, count, kth, thr, flags = topn_merge(all_score, all_row, k, n, q_mx=local[0][2])
        flagged = np.nonzero(flags & 1)[0]
        if flagged.size:
            per_rows, per_count = [], []
            for s in shards:
                lr = np.full((len(flags), k), -1, dtype=np.int64)
                lc = np.zeros(len(flags), dtype=np.int32)
                s.topn_rescan(q_ptr, q_cols, local[0][2], thr, flags, k, lr, lc)
                per_rows.append(lr[flagged])
                per_count.append(lc[flagged])
            fixed, fixed_count = sharded.combine_rescans(torch.as_tensor(np.stack(per_rows)), torch.as_tensor(np.stack(per_count)), k)
            rows[flagged] = fixed.numpy()
            count[flagged] = fixed_count.numpy()
        assert np.array_equal(rows, g[key][:n_q].astype(np.int64))
        assert (count == k).all()
        for s in shards:
            s.close()


def test_pair_kernels_with_device_resident_tables(golden_pairs):
    """K2 / K3 table form (k_indel_groups on candidate runs, the sorted class pipeline, k_feature_words) with every buffer
    in device memory."""
    from doppelspeller_b200 import feature_engineering as fe
    from tests.conftest import features_equal
    g = golden_pairs
    titles = [str(t) for t in g['feat_titles']]
    truths = [str(t) for t in g['feat_truths']]
    n = len(titles)
    idx = np.arange(n, dtype=np.int32)
    codes_a, off_a = fe.encode_titles(titles)
    codes_b, off_b = fe.encode_titles(truths)
    held = [DeviceArray.of(x) for x in (codes_a, off_a, codes_b, off_b, g['feat_counts'], idx)]
    d = [h.array for h in held]
    out = DeviceArray((n, 66), np.float32)
    nat.check(nat.lib.ds_construct_features_pairs(nat.ptr(d[0]), nat.ptr(d[1]), n, nat.ptr(d[2]), nat.ptr(d[3]), n, nat.ptr(d[4]),
                                                  nat.ptr(d[5]), nat.ptr(d[5]), 1, int(g['feat_n_truth']), n, nat.ptr(out.array), None))
    assert features_equal(out.array, g['feat_out'])
    # candidate-list shape: runs of 10 pairs sharing their first title -> k_indel_groups
    run = 10
    n_runs = n // run
    idx_a = np.repeat(np.arange(n_runs, dtype=np.int32), run)
    idx_b = np.arange(n_runs * run, dtype=np.int32)
    d_a, d_b = DeviceArray.of(idx_a), DeviceArray.of(idx_b)
    ratio = DeviceArray(n_runs * run, np.uint8)
    nat.check(nat.lib.ds_indel_ratio_pairs(nat.ptr(d[0]), nat.ptr(d[1]), n, nat.ptr(d[2]), nat.ptr(d[3]), n, nat.ptr(d_a.array),
                                           nat.ptr(d_b.array), n_runs * run, nat.ptr(ratio.array), None, None))
    from oracle import oracle
    a = np.vstack([fe.encode_title(titles[i]) for i in idx_a])
    b = np.vstack([fe.encode_title(truths[i]) for i in idx_b])
    la = np.array([len(titles[i]) for i in idx_a], np.uint8)
    lb = np.array([len(truths[i]) for i in idx_b], np.uint8)
    assert np.array_equal(ratio.array, oracle.indel_ratio_u8_batch(a, b, la, lb))


def test_no_shuffle_read_an_absent_lane_and_nothing_leaked():
    """Runs last (file order): across everything this process executed, no shuffle read a lane that was not part of the
    collective (an undefined value on the GPU) and every workspace was released."""
    import gc
    gc.collect()
    assert nat.lib.ds_emu_reads_of_absent_lanes() == 0
    assert nat.lib.ds_emu_live_device_blocks() == 0
