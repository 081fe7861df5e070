"""Profiling driver for the pair kernels on CANDIDATE pairs (the top_n candidates of every test title, like
predict.py:129-136 produces them): 20,000 x 100,000 synthetic titles, top-10 -> 200,000 pairs through the InDel-ratio
kernels (K2) and the 66-feature kernels (K3), twice.  Used under ncu (`-k regex:k_indel_chunks`, `-k regex:k_feature`)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from doppelspeller_b200 import _native as nat
    from doppelspeller_b200 import encode, synthetic
    from doppelspeller_b200 import feature_engineering as fe
    from doppelspeller_b200 import pipeline as pl
    from doppelspeller_b200.index import TruthIndex
    n_q, n_truth, k = 20000, 100000, 10
    device = torch.device('cuda', 0)
    truth = synthetic.generate_truth_titles(n_truth)
    test, _ = synthetic.generate_test_titles(truth, n_q)
    enc = encode.encode_canonical(test, truth)
    index = TruthIndex(enc['t_ptr'], enc['t_cols'], enc['idf64'], device=0)
    rows, _ = index.topn(torch.as_tensor(enc['q_ptr']).to(device), torch.as_tensor(enc['q_cols']).to(device), k)
    codes_a, off_a = fe.encode_titles(test)
    codes_b, off_b = fe.encode_titles(truth)
    counts = pl.truth_word_counts(truth)
    dev = lambda x: torch.as_tensor(x).to(device)   # noqa: E731
    a, oa, b, ob, c = dev(codes_a), dev(off_a), dev(codes_b), dev(off_b), dev(counts.view(np.int32))
    ia = torch.arange(n_q, device=device, dtype=torch.int32).repeat_interleave(k)
    ib = rows.reshape(-1).clamp(min=0).to(torch.int32)
    n = n_q * k
    ratio = torch.empty(n, dtype=torch.uint8, device=device)
    for _ in range(2):
        nat.check(nat.lib.ds_indel_ratio_pairs(nat.ptr(a), nat.ptr(oa), n_q, nat.ptr(b), nat.ptr(ob), n_truth, nat.ptr(ia), nat.ptr(ib), n,
                                               nat.ptr(ratio), None, nat.stream_for(ratio)))
        feats = fe.construct_features_pairs((a, oa), (b, ob), c, ia, ib, fe.SPACE_CODE, n_truth)
    torch.cuda.synchronize()
    print('ok', int(ratio.sum().item()), float(torch.nan_to_num(feats).sum().item()))


if __name__ == '__main__':
    main()
