"""The whole hot path as one GPU-resident pass: titles in, candidate rows + 66 features per candidate out.

    encode (csrc/ds_encode.cu) -> top-n candidates (csrc/ds_topn.cu) -> features of every (title, candidate)
    pair (csrc/ds_pairs.cu)

This is what `Prediction.generate_test_predictions` does between the exact-match join and `model.predict`
(/root/reference/doppelspeller/predict.py:286-219): MatchMaker(...) + get_closest_matches per row +
construct_features on the title/candidate pairs - without the pandas glue and without materialising
[P, 255] arrays.  Intermediate results never leave the device.
"""
from collections import Counter

import numpy as np

from . import encode
from . import feature_engineering as fe
from .index import TruthIndex


def truth_word_counts(truth_titles):
    """[n_truth, 15] uint32: document frequency of each of the first 15 words of every truth title
    (common.get_words_counter common.py:140-142 + FeatureEngineering.get_truth_words_counts :309-319)."""
    counter = Counter(w for t in truth_titles for w in set(t.split()))
    counts = np.zeros((len(truth_titles), fe.NUMBER_OF_WORDS_FEATURES), dtype=np.uint32)
    for i, t in enumerate(truth_titles):
        ws = [counter[w] for w in t.split()[:fe.NUMBER_OF_WORDS_FEATURES]]
        counts[i, :len(ws)] = ws
    return counts


def title_features_device(table, device, codes=True, word_counts=True):
    """encode_title codes and get_truth_words_counts vectors of a device-resident title table (`encode.title_table` /
    `common.transform_titles_table` output: (bytes uint8, offsets int64) CUDA tensors), computed on the GPU
    (ds_title_features) -> (codes uint8[total] or None, counts int32[n, 15] (uint32 bit pattern) or None)."""
    import torch

    from . import _native as nat
    data, offsets = table
    dev = torch.device('cuda', device)
    n = int(offsets.shape[0]) - 1
    out_codes = torch.empty(max(1, int(data.shape[0])), dtype=torch.uint8, device=dev) if codes else None
    out_counts = torch.empty((n, fe.NUMBER_OF_WORDS_FEATURES), dtype=torch.int32, device=dev) if word_counts else None
    with torch.cuda.device(dev):
        nat.check(nat.lib.ds_title_features(nat.ptr(data), nat.ptr(offsets), n, nat.ptr(out_codes), nat.ptr(out_counts), int(device),
                                            torch.cuda.current_stream(dev).cuda_stream))
    return out_codes, out_counts


class CandidatePipeline:
    """Holds the truth side on the GPU (index, code table, word counts); `run(test_titles, top_n)` returns
    (rows int64[Q, top_n] descending truth rows, features float32[Q * top_n, 66]) as CUDA tensors."""

    def __init__(self, truth_titles, device=0, raw=False):
        """`raw=True`: the titles are raw input strings; common.transform_title (common.py:20-47) runs on the GPU first."""
        import torch
        self.device = torch.device('cuda', device)
        if raw:
            from . import common
            with torch.cuda.device(self.device):
                truth_titles = common.transform_titles(truth_titles)
        self.truth_titles = truth_titles
        raw, raw_offsets = encode.title_table(truth_titles)            # ASCII bytes: the only host pass over the titles
        self.truth_table = (torch.as_tensor(raw).to(self.device), torch.as_tensor(raw_offsets).to(self.device))
        # encode_title codes and the word document-frequency vectors come from the same table, on the device (f3)
        self.truth_codes, self.word_counts = title_features_device(self.truth_table, device)
        self.truth_offsets = self.truth_table[1]
        self._prematch_truth = None

    def run(self, test_titles, top_n, with_prematch=False, raw=False):
        """`raw=True`: `test_titles` are raw input strings (transformed on the GPU first).
        -> (rows, count, features) or, with_prematch, (rows, count, features, ratios) where ratios int32[Q * top_n]
        is Prediction._get_levenshtein_ratio of every (title, candidate) pair (predict.py:147-156, the input of the
        "very close match" selection :172-176)."""
        import torch
        if raw:
            from . import common
            with torch.cuda.device(self.device):
                test_titles = common.transform_titles(test_titles)
        test_raw, test_raw_offsets = encode.title_table(test_titles)
        test_table = (torch.as_tensor(test_raw).to(self.device), torch.as_tensor(test_raw_offsets).to(self.device))
        enc = encode.encode_canonical_device(None, None, device=self.device.index, truth_table=self.truth_table, query_table=test_table)
        index = TruthIndex(enc['t_ptr'], enc['t_cols'], enc['idf64'], device=self.device.index)
        rows, count = index.topn(enc['q_ptr'], enc['q_cols'], top_n)
        index.close()
        test_codes, _ = title_features_device(test_table, self.device.index, word_counts=False)
        test_offsets = test_table[1]
        n_q = len(test_titles)
        title_index = torch.arange(n_q, device=self.device, dtype=torch.int32).repeat_interleave(top_n)
        truth_index = rows.reshape(-1).clamp(min=0).to(torch.int32)
        features = fe.construct_features_pairs((test_codes, test_offsets), (self.truth_codes, self.truth_offsets), self.word_counts,
                                               title_index, truth_index, fe.SPACE_CODE, len(self.truth_titles))
        if not with_prematch:
            return rows, count, features
        from . import predict
        if self._prematch_truth is None:
            self._prematch_truth = predict.PrematchTables(self.truth_titles, device=self.device.index)
        ratios = predict.get_levenshtein_ratios_indexed(predict.PrematchTables(test_titles, device=self.device.index),
                                                        self._prematch_truth, title_index, truth_index)
        return rows, count, features, ratios

    def predict(self, test_titles, top_n, model, raw=False):
        """Prediction.generate_test_predictions after the exact-match join (predict.py:286-219,229-252), device resident up
        to the final selection: candidates -> "very close" matches by the fuzzy ratio cascade (ratio > 94, the maximum of
        its title, attained once: :158-176) -> for the remaining titles the model's probability over the 66 features
        (`model`: gbdt.GbdtModel) > 0.9, the maximum of its title, attained once (:242-249).
        -> dict(rows int64[Q, top_n], match_row int64[Q] (truth row or -1), match_kind uint8[Q] (0 none, 1 close match,
        2 model), prediction float32[Q] (1.0 for close matches, the model's probability otherwise))."""
        import torch

        from . import gbdt, predict
        rows, count, features, ratios = self.run(test_titles, top_n, with_prematch=True, raw=raw)
        n_q = rows.shape[0]
        probabilities = model.predict(features).cpu().numpy()
        rows_host = rows.cpu().numpy()
        test_index = np.repeat(np.arange(n_q), top_n)
        valid = (rows_host.reshape(-1) >= 0)
        match_row = np.full(n_q, -1, dtype=np.int64)
        match_kind = np.zeros(n_q, dtype=np.uint8)
        prediction = np.zeros(n_q, dtype=np.float32)
        # predict.py:158-176 on the device: the pair of every title whose ratio is > 94 and alone attains the title's maximum
        invalid = (rows.reshape(-1) < 0).to(torch.uint8).contiguous()
        chosen_pair = predict.select_close_matches_grouped(ratios.contiguous(), n_q, top_n, invalid=invalid).cpu().numpy()
        close = chosen_pair[chosen_pair >= 0]
        match_row[test_index[close]] = rows_host.reshape(-1)[close]
        match_kind[test_index[close]] = 1
        prediction[test_index[close]] = 1.0
        open_pairs = valid & (match_kind[test_index] == 0)                       # predict.py:176: titles matched so far drop out
        chosen = gbdt.select_model_matches(test_index, np.where(open_pairs, probabilities, -1.0))
        match_row[test_index[chosen]] = rows_host.reshape(-1)[chosen]
        match_kind[test_index[chosen]] = 2
        prediction[test_index[chosen]] = probabilities[chosen]
        return dict(rows=rows, match_row=match_row, match_kind=match_kind, prediction=prediction)
