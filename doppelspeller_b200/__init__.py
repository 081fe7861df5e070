"""doppelspeller_b200 - B200-native candidate generation + pair scoring for DoppelSpeller.

Mirrors the reference's Python interface for its hot path only:
  match_maker.MatchMaker                      <- doppelspeller/match_maker.py
  feature_engineering.construct_features      <- doppelspeller/feature_engineering.py:75-169
  feature_engineering.fast_levenshtein_ratio  <- doppelspeller/feature_engineering.py:25-63
  common.levenshtein_ratio / _token_sort_ratio <- doppelspeller/common.py:161-167
  predict.get_levenshtein_ratios              <- Prediction._get_levenshtein_ratio (predict.py:140-156)
Everything computes in hand-written sm_100a CUDA kernels behind the C ABI of
include/doppelspeller_b200.h; there is no CPU fallback.
"""
__version__ = '0.1.0'
