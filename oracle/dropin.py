"""Drop-in proof harness - TEST INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline leg).

Runs the reference's own `Prediction(...).generate_test_predictions()` (predict.py:274-317; the code behind the CLI's
`generate-predictions`, cli.py:52-61, and `closest-search-single-title`, cli.py:64-83) twice in one process:

  * unpatched - the reference exactly as staged under oracle/_ref;
  * patched   - the same code with ONLY the three imports of INTEGRATION.md section 1 swapped inside
                `doppelspeller.predict`: MatchMaker, construct_features (+ FEATURES_COUNT) and the two fuzzy ratios;

and records what crosses the hot-path boundary in each run: the candidate list of every row, the pre-match ratio of
every (title, candidate) pair, the feature matrix handed to the model, the predictions frame and the output file.

Two things the reference needs that this image lacks are shimmed identically in BOTH runs (they are outside the path):
`Levenshtein.ratio` (python-levenshtein 0.12.0: restated InDel ratio, C) and xgboost + the pickled model (a
deterministic stand-in model over the feature matrix, so the selection stages after the model still run).
"""
import contextlib
import os
import sys
import types

import numpy as np


class StandInModel:
    """Deterministic stand-in for the pickled xgboost Booster (predict.py:80-82, :233): a logistic over a few
    always-finite features.  Identical in the patched and the unpatched run."""
    best_ntree_limit = 0

    @staticmethod
    def predict(matrix, ntree_limit=0):
        x = np.asarray(matrix.data, dtype=np.float32)
        best = np.nan_to_num(x[:, 6:21], nan=0.0).max(axis=1)
        z = (x[:, 4] - 82.0) / 6.0 + (x[:, 5] - 80.0) / 12.0 + (best - 90.0) / 20.0
        return (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)


class _DMatrix:
    captured = None   # list the feature matrices are appended to while a run is recorded

    def __init__(self, data):
        self.data = np.array(data, copy=True)
        if _DMatrix.captured is not None:
            _DMatrix.captured.append(self.data)


def _c_levenshtein_ratio(a, b):
    """python-levenshtein's ratio(a, b) = (la + lb - indel) / (la + lb), 1.0 for two empty strings (restated)."""
    import ctypes
    from oracle import oracle
    la, lb = len(a), len(b)
    if la + lb == 0:
        return 1.0
    ba = np.frombuffer(a.encode('latin-1', 'replace'), dtype=np.uint8)
    bb = np.frombuffer(b.encode('latin-1', 'replace'), dtype=np.uint8)
    d = oracle.lib().orc_indel_distance(ba.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.c_int(la),
                                        bb.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.c_int(lb))
    return (la + lb - d) / (la + lb)


def load_reference():
    """The staged reference with the two out-of-path shims installed (see the module docstring)."""
    if 'Levenshtein' not in sys.modules or not hasattr(sys.modules['Levenshtein'], '_ds_shim'):
        lev = types.ModuleType('Levenshtein')
        lev.ratio = _c_levenshtein_ratio
        lev._ds_shim = True
        sys.modules['Levenshtein'] = lev
    if 'xgboost' not in sys.modules or not hasattr(sys.modules['xgboost'], 'DMatrix'):
        xgb = types.ModuleType('xgboost')
        xgb.DMatrix = _DMatrix
        sys.modules['xgboost'] = xgb
    from oracle import ref_import
    ref = ref_import.import_reference()
    ref.common.ratio = _c_levenshtein_ratio                  # common.py:8 bound the name at import time
    ref.predict.xgb = sys.modules['xgboost']
    return ref


@contextlib.contextmanager
def swapped_imports(ref, patched):
    """the three import lines of doppelspeller/predict.py, swapped and restored."""
    names = ('MatchMaker', 'construct_features', 'FEATURES_COUNT', 'levenshtein_ratio', 'levenshtein_token_sort_ratio')
    saved = {name: getattr(ref.predict, name) for name in names}
    try:
        if patched:
            from doppelspeller_b200.common import levenshtein_ratio, levenshtein_token_sort_ratio
            from doppelspeller_b200.feature_engineering import FEATURES_COUNT, construct_features
            from doppelspeller_b200.match_maker import MatchMaker
            ref.predict.MatchMaker = MatchMaker
            ref.predict.construct_features = construct_features
            ref.predict.FEATURES_COUNT = FEATURES_COUNT
            ref.predict.levenshtein_ratio = levenshtein_ratio
            ref.predict.levenshtein_token_sort_ratio = levenshtein_token_sort_ratio
        yield
    finally:
        for name, value in saved.items():
            setattr(ref.predict, name, value)


def _limit_test_file(ref, n_rows):
    """Keeps the first n_rows data rows of the staged example_test.csv copy (settings.TEST_FILE)."""
    from oracle import ref_import
    source = os.path.join(ref_import.STAGED_ROOT, 'example_dataset', os.path.basename(ref.settings.TEST_FILE))
    with open(source) as fin:
        lines = fin.readlines()
    with open(ref.settings.TEST_FILE, 'w') as fout:
        fout.writelines(lines if n_rows is None else lines[:1 + n_rows])


def run_prediction(ref, patched, n_test_rows=None, title=None):
    """One recorded run of Prediction.generate_test_predictions (all test rows of the example set, its first
    n_test_rows, or the single `title` of closest-search-single-title).  Returns a dict of what crossed the boundary."""
    import pandas as pd
    predict = ref.predict
    record = {'candidates': {}, 'ratios': [], 'features': []}
    _limit_test_file(ref, n_test_rows)
    with swapped_imports(ref, patched):
        class Recorded(predict.Prediction):
            @staticmethod
            def _load_model():
                return StandInModel()

            def _combine_titles_with_matches(self):
                inner = self.match_maker.get_closest_matches

                def recording(row_number):
                    out = inner(row_number)
                    record['candidates'][int(row_number)] = list(out)
                    return out
                self.match_maker.get_closest_matches = recording
                try:
                    return super()._combine_titles_with_matches()
                finally:
                    self.match_maker.get_closest_matches = inner

            def _find_close_matches(self):
                remaining = super()._find_close_matches()
                return remaining

            @classmethod
            def _get_levenshtein_ratio(cls, x, y):
                value = super()._get_levenshtein_ratio(x, y)
                record['ratios'].append(int(value))
                return value

        _DMatrix.captured = record['features']
        try:
            with np.errstate(all='ignore'):
                if title is not None:
                    prediction = Recorded(ref.constants.DATA_TYPE_SINGLE, title=title)
                    record['single'] = prediction.generate_test_predictions(single_prediction=True)
                else:
                    prediction = Recorded(ref.constants.DATA_TYPE_TEST)
                    output = prediction.generate_test_predictions()
                    record['output'] = pd.read_csv(output, sep=ref.settings.TEST_FILE_DELIMITER)
        finally:
            _DMatrix.captured = None
        record['predictions'] = prediction.predictions.copy(deep=True)
    record['ratios'] = np.array(record['ratios'], dtype=np.int32)
    record['features'] = np.vstack(record['features']) if record['features'] else np.zeros((0, 66), np.float32)
    _limit_test_file(ref, None)
    return record


def compare_runs(reference_run, patched_run, rtol=1e-6):
    """-> dict of mismatch counts between two recorded runs (all zero = the swap changed nothing observable)."""
    a, b = reference_run, patched_run
    rows = sorted(a['candidates'])
    out = {'rows': len(rows), 'pairs': int(a['ratios'].shape[0]), 'feature_rows': int(a['features'].shape[0])}
    out['candidate_row_sets_differ'] = int(sorted(b['candidates']) != rows)
    out['candidate_list_mismatches'] = sum(1 for r in rows if a['candidates'][r] != b['candidates'].get(r))
    out['ratio_mismatches'] = int((a['ratios'] != b['ratios']).sum()) if a['ratios'].shape == b['ratios'].shape else -1
    fa, fb = a['features'], b['features']
    if fa.shape != fb.shape:
        out['integer_feature_mismatches'] = out['float_feature_mismatches'] = -1
    else:
        exact = (fa[:, :36] == fb[:, :36]) | (np.isnan(fa[:, :36]) & np.isnan(fb[:, :36]))
        with np.errstate(all='ignore'):
            close = np.isclose(fa[:, 36:], fb[:, 36:], rtol=rtol, atol=0, equal_nan=True)
        out['integer_feature_mismatches'] = int((~exact).sum())
        out['float_feature_mismatches'] = int((~close).sum())
    pa = a['predictions'].reset_index(drop=True)
    pb = b['predictions'].reset_index(drop=True)
    same_frame = pa.shape == pb.shape and all(
        np.array_equal(pa[col].to_numpy(), pb[col].to_numpy()) for col in pa.columns if col in pb.columns)
    out['prediction_frames_differ'] = int(not same_frame)
    if 'output' in a:
        out['output_files_differ'] = int(not a['output'].equals(b['output']))
    if 'single' in a:
        out['single_results_differ'] = int({k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in a['single'].items()} !=
                                           {k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in b['single'].items()})
    return out
