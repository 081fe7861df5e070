"""Loader for the UNMODIFIED reference (test infrastructure, never shipped).

Imports the reference's `doppelspeller` package from `oracle/_ref` (staged by `oracle/stage_reference.py`, an
offline `pip install --target` of /root/reference; git-ignored, travels to the GPU box) so that the oracle
restatement (`oracle/ds_oracle.c`, `oracle/oracle.py`) can be pinned against the reference's own numba kernels,
`tests/golden/make_golden.py` can mint golden vectors, `oracle/dropin.py` can run the reference's `Prediction`
with and without the three swapped imports of INTEGRATION.md, and `bench.py --impl reference` can time the
reference's own kernels.  `/root/reference` itself does not exist on the GPU box and is never read at run time
there: only the staged copy is.

Shims needed:
  * `Levenshtein` (python-levenshtein==0.12.0, requirements.txt:9) is not installed and
    `doppelspeller/common.py:8` imports it at module import time -> a stub exposing `ratio`
    backed by the restated InDel ratio (oracle.lev_ratio_float).
  * `xgboost` (predict.py:6, train.py) is not installed -> empty stub module.
  * `PROJECT_DATA_PATH` (settings.py:8-12) -> a temp dir with the gunzipped example CSVs.
"""
import os
import shutil
import sys
import tempfile
import types

REFERENCE_ROOT = '/root/reference'
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def _staged():
    return os.path.isfile(os.path.join(STAGED_ROOT, 'doppelspeller', 'match_maker.py')) and \
        os.path.isfile(os.path.join(STAGED_ROOT, 'example_dataset', 'example_truth.csv'))


def reference_available():
    """True when the staged copy exists, or can be made (the build container holds /root/reference)."""
    if _staged():
        return True
    if os.path.isdir(os.path.join(REFERENCE_ROOT, 'doppelspeller')):
        try:
            from oracle import stage_reference
            stage_reference.stage()
        except Exception:
            return False
        return _staged()
    return False


def _indel_ratio_py(a, b):
    """ratio = (la + lb - indel) / (la + lb), 1.0 when both empty (python-levenshtein semantics)."""
    la, lb = len(a), len(b)
    if la + lb == 0:
        return 1.0
    prev = [0] * (lb + 1)
    for i in range(1, la + 1):
        cur = [0] * (lb + 1)
        ai = a[i - 1]
        for j in range(1, lb + 1):
            if ai == b[j - 1]:
                cur[j] = prev[j - 1] + 1
            else:
                cur[j] = prev[j] if prev[j] >= cur[j - 1] else cur[j - 1]
        prev = cur
    lcs = prev[lb]
    return (la + lb - (la + lb - 2 * lcs)) / (la + lb)


_DATA_DIR = None


def stage_example_data():
    """A private copy of the staged example CSVs (settings.py:18,22,36-37; the reference writes its outputs beside
    them) - returns the directory `PROJECT_DATA_PATH` must point at."""
    global _DATA_DIR
    if _DATA_DIR is not None:
        return _DATA_DIR
    target = tempfile.mkdtemp(prefix='ds_ref_data_')
    source = os.path.join(STAGED_ROOT, 'example_dataset')
    for name in os.listdir(source):
        if name.endswith('.csv'):
            shutil.copyfile(os.path.join(source, name), os.path.join(target, name))
    _DATA_DIR = target
    return target


def _prepare():
    if not reference_available():
        raise RuntimeError('the reference is not staged under oracle/_ref (run oracle/stage_reference.py in the build container)')
    if 'Levenshtein' not in sys.modules:
        lev = types.ModuleType('Levenshtein')
        lev.ratio = _indel_ratio_py
        sys.modules['Levenshtein'] = lev
    if 'xgboost' not in sys.modules:
        sys.modules['xgboost'] = types.ModuleType('xgboost')
    os.environ['PROJECT_DATA_PATH'] = stage_example_data()
    if STAGED_ROOT not in sys.path:
        sys.path.insert(0, STAGED_ROOT)


def import_match_maker():
    """Only settings / constants / common / match_maker (skips feature_engineering's eager 11 s gufunc compile)."""
    _prepare()
    import doppelspeller.settings as settings
    import doppelspeller.constants as constants
    import doppelspeller.common as common
    import doppelspeller.match_maker as match_maker
    return types.SimpleNamespace(settings=settings, constants=constants, common=common, match_maker=match_maker)


def import_reference():
    """Returns the imported reference modules as a namespace (common, match_maker, feature_engineering,
    predict, settings, constants)."""
    _prepare()
    import doppelspeller.settings as settings
    import doppelspeller.constants as constants
    import doppelspeller.common as common
    import doppelspeller.match_maker as match_maker
    import doppelspeller.feature_engineering as feature_engineering
    import doppelspeller.predict as predict
    return types.SimpleNamespace(
        settings=settings, constants=constants, common=common, match_maker=match_maker,
        feature_engineering=feature_engineering, predict=predict)
