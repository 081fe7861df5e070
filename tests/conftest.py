import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def _built():
    """The native pieces are built once per session (no-op when up to date)."""
    if 'libds_emu' in os.path.basename(os.environ.get('DOPPELSPELLER_B200_LIB', '')):
        return      # subprocess of tests/test_emulated_kernels.py: it brings its own (host-emulated) library, and nvcc
                    # must not run under the sanitizer runtime that process preloads
    import __graft_entry__ as entry
    entry.build()


@pytest.fixture(scope='session')
def golden_matchmaker():
    return dict(np.load(os.path.join(GOLDEN, 'matchmaker_example.npz')))


@pytest.fixture(scope='session')
def golden_pairs():
    return dict(np.load(os.path.join(GOLDEN, 'pairs.npz')))


@pytest.fixture(scope='session')
def golden_topk():
    return dict(np.load(os.path.join(GOLDEN, 'topk_vectors.npz')))


@pytest.fixture(scope='session')
def example_titles():
    raw = np.load(os.path.join(GOLDEN, 'example_titles.npz'))
    return dict(truth_titles=[str(t) for t in raw['truth_titles']], truth_title_ids=raw['truth_title_ids'],
                test_titles=[str(t) for t in raw['test_titles']], test_index=raw['test_index'])


@pytest.fixture(scope='session')
def golden_transform():
    """(raw titles, reference transform_title outputs) minted by tests/golden/make_transform_golden.py"""
    g = np.load(os.path.join(GOLDEN, 'transform_titles.npz'))
    ib, io, ob, oo = g['in_bytes'], g['in_off'], g['out_bytes'], g['out_off']
    titles = [ib[io[i]:io[i + 1]].tobytes().decode('utf-8', 'surrogatepass') for i in range(len(io) - 1)]
    outputs = [ob[oo[i]:oo[i + 1]].tobytes().decode('latin-1') for i in range(len(oo) - 1)]
    return titles, outputs


def oracle_index_from_golden(g):
    """Finished oracle index from the golden fixture's explicit column ids / set order."""
    from oracle import oracle
    w64 = g['w64']
    enc = dict(n_truth=int(g['t_ptr'].shape[0]) - 1, w64=w64, w32=w64.astype(np.float32), t_ptr=g['t_ptr'],
               t_cols=g['t_cols'].astype(np.int32), q_ptr=g['q_ptr'], q_cols=g['q_cols'].astype(np.int32))
    return oracle.finish_index(enc)


def oracle_index_from_encoded(enc):
    from oracle import oracle
    return oracle.finish_index(dict(
        n_truth=int(enc['t_ptr'].shape[0]) - 1, w64=enc['idf64'], w32=enc['idf64'].astype(np.float32),
        t_ptr=enc['t_ptr'], t_cols=enc['t_cols'].astype(np.int32), q_ptr=enc['q_ptr'], q_cols=enc['q_cols'].astype(np.int32)))


def features_equal(got, want, rtol=1e-6):
    """Integer features (columns 0..35) exact; idf / rank columns within rtol with identical NaN / inf masks."""
    got, want = np.asarray(got), np.asarray(want)
    nan_ok = np.isnan(got) == np.isnan(want)
    inf_ok = (np.isinf(got) == np.isinf(want)) & (~np.isinf(got) | (np.sign(got) == np.sign(want)))
    exact = (got[:, :36] == want[:, :36]) | (np.isnan(got[:, :36]) & np.isnan(want[:, :36]))
    with np.errstate(all='ignore'):
        close = np.isclose(got[:, 36:], want[:, 36:], rtol=rtol, atol=0, equal_nan=True)
    return bool(nan_ok.all() and inf_ok.all() and exact.all() and close.all())
