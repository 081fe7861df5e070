"""GPU: the drop-in proof.  The reference's own `Prediction.generate_test_predictions` (predict.py:274-317 - the code
behind `generate-predictions`, cli.py:52-61, and `closest-search-single-title`, cli.py:64-83) is run from the staged
copy under oracle/_ref twice in this process: unmodified, and with ONLY the three imports of INTEGRATION.md section 1
swapped for doppelspeller_b200's.  Candidate lists, pre-match ratios, the feature matrix handed to the model, the
predictions frame and the output file must come out the same (BASELINE configs C1 and C2)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ref():
    from oracle import dropin, ref_import
    if not ref_import.reference_available():
        pytest.skip('the reference is not staged under oracle/_ref (python oracle/stage_reference.py)')
    return dropin.load_reference()


def _assert_same(report):
    bad = {k: v for k, v in report.items() if k not in ('rows', 'pairs', 'feature_rows') and v != 0}
    assert not bad, report


def test_generate_predictions_with_swapped_imports(ref):
    """C1 on the first 1,200 rows of example_test.csv against all 30,000 truth titles, top_n = 100."""
    from oracle import dropin
    reference_run = dropin.run_prediction(ref, patched=False, n_test_rows=1200)
    patched_run = dropin.run_prediction(ref, patched=True, n_test_rows=1200)
    report = dropin.compare_runs(reference_run, patched_run)
    assert report['rows'] > 1000 and report['pairs'] == report['rows'] * 100 and report['feature_rows'] > 50000
    _assert_same(report)


def test_single_title_search_with_swapped_imports(ref):
    """C2: closest-search-single-title - a one-row MatchMaker over the whole truth DB, features of its 100 candidates."""
    from oracle import dropin
    title = 'graet expectatoins minstries intl'
    reference_run = dropin.run_prediction(ref, patched=False, title=title)
    patched_run = dropin.run_prediction(ref, patched=True, title=title)
    report = dropin.compare_runs(reference_run, patched_run)
    assert report['rows'] == 1 and report['feature_rows'] == 100
    assert reference_run['single']['match_title_id'] == 13672
    _assert_same(report)


def test_changed_top_n_is_honoured(ref):
    """match_maker.py:187 reads self.top_n on every call: changing it after the first call recomputes."""
    import numpy as np
    from doppelspeller_b200.match_maker import MatchMaker
    c = ref.constants
    truth = ref.common.get_ground_truth()
    test = ref.common.get_test_data().iloc[:50].copy()
    ours = MatchMaker(test.copy(), truth.copy(), 100)
    theirs = ref.match_maker.MatchMaker(test.copy(), truth.copy(), 100)
    for top_n in (100, 10, 25):
        ours.top_n = theirs.top_n = top_n
        for row in range(0, 50, 7):
            assert ours.get_closest_matches(row) == theirs.get_closest_matches(row)
    assert np.array_equal(ours.sums_matrix_truth.view(np.uint32), theirs.sums_matrix_truth.view(np.uint32))
    assert c.COLUMN_TITLE_ID in truth
