"""TEST INFRASTRUCTURE ONLY: pytest plugin of the subprocess that runs the `gpu` parity tests against the HOST
EMULATION of the kernels (tests/emu/cuda_emu.h; library chosen through DOPPELSPELLER_B200_LIB).

The emulated "device" takes host buffers only, so torch is told that one CUDA device exists (the drop-in classes refuse
to start without one - there is no CPU fallback in the product) and tests that need real CUDA tensors are deselected.
"""
import os

import pytest

NEEDS_CUDA_TENSORS = (
    # tests/test_gpu_parity.py
    'test_device_resident_inputs_match_host_inputs', 'test_construct_features_device_resident_padded_layout',
    'test_gpu_trigram_encoder_matches_host_encoder', 'test_transform_titles_device_table_matches_reference',
    'test_gbdt_predict_device_resident_features',
    # tests/test_gpu_dataframe_api.py
    'test_matchmaker_canonical_order_gpu_index_build', 'test_candidate_pipeline_end_to_end', 'test_candidate_pipeline_from_raw_titles',
    'test_pipeline_predict_follows_reference_selection', 'test_indexed_prematch_matches_per_pair_form',
    'test_title_features_on_the_device',
    # tests/test_gpu_full_size.py: BASELINE sizes, minutes of CPU time
    'test_c3_properties_and_sampled_parity', 'test_c4_sampled_pairs',
)


class _Stream:
    cuda_stream = 0


def pytest_configure(config):
    library = os.environ.get('DOPPELSPELLER_B200_LIB', '')
    if 'libds_emu' not in os.path.basename(library):
        raise pytest.UsageError('tests.emu.plugin is only for runs against the emulated library (DOPPELSPELLER_B200_LIB)')
    import torch
    torch.cuda.is_available = lambda: True
    torch.cuda.current_device = lambda: 0
    torch.cuda.current_stream = lambda device=None: _Stream()


def pytest_collection_modifyitems(config, items):
    keep, drop = [], []
    for item in items:
        (drop if item.name.split('[')[0] in NEEDS_CUDA_TENSORS else keep).append(item)
    if drop:
        config.hook.pytest_deselected(items=drop)
        items[:] = keep
