"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and refuses to compute
without a GPU (no CPU fallback)."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT


def _declared_symbols():
    names = set()
    for header in glob.glob(os.path.join(ROOT, 'include', '*.h')):
        text = open(header).read()
        text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
        names.update(re.findall(r'\b(ds_[a-z0-9_]+)\s*\(', text))
    return names


def test_library_exports_every_declared_symbol():
    import doppelspeller_b200._native as nat
    declared = _declared_symbols()
    assert declared, 'no declarations found in include/*.h'
    assert declared == set(nat.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} is declared in include/ but not exported'
    assert nat.lib.ds_version() == 100
    assert nat.topn_retained(10) == 64 and nat.topn_retained(100) == 160


def test_library_targets_sm_100a_only():
    import doppelspeller_b200._native as nat
    import subprocess
    out = subprocess.run(['/usr/local/cuda/bin/cuobjdump', '-lelf', nat.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    archs = set(re.findall(r'sm_\d+a?', out.stdout))
    assert archs == {'sm_100a'}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import doppelspeller_b200._native as nat
    from doppelspeller_b200 import feature_engineering as fe
    from doppelspeller_b200.match_maker import MatchMaker
    a = np.zeros((2, 255), dtype=np.uint8)
    with pytest.raises(nat.DoppelSpellerError):
        fe.fast_levenshtein_ratio_batch(a, a, np.array([3, 3], np.uint8), np.array([3, 3], np.uint8))
    with pytest.raises(nat.DoppelSpellerError):
        MatchMaker.from_encoded(np.ones(4), np.array([0, 2], np.int64), np.array([0, 1], np.uint16),
                                np.array([0, 1], np.int64), np.array([1], np.uint16), np.arange(1), 1)


def test_product_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, 'doppelspeller_b200', '**', '*.py'), recursive=True):
        text = open(path).read()
        assert 'oracle' not in re.sub(r'#.*', '', text).replace('"""', ''), f'{path} mentions the oracle'
    for path in glob.glob(os.path.join(ROOT, 'doppelspeller_b200', 'csrc', '*')):
        assert 'oracle' not in open(path, errors='replace').read(), f'{path} mentions the oracle'


# ---------------------------------------------------------------------------------------------------
# the boundary from plain C: tests/abi_client.c (ISO C99, no Python / torch / C++ on its side of the ABI)
# ---------------------------------------------------------------------------------------------------
def _run_c_client(tmp_path, library):
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    directory, name = os.path.dirname(os.path.abspath(library)), os.path.basename(library)
    binary = os.path.join(str(tmp_path), 'abi_client')
    build = subprocess.run([gcc, '-std=c99', '-pedantic', '-Wall', '-Wextra', '-Werror', '-I', os.path.join(ROOT, 'include'),
                            os.path.join(ROOT, 'tests', 'abi_client.c'), '-o', binary, f'-L{directory}', f'-l:{name}',
                            f'-Wl,-rpath,{directory}', '-lm'], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-3000:]
    return subprocess.run([binary], capture_output=True, text=True, timeout=600)


def test_c_client_compiles_as_c99_and_gets_no_cpu_fallback(tmp_path):
    """The header is ISO C, the library links with nothing but itself, and without a GPU a C caller gets DS_ERR_CUDA with a
    message (exit status 3 of the client) - never an answer computed on the CPU."""
    import torch

    import doppelspeller_b200._native as nat
    run = _run_c_client(tmp_path, nat.LIB_PATH)
    if torch.cuda.is_available():
        assert run.returncode == 0, run.stdout + run.stderr
    else:
        assert run.returncode == 3 and 'status -2' in run.stdout and 'no CPU fallback' in run.stdout, run.stdout + run.stderr


def test_c_client_answers_on_the_emulated_kernels(tmp_path):
    """The same C program against the host emulation of the kernels (tests/emu): top-n rows, InDel ratio and the 66
    features of the reference's docstring pair come back right through plain C."""
    from tests.emu import build as emu_build
    run = _run_c_client(tmp_path, emu_build.build())
    assert run.returncode == 0 and 'construct_features ok' in run.stdout, run.stdout + run.stderr


@pytest.mark.gpu
def test_c_client_answers_on_the_gpu(tmp_path):
    import doppelspeller_b200._native as nat
    run = _run_c_client(tmp_path, nat.LIB_PATH)
    assert run.returncode == 0 and 'construct_features ok' in run.stdout, run.stdout + run.stderr
