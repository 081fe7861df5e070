"""ctypes binding of the C ABI declared in include/doppelspeller_b200.h.

There is NO fallback: if the CUDA library has not been built (`python -c "import __graft_entry__ as g;
g.build()"` or `make -C doppelspeller_b200/csrc`) importing this module raises, and every compute
call raises `DoppelSpellerError` when no sm_100 device is usable.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DOPPELSPELLER_B200_LIB: explicit path of the built library (deployments that keep it elsewhere, kernel A/B builds)
LIB_PATH = os.environ.get('DOPPELSPELLER_B200_LIB') or os.path.join(_HERE, '_lib', 'libdoppelspeller_b200.so')

DS_OK = 0
DS_FLAG_RESCAN = 1
DS_FLAG_FEW_POSITIVE = 2
DS_MX_PY312_COMPENSATED = 0
DS_MX_NAIVE = 1
N_WORDS = 15
N_FEATURES = 66
MAX_TITLE = 255

# every symbol include/doppelspeller_b200.h declares (tests check that the library exports them all)
EXPORTED_SYMBOLS = (
    'ds_version', 'ds_last_error', 'ds_kernel_launches', 'ds_trim', 'ds_profile_begin', 'ds_profile_end', 'ds_profile_end_split', 'ds_transform_titles',
    'ds_index_create', 'ds_index_destroy', 'ds_index_get_sums',
    'ds_topn', 'ds_topn_retained', 'ds_topn_local', 'ds_topn_local_shared', 'ds_topn_merge', 'ds_topn_rescan',
    'ds_indel_ratio_u8', 'ds_indel_ratio_pairs', 'ds_levenshtein_ratio_pairs', 'ds_prematch_pairs', 'ds_select_close_matches',
    'ds_construct_features', 'ds_construct_features_pairs',
    'ds_encode_max_vocab', 'ds_encode_trigrams', 'ds_title_features', 'ds_gbdt_predict',
)


class DoppelSpellerError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f'doppelspeller_b200 native call failed (status {status}): {message}')
        self.status = status


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} is missing: build the CUDA extension first (python -c "import __graft_entry__ as g; '
        f'g.build()" or make -C doppelspeller_b200/csrc).  doppelspeller_b200 has no CPU fallback.')

lib = ctypes.CDLL(LIB_PATH)

_vp, _i64, _i32, _u8, _u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint8, ctypes.c_uint32

lib.ds_version.restype = ctypes.c_int
lib.ds_last_error.restype = ctypes.c_char_p
lib.ds_kernel_launches.restype = _i64
lib.ds_trim.argtypes = [ctypes.c_int]
lib.ds_topn_retained.restype = _i32
lib.ds_topn_retained.argtypes = [_i32]
lib.ds_encode_max_vocab.restype = _i32
lib.ds_encode_trigrams.argtypes = [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_i32),
                                   ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.c_int, _vp]
lib.ds_title_features.argtypes = [_vp, _vp, _i64, _vp, _vp, ctypes.c_int, _vp]
lib.ds_transform_titles.argtypes = [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, ctypes.c_int, _vp]
lib.ds_gbdt_predict.argtypes = [_vp, _i64, _i32, _vp, _vp, _i32, ctypes.c_float, _i32, _vp, _vp]
lib.ds_index_create.argtypes = [ctypes.POINTER(_vp), ctypes.c_int, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _vp]
lib.ds_index_destroy.argtypes = [_vp]
lib.ds_index_get_sums.argtypes = [_vp, _vp, _vp]
lib.ds_topn.argtypes = [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]
lib.ds_topn_local.argtypes = [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]
lib.ds_topn_local_shared.argtypes = [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, ctypes.POINTER(_vp), _i32, _vp]
lib.ds_topn_merge.argtypes = [_i32, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp]
lib.ds_topn_rescan.argtypes = [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]
lib.ds_indel_ratio_u8.argtypes = [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp]
lib.ds_indel_ratio_pairs.argtypes = [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp]
lib.ds_levenshtein_ratio_pairs.argtypes = [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp]
lib.ds_prematch_pairs.argtypes = [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i32, _vp, _vp]
lib.ds_select_close_matches.argtypes = [_vp, _vp, _i64, _i32, _i32, _vp, _vp]
lib.ds_construct_features.argtypes = [_vp, _vp, _vp, _vp, _i64, _vp, _u8, _u32, _i64, _vp, _vp]
lib.ds_construct_features_pairs.argtypes = [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _u8, _u32, _i64, _vp, _vp]
for _name in EXPORTED_SYMBOLS:
    if _name not in ('ds_last_error', 'ds_kernel_launches', 'ds_topn_retained', 'ds_version', 'ds_encode_max_vocab'):
        getattr(lib, _name).restype = ctypes.c_int


def check(status):
    if status != DS_OK:
        raise DoppelSpellerError(status, lib.ds_last_error().decode('utf-8', 'replace'))


def kernel_launches():
    return int(lib.ds_kernel_launches())


def profile_begin():
    check(lib.ds_profile_begin())


def profile_end():
    """-> (summed k_scan device milliseconds, k_scan launches, (query, truth) pairs scanned)"""
    ms, launches, pairs = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_double(0)
    check(lib.ds_profile_end(ctypes.byref(ms), ctypes.byref(launches), ctypes.byref(pairs)))
    return ms.value, launches.value, pairs.value


def profile_end_split():
    """-> {'k_scan': (device ms, launches, pairs), 'k_post': (...)}: the two forms of the K1 scan, timed separately"""
    ms, launches, pairs = (ctypes.c_double * 2)(), (ctypes.c_int64 * 2)(), (ctypes.c_double * 2)()
    check(lib.ds_profile_end_split(ms, launches, pairs))
    return {name: (ms[i], launches[i], pairs[i]) for i, name in enumerate(('k_scan', 'k_post'))}


def topn_retained(k):
    return int(lib.ds_topn_retained(int(k)))


def ptr(array):
    """Address of a numpy array (host) or a torch tensor (host or device); None -> NULL."""
    if array is None:
        return None
    if isinstance(array, np.ndarray):
        if not array.flags['C_CONTIGUOUS']:
            raise ValueError('numpy arguments must be C-contiguous')
        return array.ctypes.data
    if hasattr(array, 'data_ptr'):
        if not array.is_contiguous():
            raise ValueError('tensor arguments must be contiguous')
        return array.data_ptr()
    raise TypeError(f'unsupported buffer type {type(array)!r}')


def current_stream():
    """cudaStream_t of torch's current stream (so torch work and native kernels stay ordered)."""
    import torch
    if not torch.cuda.is_available():
        return None
    return torch.cuda.current_stream().cuda_stream


def stream_for(*arrays):
    """cudaStream_t the call must be ordered on: torch's current stream OF THE DEVICE that owns the first CUDA tensor
    among `arrays` (not of torch's current device - the native entry points run where their buffers live)."""
    import torch
    for array in arrays:
        if hasattr(array, 'is_cuda') and array.is_cuda:
            return torch.cuda.current_stream(array.device).cuda_stream
    return current_stream()


def current_device():
    """ordinal of torch's current CUDA device (0 when no GPU is usable: the native call then reports DS_ERR_CUDA)"""
    import torch
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


def expect(array, dtype, name):
    kind = str(array.dtype).replace('torch.', '')
    if kind != dtype:
        raise TypeError(f'{name} must have dtype {dtype}, got {kind}')
    return array
