"""Drop-in `MatchMaker` (reference: /root/reference/doppelspeller/match_maker.py:74-203).

Same constructor, same `get_closest_matches(row_number) -> list[title_id]`, same `top_n` attribute and
the same exception when fewer than `top_n` rows exist - but the IDF-weighted Jaccard scan and the
top-n selection run on the GPU (csrc/ds_topn.cu) for ALL rows of `data` at once: the reference's
caller asks one row at a time (predict.py:126-127, feature_engineering_prepare.py:37-43), so the
first call computes every row and later calls are served from the cached [Q, top_n] result.

The host side below only assigns column ids and idf weights (match_maker.py:91-96,135-153).  Column
ids and the order in which a truth row's weights are summed depend on python set iteration order in
the reference (:144-147, :172-174); iterating the very same set objects in the same process
reproduces both, which is what makes the results bit-identical.
"""
import logging
import math
from collections import Counter
from itertools import chain

import numpy as np
import pandas as pd

from . import _native as nat
from .index import TruthIndex

LOGGER = logging.getLogger(__name__)

COLUMN_N_GRAMS = 'n_grams'      # constants.py:6
COLUMN_TITLE_ID = 'title_id'    # constants.py:2
COLUMN_TRANSFORMED_TITLE = 'transformed_title'   # constants.py:4
TOO_FEW_ROWS_MESSAGE = 'top_matches.shape[0] != self.top_n'   # match_maker.py:189


def _document_frequencies(n_gram_sets):
    """common.py:145-147 get_n_grams_counter: document frequency over per-title n-gram sets.  The same sequence as the
    reference's `Counter(x for y in n_grams for x in set(y))` (first-seen key order feeds the column numbering), produced
    by C-level iterators instead of a generator frame per element."""
    return Counter(chain.from_iterable(map(set, n_gram_sets)))


def _rows_to_csr(rows, encoding):
    """CSR of the rows with every set walked in its own iteration order (match_maker.py:172-174)."""
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(np.fromiter(map(len, rows), dtype=np.int64, count=len(rows)), out=ptr[1:])
    cols = np.fromiter(map(encoding.__getitem__, chain.from_iterable(rows)), dtype=np.uint16, count=int(ptr[-1]))
    return ptr, cols


class MatchMaker:
    """
    :param data: (dataframe) titles to find the closest truth titles for (column `n_grams`: set of str)
    :param truth_data: (dataframe) the "truth" database (columns `n_grams`, `title_id`)
    :param top_n: (int) number of nearest titles to fetch
    :param device: CUDA device ordinal (default: torch's current device)
    :param mx_mode: how `max_intersection_possible` is summed (match_maker.py:197): the default follows
        CPython >= 3.12's compensated builtin sum(), DS_MX_NAIVE the plain sum of older interpreters
    :param order: 'reference' (default) numbers the n-gram columns and sums each truth title's weights in the
        reference's python-set iteration order - bit-identical to a reference MatchMaker of the same process;
        'canonical' builds the whole index on the GPU from the `transformed_title` column (csrc/ds_encode.cu,
        milliseconds instead of seconds): same arithmetic, hash-seed independent column order, so scores may
        differ from one particular reference process in the last float bits (like two reference runs with
        different PYTHONHASHSEED do)
    """

    def __init__(self, data, truth_data, top_n, device=None, mx_mode=nat.DS_MX_PY312_COMPENSATED, order='reference'):
        self.top_n = top_n
        self.mx_mode = mx_mode
        LOGGER.info(f'[{self.__class__.__name__}] Loading pre-requisite data!')
        if order == 'canonical':
            self._init_canonical(data, truth_data, device)
            LOGGER.info(f'[{self.__class__.__name__}] Loaded pre-requisite data!')
            return
        if order != 'reference':
            raise ValueError("order must be 'reference' or 'canonical'")
        data_n_grams = list(data[COLUMN_N_GRAMS])
        truth_n_grams = list(truth_data[COLUMN_N_GRAMS])

        self.n_grams_counter = _document_frequencies(data_n_grams)                 # :91
        self.n_grams_counter_truth = _document_frequencies(truth_n_grams)          # :92
        self.number_of_truth_titles = len(truth_data)                              # :93
        n_truth = self.number_of_truth_titles
        self.idf_s_mapping = {key: math.log(n_truth / count)                       # :94, :135-142
                              for key, count in self.n_grams_counter_truth.items()}
        self.max_idf_value = max(self.idf_s_mapping.values())                      # :95
        all_n_grams = set(list(self.n_grams_counter.keys()) + list(self.n_grams_counter_truth.keys()))   # :144-147
        self.n_grams_decoding = {index: n_gram for index, n_gram in enumerate(all_n_grams)}
        self.n_grams_encoding = {v: k for k, v in self.n_grams_decoding.items()}
        if len(self.n_grams_encoding) > 65535:
            raise Exception(f'{len(self.n_grams_encoding)} distinct n-grams: more than the 65535 u16 column ids')
        idf64 = np.array([self.idf_s_mapping.get(self.n_grams_decoding[i], self.max_idf_value)   # :151, :180-181
                          for i in range(len(self.n_grams_decoding))], dtype=np.float64)

        t_ptr, t_cols = _rows_to_csr(truth_n_grams, self.n_grams_encoding)         # :167-178
        q_ptr, q_cols = _rows_to_csr(data_n_grams, self.n_grams_encoding)          # :155-165
        self.truth_data = truth_data.loc[:, [COLUMN_TITLE_ID]]                     # :104
        self._init_device(idf64, t_ptr, t_cols, q_ptr, q_cols, device)
        LOGGER.info(f'[{self.__class__.__name__}] Loaded pre-requisite data!')

    @classmethod
    def from_encoded(cls, idf64_by_col, t_row_ptr, t_col_ids, q_row_ptr, q_col_ids, title_ids, top_n, device=None,
                     mx_mode=nat.DS_MX_PY312_COMPENSATED, sums=None):
        """Builds a MatchMaker from already assigned column ids (fixtures, synthetic data, the canonical
        encoder in `encode.py`): truth / query CSR with u16 column ids, idf per column, title id per truth row."""
        self = cls.__new__(cls)
        self.top_n = top_n
        self.mx_mode = mx_mode
        self.number_of_truth_titles = int(t_row_ptr.shape[0]) - 1
        self.truth_data = None
        self._title_ids = np.asarray(title_ids)
        self._init_device(np.ascontiguousarray(idf64_by_col, dtype=np.float64), t_row_ptr, t_col_ids, q_row_ptr, q_col_ids,
                          device, sums=sums)
        return self

    def _init_canonical(self, data, truth_data, device):
        import torch
        from . import encode
        if not torch.cuda.is_available():
            raise nat.DoppelSpellerError(-2, 'no CUDA device available (doppelspeller_b200 has no CPU fallback)')
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.number_of_truth_titles = len(truth_data)
        enc = encode.encode_canonical_device(list(data[COLUMN_TRANSFORMED_TITLE]), list(truth_data[COLUMN_TRANSFORMED_TITLE]),
                                             device=self.device)
        self.truth_data = truth_data.loc[:, [COLUMN_TITLE_ID]]
        self.n_grams_decoding = None          # decoded lazily: vocab_codes holds the canonical trigram codes
        self._vocab_codes = enc['vocab_codes']
        self._index = TruthIndex(enc['t_ptr'], enc['t_cols'], enc['idf64'], device=self.device)
        self._q_ptr, self._q_cols = enc['q_ptr'], enc['q_cols']
        self._rows = self._count = self._flags = None

    def _init_device(self, idf64, t_ptr, t_cols, q_ptr, q_cols, device, sums=None):
        import torch
        if not torch.cuda.is_available():
            raise nat.DoppelSpellerError(-2, 'no CUDA device available (doppelspeller_b200 has no CPU fallback)')
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._index = TruthIndex(t_ptr, t_cols, idf64, sums=sums, device=self.device)
        self._q_ptr = np.ascontiguousarray(q_ptr, dtype=np.int64)
        self._q_cols = np.ascontiguousarray(q_cols, dtype=np.uint16)
        self._rows = None
        self._count = None
        self._flags = None

    @property
    def sums_matrix_truth(self):
        """match_maker.py:100,:174 - computed on the device in the reference's accumulation order."""
        return self._index.sums()

    def _compute_all(self):
        rows, count, _, flags = self._index.topn(self._q_ptr, self._q_cols, self.top_n, mx_mode=self.mx_mode,
                                                 with_details=True)
        if hasattr(rows, 'cpu'):              # device-resident queries (canonical order): results to the host once
            rows, count, flags = rows.cpu().numpy(), count.cpu().numpy(), flags.cpu().numpy()
        self._rows, self._count, self._flags = rows, count, flags
        self._rows_top_n = self.top_n
        self._ids = None

    def closest_rows(self):
        """All queries at once: (truth row indexes int64[Q, top_n] in descending row order, count[Q]).
        `top_n` is read on every call like the reference does (match_maker.py:187): changing it recomputes."""
        if self._rows is None or getattr(self, '_rows_top_n', None) != self.top_n:
            self._compute_all()
        return self._rows, self._count

    def _title_ids_of(self, top_matches):
        if self.truth_data is not None:
            return self.truth_data.loc[top_matches, COLUMN_TITLE_ID].tolist()      # match_maker.py:190
        return self._title_ids[top_matches].tolist()

    def get_closest_matches(self, row_number):
        """Given the "row_number" of the data, gets the closest (self.top_n) titles in the truth data
        (match_maker.py:192-203)."""
        rows, count = self.closest_rows()
        if count[row_number] != self.top_n:
            raise Exception(TOO_FEW_ROWS_MESSAGE)
        if self._ids is None:
            # truth_data.loc[top_matches, 'title_id'] of match_maker.py:190 for every row at once (one gather instead of a
            # pandas label lookup per call: 140 us -> 2 us per call at top_n = 100)
            if self.truth_data is not None:
                by_label = self.truth_data[COLUMN_TITLE_ID]
                positional = isinstance(by_label.index, pd.RangeIndex) and by_label.index.start == 0 and by_label.index.step == 1
                lookup = by_label.to_numpy() if positional else None
            else:
                lookup = self._title_ids
            if lookup is None:
                return self._title_ids_of(rows[row_number])
            self._ids = lookup[np.where(rows >= 0, rows, 0)]
        return self._ids[row_number].tolist()

    def get_closest_matches_batch(self):
        """[Q, top_n] title ids for every row of the data (vectorised form of the caller's loop)."""
        rows, count = self.closest_rows()
        if (count != self.top_n).any():
            raise Exception(TOO_FEW_ROWS_MESSAGE)
        flat = rows.reshape(-1)
        if self.truth_data is not None:
            ids = self.truth_data.loc[flat, COLUMN_TITLE_ID].to_numpy()
        else:
            ids = self._title_ids[flat]
        return ids.reshape(rows.shape)
