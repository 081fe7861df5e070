"""Gradient-boosted tree inference on the GPU: the replacement of
`self.model.predict(xgb.DMatrix(features), ntree_limit=...)` in /root/reference/doppelspeller/predict.py:229-233.

The reference trains and pickles an xgboost==0.90 Booster (train.py:99-121, requirements.txt:8).  xgboost is not
part of the reference tree and is not installed here, so `GbdtModel` takes the trees as plain arrays; a maintainer
exports them from the Booster once:

    model = GbdtModel.from_xgboost_dump(booster.get_dump(dump_format='json'), base_score=0.5, objective='reg:logistic',
                                        ntree_limit=booster.best_ntree_limit)

(`get_dump` prints split thresholds with limited decimal precision in 0.90; exporting from the binary model keeps
every bit.)  Parity with a real model file is unpinned - no xgboost, no model file, nothing to train with; the
arithmetic follows xgboost 0.90's CPU predictor (the CPU restatement the tests compare against lives outside this package).
"""
import json
import math

import numpy as np

from . import _native as nat

NODE_DTYPE = np.dtype([('feature', '<i4'), ('value', '<f4'), ('yes', '<u2'), ('no', '<u2'), ('missing', '<u2'), ('reserved', '<u2')])
MARGIN, LOGISTIC = 0, 1
PREDICTION_PROBABILITY_THRESHOLD = 0.9      # settings.py:76


class GbdtModel:
    """nodes: structured array (NODE_DTYPE), children as node indexes inside their tree and after their parent;
    tree_offsets int32[n_trees + 1]; base_margin float; transform MARGIN / LOGISTIC."""

    def __init__(self, nodes, tree_offsets, base_margin=0.0, transform=LOGISTIC):
        self.nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        self.tree_offsets = np.ascontiguousarray(tree_offsets, dtype=np.int32)
        self.base_margin = float(np.float32(base_margin))
        self.transform = int(transform)
        self._device_copy = {}
        self.n_features_needed = self._validate()

    def _validate(self):
        """Every walk must stay inside its tree and end: children inside the tree and after their parent, at most 65,535
        nodes per tree (u16 child indexes).  Checked once on the host - the kernel walks device copies unchecked.
        Returns 1 + the largest feature index a split uses."""
        offsets, nodes = self.tree_offsets, self.nodes
        if offsets.ndim != 1 or offsets.shape[0] < 1 or offsets[0] != 0 or (np.diff(offsets) < 0).any() or int(offsets[-1]) != nodes.shape[0]:
            raise ValueError('tree_offsets must start at 0, be non-decreasing and end at len(nodes)')
        sizes = np.diff(offsets).astype(np.int64)
        if sizes.size and int(sizes.max()) > 65535:
            raise ValueError(f'a tree has {int(sizes.max())} nodes: more than the 65,535 a u16 child index can address')
        if sizes.size and (sizes == 0).any():
            raise ValueError('empty tree')
        local = np.arange(nodes.shape[0], dtype=np.int64) - np.repeat(offsets[:-1].astype(np.int64), sizes)
        size_of = np.repeat(sizes, sizes)
        split = nodes['feature'] >= 0
        for name in ('yes', 'no', 'missing'):
            child = nodes[name].astype(np.int64)
            bad = split & ((child <= local) | (child >= size_of))
            if bad.any():
                at = int(np.nonzero(bad)[0][0])
                raise ValueError(f'node {at}: `{name}` child {int(child[at])} is not after its parent inside the tree')
        if (nodes['feature'] < -1).any():
            raise ValueError('feature index below -1')
        return int(nodes['feature'].max()) + 1 if nodes.shape[0] else 0

    @property
    def n_trees(self):
        return int(self.tree_offsets.shape[0]) - 1

    @classmethod
    def from_trees(cls, trees, base_margin=0.0, transform=LOGISTIC):
        """trees: list of lists of (feature, value, yes, no, missing) tuples, feature -1 = leaf (then value = weight)."""
        offsets = np.zeros(len(trees) + 1, dtype=np.int32)
        np.cumsum([len(t) for t in trees], out=offsets[1:])
        nodes = np.zeros(int(offsets[-1]), dtype=NODE_DTYPE)
        i = 0
        for tree in trees:
            for feature, value, yes, no, missing in tree:
                nodes[i] = (feature, value, yes, no, missing, 0)
                i += 1
        return cls(nodes, offsets, base_margin, transform)

    @classmethod
    def from_xgboost_dump(cls, dumped_trees, base_score=0.5, objective='reg:logistic', ntree_limit=0, feature_names=None):
        """dumped_trees: `booster.get_dump(dump_format='json')` - one JSON document per tree with nested `children`,
        `nodeid`, `split` ("f12" or a feature name), `split_condition`, `yes`, `no`, `missing` and `leaf`."""
        names = {name: i for i, name in enumerate(feature_names)} if feature_names else {}
        if ntree_limit:
            dumped_trees = dumped_trees[:ntree_limit]
        trees = []
        for text in dumped_trees:
            flat = {}
            stack = [json.loads(text) if isinstance(text, str) else text]
            while stack:
                node = stack.pop()
                flat[int(node['nodeid'])] = node
                stack.extend(node.get('children', []))
            # renumber in breadth-first order from the root so that children follow their parent
            order, queue = [], [0]
            while queue:
                nodeid = queue.pop(0)
                order.append(nodeid)
                node = flat[nodeid]
                if 'leaf' not in node:
                    queue.extend([int(node['yes']), int(node['no'])])
            position = {nodeid: i for i, nodeid in enumerate(order)}
            tree = []
            for nodeid in order:
                node = flat[nodeid]
                if 'leaf' in node:
                    tree.append((-1, float(node['leaf']), 0, 0, 0))
                else:
                    split = node['split']
                    feature = names[split] if split in names else int(str(split).lstrip('f'))
                    tree.append((feature, float(node['split_condition']), position[int(node['yes'])], position[int(node['no'])],
                                 position[int(node['missing'])]))
            trees.append(tree)
        logistic = objective in ('reg:logistic', 'binary:logistic')
        base_margin = math.log(base_score / (1.0 - base_score)) if logistic else base_score
        return cls.from_trees(trees, base_margin, LOGISTIC if logistic else MARGIN)

    def predict(self, features):
        """features: float32 [n_rows, n_features] numpy array or CUDA tensor -> predictions float32 [n_rows] of the same kind."""
        n_rows, n_features = int(features.shape[0]), int(features.shape[1])
        if n_features < self.n_features_needed:
            raise ValueError(f'the model splits on feature {self.n_features_needed - 1}, the matrix has {n_features} columns')
        if hasattr(features, 'data_ptr'):
            import torch
            nat.expect(features, 'float32', 'features')
            device = features.device
            if device not in self._device_copy:
                self._device_copy[device] = (torch.as_tensor(self.nodes.view(np.uint8).reshape(-1, 16).copy()).to(device),
                                             torch.as_tensor(self.tree_offsets).to(device))
            nodes, offsets = self._device_copy[device]
            out = torch.empty(n_rows, dtype=torch.float32, device=device)
            with torch.cuda.device(device):
                nat.check(nat.lib.ds_gbdt_predict(nat.ptr(features.contiguous()), n_rows, n_features, nat.ptr(nodes), nat.ptr(offsets),
                                                  self.n_trees, self.base_margin, self.transform, nat.ptr(out),
                                                  torch.cuda.current_stream(device).cuda_stream))
            return out
        features = np.ascontiguousarray(features, dtype=np.float32)
        out = np.empty(n_rows, dtype=np.float32)
        nat.check(nat.lib.ds_gbdt_predict(nat.ptr(features), n_rows, n_features, nat.ptr(self.nodes), nat.ptr(self.tree_offsets),
                                          self.n_trees, self.base_margin, self.transform, nat.ptr(out), nat.current_stream()))
        return out


def select_model_matches(test_index, predictions, threshold=PREDICTION_PROBABILITY_THRESHOLD):
    """predict.py:242-249: positions of the pairs kept as matches - the prediction equals the maximum of its test_index,
    exceeds the probability threshold, and no other pair of that test_index shares it (_remove_duplicated_matches)."""
    from .predict import select_close_matches
    return select_close_matches(test_index, predictions, threshold=threshold)
