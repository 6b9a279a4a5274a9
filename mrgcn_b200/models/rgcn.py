"""RGCN: stack of GraphConvolution layers — drop-in for `mrgcn.models.rgcn.RGCN`
(/root/reference/mrgcn/models/rgcn.py:11-132).  Same constructor, attributes (`layers`, `activations`,
`num_layers`, `relations`), parameter names and forward dispatch; ReLU and the row-wise "node dropout"
mask are fused into the layer kernel's epilogue."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn.functional import dropout

from ..data.batch import A_Batch, getAdjacencyNodeColumnIdx
from ..layers.graph import GraphConvolution


class RGCN(nn.Module):
    def __init__(self, modules, num_relations, num_nodes, num_bases, p_dropout, featureless, bias,
                 link_prediction):
        super().__init__()
        assert len(modules) > 0
        self.num_nodes = num_nodes
        self.p_dropout = p_dropout
        self.layers = nn.ModuleDict()
        self.activations = nn.ModuleDict()
        # layer_0 is the input layer (identity term, optionally featureless); the rest are hidden layers (rgcn.py:25-50)
        for k, (indim, outdim, _ltype, f_activation) in enumerate(modules):
            name = "layer_%d" % k
            self.layers[name] = GraphConvolution(indim=indim, outdim=outdim, num_relations=num_relations,
                                                 num_nodes=num_nodes, num_bases=num_bases, bias=bias,
                                                 input_layer=(k == 0), featureless=(featureless and k == 0))
            self.activations[name] = f_activation
        self.num_layers = len(self.layers)
        if link_prediction:
            # DistMult relation embeddings, one row per relation block (rgcn.py:54-61)
            self.relations = nn.Parameter(torch.empty((num_relations, modules[-1][1])))
            self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.relations)

    def forward(self, X, A):
        if isinstance(A, A_Batch) or type(A).__name__ == "A_Batch":   # ours (host or device built) or the reference's
            return self._forward_mini_batch(X, A)
        return self._forward_full_batch(X, A)

    def _row_mask(self, n):
        # rgcn.py:78-84: dropout applied to a vector of ones (functional default training=True, CPU RNG stream)
        if self.p_dropout > 0.0:
            return dropout(torch.ones(n), p=self.p_dropout)
        return None

    def _layer_step(self, layer, f_activation, X, A, A_idx, n_rows):
        mask = self._row_mask(n_rows)
        fuse_relu = isinstance(f_activation, nn.ReLU)
        X = layer(X, A, A_idx, row_mask=mask, relu=fuse_relu)
        if f_activation is not None and not fuse_relu:
            X = f_activation(X)
        return X

    def _forward_full_batch(self, X, A):
        for layer, f_activation in zip(self.layers.values(), self.activations.values()):
            if type(layer) is GraphConvolution:
                X = self._layer_step(layer, f_activation, X, A, None, self.num_nodes)
            else:
                X = layer(X)
                if f_activation is not None:
                    X = f_activation(X)
        return X

    def _forward_mini_batch(self, X, A):
        for layer_idx, (layer, f_activation) in enumerate(zip(self.layers.values(), self.activations.values())):
            i = self.num_layers - (layer_idx + 1)        # most distant nodes first (rgcn.py:101-102)
            A_slices = A.row[i]
            if layer.input_layer and layer.featureless:
                X = self._layer_step(layer, f_activation, None, A_slices, None, A_slices.shape[0])
            else:
                A_idx = getAdjacencyNodeColumnIdx(A.neighbours[i], layer.num_nodes, layer.num_relations)
                X = self._layer_step(layer, f_activation, X, A_slices, A_idx, A_slices.shape[0])
        return X
