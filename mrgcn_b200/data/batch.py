"""Batch containers with the attribute layout the reference's task loops and models rely on
(/root/reference/mrgcn/data/batch.py).  Only what sits on the hot path is mirrored: the adjacency
hand-off (FullBatch.as_tensors_, :144-149), A_Batch's k-hop row slices (:168-226) and the column helpers
(:245-263).  Padding / subsetting of raw literal encodings is out of scope."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch


def scipy_sparse_to_pytorch_sparse(sp_input, dtype=torch.float32):
    """/root/reference/mrgcn/data/utils.py:165-170 — COO indices from `.nonzero()`, values cast to `dtype`
    (int8 in the reference's batches, which truncates 1/deg to 0 for deg >= 2)."""
    indices = np.array(sp_input.nonzero())
    return torch.sparse_coo_tensor(torch.LongTensor(indices), torch.Tensor(sp_input.data), sp_input.shape,
                                   dtype=dtype)


class Batch:
    A = None
    X = None
    node_index = None
    device = None

    def __init__(self, batch_node_idx=None):
        self.device = torch.device("cpu")
        if batch_node_idx is not None:
            self.node_index = np.copy(batch_node_idx)

    def as_tensors_(self):
        if self.node_index is not None and not torch.is_tensor(self.node_index):
            self.node_index = torch.from_numpy(np.asarray(self.node_index))
        if self.X is not None and not torch.is_tensor(self.X[0]):
            self.X[0] = torch.from_numpy(np.asarray(self.X[0]))

    def to(self, devices):
        return self


class FullBatch(Batch):
    def __init__(self, A=None, X=None, batch_node_idx=None, value_dtype=torch.int8):
        super().__init__(batch_node_idx)
        self.value_dtype = value_dtype
        if A is not None:
            self.A = A
        if X is not None:
            self.X = X

    def as_tensors_(self):
        super().as_tensors_()
        if sp.issparse(self.A):
            self.A = scipy_sparse_to_pytorch_sparse(self.A, dtype=self.value_dtype)


class A_Batch:
    """k-hop row slices of A for a batch of target nodes (batch.py:161-226)."""
    node_index = None
    neighbours = None
    row = None
    device = None

    def __init__(self, A=None, batch_idx=None, num_layers=0):
        self.neighbours, self.row = [], []
        self.device = torch.device("cpu")
        if batch_idx is not None:
            self.node_index = np.copy(batch_idx)
        if A is not None:
            sample_idx = self.node_index
            for _ in range(num_layers):
                self.row.append(A[sample_idx])
                sample_idx = getNeighboursSparse(A, sample_idx)
                self.neighbours.append(sample_idx)

    def as_tensors_(self):
        self.node_index = torch.from_numpy(self.node_index)
        self.row = [scipy_sparse_to_pytorch_sparse(a, dtype=torch.int8) for a in self.row]
        self.neighbours = [torch.from_numpy(a) for a in self.neighbours]


class DeviceABatch(A_Batch):
    """A_Batch built on the GPU from a device-resident RelGraph (mrgcn_b200.graph): frontier expansion
    (`getNeighboursSparse`, batch.py:228-243) and row slices (`A[sample_idx]`, :190) are gathers over the destination-major
    edge order, no Python loop over nodes and no scipy.  Values are truncated to int8 exactly as A_Batch.as_tensors_ does
    (batch.py:223-226) unless value_dtype says otherwise."""

    def __init__(self, graph, batch_idx, num_layers, value_dtype=torch.int8):
        self.neighbours, self.row = [], []
        self.device = graph.device
        self.node_index = torch.as_tensor(batch_idx, device=graph.device).long()
        sample = self.node_index
        for _ in range(num_layers):
            self.row.append(graph.row_slice(sample, value_dtype))
            sample = graph.neighbours(sample)
            self.neighbours.append(sample)

    def as_tensors_(self):
        pass


class MiniBatch(Batch):
    def __init__(self, A=None, X=None, batch_node_idx=None, num_layers=None):
        super().__init__(batch_node_idx)
        if A is not None:
            self.A = A_Batch(A, self.node_index, num_layers)
            if X is not None:
                outer = self.A.neighbours[-1]
                self.X = [np.asarray(X[0])[outer]] + list(X[1:])

    def as_tensors_(self):
        super().as_tensors_()
        self.A.as_tensors_()


def getNeighboursSparse(A, idx):
    """batch.py:228-243 — union of the source nodes of the rows in idx, irrespective of relation."""
    assert isinstance(A, sp.csr_matrix)
    num_nodes = A.shape[0]
    cols = [A.indices[A.indptr[i]:A.indptr[i + 1]] for i in idx]
    return np.unique(np.concatenate(cols) % num_nodes) if len(cols) else np.empty(0, dtype=np.int64)


def getAdjacencyNodeColumnIdx(idx, num_nodes, num_relations):
    """batch.py:245-250 — column ids r*N + i for every relation r and node i in idx (vectorised)."""
    idx = torch.as_tensor(idx).long()
    rel = torch.arange(num_relations, dtype=torch.int64, device=idx.device).view(-1, 1)
    return (rel * num_nodes + idx.view(1, -1)).reshape(-1)
