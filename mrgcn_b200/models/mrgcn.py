"""MRGCN: drop-in for `mrgcn.models.mrgcn.MRGCN` (/root/reference/mrgcn/models/mrgcn.py:25-305).

Same constructor signature, attributes read by the task loops (`devices`, `gate_weights`, `gate_map`,
`module_dict`, `rgcn`) and forward(batch) contract.  On the hot path: the gated scatter of modality
embeddings into the node-feature matrix (mrgcn.py:250-305) and the RGCN stack.  The modality encoders
themselves are out of scope: the numeric/temporal MLP is mirrored because it is
three lines; string / image / geometry encoders are taken from the reference package when it is
importable and refused otherwise.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn

from ..data.batch import MiniBatch
from .rgcn import RGCN


class MLP(nn.Module):
    """Literal encoder for numeric / temporal vectors (mrgcn/models/perceptron.py:6-46)."""

    def __init__(self, input_dim, output_dim, num_layers=3, p_dropout=0.0, bias=True):
        super().__init__()
        self.input_dim, self.output_dim, self.p_dropout = input_dim, output_dim, p_dropout
        step = (input_dim - output_dim) // num_layers
        dims = [output_dim + i * step for i in reversed(range(num_layers))]
        seq, d_in = [], input_dim
        for d in dims:
            seq += [nn.Linear(d_in, d, bias), nn.Dropout(p=p_dropout, inplace=True), nn.ReLU()]
            d_in = d
        self.mlp = nn.Sequential(*seq)
        for p in self.parameters():
            nn.init.uniform_(p)

    def forward(self, X):
        return self.mlp(X)


def _reference_encoder(datatype, args):
    try:
        if datatype in ("xsd.string", "xsd.anyURI"):
            from mrgcn.models.transformer import Transformer
            from mrgcn.models.utils import loadFromHub
            cfg, dim_out, drop = args
            return Transformer(loadFromHub(cfg), output_dim=dim_out, p_dropout=drop), dim_out, -1
        if datatype == "blob.image":
            from mrgcn.models.imagecnn import ImageCNN
            from mrgcn.models.utils import loadFromHub
            cfg, _tcfg, dim_out, drop = args
            return ImageCNN(loadFromHub(cfg), output_dim=dim_out, p_dropout=drop), dim_out, -1
        if datatype == "ogc.wktLiteral":
            from mrgcn.models.temporal_cnn import TCNN
            nrows, dim_out, size, drop = args
            m = TCNN(features_in=nrows, features_out=dim_out, p_dropout=drop, size=size)
            return m, dim_out, m.minimal_length
    except ImportError as e:   # pragma: no cover
        raise NotImplementedError("encoder for %s lives in the reference package (out of scope here): %s"
                                  % (datatype, e))
    raise Exception("Datatype not supported: " + datatype)


class MRGCN(nn.Module):
    def __init__(self, modules, embedding_modules, num_relations, num_nodes, num_bases=-1, p_dropout=0.0,
                 featureless=False, bias=False, link_prediction=False, gcn_gpu_acceleration=False, gated=True):
        super().__init__()
        assert len(modules) > 0
        self.num_nodes = num_nodes
        self.p_dropout = p_dropout
        self.module_dict = nn.ModuleDict()
        self.devices = dict()
        self.gate_map = dict()
        self.modality_modules = dict()
        self.modality_out_dim = 0
        self.compute_modality_embeddings = False
        self.im_norm = None
        counters = dict()
        i_gate = 0
        for datatype, args, gpu_acceleration in embedding_modules:
            if datatype in ("xsd.boolean", "xsd.numeric"):
                ncols, dim_out, drop = args
                module, seq_length, group = MLP(ncols, dim_out, num_layers=1, p_dropout=drop), -1, "num"
            elif datatype in ("xsd.date", "xsd.dateTime", "xsd.gYear"):
                ncols, dim_out, drop = args
                module, seq_length, group = MLP(ncols, dim_out, num_layers=2, p_dropout=drop), -1, "temp"
            else:
                module, dim_out, seq_length = _reference_encoder(datatype, args)
                group = {"xsd.string": "llm", "xsd.anyURI": "llm", "blob.image": "img"}.get(datatype, "geo")
                if datatype == "blob.image":
                    tcfg = args[1]
                    if "mean" in tcfg and "std" in tcfg:
                        from mrgcn.encodings.blob.image import Normalizer
                        self.im_norm = Normalizer(tcfg["mean"], tcfg["std"])
            k = counters.get(group, 0)
            counters[group] = k + 1
            mod_name = datatype.replace(".", "_") + "_" + str(k)       # mrgcn.py:69,80,94,108,121
            self.module_dict[mod_name] = module
            self.modality_modules.setdefault(datatype, []).append((module, seq_length, dim_out, i_gate))
            self.modality_out_dim += dim_out
            self.compute_modality_embeddings = True
            self.gate_map[mod_name] = i_gate
            i_gate += 1
            device = torch.device("cpu")
            if gpu_acceleration:
                if torch.cuda.is_available():
                    device = torch.device("cuda")
                else:
                    warnings.warn("CUDA Resource not available", ResourceWarning)
            self.devices[datatype] = device
            module.to(device)

        self.gate_weights = torch.ones(i_gate)
        if gated and i_gate > 0:
            self.gate_weights = nn.Parameter(torch.mul(self.gate_weights, 0.1))     # mrgcn.py:151-154
        else:
            self.gate_weights.requires_grad = False

        # One process per GPU under torchrun: the relational part is node-partitioned (mrgcn_b200/partition.py).  The ranges
        # are fixed here, by node count, because the task loops build their optimizer from the parameters right after
        # construction (node_classification.py:26-45) - before the model has seen the graph.
        import torch.distributed as dist
        self.partitioned = bool(dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
        if self.partitioned:
            from ..partition import PartitionedRGCN, equal_bounds
            assert p_dropout == 0.0, "node dropout is not supported by the partitioned model"
            self.rgcn = PartitionedRGCN(modules, num_relations, num_nodes, num_bases, featureless, bias, link_prediction,
                                        equal_bounds(num_nodes, dist.get_world_size()), dist.get_rank())
            self.rgcn.enable_grad_hooks()
        else:
            self.rgcn = RGCN(modules, num_relations, num_nodes, num_bases, p_dropout, featureless, bias, link_prediction)

        # The kernels are CUDA-only: the relational part always lives on the GPU.  `gcn_gpu_acceleration`
        # (node_classification.py:390-392; never forwarded by link_prediction.py:466-471) is accepted and
        # implied; without a device the model can be built (state_dict work) but not run.
        device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        if device.type == "cpu":
            warnings.warn("mrgcn_b200: no CUDA device - the model can be constructed but not run", ResourceWarning)
        self.devices["relational"] = device
        self.rgcn.to(device)
        self.X_device = device if all(d.type == device.type for d in self.devices.values()) else torch.device("cpu")
        from ..upload import FeaturePrefetcher
        self._prefetcher = FeaturePrefetcher()

    # ------------------------------------------------------------------------------------------
    def prefetch(self, batch):
        """Start the host -> device copy of batch.X[0] for a LATER forward(batch) on a copy stream (mrgcn_b200/upload.py), so
        that it overlaps the step that is computing now.  Optional: without it forward uploads in line, as the reference does
        (mrgcn.py:199-204).  Only the encoder-free feature path is prefetched (with modality encoders the matrix is assembled
        on the device anyway).  Returns True when a copy was started."""
        if self.compute_modality_embeddings or self.rgcn.layers["layer_0"].featureless:
            return False
        X = self._own_rows(torch.as_tensor(batch.X[0]))
        if X.device.type != "cpu":
            return False
        return self._prefetcher.start(X, self.devices["relational"])

    def upload(self, batch):
        """The device copy of batch.X[0] (the rank's rows under a node partition): the one prefetch() started, else a copy
        made now.  For callers that feed a captured step (mrgcn_b200/stepping.py) through a static input buffer:
        `X_static.copy_(model.upload(batch)); step()`."""
        return self._upload_features(self._own_rows(torch.as_tensor(batch.X[0])), self.devices["relational"])

    def _own_rows(self, X):
        """The rows of the host matrix this process uploads: all of them, or the rank's range when layer 0 is
        source-partitioned."""
        if self.partitioned and self.rgcn.layer0_is_source_partitioned():
            return X[self.rgcn.lay.lo:self.rgcn.lay.hi]
        return X

    def forward(self, batch):
        if isinstance(batch, MiniBatch) or type(batch).__name__ == "MiniBatch":
            return self._forward(batch, batch.A.neighbours[-1])
        return self._forward(batch, None)

    def _forward(self, batch, outer_idx):
        X, F = batch.X[0], batch.X[1:]
        rgcn_device = self.devices["relational"]
        X_dev = None
        if self.compute_modality_embeddings:
            batch_idx = torch.arange(self.num_nodes) if outer_idx is None else torch.as_tensor(outer_idx)
            XF = self._compute_modality_embeddings(F, batch_idx)
            X = torch.as_tensor(X).to(XF.device)
            X_dev = torch.cat([X.to(XF.dtype), XF], dim=1).to(rgcn_device)           # mrgcn.py:199-204
        elif not self.rgcn.layers["layer_0"].featureless and not self.partitioned:
            # extension: pre-computed node features handed over as batch.X[0] (BASELINE.json config 3);
            # the reference has no encoder-free feature path (mrgcn.py:192-207 leaves X_dev = None)
            X_dev = self._upload_features(torch.as_tensor(X), rgcn_device)
        if self.partitioned:
            return self._forward_partitioned(X if X_dev is None else X_dev, batch.A)
        if X_dev is not None:
            X_dev = X_dev.float()
        return self.rgcn(X_dev, batch.A)

    def _forward_partitioned(self, X, A):
        """Node-partitioned forward: every rank hands over the SAME batch (as the reference's single process would); the rank
        builds its share of the graph on first sight of A, uploads only the feature rows it needs and returns the logits of
        ALL nodes (true node order), so that the caller's indexing with global node ids keeps working."""
        rg, dev = self.rgcn, self.devices["relational"]
        lay = rg.lay
        if rg.gF is None or getattr(self, "_part_A", None) is not A:
            idx = A._indices().to(dev)
            rg.set_graph(idx[0], idx[1], A._values().to(dev).float())
            self._part_A = A
        if rg.layers["layer_0"].featureless:
            return rg.forward_all(None)
        X_in = self._upload_features(self._own_rows(torch.as_tensor(X)), dev).float()          # own rows only, or all
        if not rg.layer0_is_source_partitioned():
            X_in = lay.to_padded(X_in)
        return rg.forward_all(X_in)

    def _upload_features(self, X, dev):
        """Host feature matrix -> device, every call (as mrgcn.py:203-204 does).  One contiguous DMA (a pitched 2-D copy of
        604-byte rows was measured at 5 GB/s, the contiguous one runs at the PCIe rate); the layer pads the rows on the
        device for the projection kernel's tensor-map loads (csrc/feat_proj.cu: k_pad_rows).  A copy started earlier by
        prefetch() is used instead when there is one."""
        if X.device.type == "cpu":
            ahead = self._prefetcher.take(X, dev)
            if ahead is not None:
                return ahead
        return X.to(dev, non_blocking=True)

    def _compute_modality_embeddings(self, F, batch_idx):
        """mrgcn.py:250-305: gate * encoder(data) scattered into the rows of the nodes that carry the modality."""
        # the feature matrix is assembled where the relational part lives: every modality block is written by one fused
        # gate-scale + row-scatter kernel (mrgcn_b200/optim.py: gated_scatter) instead of mul + masked assignment
        dev = self.devices["relational"] if self.devices["relational"].type == "cuda" else self.X_device
        X = torch.zeros((len(batch_idx), self.modality_out_dim), dtype=torch.float32, device=dev)
        batch_idx = torch.as_tensor(batch_idx)
        offset = 0
        for datatype, encoding_sets, _ in F:
            if datatype not in self.modality_modules:
                continue
            for i, (encodings, node_idx, _) in enumerate(encoding_sets):
                module, _, out_dim, i_gate = self.modality_modules[datatype][i]
                if torch.isclose(self.gate_weights[i_gate].detach().cpu(), torch.tensor(0.)):
                    offset += out_dim
                    continue
                node_idx = torch.as_tensor(node_idx)
                F_mask = torch.isin(node_idx, batch_idx)       # == isin(node_idx, intersect1d(node_idx, batch_idx))
                if not bool(F_mask.any()):
                    offset += out_dim
                    continue
                X_mask = torch.isin(batch_idx, node_idx)
                enc = torch.as_tensor(encodings)
                mod_dev = self.devices[datatype]
                if datatype in ("xsd.string", "xsd.anyURI"):
                    data = enc[F_mask].int()
                elif datatype == "blob.image":
                    data = self.im_norm.normalize_(enc[F_mask])
                else:
                    data = enc[F_mask].float()
                out = module(data.to(mod_dev))
                if dev.type == "cuda":
                    from ..optim import gated_scatter
                    X = gated_scatter(X, out, torch.nonzero(X_mask).flatten(), self.gate_weights[i_gate], offset)
                else:
                    out = torch.mul(out, self.gate_weights[i_gate].to(out.device))
                    X[X_mask.to(dev), offset:offset + out_dim] = out.to(dev)
                offset += out_dim
        return X
