"""mrgcn_b200 — B200-native R-GCN / DistMult hot path of wxwilcke/mrgcn behind a C ABI (include/mrgcn_b200.h).

Host-side mirror of the reference interface: `layers.graph.GraphConvolution`, `models.rgcn.RGCN`,
`models.mrgcn.MRGCN`, `tasks.link_prediction.{score_distmult_bc, negative_samples, compute_ranks_fast}`,
`data.batch.{FullBatch, MiniBatch, A_Batch}`; `graph.RelGraph` (device edge orders), `partition.PartitionedRGCN`
(1-D node partition over NCCL), `dropin.install()` (rebinds the reference's import paths).
CUDA only: there is no CPU fallback."""

__version__ = "0.1.0"
