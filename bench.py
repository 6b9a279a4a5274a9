#!/usr/bin/env python
"""bench.py — R-GCN fwd+bwd edges/s on the shapes BASELINE.json names (default: AM), 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--shape am|aifb|aifb_b40|synth|fb15k237|yago3-10+]

A step = one full-batch training step of the configured model WITHOUT the optimizer.
  node classification (am, aifb, synth): forward of the 2-layer R-GCN (am.toml: 151 -> 10 -> 11, 40 bases, identity + feature
      terms in layer 0), cross-entropy on the labelled nodes, backward to every parameter;
  link prediction (fb15k237, yago3-10+): 1-layer encoder (h = 200, 2 bases, ReLU) over the full graph, DistMult scores of 500
      positives + 100 in-batch negatives, BCE, backward (/root/reference/mrgcn/tasks/link_prediction.py:244-326); the line
      also carries `lp.rank_scores_per_s` for compute_ranks_fast on 500 facts (:593-643).
edges = nnz of the stacked adjacency (forward + inverse + self-loop blocks).  Prints ONE JSON line (rank 0).

  value    : device-resident throughput (features already in HBM), CUDA events, max over ranks; the K timed steps are
             replays of ONE captured CUDA graph of the step (static in full-batch training); --no-graph times eager launches
  e2e      : the same step through the public module call `MRGCN.forward(batch)` with the feature matrix in pinned HOST
             memory (copied to the device every step, as the reference's forward does, mrgcn/models/mrgcn.py:203-204) and
             the loss read back to the host; under torchrun the same call runs the node-partitioned model (every rank
             uploads only the feature rows it owns)
  roofline : dominant kernel of the step, timed live with CUDA events inside the library (mrgcn_profile_enable) in an
             eager pass of the same K steps, against MEASURED_PEAKS.json; `traffic` = its DRAM bytes per launch from the
             ncu capture recorded in profiles/ncu_traffic.json
  cpu_baseline / --impl reference : the UNMODIFIED reference modules (vendored by __graft_entry__.build() into the
             git-ignored baseline/_ref; `kind: "reference"`) - or the oracle port when that install is absent (`kind:
             "port"`) - on a sample of the same workload sized to the host's free RAM and to a time budget
  parity_vs_n1 (N > 1): loss and small-parameter gradients of the partitioned step against the single-GPU model with the
             same weights, computed on rank 0 inside the run (max relative error)
  collectives (N > 1): name, bytes, calls per step and milliseconds of every collective of the step

N > 1 (torchrun): the graph is 1-D node-partitioned (mrgcn_b200/partition.py); total work is fixed ("scaling": "strong").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

METRIC = "rgcn_fwd_bwd_edges_per_s"
UNIT = "edges/s"
NUM_LABELLED = 10000
LP_POS = 500          # fb15k-237.toml / yago3-10+.toml: test_batchsize = 500 positives per step, + 20 % negatives
SURVEY_B_PER_EDGE = {"am": 5012.0, "fb15k237": 4891.0}      # per-edge-gather formulation of the step


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_ram_gb():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable"):
                    return int(ln.split()[1]) / 1e6
    except OSError:
        pass
    return 16.0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
def make_workload(shape_name, scale, seed=1):
    from mrgcn_b200.synth import SHAPES, synth_graph
    shp = SHAPES[shape_name]
    n, tr = synth_graph(shp, seed=seed, scale=scale)
    return shp, n, tr


def labelled_nodes(n, num_classes, seed=1):
    rng = np.random.default_rng(seed + 7)
    idx = np.sort(rng.choice(n, size=min(n, NUM_LABELLED), replace=False))
    return idx, rng.integers(0, num_classes, size=len(idx))


def lp_batch(tr, seed=1):
    """One training batch of the LP loop: the first LP_POS triples + in-batch negatives made exactly as
    link_prediction.py:247-264 does (host NumPy RNG), labels 1 / 0."""
    from mrgcn_b200.tasks.link_prediction import negative_samples
    data = tr[:LP_POS].astype(np.int64)
    corrupted, Y = negative_samples(data, np.random.RandomState(seed))
    trip = np.concatenate([data, np.asarray(corrupted, dtype=np.int64)], 0)
    return trip, Y


def modules_of(shp):
    dims = shp.dims
    if shp.task == "lp":      # link_prediction.py:449-464: ReLU on every layer, the last one included
        return [(dims[k], dims[k + 1], "mrgcn", nn.ReLU()) for k in range(len(dims) - 1)]
    return [(dims[k], dims[k + 1], "mrgcn", nn.ReLU() if k + 2 < len(dims) else None) for k in range(len(dims) - 1)]


def algorithmic_bytes(g, in0, dims, B, fused):
    """Algorithmic bytes per launch of the kernels of one step: 4 B per structure word read, table / feature
    rows counted once per pass where a kernel keeps them on chip and per edge where it gathers them, every output written
    once.  Returns {kernel: [bytes per launch...]} in launch order.  g: sizes of the (rank's) graph and work plan."""
    E, ND, NS, R = g["E"], g["ND"], g["NS"], g["R"]
    nch, ntask, npc, nblk = g["n_chunks"], g.get("n_tasks", 0), g.get("n_pieces", 0), g.get("n_blks", 0)
    h = dims[0]
    c = dims[1] if len(dims) > 1 else None
    kp = (in0 + 31) // 32 * 32
    out = {}
    ms = lambda d: 4 if d <= 4 else 8 if d <= 8 else (d + 15) // 16 * 16      # padded message row (mrgcn_msg_stride)

    def add(k, v):
        out.setdefault(k, []).append(float(v))
    Bn = max(B, 0)
    if Bn:
        if fused:
            add("vcat_split", Bn * in0 * h * 4 + 2 * Bn * h * kp * 4)
            add("feat_proj", NS * kp * 4 + NS * Bn * h * 4 + 2 * Bn * h * kp * 4)
        add("tab_msg_fwd", Bn * NS * h * 4 * (2 if fused else 1) + E * 8 + ntask * 8 + NS * 4 + E * h * 4 + R * Bn * 4 * (2 if fused else 1))
        add("tab_bwd_w", E * (12 + h * 4) + NS * 8 + Bn * NS * h * 4 + R * Bn * 4)
        add("tab_bwd_c", Bn * NS * h * 4 + E * (8 + h * 4) + ntask * 8 + E * 4 + npc * (4 + Bn * 4))
        # the tile-staging pair (default for narrow outputs): same algorithmic work; ident_bwd_c's E x B scratch (`cbuf`) and
        # its reduction are implementation traffic, not algorithmic bytes, so they lower its achieved figure
        add("ident_bwd_w", E * (12 + h * 4) + NS * 8 + Bn * NS * h * 4 + R * Bn * 4)
        add("ident_bwd_c", Bn * NS * h * 4 + E * (12 + h * 4) + NS * 4 + R * Bn * 4)
        # one pass for both (ident_bwd.cu): table read once, gradient written once, edges and gact rows once; its scratch rows
        # (E x B, written in E3 order and streamed by comp_chunk_reduce) are again implementation traffic
        add("ident_bwd_fused", 2 * Bn * NS * h * 4 + E * (16 + h * 4) + NS * 4 + R * Bn * 4)
        add("comp_block_reduce", npc * (4 + Bn * 4) + nblk * Bn * 4)
        add("comp_reduce", nblk * Bn * 4 + R * Bn * 4)
    if in0 and not fused:
        add("feat_msg_fwd", E * (8 + in0 * 4) + R * in0 * h * 4 + E * ms(h) * 4)
    n_msgs0 = (1 if Bn else 0) + (1 if (in0 and not fused) else 0)
    add("agg_fwd", ND * 4 + E * n_msgs0 * (4 + ms(h) * 4) + (E * (12 + h * 4) if not Bn else 0) + ND * h * 4)
    if in0:
        add("feat_bwd_w", E * (12 + in0 * 4 + h * 4) + nch * in0 * h * 4)
        add("feat_w_reduce", nch * in0 * h * 4 + R * in0 * h * 4)
    add("act_bwd", 3 * ND * h * 4)
    if c is not None:
        add("feat_msg_fwd", E * (8 + h * 4) + R * h * c * 4 + E * ms(c) * 4)
        add("agg_fwd", ND * 4 + E * (4 + ms(c) * 4) + ND * c * 4)
        add("act_bwd", 2 * ND * c * 4)
        add("feat_bwd_w", E * (12 + h * 4 + c * 4) + nch * h * c * 4)
        add("feat_w_reduce", nch * h * c * 4 + R * h * c * 4)
        add("feat_bwd_x_msg", E * (8 + c * 4) + R * h * c * 4 + E * ms(h) * 4)
        add("feat_bwd_x_agg", NS * 4 + E * (4 + ms(h) * 4) + NS * h * 4)
    return out


# ---------------------------------------------------------------------------------------------------------
def _reference_modules():
    """The unmodified reference (baseline/_ref, installed by __graft_entry__.build()) with the import-only rdflib stand-in
    of tests/golden/_stubs; None when the install is absent."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "mrgcn")):
        return None
    for p in (os.path.join(ROOT, "tests", "golden", "_stubs"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        from mrgcn.models.rgcn import RGCN
        from mrgcn.tasks import link_prediction as lp
        return RGCN, lp
    except Exception as exc:      # pragma: no cover
        print("bench: reference import failed (%s); using the oracle port" % exc, file=sys.stderr)
        return None


def cpu_reference_step(shape_name, scale, steps, warmup, threads):
    """One training step (no optimizer) of the reference's CPU path on the workload scaled by `scale` (N and triples
    together; R, bases and layer widths unchanged).  Returns (nnz, [seconds per step], N, kind)."""
    from oracle import reference_port as rp
    torch.set_num_threads(threads)
    shp, n, tr = make_workload(shape_name, scale)
    R, B = shp.num_relations, shp.num_bases if shp.num_bases > 0 else -1
    dims = shp.dims
    fl = dims[0] == 0
    A = rp.csr_to_coo(rp.as_float32(rp.stacked_adjacency(tr, n, shp.num_props)), torch.float32)
    nnz = A._nnz()
    mods = modules_of(shp)
    ref = _reference_modules()
    torch.manual_seed(1)
    X = torch.randn(n, dims[0]) if not fl else None
    if shp.task == "nc":
        idx, y = labelled_nodes(n, dims[-1])
        idx, y = torch.from_numpy(idx), torch.from_numpy(y)
    else:
        trip, Y = lp_batch(tr)
        trip = torch.from_numpy(trip)
    if ref is not None:
        RefRGCN, ref_lp = ref
        model = RefRGCN(mods, R, n, B, 0.0, fl, False, shp.task == "lp")
        params = list(model.parameters())

        def step():
            for p in params:
                p.grad = None
            out = model(X, A)
            if shp.task == "nc":
                loss = nn.functional.cross_entropy(out[idx], y)
            else:
                sc = ref_lp.score_distmult_bc((trip[:, 0], trip[:, 1], trip[:, 2]), out, model.relations)
                loss = nn.functional.binary_cross_entropy_with_logits(sc, Y)
            loss.backward()
        kind = "reference"
    else:
        layers, rel = rp.init_rgcn_params(mods, R, n, B, fl, False, shp.task == "lp")
        acts = ["relu" if m[3] is not None else None for m in mods]
        params = [v for l in layers for v in l.values()] + ([rel] if rel is not None else [])
        for v in params:
            v.requires_grad_(True)

        def step():
            for p in params:
                p.grad = None
            out = rp.rgcn_forward(layers, acts, X, A, num_nodes=n, num_relations=R, num_bases=B, featureless=fl)
            if shp.task == "nc":
                loss = rp.nc_loss(out, idx, y)
            else:
                loss = rp.lp_loss(rp.distmult_score((trip[:, 0], trip[:, 1], trip[:, 2]), out, rel), Y)
            loss.backward()
        kind = "port"
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return nnz, times, n, kind


def pick_cpu_scale(shape_name, budget_s, steps, threads):
    """Largest power-of-two fraction of the workload whose reference step fits the host (free RAM: the reference holds
    ~7 tensors of R*N*out floats, measured 4.2 GB at AM/16) and the time budget (cost is ~linear in N: calibrated with
    one step at a small scale)."""
    from mrgcn_b200.synth import SHAPES
    shp = SHAPES[shape_name]
    free = host_ram_gb()
    copies = 7 if shp.task == "nc" else 4
    probe = 1.0 / 64 if shp.num_nodes > 100000 else 1.0 / 8
    nnz, times, _, _ = cpu_reference_step(shape_name, probe, 1, 1, threads)
    rate = nnz / times[0]                                    # edges/s at the probe scale
    full_nnz = shp.nnz
    scale = 1.0
    while scale > probe:
        need_gb = copies * 4e-9 * shp.num_relations * shp.num_nodes * scale * max(shp.dims[1:])
        t_est = full_nnz * scale / rate * (steps + 1)
        if need_gb < 0.6 * free and t_est < budget_s:
            break
        scale /= 2
    return max(scale, probe), free


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 3))
    warm = 1
    scale, free = (args.cpu_sample_scale, host_ram_gb()) if args.cpu_sample_scale else pick_cpu_scale(args.shape, 150.0, steps, threads)
    scale *= args.scale
    nnz, times, n, kind = cpu_reference_step(args.shape, scale, steps, warm, threads)
    ms = 1e3 * float(np.mean(times))
    val = nnz / (ms / 1e3)
    sample = ("%s-shape scaled x%g (N=%d, nnz=%d; R, bases, dims unchanged), 1 warm-up + %d timed step(s), host RAM free %.0f GB"
              % (args.shape, scale, n, nnz, len(times), free))
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": len(times),
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.shape), "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(shape):
    from mrgcn_b200.synth import SHAPES
    s = SHAPES[shape]
    if s.task == "lp":
        return ("%s-shape link prediction: N=%d, R=%d, nnz=%d (synthetic power-law graph), R-GCN encoder %s (ReLU), %d bases, "
                "full batch, DistMult on %d positives + %d negatives, BCE" % (shape.upper(), s.num_nodes, s.num_relations, s.nnz,
                                                                              "->".join(str(d) for d in s.dims), s.num_bases,
                                                                              LP_POS, LP_POS // 5))
    return ("%s-shape node classification: N=%d, R=%d, nnz=%d (synthetic power-law graph), R-GCN %s, %d bases, "
            "full batch, CE on %d labelled nodes" % (shape.upper(), s.num_nodes, s.num_relations, s.nnz,
                                                     "->".join(str(d) for d in s.dims), s.num_bases, NUM_LABELLED))


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default="am")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink N and triples together (debug)")
    ap.add_argument("--cpu-sample-scale", type=float, default=0.0, help="0 = sized to host RAM and a time budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    ap.add_argument("--no-prefetch", action="store_true", help="e2e: upload the features in line (serial with the step) "
                    "instead of one step ahead on the copy stream")
    ap.add_argument("--no-parity", action="store_true", help="skip the N>1 vs N=1 parity figure")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from mrgcn_b200 import _native as nv
    from mrgcn_b200 import partition as part
    from mrgcn_b200.graph import RelGraph
    from mrgcn_b200.data.batch import FullBatch
    from mrgcn_b200.layers.graph import fused_projection_pitch, padded_features
    from mrgcn_b200.models.mrgcn import MRGCN
    from mrgcn_b200.models.rgcn import RGCN
    from mrgcn_b200.tasks.link_prediction import compute_ranks_fast, score_distmult_bc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus

    shp, N, tr = make_workload(args.shape, args.scale)
    R, B = shp.num_relations, shp.num_bases if shp.num_bases > 0 else -1
    dims = shp.dims
    featureless = dims[0] == 0
    is_lp = shp.task == "lp"
    modules = modules_of(shp)
    full = RelGraph.from_triples(tr, N, shp.num_props, device=dev)
    nnz = full.E
    fused = bool(not featureless and fused_projection_pitch(dims[0], B, dims[1]))
    Xh = None
    if not featureless:
        Xh = torch.empty((N, dims[0]), dtype=torch.float32).pin_memory()
        Xh.normal_(generator=torch.Generator().manual_seed(1))      # identical features on every rank
    ce = nn.CrossEntropyLoss(reduction="sum")
    bce = nn.BCEWithLogitsLoss(reduction="sum")
    if is_lp:
        trip_np, Yh = lp_batch(tr)
        n_lab = len(trip_np)
    else:
        lab_idx, lab_y = labelled_nodes(N, dims[-1])
        n_lab = len(lab_idx)

    # the model, built through the drop-in class: under torchrun it is node-partitioned (equal node ranges)
    torch.manual_seed(1)
    model = MRGCN(modules, [], R, N, num_bases=B, p_dropout=0.0, featureless=featureless, bias=False, link_prediction=is_lp)
    model.to(dev)
    rg = model.rgcn
    params = list(model.parameters())
    row, col, val = full.coo
    A_coo = torch.sparse_coo_tensor(torch.stack([row, col]), val, (N, R * N))      # what the reference's batch carries
    batch = FullBatch(full if world == 1 else A_coo, [Xh if Xh is not None else torch.empty((N, 0))], np.arange(N))

    parity = None
    if world == 1:
        graph = full
        Xd = padded_features(Xh.to(dev)) if Xh is not None else None      # rows of ceil32(in) floats (tensor-map loads)
        if is_lp:
            trip_d, Y_d = torch.from_numpy(trip_np).to(dev), Yh.to(dev)
        else:
            idx_d, y_d = torch.from_numpy(lab_idx).to(dev), torch.from_numpy(lab_y).to(dev)

        def loss_of(out):
            if is_lp:
                return bce(score_distmult_bc((trip_d[:, 0], trip_d[:, 1], trip_d[:, 2]), out, rg.relations), Y_d) / n_lab
            return ce(out[idx_d], y_d) / n_lab

        def step_loss():
            return loss_of(rg(Xd, graph))
        step_after = None

        def step_device():
            for p in params:
                p.grad = None
            loss = step_loss()
            loss.backward()
            return loss

        def step_e2e():
            for p in params:
                p.grad = None
            if not args.no_prefetch:
                model.prefetch(batch)                   # the NEXT step's upload starts now, on the copy stream
            loss = loss_of(model(batch))                # host features -> device inside the call (or the copy started a step ago)
            loss.backward()
            return float(loss.item())                   # device -> host read of the step's result
        g_meta = dict(E=full.E, ND=full.ND, NS=full.NS, R=R, n_chunks=full.n_chunks)
        plan = full._tab[1] if getattr(full, "_tab", None) else None
    else:
        lay = rg.lay
        lo, hi = lay.lo, lay.hi
        # ---- N > 1 vs N = 1: the same weights on one GPU (rank 0), loss and small-parameter gradients compared
        ref_state = None
        if not args.no_parity:
            torch.manual_seed(1)
            ref = RGCN(modules, R, N, B, 0.0, featureless, False, is_lp)      # identical on every rank (same seed)
            ref_state = {k: v.clone() for k, v in ref.state_dict().items()}
            rg.load_full_state(ref_state)
        rg.to(dev)
        rg.set_graph(row, col, val)
        model._part_A = A_coo
        src_part = rg.layer0_is_source_partitioned()
        Xd = None
        if Xh is not None:
            Xd = padded_features(Xh[lo:hi].to(dev)) if src_part else lay.to_padded(Xh.to(dev))
        if is_lp:
            mine = np.arange(n_lab) % world == rank                         # the rank's shard of the batch's triples
            t_own = torch.from_numpy(trip_np[mine]).to(dev)
            trip_d = torch.stack([lay.pad_ids(t_own[:, 0]), t_own[:, 1], lay.pad_ids(t_own[:, 2])], 1)
            Y_d = Yh[torch.from_numpy(mine)].to(dev)
        else:
            m = (lab_idx >= lo) & (lab_idx < hi)
            idx_d, y_d = torch.from_numpy(lab_idx[m] - lo).to(dev), torch.from_numpy(lab_y[m]).to(dev)
            idx_all, y_all = torch.from_numpy(lab_idx).to(dev), torch.from_numpy(lab_y).to(dev)

        def step_loss():
            rg.hooks_enabled = False
            H = rg(Xd)                                                       # the rank's rows
            if is_lp:      # all-gather E once, score the rank's shard of the triples, reduce-scatter dE in backward
                E_all = part.GatherRows.apply(H, lay)
                return bce(score_distmult_bc((trip_d[:, 0], trip_d[:, 1], trip_d[:, 2]), E_all, rg.relations), Y_d) / n_lab
            return ce(H[idx_d], y_d) / n_lab
        step_after = rg.sync_grads

        def step_device():
            for p in params:
                p.grad = None
            loss = step_loss()
            loss.backward()
            rg.sync_grads()
            return loss

        def step_e2e():
            rg.hooks_enabled = True                       # the drop-in path: gradients of replicated weights all-reduced by hooks
            for p in params:
                p.grad = None
            if not args.no_prefetch:
                model.prefetch(batch)                     # the NEXT step's upload of the rank's rows, on the copy stream
            out = model(batch)                            # every rank uploads the feature rows it owns; logits of all nodes
            if is_lp:
                t = torch.from_numpy(trip_np).to(dev)
                loss = bce(score_distmult_bc((t[:, 0], t[:, 1], t[:, 2]), out, rg.relations), Yh.to(dev)) / n_lab
            else:
                loss = ce(out[idx_all], y_all) / n_lab
            loss.backward()
            return float(loss.item())
        if ref_state is not None:
            loss_p = step_device().detach().clone()
            dist.all_reduce(loss_p)
            small = {n: p.grad.clone() for n, p in rg.named_parameters() if n != "layers.layer_0.weight_I" and p.grad is not None}
            if rank == 0:
                ref.to(dev)
                Xr = padded_features(Xh.to(dev)) if Xh is not None else None
                out = ref(Xr, full)
                if is_lp:
                    t = torch.from_numpy(trip_np).to(dev)
                    lr = bce(score_distmult_bc((t[:, 0], t[:, 1], t[:, 2]), out, ref.relations), Yh.to(dev)) / n_lab
                else:
                    lr = ce(out[torch.from_numpy(lab_idx).to(dev)], torch.from_numpy(lab_y).to(dev)) / n_lab
                lr.backward()
                errs = {"loss": abs(float(loss_p) - float(lr)) / max(abs(float(lr)), 1e-30)}
                for n, p in ref.named_parameters():
                    if n in small:
                        errs[n] = float((small[n] - p.grad).abs().max()) / max(float(p.grad.abs().max()), 1e-30)
                parity = {"max_rel_err": max(errs.values()), "per_tensor": errs, "loss_n1": float(lr), "loss": float(loss_p)}
                del ref, out, Xr
            del ref_state
            torch.cuda.empty_cache()
        g_meta = dict(E=rg.gI.E, ND=rg.gI.ND, NS=rg.gI.NS, R=R, n_chunks=rg.gI.n_chunks)
        plan = rg.gI._tab[1] if getattr(rg.gI, "_tab", None) else None
        del full
        torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        barrier()
        if profile:
            nv.profile_dump()
            nv.profile_enable(True)
        l0 = nv.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        prof = None
        if profile:
            nv.profile_enable(False)
            prof = nv.profile_dump()
        launches = nv.launch_count() - l0
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), launches, prof

    # eager pass: per-kernel CUDA-event profile (kernel table, roofline) and the launch count
    _, _, launches, prof = timed(step_device, args.steps, args.warmup, profile=True)
    launches_per_step = launches / max(args.steps, 1)
    if plan is None:
        g_ = full if world == 1 else rg.gI
        plan = g_._tab[1] if getattr(g_, "_tab", None) else None
    if plan is not None:
        g_meta.update(n_tasks=plan["n_tasks"], n_pieces=plan["n_pieces"], n_blks=plan["n_blks"])

    # collectives of one step (N > 1): sizes from the partition's log, each timed on its own
    collectives = None
    if world > 1:
        part.COMM_LOG = []
        step_device()
        torch.cuda.synchronize()
        log, part.COMM_LOG = part.COMM_LOG, None
        collectives = []
        for name in ("reduce_scatter", "all_gather", "all_reduce"):
            sizes = [b for n_, b in log if n_ == name]
            if not sizes:
                continue
            nbytes = max(sizes)
            buf = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
            piece = torch.empty(max(nbytes // 4 // world, 1), dtype=torch.float32, device=dev)

            def run():
                if name == "reduce_scatter":
                    dist.reduce_scatter_tensor(piece, buf[:piece.numel() * world])
                elif name == "all_gather":
                    dist.all_gather_into_tensor(buf[:piece.numel() * world], piece)
                else:
                    dist.all_reduce(buf)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                run()
            b.record()
            torch.cuda.synchronize()
            collectives.append({"name": name, "bytes": int(nbytes), "calls_per_step": len(sizes), "ms_per_call": a.elapsed_time(b) / 10})

    # timed pass: the same step captured once into a CUDA graph and replayed (the step is static in full-batch
    # training: same graph, same shapes, same buffers every epoch), which removes the host launch overhead that
    # dominates once the partitioned step drops to a few ms.  Falls back to eager launches if capture fails.
    run_step, graphed = step_device, False
    if not args.no_graph:
        try:
            from mrgcn_b200.stepping import GraphedStep      # the product API: capture once, replay per epoch
            run_step, graphed = GraphedStep(step_loss, params, after_backward=step_after, barrier=barrier), True
        except Exception as exc:     # pragma: no cover
            if rank == 0:
                print("bench: CUDA graph capture failed (%s); timing eager launches" % str(exc).splitlines()[0], file=sys.stderr)
            torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        ms_total, _, _, _ = timed(run_step, args.steps, args.warmup)
    clocks = clk.summary()
    ms_step = ms_total / args.steps
    value = nnz / (ms_step / 1e3)
    launches = int(round(launches_per_step * args.steps))

    # end to end: host features copied every step, loss read back every step
    e_steps = max(3, min(args.steps, 10))
    if not args.no_prefetch and Xh is not None:
        model.prefetch(batch)          # prime the pipeline: from here on every step starts the copy the NEXT step consumes
    ms_e2e, wall_e2e, _, _ = timed(step_e2e, e_steps, 2)
    ms_e2e_eager = max(ms_e2e, wall_e2e) / e_steps
    ms_e2e_step, e2e_mode = ms_e2e_eager, "eager MRGCN.forward(batch) + backward"
    x_rows_static = Xh is not None and (world == 1 or src_part)      # the captured step reads the rank's own rows from Xd
    if graphed and x_rows_static and not args.no_prefetch:
        # the same step as the device-timed one (captured once, GraphedStep), fed from the host every step: the upload
        # started a step ago is waited for and copied into the captured step's input buffer, the graph replayed, the
        # loss read back.  One host call instead of ~100 launches: this is what lets e2e follow the step time at N > 1.
        cols = Xh.shape[1]

        def step_e2e_graphed():
            model.prefetch(batch)
            Xd[:, :cols].copy_(model.upload(batch), non_blocking=True)
            return float(run_step().item())
        barrier()
        ms_g, wall_g, _, _ = timed(step_e2e_graphed, e_steps, 2)
        ms_e2e_step, e2e_mode = max(ms_g, wall_g) / e_steps, "GraphedStep replay fed by MRGCN.prefetch/upload"
    h2d = int(Xh.numel() * 4) if Xh is not None else 0      # whole job: the N ranks together copy the matrix once
    if is_lp and e2e_mode.startswith("eager"):
        h2d += int(trip_np.nbytes + Yh.numel() * 4)      # the graphed step keeps the (fixed) triples and labels on the device

    # ranking throughput (LP shapes, 1 GPU): compute_ranks_fast on one test batch, raw and filtered
    lp_extra = None
    if is_lp and world == 1:
        with torch.no_grad():
            emb = rg(Xd, graph).detach()
            n_rank = min(len(tr), 20000)                 # one test split's worth of facts (FB15k-237 test: 20 466)
            facts = torch.from_numpy(tr[:n_rank].astype(np.int64)).to(dev)
            lp_extra = {}
            for flt in (False, True):
                compute_ranks_fast(facts, emb, rg.relations, 50, flt)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(3):
                    compute_ranks_fast(facts, emb, rg.relations, 50, flt)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / 3
                lp_extra["filtered" if flt else "raw"] = {"facts": n_rank, "candidates": N, "ms": dt * 1e3,
                                                           "rank_scores_per_s": 2 * n_rank * N / dt,
                                                           "tflops": 2 * n_rank * N * 3 * dims[-1] / dt / 1e12}

    if rank == 0:
        peak, peak_src = peaks()
        roof = None
        kernels = {}
        if prof:
            # per-kernel table; the dominant kernel (largest share of the step) is the one reported
            alg = algorithmic_bytes(g_meta, dims[0], dims[1:], max(B, 0), fused)
            tot = sum(v[1] for v in prof.values())
            for name, (n, ms) in prof.items():
                per_step = n // args.steps if args.steps else n
                ab = alg.get(name)
                gb = (sum(ab) / 1e9) if ab and len(ab) == per_step else None
                kernels[name] = {"launches_per_step": per_step, "ms_per_step": ms / args.steps, "share": ms / tot if tot else None,
                                 "alg_gb_per_step": gb, "gbps": (gb / (ms / args.steps / 1e3)) if gb else None}
            top = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
            k = kernels[top]
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tpath) and args.shape == "am" and args.scale == 1.0 and world == 1:
                with open(tpath) as f:
                    t = json.load(f).get(top)
                if t and k["launches_per_step"]:
                    traffic = t["bytes_per_step"] / k["launches_per_step"]      # per launch, like `achieved`
            if k["gbps"]:
                step_gb = sum(v["alg_gb_per_step"] or 0 for v in kernels.values())
                roof = {"kernel": top, "bound": "hbm", "achieved": k["gbps"], "peak": peak, "unit": "GB/s",
                        "frac": k["gbps"] / peak, "traffic": traffic, "peak_source": peak_src,
                        "alg_bytes_per_launch": k["alg_gb_per_step"] * 1e9 / max(k["launches_per_step"], 1),
                        "launches_per_step": k["launches_per_step"], "ms_per_step": k["ms_per_step"], "share_of_step": k["share"],
                        "step_alg_gb": step_gb, "step_gbps": step_gb / (ms_step / 1e3), "step_frac": step_gb / (ms_step / 1e3) / peak}
                if args.shape in SURVEY_B_PER_EDGE and world == 1 and args.scale == 1.0:
                    sgb = SURVEY_B_PER_EDGE[args.shape] * nnz / 1e9      # the survey's per-edge-gather formulation of the same step
                    roof.update(survey_step_gb=sgb, survey_step_frac=sgb / (ms_step / 1e3) / peak)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cscale, free = (args.cpu_sample_scale, host_ram_gb()) if args.cpu_sample_scale else pick_cpu_scale(args.shape, 25.0, 2, threads)
            cnnz, times, cn, kind = cpu_reference_step(args.shape, cscale * args.scale, 2, 1, threads)
            cpu = {"value": cnnz / float(np.mean(times)), "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": "%s-shape scaled x%g (N=%d, nnz=%d; R, bases, dims unchanged), 1 warm-up + 2 timed steps, host RAM free %.0f GB" % (
                       args.shape, cscale * args.scale, cn, cnnz, free)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args.shape) + (" (scaled x%g)" % args.scale if args.scale != 1.0 else ""),
                           "nnz": nnz, "parallelism": "1 GPU" if world == 1 else "1-D node partition x%d (equal node ranges)" % world,
                           "launch": "CUDA graph replay of one captured step" if graphed else "eager launches",
                           "feature_term": "per-basis projection on tcgen05 + table mixing" if fused else "per-edge messages",
                           "l2": "working set (identity table, features, edge lists: GBs) exceeds the 126 MB L2; no flush needed"
                                 if nnz > 5e6 else "working set fits the 126 MB L2 (small graph): numbers are L2-resident"},
                "e2e": {"value": nnz / (ms_e2e_step / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e_step, "steps": e_steps, "mode": e2e_mode, "ms_per_step_eager": ms_e2e_eager,
                        "upload": "in line on the compute stream" if args.no_prefetch or Xh is None else
                                  "MRGCN.prefetch: step i+1's copy (one per step, all inside the timed region) overlaps step i"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels}
        if parity is not None:
            line["parity_vs_n1"] = parity["max_rel_err"]
            line["parity_detail"] = parity
        if collectives is not None:
            line["collectives"] = collectives
        if lp_extra is not None:
            line["lp"] = lp_extra
        print(json.dumps(line), flush=True)
    if world > 1:
        # A captured CUDA graph holds NCCL kernels of this communicator; tearing the process group down with the graph
        # alive can block.  Everything has been measured and printed: drain, meet the other ranks, leave.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
