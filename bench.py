#!/usr/bin/env python
"""bench.py — R-GCN fwd+bwd edges/s on the AM-shape workload (BASELINE.json metric), 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--shape am]

A step = one full-batch training step of the configured model WITHOUT the optimizer: forward of the
2-layer R-GCN (am.toml: 151 -> 10 -> 11, 40 bases, identity + feature terms in layer 0), cross-entropy on
the labelled nodes, backward to every parameter.  edges = nnz of the stacked adjacency (forward + inverse +
self-loop blocks).  Prints ONE JSON line (rank 0).

  value    : device-resident throughput (features already in HBM), CUDA events, max over ranks; the K timed steps are
             replays of ONE captured CUDA graph of the step (static in full-batch training); --no-graph times eager launches
  e2e      : the same step through the public module call `MRGCN.forward(batch)` with the feature matrix in
             pinned HOST memory (copied to the device every step, as the reference's forward does,
             mrgcn/models/mrgcn.py:203-204) and the loss read back to the host
  roofline : dominant kernel of the step, timed live with CUDA events inside the library (mrgcn_profile_enable) in an
             eager pass of the same K steps, against MEASURED_PEAKS.json; `traffic` = its DRAM bytes per launch from the
             ncu capture recorded in profiles/ncu_traffic.json
  cpu_baseline / --impl reference : the CPU oracle (the reference's own torch.sparse op sequence,
             oracle/reference_port.py) on a bounded sample of the same workload, on this box's host cores

N > 1 (torchrun): the graph is 1-D node-partitioned (mrgcn_b200/partition.py); total work is fixed
("scaling": "strong").
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


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def algorithmic_bytes(g, in0, dims, B, P=1):
    """Algorithmic bytes per launch of every kernel of one step: 4 B per structure word read,
    gathered operand rows counted per edge, every output written once.  Returns {kernel: [bytes per launch...]}
    in launch order (layer 0 fwd, layer 1 fwd, layer 1 bwd, layer 0 bwd)."""
    E, ND, NS, R = g["E"], g["ND"], g["NS"], g["R"]
    nch = g["n_chunks"]
    h, c = dims
    out = {}
    ms = lambda d: 4 if d <= 4 else 8 if d <= 8 else (d + 15) // 16 * 16      # padded message row (mrgcn_msg_stride)

    def add(k, v):
        out.setdefault(k, []).append(float(v))
    # ---- layer 0 forward (identity + feature)
    add("ident_msg_fwd", B * NS * h * 4 + E * 12 + NS * 4 + E * ms(h) * 4 + R * B * 4)
    add("basis_mix_fwd", B * in0 * h * 4 + R * B * 4 + R * in0 * h * 4)
    add("feat_msg_fwd", E * (8 + in0 * 4) + R * in0 * h * 4 + E * ms(h) * 4)
    add("agg_fwd", ND * 4 + E * 2 * (4 + ms(h) * 4) + ND * h * 4)
    # ---- layer 1 forward (feature only)
    add("basis_mix_fwd", B * h * c * 4 + R * B * 4 + R * h * c * 4)
    add("feat_msg_fwd", E * (8 + h * 4) + R * h * c * 4 + E * ms(c) * 4)
    add("agg_fwd", ND * 4 + E * (4 + ms(c) * 4) + ND * c * 4)
    # ---- layer 1 backward
    add("act_bwd", 2 * ND * c * 4)
    add("feat_bwd_w", E * (12 + h * 4 + c * 4) + nch * h * c * 4)
    add("feat_w_reduce", nch * h * c * 4 + R * h * c * 4)
    add("basis_mix_bwd_v", R * h * c * 4 + B * h * c * 4)
    add("basis_mix_bwd_c", R * h * c * 4 + B * h * c * 4 + R * B * 4)
    add("feat_bwd_x_msg", E * (8 + c * 4) + R * h * c * 4 + E * ms(h) * 4)
    add("feat_bwd_x_agg", NS * 4 + E * (4 + ms(h) * 4) + NS * h * 4)
    # ---- layer 0 backward
    add("act_bwd", 3 * ND * h * 4)
    add("ident_bwd_w", E * (12 + h * 4) + NS * 4 + B * NS * h * 4 + R * B * 4)
    add("ident_bwd_c", B * NS * h * 4 + E * (12 + h * 4) + NS * 4 + E * B * 4)
    add("comp_chunk_reduce", E * (4 + B * 4) + nch * B * 4)
    add("comp_reduce", nch * B * 4 + R * B * 4)
    add("feat_bwd_w", E * (12 + in0 * 4 + h * 4) + nch * in0 * h * 4)
    add("feat_w_reduce", nch * in0 * h * 4 + R * in0 * h * 4)
    add("basis_mix_bwd_v", R * in0 * h * 4 + B * in0 * h * 4)
    add("basis_mix_bwd_c", R * in0 * h * 4 + B * in0 * h * 4 + R * B * 4)
    return out


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_step(shape_name, sample_scale, steps, warmup, threads):
    """The reference's CPU path (oracle port: same scipy/torch-CPU calls, oracle/reference_port.py) on a bounded
    sample of the workload: N and triples scaled by `sample_scale`, R / bases / dims unchanged."""
    from oracle import reference_port as rp
    torch.set_num_threads(threads)
    shp, n, tr = make_workload(shape_name, sample_scale)
    R = shp.num_relations
    A = rp.csr_to_coo(rp.as_float32(rp.stacked_adjacency(tr, n, shp.num_props)), torch.int8 if False else torch.float32)
    nnz = A._nnz()
    dims = shp.dims
    torch.manual_seed(1)
    modules = [(dims[k], dims[k + 1], "mrgcn", "relu" if k + 2 < len(dims) else None) for k in range(len(dims) - 1)]
    layers, _ = rp.init_rgcn_params(modules, R, n, shp.num_bases if shp.num_bases > 0 else -1, dims[0] == 0, False, False)
    for l in layers:
        for v in l.values():
            v.requires_grad_(True)
    X = torch.randn(n, dims[0]) if dims[0] > 0 else None
    idx, y = labelled_nodes(n, dims[-1])
    idx, y = torch.from_numpy(idx), torch.from_numpy(y)
    times = []
    for it in range(warmup + steps):
        for l in layers:
            for v in l.values():
                v.grad = None
        t0 = time.perf_counter()
        out = rp.rgcn_forward(layers, [m[3] for m in modules], X, A, num_nodes=n, num_relations=R,
                              num_bases=shp.num_bases if shp.num_bases > 0 else -1, featureless=dims[0] == 0)
        loss = rp.nc_loss(out, idx, y)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return nnz, times, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    scale = args.cpu_sample_scale
    nnz, times, n = cpu_reference_step(args.shape, scale, max(1, args.steps), max(0, min(args.warmup, 1)), threads)
    ms = 1e3 * float(np.mean(times))
    val = nnz / (ms / 1e3)
    sample = "%s-shape scaled x%g (N=%d, nnz=%d; R, bases, dims unchanged), %d timed step(s)" % (args.shape, scale, n, nnz, len(times))
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": len(times),
            "warmup": max(0, min(args.warmup, 1)), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.shape), "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(shape):
    from mrgcn_b200.synth import SHAPES
    s = SHAPES[shape]
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
    ap.add_argument("--cpu-sample-scale", type=float, default=1.0 / 16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from mrgcn_b200 import _native as nv
    from mrgcn_b200.graph import RelGraph
    from mrgcn_b200.data.batch import FullBatch
    from mrgcn_b200.models.mrgcn import MRGCN
    from mrgcn_b200.partition import PartitionedRGCN, balanced_bounds, node_weights

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
    modules = [(dims[k], dims[k + 1], "mrgcn", nn.ReLU() if k + 2 < len(dims) else None) for k in range(len(dims) - 1)]
    full = RelGraph.from_triples(tr, N, shp.num_props, device=dev)
    nnz = full.E
    lab_idx, lab_y = labelled_nodes(N, dims[-1])
    torch.manual_seed(1 + rank)
    Xh = None
    if not featureless:
        Xh = torch.empty((N, dims[0]), dtype=torch.float32).pin_memory()
        g = torch.Generator().manual_seed(1)       # identical features on every rank
        Xh.normal_(generator=g)
    # device-resident features live in rows of ceil32(in) floats (what the projection kernel's tensor-map loads read);
    # the e2e path uploads into the same layout every step (MRGCN._upload_features)
    from mrgcn_b200.layers.graph import padded_features
    Xd = padded_features(Xh.to(dev)) if Xh is not None else None
    ce = nn.CrossEntropyLoss(reduction="sum")

    if world == 1:
        model = MRGCN(modules, [], R, N, num_bases=B, p_dropout=0.0, featureless=featureless, bias=False)
        model.to(dev)
        graph = full
        idx_d, y_d = torch.from_numpy(lab_idx).to(dev), torch.from_numpy(lab_y).to(dev)
        n_lab = len(lab_idx)
        params = list(model.parameters())

        def step_device():
            for p in params:
                p.grad = None
            out = model.rgcn(Xd, graph)
            loss = ce(out[idx_d], y_d) / n_lab
            loss.backward()
            return loss

        batch = FullBatch(graph, [Xh if Xh is not None else torch.empty((N, 0))], np.arange(N))

        def step_e2e():
            for p in params:
                p.grad = None
            out = model(batch)                         # host features -> device inside the call
            loss = ce(out[idx_d], y_d) / n_lab
            loss.backward()
            return float(loss.item())                  # device -> host read of the step's result
        g_meta = dict(E=full.E, ND=full.ND, NS=full.NS, R=R, n_chunks=full.n_chunks)
    else:
        row, col, val = full.coo
        bounds = balanced_bounds(node_weights(row, col, N), world)
        model = PartitionedRGCN(modules, R, N, B, featureless, False, False, bounds, rank)
        model.to(dev)
        model.set_graph(row, col, val)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        m = (lab_idx >= lo) & (lab_idx < hi)
        idx_d, y_d = torch.from_numpy(lab_idx[m] - lo).to(dev), torch.from_numpy(lab_y[m]).to(dev)
        n_lab = len(lab_idx)
        params = list(model.parameters())
        del full, row, col, val
        torch.cuda.empty_cache()

        def step_device(X=None):
            for p in params:
                p.grad = None
            out = model(Xd if X is None else X)
            loss = ce(out[idx_d], y_d) / n_lab
            loss.backward()
            model.sync_grads()
            return loss

        from mrgcn_b200.partition import gather_rows

        def step_e2e():
            # every rank copies only the rows it owns from pinned host memory; NVLink all-gather rebuilds the matrix
            Xfull = gather_rows(Xh[lo:hi].to(dev, non_blocking=True), model.lay) if Xh is not None else None
            loss = step_device(Xfull)
            dist.all_reduce(loss)
            return float(loss.item())
        g_meta = dict(E=model.gF.E, ND=model.gF.ND, NS=model.gF.NS, R=R, n_chunks=model.gF.n_chunks)

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
    # timed pass: the same step captured once into a CUDA graph and replayed (the step is static in full-batch
    # training: same graph, same shapes, same buffers every epoch), which removes the host launch overhead that
    # dominates once the partitioned step drops to a few ms.  Falls back to eager launches if capture fails.
    run_step, graphed = step_device, False
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step_device()
            torch.cuda.current_stream().wait_stream(side)
            barrier()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                step_device()
            run_step, graphed = cg.replay, True
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
    ms_e2e, wall_e2e, _, _ = timed(step_e2e, e_steps, 2)
    ms_e2e_step = max(ms_e2e, wall_e2e) / e_steps
    h2d = int(Xh.numel() * 4) if Xh is not None else 0      # whole job: the N ranks together copy the matrix once

    if rank == 0:
        peak, peak_src = peaks()
        roof = None
        kernels = {}
        if prof:
            # per-kernel table; the dominant kernel (largest share of the step) is the one reported
            alg = algorithmic_bytes(g_meta, dims[0], dims[1:], max(B, 0)) if (len(dims) == 3 and not featureless and B > 0) else {}
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
                roof = {"kernel": top, "bound": "hbm", "achieved": k["gbps"], "peak": peak, "unit": "GB/s",
                        "frac": k["gbps"] / peak, "traffic": traffic, "peak_source": peak_src,
                        "alg_bytes_per_launch": k["alg_gb_per_step"] * 1e9 / max(k["launches_per_step"], 1),
                        "launches_per_step": k["launches_per_step"], "ms_per_step": k["ms_per_step"], "share_of_step": k["share"],
                        "step_alg_gb": sum(v["alg_gb_per_step"] or 0 for v in kernels.values()),
                        "step_gbps": sum(v["alg_gb_per_step"] or 0 for v in kernels.values()) / (ms_step / 1e3)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cnnz, times, cn = cpu_reference_step(args.shape, args.cpu_sample_scale * args.scale, 2, 1, threads)
            cpu = {"value": cnnz / float(np.mean(times)), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "%s-shape scaled x%g (N=%d, nnz=%d; R, bases, dims unchanged), 1 warm-up + 2 timed steps" % (
                       args.shape, args.cpu_sample_scale * args.scale, cn, cnnz)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args.shape) + (" (scaled x%g)" % args.scale if args.scale != 1.0 else ""),
                           "nnz": nnz, "parallelism": "1 GPU" if world == 1 else "1-D node partition x%d" % world,
                           "launch": "CUDA graph replay of one captured step" if graphed else "eager launches",
                           "l2": "working set (weight_I 2.67 GB, X 1.0 GB, edge lists) exceeds the 126 MB L2; no flush needed"},
                "e2e": {"value": nnz / (ms_e2e_step / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e_step, "steps": e_steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels}
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
