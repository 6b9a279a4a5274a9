"""Host-side logic that needs no GPU: work lists of the relation-major kernels, hub segments, synthetic workloads,
the algorithmic-bytes model and the JSON contract of bench.py's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunk_worklist_tiles_groups_without_crossing_relations():
    from mrgcn_b200.graph import chunk_worklist
    R = 5
    rng = np.random.default_rng(0)
    cnt = rng.integers(0, 700, size=3 * R)            # 3 slabs x 5 relations, some groups empty
    cnt[[2, 7]] = 0
    grpptr = np.concatenate([[0], np.cumsum(cnt)])
    rel, ptr, rcp, rci = chunk_worklist(grpptr, R, 128)
    n = len(rel)
    assert ptr[0] == 0 and ptr[-1] == grpptr[-1] and np.all(np.diff(ptr) > 0) and np.all(np.diff(ptr) <= 128)
    grp_of_edge = np.repeat(np.arange(3 * R), cnt)
    for c in range(n):                                 # one group, hence one relation, per chunk
        g = grp_of_edge[ptr[c]:ptr[c + 1]]
        assert np.all(g == g[0]) and g[0] % R == rel[c]
    assert np.array_equal(np.sort(rci), np.arange(n))
    for r in range(R):
        mine = rci[rcp[r]:rcp[r + 1]]
        assert np.all(rel[mine] == r) and np.all(np.diff(mine) > 0)     # slab order = ascending chunk id
    assert rcp[-1] == n
    # degenerate: no edges at all
    rel0, ptr0, rcp0, rci0 = chunk_worklist(np.zeros(R + 1, dtype=np.int64), R, 128)
    assert len(rel0) == 0 and list(ptr0) == [0] and rcp0[-1] == 0


def test_hub_segments():
    from mrgcn_b200.graph import hub_segments
    hub, first = hub_segments([513, 512, 1, 2000], 512)
    assert list(first) == [0, 2, 3, 4, 8] and list(hub) == [0, 0, 1, 2, 3, 3, 3, 3]
    hub, first = hub_segments([], 512)
    assert len(hub) == 0 and list(first) == [0]


def test_synthetic_shapes_follow_the_named_configs():
    from mrgcn_b200.synth import SHAPES, synth_graph
    am = SHAPES["am"]
    assert am.num_relations == 267 and am.dims == (151, 10, 11) and am.num_bases == 40 and am.nnz == 13466764
    n, tr = synth_graph("aifb", seed=0)
    assert n == 8285 and tr.shape[1] == 3 and tr.dtype == np.int32
    assert len(np.unique(tr, axis=0)) == len(tr)                       # RDF graphs are sets
    assert set(np.unique(tr[:, 1])) == set(range(SHAPES["aifb"].num_props))   # every property occurs
    n2, tr2 = synth_graph("aifb", seed=0)
    assert np.array_equal(tr, tr2)                                      # seeded


def test_algorithmic_bytes_model():
    sys.path.insert(0, ROOT)
    import bench
    g = dict(E=13466744, ND=1666764, NS=1666764, R=267, n_chunks=17000, n_tasks=1750000, n_pieces=3000000, n_blks=24000)
    # round-1 formulation (per-edge feature messages in layer 0) and round-2 (per-basis projection + table mixing)
    old = bench.algorithmic_bytes(g, 151, (10, 11), 40, False)
    new = bench.algorithmic_bytes(g, 151, (10, 11), 40, True)
    assert len(old["feat_msg_fwd"]) == 2 and old["feat_msg_fwd"][0] > 8e9   # the layer-0 launch gathers E*in*4 bytes
    assert len(new["feat_msg_fwd"]) == 1 and "feat_proj" in new
    # projection: X read once + P written once; the mixing pass reads both tables once
    assert abs(new["feat_proj"][0] - (1666764 * 160 * 4 + 1666764 * 400 * 4 + 2 * 400 * 160 * 4)) < 1e6
    # identity backward in one pass: table read once + gradient written once + per edge 16 B of structure and one gact row
    assert abs(new["ident_bwd_fused"][0] - (2 * 40 * 1666764 * 10 * 4 + 13466744 * (16 + 40) + 1666764 * 4 + 267 * 40 * 4)) < 1e6
    # the dictionary lists alternatives of the same work (table kernels, the separate round-1 kernels): a step runs one of them
    alt = ("tab_bwd_w", "tab_bwd_c", "ident_bwd_w", "ident_bwd_c", "comp_block_reduce")
    tot_old, tot_new = (sum(sum(v) for k, v in d.items() if k not in alt) for d in (old, new))
    assert 20e9 < tot_new < tot_old < 45e9


def test_bench_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-scale", "0.002"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "edges/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_bench_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_tab_plan_invariants_and_reduction_semantics():
    """Work plan of the table-term kernels (mrgcn_b200/graph.py: build_tab_plan, include/mrgcn_b200.h: mrgcn_tab_plan):
    tasks tile E2 exactly, tiles tile the tasks, every (tile, relation) piece holds <= 32 edges of one relation, and the
    two-level reduction the comp-gradient kernel performs with it equals the direct per-relation sum."""
    import torch
    from mrgcn_b200.graph import build_tab_plan
    torch.manual_seed(0)
    NS, R, thresh = 900, 11, 128
    deg = torch.randint(0, 14, (NS,))
    deg[5], deg[77], deg[300], deg[301] = 700, 129, 0, 128
    colptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(deg, 0)])
    E = int(colptr[-1])
    rel = torch.cat([torch.sort(torch.randint(0, R, (int(d),))).values for d in deg])
    p = build_tab_plan(colptr.int(), rel.int(), E, NS, R, thresh)
    lt = p["lt"]
    src, lo = p["task_src"].long(), p["task_lo"].long()
    ln = torch.minimum(torch.full_like(lo, lt), colptr[src + 1] - lo)
    assert int(ln.min()) >= 1 and int(ln.sum()) == E
    assert torch.equal(lo[1:], (lo + ln)[:-1]) and int(lo[0]) == 0                 # tasks are consecutive E2 ranges
    assert bool(((lo >= colptr[src]) & (lo + ln <= colptr[src + 1])).all())        # ... inside their source
    # basis-gradient tasks: every non-hub source once, in node order
    assert torch.equal(p["wsrc"].long(), torch.nonzero(deg <= thresh).flatten())
    # tiles
    ttp, te0 = p["tile_task_ptr"].long(), p["tile_e0"].long()
    assert int(ttp[0]) == 0 and int(ttp[-1]) == p["n_tasks"] and bool((ttp[1:] > ttp[:-1]).all())
    assert torch.equal(te0, lo[ttp[:-1]])
    tend = torch.cat([te0[1:], torch.tensor([E])])
    assert int((tend - te0).max()) == p["tile_slots"] <= 224 + lt - 1
    # pieces: tperm restricted to a tile is a permutation of its slots; a piece has one relation and <= 32 edges
    tperm, pp, tpp = p["tperm"].long(), p["piece_ptr"].long(), p["tile_piece_ptr"].long()
    assert int(pp[0]) == 0 and int(pp[-1]) == E and int((pp[1:] - pp[:-1]).max()) <= 32 and int(tpp[-1]) == p["n_pieces"]
    vals = torch.randn(E, 3, dtype=torch.float64)
    rec = torch.zeros(p["n_pieces"], 3, dtype=torch.float64)
    piece_rel = torch.empty(p["n_pieces"], dtype=torch.long)
    for t in range(p["n_tiles"]):
        q0, q1 = int(pp[tpp[t]]), int(pp[tpp[t + 1]])
        assert (q0, q1) == (int(te0[t]), int(tend[t]))
        assert torch.equal(torch.sort(tperm[q0:q1]).values, torch.arange(q1 - q0))
        for pc in range(int(tpp[t]), int(tpp[t + 1])):
            e = te0[t] + tperm[pp[pc]:pp[pc + 1]]
            assert len(torch.unique(rel[e])) == 1
            piece_rel[pc] = rel[e[0]]
            rec[pc] = vals[e].sum(0)
    rpp, rpi = p["rel_piece_ptr"].long(), p["rel_piece_idx"].long()
    assert torch.equal(torch.sort(rpi).values, torch.arange(p["n_pieces"]))
    for r in range(R):
        mine = rpi[rpp[r]:rpp[r + 1]]
        assert bool((piece_rel[mine] == r).all())
        assert torch.allclose(rec[mine].sum(0), vals[rel == r].sum(0), atol=1e-9)
    # two-stage sum: blocks of <= 128 pieces of one relation tile every relation's piece list
    bp, rbp = p["blk_ptr"].long(), p["rel_blk_ptr"].long()
    assert int(bp[0]) == 0 and int(bp[-1]) == p["n_pieces"] and int((bp[1:] - bp[:-1]).max()) <= 128 and int(rbp[-1]) == p["n_blks"]
    for r in range(R):
        assert int(bp[rbp[r]]) == int(rpp[r]) if rbp[r] < rbp[r + 1] else True
        assert int(bp[rbp[r + 1]]) == int(rpp[r + 1])
    # an empty graph gives an empty plan
    z = build_tab_plan(torch.zeros(5, dtype=torch.int32), torch.zeros(1, dtype=torch.int32), 0, 4, R, thresh)
    assert z["n_tasks"] == 0 and z["n_tiles"] == 0 and z["n_pieces"] == 0 and z["n_wsrc"] == 4


def test_device_worklists_match_the_numpy_ones():
    """chunk_worklist_device / hub_segments_device (torch ops, used at graph build) against the NumPy definitions."""
    import torch
    from mrgcn_b200.graph import chunk_worklist, chunk_worklist_device, hub_segments, hub_segments_device
    rng = np.random.default_rng(3)
    R = 7
    cnt = rng.integers(0, 900, size=3 * R)
    grp = np.concatenate([[0], np.cumsum(cnt)])
    want = chunk_worklist(grp, R, 128)
    got = chunk_worklist_device(torch.from_numpy(grp), R, 128)
    for a, b in zip(want, got):
        assert np.array_equal(np.asarray(a), b.numpy())
    deg = rng.integers(129, 5000, size=17)
    hub, first = hub_segments(deg, 512)
    hub2, first2 = hub_segments_device(torch.from_numpy(deg), 512)
    assert np.array_equal(hub, hub2.numpy()) and np.array_equal(first, first2.numpy())


def test_filter_csr_matches_reference_dictionaries():
    """filter_csr_device == the reference's truedicts lists (link_prediction.py:575-591), as sets, for both sides,
    including repeated facts."""
    import torch
    from mrgcn_b200.tasks.link_prediction import filter_csr_device
    from oracle import reference_port as rp
    rng = np.random.default_rng(0)
    data = np.stack([rng.integers(0, 30, 200), rng.integers(0, 4, 200), rng.integers(0, 30, 200)], 1)
    data[50:60] = data[40:50]
    heads, tails = rp.true_dicts(torch.from_numpy(data))
    facts = torch.from_numpy(data).long()
    for head in (False, True):
        ptr, idx = filter_csr_device(facts, head)
        assert ptr.dtype == torch.int32 and idx.dtype == torch.int32
        for f, (s, p, o) in enumerate(data.tolist()):
            want = sorted(set(heads[(p, o)] if head else tails[(s, p)]))
            assert idx[ptr[f]:ptr[f + 1]].tolist() == want
    ptr, idx = filter_csr_device(facts[:0], True)
    assert ptr.tolist() == [0]
