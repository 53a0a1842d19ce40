"""Parity of the CUDA path (through the C ABI) against the golden fixtures the
real reference produced and against the oracle on seeded graphs.  Integer /
index outputs must be bit-exact; modularity of integer-weight graphs is
bit-exact too, real-weight graphs within 1e-12 (the north-star bar is 1e-6)."""
import numpy as np
import pytest
from conftest import GOLDEN as ALL_GOLDEN, load_golden

pytestmark = pytest.mark.gpu

# the non-simple input (parallel entries, self-loops, unsorted rows) has its own test at the end of the file
GOLDEN = [n for n in ALL_GOLDEN if n != "multi9"]


def _csr(g):
    return g["offsets"], g["tails"].astype(np.int64), g["weights"]


@pytest.mark.parametrize("name", GOLDEN)
def test_sweep_by_sweep_matches_reference(gpu, name):
    g = load_golden(name)
    o, t, w = _csr(g)
    exact = name != "rmat9_realw"
    gpu.upload(o, t, w)
    assert gpu.info().weight_mode == (1 if exact else 2)
    gpu.phase_begin()
    for k in range(int(g["iters"])):
        gpu.sweep()
        assert np.array_equal(gpu.labels(2), g["targets"][k]), f"targetComm differs in iteration {k + 1}"
        cw = gpu.cluster_weight()
        if exact:
            assert np.array_equal(cw, g["cw"][k])
        else:
            np.testing.assert_allclose(cw, g["cw"][k], rtol=1e-13, atol=0)
        sums = gpu.apply_and_reduce()
        q = gpu.modularity_from_sums(sums)
        if exact:
            assert q == g["qs"][k], (k, q, g["qs"][k])
        else:
            assert abs(q - g["qs"][k]) < 1e-12
        size, deg = gpu.cinfo()
        assert np.array_equal(size, g["csize"][k])
        if exact:
            assert np.array_equal(deg, g["cdeg"][k])
        else:
            np.testing.assert_allclose(deg, g["cdeg"][k], rtol=1e-12, atol=1e-12)
        gpu.rotate_labels()


@pytest.mark.parametrize("name", GOLDEN)
def test_phase_matches_reference(gpu, name):
    from cuvite_b200.engine import distLouvainMethod
    g = load_golden(name)
    q, cvect, iters = distLouvainMethod(*_csr(g), -1.0, 1e-6, gpu)
    assert iters == int(g["iters"])
    assert np.array_equal(cvect, g["cvect"])
    if name != "rmat9_realw":
        assert q == float(g["q"])
    else:
        assert abs(q - float(g["q"])) < 1e-12


def test_aos_upload_equals_soa(gpu):
    g = load_golden("rmat10_intw")
    o, t, w = _csr(g)
    edges = np.empty(t.size, dtype=[("tail", "<i8"), ("weight", "<f8")])
    edges["tail"], edges["weight"] = t, w
    gpu.upload_aos(o, edges)
    q, iters = gpu.phase_run()
    assert q == float(g["q"]) and iters == int(g["iters"]) and np.array_equal(gpu.labels(0), g["cvect"])


def test_plugin_call_one_rank(gpu, oracle):
    """gpu_for_louvain_iteration + gpu_reduce_mod driven exactly like louvain.cpp:638-739."""
    from cuvite_b200.engine import gpu_for_louvain_iteration, gpu_reduce_mod
    g = load_golden("hub30k")
    o, t, w = _csr(g)
    nv = o.size - 1
    edges = np.empty(t.size, dtype=[("tail", "<i8"), ("weight", "<f8")])
    edges["tail"], edges["weight"] = t, w
    vdeg = oracle.vertex_degree(o, w)
    csize, cdeg = np.ones(nv, np.int64), vdeg.copy()
    const = 1.0 / vdeg.sum()
    past, curr, target = np.arange(nv, dtype=np.int64), np.arange(nv, dtype=np.int64), np.empty(nv, np.int64)
    prev, iters = -1.0, 0
    while True:
        iters += 1
        cw = np.zeros(nv)
        usz, udg = np.zeros(nv, np.int64), np.zeros(nv)  # distCleanCWandCU
        assert gpu_for_louvain_iteration(nv, o, edges, curr, target, vdeg, csize, cdeg, usz, udg, {}, {}, {},
                                         const, cw, 0, iters, gpu) == 1
        assert np.array_equal(target, g["targets"][iters - 1])
        assert np.array_equal(cw, g["cw"][iters - 1])
        csize += usz
        cdeg += udg  # distUpdateLocalCinfo
        assert np.array_equal(csize, g["csize"][iters - 1]) and np.array_equal(cdeg, g["cdeg"][iters - 1])
        le = gpu_reduce_mod(cdeg, cw, gpu, nv)
        q = le[0] * const - le[1] * const * const
        assert q == g["qs"][iters - 1]
        if q - prev < 1e-6:
            break
        prev = q
        past, curr, target = curr, target, past
    assert iters == int(g["iters"]) and np.array_equal(past, g["cvect"])


def test_plugin_call_partitioned_with_ghosts(gpu, oracle):
    """Three vertex-range partitions, each through the plugin call with its ghost
    labels / remote community info as maps (fillRemoteCommunities' outputs)."""
    from cuvite_b200.engine import gpu_for_louvain_iteration
    g = load_golden("rmat10_intw")
    o, t, w = _csr(g)
    nv = o.size - 1
    vdeg = oracle.vertex_degree(o, w)
    # state after one reference iteration, so that labels/Cinfo are non-trivial
    curr = g["targets"][0].astype(np.int64)
    csize, cdeg = g["csize"][0].astype(np.int64), g["cdeg"][0]
    const = 1.0 / vdeg.sum()
    want_t, want_cw, want_dd, want_ds = oracle.sweep(nv, 0, nv, o, t, w, curr, vdeg, cdeg, csize, const)
    parts = [(nv * p) // 3 for p in range(4)]
    got_t, got_cw = np.empty(nv, np.int64), np.empty(nv)
    dd, ds = np.zeros(nv), np.zeros(nv, np.int64)
    for p in range(3):
        b0, b1 = parts[p], parts[p + 1]
        oo = o[b0:b1 + 1] - o[b0]
        tt, ww = t[o[b0]:o[b1]], w[o[b0]:o[b1]]
        edges = np.empty(tt.size, dtype=[("tail", "<i8"), ("weight", "<f8")])
        edges["tail"], edges["weight"] = tt, ww
        ghosts = np.unique(tt[(tt < b0) | (tt >= b1)])
        remoteComm = {int(v): int(curr[v]) for v in ghosts}
        comms = np.unique(np.concatenate([curr[b0:b1], curr[ghosts]]))
        rc = comms[(comms < b0) | (comms >= b1)]
        remoteCinfo = {int(c): (int(csize[c]), float(cdeg[c])) for c in rc}
        remoteCupdate = {int(c): [0, 0.0] for c in rc}
        tg, cw = np.empty(b1 - b0, np.int64), np.zeros(b1 - b0)
        usz, udg = np.zeros(b1 - b0, np.int64), np.zeros(b1 - b0)
        assert gpu_for_louvain_iteration(b1 - b0, oo, edges, curr[b0:b1].copy(), tg, vdeg[b0:b1], csize[b0:b1].copy(),
                                         cdeg[b0:b1].copy(), usz, udg, remoteComm, remoteCinfo, remoteCupdate,
                                         const, cw, p, 1, gpu, base=b0, nv_global=nv) == 1
        got_t[b0:b1], got_cw[b0:b1] = tg, cw
        ds[b0:b1] += usz
        dd[b0:b1] += udg
        for c, (s, d) in remoteCupdate.items():
            ds[c] += s
            dd[c] += d
    assert np.array_equal(got_t, want_t) and np.array_equal(got_cw, want_cw)
    assert np.array_equal(ds, want_ds) and np.array_equal(dd, want_dd)


@pytest.mark.parametrize("scale", [16, 18])
def test_rmat_phase_matches_oracle(gpu, oracle, scale):
    from cuvite_b200 import graphgen as gg
    g = gg.rmat(scale, seed=2)
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy()
    want = oracle.phase(o, t, w, trace=3)
    gpu.upload(o, t, w)
    gpu.phase_begin()
    for k in range(min(3, want.iters)):
        gpu.sweep()
        assert np.array_equal(gpu.labels(2), want.targets[k])
        sums = gpu.apply_and_reduce()
        assert gpu.modularity_from_sums(sums) == want.qs[k]
        gpu.rotate_labels()
    q, iters = gpu.phase_run()
    assert (q, iters) == (want.q, want.iters)
    assert np.array_equal(gpu.labels(0), want.cvect)


def test_full_size_properties_rmat20(gpu):
    """BASELINE config C2 size: invariants that need no oracle (conservation of
    community size/degree, zero-sum deltas, targets drawn from the neighbourhood)."""
    import torch
    from cuvite_b200 import graphgen as gg
    g = gg.rmat(20, seed=2, device="cuda")
    gpu.upload_device(g.offsets, g.tails, g.weights)
    info = gpu.info()
    assert info.nv_local == 1 << 20 and info.weight_mode == 1 and info.local_weight == g.ne
    assert sum(info.class_count[:info.num_classes]) == info.nv_local
    assert sum(info.class_entries[:info.num_classes]) == g.ne
    gpu.phase_begin()
    labels = torch.arange(g.nv, device="cuda")
    src = torch.repeat_interleave(torch.arange(g.nv, device="cuda"), g.offsets[1:] - g.offsets[:-1])
    for _ in range(3):
        gpu.sweep()
        gpu.sync()
        dsize = torch.as_tensor(gpu.buffer("dsize"), device="cuda")
        ddeg = torch.as_tensor(gpu.buffer("ddeg"), device="cuda")
        assert int(dsize.sum()) == 0 and float(ddeg.sum()) == 0.0
        target = torch.as_tensor(gpu.buffer("labels_target"), device="cuda").long()
        # every target is the own community or the community of a neighbour
        ok = torch.zeros(g.nv, dtype=torch.bool, device="cuda")
        ok |= target == labels
        hit = labels[g.tails] == target[src]
        ok.index_put_((src[hit],), torch.ones((), dtype=torch.bool, device="cuda"))
        assert bool(ok.all())
        sums = gpu.apply_and_reduce()
        csize = torch.as_tensor(gpu.buffer("csize"), device="cuda")
        cdeg = torch.as_tensor(gpu.buffer("cdeg"), device="cuda")
        assert int(csize.sum()) == g.nv and float(cdeg.sum()) == float(g.ne)
        assert int((csize < 0).sum()) == 0
        # community info recomputed from the labels equals the incrementally updated one
        deg = (g.offsets[1:] - g.offsets[:-1]).double()
        assert torch.equal(torch.zeros(g.nv, dtype=torch.float64, device="cuda").index_add_(0, target, deg), cdeg)
        assert 0.0 <= sums[0] <= g.ne
        gpu.rotate_labels()
        labels = target


@pytest.mark.parametrize("name", GOLDEN)
def test_coarsen_matches_reference_rebuild(gpu, name):
    """distbuildNextLevelGraph (rebuild.cpp): same renumbering, same CSR, same weights."""
    g = load_golden(name)
    gpu.upload(*_csr(g))
    gpu.phase_run()
    nnv = gpu.coarsen()
    o, t, w = gpu.download_graph()
    assert nnv == g["co_offsets"].size - 1
    assert np.array_equal(o, g["co_offsets"]) and np.array_equal(t, g["co_tails"])
    if name != "rmat9_realw":
        assert np.array_equal(w, g["co_weights"])
    else:
        np.testing.assert_allclose(w, g["co_weights"], rtol=1e-13)


@pytest.mark.parametrize("name", GOLDEN)
def test_full_run_matches_reference_driver(gpu, name):
    """main.cpp's phase loop: per-level modularity, cumulative iterations, level sizes, final assignment."""
    g = load_golden(name)
    gpu.upload(*_csr(g))
    lq, li, ln, comm = gpu.run()
    assert li == list(g["level_iters"]) and ln == list(g["level_nv"])
    if name != "rmat9_realw":
        assert lq == list(g["level_q"])
    else:
        np.testing.assert_allclose(lq, g["level_q"], rtol=0, atol=1e-12)
    assert np.array_equal(comm, g["comm"])


def test_full_run_rmat18_matches_oracle(gpu, oracle):
    from cuvite_b200 import graphgen as gg
    g = gg.rmat(18, seed=2)
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy()
    want = oracle.run(o, t, w)
    gpu.upload(o, t, w)
    lq, li, ln, comm = gpu.run()
    assert li == want.level_iters and ln == want.level_nv
    np.testing.assert_allclose(lq, want.level_q, rtol=0, atol=1e-12)  # sum of squares may exceed 2^53
    assert np.array_equal(comm, want.comm)
    # NB: the reference's reported level Q is NOT the modularity of `comm`: it mixes
    # iteration k's intra-community weight with iteration k+1's degrees and returns the
    # labels of the iteration before (louvain.cpp:521-534, 576) - mirrored, not fixed.
    assert len(np.unique(comm)) == ln[-1] if len(ln) == 1 else len(np.unique(comm)) <= ln[-1]


def test_reference_main_runs_on_our_plugin(tmp_path):
    """The reference's UNMODIFIED main.cpp/louvain.cpp/rebuild.cpp, built with -DRUNONGPU=1
    against include/GpuGraph.cuh + adapter.cpp + libcuvite_b200.so (oracle/_ref/gc_gpu_dropin),
    must print the same per-level lines and write the same .communities file as the
    reference's CPU build (oracle/_ref/gc_cpu64) on the same Vite file — through the
    per-iteration plugin (gc_gpu_dropin) and through the resident outer boundary
    (gc_gpu_resident: adapter_resident.cpp supplies distLouvainMethod(.., GpuGraph&))."""
    import os
    import re
    import shutil
    import subprocess
    from cuvite_b200 import graphgen as gg, vite_io
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    cpu, dropin = os.path.join(ref, "gc_cpu64"), os.path.join(ref, "gc_gpu_dropin")
    resident = os.path.join(ref, "gc_gpu_resident")  # outer boundary: distLouvainMethod(.., GpuGraph&) is ours
    if not (os.path.exists(cpu) and os.path.exists(dropin) and os.path.exists(resident)):
        pytest.skip("oracle/_ref binaries not built (needs /root/reference at build time)")
    g = gg.rmat(14, seed=2)
    outs = {}
    for tag, exe in (("cpu", cpu), ("gpu", dropin), ("resident", resident)):
        d = tmp_path / tag
        d.mkdir()
        f = str(d / "g.bin")
        vite_io.write_vite(f, g)
        r = subprocess.run([exe, "-f", f, "-o"], capture_output=True, text=True, timeout=600, cwd=str(d),
                           env=dict(os.environ, OMP_NUM_THREADS="4"))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        levels = re.findall(r"Level (\d+), Modularity: ([0-9.eE+-]+), Clustering time: [^,]*, Iterations: (\d+)", r.stdout)
        assert levels, r.stdout[-2000:]
        outs[tag] = (levels, open(f + ".communities").read())
    assert outs["cpu"][0] == outs["gpu"][0] == outs["resident"][0]
    assert outs["cpu"][1] == outs["gpu"][1] == outs["resident"][1]


def test_tma_staged_variant_matches_reference(monkeypatch):
    """CVB_TMA=1: block-per-vertex classes read their tails/weights tile from a cp.async.bulk +
    mbarrier staged shared-memory buffer; same results as the golden fixture."""
    from cuvite_b200.engine import GpuGraph
    monkeypatch.setenv("CVB_TMA", "1")
    gpu = GpuGraph()
    g = load_golden("hub30k")  # has vertices in the 513..8192 block classes
    gpu.upload(*_csr(g))
    gpu.phase_begin()
    for k in range(int(g["iters"])):
        gpu.sweep()
        assert np.array_equal(gpu.labels(2), g["targets"][k])
        gpu.apply_and_reduce()
        gpu.rotate_labels()
    gpu.close()


@pytest.mark.parametrize("big,mp,lm_mask", [("1", "0", None), ("1", "0", "0x1ffc"), ("0", "1", None)])
def test_big_vertex_paths_match_oracle(oracle, big, mp, lm_mask, monkeypatch):
    """Round-2 paths for vertices above 8192 entries on the hub graph below (one vertex above 65536 entries):
    the default partition-then-hash path (k_big_stage / k_big_pass / k_big_final), the same with every class
    on the second-generation kernels (CVB_LM_MASK=0x1ffc), and the serial multi-pass kernel (CVB_BIG=0
    CVB_MP=1) with the round-1 global tables above 65536."""
    monkeypatch.setenv("CVB_BIG", big)
    monkeypatch.setenv("CVB_MP", mp)
    if lm_mask:
        monkeypatch.setenv("CVB_LM_MASK", lm_mask)
    _giant_hub_case(oracle)


@pytest.mark.parametrize("cluster,arena_mb,twopass,dense", [(False, 0, True, True), (True, 0, True, True),
                                                            (False, 1, True, False), (False, 0, False, False),
                                                            (False, 0, False, True)])
def test_true_giant_hub_matches_oracle(oracle, cluster, arena_mb, twopass, dense, monkeypatch):
    """ROUND-1 A/B variants (CVB_BIG=0: the partition-then-hash path off).
    Vertices above 8192 entries: the two-pass shared-memory kernel for 8193..16384 (default;
    twopass=False sends that class through the global tables too), the global-table path above
    (arena_mb=1 forces one batch per big vertex through the shared table arena) and the
    thread-block-cluster / distributed-shared-memory path (CVB_CLUSTER=1), incl. one vertex
    above 65536 entries.  dense=True (default): the vertices above 16384 entries accumulate into
    dense per-vertex arrays with L2 reductions; dense=False: open-addressing tables in global
    memory."""
    monkeypatch.setenv("CVB_BIG", "0")
    monkeypatch.setenv("CVB_CLUSTER", "1" if cluster else "0")
    monkeypatch.setenv("CVB_NO_TWOPASS", "0" if twopass else "1")
    monkeypatch.setenv("CVB_BIG_DENSE", "1" if dense else "0")
    if arena_mb:
        monkeypatch.setenv("CVB_GIANT_ARENA_MB", str(arena_mb))
    _giant_hub_case(oracle)


def _giant_hub_case(oracle):
    import torch
    from cuvite_b200 import graphgen as gg
    from cuvite_b200.engine import GpuGraph
    gpu = GpuGraph()
    nv = 90000
    bg = gg.erdos_renyi(nv, 6.0, seed=11)
    src = np.repeat(np.arange(nv, dtype=np.int64), np.diff(bg.offsets.numpy()))
    dst = bg.tails.numpy()
    es, ed = [src], [dst]
    for hub, deg in ((0, nv - 1), (1, 40000), (2, 20000), (3, 10000), (4, 16000), (5, 8300)):
        es.append(np.full(deg, hub, np.int64))
        ed.append((np.arange(deg, dtype=np.int64) * 7919 + 11 * hub + 13) % nv)
    g = gg.csr_from_pairs(nv, torch.from_numpy(np.concatenate(es)), torch.from_numpy(np.concatenate(ed)))
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy()
    want = oracle.phase(o, t, w, trace=8)
    gpu.upload(o, t, w)
    info = gpu.info()
    from cuvite_b200.engine import CLASS_NAMES
    counts = dict(zip(CLASS_NAMES, list(info.class_count)[:info.num_classes]))
    assert counts["giant"] >= 1 and counts["c64k"] >= 1 and counts["c32k"] >= 1 and counts["c16k"] >= 1, counts
    gpu.phase_begin()
    for k in range(want.iters):
        gpu.sweep()
        assert np.array_equal(gpu.labels(2), want.targets[k]), f"iteration {k + 1}"
        sums = gpu.apply_and_reduce()
        assert gpu.modularity_from_sums(sums) == want.qs[k]
        gpu.rotate_labels()
    q, iters = gpu.phase_run()
    assert (q, iters) == (want.q, want.iters) and np.array_equal(gpu.labels(0), want.cvect)
    gpu.close()


@pytest.mark.parametrize("name", ["rmat10_intw", "hub30k", "er1k", "ring6x5", "empty_rows"])
def test_staged_coarsening_driver_matches_golden(name):
    """DistLouvain.run on one rank drives cvb_renumber / cvb_contract_local /
    cvb_graph_from_pairs (the N-rank coarsening stages, here with the merge of stage 3
    exercised on already-distinct pairs) and must reproduce the reference's levels."""
    from cuvite_b200.dist import DistLouvain
    from cuvite_b200.engine import GpuGraph
    g = load_golden(name)
    gpu = GpuGraph()
    gpu.upload(*_csr(g))
    nv = g["offsets"].size - 1
    lq, li, ln, comm = DistLouvain(gpu, 0, 1, nv).run()
    assert lq == list(g["level_q"]) and li == list(g["level_iters"]) and ln == list(g["level_nv"])
    assert np.array_equal(comm.cpu().numpy(), g["comm"])
    gpu.close()


def test_threshold_cycling_matches_reference_cli(tmp_path):
    """cvb_louvain_run_ex(CVB_RUN_THRESHOLD_CYCLING) == the reference CLI with -i (main.cpp:225-239) on
    a graph where -i changes the trajectory: Q as printed (6 significant digits) and cumulative
    iterations per level."""
    import os
    import re
    import subprocess
    from cuvite_b200 import graphgen as gg, vite_io
    from cuvite_b200.engine import GpuGraph
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "gc_cpu64")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gc_cpu64 not built (needs /root/reference at build time)")
    g = gg.erdos_renyi(3000, 8.0, seed=1)
    f = str(tmp_path / "h.bin")
    vite_io.write_vite(f, g)
    r = subprocess.run([exe, "-f", f, "-i"], capture_output=True, text=True, timeout=600, cwd=str(tmp_path),
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want = re.findall(r"Level \d+, Modularity: ([0-9.eE+-]+), Clustering time: [^,]*, Iterations: (\d+)", r.stdout)
    assert want
    gpu = GpuGraph()
    gpu.upload(g.offsets.numpy(), g.tails.numpy(), g.weights.numpy())
    lq, li, _, _ = gpu.run(threshold_cycling=True)
    gpu.close()
    assert [f"{q:.6g}" for q in lq] == [f"{float(q):.6g}" for q, _ in want]
    assert li == [int(i) for _, i in want]


def test_non_simple_input_matches_reference():
    """A Vite input that is NOT a simple graph (parallel entries in both directions, self-loops stored once
    or twice, rows in shuffled order; fixture produced by the real reference): the table-free first-sweep
    kernels must step aside (k_check_simple_rows) and every sweep, the phase, the coarsened graph and the
    full run must still equal the reference bit for bit."""
    from cuvite_b200.engine import GpuGraph
    g = load_golden("multi9")
    o, t, w = _csr(g)
    gpu = GpuGraph()
    gpu.upload(o, t, w)
    assert gpu.info().weight_mode == 1
    gpu.phase_begin()
    for k in range(int(g["iters"])):
        gpu.sweep()
        assert np.array_equal(gpu.labels(2), g["targets"][k]), f"targetComm differs in iteration {k + 1}"
        assert np.array_equal(gpu.cluster_weight(), g["cw"][k])
        q = gpu.modularity_from_sums(gpu.apply_and_reduce())
        assert q == g["qs"][k], (k, q, g["qs"][k])
        size, deg = gpu.cinfo()
        assert np.array_equal(size, g["csize"][k]) and np.array_equal(deg, g["cdeg"][k])
        gpu.rotate_labels()
    q, iters = gpu.phase_run()
    assert (q, iters) == (float(g["q"]), int(g["iters"])) and np.array_equal(gpu.labels(0), g["cvect"])
    nnv = gpu.coarsen()
    co, ct, cw = gpu.download_graph()
    assert nnv == g["co_offsets"].size - 1 and np.array_equal(co, g["co_offsets"])
    assert np.array_equal(ct, g["co_tails"]) and np.array_equal(cw, g["co_weights"])
    gpu.upload(o, t, w)
    lq, li, ln, comm = gpu.run()
    assert lq == list(g["level_q"]) and li == list(g["level_iters"]) and ln == list(g["level_nv"])
    assert np.array_equal(comm, g["comm"])
    gpu.close()


def test_config_c1_er16k_full_run_matches_oracle(oracle):
    """BASELINE config 1 (16K-vertex Erdos-Renyi, p = 16/n): full multi-phase run on the GPU equals the
    oracle level by level and vertex by vertex (the oracle equals the reference CLI on this graph:
    tests/test_oracle_pin.py::test_config_c1_er16k_reference_cli_vs_oracle)."""
    from cuvite_b200 import graphgen as gg
    from cuvite_b200.engine import GpuGraph
    g = gg.erdos_renyi(16384, 16.0, seed=1)
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy()
    want = oracle.run(o, t, w)
    gpu = GpuGraph()
    gpu.upload(o, t, w)
    lq, li, ln, comm = gpu.run()
    gpu.close()
    assert li == want.level_iters and ln == want.level_nv
    np.testing.assert_allclose(lq, want.level_q, rtol=0, atol=1e-12)
    assert np.array_equal(comm, want.comm)


def test_native_vite_load_matches_upload(tmp_path):
    """cvb_graph_load_vite (file -> pinned staging -> device, reference: loadDistGraphMPIIO + first plugin upload):
    same phase result as uploading the arrays; a partitioned load reports the reference's ranges."""
    from cuvite_b200 import vite_io
    from cuvite_b200.engine import GpuGraph
    from cuvite_b200.graphgen import CSRGraph
    import torch
    g = load_golden("rmat10_intw")
    o, t, w = _csr(g)
    f = str(tmp_path / "g.bin")
    vite_io.write_vite(f, CSRGraph(o.size - 1, torch.from_numpy(o), torch.from_numpy(t), torch.from_numpy(w)))
    gpu = GpuGraph()
    part = gpu.load_vite(f)
    assert (part.nv_global, part.base, part.nv_local, part.ne_local) == (o.size - 1, 0, o.size - 1, t.size)
    q, iters = gpu.phase_run()
    assert (q, iters) == (float(g["q"]), int(g["iters"])) and np.array_equal(gpu.labels(0), g["cvect"])
    part = gpu.load_vite(f, rank=1, nparts=3, balanced=True)
    parts = vite_io.balanced_ranges(o, 3)
    info = gpu.info()
    assert (part.base, part.nv_local) == (parts[1], parts[2] - parts[1])
    assert (info.base, info.nv_local, info.nv_global, info.ne_local) == (parts[1], parts[2] - parts[1], o.size - 1,
                                                                        int(o[parts[2]] - o[parts[1]]))
    gpu.close()


@pytest.mark.xfail(strict=False, reason="CVB_SPLIT=1 / CVB_L2_PERSIST=1 are A/B variants written after the last GPU "
                                        "run of round 1 (no GPU minutes left): first executed by this test")
@pytest.mark.parametrize("name", ["rmat10_intw", "hub30k"])
def test_split_gather_variant_matches_reference(name, monkeypatch):
    """CVB_SPLIT=1: the hash classes gather 4-byte labels per entry and community degrees per distinct
    community; CVB_L2_PERSIST=1: access-policy window on the packed array.  Same results as the fixture."""
    from cuvite_b200.engine import GpuGraph
    monkeypatch.setenv("CVB_SPLIT", "1")
    monkeypatch.setenv("CVB_L2_PERSIST", "1")
    g = load_golden(name)
    gpu = GpuGraph()
    gpu.upload(*_csr(g))
    gpu.phase_begin()
    for k in range(int(g["iters"])):
        gpu.sweep()
        assert np.array_equal(gpu.labels(2), g["targets"][k]), f"targetComm differs in iteration {k + 1}"
        assert np.array_equal(gpu.cluster_weight(), g["cw"][k])
        assert gpu.modularity_from_sums(gpu.apply_and_reduce()) == g["qs"][k]
        gpu.rotate_labels()
    gpu.close()


# ---- BASELINE.json configs at their stated sizes -------------------------------
def test_config_c2_rmat20_sweeps_and_phase_match_oracle(oracle):
    """C2: R-MAT scale-20 on one GPU.  The first two sweeps (targets, cluster weights, modularity) and the
    whole phase 0 against the CPU checker on the same graph, plus the committed golden phase result."""
    import json
    import os
    import torch
    from cuvite_b200 import graphgen as gg
    from cuvite_b200.engine import GpuGraph
    g = gg.rmat(20, seed=2, device="cuda")
    o, t, w = g.offsets.cpu().numpy(), g.tails.cpu().numpy(), g.weights.cpu().numpy()
    want = oracle.phase(o, t, w, trace=2)
    gpu = GpuGraph()
    try:
        gpu.upload_device(g.offsets, g.tails, g.weights)
        del g
        torch.cuda.empty_cache()
        gpu.phase_begin()
        for k in range(2):
            gpu.sweep()
            assert np.array_equal(gpu.labels(2), want.targets[k]), f"targetComm differs in sweep {k + 1}"
            assert np.array_equal(gpu.cluster_weight(), want.cw[k])
            q = gpu.modularity_from_sums(gpu.apply_and_reduce())
            assert q == want.qs[k], (k, q, want.qs[k])
            gpu.rotate_labels()
        q, iters = gpu.phase_run()
        assert (q, iters) == (want.q, want.iters)
        assert np.array_equal(gpu.labels(0), want.cvect)
        gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bench_q.json")))["rmat20"]
        assert (q, iters) == (gold["q"], gold["iters"])
    finally:
        gpu.close()


def test_config_c3_rmat22_full_run_matches_golden_summary():
    """C3: R-MAT scale-22, full multi-phase Louvain with coarsening to convergence on one GPU, against the
    summary the REFERENCE produced for the same graph (tools/run_configs.py c3 -> tests/golden/c3_rmat22.json:
    per-level Q / cumulative iterations / sizes and a hash of the final assignment)."""
    import hashlib
    import json
    import os
    import torch
    from cuvite_b200 import graphgen as gg
    from cuvite_b200.engine import GpuGraph
    path = os.path.join(os.path.dirname(__file__), "golden", "c3_rmat22.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c3_rmat22.json not generated yet (tools/run_configs.py c3)")
    gold = json.load(open(path))
    g = gg.rmat(22, seed=2, device="cuda")
    gpu = GpuGraph()
    try:
        gpu.upload_device(g.offsets, g.tails, g.weights)
        del g
        torch.cuda.empty_cache()
        lq, li, ln, comm = gpu.run()
        assert lq == gold["level_q"], (lq, gold["level_q"])
        assert li == gold["level_iters"] and ln == gold["level_nv"]
        assert hashlib.sha256(np.ascontiguousarray(comm, dtype=np.int64).tobytes()).hexdigest() == gold["comm_sha256"]
    finally:
        gpu.close()


def test_real_weight_hub_graph_f64_mode_within_tolerance(oracle):
    """F64 mode above degree 512 accumulates with atomics (order not fixed): on a real-weight graph with
    hubs in every degree class the modularity must stay within the north-star tolerance (1e-6; observed
    ~1e-15) of the CPU checker, per sweep and for the phase, and the iteration count must agree."""
    from graphs import hub_graph, with_real_weights
    from cuvite_b200.engine import GpuGraph
    o, t, w = with_real_weights(*hub_graph())
    want = oracle.phase(o, t, w, trace=16)
    gpu = GpuGraph()
    try:
        gpu.upload(o, t, w)
        assert gpu.info().weight_mode == 2
        gpu.phase_begin()
        same_labels = True
        for k in range(min(want.iters, 16)):
            gpu.sweep()
            same_labels &= bool(np.array_equal(gpu.labels(2), want.targets[k]))
            q = gpu.modularity_from_sums(gpu.apply_and_reduce())
            assert abs(q - want.qs[k]) <= 1e-6, (k, q, want.qs[k])
            gpu.rotate_labels()
        q, iters = gpu.phase_run()
        assert iters == want.iters and abs(q - want.q) <= 1e-6
        # the labels are expected to agree too (sums differ in the last bits only); a tie flipped by
        # rounding would show up here without breaking the modularity bound
        assert same_labels, "labels differ from the CPU checker (modularity still within 1e-6)"
    finally:
        gpu.close()


@pytest.mark.parametrize("graph", ["er3000", "rmat13"])
@pytest.mark.parametrize("et_type,delta", [(1, 1.0), (3, 1.0), (2, 0.9), (4, 0.9), (2, 0.25), (4, 1.0)])
def test_early_termination_matches_reference_cli(tmp_path, graph, et_type, delta):
    """cvb_louvain_run_et == the reference CLI with -t <type> [-a <delta>] (louvain.cpp:7-423, main.cpp:326-341)
    on graphs where early termination changes the trajectory: Q as printed (6 significant digits) and
    cumulative iterations per level.  The freeze rules are mirrored bug for bug (k_et_rotate)."""
    import os
    import re
    import subprocess
    from cuvite_b200 import graphgen as gg, vite_io
    from cuvite_b200.engine import GpuGraph
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "gc_cpu64")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gc_cpu64 not built (needs /root/reference at build time)")
    g = gg.erdos_renyi(3000, 8.0, seed=1) if graph == "er3000" else gg.rmat(13, seed=2)
    f = str(tmp_path / "h.bin")
    vite_io.write_vite(f, g)
    opts = ["-t", str(et_type)] + (["-a", str(delta)] if et_type in (2, 4) else [])
    # ONE host thread: the reference keeps vActive in a std::vector<bool> and clears bits of the same word from
    # several OpenMP threads (louvain.cpp:174-186, 375-398) — with more than one thread freezes get lost and its
    # own output changes from run to run (observed: three different level-5 modularities in three runs)
    r = subprocess.run([exe, "-f", f] + opts, capture_output=True, text=True, timeout=600, cwd=str(tmp_path),
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want = re.findall(r"Level \d+, Modularity: ([0-9.eE+-]+), Clustering time: [^,]*, Iterations: (\d+)", r.stdout)
    assert want
    gpu = GpuGraph()
    try:
        gpu.upload(g.offsets.numpy(), g.tails.numpy(), g.weights.numpy())
        lq, li, _, _ = gpu.run(et_type=et_type, et_delta=delta)
    finally:
        gpu.close()
    assert [f"{q:.6g}" for q in lq] == [f"{float(q):.6g}" for q, _ in want]
    assert li == [int(i) for _, i in want]


@pytest.mark.parametrize("k", [1, 2, 3])
def test_phase_run_max_iters_returns_matching_labels(gpu, k):
    """on the max_iters exit the returned modularity (iteration k's) must belong to the labels in
    LABELS_PAST — the input of iteration k, i.e. iteration k-1's targets."""
    g = load_golden("rmat10_intw")
    assert int(g["iters"]) > k
    gpu.upload(*_csr(g))
    q, iters = gpu.phase_run(max_iters=k)
    assert iters == k and q == g["qs"][k - 1]
    want = g["targets"][k - 2] if k >= 2 else np.arange(g["offsets"].size - 1)
    assert np.array_equal(gpu.labels(0), want)
