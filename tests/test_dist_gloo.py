"""World-size > 1 on CPU (gloo): the distributed phase orchestration of
cuvite_b200.dist (partitioning, delta all-reduce, label all-gather, modularity
all-reduce, convergence) gives the one-rank reference result for any rank count."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, exchange, ret, balanced=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuvite_b200.dist import DistLouvain
    from cuvite_b200.vite_io import partition_ranges
    from fake_engine import OracleEngine
    from oracle.binding import Oracle
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
    o, t, w = g["offsets"], g["tails"].astype(np.int64), g["weights"]
    nv = o.size - 1
    from cuvite_b200.vite_io import balanced_ranges
    parts = balanced_ranges(o, world) if balanced else partition_ranges(nv, world)
    b0, b1 = parts[rank], parts[rank + 1]
    orc = Oracle()
    orc.set_num_threads(1)
    eng = OracleEngine(orc).upload(o[b0:b1 + 1] - o[b0], t[o[b0]:o[b1]], w[o[b0]:o[b1]], base=b0, nv_global=nv)
    dl = DistLouvain(eng, rank, world, nv, exchange=exchange, parts=parts if balanced else None)
    q, iters = dl.run_phase()
    past = eng.tensor("labels_past").numpy()
    ok = (iters == int(g["iters"]) and abs(q - float(g["q"])) < 1e-12 and np.array_equal(past, g["cvect"][b0:b1]))
    ret[rank] = (bool(ok), q, iters)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name,exchange", [(2, "rmat10_intw", "labels"), (3, "er1k", "labels"),
                                                 (2, "hub30k", "labels"), (3, "rmat10_intw", "deltas")])
def test_distributed_phase_equals_one_rank_reference(oracle, world, name, exchange):
    port = 29600 + world * 7 + len(name) + (50 if exchange == "deltas" else 0)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, name, exchange, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r][0], f"rank {r}: {ret[r]}"


def test_distributed_phase_on_edge_balanced_ranges(oracle):
    """The reference's -b ranges (uneven vertex counts, hub30k puts the hub rows on rank 0)."""
    world, name = 3, "hub30k"
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29877, name, "labels", ret, True), nprocs=world, join=True)
    for r in range(world):
        assert ret[r][0], f"rank {r}: {ret[r]}"


def _run_worker(rank, world, port, name, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuvite_b200.dist import DistLouvain
    from cuvite_b200.vite_io import partition_ranges
    from fake_engine import OracleEngine
    from oracle.binding import Oracle
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
    o, t, w = g["offsets"], g["tails"].astype(np.int64), g["weights"]
    nv = o.size - 1
    parts = partition_ranges(nv, world)
    b0, b1 = parts[rank], parts[rank + 1]
    orc = Oracle()
    orc.set_num_threads(1)
    eng = OracleEngine(orc).upload(o[b0:b1 + 1] - o[b0], t[o[b0]:o[b1]], w[o[b0]:o[b1]], base=b0, nv_global=nv)
    lq, li, ln, comm = DistLouvain(eng, rank, world, nv).run()
    ok = (lq == list(g["level_q"]) and li == list(g["level_iters"]) and ln == list(g["level_nv"])
          and np.array_equal(comm.numpy(), g["comm"]))
    ret[rank] = (bool(ok), lq, li, ln)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "rmat10_intw"), (3, "er1k"), (2, "ring6x5"), (3, "hub30k"), (2, "multi9")])
def test_distributed_multilevel_run_equals_one_rank_reference(oracle, world, name):
    """Phases + distributed coarsening (all-gathered cvect, 1-rank numbering, all-to-all of the
    contracted pairs, owner-side merge) reproduce the one-rank reference's levels and final
    assignment on any rank count."""
    port = 29800 + world * 11 + len(name)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run_worker, args=(world, port, name, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r][0], f"rank {r}: {ret[r]}"


def _communities_worker(rank, world, port, path, ret, cycling=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuvite_b200.dist import DistLouvain
    from cuvite_b200 import vite_io
    from fake_engine import OracleEngine
    from oracle.binding import Oracle
    loc, base, nv = vite_io.read_vite(path, rank, world)  # every rank reads its own rows of the Vite file
    orc = Oracle()
    orc.set_num_threads(1)
    eng = OracleEngine(orc).upload(loc.offsets.numpy(), loc.tails.numpy(), loc.weights.numpy(), base=base, nv_global=nv)
    lq, li, _, comm = DistLouvain(eng, rank, world, nv).run(reference_numbering=True, threshold_cycling=cycling)
    if rank == 0:
        vite_io.write_communities(path + ".ours", comm)
    ret[rank] = (lq, li)
    dist.barrier()
    dist.destroy_process_group()


def test_communities_file_equals_reference_cli(oracle, tmp_path):
    """Vite file in -> <file>.communities out: the N-rank driver with reference_numbering writes
    byte for byte what the reference's own CLI (oracle/_ref/gc_cpu64 -f file -o) writes."""
    import subprocess
    from cuvite_b200 import graphgen as gg, vite_io
    exe = os.path.join(ROOT, "oracle", "_ref", "gc_cpu64")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gc_cpu64 not built (needs /root/reference at build time)")
    f = str(tmp_path / "g.bin")
    vite_io.write_vite(f, gg.rmat(10, seed=2))
    r = subprocess.run([exe, "-f", f, "-o"], capture_output=True, text=True, timeout=600, cwd=str(tmp_path),
                       env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_communities_worker, args=(2, 29933, f, ret), nprocs=2, join=True)
    assert len(ret) == 2
    assert open(f + ".ours").read() == open(f + ".communities").read()


def test_threshold_cycling_equals_reference_cli(oracle, tmp_path):
    """-i (threshold cycling, main.cpp:225-239): level lines (Q as printed, cumulative iterations) and
    the community dump of the reference CLI, on a graph where -i changes the trajectory."""
    import re
    import subprocess
    from cuvite_b200 import graphgen as gg, vite_io
    exe = os.path.join(ROOT, "oracle", "_ref", "gc_cpu64")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gc_cpu64 not built (needs /root/reference at build time)")
    f = str(tmp_path / "h.bin")
    vite_io.write_vite(f, gg.erdos_renyi(3000, 8.0, seed=1))
    outs = {}
    for flag in ((), ("-i",)):
        r = subprocess.run([exe, "-f", f, "-o", *flag], capture_output=True, text=True, timeout=600, cwd=str(tmp_path),
                           env=dict(os.environ, OMP_NUM_THREADS="2"))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[flag] = (re.findall(r"Level \d+, Modularity: ([0-9.eE+-]+), Clustering time: [^,]*, Iterations: (\d+)", r.stdout),
                      open(f + ".communities").read())
    assert outs[()][0] != outs[("-i",)][0], "pick a graph where -i matters"
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_communities_worker, args=(2, 29944, f, ret, True), nprocs=2, join=True)
    lq, li = ret[0]
    want = outs[("-i",)]
    assert [f"{q:.6g}" for q in lq] == [f"{float(q):.6g}" for q, _ in want[0]]
    assert li == [int(i) for _, i in want[0]]
    assert open(f + ".ours").read() == want[1]


def test_partition_ranges_match_reference_formula():
    from cuvite_b200.vite_io import partition_ranges
    for nv, p in [(1024, 3), (7, 8), (1 << 20, 8), (30000, 7)]:
        parts = partition_ranges(nv, p)
        assert parts[0] == 0 and parts[-1] == nv
        assert all(parts[i] == (nv * i) // p for i in range(p + 1))  # distgraph.cpp:118-119
