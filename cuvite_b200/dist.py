"""Multi-GPU phase driver: one process per GPU, vertex-range partition, NCCL.

Replaces the reference's MPI ghost exchange (louvain.cpp: exchangeVertexReqs
3118-3264, fillRemoteCommunities 2588-2959, updateRemoteCommunities 2983-3116,
the Allreduce of distComputeModularity 2466/2522 and of
distCalcConstantForSecondTerm 2175) with collectives over NVLink, keeping the
reference's partition ``parts[p] = nv*p // P`` (distgraph.cpp:118-119).

Per sweep, on every rank (default "labels-only" exchange):
  1. local sweep kernels read the replicated label / community arrays;
  2. ONE ``all_gather`` of the new local labels (4 bytes per vertex) into the
     "next" replicated label array — replaces the ghost-label pull;
  3. every rank rebuilds the replicated community info from those labels and the
     replicated vertex degrees on its own GPU (``cvb_rebuild_and_reduce``) — the
     Cupdate push and the remote-community request/reply are gone, not just
     re-expressed; ``all_reduce(SUM)`` of the two modularity sums;
  4. accept: rotate, which swaps the two replicated label arrays.
(``exchange="deltas"`` keeps the first version: dense ``all_reduce`` of the
per-sweep degree/size deltas + label all-gather after the rotation.)
With integer weights every sum is exact, so an N-GPU phase is bit-identical to
the 1-GPU phase (tests/test_dist_gloo.py checks the orchestration on CPU with a
stand-in local engine; tests/test_gpu_parity.py checks it on GPUs).

The local engine is anything with the small interface ``GpuGraph`` offers
(``phase_begin, sweep, apply_and_reduce, rotate_labels, tensor(name), info()``).
"""
from __future__ import annotations

import json
import os
import statistics
import time

import torch
import torch.distributed as dist

from .vite_io import partition_ranges


def tensor_of(engine, name: str) -> torch.Tensor:
    """torch view of an engine buffer (zero copy)."""
    if hasattr(engine, "tensor"):
        return engine.tensor(name)
    buf = engine.buffer(name)
    if buf.shape[0] == 0:
        return torch.empty(0, device="cuda")
    return torch.as_tensor(buf, device="cuda")


class DistLouvain:
    """Distributed Louvain over ``world`` vertex-range partitions: phases (``run_phase``),
    the coarsening between them (``coarsen``) and the level loop (``run``)."""

    def __init__(self, engine, rank: int, world: int, nv_global: int, group=None, exchange: str = "labels",
                 parts=None):
        """``parts``: explicit vertex ranges of the CURRENT graph (e.g. the reference's -b
        edge-balanced ranges, vite_io.balanced_ranges); default parts[p] = nv*p // P.  Coarsened
        levels always use the default, like rebuild.cpp:294-295."""
        self.e, self.rank, self.world, self.group = engine, rank, world, group
        assert exchange in ("labels", "deltas")
        self.labels_only = exchange == "labels"
        self.constant = None
        self.timing = [] if os.environ.get("CVB_DIST_TIMING") == "1" else None
        self._pinned = None
        self._set_nv_global(nv_global, parts)
        # The real engine runs the phase natively (cvb_dist_*: NCCL issued from C++ on the engine
        # stream); the stand-in engines of the CPU tests keep the torch.distributed path below, which
        # is the same sequence of collectives.
        self.native = bool(self.labels_only and world > 1 and hasattr(engine, "dist_init")
                           and os.environ.get("CVB_DIST_NATIVE", "1") != "0")
        if self.native:
            uid = [engine.dist_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0, group=group)
            engine.dist_init(uid[0], rank, world)
            engine.dist_set_parts(self.parts)
            engine.dist_agree()
        elif world > 1 and hasattr(engine, "stream"):
            # the collectives below run on torch's current stream; the engine must too
            assert engine.stream == torch.cuda.current_stream().cuda_stream, \
                "DistLouvain: create the engine on torch's current stream (GpuGraph(stream=...))"

    def _set_nv_global(self, nv_global: int, parts=None):
        self.nvg = nv_global
        self.parts = list(parts) if parts is not None else partition_ranges(nv_global, self.world)
        assert len(self.parts) == self.world + 1 and self.parts[0] == 0 and self.parts[-1] == nv_global
        self.base, self.bound = self.parts[self.rank], self.parts[self.rank + 1]
        self.even = len({self.parts[p + 1] - self.parts[p] for p in range(self.world)}) == 1

    # -- collectives -------------------------------------------------------------
    def _all_gather_rows(self, global_t: torch.Tensor, local_t: torch.Tensor):
        """global_t[parts[p]:parts[p+1]] <- rank p's local_t, for every p."""
        if self.world == 1:
            if global_t.data_ptr() != local_t.data_ptr():
                global_t[self.base:self.bound].copy_(local_t)
            return
        if self.even:
            dist.all_gather_into_tensor(global_t, local_t.contiguous(), group=self.group)
        else:  # uneven ranges (coarsened levels, -b ranges): ONE all-gather of max-size chunks, then unpack
            m = max(self.parts[p + 1] - self.parts[p] for p in range(self.world))
            pad = torch.empty(m, dtype=local_t.dtype, device=local_t.device)
            pad[:local_t.numel()].copy_(local_t)
            buf = torch.empty(self.world * m, dtype=local_t.dtype, device=local_t.device)
            dist.all_gather_into_tensor(buf, pad, group=self.group)
            for p in range(self.world):
                a, b = self.parts[p], self.parts[p + 1]
                global_t[a:b].copy_(buf[p * m:p * m + (b - a)])

    def _all_reduce(self, t: torch.Tensor, op=dist.ReduceOp.SUM):
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)

    # -- phase -------------------------------------------------------------------
    def phase_begin(self):
        """distInitLouvain across ranks: 2m by all-reduce, community degrees by all-gather."""
        e = self.e
        if self.native:
            e.dist_phase_begin()
            self.constant = None
            return
        info = e.info()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if hasattr(e, "buffer") else torch.device("cpu")
        tot = torch.tensor([info.local_weight], dtype=torch.float64, device=self.dev)
        self._all_reduce(tot)
        self.total_weight = float(tot.item())
        e.phase_begin(self.total_weight)
        self.constant = 1.0 / self.total_weight
        if self.world > 1:
            self._all_gather_rows(tensor_of(e, "cdeg"), tensor_of(e, "vdeg"))
            if self.labels_only:
                tensor_of(e, "vdeg_global").copy_(tensor_of(e, "cdeg"))

    def _mark(self, name):
        """CVB_DIST_TIMING=1: CUDA-event stamps between the stages of a sweep (bench diagnostics)."""
        if self.timing is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timing.append((name, ev))

    def stage_times_ms(self):
        """Mean milliseconds per stage over the sweeps stamped so far (after a synchronize)."""
        if self.native:
            return self.e.dist_stage_ms()
        out, prev = {}, None
        for name, ev in self.timing or []:
            if name != "begin" and prev is not None:
                out.setdefault(name, []).append(prev.elapsed_time(ev))
            prev = ev
        return {k: sum(v) / len(v) for k, v in out.items()}

    def sweep(self):
        """One synchronous sweep + exchange.  Returns the modularity of this iteration."""
        e = self.e
        if self.native:
            return e.dist_sweep()
        self._mark("begin")
        e.sweep()
        self._mark("sweep_kernels")
        if self.world > 1 and self.labels_only:
            self._all_gather_rows(tensor_of(e, "labels_global_next"), tensor_of(e, "labels_target"))
            self._mark("all_gather_labels")
            if hasattr(e, "rebuild_and_reduce_async"):  # sums stay on the device: one read-back per sweep
                e.rebuild_and_reduce_async()
                self._mark("label_diff_and_reduce")
                s = tensor_of(e, "sums")
                self._all_reduce(s)
                self._mark("all_reduce_sums")
                if self._pinned is None:
                    self._pinned = torch.empty(2, dtype=torch.float64).pin_memory()
                self._pinned.copy_(s, non_blocking=True)  # pinned + stream sync: no pageable staging copy
                torch.cuda.current_stream().synchronize()
                sums = self._pinned.numpy()
                self._mark("read_back")
                c = self.constant
                return (sums[0] * c) - (sums[1] * c * c)
            sums = e.rebuild_and_reduce()
        else:
            if self.world > 1:
                self._all_reduce(tensor_of(e, "ddeg"))
                self._all_reduce(tensor_of(e, "dsize"))
            sums = e.apply_and_reduce()
        if self.world > 1:
            s = torch.tensor(sums, dtype=torch.float64, device=self.dev)
            self._all_reduce(s)
            sums = s.cpu().numpy()
        c = self.constant
        return (sums[0] * c) - (sums[1] * c * c)  # louvain.cpp:2472-2473

    def advance(self):
        """Accept the sweep: rotate (louvain.cpp:562-567) and publish the new labels."""
        e = self.e
        e.rotate_labels()  # labels-only mode: also swaps the two replicated label arrays
        if self.world > 1 and not self.labels_only:
            self._all_gather_rows(tensor_of(e, "labels_global"), tensor_of(e, "labels_curr"))

    def run_phase(self, lower: float = -1.0, thresh: float = 1e-6, max_iters: int = 0):
        """distLouvainMethod's loop (louvain.cpp:638-739).  Returns (Q, iterations)."""
        self.phase_begin()
        prev, iters = lower, 0
        while True:
            iters += 1
            cur = self.sweep()
            if (cur - prev) < thresh:
                break
            prev = max(cur, lower)
            self.advance()  # also on the max_iters exit: the returned Q belongs to the past labels
            if max_iters and iters >= max_iters:
                break
        return prev, iters


    # -- coarsening between phases -------------------------------------------------
    def coarsen(self):
        """distbuildNextLevelGraph across ranks (rebuild.cpp:27-454), with the ONE-rank
        community numbering so that an N-rank run reproduces the 1-rank reference:

        1. all-gather cvect (= the local PAST labels) and renumber the communities in order of
           first appearance — every rank computes the same numbering from the same array;
        2. contract the local rows to distinct (new src, new dst, weight) pairs on the device;
        3. one all-to-all moves every pair to the owner of its new source under the
           reference's partition of the new graph, ``parts[p] = nc*p // P`` (rebuild.cpp:294-295);
        4. the owner merges what it received and emits its CSR rows (sorted by tail).

        Returns (newid, nc): newid[i] = new vertex of old GLOBAL vertex i (replicated tensor);
        the all-gathered cvect of the level is left in ``self.last_cvect``."""
        e = self.e
        stages = getattr(self, "coarsen_seconds", None)
        if stages is None:
            stages = self.coarsen_seconds = {}
        clock = [time.perf_counter()]

        def mark(name):  # CVB_DIST_TIMING=1: per-stage host seconds (with a device sync per stage)
            if self.timing is not None:
                torch.cuda.synchronize()
                now = time.perf_counter()
                stages[name] = stages.get(name, 0.0) + now - clock[0]
                clock[0] = now

        cv_local = tensor_of(e, "labels_past")
        cv_global = torch.empty(self.nvg, dtype=cv_local.dtype, device=cv_local.device)
        self._all_gather_rows(cv_global, cv_local)
        self.last_cvect = cv_global
        mark("allgather_cvect")
        newid, nc = e.renumber(cv_global)
        mark("renumber")
        keys, w = e.contract_local(newid, nc)
        mark("contract_local")
        parts = partition_ranges(nc, self.world)
        if self.world > 1:
            cuts = torch.tensor(parts[1:-1], dtype=torch.int64, device=keys.device) << 32
            idx = torch.searchsorted(keys, cuts).tolist()  # keys ascend, source in the high word
            edges = [0] + idx + [keys.numel()]
            send = [edges[p + 1] - edges[p] for p in range(self.world)]
            recv_t = torch.empty(self.world, dtype=torch.int64, device=keys.device)
            dist.all_to_all_single(recv_t, torch.tensor(send, dtype=torch.int64, device=keys.device), group=self.group)
            recv = recv_t.tolist()
            rkeys = torch.empty(sum(recv), dtype=keys.dtype, device=keys.device)
            rw = torch.empty(sum(recv), dtype=w.dtype, device=w.device)
            dist.all_to_all_single(rkeys, keys.contiguous(), recv, send, group=self.group)
            dist.all_to_all_single(rw, w.contiguous(), recv, send, group=self.group)
            keys, w = rkeys, rw
        mark("all_to_all_pairs")
        e.graph_from_pairs(parts[self.rank + 1] - parts[self.rank], parts[self.rank], nc, keys, w)
        mark("graph_from_pairs")
        self._set_nv_global(nc)
        if self.native:
            e.dist_set_parts(self.parts)
            e.dist_agree()
        mark("set_parts_agree")
        return newid, nc

    @staticmethod
    def cycled_threshold(short_phase: int) -> float:
        """Threshold cycling (option -i), main.cpp:225-236: coarse thresholds first, 1e-6 last,
        on a 13-phase cycle."""
        sp = short_phase % 13
        return 1.0e-3 if sp <= 2 else 1.0e-4 if sp <= 6 else 1.0e-5 if sp <= 9 else 1.0e-6

    def run(self, max_levels: int = 200, threshold: float = 1.0e-6, reference_numbering: bool = False,
            threshold_cycling: bool = False):
        """The default phase loop of main.cpp:218-495 across ranks.  Returns (level_q,
        level_iters (cumulative, as the reference prints them), level_nv, comm) where
        comm[i] is the final community of original GLOBAL vertex i (replicated tensor).

        Community ids: by default the coarsening's own numbering (first appearance,
        rebuild.cpp:72-85 — the numbering of the golden fixtures).  ``reference_numbering``
        gives the ids main.cpp writes to ``<input>.communities``: at every level the cvect values
        are replaced by their rank among the sorted distinct values ("ensure clusters are
        contiguous", main.cpp:375-396) before composing (main.cpp:397-404) — including the
        reference's quirk that this composition indexes level l's vertices (numbered by rebuild in
        first-appearance order) with level l-1's sorted-rank ids, so beyond the first level the
        dumped ids describe the reference's output, not necessarily the partition found.

        ``threshold_cycling`` is the reference's -i (main.cpp:225-239, 482-483): the phase threshold —
        both the sweep-loop exit and the level-acceptance test — follows ``cycled_threshold``.  (The
        extra 1e-6 pass main.cpp:436-441 runs right before exiting changes no reported level and no
        output file, so it is not run.)"""
        level_q, level_iters, level_nv = [], [], []
        self.seconds = {"phases": 0.0, "coarsening": 0.0}  # host clock; both stages end with a host sync
        comm, comm_ref, nv0 = None, None, self.nvg
        curr, prev, tot_iters, phase = -1.0, -1.0, 0, 0
        while True:
            nv = self.nvg
            if threshold_cycling:
                threshold = self.cycled_threshold(phase)  # short_phase == phase until the 200-phase cap
            t0 = time.perf_counter()
            curr, iters = self.run_phase(lower=curr, thresh=threshold)
            t1 = time.perf_counter()
            self.seconds["phases"] += t1 - t0
            tot_iters += iters
            if (curr - prev) > threshold:  # main.cpp:361
                newid, _ = self.coarsen()  # main.cpp:413
                self.seconds["coarsening"] += time.perf_counter() - t1
                if reference_numbering:
                    cv = self.last_cvect.long()
                    ranks = torch.searchsorted(torch.unique(cv), cv)  # unique() returns the values sorted
                    # NB main.cpp:402 indexes the level's cvect with the PREVIOUS level's sorted-rank ids
                    # although the level's vertices carry rebuild's first-appearance ids; mirrored as is
                    comm_ref = ranks if comm_ref is None else ranks[comm_ref]
                comm = newid.clone() if comm is None else newid[comm.long()]  # main.cpp:401-404
            else:
                break
            if phase < max_levels:
                level_q.append(curr)
                level_iters.append(tot_iters)
                level_nv.append(nv)
            prev, curr = curr, -1.0
            phase += 1
            if phase == 200 or tot_iters > 10000:  # main.cpp:486
                break
        if comm is None:  # no level improved: main.cpp:196 leaves commAll at its initial -1
            comm = torch.arange(nv0, dtype=torch.int64)
            comm_ref = torch.full((nv0,), -1, dtype=torch.int64)
        return level_q, level_iters, level_nv, (comm_ref if reference_numbering else comm)


def init_process_group(backend: str | None = None):
    """Read torchrun's environment; returns (rank, local_rank, world)."""
    # before the first NCCL communicator of the process exists: see cvb_dist_init (engine.cu) for the measurement
    os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def load_partition(engine, g, rank: int, world: int, parts=None):
    """Upload rows [parts[rank], parts[rank+1]) of an in-memory (device) CSRGraph."""
    parts = partition_ranges(g.nv, world) if parts is None else parts
    b0, b1 = parts[rank], parts[rank + 1]
    e0, e1 = int(g.offsets[b0]), int(g.offsets[b1])
    offs = (g.offsets[b0:b1 + 1] - e0).contiguous()
    tails = g.tails[e0:e1].contiguous()
    w = g.weights[e0:e1].contiguous()
    if offs.is_cuda:
        engine.upload_device(offs, tails, w, base=b0, nv_global=g.nv)
    else:
        engine.upload(offs.numpy(), tails.numpy(), w.numpy(), base=b0, nv_global=g.nv)
    return b0, b1, e1 - e0


def bench_multi(args, metric, unit, base_scale, measured_peak_gbs, ClockSampler, workload_name=None, parity_record=None):
    """bench.py at N > 1: weak scaling, R-MAT scale base+log2(N), hubs spread by a fixed permutation."""
    from . import build
    build.build()
    from .engine import CLASS_NAMES, GpuGraph, set_gpuDevices
    from .graphgen import rmat
    rank, local, world = init_process_group("nccl")
    set_gpuDevices(local)
    torch.cuda.set_device(local)
    scale = (base_scale + (world - 1).bit_length()) if args.scale is None else args.scale
    g = rmat(scale, seed=2, device="cuda", permute_seed=3)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()  # engine kernels and NCCL calls share this (non-default) stream
    torch.cuda.set_stream(side)
    gpu = GpuGraph(stream=side.cuda_stream)
    b0, b1, ne_local = load_partition(gpu, g, rank, world)
    nv, ne = g.nv, g.ne
    # this rank's rows as the reference holds them (Graph::edgeListIndexes + Edge{int64 tail; double weight}
    # AoS, distgraph.hpp) in pinned host memory: the input of the end-to-end leg below
    import numpy as np
    e_lo, e_hi = int(g.offsets[b0]), int(g.offsets[b1])
    from .hostmem import near_gpu
    edges = np.empty(e_hi - e_lo, dtype=[("tail", "<i8"), ("weight", "<f8")])
    edges["tail"] = g.tails[e_lo:e_hi].cpu().numpy()
    edges["weight"] = g.weights[e_lo:e_hi].cpu().numpy()
    with near_gpu(local) as numa_local:  # every rank's pinned rows in the memory of its GPU's socket
        o_pin = torch.from_numpy((g.offsets[b0:b1 + 1] - e_lo).cpu().numpy()).pin_memory().numpy()
        edges_pin = torch.from_numpy(edges.view(np.uint8)).pin_memory()
    edges_p = edges_pin.numpy().view(edges.dtype)
    del g, edges
    torch.cuda.empty_cache()
    dl = DistLouvain(gpu, rank, world, nv)

    def step(state):
        q = dl.sweep()
        state["sweep_ms"].append(gpu.last_sweep_ms())
        state["it"] += 1
        if state["phases"] == 0:
            state["qs"].append(q)
        if q - state["prev"] < 1e-6:
            if state["phases"] == 0:
                state["phase_q"], state["phase_iters"] = state["prev"], state["it"]
            dl.phase_begin()  # converged: restart from singletons inside the timed region
            state["prev"] = -1.0
            state["phases"] += 1
            state["it"] = 0
        else:
            state["prev"] = q
            dl.advance()

    def fresh():
        return {"prev": -1.0, "sweep_ms": [], "phases": 0, "it": 0, "qs": [], "phase_q": None, "phase_iters": None}

    dl.phase_begin()
    st = fresh()
    for _ in range(max(args.warmup, 3)):
        step(st)
    dl.phase_begin()
    st = fresh()
    launches0 = gpu.launch_count
    if dl.timing is not None:
        dl.timing = []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(st)
    ev1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sw = torch.tensor([statistics.mean(st["sweep_ms"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(sw, op=dist.ReduceOp.MAX)
    launches = torch.tensor([gpu.launch_count - launches0], dtype=torch.int64, device="cuda")
    dist.all_reduce(launches)
    ne_max = torch.tensor([ne_local], dtype=torch.int64, device="cuda")
    dist.all_reduce(ne_max, op=dist.ReduceOp.MAX)

    # ---- end to end: the N-rank distLouvainMethod(.., GpuGraph&) of adapter_resident.cpp, per phase: every rank
    # ---- uploads ITS rows from pinned host memory (AoS), the ranks agree on ranges and weight mode, the phase
    # ---- runs through cvb_dist_phase_run, and each rank reads its cvect back.  Host clock, max over ranks.
    e2e = {"value": None, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16,
           "api": "device-resident distributed phase; see N=1 for the host-buffer plugin path"}
    if dl.native:
        def resident_phases(min_sweeps):
            sweeps = phases = 0
            q = None
            while sweeps < min_sweeps:
                gpu.upload_aos(o_pin, edges_p, base=b0, nv_global=nv)
                gpu.dist_set_parts(dl.parts)
                gpu.dist_agree()
                q, iters = gpu.dist_phase_run(-1.0, 1e-6)
                gpu.labels(0)
                sweeps += iters
                phases += 1
            return sweeps, phases, q

        resident_phases(1)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res_sweeps, res_phases, res_q = resident_phases(args.steps)
        torch.cuda.synchronize()
        dist.barrier()
        secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        h2d = torch.tensor([edges_p.nbytes + o_pin.nbytes], dtype=torch.int64, device="cuda")
        dist.all_reduce(h2d)
        e2e = {"value": (ne / 2) * res_sweeps / float(secs.item()), "unit": unit, "sweeps": res_sweeps,
               "h2d_bytes_per_step": int(int(h2d.item()) * res_phases / res_sweeps),
               "d2h_bytes_per_step": int(nv * 8 * res_phases / res_sweeps), "phase_q": res_q,
               "pinned_buffers_numa_local": bool(numa_local),
               "phase_q_equals_device_resident_run": bool(res_q == st["phase_q"]),
               "api": "N-rank distLouvainMethod(.., GpuGraph&) (adapter_resident.cpp): per phase every rank runs "
                      "cvb_graph_upload_host on its rows (pinned AoS) + cvb_dist_set_parts/agree + "
                      "cvb_dist_phase_run + cvb_get_labels; bytes are the sum over ranks"}
    if rank == 0:
        clocks = sampler.stop()
        ms_per_step = float(ms.item()) / args.steps
        peak, peak_src = measured_peak_gbs()
        sweep_ms = float(sw.item())
        bytes_rank = 12 * int(ne_max.item()) + 40 * (b1 - b0)
        achieved = bytes_rank / (sweep_ms * 1e-3) / 1e9
        line = {"metric": metric, "value": (ne / 2) / (ms_per_step * 1e-3), "unit": unit, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32 weights / f64 gains", "data": "synthetic",
                "config": {"workload": workload_name(scale, True),
                           "nv": nv, "csr_entries": ne, "undirected_edges": ne // 2,
                           "parallelism": f"{world} GPUs, vertex-range partition parts[p]=nv*p//P; per sweep: NCCL "
                                          "allgather(labels) + local label-diff community rebuild + "
                                          "allreduce(2 sums), issued from the C ABI (cvb_dist_sweep)",
                           "l2": "per-GPU inputs larger than L2; no flush",
                           "sweep_kernels_ms_max_rank": sweep_ms, "max_local_entries": int(ne_max.item()),
                           "phase_restarts_in_timed_region": st["phases"],
                           **({"stage_ms_rank0": dl.stage_times_ms()} if dl.timing is not None else {})},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                             "kernel": "sweep kernel group of the slowest rank, 12E+40V local bytes"},
                "parity": ({**parity_record(scale, True, st["phase_q"], st["phase_iters"], None),
                            "q_per_iteration_gpu": st["qs"]} if parity_record else None),
                "cpu_baseline": None,
                "e2e": e2e,
                "gpu_launches": int(launches.item()), "clocks": clocks}
        print(json.dumps(line), flush=True)
    bad = bool(parity_record and parity_record(scale, True, st["phase_q"], st["phase_iters"], None)["ok"] is False)
    gpu.close()
    dist.barrier()
    dist.destroy_process_group()
    if bad:
        raise SystemExit("bench.py: the N-GPU modularity differs from the reference — see the parity record")
