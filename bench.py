#!/usr/bin/env python
"""bench.py — edges/sec per Louvain sweep (BASELINE.json metric) on N B200s.

A *step* is one synchronous local-move sweep of phase 0 — neighbour scan +
per-community accumulation + gain argmax + community-delta commit, followed by
the community-info apply + modularity reduction (and, at N > 1, the label /
delta exchange) — on BASELINE.json configs[1]: R-MAT scale-20 (scale-20+log2 N
at N GPUs: weak scaling, per-GPU work fixed).  Steps walk the real algorithm:
iteration 1, 2, ... of the phase; when the phase converges (dQ < 1e-6) the
state is reset to singletons inside the timed region and the walk continues.

  python bench.py --gpus 1 --steps K --warmup W            # this framework
  python bench.py --impl reference --gpus 1 --steps K ...   # reference CPU path (oracle/_ref)

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "undirected_edges_per_second_per_louvain_sweep"
UNIT = "edges/s"
BASE_SCALE = 20  # BASELINE.json configs[1]


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_name(scale, permuted):
    """One name per workload, shared by both arms (the driver compares the strings)."""
    return (f"R-MAT scale-{scale} ef16 (a,b,c,d=.57,.19,.19,.05) seed 2" + (", vertex ids permuted (seed 3)" if permuted else "")
            + ", phase-0 local-move sweeps; metric = undirected edges per second per sweep")


def make_graph(scale, device, permute):
    from cuvite_b200 import graphgen as gg
    return gg.rmat(scale, seed=2, device=device, permute_seed=3 if permute else None)


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU louvain.cpp path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm is the CPU path on ALL host
    # threads at every N (r01: the N>1 reference lines had silently run on one core)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    import numpy as np
    from oracle.binding import Oracle, Reference
    import torch
    scale = BASE_SCALE + (args.gpus - 1).bit_length() if args.scale is None else args.scale
    # same deterministic generator on either device; the GPU only builds the input faster
    g = make_graph(scale, "cuda" if torch.cuda.is_available() else "cpu", args.gpus > 1).cpu()
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy()
    if Reference.available():
        impl, kind = Reference(), "reference"
        h = impl.graph(o, t, w)
        run = lambda: impl.phase_on(h)  # noqa: E731
    else:
        impl, kind = Oracle(), "port"
        run = lambda: impl.phase(o, t, w)  # noqa: E731
    # Thread count: every hardware thread, or one per physical core when that is faster for this memory-bound
    # loop (measured on the round-2 box: 128 SMT threads 13.4 M edges/s, 64 threads 29.7 M) — one untimed run
    # each, the faster setting is the one timed, so the reference arm is the CPU path at its best
    tried = {}
    try:
        import psutil
        phys = psutil.cpu_count(logical=False) or ncpu
    except Exception:
        phys = ncpu
    candidates = sorted({ncpu, min(phys, ncpu)}, reverse=True)
    warm = 0
    try:
        for n in candidates:
            impl.set_num_threads(n)
            pw = run()
            warm += 1
            tried[n] = pw.seconds / max(pw.iters, 1)
        impl.set_num_threads(min(tried, key=tried.get))
    except Exception:
        if not warm:
            run()
            warm = 1
    cores = impl.num_threads()
    steps, secs, iters = 0, 0.0, 0
    t_begin = time.perf_counter()
    while steps < args.steps and (time.perf_counter() - t_begin) < 120.0:
        p = run()
        steps += 1
        secs += p.seconds
        iters += p.iters
    value = (g.ne / 2) * iters / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * secs / iters, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(scale, args.gpus > 1),
                       "nv": g.nv, "csr_entries": g.ne, "sweeps_timed": iters,
                       "step": "one full phase-0 run of louvain.cpp (all its sweeps); ms_per_step is per sweep"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{steps} x phase 0 of the same graph ({iters} sweeps)",
                             "seconds_per_sweep_by_threads": {str(k): v for k, v in tried.items()}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


GOLDEN_Q = os.path.join(ROOT, "tests", "golden", "bench_q.json")


def golden_phase(scale, permuted):
    """Phase-0 (Q, iterations) of a bench graph from the CPU checker (tools/make_bench_golden.py)."""
    try:
        return json.load(open(GOLDEN_Q)).get(f"rmat{scale}{'p' if permuted else ''}")
    except Exception:
        return None


def device_steps(gpu, steps, warmup, ne, nv):
    """W untimed + K timed device-resident steps on the graph `gpu` holds.  One step = one sweep +
    community apply + modularity reduction; a converged phase restarts from singletons inside the
    timed region.  The modularity of every iteration is kept: the first phase's trajectory is the
    parity record of the run."""
    import torch

    def step(state):
        gpu.sweep()
        sums = gpu.apply_and_reduce()
        q = gpu.modularity_from_sums(sums)
        state["sweep_ms"].append(gpu.last_sweep_ms())
        state["it"] += 1
        state["iter_in_phase"].append(state["it"])
        if state["phases"] == 0:
            state["q_first_phase"].append(q)
        if q - state["prev"] < 1e-6:  # phase converged: restart from singletons (inside the timed region)
            if state["phases"] == 0:
                state["phase_q"], state["phase_iters"] = max(state["prev"], -1.0), state["it"]
            gpu.phase_reset()
            state["prev"] = -1.0
            state["phases"] += 1
            state["it"] = 0
        else:
            state["prev"] = q
            gpu.rotate_labels()

    def fresh_state():
        return {"prev": -1.0, "sweep_ms": [], "phases": 0, "it": 0, "iter_in_phase": [], "q_first_phase": [],
                "phase_q": None, "phase_iters": None}

    gpu.phase_begin()
    st = fresh_state()
    for _ in range(warmup):
        step(st)
    gpu.phase_reset()
    st = fresh_state()
    launches0 = gpu.launch_count
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(st)  # every step ends with a stream synchronise (the 2 sums reach the host)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    ms_per_step = wall_ms / steps  # host clock around synchronising steps == device-complete time
    first_ms = [m for i, m in enumerate(st["sweep_ms"]) if st["iter_in_phase"][i] == 1]
    later_ms = [m for i, m in enumerate(st["sweep_ms"]) if st["iter_in_phase"][i] > 1]
    return {"state": st, "ms_per_step": ms_per_step, "value": (ne / 2) / (ms_per_step * 1e-3),
            "sweep_ms": statistics.mean(st["sweep_ms"]), "launches": gpu.launch_count - launches0, "clocks": clocks,
            "first_ms": first_ms, "later_ms": later_ms}


def parity_record(scale, permuted, q_gpu, iters_gpu, cpu):
    """GPU phase-0 result against the reference: the live cpu_baseline leg when it ran, and the
    committed golden value of this graph.  Integer weights: every sum is exact, so the modularity
    must be EQUAL, not close."""
    rec = {"q_gpu": q_gpu, "iters_gpu": iters_gpu}
    ok = q_gpu is not None
    if cpu and cpu.get("q") is not None:
        rec.update(q_reference_live=cpu["q"], iters_reference_live=cpu.get("iters"))
        ok = ok and cpu["q"] == q_gpu and (cpu.get("iters") in (None, iters_gpu))
    gold = golden_phase(scale, permuted)
    if gold:
        rec.update(q_golden=gold["q"], iters_golden=gold["iters"], golden_checker=gold["checker"])
        ok = ok and gold["q"] == q_gpu and gold["iters"] == iters_gpu
    rec["compared"] = bool((cpu and cpu.get("q") is not None) or gold)
    rec["ok"] = bool(ok) if rec["compared"] else None
    return rec


# --------------------------------------------------------------------------- own arm, one GPU
def bench_single(args):
    import numpy as np
    import torch
    from cuvite_b200 import build
    build.build()
    from cuvite_b200.engine import CLASS_NAMES, GpuGraph, gpu_for_louvain_iteration, gpu_reduce_mod, set_gpuDevices
    set_gpuDevices(0)
    torch.cuda.set_device(0)
    scale = BASE_SCALE if args.scale is None else args.scale
    g = make_graph(scale, "cuda", False)
    nv, ne = g.nv, g.ne
    gpu = GpuGraph()
    gpu.upload_device(g.offsets, g.tails, g.weights)
    info = gpu.info()

    # ---- device-resident steps ------------------------------------------------
    dev = device_steps(gpu, args.steps, max(args.warmup, 3), ne, nv)
    st, ms_per_step, value, sweep_ms, launches, clocks = (dev["state"], dev["ms_per_step"], dev["value"],
                                                          dev["sweep_ms"], dev["launches"], dev["clocks"])
    bytes_sweep = 12 * ne + 40 * nv
    peak, peak_src = measured_peak_gbs()
    achieved = bytes_sweep / (sweep_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    for tname in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if scale == BASE_SCALE and os.path.exists(tpath):  # ncu dram__bytes of the same command, committed
            tj = json.load(open(tpath))
            traffic = tj["mean_sweep_dram_bytes"]
            traffic_src = (f"profiles/{tname}: a constant from a committed `ncu --set full` capture of this "
                           "command (dram__bytes_read+write per sweep kernel group), not measured in this run")
            break
    first_ms, later_ms = dev["first_ms"], dev["later_ms"]

    # ---- end to end through the plugin call with pinned host buffers -----------
    o_h = g.offsets.cpu().numpy()
    edges = np.empty(ne, dtype=[("tail", "<i8"), ("weight", "<f8")])
    edges["tail"] = g.tails.cpu().numpy()
    edges["weight"] = g.weights.cpu().numpy()
    edges_pin = torch.from_numpy(edges.view(np.uint8)).pin_memory()
    edges_p = edges_pin.numpy().view(edges.dtype)

    def pinned(a):
        return torch.from_numpy(a).pin_memory().numpy()

    vdeg = pinned((g.offsets[1:] - g.offsets[:-1]).double().cpu().numpy())
    const = 1.0 / float(ne)
    bufs = {k: pinned(np.zeros(nv, dt)) for k, dt in (("past", np.int64), ("curr", np.int64), ("target", np.int64),
                                                       ("csize", np.int64), ("usz", np.int64), ("cdeg", np.float64),
                                                       ("udg", np.float64), ("cw", np.float64))}
    o_pin = pinned(o_h)

    use_torch_host = os.environ.get("BENCH_E2E_HOST", "numpy") == "torch"
    tb = {k: torch.from_numpy(v) for k, v in bufs.items()}  # same memory
    t_vdeg, t_iota = torch.from_numpy(vdeg), torch.arange(nv, dtype=torch.int64)

    def e2e_reset():
        tb["past"].copy_(t_iota)
        tb["curr"].copy_(t_iota)
        tb["csize"].fill_(1)
        tb["cdeg"].copy_(t_vdeg)

    def e2e_steps(k):
        prev, it, uploads = -1.0, 0, 0
        e2e_reset()
        for _ in range(k):
            it += 1
            uploads += it == 1
            if use_torch_host:  # distCleanCWandCU, louvain.cpp:664
                tb["cw"].zero_(), tb["usz"].zero_(), tb["udg"].zero_()
            else:
                bufs["cw"].fill(0), bufs["usz"].fill(0), bufs["udg"].fill(0)
            gpu_for_louvain_iteration(nv, o_pin, edges_p, bufs["curr"], bufs["target"], vdeg, bufs["csize"],
                                      bufs["cdeg"], bufs["usz"], bufs["udg"], {}, {}, {}, const, bufs["cw"], 0,
                                      it, gpu)
            if use_torch_host:  # distUpdateLocalCinfo, louvain.cpp:679
                tb["csize"].add_(tb["usz"]), tb["cdeg"].add_(tb["udg"])
            else:
                np.add(bufs["csize"], bufs["usz"], out=bufs["csize"])
                np.add(bufs["cdeg"], bufs["udg"], out=bufs["cdeg"])
            le = gpu_reduce_mod(bufs["cdeg"], bufs["cw"], gpu, nv)
            q = le[0] * const - le[1] * const * const
            if q - prev < 1e-6:
                e2e_reset()
                prev, it = -1.0, 0
            else:
                prev = q
                bufs["past"], bufs["curr"], bufs["target"] = bufs["curr"], bufs["target"], bufs["past"]
                tb["past"], tb["curr"], tb["target"] = tb["curr"], tb["target"], tb["past"]
        return uploads

    e2e_steps(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    csr_uploads = e2e_steps(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = (ne / 2) * args.steps / e2e_s
    # bytes per step: per-iteration arrays every step; the CSR once per phase restart (num_iters == 1)
    h2d = nv * (8 + 8 + 8) + nv * 16  # plugin inputs + gpu_reduce_mod inputs
    d2h = nv * (8 + 8 + 16) + 16
    csr = ne * 16 + (nv + 1) * 8 + nv * 8
    h2d_step = h2d + csr * csr_uploads / args.steps

    # ---- the same through the OUTER boundary (adapter_resident.cpp's distLouvainMethod(.., GpuGraph&)): one
    # ---- AoS upload from pinned host memory per phase, all sweeps on the device, cvect back to the host
    def resident_phases(min_sweeps):
        sweeps = phases = 0
        while sweeps < min_sweeps:
            gpu.upload_aos(o_pin, edges_p)
            _, iters = gpu.phase_run(-1.0, 1e-6)
            gpu.labels(0)
            sweeps += iters
            phases += 1
        return sweeps, phases

    resident_phases(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res_sweeps, res_phases = resident_phases(args.steps)
    torch.cuda.synchronize()
    res_s = time.perf_counter() - t0
    resident = {"value": (ne / 2) * res_sweeps / res_s, "unit": UNIT, "sweeps": res_sweeps,
                "h2d_bytes_per_step": int(csr * res_phases / res_sweeps), "d2h_bytes_per_step": int(nv * 8 * res_phases / res_sweeps),
                "api": "distLouvainMethod(.., GpuGraph&) on the resident engine: cvb_graph_upload_host (pinned AoS) + "
                       "cvb_phase_run + cvb_get_labels per phase"}

    # ---- CPU baseline beside it (bounded sample: phase 0 of the same graph) -----
    cpu = None
    try:
        if args.no_cpu_baseline:
            raise RuntimeError("skipped by --no-cpu-baseline")
        from oracle.binding import Oracle, Reference
        t_cpu = g.tails.cpu().numpy()
        w_cpu = g.weights.cpu().numpy()
        if Reference.available():
            R = Reference()
            h = R.graph(o_h, t_cpu, w_cpu)
            p = R.phase_on(h)
            R.destroy(h)
            kind, cores = "reference", R.num_threads()
        else:
            O = Oracle()
            p = O.phase(o_h, t_cpu, w_cpu)
            kind, cores = "port", O.num_threads()
        cpu = {"value": (ne / 2) * p.iters / p.seconds, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"1 x phase 0 of the same graph ({p.iters} sweeps, {p.seconds:.2f} s)", "q": p.q,
               "iters": p.iters}
    except Exception as ex:  # the checker is optional for the number itself
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    parity = parity_record(scale, False, st["phase_q"], st["phase_iters"], cpu)
    parity["q_per_iteration_gpu"] = st["q_first_phase"]

    # ---- the north-star configuration on ONE GPU: R-MAT scale-24, same steps, same timing ----------
    scale24 = None
    if args.scale is None and not args.no_scale24:
        try:
            del g
            torch.cuda.empty_cache()
            g24 = make_graph(24, "cuda", False)
            gpu.upload_device(g24.offsets, g24.tails, g24.weights)
            nv24, ne24 = g24.nv, g24.ne
            del g24
            torch.cuda.empty_cache()
            d24 = device_steps(gpu, 8, 3, ne24, nv24)
            b24 = 12 * ne24 + 40 * nv24
            a24 = b24 / (d24["sweep_ms"] * 1e-3) / 1e9
            p24 = parity_record(24, False, d24["state"]["phase_q"], d24["state"]["phase_iters"], None)
            scale24 = {"workload": "R-MAT scale-24 ef16 seed 2 on one GPU (BASELINE.json north_star target config), "
                                   "phase-0 sweeps, 8 timed steps after 3 warm-up steps",
                       "nv": nv24, "csr_entries": ne24, "value": d24["value"], "unit": UNIT,
                       "ms_per_step": d24["ms_per_step"], "sweep_kernels_ms": d24["sweep_ms"],
                       "first_iteration_sweeps_ms": statistics.mean(d24["first_ms"]) if d24["first_ms"] else None,
                       "later_iteration_sweeps_ms": statistics.mean(d24["later_ms"]) if d24["later_ms"] else None,
                       "bound": "hbm", "achieved": a24, "peak": peak, "frac": a24 / peak, "traffic": None,
                       "parity": p24, "clocks": d24["clocks"]}
        except Exception as ex:  # reported, never fatal for the headline line
            scale24 = {"error": repr(ex)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 weights / f64 gains" if info.weight_mode == 1 else "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(scale, False), "nv": nv, "csr_entries": ne,
                       "undirected_edges": ne // 2, "parallelism": "1 GPU",
                       "l2": "inputs (8 B x entries = %d MB) larger than the 126 MB L2; no flush" % (8 * ne >> 20),
                       "sweep_kernels_ms": sweep_ms, "first_sweep_ms": st["sweep_ms"][0],
                       "first_iteration_sweeps_ms": statistics.mean(first_ms) if first_ms else None,
                       "later_iteration_sweeps_ms": statistics.mean(later_ms) if later_ms else None,
                       "phase_restarts_in_timed_region": st["phases"],
                       "classes": {n: int(info.class_count[i]) for i, n in
                                   enumerate(CLASS_NAMES[:info.num_classes])}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "sweep kernel group (label pack + all degree-class kernels of one sweep), "
                                   "12E+40V algorithmic bytes, CUDA events on the engine stream"},
            "roofline_scale24": scale24,
            "parity": parity,
            "cpu_baseline": cpu,
            # e2e = the phase-level reference-facing call, distLouvainMethod(.., GpuGraph&) (louvain.hpp:93-97,
            # main.cpp:344) — the same boundary the reference arm (`--impl reference`: louvain.cpp's
            # distLouvainMethod on the host cores) is timed at; the per-iteration plugin calls are reported beside it
            "e2e": {"value": resident["value"], "unit": UNIT, "h2d_bytes_per_step": resident["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": resident["d2h_bytes_per_step"], "api": resident["api"],
                    "sweeps": resident["sweeps"],
                    "per_iteration_plugin": {
                        "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step),
                        "d2h_bytes_per_step": int(d2h),
                        "api": "gpu_for_louvain_iteration + gpu_reduce_mod (plugin calls, pinned host buffers)"}},
            "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line), flush=True)
    gpu.close()
    bad = parity["ok"] is False or (isinstance(scale24, dict) and scale24.get("parity", {}).get("ok") is False)
    if bad:
        sys.stderr.write("bench.py: GPU modularity differs from the reference — see the parity record\n")
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuvite_b200", choices=["cuvite_b200", "reference"])
    ap.add_argument("--scale", type=int, default=None, help="override the R-MAT scale (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (debug)")
    ap.add_argument("--no-scale24", action="store_true", help="skip the roofline_scale24 block (debug)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus == 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return bench_single(args)
    from cuvite_b200 import dist
    return dist.bench_multi(args, METRIC, UNIT, BASE_SCALE, measured_peak_gbs, ClockSampler, workload_name, parity_record)


if __name__ == "__main__":
    main()
