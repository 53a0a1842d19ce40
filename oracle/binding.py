"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle``  — oracle/liblouvain_oracle.so, the plain-C restatement
  (louvain_oracle.c) of the reference's louvain.cpp / rebuild.cpp one-rank path.
* ``Reference`` — oracle/_ref/libcuvite_ref.so, the *real* reference sources
  compiled by build_ref.sh (present wherever it was built; travels to the GPU
  box as a prebuilt binary).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _csr(offsets, tails, weights):
    return (np.ascontiguousarray(np.asarray(offsets), dtype=np.int64),
            np.ascontiguousarray(np.asarray(tails), dtype=np.int64),
            np.ascontiguousarray(np.asarray(weights), dtype=np.float64))


@dataclass
class PhaseResult:
    q: float
    iters: int
    cvect: np.ndarray
    seconds: float = 0.0
    targets: np.ndarray | None = None  # [iters_traced, nv]
    cw: np.ndarray | None = None
    qs: np.ndarray | None = None
    csize: np.ndarray | None = None
    cdeg: np.ndarray | None = None


@dataclass
class RunResult:
    level_q: list = field(default_factory=list)
    level_iters: list = field(default_factory=list)  # cumulative, like main.cpp:450
    level_nv: list = field(default_factory=list)
    level_secs: list = field(default_factory=list)
    level_rebuild_secs: list = field(default_factory=list)
    comm: np.ndarray | None = None  # composed assignment of the original vertices


class Oracle:
    """The C restatement."""

    path = os.path.join(_HERE, "liblouvain_oracle.so")

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(f"{self.path} missing: run `make -C oracle liblouvain_oracle.so`")
        L = self.lib = C.CDLL(self.path)
        L.orc_phase.restype = C.c_double
        L.orc_phase.argtypes = [C.c_int64, _i64p, _i64p, _f64p, C.c_double, C.c_double, _i64p,
                                C.POINTER(C.c_int), C.c_int] + [C.c_void_p] * 5
        L.orc_renumber.restype = C.c_int64
        L.orc_renumber.argtypes = [C.c_int64, _i64p, _i64p]
        L.orc_rebuild.restype = C.c_int64
        L.orc_rebuild.argtypes = [C.c_int64, _i64p, _i64p, _f64p, _i64p, _i64p, C.c_int64, _i64p, _i64p, _f64p]
        L.orc_sweep.restype = None
        L.orc_sweep.argtypes = [C.c_int64, C.c_int64, C.c_int64, _i64p, _i64p, _f64p, _i64p, _f64p, _f64p,
                                _i64p, C.c_double, _i64p, _f64p, _f64p, _i64p]
        L.orc_sum_vertex_degree.argtypes = [C.c_int64, _i64p, _f64p, _f64p]
        L.orc_modularity.restype = C.c_double
        L.orc_modularity.argtypes = [C.c_int64, _f64p, _f64p, C.c_double, _f64p]
        L.orc_num_threads.restype = C.c_int

    def num_threads(self) -> int:
        return self.lib.orc_num_threads()

    def set_num_threads(self, n: int):
        self.lib.orc_set_num_threads(int(n))

    def phase(self, offsets, tails, weights, lower=-1.0, thresh=1e-6, trace=0) -> PhaseResult:
        import time
        offsets, tails, weights = _csr(offsets, tails, weights)
        nv = offsets.size - 1
        cvect = np.empty(nv, np.int64)
        it = C.c_int(0)
        tg = np.empty((trace, nv), np.int64) if trace else None
        cw = np.empty((trace, nv), np.float64) if trace else None
        qs = np.full(trace, np.nan) if trace else None
        cs = np.empty((trace, nv), np.int64) if trace else None
        cd = np.empty((trace, nv), np.float64) if trace else None
        t0 = time.perf_counter()
        q = self.lib.orc_phase(nv, offsets, tails, weights, lower, thresh, cvect, C.byref(it), trace,
                               _opt(tg), _opt(cw), _opt(qs), _opt(cs), _opt(cd))
        dt = time.perf_counter() - t0
        n = min(it.value, trace)
        return PhaseResult(q, it.value, cvect, dt, None if tg is None else tg[:n], None if cw is None else cw[:n],
                           None if qs is None else qs[:n], None if cs is None else cs[:n],
                           None if cd is None else cd[:n])

    def sweep(self, nvl, base, nvg, offsets, tails, weights, curr, vdeg, cdeg, csize, constant):
        """One synchronous sweep over rows [base, base+nvl); returns target, cw, ddeg, dsize."""
        offsets, tails, weights = _csr(offsets, tails, weights)
        target = np.empty(nvl, np.int64)
        cw = np.empty(nvl, np.float64)
        ddeg = np.empty(nvg, np.float64)
        dsize = np.empty(nvg, np.int64)
        self.lib.orc_sweep(nvl, base, nvg, offsets, tails, weights, np.ascontiguousarray(curr, np.int64),
                           np.ascontiguousarray(vdeg, np.float64), np.ascontiguousarray(cdeg, np.float64),
                           np.ascontiguousarray(csize, np.int64), constant, target, cw, ddeg, dsize)
        return target, cw, ddeg, dsize

    def vertex_degree(self, offsets, weights):
        offsets = np.ascontiguousarray(offsets, np.int64)
        out = np.empty(offsets.size - 1, np.float64)
        self.lib.orc_sum_vertex_degree(offsets.size - 1, offsets, np.ascontiguousarray(weights, np.float64), out)
        return out

    def modularity(self, cw, cdeg, constant):
        sums = np.empty(2, np.float64)
        q = self.lib.orc_modularity(len(cw), np.ascontiguousarray(cw, np.float64),
                                    np.ascontiguousarray(cdeg, np.float64), constant, sums)
        return q, sums

    def rebuild(self, offsets, tails, weights, cvect):
        """Returns (new offsets, tails, weights, newid[old community] or -1)."""
        offsets, tails, weights = _csr(offsets, tails, weights)
        nv = offsets.size - 1
        cvect = np.ascontiguousarray(cvect, np.int64)
        newid = np.empty(nv, np.int64)
        nnv = self.lib.orc_renumber(nv, cvect, newid)
        no = np.empty(nnv + 1, np.int64)
        nt = np.empty(tails.size, np.int64)
        nw = np.empty(tails.size, np.float64)
        nne = self.lib.orc_rebuild(nv, offsets, tails, weights, cvect, newid, nnv, no, nt, nw)
        return no, nt[:nne].copy(), nw[:nne].copy(), newid

    def run(self, offsets, tails, weights, max_levels=200) -> RunResult:
        """The default-flag phase driver of main.cpp:218-495 (see ref_harness.cpp:ref_louvain_run)."""
        offsets, tails, weights = _csr(offsets, tails, weights)
        nv0 = offsets.size - 1
        res = RunResult(comm=np.arange(nv0, dtype=np.int64))
        curr_mod, prev_mod, tot = -1.0, -1.0, 0
        import time
        while True:
            p = self.phase(offsets, tails, weights, lower=curr_mod, thresh=1e-6)
            curr_mod = p.q
            tot += p.iters
            if curr_mod - prev_mod > 1e-6:
                t0 = time.perf_counter()
                nvl = offsets.size - 1
                offsets, tails, weights, newid = self.rebuild(offsets, tails, weights, p.cvect)
                res.comm = newid[p.cvect[res.comm]]
                res.level_rebuild_secs.append(time.perf_counter() - t0)
            else:
                break
            res.level_q.append(curr_mod)
            res.level_iters.append(tot)
            res.level_nv.append(nvl)
            res.level_secs.append(p.seconds)
            prev_mod, curr_mod = curr_mod, -1.0
            if len(res.level_q) == max_levels or tot > 10000:
                break
        return res


class Reference:
    """The real reference CPU path (oracle/_ref/libcuvite_ref.so)."""

    path = os.path.join(_HERE, "_ref", "libcuvite_ref.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.path)

    def __init__(self):
        if not self.available():
            raise FileNotFoundError(f"{self.path} missing: run oracle/build_ref.sh where /root/reference exists")
        L = self.lib = C.CDLL(self.path)
        L.ref_graph_create.restype = C.c_void_p
        L.ref_graph_create.argtypes = [C.c_int64, C.c_int64, _i64p, _i64p, _f64p]
        L.ref_graph_load.restype = C.c_void_p
        L.ref_graph_load.argtypes = [C.c_char_p]
        L.ref_graph_destroy.argtypes = [C.c_void_p]
        L.ref_graph_nv.restype = C.c_int64
        L.ref_graph_nv.argtypes = [C.c_void_p]
        L.ref_graph_ne.restype = C.c_int64
        L.ref_graph_ne.argtypes = [C.c_void_p]
        L.ref_graph_get.argtypes = [C.c_void_p, _i64p, _i64p, _f64p]
        L.ref_louvain_phase.restype = C.c_double
        L.ref_louvain_phase.argtypes = [C.c_void_p, C.c_double, C.c_double, _i64p, C.POINTER(C.c_int),
                                        C.POINTER(C.c_double)]
        L.ref_phase_trace.restype = C.c_int
        L.ref_phase_trace.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 7
        L.ref_rebuild.argtypes = [C.c_void_p, _i64p, C.POINTER(C.c_double)]
        L.ref_louvain_run.restype = C.c_int
        L.ref_louvain_run.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.ref_num_threads.restype = C.c_int

    def num_threads(self) -> int:
        return self.lib.ref_num_threads()

    def set_num_threads(self, n: int):
        self.lib.ref_set_num_threads(int(n))

    def graph(self, offsets, tails, weights):
        offsets, tails, weights = _csr(offsets, tails, weights)
        return self.lib.ref_graph_create(offsets.size - 1, tails.size, offsets, tails, weights)

    def balance_edges(self, path: str, nprocs: int) -> np.ndarray:
        """balanceEdges (distgraph.cpp:22-66) on a Vite file: the -b vertex ranges."""
        out = np.zeros(nprocs + 1, np.int64)
        self.lib.ref_balance_edges.argtypes = [C.c_int, C.c_char_p, _i64p]
        self.lib.ref_balance_edges.restype = None
        self.lib.ref_balance_edges(nprocs, path.encode(), out)
        return out

    def graph_load(self, path: str):
        return self.lib.ref_graph_load(path.encode())

    def graph_arrays(self, h):
        nv, ne = self.lib.ref_graph_nv(h), self.lib.ref_graph_ne(h)
        o, t, w = np.empty(nv + 1, np.int64), np.empty(ne, np.int64), np.empty(ne, np.float64)
        self.lib.ref_graph_get(h, o, t, w)
        return o, t, w

    def destroy(self, h):
        self.lib.ref_graph_destroy(h)

    def phase_on(self, h, lower=-1.0, thresh=1e-6) -> PhaseResult:
        nv = self.lib.ref_graph_nv(h)
        cvect = np.empty(nv, np.int64)
        it, secs = C.c_int(0), C.c_double(0)
        q = self.lib.ref_louvain_phase(h, lower, thresh, cvect, C.byref(it), C.byref(secs))
        return PhaseResult(q, it.value, cvect, secs.value)

    def phase(self, offsets, tails, weights, lower=-1.0, thresh=1e-6, trace=0) -> PhaseResult:
        h = self.graph(offsets, tails, weights)
        try:
            if not trace:
                return self.phase_on(h, lower, thresh)
            nv = self.lib.ref_graph_nv(h)
            tg = np.empty((trace, nv), np.int64)
            cw = np.empty((trace, nv), np.float64)
            qs = np.full(trace, np.nan)
            cs = np.empty((trace, nv), np.int64)
            cd = np.empty((trace, nv), np.float64)
            cvect = np.empty(nv, np.int64)
            qret = C.c_double(0)
            it = self.lib.ref_phase_trace(h, lower, thresh, trace, _opt(tg), _opt(cw), _opt(qs), _opt(cs),
                                          _opt(cd), _opt(cvect), C.cast(C.byref(qret), C.c_void_p))
            n = min(it, trace)
            return PhaseResult(qret.value, it, cvect, 0.0, tg[:n], cw[:n], qs[:n], cs[:n], cd[:n])
        finally:
            self.destroy(h)

    def rebuild(self, offsets, tails, weights, cvect):
        """Phase (to set up the ghost lists) is NOT re-run: one rank has none."""
        h = self.graph(offsets, tails, weights)
        try:
            # exchangeVertexReqs leaves ssizes/rsizes sized nprocs; a phase call sets them up
            self.phase_on(h)
            self.lib.ref_rebuild(h, np.ascontiguousarray(cvect, np.int64), None)
            return self.graph_arrays(h)
        finally:
            self.destroy(h)

    def run(self, offsets, tails, weights, max_levels=200) -> RunResult:
        h = self.graph(offsets, tails, weights)
        try:
            nv0 = self.lib.ref_graph_nv(h)
            lq = np.zeros(max_levels)
            li = np.zeros(max_levels, np.int32)
            ln = np.zeros(max_levels, np.int64)
            ls = np.zeros(max_levels)
            lr = np.zeros(max_levels)
            comm = np.empty(nv0, np.int64)
            n = self.lib.ref_louvain_run(h, max_levels, _opt(lq), _opt(li), _opt(ln), _opt(ls), _opt(lr), _opt(comm))
            return RunResult(list(lq[:n]), [int(x) for x in li[:n]], [int(x) for x in ln[:n]], list(ls[:n]),
                             list(lr[:n]), comm)
        finally:
            self.destroy(h)
