"""Host-side mirror of the reference's GPU interface, over the C ABI.

``GpuGraph`` plays the role of ``CuVite::GpuGraph`` (GpuGraph.cuh:31-225): it
owns the device state of one rank.  The module-level functions keep the names
and argument meaning of the reference's plugin entry points:

* ``set_gpuDevices(me)``               — louvain_cuda.cu:1634-1669
* ``gpu_for_louvain_iteration(...)``   — louvain_cuda.cu:1671-1686, 2235-2735
* ``gpu_reduce_mod(...)``              — modularity.cu:102-141
* ``distLouvainMethod(...)``           — louvain.cpp:591-753 (GPU overload)

so the parity tests read like a test of the reference's own interface.  STL
containers become numpy arrays / dicts; everything runs through
libcuvite_b200.so — there is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import GraphInfo, check

WEIGHTS_U32, WEIGHTS_F64 = 1, 2
LABELS_PAST, LABELS_CURR, LABELS_TARGET = 0, 1, 2
BUF = dict(labels_global=0, labels_curr=1, labels_target=2, labels_past=3, cdeg=4, csize=5, ddeg=6, dsize=7,
           vdeg=8, cw=9, offsets=10, tails=11, weights=12, labels_global_next=13, vdeg_global=14, sums=15,
           pair_keys=16, pair_weights=17)
CLASS_NAMES = ["deg0", "d4", "d8", "d16", "d32", "d64", "d128", "d256", "d512", "d1k", "d2k", "d4k", "d8k", "c16k", "c32k", "c64k", "giant"]


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def set_gpuDevices(me: int) -> None:
    check(_lib.load().cvb_set_device(int(me)), "set_gpuDevices")


def device_count() -> int:
    return _lib.load().cvb_device_count()


@dataclass
class DeviceArray:
    """Raw device buffer exposed through __cuda_array_interface__ (for torch.as_tensor)."""
    ptr: int
    shape: tuple
    typestr: str

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.typestr, "data": (self.ptr, False), "version": 3,
                "strides": None}


class GpuGraph:
    """Device state of one rank (one GPU, one vertex-range partition)."""

    def __init__(self, stream: int | None = None):
        self._L = _lib.load()
        self._h = self._L.cvb_create(C.c_void_p(stream) if stream else None)
        if not self._h:
            raise _lib.CuviteError("cvb_create failed: " + self._L.cvb_last_error().decode())
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.cvb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- graph -----------------------------------------------------------
    def upload(self, offsets, tails, weights, base: int = 0, nv_global: int | None = None):
        """CSR rows [base, base+nv) in the boundary layout (int64 / int64 / float64, host)."""
        offsets, tails, weights = _arr(offsets, np.int64), _arr(tails, np.int64), _arr(weights, np.float64)
        nv = offsets.size - 1
        check(self._L.cvb_graph_upload_host(self._h, nv, base, nv if nv_global is None else nv_global,
                                            _ptr(offsets), None, _ptr(tails), _ptr(weights)), "upload")
        return self

    def upload_aos(self, offsets, edges, base: int = 0, nv_global: int | None = None):
        """Same, with the reference's Edge{int64 tail; double weight} AoS array."""
        offsets = _arr(offsets, np.int64)
        nv = offsets.size - 1
        assert edges.dtype.itemsize == 16
        check(self._L.cvb_graph_upload_host(self._h, nv, base, nv if nv_global is None else nv_global,
                                            _ptr(offsets), _ptr(np.ascontiguousarray(edges)), None, None),
              "upload_aos")
        return self

    def load_vite(self, path: str, rank: int = 0, nparts: int = 1, balanced: bool = False):
        """cvb_graph_load_vite: this rank's rows of a Vite file -> pinned staging -> device.  Returns the
        partition record (nv_global, ne_global, base, nv_local, first_entry, ne_local)."""
        from ._lib import VitePart
        part = VitePart()
        check(self._L.cvb_graph_load_vite(self._h, path.encode(), rank, nparts, 1 if balanced else 0, C.byref(part)),
              "graph_load_vite")
        return part

    def upload_device(self, offsets, tails, weights, base: int = 0, nv_global: int | None = None):
        """Device-resident CSR (torch CUDA tensors: int64, int64, float64)."""
        import torch
        torch.cuda.synchronize(offsets.device)  # the engine reads on its own stream: producers must be done
        nv = offsets.numel() - 1
        check(self._L.cvb_graph_upload_device(self._h, nv, base, nv if nv_global is None else nv_global,
                                              C.c_void_p(offsets.data_ptr()), C.c_void_p(tails.data_ptr()),
                                              C.c_void_p(weights.data_ptr())), "upload_device")
        return self

    def info(self) -> GraphInfo:
        gi = GraphInfo()
        check(self._L.cvb_graph_get_info(self._h, C.byref(gi)), "graph_get_info")
        return gi

    # ---- phase, stepwise ---------------------------------------------------
    def phase_begin(self, total_weight: float = 0.0):
        check(self._L.cvb_phase_begin(self._h, float(total_weight)), "phase_begin")

    def phase_reset(self):
        check(self._L.cvb_phase_reset(self._h), "phase_reset")

    def sweep(self):
        check(self._L.cvb_sweep(self._h), "sweep")

    def apply_and_reduce(self) -> np.ndarray:
        sums = (C.c_double * 2)()
        check(self._L.cvb_apply_and_reduce(self._h, sums), "apply_and_reduce")
        return np.array([sums[0], sums[1]])

    def rebuild_and_reduce(self) -> np.ndarray:
        sums = (C.c_double * 2)()
        check(self._L.cvb_rebuild_and_reduce(self._h, sums), "rebuild_and_reduce")
        return np.array([sums[0], sums[1]])

    def rebuild_and_reduce_async(self):
        check(self._L.cvb_rebuild_and_reduce_async(self._h), "rebuild_and_reduce_async")

    def modularity_from_sums(self, sums) -> float:
        s = (C.c_double * 2)(float(sums[0]), float(sums[1]))
        return self._L.cvb_modularity_from_sums(self._h, s)

    def rotate_labels(self):
        check(self._L.cvb_rotate_labels(self._h), "rotate_labels")

    def sync(self):
        check(self._L.cvb_sync(self._h), "sync")

    # ---- N > 1 ranks: the phase over NCCL, behind the C ABI (cvb_dist_*) --------
    @staticmethod
    def dist_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().cvb_dist_unique_id(buf, 128), "dist_unique_id")
        return buf.raw

    def dist_init(self, uid: bytes | None, rank: int, world: int):
        b = C.create_string_buffer(uid, 128) if uid is not None else None
        check(self._L.cvb_dist_init(self._h, b, 128 if uid is not None else 0, rank, world), "dist_init")

    def dist_set_parts(self, parts=None):
        a = None if parts is None else np.ascontiguousarray(np.asarray(parts, dtype=np.int64))
        check(self._L.cvb_dist_set_parts(self._h, None if a is None else _ptr(a)), "dist_set_parts")

    def dist_agree(self):
        check(self._L.cvb_dist_agree(self._h), "dist_agree")

    def dist_phase_begin(self):
        check(self._L.cvb_dist_phase_begin(self._h), "dist_phase_begin")

    def dist_sweep(self) -> float:
        q = C.c_double(0)
        check(self._L.cvb_dist_sweep(self._h, C.byref(q)), "dist_sweep")
        return q.value

    def dist_advance(self):
        check(self._L.cvb_dist_advance(self._h), "dist_advance")

    def dist_phase_run(self, lower: float = -1.0, thresh: float = 1e-6, max_iters: int = 0):
        q, it = C.c_double(0), C.c_int(0)
        check(self._L.cvb_dist_phase_run(self._h, lower, thresh, max_iters, C.byref(q), C.byref(it)), "dist_phase_run")
        return q.value, it.value

    def dist_stage_ms(self):
        out = (C.c_double * 5)()
        check(self._L.cvb_dist_stage_ms(self._h, out), "dist_stage_ms")
        return dict(zip(("sweep_kernels", "all_gather_labels", "label_diff_and_reduce", "all_reduce_sums", "read_back"),
                        [float(x) for x in out]))

    def dist_finalize(self):
        check(self._L.cvb_dist_finalize(self._h), "dist_finalize")

    # ---- phase, whole ------------------------------------------------------
    def phase_run(self, lower: float = -1.0, thresh: float = 1e-6, max_iters: int = 0):
        q, it = C.c_double(0), C.c_int(0)
        check(self._L.cvb_phase_run(self._h, lower, thresh, max_iters, C.byref(q), C.byref(it)), "phase_run")
        return q.value, it.value

    def coarsen(self) -> int:
        n = C.c_int64(0)
        check(self._L.cvb_coarsen(self._h, C.byref(n)), "coarsen")
        return n.value

    def phase_run_et(self, et_type: int, et_delta: float = 1.0, lower: float = -1.0, thresh: float = 1e-6):
        """cvb_phase_run_et: one phase with early termination (options -t et_type, -a et_delta)."""
        q, it = C.c_double(0), C.c_int(0)
        check(self._L.cvb_phase_run_et(self._h, lower, thresh, et_type, et_delta, C.byref(q), C.byref(it)), "phase_run_et")
        return q.value, it.value

    def phase_run_color(self, max_colors: int, lower: float = -1.0, thresh: float = 1e-6, et_type: int = 0,
                        et_delta: float = 1.0):
        """cvb_phase_run_color[_et]: colouring (option -c max_colors) + the coloured phase, optionally with early
        termination (-t et_type, -a et_delta).  Returns (Q, iterations, number of colours as the reference
        prints it)."""
        q, it, nc = C.c_double(0), C.c_int(0), C.c_int(0)
        check(self._L.cvb_phase_run_color_et(self._h, lower, thresh, max_colors, et_type, et_delta, C.byref(q),
                                             C.byref(it), C.byref(nc)), "phase_run_color")
        return q.value, it.value, nc.value

    def color_graph(self, max_colors: int) -> int:
        """cvb_color_graph: the colouring alone (what option -d adds to the plain phase on one rank).  Returns the
        number of colours as the reference prints it."""
        nc = C.c_int(0)
        check(self._L.cvb_color_graph(self._h, max_colors, C.byref(nc)), "color_graph")
        return nc.value

    def colors(self) -> np.ndarray:
        """cvb_get_colors: the colours of the last coloured phase (-1 = uncoloured)."""
        out = np.empty(self.info().nv_local, np.int32)
        check(self._L.cvb_get_colors(self._h, _ptr(out)), "get_colors")
        return out

    def run(self, max_levels: int = 200, threshold_cycling: bool = False, et_type: int = 0, et_delta: float = 1.0,
            max_colors: int | None = None):
        """cvb_louvain_run[_ex/_et/_color]: the phase loop of main.cpp:218-495 (threshold_cycling = option -i,
        et_type / et_delta = options -t / -a, max_colors = option -c)."""
        nv0 = self.info().nv_local
        lq = np.zeros(max_levels)
        li = np.zeros(max_levels, np.int32)
        ln = np.zeros(max_levels, np.int64)
        comm = np.empty(nv0, np.int64)
        if max_colors is not None:
            n = self._L.cvb_louvain_run_color_et(self._h, max_levels, 1 if threshold_cycling else 0, max_colors,
                                                 et_type, et_delta, _ptr(lq), _ptr(li), _ptr(ln), _ptr(comm))
        else:
            n = self._L.cvb_louvain_run_et(self._h, max_levels, 1 if threshold_cycling else 0, et_type, et_delta,
                                           _ptr(lq), _ptr(li), _ptr(ln), _ptr(comm))
        if n < 0:
            check(n, "louvain_run")
        return list(lq[:n]), [int(x) for x in li[:n]], [int(x) for x in ln[:n]], comm

    # ---- N-rank coarsening stages (device tensors in, device tensors out) ------
    def renumber(self, labels):
        """cvb_renumber: ``labels`` is a CUDA int32 tensor (the all-gathered cvect).  Returns
        (newid int32 tensor, number of communities)."""
        import torch
        assert labels.is_cuda and labels.dtype == torch.int32 and labels.is_contiguous()
        newid = torch.empty_like(labels)
        torch.cuda.current_stream().synchronize()  # the engine stream may differ from torch's
        n = C.c_int64(0)
        check(self._L.cvb_renumber(self._h, labels.data_ptr(), labels.numel(), newid.data_ptr(), C.byref(n)), "renumber")
        return newid, n.value

    def contract_local(self, newid, n_comm: int):
        """cvb_contract_local: distinct (src << 32 | dst) pairs of this engine's rows under
        ``newid``.  Returns (keys int64 tensor, weights tensor) — zero-copy views of engine
        buffers, valid until the next coarsening call."""
        import torch
        m = C.c_int64(0)
        torch.cuda.current_stream().synchronize()
        check(self._L.cvb_contract_local(self._h, newid.data_ptr(), n_comm, C.byref(m)), "contract_local")
        wt = "<u4" if self.info().weight_mode == WEIGHTS_U32 else "<f8"
        if m.value == 0:
            return (torch.empty(0, dtype=torch.int64, device="cuda"),
                    torch.empty(0, dtype=torch.int32 if wt == "<u4" else torch.float64, device="cuda"))
        keys = torch.as_tensor(DeviceArray(int(self._L.cvb_buffer(self._h, BUF["pair_keys"])), (m.value,), "<i8"),
                               device="cuda")
        w = torch.as_tensor(DeviceArray(int(self._L.cvb_buffer(self._h, BUF["pair_weights"])), (m.value,),
                                        "<i4" if wt == "<u4" else "<f8"), device="cuda")
        return keys, w

    def graph_from_pairs(self, nv_local: int, base: int, nv_global: int, keys, weights):
        """cvb_graph_from_pairs: the pairs routed to this rank become the engine's graph."""
        import torch
        torch.cuda.current_stream().synchronize()
        keys, weights = keys.contiguous(), weights.contiguous()
        check(self._L.cvb_graph_from_pairs(self._h, nv_local, base, nv_global, keys.data_ptr() if keys.numel() else None,
                                           weights.data_ptr() if keys.numel() else None, keys.numel()),
              "graph_from_pairs")

    def download_graph(self):
        gi = self.info()
        o = np.empty(gi.nv_local + 1, np.int64)
        t = np.empty(gi.ne_local, np.int64)
        w = np.empty(gi.ne_local, np.float64)
        check(self._L.cvb_graph_download(self._h, _ptr(o), _ptr(t), _ptr(w)), "graph_download")
        return o, t, w

    # ---- results -------------------------------------------------------------
    def labels(self, which: int = LABELS_PAST) -> np.ndarray:
        out = np.empty(self.info().nv_local, np.int64)
        check(self._L.cvb_get_labels(self._h, which, _ptr(out)), "get_labels")
        return out

    def cluster_weight(self) -> np.ndarray:
        out = np.empty(self.info().nv_local, np.float64)
        check(self._L.cvb_get_cluster_weight(self._h, _ptr(out)), "get_cluster_weight")
        return out

    def cinfo(self):
        n = self.info().nv_global
        size, deg = np.empty(n, np.int64), np.empty(n, np.float64)
        check(self._L.cvb_get_cinfo(self._h, _ptr(size), _ptr(deg)), "get_cinfo")
        return size, deg

    # ---- raw buffers / instrumentation ----------------------------------------
    def buffer(self, name: str) -> DeviceArray:
        gi = self.info()
        shapes = dict(labels_global=(gi.nv_global, "<i4"), labels_curr=(gi.nv_local, "<i4"),
                      labels_target=(gi.nv_local, "<i4"), labels_past=(gi.nv_local, "<i4"),
                      cdeg=(gi.nv_global, "<f8"), csize=(gi.nv_global, "<i4"), ddeg=(gi.nv_global, "<f8"),
                      dsize=(gi.nv_global, "<i4"), vdeg=(gi.nv_local, "<f8"), cw=(gi.nv_local, "<f8"),
                      labels_global_next=(gi.nv_global, "<i4"), vdeg_global=(gi.nv_global, "<f8"), sums=(2, "<f8"),
                      offsets=(gi.nv_local + 1, "<i8"), tails=(gi.ne_local, "<i4"),
                      weights=(gi.ne_local, "<u4" if gi.weight_mode == WEIGHTS_U32 else "<f8"))
        n, ts = shapes[name]
        return DeviceArray(int(self._L.cvb_buffer(self._h, BUF[name]) or 0), (int(n),), ts)

    @property
    def stream(self) -> int:
        return int(self._L.cvb_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._L.cvb_launch_count(self._h))

    def last_sweep_ms(self) -> float:
        return float(self._L.cvb_last_sweep_ms(self._h))

    @property
    def handle(self):
        return self._h


def vite_partition(path: str, rank: int, nparts: int, balanced: bool = False):
    """cvb_vite_partition (pure host code of the C ABI): (partition record, all range starts)."""
    from ._lib import VitePart, load
    part, parts = VitePart(), np.zeros(nparts + 1, np.int64)
    check(load().cvb_vite_partition(path.encode(), rank, nparts, 1 if balanced else 0, C.byref(part), _ptr(parts)),
          "vite_partition")
    return part, parts


def vite_read_rows(path: str, part):
    """cvb_vite_read_rows (pure host code): (offsets rebased to 0, AoS edges) of the partition."""
    from ._lib import load
    off = np.empty(part.nv_local + 1, np.int64)
    edges = np.empty(part.ne_local, dtype=[("tail", "<i8"), ("weight", "<f8")])
    check(load().cvb_vite_read_rows(path.encode(), C.byref(part), _ptr(off), _ptr(edges) if part.ne_local else None),
          "vite_read_rows")
    return off, edges


def gpu_for_louvain_iteration(nv, offsets, edges_aos, currComm, targetComm, vDegree, localCinfo_size,
                              localCinfo_degree, localCupdate_size, localCupdate_degree, remoteComm: dict,
                              remoteCinfo: dict, remoteCupdate: dict, constantForSecondTerm, clusterWeight,
                              me, numIters, gpu_graph: GpuGraph, base=0, nv_global=None) -> int:
    """The reference's per-iteration plugin call (louvain.hpp:232-245) with the
    STL containers as numpy arrays / dicts: remoteComm {ghost vertex: community},
    remoteCinfo {community: (size, degree)} (iterated in ascending id like
    std::map), remoteCupdate {community: [size, degree]} updated in place.
    targetComm, clusterWeight and the Cupdate arrays are written in place."""
    L = _lib.load()
    gv = np.fromiter(remoteComm.keys(), np.int64, len(remoteComm))
    gc = np.fromiter(remoteComm.values(), np.int64, len(remoteComm))
    rid = np.array(sorted(remoteCinfo.keys()), np.int64)
    rsz = np.array([remoteCinfo[k][0] for k in rid], np.int64)
    rdg = np.array([remoteCinfo[k][1] for k in rid], np.float64)
    usz = np.zeros(rid.size, np.int64)
    udg = np.zeros(rid.size, np.float64)
    nvg = nv if nv_global is None else nv_global
    rc = L.cvb_louvain_iteration(gpu_graph.handle, nv, base, nvg, _ptr(_arr(offsets, np.int64)), _ptr(edges_aos),
                                 _ptr(_arr(currComm, np.int64)), _ptr(_arr(vDegree, np.float64)),
                                 _ptr(_arr(localCinfo_size, np.int64)), _ptr(_arr(localCinfo_degree, np.float64)),
                                 gv.size, _ptr(gv), _ptr(gc), rid.size, _ptr(rid), _ptr(rsz), _ptr(rdg),
                                 float(constantForSecondTerm), int(numIters), _ptr(targetComm),
                                 _ptr(clusterWeight), _ptr(localCupdate_size), _ptr(localCupdate_degree),
                                 _ptr(usz), _ptr(udg))
    if rc <= 0:
        check(rc if rc < 0 else -4, "gpu_for_louvain_iteration")
    for k, s, d in zip(rid.tolist(), usz.tolist(), udg.tolist()):
        remoteCupdate[k][0] += s
        remoteCupdate[k][1] += d
    return rc


def gpu_reduce_mod(localCinfo_degree, clusterWeight, gpu_graph: GpuGraph, nv: int) -> np.ndarray:
    """modularity.cu:102-141: returns le_la_xx = [sum clusterWeight, sum degree^2]."""
    out = (C.c_double * 2)()
    check(_lib.load().cvb_reduce_mod(gpu_graph.handle, nv, _ptr(_arr(localCinfo_degree, np.float64)),
                                     _ptr(_arr(clusterWeight, np.float64)), out), "gpu_reduce_mod")
    return np.array([out[0], out[1]])


def distLouvainMethod(offsets, tails, weights, lower: float, thresh: float, gpu_graph: GpuGraph):
    """One-rank GPU overload of the phase (louvain.cpp:591-753): returns
    (modularity, cvect, iters)."""
    gpu_graph.upload(offsets, tails, weights)
    q, iters = gpu_graph.phase_run(lower, thresh)
    return q, gpu_graph.labels(LABELS_PAST), iters
