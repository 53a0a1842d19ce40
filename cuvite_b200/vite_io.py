"""Vite binary graph format reader / writer.

Layout (little-endian, no padding; /root/reference/distgraph.cpp:99-100,131,170
and SURVEY.md A4), 64-bit build: ``int64 nv; int64 ne; int64 offsets[nv+1];
struct {int64 tail; double weight;} edges[ne]``.  ``ne`` counts directed CSR
entries.  The reader also implements the reference's vertex-range partition
(``parts[p] = nv*p // P``, distgraph.cpp:102,118-119) so that rank ``p`` of
``P`` loads only its rows, with offsets rebased to 0 like distgraph.cpp:194-195.
"""
from __future__ import annotations

import numpy as np
import torch

from .graphgen import CSRGraph

_EDGE_DT = np.dtype([("tail", "<i8"), ("weight", "<f8")])


def write_vite(path: str, g: CSRGraph) -> None:
    g = g.cpu()
    with open(path, "wb") as f:
        np.array([g.nv, g.ne], dtype="<i8").tofile(f)
        g.offsets.numpy().astype("<i8", copy=False).tofile(f)
        edges = np.empty(g.ne, dtype=_EDGE_DT)
        edges["tail"] = g.tails.numpy()
        edges["weight"] = g.weights.numpy()
        edges.tofile(f)


def partition_ranges(nv: int, nparts: int) -> list[int]:
    """parts[p] = nv*p // P (distgraph.cpp:118-119, rebuild.cpp:294-295)."""
    return [(nv * p) // nparts for p in range(nparts + 1)]


def balanced_ranges(offsets, nparts: int) -> list[int]:
    """The reference's edge-balanced vertex ranges (option -b): balanceEdges, distgraph.cpp:22-66.

    The reference walks the first nv offsets once: bin p takes vertices until the entries it
    has SEEN reach ne // P — where vertex m is charged offsets[m] - offsets[m-1], i.e. the row of
    vertex m-1 (its loop reads offsets[m], not offsets[m+1]) — and the vertex that fills the bin
    already belongs to bin p+1 (distgraph.cpp:48-55).  Restated with one binary search per bin."""
    off = np.asarray(offsets.cpu() if hasattr(offsets, "cpu") else offsets, dtype=np.int64)
    nv = off.size - 1
    cap = int(off[nv]) // nparts
    seen = off[:nv]  # what the loop reads for m = 0 .. nv-1
    parts, start, charged = [0], 0, 0
    for _ in range(nparts - 1):
        # first m >= start with seen[m] - charged >= cap; vertex m opens the next bin
        m = start + int(np.searchsorted(seen[start:], charged + cap, side="left"))
        if m >= nv:
            parts.append(nv)
            start = nv
            continue
        parts.append(m)
        charged = int(seen[m])
        start = m + 1
    parts.append(nv)
    return parts


def read_vite(path: str, rank: int = 0, nparts: int = 1, balanced: bool = False) -> tuple[CSRGraph, int, int]:
    """Returns (local rows as a CSRGraph with global tails, base, global nv).  ``balanced``
    selects the reference's -b edge-balanced ranges (loadDistGraphMPIIOBalanced,
    distgraph.cpp:205-337) instead of parts[p] = nv*p // P."""
    with open(path, "rb") as f:
        nv, ne = (int(x) for x in np.fromfile(f, dtype="<i8", count=2))
        if balanced:
            parts = balanced_ranges(np.fromfile(f, dtype="<i8", count=nv + 1), nparts)
        else:
            parts = partition_ranges(nv, nparts)
        base, bound = parts[rank], parts[rank + 1]
        f.seek(16 + base * 8)
        offsets = np.fromfile(f, dtype="<i8", count=bound - base + 1)
        e0, e1 = int(offsets[0]), int(offsets[-1])
        f.seek(16 + (nv + 1) * 8 + e0 * _EDGE_DT.itemsize)
        edges = np.fromfile(f, dtype=_EDGE_DT, count=e1 - e0)
    g = CSRGraph(bound - base, torch.from_numpy(offsets - e0),
                 torch.from_numpy(np.ascontiguousarray(edges["tail"])),
                 torch.from_numpy(np.ascontiguousarray(edges["weight"])))
    return g, base, nv


def slice_rows(g: CSRGraph, base: int, bound: int) -> CSRGraph:
    """Rows [base, bound) of an in-memory graph, offsets rebased, tails global."""
    e0, e1 = int(g.offsets[base]), int(g.offsets[bound])
    return CSRGraph(bound - base, g.offsets[base:bound + 1] - e0, g.tails[e0:e1], g.weights[e0:e1])


def write_communities(path: str, comm) -> None:
    """The reference's community dump (main.cpp:521-541): one community id per line, vertex order.
    ``DistLouvain.run(reference_numbering=True)`` yields the ids main.cpp would write."""
    a = np.asarray(comm.cpu() if hasattr(comm, "cpu") else comm, dtype=np.int64)
    with open(path, "w") as f:
        f.write("\n".join(map(str, a.tolist())))
        if a.size:
            f.write("\n")
