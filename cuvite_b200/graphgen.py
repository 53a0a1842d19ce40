"""Deterministic synthetic graph generators (R-MAT, Erdős–Rényi, power-law).

The reference ships only a random-geometric-graph generator
(/root/reference/distgraph.cpp:341-933); BASELINE.json's configs need R-MAT,
Erdős–Rényi and power-law inputs, so these are new.  Every generator is
counter-based: random numbers are a pure function of (seed, edge index, draw
index) through a splitmix64 finaliser evaluated in wrapping int64 arithmetic,
and every real-valued threshold is an integer cut on a 32-bit draw, so the
graph is bit-identical on CPU and GPU and for any chunking.

Output is the Vite CSR layout the reference consumes (graph.hpp:35-36,
distgraph.cpp:99-190): ``offsets[nv+1]`` (int64), ``tails[ne]`` (int64 global
ids), ``weights[ne]`` (float64, 1.0).  Graphs are symmetrised, self-loops and
duplicate edges removed, and each adjacency list is sorted by tail — the order
rebuild.cpp:413-426 emits for coarsened graphs.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

_MASK32 = 0xFFFFFFFF
_GOLD = -7046029254386353131  # 0x9E3779B97F4A7C15 as int64
_M1 = -4658895280553007687  # 0xBF58476D1CE4E5B9
_M2 = -7723592293110705685  # 0x94D049BB133111EB


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    """Logical shift right on int64 (torch's >> is arithmetic)."""
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors with two's-complement wrap."""
    z = x + _GOLD
    z = (z ^ _lsr(z, 30)) * _M1
    z = (z ^ _lsr(z, 27)) * _M2
    return z ^ _lsr(z, 31)


def _splitmix64_int(x: int) -> int:
    """Scalar splitmix64 on Python ints; returns the value as a signed int64."""
    m = (1 << 64) - 1
    z = (x + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    z ^= z >> 31
    return z - (1 << 64) if z >= (1 << 63) else z


def _draws(seed: int, idx: torch.Tensor, draw: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Two independent 32-bit draws (as int64 in [0, 2^32)) per index."""
    key = _splitmix64_int((seed * 0x632BE5AB + draw * 0x2545F491) & 0x7FFFFFFFFFFFFFFF)
    h = splitmix64(idx ^ key)
    return _lsr(h, 32), h & _MASK32


@dataclass
class CSRGraph:
    """Undirected graph in the reference's CSR layout (both directions stored)."""

    nv: int
    offsets: torch.Tensor  # int64 [nv+1]
    tails: torch.Tensor  # int64 [ne]
    weights: torch.Tensor  # float64 [ne]

    @property
    def ne(self) -> int:
        return int(self.tails.numel())

    def cpu(self) -> "CSRGraph":
        return CSRGraph(self.nv, self.offsets.cpu(), self.tails.cpu(), self.weights.cpu())


def csr_from_pairs(nv: int, src: torch.Tensor, dst: torch.Tensor) -> CSRGraph:
    """Symmetrise, drop self-loops and duplicates, sort by (src, dst), build CSR."""
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = torch.cat([src * nv + dst, dst * nv + src])
    del src, dst
    key = torch.unique(key)  # sorted ascending => rows ascending, tails ascending inside a row
    rows = torch.div(key, nv, rounding_mode="floor")
    tails = key - rows * nv
    counts = torch.bincount(rows, minlength=nv)
    offsets = torch.zeros(nv + 1, dtype=torch.int64, device=key.device)
    torch.cumsum(counts, 0, out=offsets[1:])
    weights = torch.ones(tails.numel(), dtype=torch.float64, device=key.device)
    return CSRGraph(nv, offsets, tails, weights)


def rmat(scale: int, edge_factor: int = 16, seed: int = 2, a: float = 0.57, b: float = 0.19,
         c: float = 0.19, device: str | torch.device = "cpu", permute_seed: int | None = None,
         chunk: int = 1 << 24) -> CSRGraph:
    """Graph500-style R-MAT (a,b,c,d = 0.57,0.19,0.19,0.05 by default), nv = 2^scale."""
    nv = 1 << scale
    m = edge_factor * nv
    t1 = int(a * 4294967296.0)
    t2 = int((a + b) * 4294967296.0)
    t3 = int((a + b + c) * 4294967296.0)
    srcs, dsts = [], []
    for lo in range(0, m, chunk):
        idx = torch.arange(lo, min(lo + chunk, m), dtype=torch.int64, device=device)
        s = torch.zeros_like(idx)
        d = torch.zeros_like(idx)
        for lvl in range(0, scale, 2):
            u0, u1 = _draws(seed, idx, lvl // 2)
            for u in ((u0,) if lvl + 1 >= scale else (u0, u1)):
                sbit = (u >= t2).to(torch.int64)  # quadrants c,d -> lower half of the rows
                dbit = (((u >= t1) & (u < t2)) | (u >= t3)).to(torch.int64)  # quadrants b,d
                s = (s << 1) | sbit
                d = (d << 1) | dbit
        srcs.append(s)
        dsts.append(d)
    src, dst = torch.cat(srcs), torch.cat(dsts)
    del srcs, dsts
    if permute_seed is not None:
        ids = torch.arange(nv, dtype=torch.int64, device=device)
        perm = torch.argsort(splitmix64(ids ^ _splitmix64_int(permute_seed)), stable=True)
        src, dst = perm[src], perm[dst]
    return csr_from_pairs(nv, src, dst)


def erdos_renyi(nv: int, avg_degree: float = 16.0, seed: int = 1,
                device: str | torch.device = "cpu") -> CSRGraph:
    """G(n, M) with M = nv*avg_degree/2 sampled pairs (duplicates/self-loops dropped)."""
    m = int(nv * avg_degree / 2)
    idx = torch.arange(m, dtype=torch.int64, device=device)
    u0, u1 = _draws(seed, idx, 0)
    v0, v1 = _draws(seed, idx, 1)
    src = ((u0 << 20) ^ v0) % nv
    dst = ((u1 << 20) ^ v1) % nv
    return csr_from_pairs(nv, src, dst)


def power_law(nv: int, num_edges: int, alpha: float = 1.8, max_degree: int = 1 << 20,
              seed: int = 4, device: str | torch.device = "cpu", chunk: int = 1 << 24) -> CSRGraph:
    """Chung–Lu graph whose expected degrees follow P(d) ∝ d^-alpha on [1, max_degree].

    Vertex v's weight is the inverse-CDF of the truncated power law at the
    (deterministic) quantile (v + 0.5)/nv, so low ids are the hubs; both
    endpoints of each of the ``num_edges`` sampled edges are drawn ∝ weight
    through an integer-quantised cumulative table (exactly reproducible).
    """
    # the weight table is built on the CPU (sequential float64 cumsum) so that
    # it does not depend on the device's reduction order
    q = (torch.arange(nv, dtype=torch.float64) + 0.5) / nv
    e = 1.0 - alpha
    lo, hi = 1.0, float(max_degree)
    # inverse CDF, descending in v (q -> 1-q) so hubs come first
    w = ((1.0 - q) * (hi ** e - lo ** e) + lo ** e) ** (1.0 / e)
    cum = torch.cumsum(w, 0)
    cum_i = torch.clamp((cum / cum[-1] * float(1 << 52)).to(torch.int64), max=(1 << 52) - 1)
    cum_i[-1] = (1 << 52) - 1
    cum_i = cum_i.to(device)
    srcs, dsts = [], []
    for lo_i in range(0, num_edges, chunk):
        idx = torch.arange(lo_i, min(lo_i + chunk, num_edges), dtype=torch.int64, device=device)
        u0, u1 = _draws(seed, idx, 0)
        v0, v1 = _draws(seed, idx, 1)
        rs = ((u0 << 20) | (v0 >> 12)) & ((1 << 52) - 1)
        rd = ((u1 << 20) | (v1 >> 12)) & ((1 << 52) - 1)
        srcs.append(torch.clamp(torch.searchsorted(cum_i, rs), max=nv - 1))
        dsts.append(torch.clamp(torch.searchsorted(cum_i, rd), max=nv - 1))
    return csr_from_pairs(nv, torch.cat(srcs), torch.cat(dsts))


def ring_of_cliques(num_cliques: int, clique_size: int, device: str | torch.device = "cpu") -> CSRGraph:
    """Tiny structured graph with a known optimum, for unit tests."""
    nv = num_cliques * clique_size
    base = torch.arange(num_cliques, dtype=torch.int64, device=device) * clique_size
    a, b = torch.meshgrid(torch.arange(clique_size, device=device), torch.arange(clique_size, device=device),
                          indexing="ij")
    src = (base[:, None, None] + a[None]).reshape(-1)
    dst = (base[:, None, None] + b[None]).reshape(-1)
    bridge_s = base + clique_size - 1
    bridge_d = (base + clique_size) % nv
    return csr_from_pairs(nv, torch.cat([src, bridge_s]), torch.cat([dst, bridge_d]))
