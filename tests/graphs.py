"""Seeded test graphs shared by the golden-fixture generator and the tests."""
from __future__ import annotations

import numpy as np
import torch

from cuvite_b200 import graphgen as gg


def _np(g):
    return g.offsets.numpy().copy(), g.tails.numpy().copy(), g.weights.numpy().copy()


def _sym_hash(o, t):
    """A symmetric per-edge hash in [0, 2^32): same value for (i,j) and (j,i)."""
    src = np.repeat(np.arange(o.size - 1, dtype=np.int64), np.diff(o))
    lo, hi = np.minimum(src, t), np.maximum(src, t)
    h = gg.splitmix64(torch.from_numpy(lo * 1000003 + hi)).numpy()
    return (h >> 32) & 0xFFFFFFFF


def with_int_weights(o, t, w, maxw=5):
    return o, t, (1 + _sym_hash(o, t) % maxw).astype(np.float64)


def with_real_weights(o, t, w):
    return o, t, 0.1 + _sym_hash(o, t).astype(np.float64) / 4294967296.0


def hub_graph(nv=30000, seed=7):
    """Background ER graph plus hubs hitting every degree class (giant, 16K, 4K, warp)."""
    bg = gg.erdos_renyi(nv, 6.0, seed=seed)
    src = np.repeat(np.arange(nv, dtype=np.int64), np.diff(bg.offsets.numpy()))
    dst = bg.tails.numpy()
    extra_s, extra_d = [], []
    for hub, deg in ((0, nv - 1), (1, 6000), (2, 1500), (3, 400), (4, 100), (5, 40)):
        others = (np.arange(deg, dtype=np.int64) * 7919 + 11 * hub + 13) % nv
        extra_s.append(np.full(deg, hub, np.int64))
        extra_d.append(others)
    s = torch.from_numpy(np.concatenate([src] + extra_s))
    d = torch.from_numpy(np.concatenate([dst] + extra_d))
    return _np(gg.csr_from_pairs(nv, s, d))


def multigraph(scale=9, seed=6):
    """A NON-simple Vite-style input: parallel entries (both directions duplicated), self-loops (some
    stored twice), integer weights 1..3 and rows in shuffled (not ascending) order."""
    o, t, _ = _np(gg.rmat(scale, seed=seed))
    nv = o.size - 1
    src = np.repeat(np.arange(nv, dtype=np.int64), np.diff(o))
    h = _sym_hash(o, t)
    w = (1 + h % 3).astype(np.float64)
    dup = (h % 5) == 0  # symmetric predicate: (i,j) and (j,i) are duplicated together
    loops = np.arange(0, nv, 7, dtype=np.int64)
    loops2 = np.arange(0, nv, 21, dtype=np.int64)
    s = np.concatenate([src, src[dup], loops, loops2])
    d = np.concatenate([t, t[dup], loops, loops2])
    ww = np.concatenate([w, w[dup], np.full(loops.size, 2.0), np.full(loops2.size, 1.0)])
    rng = np.random.default_rng(seed)
    order = np.lexsort((rng.random(s.size), s))  # rows contiguous, entries of a row shuffled
    s, d, ww = s[order], d[order], ww[order]
    off = np.zeros(nv + 1, np.int64)
    np.cumsum(np.bincount(s, minlength=nv), out=off[1:])
    return off, d, ww


def small_cases():
    """name -> (offsets, tails, weights); all small enough for committed fixtures."""
    cases = {
        "ring6x5": _np(gg.ring_of_cliques(6, 5)),
        "er1k": _np(gg.erdos_renyi(1024, 8.0, seed=1)),
        "rmat10": _np(gg.rmat(10, seed=2)),
        "rmat10_intw": with_int_weights(*_np(gg.rmat(10, seed=2))),
        "rmat9_realw": with_real_weights(*_np(gg.rmat(9, seed=5))),
        "hub30k": hub_graph(),
        "empty_rows": _np(gg.csr_from_pairs(64, torch.tensor([1, 5, 9, 9]), torch.tensor([2, 6, 10, 11]))),
        "multi9": multigraph(),
    }
    return cases
