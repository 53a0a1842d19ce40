"""Independent check of a community assignment: the true modularity of ``comm`` on a CSR graph.

    Q = sum_c [ e_cc / 2m  -  (a_c / 2m)^2 ],   2m = sum of all CSR entry weights,
    e_cc = weight of the entries whose two endpoints lie in c (a stored self-loop counts once,
    like the reference's vertex degree, louvain.cpp:2126-2151), a_c = sum of the degrees in c.

This is NOT the number the reference prints per level: that one mixes iteration k's intra-community
weights with iteration k+1's degrees (louvain.cpp:521-534).  It is what SURVEY.md
section 8(f) next-3 asks for as a check of the final assignment.  torch ops only (CPU or CUDA tensors);
not on the sweep hot path.
"""
from __future__ import annotations

import torch


def modularity(offsets: torch.Tensor, tails: torch.Tensor, weights: torch.Tensor, comm: torch.Tensor) -> float:
    offsets, tails, comm = offsets.long(), tails.long(), comm.long().to(offsets.device)
    w = weights.double()
    nv = offsets.numel() - 1
    src = torch.repeat_interleave(torch.arange(nv, device=offsets.device), offsets[1:] - offsets[:-1])
    two_m = w.sum()
    if float(two_m) == 0.0:
        return 0.0
    nc = int(comm.max().item()) + 1 if nv else 0
    cs, ct = comm[src], comm[tails]
    intra = torch.zeros(nc, dtype=torch.float64, device=w.device).index_add_(0, cs, torch.where(cs == ct, w, torch.zeros_like(w)))
    deg = torch.zeros(nc, dtype=torch.float64, device=w.device).index_add_(0, cs, w)
    return float((intra / two_m - (deg / two_m) ** 2).sum().item())
