"""CPU stand-in for cuvite_b200.engine.GpuGraph used by the gloo tests: the same
small interface DistLouvain drives, with the local sweep done by the oracle
(oracle/louvain_oracle.c orc_sweep on a vertex-range view).  TEST ONLY."""
from types import SimpleNamespace

import numpy as np
import torch


class OracleEngine:
    def __init__(self, oracle):
        self.o = oracle

    def upload(self, offsets, tails, weights, base=0, nv_global=None):
        self.off, self.tl, self.w = (np.ascontiguousarray(offsets, np.int64), np.ascontiguousarray(tails, np.int64),
                                     np.ascontiguousarray(weights, np.float64))
        self.nvl, self.base = self.off.size - 1, base
        self.nvg = self.nvl if nv_global is None else nv_global
        self.vdeg = self.o.vertex_degree(self.off, self.w)
        return self

    def info(self):
        return SimpleNamespace(nv_local=self.nvl, base=self.base, nv_global=self.nvg, ne_local=self.tl.size,
                               local_weight=float(self.vdeg.sum()))

    def phase_begin(self, total_weight=0.0):
        self.constant = 1.0 / (total_weight if total_weight > 0 else self.vdeg.sum())
        self.T = {"labels_global": torch.arange(self.nvg, dtype=torch.int64),
                  "cdeg": torch.zeros(self.nvg, dtype=torch.float64), "csize": torch.ones(self.nvg, dtype=torch.int64),
                  "ddeg": torch.zeros(self.nvg, dtype=torch.float64), "dsize": torch.zeros(self.nvg, dtype=torch.int64),
                  "vdeg": torch.from_numpy(self.vdeg)}
        self.T["cdeg"][self.base:self.base + self.nvl] = self.T["vdeg"]
        self.T["vdeg_global"] = torch.zeros(self.nvg, dtype=torch.float64)
        self.T["labels_global_next"] = torch.arange(self.nvg, dtype=torch.int64)
        self.labels_only = False
        ids = torch.arange(self.base, self.base + self.nvl, dtype=torch.int64)
        self.lab = [ids.clone(), ids.clone(), ids.clone()]
        self.past, self.curr, self.targ = 0, 1, 2
        self.cw = np.zeros(self.nvl)

    def tensor(self, name):
        if name == "labels_curr":
            return self.lab[self.curr]
        if name == "labels_target":
            return self.lab[self.targ]
        if name == "labels_past":
            return self.lab[self.past]
        if name == "labels_global" and self.nvl == self.nvg:
            return self.lab[self.curr]
        return self.T[name]

    def sweep(self):
        labels = self.tensor("labels_global").numpy()
        tgt, cw, dd, ds = self.o.sweep(self.nvl, self.base, self.nvg, self.off, self.tl, self.w, labels, self.vdeg,
                                       self.T["cdeg"].numpy(), self.T["csize"].numpy(), self.constant)
        self.lab[self.targ] = torch.from_numpy(tgt)
        self.cw = cw
        self.T["ddeg"] += torch.from_numpy(dd)
        self.T["dsize"] += torch.from_numpy(ds)

    def apply_and_reduce(self):
        self.T["cdeg"] += self.T["ddeg"]
        self.T["csize"] += self.T["dsize"]
        self.T["ddeg"].zero_()
        self.T["dsize"].zero_()
        own = self.T["cdeg"][self.base:self.base + self.nvl].numpy()
        return np.array([self.cw.sum(), (own * own).sum()])

    def rebuild_and_reduce(self):
        """cvb_rebuild_and_reduce: community info from the gathered next labels."""
        self.labels_only = True
        nxt = self.T["labels_global_next"]
        self.T["cdeg"].zero_().index_add_(0, nxt, self.T["vdeg_global"])
        self.T["csize"].zero_().index_add_(0, nxt, torch.ones(self.nvg, dtype=torch.int64))
        self.T["ddeg"].zero_()
        self.T["dsize"].zero_()
        own = self.T["cdeg"][self.base:self.base + self.nvl].numpy()
        return np.array([self.cw.sum(), (own * own).sum()])

    def rotate_labels(self):
        self.past, self.curr, self.targ = self.curr, self.targ, self.past
        if self.labels_only:
            self.T["labels_global"], self.T["labels_global_next"] = self.T["labels_global_next"], self.T["labels_global"]

    def last_sweep_ms(self):
        return 0.0

    # ---- N-rank coarsening stages (numpy stand-ins for cvb_renumber / cvb_contract_local /
    # ---- cvb_graph_from_pairs) -----------------------------------------------------------
    def renumber(self, labels):
        lab = labels.numpy()
        uniq, first = np.unique(lab, return_index=True)
        rank = np.empty(uniq.size, np.int64)
        rank[np.argsort(first, kind="stable")] = np.arange(uniq.size)
        return torch.from_numpy(rank[np.searchsorted(uniq, lab)]), int(uniq.size)

    @staticmethod
    def _merge(keys, w):
        uk, inv = np.unique(keys, return_inverse=True)
        return uk, np.bincount(inv, weights=w, minlength=uk.size)

    def contract_local(self, newid, n_comm):
        nid = newid.numpy()
        src = np.repeat(nid[self.base:self.base + self.nvl], np.diff(self.off))
        keys = (src.astype(np.int64) << 32) | nid[self.tl]
        uk, uw = self._merge(keys, self.w)
        return torch.from_numpy(uk), torch.from_numpy(uw)

    def graph_from_pairs(self, nv_local, base, nv_global, keys, weights):
        uk, uw = self._merge(keys.numpy(), weights.numpy())
        src, dst = (uk >> 32) - base, uk & 0xFFFFFFFF
        assert uk.size == 0 or (src.min() >= 0 and src.max() < nv_local)
        off = np.zeros(nv_local + 1, np.int64)
        np.cumsum(np.bincount(src, minlength=nv_local), out=off[1:])
        self.upload(off, dst, uw, base=base, nv_global=nv_global)
