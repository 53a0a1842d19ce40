"""CPU emulation of the reference's colouring option (-c) for the tests: distColoringMultiHashMinMax
(coloring.cpp:3-200) in numpy and distLouvainMethodWithColoring (louvain.cpp:756-949) on top of the
oracle's single-sweep function.  Test infrastructure only — pinned against the reference CLI in
tests/test_coloring.py (number of colours, level-0 modularity and iterations)."""
import numpy as np

M32 = 0xffffffff


def h(a, seed):
    a = (a ^ seed) & M32
    a = ((a + 0x7ed55d16) + (a << 12)) & M32
    a = ((a ^ 0xc761c23c) + (a >> 19)) & M32
    a = ((a + 0x165667b1) + (a << 5)) & M32
    a = ((a ^ 0xd3a2646c) + (a << 9)) & M32
    a = ((a + 0xfd7046c5) + (a << 3)) & M32
    a = ((a ^ 0xb55a4f09) + (a >> 16)) & M32
    return a
def hv(a, seed):
    a = a.astype(np.uint64)
    a = (a ^ np.uint64(seed)) & M32
    a = ((a + 0x7ed55d16) + (a << 12)) & M32
    a = ((a ^ 0xc761c23c) + (a >> 19)) & M32
    a = ((a + 0x165667b1) + (a << 5)) & M32
    a = ((a ^ 0xd3a2646c) + (a << 9)) & M32
    a = ((a + 0xfd7046c5) + (a << 3)) & M32
    a = ((a ^ 0xb55a4f09) + (a >> 16)) & M32
    return a
def colour(o, t, nhash, pct=70):
    nv = o.size - 1
    color = np.full(nv, -1, np.int64)
    seed, nxt, last = 1012, 0, 0
    src = np.repeat(np.arange(nv), np.diff(o))
    while True:
        unc = color == -1
        live = (t != src) & ~((color[t] != -1) & (color[t] < nxt)) & unc[src]
        not_min = np.zeros(nv, np.int64); not_max = np.zeros(nv, np.int64)
        for k in range(nhash):
            s = (seed + 1043 * k) & M32
            hvx = hv(np.arange(nv), s)
            a, b = hvx[src[live]], hvx[t[live]]
            nm = np.zeros(nv, bool); nx = np.zeros(nv, bool)
            nx[src[live][a <= b]] = True
            nm[src[live][a >= b]] = True
            not_min |= nm.astype(np.int64) << k; not_max |= nx.astype(np.int64) << k
        newc = color.copy()
        for v in np.nonzero(unc)[0]:
            poss = 2 * nhash - bin(not_min[v]).count('1') - bin(not_max[v]).count('1')
            if poss <= 0: continue
            cid, tid = v % poss, 0
            for k in range(nhash):
                if not (not_min[v] >> k) & 1 and cid == tid: newc[v] = 2 * k + nxt; break
                tid += not (not_min[v] >> k) & 1
                if not (not_max[v] >> k) & 1 and cid == tid: newc[v] = 2 * k + 1 + nxt; break
                tid += not (not_max[v] >> k) & 1
        color = newc
        cur = int((color != -1).sum())
        fin = cur >= (nv * pct) // 100 or cur == last
        nxt += 2 * nhash; seed = h(seed, 0); last = cur
        if fin: break
    return color, nxt
def phase_color(orc, o, t, w, color, numcolor, lower=-1.0, thresh=1e-6):
    nv = o.size - 1
    vdeg = orc.vertex_degree(o, w)
    const = 1.0 / vdeg.sum()
    past = np.arange(nv); curr = np.arange(nv); target = np.zeros(nv, np.int64)
    cdeg = vdeg.copy(); csize = np.ones(nv, np.int64)
    bucket = np.where(color < 0, numcolor - 1, color)
    prev, iters = lower, 0
    while True:
        iters += 1
        cw = np.zeros(nv)
        for ci in range(numcolor):
            sel = bucket == ci
            if not sel.any(): continue
            tg, cwv, _, _ = orc.sweep(nv, 0, nv, o, t, w, curr, vdeg, cdeg, csize, const)
            target[sel] = tg[sel]; cw[sel] = cwv[sel]
            mv = sel & (tg != curr)
            np.add.at(cdeg, tg[mv], vdeg[mv]); np.add.at(cdeg, curr[mv], -vdeg[mv])
            np.add.at(csize, tg[mv], 1); np.add.at(csize, curr[mv], -1)
        q, _ = orc.modularity(cw, cdeg, const)
        if q - prev < thresh: break
        prev = max(q, lower)
        past, curr, target = curr, target.copy(), past.copy()
    return prev, iters


def phase_color_et(orc, o, t, w, color, numcolor, et_type, delta, lower=-1.0, thresh=1e-6):
    """distLouvainMethodWithColoring with early termination: louvain.cpp:951-1178 (et_type 1, 3) and 1180-1426
    (et_type 2, 4), freeze rules as written there (incl. `target == curr == past` and the never-initialised p_curr)."""
    prob, cutoff = et_type in (2, 4), et_type in (3, 4)
    nv = o.size - 1
    vdeg = orc.vertex_degree(o, w)
    const = 1.0 / vdeg.sum()
    past = np.arange(nv)
    curr = np.arange(nv)
    target = np.zeros(nv, np.int64)
    cdeg = vdeg.copy()
    csize = np.ones(nv, np.int64)
    bucket = np.where(color < 0, numcolor - 1, color)
    active = np.ones(nv, bool)
    frozen_cw = np.zeros(nv)
    p_prev, p_curr = np.ones(nv), np.zeros(nv)
    prev, iters = lower, 0
    while True:
        iters += 1
        cw = np.zeros(nv)
        vc = 0
        for ci in range(numcolor):
            sel = bucket == ci
            if not sel.any():
                continue
            tg, cwv, _, _ = orc.sweep(nv, 0, nv, o, t, w, curr, vdeg, cdeg, csize, const)
            a, f = sel & active, sel & ~active
            cw[f] = frozen_cw[f]
            vc += int(f.sum())
            target[a] = tg[a]
            cw[a] = cwv[a]
            frozen_cw[a] = cwv[a]
            mv = a & (tg != curr)
            np.add.at(cdeg, tg[mv], vdeg[mv])
            np.add.at(cdeg, curr[mv], -vdeg[mv])
            np.add.at(csize, tg[mv], 1)
            np.add.at(csize, curr[mv], -1)
        if cutoff and vc >= 90:
            break
        q, _ = orc.modularity(cw, cdeg, const)
        if q - prev < thresh:
            break
        prev = max(q, lower)
        if not prob:
            freeze = (iters > 2) & ((target == curr).astype(np.int64) == past)
            active[freeze] = False
            rot = ~freeze
        else:
            hit = active & (iters > 2) & (curr == past)
            p_curr[hit] = p_prev[hit] * (1.0 - delta)
            active[hit & (p_curr <= 0.02)] = False
            rot = active.copy()
        tmp = past[rot].copy()
        past[rot] = curr[rot]
        curr[rot] = target[rot]
        target[rot] = tmp
        p_curr, p_prev = p_prev, p_curr
    return prev, iters
