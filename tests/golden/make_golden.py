"""Generate tests/golden/*.npz from the REAL reference (oracle/_ref/libcuvite_ref.so).

Run in the build container (where /root/reference exists and oracle/build_ref.sh
has been run):   python tests/golden/make_golden.py
Each fixture holds the input CSR and what pnnl/cuVite's own louvain.cpp /
rebuild.cpp / main.cpp-loop produce for it: per-iteration targetComm,
clusterWeight, modularity and community info of phase 0, the phase result, the
coarsened graph, and the per-level results of the full run.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from graphs import small_cases  # noqa: E402
from oracle.binding import Reference  # noqa: E402


def main():
    R = Reference()
    R.set_num_threads(1)
    out = os.path.dirname(os.path.abspath(__file__))
    for name, (o, t, w) in small_cases().items():
        p = R.phase(o, t, w, trace=32)
        p2 = R.phase(o, t, w)
        assert p.iters == p2.iters and p.q == p2.q and (p.cvect == p2.cvect).all()
        no, nt, nw = R.rebuild(o, t, w, p.cvect)
        run = R.run(o, t, w)
        np.savez_compressed(
            os.path.join(out, f"{name}.npz"), offsets=o, tails=t.astype(np.int32), weights=w,
            targets=p.targets.astype(np.int32), cw=p.cw, qs=p.qs, csize=p.csize.astype(np.int32), cdeg=p.cdeg,
            cvect=p.cvect.astype(np.int32), q=p.q, iters=p.iters, co_offsets=no, co_tails=nt.astype(np.int32),
            co_weights=nw, level_q=np.array(run.level_q), level_iters=np.array(run.level_iters),
            level_nv=np.array(run.level_nv), comm=run.comm.astype(np.int32))
        print(f"{name}: nv={o.size - 1} ne={t.size} iters={p.iters} q={p.q:.12f} levels={len(run.level_q)}")


def balance_fixture():
    """-b vertex ranges of the reference (balanceEdges, distgraph.cpp:22-66) for a few graphs."""
    import tempfile
    from cuvite_b200 import graphgen as gg, vite_io
    R = Reference()
    out = os.path.dirname(os.path.abspath(__file__))
    cases = {"rmat12": gg.rmat(12, seed=2), "er3k": gg.erdos_renyi(3000, 8.0, seed=1),
             "pl5k": gg.power_law(5000, 40000, seed=4), "ring6x5": gg.ring_of_cliques(6, 5)}
    data = {}
    with tempfile.TemporaryDirectory() as d:
        for name, g in cases.items():
            f = os.path.join(d, name + ".bin")
            vite_io.write_vite(f, g)
            data[name + "_offsets"] = g.offsets.numpy()
            for P in (1, 2, 3, 4, 7, 8, 16, 64):
                data[f"{name}_P{P}"] = R.balance_edges(f, P)
    np.savez_compressed(os.path.join(out, "balance.npz"), **data)
    print("balance.npz:", len(data), "arrays")


if __name__ == "__main__":
    balance_fixture()
    main()
