"""Golden phase-0 results of the bench graphs from the CPU checker (the reference's louvain.cpp when
oracle/_ref is built, else the pinned C restatement): tests/golden/bench_q.json.
bench.py compares the GPU phase against these (and against the live cpu_baseline leg when it runs)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuvite_b200 import graphgen as gg  # noqa: E402
from oracle.binding import Oracle, Reference  # noqa: E402

out_path = os.path.join(ROOT, "tests", "golden", "bench_q.json")
res = json.load(open(out_path)) if os.path.exists(out_path) else {}
cases = [(int(a.split(":")[0]), a.split(":")[1] == "p") for a in sys.argv[1:]] or [(20, False), (21, True)]
for scale, perm in cases:
    key = f"rmat{scale}{'p' if perm else ''}"
    if key in res:
        continue
    t0 = time.time()
    g = gg.rmat(scale, seed=2, permute_seed=3 if perm else None)
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy()
    if Reference.available():
        R = Reference()
        h = R.graph(o, t, w)
        p = R.phase_on(h)
        R.destroy(h)
        kind = "reference"
    else:
        p = Oracle().phase(o, t, w)
        kind = "port"
    res[key] = {"scale": scale, "permuted": perm, "nv": g.nv, "csr_entries": g.ne, "q": p.q, "iters": p.iters,
                "checker": kind, "seconds": round(time.time() - t0, 1)}
    print(key, res[key], flush=True)
    json.dump(res, open(out_path, "w"), indent=1, sort_keys=True)
