"""Option -c (distance-1 colouring + coloured phase 0): the emulation against the reference CLI (CPU),
the GPU path against both."""
import os
import re
import subprocess

import numpy as np
import pytest

import color_emulation as ce

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "gc_cpu64")


def _graph(name):
    from cuvite_b200 import graphgen as gg
    return gg.erdos_renyi(3000, 8.0, seed=1) if name == "er3000" else gg.rmat(13, seed=2)


def _cli(tmp_path, g, max_colors):
    from cuvite_b200 import vite_io
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gc_cpu64 not built (needs /root/reference at build time)")
    f = str(tmp_path / "g.bin")
    vite_io.write_vite(f, g)
    r = subprocess.run([EXE, "-f", f, "-c", str(max_colors)], capture_output=True, text=True, timeout=600,
                       cwd=str(tmp_path), env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ncol = int(re.search(r"Number of colors \(2\*nHash\): (\d+)", r.stdout).group(1))
    levels = re.findall(r"Level \d+, Modularity: ([0-9.eE+-]+), Clustering time: [^,]*, Iterations: (\d+)", r.stdout)
    assert levels
    return ncol, [(float(q), int(i)) for q, i in levels]


@pytest.mark.parametrize("graph", ["er3000", "rmat13"])
@pytest.mark.parametrize("max_colors", [8, 4, 16, 3, 32])
def test_emulation_matches_reference_cli(tmp_path, graph, max_colors):
    """Pins the emulation: number of colours, level-0 modularity (6 digits, as printed) and iterations of
    the reference's own `-c` run."""
    from oracle.binding import Oracle
    g = _graph(graph)
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy().astype(np.float64)
    ncol, levels = _cli(tmp_path, g, max_colors)
    color, n = ce.colour(o, t, max_colors // 2)
    assert n == ncol
    q, it = ce.phase_color(Oracle(), o, t, w, color, n + 1)
    assert f"{q:.6g}" == f"{levels[0][0]:.6g}" and it == levels[0][1]


@pytest.mark.gpu
@pytest.mark.parametrize("graph", ["er3000", "rmat13"])
@pytest.mark.parametrize("max_colors", [8, 4, 2, 1, 16, 3, 32])
def test_gpu_colouring_and_coloured_phase_match_emulation(graph, max_colors):
    """Colours bit for bit, number of colours, and the coloured phase (modularity EXACTLY — unit weights make
    every sum an integer — and iterations) against the CPU emulation."""
    from cuvite_b200.engine import GpuGraph
    from oracle.binding import Oracle
    g = _graph(graph)
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy().astype(np.float64)
    color, n = ce.colour(o, t, max_colors // 2)
    orc = Oracle()
    if n == 0:  # no hash function: the reference falls back to the plain phase (louvain.cpp:764-768)
        want = orc.phase(o, t, w)
        wq, wit = want.q, want.iters
    else:
        wq, wit = ce.phase_color(orc, o, t, w, color, n + 1)
    gpu = GpuGraph()
    try:
        gpu.upload(o, t, w)
        q, it, ncol = gpu.phase_run_color(max_colors)
        assert ncol == n
        assert np.array_equal(gpu.colors(), color)
        assert (q, it) == (wq, wit)
    finally:
        gpu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("graph", ["er3000", "rmat13"])
@pytest.mark.parametrize("max_colors", [8, 4])
def test_gpu_full_run_with_colouring_matches_reference_cli(tmp_path, graph, max_colors):
    """cvb_louvain_run_color == the reference CLI with -c: every level's modularity (as printed) and
    cumulative iterations; only phase 0 is coloured (main.cpp:242)."""
    from cuvite_b200.engine import GpuGraph
    g = _graph(graph)
    _, levels = _cli(tmp_path, g, max_colors)
    gpu = GpuGraph()
    try:
        gpu.upload(g.offsets.numpy(), g.tails.numpy(), g.weights.numpy())
        lq, li, _, _ = gpu.run(max_colors=max_colors)
    finally:
        gpu.close()
    assert [f"{q:.6g}" for q in lq] == [f"{q:.6g}" for q, _ in levels]
    assert li == [i for _, i in levels]


@pytest.mark.gpu
@pytest.mark.parametrize("real_weights", [False, True])
def test_gpu_coloured_phase_on_hub_graph(real_weights):
    """Hubs in every degree class, including the partition-then-hash path above 8192 entries (evaluated for all
    big vertices in every colour step, kept for the step's own colour) and, with real weights, the F64 kernels
    over the per-colour lists."""
    from graphs import hub_graph, with_real_weights
    from cuvite_b200.engine import GpuGraph
    from oracle.binding import Oracle
    o, t, w = hub_graph()
    if real_weights:
        o, t, w = with_real_weights(o, t, w)
    w = np.asarray(w, np.float64)
    color, n = ce.colour(np.asarray(o), np.asarray(t), 4)
    wq, wit = ce.phase_color(Oracle(), np.asarray(o), np.asarray(t), w, color, n + 1)
    gpu = GpuGraph()
    try:
        gpu.upload(o, t, w)
        q, it, ncol = gpu.phase_run_color(8)
        assert ncol == n and np.array_equal(gpu.colors(), color)
        assert it == wit
        assert (q == wq) if not real_weights else abs(q - wq) <= 1e-6
    finally:
        gpu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ring6x5", "er1k", "rmat10", "rmat10_intw", "rmat9_realw", "hub30k", "empty_rows", "multi9"])
@pytest.mark.parametrize("max_colors", [8, 2])
def test_gpu_coloured_phase_on_golden_inputs(name, max_colors):
    """The edge cases of the golden set under -c: empty rows, repeated tails (multi9), integer weights other
    than 1 (rmat10_intw), real weights (rmat9_realw, tolerance 1e-6), a 30 k-entry hub."""
    from conftest import load_golden
    from cuvite_b200.engine import GpuGraph
    from oracle.binding import Oracle
    g = load_golden(name)
    o, t, w = g["offsets"], g["tails"].astype(np.int64), np.asarray(g["weights"], np.float64)
    color, n = ce.colour(o, t, max_colors // 2)
    wq, wit = ce.phase_color(Oracle(), o, t, w, color, n + 1)
    gpu = GpuGraph()
    try:
        gpu.upload(o, t, w)
        q, it, ncol = gpu.phase_run_color(max_colors)
        assert ncol == n and np.array_equal(gpu.colors(), color)
        assert it == wit
        assert (q == wq) if name != "rmat9_realw" else abs(q - wq) <= 1e-6
    finally:
        gpu.close()


# ---- -c together with -t / -a (louvain.cpp:951-1426) ---------------------------------------------------------
ET_CASES = [(1, 1.0), (3, 1.0), (2, 0.9), (4, 0.9), (2, 0.25)]


def _cli_et(tmp_path, g, max_colors, et_type, delta):
    """The reference CLI with -c and -t [-a], ONE host thread (its vector<bool> freeze mask races otherwise)."""
    from cuvite_b200 import vite_io
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gc_cpu64 not built (needs /root/reference at build time)")
    f = str(tmp_path / "g.bin")
    vite_io.write_vite(f, g)
    opts = ["-c", str(max_colors), "-t", str(et_type)] + (["-a", str(delta)] if et_type in (2, 4) else [])
    r = subprocess.run([EXE, "-f", f] + opts, capture_output=True, text=True, timeout=600, cwd=str(tmp_path),
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    levels = re.findall(r"Level \d+, Modularity: ([0-9.eE+-]+), Clustering time: [^,]*, Iterations: (\d+)", r.stdout)
    assert levels
    return [(float(q), int(i)) for q, i in levels]


@pytest.mark.parametrize("graph", ["er3000", "rmat13"])
@pytest.mark.parametrize("et_type,delta", ET_CASES)
def test_emulation_with_early_termination_matches_reference_cli(tmp_path, graph, et_type, delta):
    from oracle.binding import Oracle
    g = _graph(graph)
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy().astype(np.float64)
    levels = _cli_et(tmp_path, g, 8, et_type, delta)
    color, n = ce.colour(o, t, 4)
    q, it = ce.phase_color_et(Oracle(), o, t, w, color, n + 1, et_type, delta)
    assert f"{q:.6g}" == f"{levels[0][0]:.6g}" and it == levels[0][1]


@pytest.mark.gpu
@pytest.mark.parametrize("graph", ["er3000", "rmat13", "hub"])
@pytest.mark.parametrize("et_type,delta", ET_CASES)
def test_gpu_coloured_phase_with_early_termination_matches_emulation(graph, et_type, delta):
    """cvb_phase_run_color_et against the emulation: modularity exactly, iterations — incl. the hub graph, whose
    vertices above 8192 entries go through k_color_fix_big's frozen / active handling."""
    from cuvite_b200.engine import GpuGraph
    from oracle.binding import Oracle
    if graph == "hub":
        from graphs import hub_graph
        o, t, w = (np.asarray(x) for x in hub_graph())
        w = w.astype(np.float64)
    else:
        g = _graph(graph)
        o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy().astype(np.float64)
    color, n = ce.colour(o, t, 4)
    wq, wit = ce.phase_color_et(Oracle(), o, t, w, color, n + 1, et_type, delta)
    gpu = GpuGraph()
    try:
        gpu.upload(o, t, w)
        q, it, ncol = gpu.phase_run_color(8, et_type=et_type, et_delta=delta)
        assert ncol == n and (q, it) == (wq, wit)
    finally:
        gpu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("graph", ["er3000", "rmat13"])
@pytest.mark.parametrize("et_type,delta", ET_CASES)
def test_gpu_full_run_with_colouring_and_early_termination_matches_reference_cli(tmp_path, graph, et_type, delta):
    """cvb_louvain_run_color_et == the reference CLI with -c 8 -t k [-a d]: phase 0 coloured with early termination,
    later phases with early termination only (main.cpp:242-341)."""
    from cuvite_b200.engine import GpuGraph
    g = _graph(graph)
    levels = _cli_et(tmp_path, g, 8, et_type, delta)
    gpu = GpuGraph()
    try:
        gpu.upload(g.offsets.numpy(), g.tails.numpy(), g.weights.numpy())
        lq, li, _, _ = gpu.run(max_colors=8, et_type=et_type, et_delta=delta)
    finally:
        gpu.close()
    assert [f"{q:.6g}" for q in lq] == [f"{q:.6g}" for q, _ in levels]
    assert li == [i for _, i in levels]


# ---- option -d (vertex ordering) -----------------------------------------------------------------------------
def _cli_plain(tmp_path, g, opts):
    from cuvite_b200 import vite_io
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gc_cpu64 not built (needs /root/reference at build time)")
    f = str(tmp_path / "g.bin")
    vite_io.write_vite(f, g)
    r = subprocess.run([EXE, "-f", f] + opts, capture_output=True, text=True, timeout=600, cwd=str(tmp_path),
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ncol = re.search(r"Number of colors \(2\*nHash\): (\d+)", r.stdout)
    levels = re.findall(r"Level \d+, Modularity: ([0-9.eE+-]+), Clustering time: [^,]*, Iterations: (\d+)", r.stdout)
    return (int(ncol.group(1)) if ncol else None), [(q, int(i)) for q, i in levels]


@pytest.mark.parametrize("graph", ["er3000", "rmat13"])
def test_reference_vertex_ordering_equals_plain_run(tmp_path, graph):
    """Pins the claim in include/cuvite_b200.h: on one rank `-d 8` prints the level lines of the plain run (the
    colour-ordered visit applies the community updates once per iteration), plus the number of colours."""
    g = _graph(graph)
    ncol, with_d = _cli_plain(tmp_path, g, ["-d", "8"])
    _, plain = _cli_plain(tmp_path, g, [])
    assert with_d == plain and ncol is not None


@pytest.mark.gpu
@pytest.mark.parametrize("graph", ["er3000", "rmat13"])
def test_gpu_vertex_ordering_is_colouring_plus_plain_run(tmp_path, graph):
    """Option -d on the GPU: cvb_color_graph gives the CLI's number of colours (and the emulation's colours), the
    run itself is cvb_louvain_run — equal to the CLI's `-d 8` level lines."""
    from cuvite_b200.engine import GpuGraph
    g = _graph(graph)
    ncol, levels = _cli_plain(tmp_path, g, ["-d", "8"])
    color, n = ce.colour(g.offsets.numpy(), g.tails.numpy(), 4)
    gpu = GpuGraph()
    try:
        gpu.upload(g.offsets.numpy(), g.tails.numpy(), g.weights.numpy())
        assert gpu.color_graph(8) == ncol == n
        assert np.array_equal(gpu.colors(), color)
        lq, li, _, _ = gpu.run()
    finally:
        gpu.close()
    assert [f"{q:.6g}" for q in lq] == [f"{float(q):.6g}" for q, _ in levels]
    assert li == [i for _, i in levels]
