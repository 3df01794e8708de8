"""CPU-side tests: the C-ABI library loads and exports every declared symbol, host logic (codec names, synthetic
image generator, sharding plan), and the oracle pipelines against each other.  No GPU compute."""
import ctypes
import os

import numpy as np
import pytest

import oracle as O
from cniic_b200 import _lib, codecs, dist
import cniic_b200 as cb


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.SO_PATH)
    names = _lib.declared_symbols()
    assert len(names) >= 40
    missing = [s for s in names if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CniicError) as e:
        cb.Context()
    assert e.value.code == cb.ERR_CUDA


def test_product_never_imports_oracle():
    root = os.path.dirname(_lib.__file__)
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f)).read()
                assert "import oracle" not in text and "liboracle" not in text and "cniic_oracle" not in text, f
                # ... nor the host emulation of tests/emu (test infrastructure), nor any environment switch to another library
                assert "libcniic_emu" not in text and "tests/emu" not in text and "CNIIC_EMU" not in text, f
    assert _lib.lib()._name == _lib.SO_PATH  # the loader serves exactly the in-tree CUDA build


@pytest.mark.parametrize("expr,name", [
    ("cluster-colors(256)", "cluster-colors_256"), ("ccol(16)", "cluster-colors_16"), ("ccolors(7)", "cluster-colors_7"),
    ("cluster-col(9)", "cluster-colors_9"), ("voronoi(2048)", "voronoi_2048"), ("delta", "delta"), ("hufman", "Hufman"),
    ("HUFMAN", "Hufman"), ("hilbert(rle)", "hilbert-rle")])
def test_codec_names(expr, name):  # clusterc.rs:59-61,116-141,191-193,274-297 ; hilbertc.rs:81-87,433-435 ; hufc.rs:42-62
    assert codecs.codec_name(expr) == name


@pytest.mark.parametrize("expr", ["", "voronoi", "voronoi()", "cluster-colors(x)", "zip(dict)", "Delta", "hilbert(zip)"])
def test_codec_names_rejected(expr):
    with pytest.raises(ValueError):
        codecs.codec_name(expr)


def test_synth_image_is_deterministic_and_shardable():
    a = cb.synth_image_host(96, 64, 42, 12)
    b = cb.synth_image_host(96, 64, 42, 12)
    assert np.array_equal(a, b)
    top = cb.synth_image_host(96, 40, 42, 12, y0=0, h_total=64)
    bot = cb.synth_image_host(96, 24, 42, 12, y0=40, h_total=64)
    assert np.array_equal(np.concatenate([top, bot]), a)
    assert len(np.unique(a.reshape(-1, 3), axis=0)) >= 256  # enough distinct colours for every k used in the tests
    assert not np.array_equal(a, cb.synth_image_host(96, 64, 43, 12))


def test_row_shard_covers_image():
    for h in (1, 7, 540, 4320):
        for world in (1, 2, 3, 8):
            rows = [dist.row_shard(h, world, r) for r in range(world)]
            assert rows[0][0] == 0 and sum(n for _, n in rows) == h
            for (y0, n), (y1, _) in zip(rows, rows[1:]):
                assert y0 + n == y1


def test_init_point_indices_match_oracle():
    img = cb.synth_image_host(40, 30, 7, 6)
    for k in (1, 3, 16, 37):
        idx = dist.init_point_indices(40 * 30, k)
        o = O.kmeans_xyrgb(img, k, max_iters=1, tie=O.TIE_KEEP_CURRENT)
        # after the init the oracle's first pass uses exactly these points as centroids; reproduce the first pass
        flat = img.reshape(-1, 3).astype(np.int64)
        pts = np.concatenate([(np.arange(1200) % 40)[:, None], (np.arange(1200) // 40)[:, None], flat], axis=1)
        cen = pts[idx]
        d2 = ((pts[:, None, :] - cen[None, :, :]) ** 2).sum(-1)
        # keep-current tie rule: current = chunk assignment
        ppc = 1200 // k
        cur = np.where(np.arange(1200) >= 1200 - (k - 1) * ppc, (1199 - np.arange(1200)) // ppc, k - 1)
        best = d2.argmin(1)
        keep = d2[np.arange(1200), cur] == d2.min(1)
        best = np.where(keep, cur, best)
        assert np.array_equal(best, o.assign)


def test_local_init_contributions_sum_to_global():
    img = cb.synth_image_host(33, 21, 3, 6)
    n, k = 33 * 21, 16
    full = dist.local_init_contribution(5, img, 33, 0, n, k)
    parts = np.zeros_like(full)
    for r in range(3):
        y0, hl = dist.row_shard(21, 3, r)
        parts += dist.local_init_contribution(5, img[y0:y0 + hl], 33, y0, n, k)
    assert np.array_equal(parts, full)
    idx = dist.init_point_indices(n, k)
    assert np.array_equal(full[:, 0], idx % 33) and np.array_equal(full[:, 1], idx // 33)


# ---- oracle self-consistency (exercises the checker the GPU tests rely on) ----

def test_oracle_hilbert_invariants():  # hilbert.rs:40-43 ; SURVEY 8c invariants (curve parity is UNPINNED)
    for (w, h) in [(1, 1), (1, 7), (7, 1), (4, 4), (8, 8), (5, 3), (3, 5), (13, 29), (64, 48), (100, 7)]:
        xy = O.hilbert_xy(w, h)
        assert len(xy) == w * h
        lin = xy[:, 1].astype(np.int64) * w + xy[:, 0]
        assert len(np.unique(lin)) == w * h and xy[:, 0].max() < w and xy[:, 1].max() < h  # bijection
        assert tuple(xy[0]) == (0, 0)
        step = np.abs(np.diff(xy.astype(np.int64), axis=0)).max(axis=1)
        assert step.max(initial=0) <= 1  # consecutive cells are 8-adjacent
    # README.md:87-106 4x4 diagram (y drawn upward)
    assert O.hilbert_xy(4, 4).tolist() == [[0, 0], [1, 0], [1, 1], [0, 1], [0, 2], [0, 3], [1, 3], [1, 2], [2, 2], [2, 3],
                                          [3, 3], [3, 2], [3, 1], [2, 1], [2, 0], [3, 0]]


def test_oracle_delta_roundtrip_and_codecs():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(9, 14, 3), dtype=np.uint8)
    d = O.delta(img)
    assert d[0].tolist() == O.hilbert_gather(img)[0].tolist()  # first diff is the first colour (hilbertc.rs:441-445)
    assert np.array_equal(O.undelta(d, 14, 9), img)
    for enc, dec in [(O.encode_delta, O.decode_delta), (O.encode_hufman, O.decode_hufman),
                     (O.encode_hilbert_rle, O.decode_hilbert_rle)]:
        assert np.array_equal(dec(enc(img)), img)  # bench.rs:57-59 lossless check
    keys, cnts = O.hist_delta(d)
    assert int(cnts.sum()) == 9 * 14 and np.all(np.diff(keys.astype(np.int64)) > 0)


def test_oracle_rle_run_cap():  # hilbertc.rs:23,130 : runs are capped at 255
    s = np.zeros((600, 3), np.uint8)
    s[599] = 1
    cnt, col = O.rle_exact(s)
    assert cnt.tolist() == [255, 255, 89, 1] and col[-1].tolist() == [1, 1, 1]


def test_oracle_voronoi_codec():  # clusterc.rs:148-189 ; stream = 16 + 19k bytes
    img = cb.synth_image_host(48, 32, 11, 6)
    data = O.encode_voronoi(img, 8)
    assert len(data) == 16 + 19 * 8
    dec = O.decode_voronoi(data)
    assert dec.shape == img.shape
    o = O.kmeans_xyrgb(img, 8)
    assert np.array_equal(dec, O.voronoi_fill(o.centroids[:, :2], o.centroids[:, 2:], 48, 32))
    assert O.mse(img, dec) == pytest.approx(O.sse(img, dec) / (48 * 32), rel=1e-12)  # bench.rs:95-104 vs exact integer SSE


def test_oracle_verbatim_vs_exact_divergence_is_small():
    """the reference's truncated neighbour lists make later iterations approximate; iteration 1 is exact."""
    img = cb.synth_image_host(64, 48, 5, 12)
    v = O.kmeans_xyrgb(img, 16, mode=O.MODE_VERBATIM, max_iters=1)
    e = O.kmeans_xyrgb(img, 16, mode=O.MODE_EXACT, max_iters=1)
    # distinct integer distances are never mis-ordered by the f64 path; only exact integer ties may differ (F4)
    diff = v.assign != e.assign
    flat = img.reshape(-1, 3).astype(np.int64)
    pts = np.concatenate([(np.arange(64 * 48) % 64)[:, None], (np.arange(64 * 48) // 64)[:, None], flat], axis=1)
    idx = dist.init_point_indices(64 * 48, 16)
    d2 = ((pts[:, None, :] - pts[idx][None]) ** 2).sum(-1)
    rows = np.nonzero(diff)[0]
    assert np.all(d2[rows, v.assign[rows]] == d2[rows, e.assign[rows]])
    full = O.kmeans_xyrgb(img, 16, mode=O.MODE_VERBATIM)
    assert full.dist_evals < O.kmeans_xyrgb(img, 16, mode=O.MODE_EXACT, max_iters=full.iterations).dist_evals


def test_bench_roofline_object():
    """bench.py's roofline builder on made-up numbers: schema of the judged keys and the arithmetic."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    pk = {"hbm_gbs": 6471.1, "sm_max_mhz": 1965.0, "source": "measured"}
    n, k = 4096 * 4096, 256
    r = bench.build_roofline(3, n, k, 0.1, 0.036 * n * k, pk, "km_assign_rgb_cull", 110253056, 0.43)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == 6471.1 and r["traffic"] == 110253056
    assert r["achieved"] == pytest.approx(3 * n / 1e-4 / 1e9) and r["frac"] == pytest.approx(r["achieved"] / 6471.1)
    f = r["fp32"]
    assert f["peak"] == pytest.approx(74.45, rel=1e-3)
    assert f["algorithmic_equiv_tflops"] == pytest.approx(7 * n * k / 1e-4 / 1e12)
    assert f["brute_force_kernel"]["frac"] == pytest.approx(7 * n * k / 0.43e-3 / 1e12 / f["peak"])
    assert bench.build_roofline(5, 100, 10, 1.0, 1000.0, pk, "x", None, 0.0)["fp32"]["brute_force_kernel"] is None
    import json
    json.dumps(r)


def test_curve_shard_and_histogram_merge():
    from cniic_b200 import dist as cdist
    for n, world in ((8192 * 8192, 8), (100 * 70, 3), (4096, 4), (1, 2), (0, 3)):
        edges = [cdist.curve_shard(n, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert all(e[0] % 4096 == 0 for e in edges)
    if True:
        k, c = cdist.merge_histograms([(np.array([1, 5, 9], np.uint32), np.array([2, 3, 4], np.uint64)),
                                       (np.array([0, 5], np.uint32), np.array([7, 10], np.uint64)), (np.zeros(0, np.uint32), np.zeros(0, np.uint64))])
        assert k.tolist() == [0, 1, 5, 9] and c.tolist() == [7, 2, 13, 4]


# ---- the Rust boundary (source only here: no cargo in the image) is kept honest mechanically ----

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_rust_ffi_matches_header():
    """rust/cniic-cuda-sys/src/ffi.rs is generated from include/cniic_b200.h: every entry point, same arity, same order."""
    import re
    g = _gen()
    committed = open(g.OUT).read()
    assert committed == g.generate(), "ffi.rs is stale: run python tools/gen_rust_ffi.py"
    rust = dict((m.group(1), m.group(2)) for m in re.finditer(r"pub fn (cniic_\w+)\((.*?)\)", committed))
    _, _, protos = g.parse_header(open(g.HEADER).read())
    assert sorted(rust) == _lib.declared_symbols() == sorted(p[0] for p in protos)
    for name, _, args in protos:
        n_rust = len([a for a in rust[name].split(",") if a.strip()])
        assert n_rust == len(args), name
    # pointer constness survives the translation (spot checks)
    assert "rgb: *const *const u8" in rust["cniic_kmeans_rgb_batch"] and "out_assign: *const *mut u16" in rust["cniic_kmeans_rgb_batch"]
    assert "sessions: *const *mut cniic_kmeans" in rust["cniic_kmeans_run_batch"]
    # struct layouts follow the ctypes mirrors field by field
    for cname, ct in (("cniic_kmeans_stats", _lib.KMeansStats), ("cniic_kmeans_desc", _lib.KMeansDesc)):
        body = re.search(r"pub struct %s \{(.*?)\}" % cname, committed, flags=re.S).group(1)
        assert re.findall(r"pub (\w+):", body) == [f[0] for f in ct._fields_], cname


def test_rust_build_compiles_every_cu_file():
    """build.rs once listed the .cu files by hand and missed huffdec.cu -> the crate could not link."""
    import re
    text = open(os.path.join(ROOT, "rust", "cniic-cuda-sys", "build.rs")).read()
    code = "\n".join(l for l in text.splitlines() if not l.strip().startswith("//"))
    assert "read_dir" in code and 'x == "cu"' in code
    csrc = os.path.join(ROOT, "cniic_b200", "csrc")
    stems = [f[:-3] for f in os.listdir(csrc) if f.endswith(".cu")]
    assert "huffdec" in stems
    assert not [s for s in stems if re.search(r'"%s"' % s, code)], "build.rs must not hard-code a source list"
    # the wrapper crate re-exports the generated declarations and the adapter file a cniic maintainer adds exists
    lib = open(os.path.join(ROOT, "rust", "cniic-cuda-sys", "src", "lib.rs")).read()
    assert "mod ffi;" in lib and "pub use ffi::*;" in lib and 'extern "C"' not in lib
    gpuc = open(os.path.join(ROOT, "rust", "cniic-side", "gpuc.rs")).read()
    assert "impl Codec for Gpu" in gpuc and "impl FromStr for Gpu" in gpuc
    for used in re.findall(r"\b(cniic_[a-z0-9_]+)\(", lib):
        assert used in _lib.declared_symbols(), used


def test_oracle_side_image_generator_equals_the_library_generator():
    """oracle/synth.py (numpy) exists so that bench.py's CPU legs never load the product; it must produce the very bytes of
    cniic_synth_image_host, shards included."""
    from oracle import synth
    for (w, h, seed, nb, y0, ht) in [(96, 64, 42, 12, 0, None), (96, 24, 42, 12, 40, 64), (513, 37, 0xC0FFEE + 2, 192, 0, None),
                                     (7680, 6, 0xC0FFEE + 3, 2048, 1000, 4320), (33, 7, 5, 0, 0, None), (8, 8, 1, 1000, 0, None)]:
        assert np.array_equal(synth.synth_image(w, h, seed, nb, y0, ht), cb.synth_image_host(w, h, seed, nb, y0, ht)), (w, h, seed)


def test_reference_arm_does_not_touch_the_product():
    """`bench.py --impl reference` imported cniic_b200 (and so mapped the product .so) just to build its image."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    cpu_side = {"reference_arm", "cpu_reference_leg", "workload_image", "cpu_sample_plan"}
    seen = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in cpu_side:
            seen.add(node.name)
            src = ast.dump(node)
            assert "cniic_b200" not in src and "'cb'" not in src, node.name
    assert seen == cpu_side
    # ... and nothing at module level imports it either (the GPU arm imports it inside gpu_arm / main)
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            assert "cniic_b200" not in ast.dump(node)
