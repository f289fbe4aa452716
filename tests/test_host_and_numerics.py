"""CPU-side tests (no GPU): the numerics contract vs glibc, the C-ABI library's exports and its
loud failure without a device, the sweep sharding logic under gloo (world_size 2), and closed-form
known-answer tests of the oracle."""
import ctypes as C
import ctypes.util
import os
import re
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle_py
    return oracle_py


@pytest.fixture(scope="module")
def api(built):
    from mcray_tracing_b200 import api as a
    return a


# ---- numerics contract vs glibc ----------------------------------------------------------------------
def _ulp32(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_contract_float_functions_match_glibc_within_one_ulp(O):
    libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    for f in ("expf", "logf"):
        getattr(libm, f).restype = C.c_float
        getattr(libm, f).argtypes = [C.c_float]
    libm.powf.restype = C.c_float
    libm.powf.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(5)
    n = 60000
    x = rng.uniform(-40, 10, n).astype(np.float32)
    ref = np.array([libm.expf(float(v)) for v in x], np.float32)
    d = _ulp32(O.numerics(0, x).astype(np.float32), ref)
    assert d.max() <= 1 and np.mean(d == 0) > 0.995
    y = np.exp(rng.uniform(-25, 5, n)).astype(np.float32)
    ref = np.array([libm.logf(float(v)) for v in y], np.float32)
    d = _ulp32(O.numerics(1, y).astype(np.float32), ref)
    assert d.max() <= 1 and np.mean(d == 0) > 0.995
    b = rng.uniform(0, 1, n).astype(np.float32)
    e = rng.choice([0.001, 0.2, 0.5, 1.0, 2.0, 3.0], n).astype(np.float32)
    ref = np.array([libm.powf(float(u), float(v)) for u, v in zip(b, e)], np.float32)
    d = _ulp32(O.numerics(2, b, e).astype(np.float32), ref)
    assert d.max() <= 1 and np.mean(d == 0) > 0.995


def test_contract_double_functions_accuracy(O):
    import math
    rng = np.random.default_rng(6)
    a = rng.uniform(0, 2 * np.pi, 20000)
    assert np.abs(O.numerics(3, a) - np.sin(a)).max() < 4e-16
    assert np.abs(O.numerics(4, a) - np.cos(a)).max() < 4e-16
    x = rng.uniform(-700, 700, 20000)
    ref = np.array([math.exp(v) for v in x])
    assert np.abs(O.numerics(7, x) / ref - 1).max() < 5e-16
    y = np.exp(rng.uniform(-700, 700, 20000))
    ref = np.array([math.log(v) for v in y])
    assert np.abs(O.numerics(8, y) - ref).max() <= 5e-16 * np.abs(ref).max()
    # special cases of pow the hot path can reach (ray.cpp:131,158,160,223)
    vals = O.numerics(6, np.array([-0.5, -0.5, -0.5, 0.0, 0.0, 2.0, 1.0, np.nan]), np.array([1.0, 2.0, 0.2, 1.0, 0.0, 0.0, np.nan, 1.0]))
    assert vals[0] == -0.5 and vals[1] == 0.25 and np.isnan(vals[2]) and vals[3] == 0.0 and vals[4] == 1.0 and vals[5] == 1.0
    assert vals[6] == 1.0 and np.isnan(vals[7])


def test_envelope_alpha_division_is_exact(O):
    """k_post_tma evaluates the envelope's alpha = (j - p) / (q - p) (rfimage.h:80) with Markstein's three-instruction division
    from a correctly rounded reciprocal instead of the IEEE division; exhaustive over every pair the kernel can meet
    (0 <= a < b <= 2048, the kernel is limited to 640 rows) the two are bit-identical."""
    b = np.arange(1, 2049, dtype=np.float64)
    A, B = np.meshgrid(np.arange(0, 2048, dtype=np.float64), b, indexing="ij")
    keep = A < B
    a, bb = A[keep], B[keep]
    got = O.numerics(9, a, bb).astype(np.float32)
    ref = a.astype(np.float32) / bb.astype(np.float32)
    assert got.size > 2_000_000 and np.array_equal(got, ref)


def test_philox_known_answer(O):
    """Philox4x32-10 known-answer vectors of Random123 (kat_vectors): zero counter/key and the
    'pi' vector; checked through an independent pure-Python restatement of the round function."""
    def philox(c, k):
        c, k = list(c), list(k)
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xffffffff, p1 & 0xffffffff, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xffffffff, p0 & 0xffffffff]
            k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
        return c
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # the contract's keyed block: counter = (frame, element, sample, bounce*8+block), key = seed halves
    a = np.array([0.0, 1.0, 123456.0, 2**31 - 1.0])
    b = np.array([0.0, 7.0, 99.0, 2**20 - 1.0])
    got = O.numerics(5, a, b)
    seed = 0x0123456789abcdef
    for i in range(len(a)):
        w = philox([int(a[i]), int(b[i]), 3, 2 * 8 + 1], [seed & 0xffffffff, seed >> 32])
        assert got[i] == float(w[0]) + 4294967296.0 * float(w[3] & 0xfffff)


# ---- the C-ABI library ---------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(api):
    header = (ROOT / "include" / "mcrt.h").read_text()
    declared = set(re.findall(r"\b(mcrt_[a-z_0-9]+)\s*\(", header))
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    L = api.lib()
    for name in sorted(declared):
        assert getattr(L, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", str(api.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mcrt_[a-z_0-9]+)", out))
    assert declared <= exported


def test_header_is_plain_c_and_links(api, tmp_path):
    """include/mcrt.h is a C header (the reference-side binding may be C): it compiles as strict C99, and a C program that takes the
    address of every declared entry point links against libmcrt.so."""
    header = (ROOT / "include" / "mcrt.h").read_text()
    names = sorted(set(re.findall(r"\b(mcrt_[a-z_0-9]+)\s*\(", header)))
    src = tmp_path / "use_all.c"
    src.write_text('#include "mcrt.h"\n#include <stdio.h>\nint main(void) {\n    mcrt_params p; mcrt_bmode_params b; mcrt_pose q; mcrt_info i; mcrt_stats s;\n'
                   '    (void)p; (void)b; (void)q; (void)i; (void)s;\n    const void* f[] = {' + ", ".join("(const void*)" + n for n in names) + '};\n'
                   '    printf("%d\\n", (int)(sizeof(f) / sizeof(f[0])));\n    return 0;\n}\n')
    exe = tmp_path / "use_all"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                        "-L", str(api.LIB_PATH.parent), "-lmcrt", "-Wl,-rpath," + str(api.LIB_PATH.parent)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and int(out.stdout) == len(names)


def test_library_contains_sm100a_code_only(api):
    out = subprocess.run(["cuobjdump", "-lelf", str(api.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_gpu_means_loud_failure_not_fallback(api, assets_dirs):
    """In this container there is no GPU: creating a context must fail with MCRT_ERR_CUDA -- the
    product has no CPU path.  (On a GPU box this test is skipped.)"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(api.McrtError) as e:
        api.Simulator(assets_dirs["sphere"] / "sphere.scene")
    assert e.value.code == api.MCRT_ERR_CUDA


def test_product_does_not_reference_the_oracle():
    """The shipped path must not import, link or call anything under oracle/."""
    pkg = ROOT / "mcray_tracing_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("Makefile")):
        text = path.read_text(errors="replace")
        assert "oracle_py" not in text and "liboracle" not in text and "mcrt_oracle" not in text, path
        assert not re.search(r"(from|import)\s+oracle\b", text), path
    out = subprocess.run(["ldd", str(pkg / "libmcrt.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_scene_errors_and_probe(api, assets_dirs, tmp_path):
    info = api.scene_probe(assets_dirs["sphere"] / "sphere.scene")
    assert info["n_triangles"] == 12 + 20480 and info["n_meshes"] == 2 and info["n_materials"] == 11
    assert np.array_equal(info["start_pose"], np.array([-13.5, 0, 0, 0, 0, -90], np.float32))
    assert api.scene_probe(assets_dirs["ircad11"] / "ircad11.scene")["n_triangles"] == 624640   # defaulted shininess/thickness
    cases = {"nojson.scene": "{", "nomesh.scene": '{"transducerPosition":[0,0,0],"origin":[0,0,0],"spacing":[1,1,1],"startingMaterial":"A","scaling":1,'
             '"materials":[{"name":"A","impedance":1,"attenuation":1,"mu0":0,"mu1":0,"sigma":0,"specularity":1}],"meshes":3}',
             "badmat.scene": '{"transducerPosition":[0,0,0],"origin":[0,0,0],"spacing":[1,1,1],"startingMaterial":"B","scaling":1,'
             '"materials":[{"name":"A","impedance":1,"attenuation":1,"mu0":0,"mu1":0,"sigma":0,"specularity":1}],"meshes":[]}'}
    for name, text in cases.items():
        (tmp_path / name).write_text(text)
        with pytest.raises(api.McrtError) as e:
            api.scene_probe(tmp_path / name)
        assert e.value.code == api.MCRT_ERR_SCENE and e.value.message.startswith("Error while loading scene: ")
    with pytest.raises(api.McrtError) as e:
        api.host_tables(api.default_params(psf_lateral=12))
    assert e.value.code == api.MCRT_ERR_INVALID


def test_runtime_sizes(api, O):
    """BASELINE configs: 256 x 16 keeps 465 rows; a finer axial grid gives config 5's ~8192 samples."""
    for kw, rows in ((dict(elements=256, samples=16), 465), (dict(elements=1024, samples=8, axial_scale=17.6), 8333)):
        info = api.host_tables(api.default_params(**kw))["info"]
        d = O.derive(O.default_params(**kw))
        assert info.rows == d.rows == rows and info.cols == kw["elements"]
        assert info.time_step_us == d.time_step_us and info.row_period_us == d.row_period_us


# ---- sweep sharding --------------------------------------------------------------------------------------
def test_shard_bounds_partition():
    from mcray_tracing_b200 import sweep
    for n in (0, 1, 7, 512, 513):
        for w in (1, 2, 3, 4, 8):
            b = [sweep.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == sweep.all_shard_sizes(n, w)
            # the round-robin deal: a partition too, rank r holds r, r + w, ...
            idx = [sweep.shard_indices(n, w, r, interleave=True) for r in range(w)]
            assert sorted(np.concatenate(idx).tolist()) == list(range(n)) and [len(i) for i in idx] == sweep.interleaved_sizes(n, w)
            assert all((i % w == r).all() for r, i in enumerate(idx))
            assert [sweep.shard_indices(n, w, r).tolist() for r in range(w)] == [list(range(*b[r])) for r in range(w)]


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from mcray_tracing_b200 import sweep
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
n = int(sys.argv[1])
poses = np.arange(n * 6, dtype=np.float32).reshape(n, 6)
stride = 1
def fake(block, first_frame, out):          # stands in for Simulator.simulate_device: a pure function of (pose, global frame)
    for i in range(len(block)):
        out[i] = torch.arange(12, dtype=torch.float32).reshape(3, 4) * float(first_frame + i * stride + 1) + float(block[i, 0])
res = sweep.run_sweep(fake, poses, (3, 4), torch.device("cpu"), seed_first_frame=100)
stride = dist.get_world_size()               # option frame_stride = G of the round-robin deal
res_il = sweep.run_sweep(fake, poses, (3, 4), torch.device("cpu"), seed_first_frame=100, interleave=True)
stride = 1
if dist.get_rank() == 0:
    single = torch.empty((n, 3, 4))
    fake(poses, 100, single)
    assert res.shape == single.shape and torch.equal(res, single), "sharded sweep differs from the single-rank sweep"
    assert res_il.shape == single.shape and torch.equal(res_il, single), "round-robin sweep differs from the single-rank sweep"
    print("OK", n)
else:
    assert res is None and res_il is None
dist.destroy_process_group()
"""


@pytest.mark.parametrize("n_poses", [8, 7])
def test_sweep_gather_world_size_2_gloo(tmp_path, n_poses):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(n_poses)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert f"OK {n_poses}" in outs[0][0]


_WORKER_SCANLINES = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from mcray_tracing_b200 import sweep
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
E, rows, kl = int(sys.argv[1]), 5, 4
raw = torch.arange(E * rows, dtype=torch.float32).reshape(E, rows) % 7.0        # what tracing scanline e yields: a pure function of e
taps = torch.tensor([0.5, 0.25, 0.125, 0.0625])
def lateral(block, first, total):                                                # forward-looking taps, raw borders (rfimage.h:111-122)
    out = block.clone()
    for c in range(block.shape[0]):
        g = first + c
        if g >= kl // 2 and g < total - kl and c + kl <= block.shape[0]:
            out[c] = sum(block[c + k] * taps[k] for k in range(kl))
    return out
def fake(first, n, out):                  # stands in for Simulator.simulate_scanlines: re-traces its right-hand halo locally
    e1 = min(first + n + kl - 1, E)
    out[:] = lateral(raw[first:e1], first, E)[:n]
res = sweep.run_frame_scanline_blocks(fake, E, rows, torch.device("cpu"))
if dist.get_rank() == 0:
    assert res.shape == (E, rows) and torch.equal(res, lateral(raw, 0, E)), "scanline-block frame differs from the whole frame"
    print("OK", E)
else:
    assert res is None
dist.destroy_process_group()
"""


@pytest.mark.parametrize("n_elements", [16, 9])
def test_frame_scanline_blocks_world_size_2_gloo(tmp_path, n_elements):
    """SURVEY 8(e) secondary partition, host logic: contiguous scanline blocks + recomputed halo + one gather."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker_scanlines.py"
    script.write_text(_WORKER_SCANLINES.format(root=str(ROOT)))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(n_elements)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert f"OK {n_elements}" in outs[0][0]


# ---- closed-form known answers for the oracle ----------------------------------------------------------
def _box_scene(O, half=2.0):
    v = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * half
    q = [[0, 3, 2, 1], [4, 5, 6, 7], [0, 1, 5, 4], [2, 3, 7, 6], [1, 2, 6, 5], [0, 4, 7, 3]]
    tris = []
    for a, b, c, d in q:
        tris += [np.concatenate([v[a], v[b], v[c]]), np.concatenate([v[a], v[c], v[d]])]
    mats = np.array([[1.5, 0.5, 0.1, 0.2, 0.1, 1, 1e6, 0], [1.5, 0.7, 0.2, 0.3, 0.2, 1, 1e6, 0], [3.0, 1.0, 0.2, 0.3, 0.2, 1, 1e6, 0]], np.float32)
    return dict(materials=mats, starting_material=0, mesh_material_inside=np.array([1], np.int32), mesh_material_outside=np.array([0], np.int32),
                mesh_vascular=np.array([0], np.int32), mesh_deltas=np.zeros((1, 3), np.float32), tri_offsets=np.array([0, 12], np.int64),
                tri_vertices=np.array(tris, np.float32), scaling=1.0, origin=np.zeros(3, np.float32), spacing=np.ones(3, np.float32))


def test_ray_box_closed_form(O):
    osc = O.OracleScene(_box_scene(O))
    rng = np.random.default_rng(2)
    for _ in range(300):
        o = np.array([-5.0, rng.uniform(-1.9, 1.9), rng.uniform(-1.9, 1.9)], np.float32)
        t = o + np.array([10.0, 0, 0], np.float32)
        for use_bvh in (False, True):
            tri, mesh, f = osc.closest_hit(o, t, use_bvh)
            assert tri >= 0 and mesh == 0
            assert abs(f[0] - 0.3) < 1e-6                                  # enters the x = -2 face at 3/10 of the segment
            assert np.allclose(f[1:4], [-2.0, o[1], o[2]], atol=1e-5)
            assert np.allclose(f[4:7], [-1.0, 0, 0], atol=1e-6)            # normal faces the ray origin
    assert osc.closest_hit([-5, 3, 0], [5, 3, 0])[0] == -1                 # passes above the box
    assert osc.closest_hit([-5, 0, 0], [-3, 0, 0])[0] == -1                # stops short


def test_snell_and_intensity_closed_forms(O):
    L = O.oracle()
    assert L.orc_reflection_intensity(0.7, 1.5, 0.8, 1.5, 0.8) == 0.0      # matched impedance: nothing reflected
    z1, z2 = 1.38, 7.8
    r = L.orc_reflection_intensity(1.0, z1, 1.0, z2, 1.0)
    assert abs(r - ((z1 - z2) / (z1 + z2)) ** 2) < 1e-6                    # normal incidence
    o3 = np.zeros(3, np.float32)
    l = np.array([1, 0, 0], np.float32); n = np.array([-1, 0, 0], np.float32)
    L.orc_snells_law(l.ctypes.data, n.ctypes.data, 1.0, 1.0, 1.0, o3.ctypes.data)
    assert np.array_equal(o3, l)                                           # no deviation at normal incidence, equal media
    oi, od = C.c_float(), C.c_double()
    L.orc_travel(0.7, 0.2, 4.5, 0.0, 30.0, C.byref(oi), C.byref(od))       # Beer-Lambert; SURVEY.md C-3c value
    assert abs(oi.value - 0.2 * np.exp(-0.7 * 0.3 * 4.5)) < 1e-7 and abs(oi.value - 0.0777359232) < 1e-8 and od.value == 30.0
    assert abs(L.orc_max_ray_length(0.7, 0.2, 4.5) - 1376.769) < 1e-2


def test_psf_of_delta_and_envelope_of_sinusoid(O):
    p = O.default_params()
    ax, lat = O.psf_taps(p)
    img = np.zeros((100, 60), np.float32)
    img[50, 30] = 1.0
    out = O.convolve(img, ax, lat)
    exp = np.zeros_like(img)
    for k in range(7):
        for m in range(13):
            exp[50 - k, 30 - m] = ax[k] * lat[m]                           # forward-looking taps: the delta spreads up/left
    assert np.allclose(out, exp, atol=1e-7)
    t = np.arange(400, dtype=np.float32)
    sig = (np.sin(2 * np.pi * t / 16.0)).astype(np.float32)
    col = np.tile(sig[:, None], (1, 4))
    env = O.envelope(col)
    inner = env[8:380, 0]
    assert inner.min() > 0.99 and inner.max() < 1.0 + 1e-6                 # envelope of a unit sinusoid ~ 1 between its peaks


# ---- the headless CLI's argument and error behaviour (no GPU needed: it fails before any CUDA call) ------------------------
def test_cli_argument_and_scene_errors(api, tmp_path):
    """`mattausch` keeps the reference's messages: wrong argument list -> "Incorrect argument list." and exit code 0
    (main.cpp:46-50); an unloadable scene -> "The program found an error and will terminate." + the scene.cpp:23-26 reason."""
    exe = Path(api.__file__).resolve().parent / "mattausch"
    assert exe.exists()
    for argv in ([], ["--frames", "2"], [str(tmp_path / "x.scene"), "--no-such-flag"]):
        out = subprocess.run([str(exe)] + argv, capture_output=True, text=True, timeout=60)
        assert out.returncode == 0 and out.stdout.strip() == "Incorrect argument list.", (argv, out.stdout, out.stderr)
    out = subprocess.run([str(exe), str(tmp_path / "missing.scene")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0
    assert "The program found an error and will terminate." in out.stdout and "Error while loading scene" in out.stdout


# ---- asset conversion (replaces utils/vtp_to_obj.py) ------------------------------------------------------
def _vtp_text(points, polys, strips, flavour):
    import base64, struct, zlib
    pts = np.asarray(points, np.float32)
    conn = np.concatenate(polys).astype(np.int64); offs = np.cumsum([len(c) for c in polys]).astype(np.int64)
    sconn = np.concatenate(strips).astype(np.int64) if strips else np.zeros(0, np.int64)
    soffs = np.cumsum([len(c) for c in strips]).astype(np.int64) if strips else np.zeros(0, np.int64)
    arrays = [("Points", None, pts.ravel(), "Float32", 3), ("Polys", "connectivity", conn, "Int64", 1), ("Polys", "offsets", offs, "Int64", 1),
              ("Strips", "connectivity", sconn, "Int64", 1), ("Strips", "offsets", soffs, "Int64", 1)]
    hdr = "<Q" if flavour == "appended_raw64" else "<I"
    attrs = 'byte_order="LittleEndian"' + (' header_type="UInt64"' if hdr == "<Q" else "") + (' compressor="vtkZLibDataCompressor"' if flavour == "zlib" else "")
    appended = b""
    def payload(a):
        raw = a.tobytes()
        if flavour == "zlib":
            comp = zlib.compress(raw)
            head = struct.pack("<4I", 1, len(raw), len(raw), len(comp))
            return base64.b64encode(head).decode() + base64.b64encode(comp).decode()
        return base64.b64encode(struct.pack(hdr, len(raw)) + raw).decode()
    def da(name, a, typ, nc):
        nonlocal appended
        nm = f' Name="{name}"' if name else ""
        if flavour == "ascii":
            return f'<DataArray type="{typ}"{nm} NumberOfComponents="{nc}" format="ascii">{" ".join(repr(float(x)) if typ.startswith("F") else str(int(x)) for x in a)}</DataArray>'
        if flavour == "appended_raw64":
            off = len(appended)
            appended += struct.pack(hdr, a.nbytes) + a.tobytes()
            return f'<DataArray type="{typ}"{nm} NumberOfComponents="{nc}" format="appended" offset="{off}"/>'
        return f'<DataArray type="{typ}"{nm} NumberOfComponents="{nc}" format="binary">{payload(a)}</DataArray>'
    sec = {"Points": "", "Polys": "", "Strips": ""}
    for s_, name, a, typ, nc in arrays:
        sec[s_] += da(name, a, typ, nc)
    body = (f'<?xml version="1.0"?><VTKFile type="PolyData" version="1.0" {attrs}><PolyData><Piece NumberOfPoints="{len(pts)}" NumberOfPolys="{len(polys)}" '
            f'NumberOfStrips="{len(strips)}"><Points>{sec["Points"]}</Points><Polys>{sec["Polys"]}</Polys><Strips>{sec["Strips"]}</Strips></Piece></PolyData>').encode()
    if flavour == "appended_raw64":
        body += b'<AppendedData encoding="raw">_' + appended + b"</AppendedData>"
    return body + b"</VTKFile>"


@pytest.mark.parametrize("flavour", ["ascii", "base64", "zlib", "appended_raw64"])
def test_vtp_to_obj_matches_the_loader(api, tmp_path, flavour):
    """convert.py reads VTK XML PolyData without vtk (ascii / base64 / zlib / appended raw) and writes OBJ files the C++ loader
    (mcrt_load_obj = tinyobj + objloader.h semantics) turns into exactly the fan-triangulated polygons."""
    from mcray_tracing_b200 import convert
    rng = np.random.default_rng(11)
    pts = rng.normal(size=(12, 3)).astype(np.float32) * 37.5
    polys = [np.array([0, 1, 2]), np.array([2, 3, 4, 5]), np.array([5, 6, 7, 8, 9]), np.array([11, 10, 0])]
    strips = [np.array([3, 4, 5, 6, 7])]
    (tmp_path / "m.vtp").write_bytes(_vtp_text(pts, polys, strips, flavour))
    rp, rpoly = convert.read_vtp(tmp_path / "m.vtp")
    assert np.array_equal(rp.astype(np.float32), pts)
    want_polys = polys + [np.array([3, 4, 5]), np.array([5, 4, 6]), np.array([5, 6, 7])]
    assert len(rpoly) == len(want_polys) and all(np.array_equal(a, b) for a, b in zip(rpoly, want_polys))
    n = convert.convert(tmp_path / "m.vtp", tmp_path / "m.obj")
    soup = api.load_obj(tmp_path / "m.obj")
    assert n == len(soup) == 1 + 2 + 3 + 1 + 3
    assert np.array_equal(soup.reshape(-1, 3, 3), convert.indexed_to_soup(pts, want_polys))


def test_stl_to_obj_weld_and_dump(api, tmp_path):
    import struct
    from mcray_tracing_b200 import convert
    rng = np.random.default_rng(12)
    v = rng.normal(size=(9, 3)).astype(np.float32)
    tri = np.array([[0, 1, 2], [2, 1, 3], [3, 4, 5], [6, 7, 8], [8, 7, 0]])
    soup = v[tri]
    with open(tmp_path / "b.stl", "wb") as f:
        f.write(b"solid looks-like-ascii".ljust(80, b" ") + struct.pack("<I", len(soup)))
        for t in soup:
            f.write(struct.pack("<12fH", 0, 0, 0, *t.ravel(), 0))
    with open(tmp_path / "a.stl", "w") as f:
        f.write("solid s\n" + "".join("facet normal 0 0 0\n outer loop\n" + "".join("  vertex %.9g %.9g %.9g\n" % tuple(p) for p in t) + " endloop\nendfacet\n"
                                       for t in soup) + "endsolid s\n")
    for name in ("a.stl", "b.stl"):
        assert np.array_equal(convert.read_stl(tmp_path / name), soup)
        assert convert.convert(tmp_path / name, tmp_path / (name + ".obj")) == len(soup)
        assert np.array_equal(api.load_obj(tmp_path / (name + ".obj")).reshape(-1, 3, 3), soup)
    verts, idx = convert.weld(soup)
    assert len(verts) == 9 and np.array_equal(verts[idx], soup) and np.array_equal(convert.indexed_to_soup(verts, idx), soup)
    convert.convert(tmp_path / "b.stl", tmp_path / "nw.obj", weld_vertices=False)
    assert sum(ln.startswith("v ") for ln in (tmp_path / "nw.obj").read_text().splitlines()) == 3 * len(soup)
    assert np.array_equal(api.load_obj(tmp_path / "nw.obj").reshape(-1, 3, 3), soup)


def test_host_sah_builder_is_valid_and_thread_independent(api):
    """The host tree of the background optimisation (mcrt.h "bvh_optimise"; sah_builder.cpp): a valid BVH2 -- pre-order ids, every triangle
    in exactly one leaf, every child box the exact union of what lies below it -- that does not depend on the number of threads
    (subtrees are forked from 8192 triangles up, so a 49 152-triangle scene forks several times)."""
    from mcray_tracing_b200 import assets
    A = assets.stress_scene_arrays(shells=3, nu=128, nv=64)
    tri = (A["tri_vertices"] * np.float32(A["scaling"])).astype(np.float32).reshape(-1, 9)
    offs = A["tri_offsets"]
    mesh = np.concatenate([np.full(int(offs[m + 1] - offs[m]), m, np.int32) for m in range(len(offs) - 1)])
    origins = (A["mesh_deltas"] * np.float32(A["scaling"])).astype(np.float32)
    origins[1] += np.float32(0.37)                                       # body origins enter the world boxes
    n = len(mesh)
    nodes1, slots1, depth1 = api.host_build_sah(tri, mesh, origins, threads=1)
    for threads in (2, 5, 0):
        nodes, slots, depth = api.host_build_sah(tri, mesh, origins, threads=threads)
        assert depth == depth1 and np.array_equal(slots, slots1) and nodes.tobytes() == nodes1.tobytes(), threads
    assert nodes1.shape == (n - 1, 16) and sorted(slots1.tolist()) == list(range(n)) and 16 <= depth1 <= 96
    child = nodes1[:, 12:14].copy().view(np.int32)
    world = tri.reshape(n, 3, 3) + origins[mesh][:, None, :]
    lo, hi = world.min(axis=1), world.max(axis=1)
    # bottom-up in reverse pre-order: children always have larger ids than their parent
    box_lo, box_hi = np.empty((n - 1, 3), np.float32), np.empty((n - 1, 3), np.float32)
    seen_nodes, seen_slots = np.zeros(n - 1, bool), np.zeros(n, bool)
    for i in range(n - 2, -1, -1):
        los, his = [], []
        for k in range(2):
            ch = int(child[i, k])
            if ch >= 0:
                assert i < ch < n - 1 and not seen_nodes[ch]
                seen_nodes[ch] = True
                l, h = box_lo[ch], box_hi[ch]
            else:
                code = -ch - 1
                slot = code >> 2
                assert code & 3 == 0 and 0 <= slot < n and not seen_slots[slot]
                seen_slots[slot] = True
                l, h = lo[slots1[slot]], hi[slots1[slot]]
            assert np.array_equal(nodes1[i, 6 * k:6 * k + 3], l) and np.array_equal(nodes1[i, 6 * k + 3:6 * k + 6], h), (i, k)
            los.append(l); his.append(h)
        box_lo[i], box_hi[i] = np.minimum(los[0], los[1]), np.maximum(his[0], his[1])
    assert seen_slots.all() and seen_nodes[1:].all() and not seen_nodes[0]
    assert int(child[0, 0]) == 1                                          # pre-order: the left child of the root follows it


def test_closest_hit_agrees_with_an_independent_moller_trumbore(O, assets_dirs):
    """The one piece of the oracle that cannot be pinned to reference code is Bullet's ray / triangle arithmetic.
    Independent geometric check: on the 20 480-triangle sphere scene the oracle's closest hit (brute force and BVH) must agree with a
    textbook Moller-Trumbore intersection evaluated in float64 over all triangles -- same triangle (unless two candidates are closer than
    fp32 can tell apart), fraction and hit point to fp32 accuracy, unit normal parallel to the triangle's and facing the ray origin."""
    A = O.load_scene_py(assets_dirs["sphere"] / "sphere.scene")
    osc = O.OracleScene(A)
    s = np.float64(A["scaling"])
    tri = np.asarray(A["tri_vertices"], np.float64).reshape(-1, 3, 3) * s
    offs = np.asarray(A["tri_offsets"])
    mesh = np.concatenate([np.full(int(offs[m + 1] - offs[m]), m) for m in range(len(offs) - 1)])
    org = np.asarray(A["mesh_deltas"], np.float64) * s * s + np.asarray(A["origin"], np.float64)[None, :]
    w = tri + org[mesh][:, None, :]
    v0, e1, e2 = w[:, 0], w[:, 1] - w[:, 0], w[:, 2] - w[:, 0]
    nrm = np.cross(e1, e2)
    lo, hi = w.reshape(-1, 3).min(0), w.reshape(-1, 3).max(0)
    c, r = 0.5 * (lo + hi), 0.5 * np.linalg.norm(hi - lo)
    rng = np.random.default_rng(11)
    checked = missed = 0
    for k in range(200):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        side = np.cross(d, rng.normal(size=3)); side /= np.linalg.norm(side)
        # three rays in four aim at the body, the fourth passes it at 1.1 - 1.6 bounding radii
        off = side * r * (rng.uniform(1.1, 1.6) if k % 4 == 3 else rng.uniform(0.0, 0.45))
        o32 = (c - d * 1.5 * r + off).astype(np.float32)
        t32 = (o32.astype(np.float64) + d * 3.0 * r).astype(np.float32)
        o, dd = o32.astype(np.float64), t32.astype(np.float64) - o32.astype(np.float64)
        p = np.cross(dd[None, :], e2)
        det = np.einsum("ij,ij->i", e1, p)
        ok = np.abs(det) > 1e-14
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        tv = o[None, :] - v0
        u = np.einsum("ij,ij->i", tv, p) * inv
        q = np.cross(tv, e1)
        v = np.einsum("j,ij->i", dd, q) * inv
        tt = np.einsum("ij,ij->i", e2, q) * inv
        m = 2e-4                                                         # stay clear of edges: Bullet's edge tolerance is not Moller-Trumbore's
        inside = ok & (u > m) & (v > m) & (u + v < 1 - m) & (tt > 1e-6) & (tt < 1.0)
        near_edge = ok & (u > -m) & (v > -m) & (u + v < 1 + m) & (tt > -1e-6) & (tt < 1.0 + 1e-6) & ~inside
        for use_bvh in (False, True):
            tid, _, f = osc.closest_hit(o32, t32, use_bvh)
            if not inside.any():
                if not near_edge.any():
                    assert tid == -1
                    missed += 1
                continue
            cand = np.where(inside)[0]
            best = cand[np.argmin(tt[cand])]
            if near_edge.any() and tt[near_edge].min() < tt[best] + 1e-5:
                continue                                                 # an edge-grazing candidate in front: either answer is defensible
            assert tid >= 0 and abs(f[0] - tt[best]) < 2e-6 * max(1.0, 1.0 / max(tt[best], 1e-3)), (tid, best, f[0], tt[best])
            others = cand[cand != best]
            if not (len(others) and tt[others].min() < tt[best] + 1e-6):
                assert tid == best
            hit = o + dd * tt[best]
            assert np.allclose(f[1:4], hit, atol=2e-4 * max(1.0, r))
            n = nrm[tid] / np.linalg.norm(nrm[tid])
            assert abs(abs(np.dot(f[4:7], n)) - 1.0) < 1e-4 and np.dot(f[4:7], dd) < 0.0
            checked += 1
    assert checked > 200 and missed > 50


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm) runs without a GPU and prints ONE JSON line with the
    contract's keys: the metric / unit / config of the GPU arm, `impl`, a `cpu_baseline` describing the run and an `e2e` with no copies."""
    import json
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b["impl"] == "reference" and b["metric"] == "rf_frames_per_s" and b["unit"] == "frames/s" and b["higher_is_better"] is True
    assert b["n_gpus"] == 1 and b["steps"] == 1 and b["value"] > 0 and b["ms_per_step"] > 0 and b["vs_baseline"] is None
    assert b["config"]["config"] == "c2" and b["config"]["elements"] == 256 and b["config"]["samples_per_element"] == 16
    cb = b["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == b["value"] and "sample" in cb
    e = b["e2e"]
    assert e["value"] == b["value"] and e["unit"] == b["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
