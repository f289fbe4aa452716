"""GPU parity at the other BASELINE configurations (parity cases, not bench lines) and edge cases:
config 4 (2 097 152-triangle nested tissue mesh, rough surfaces, 10 bounces), config 5 (long
scanlines / large PSF), degenerate scenes (no mesh, one triangle, two triangles)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tol(ref):
    return 1e-4 * np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle_py
    return oracle_py


@pytest.fixture(scope="module")
def api(built):
    from mcray_tracing_b200 import api as a
    return a


def _segments_equal(gs, gn, os_, on):
    assert np.array_equal(gn, on)
    valid = np.arange(gs.shape[-1])[None, None, :] < on[:, :, None]
    for name in gs.dtype.names:
        a, b = gs[name][valid], os_[name][valid]
        same = (a == b) | (np.isnan(a) & np.isnan(b)) if a.dtype.kind == "f" else (a == b)
        assert np.all(same), f"{name}: {np.count_nonzero(~same)} of {same.size} differ"


def _tiny_scene(tris):
    tris = np.asarray(tris, np.float32).reshape(-1, 9)
    mats = np.array([[1.5, 0.5, 0.1, 0.2, 0.1, 1, 1e6, 0], [1.6, 0.7, 0.2, 0.3, 0.2, 1, 1e6, 0]], np.float32)
    n_mesh = 1 if len(tris) else 0
    return dict(materials=mats, starting_material=0, mesh_material_inside=np.array([1] * n_mesh, np.int32),
                mesh_material_outside=np.array([0] * n_mesh, np.int32), mesh_vascular=np.array([0] * n_mesh, np.int32),
                mesh_deltas=np.zeros((n_mesh, 3), np.float32), tri_offsets=np.array([0, len(tris)][: n_mesh + 1], np.int64),
                tri_vertices=tris, scaling=1.0, origin=np.zeros(3, np.float32), spacing=np.ones(3, np.float32))


@pytest.mark.parametrize("tris", [
    [],                                                                                   # no mesh at all: every ray misses
    [[-10.0, -5, -5, -10.0, 5, -5, -10.0, 0, 6]],                                         # one triangle (BVH without inner nodes)
    [[-10.0, -5, -5, -10.0, 5, -5, -10.0, 0, 6], [-8.0, -5, -5, -8.0, 5, -5, -8.0, 0, 6]],  # two triangles (single inner node)
])
def test_degenerate_scenes(api, O, tris):
    A = _tiny_scene(tris)
    pose = np.array([-13.5, 0, 0, 0, 0, -90], np.float32)
    gp = api.default_params(elements=64, samples=2, deterministic=1)
    op = O.default_params(elements=64, samples=2, deterministic=1)
    osc = O.OracleScene(A)
    with api.Simulator(A, gp) as sim:
        gs, gn = sim.cast_rays(pose, seed=3)
        rf = sim.simulate(pose[None, :], seed=3)[0]
    os_, on, _ = osc.cast_rays(op, pose[:3], pose[3:], seed=3, use_bvh=False)
    _segments_equal(gs, gn, os_, on)
    ref = osc.simulate_frame(op, pose[:3], pose[3:], seed=3)["rf"].T
    assert np.all(np.abs(rf - ref) <= _tol(ref) + 1e-12)
    if len(tris) == 0:
        assert np.all(gn == 1) and np.all(gs["tri_id"][:, :, 0] == -1)
    else:
        assert (gs["tri_id"] >= 0).any()


def test_config4_small_rough_nested_shells_path_for_path(api, O):
    """Config 4 in miniature (8 nested shells, 32 768 triangles, shininess 2, thickness 0.5, vascular
    shells): 10-bounce stochastic paths identical to the oracle's brute-force closest hit."""
    from mcray_tracing_b200 import assets
    A = assets.stress_scene_arrays(shells=8, nu=64, nv=32)
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]])
    gp = api.default_params(elements=64, samples=8)
    op = O.default_params(elements=64, samples=8)
    osc = O.OracleScene(A)
    with api.Simulator(A, gp) as sim:
        gs, gn = sim.cast_rays(pose, seed=99, frame=2)
        rf = sim.simulate(pose[None, :], seed=99, first_frame=2)[0]
    os_, on, _ = osc.cast_rays(op, pose[:3], pose[3:], seed=99, frame=2, use_bvh=False)
    _segments_equal(gs, gn, os_, on)
    assert on.max() == 10 and on.mean() > 4                       # paths really survive many bounces
    assert len(np.unique(gs["media_id"][gs["tri_id"] >= 0])) >= 3
    ref = osc.simulate_frame(op, pose[:3], pose[3:], seed=99, frame=2)["rf"].T
    assert np.all(np.abs(rf - ref) <= _tol(ref))


def test_config4_full_size_2m_triangles(api, O):
    """Config 4 at full size (2 097 152 triangles): closest hits and whole paths against the oracle's
    own BVH (validated against brute force on the small variant and on a subset here)."""
    from mcray_tracing_b200 import assets
    A = assets.stress_scene_arrays()
    assert len(A["tri_vertices"]) == 2097152
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]])
    osc = O.OracleScene(A)
    rng = np.random.default_rng(8)
    o = rng.normal(size=(3000, 3)) * 5.0
    d = rng.normal(size=(3000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    f, t = o.astype(np.float32), (o + d * 40.0).astype(np.float32)
    gp = api.default_params(elements=64, samples=8)
    op = O.default_params(elements=64, samples=8)
    with api.Simulator(A, gp) as sim:
        assert sim.info.n_triangles == 2097152
        tri, mesh, frac, pt, nr = sim.closest_hit(f, t)
        gs, gn = sim.cast_rays(pose, seed=99, frame=0)
    for i in range(len(f)):
        otri, omesh, o7 = osc.closest_hit(f[i], t[i], use_bvh=True)
        assert tri[i] == otri and frac[i] == o7[0], i
    for i in range(40):
        a = osc.closest_hit(f[i], t[i], use_bvh=False)
        assert a[0] == tri[i]
    os_, on, _ = osc.cast_rays(op, pose[:3], pose[3:], seed=99, frame=0, use_bvh=True)
    _segments_equal(gs, gn, os_, on)
    assert on.max() == 10


def test_config5_long_scanlines_and_large_psf(api, O, assets_dirs):
    """Config 5: a 17.6x finer axial grid (8333 RF samples per scanline; private accumulate columns of
    33 KB per path) and a 63 x 31 PSF: full frame within tolerance, PSF/envelope bit-exact."""
    path = assets_dirs["ircad11"] / "santi-liver.scene"
    A = O.load_scene_py(path)
    osc = O.OracleScene(A)
    kw = dict(elements=64, samples=8, axial_scale=17.6, psf_axial=63, psf_lateral=31)
    gp, op = api.default_params(**kw), O.default_params(**kw)
    with api.Simulator(path, gp) as sim:
        assert sim.rows == 8333
        pose = sim.start_pose
        rf = sim.simulate(pose[None, :], seed=7, first_frame=1)[0]
        st = sim.stats()
        rng = np.random.default_rng(5)
        img = rng.normal(size=(8192, 256)).astype(np.float32)        # [rows][cols], oracle layout
        ax = rng.normal(size=63).astype(np.float32); lat = rng.random(31).astype(np.float32)
        g = sim.postprocess(img.T, ax, lat)
    o = osc.simulate_frame(op, pose[:3], pose[3:], seed=7, frame=1)
    assert st.march_steps == o["steps"] and st.segments == o["tests"]
    ref = o["rf"].T
    assert np.all(np.abs(rf - ref) <= _tol(ref)), np.abs(rf - ref).max()
    assert np.array_equal(g.T, O.envelope(O.convolve(img, ax, lat)))


def test_traversal_options_do_not_change_results(api, O):
    """The host SAH tree and the two-stream pipeline only change
    scheduling / the acceleration structure: RF frames and segments stay bit-identical."""
    from mcray_tracing_b200 import assets
    A = assets.stress_scene_arrays(shells=8, nu=64, nv=32)
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]])
    poses = np.repeat(pose[None, :], 20, axis=0)
    with api.Simulator(A, api.default_params(elements=64, samples=8)) as sim:
        base = sim.simulate(poses, seed=5, first_frame=7)
        segs0, n0 = sim.cast_rays(pose, seed=5, frame=7)
        for opt, val in (("tail_merge", 0), ("tail_merge", 1), ("first_hit_dedup", 2), ("ordered_compaction", 2), ("group_histories", 1), ("bvh_builder", 1), ("bvh_builder", 2), ("overlap", 1),
                         ("post_tma", 0)):       # round 1's fused post kernel instead of the TMA-staged one
            sim.set_option(opt, val)
            assert np.array_equal(sim.simulate(poses, seed=5, first_frame=7), base), opt
            segs, n = sim.cast_rays(pose, seed=5, frame=7)
            assert np.array_equal(n, n0) and np.array_equal(segs["tri_id"], segs0["tri_id"]), opt
        sim.set_option("count_traversal", 1)
        sim.simulate(poses, seed=5, first_frame=7)
        st = sim.stats()
        assert st.bvh_node_visits > st.segments and st.bvh_triangle_tests > 0


def test_background_tree_optimisation(api, tmp_path, monkeypatch):
    """mcrt.h "bvh_optimise": the device LBVH serves a new or changed scene at once, a host thread builds the binned-SAH tree and a later
    compute call adopts it.  Frames never depend on which tree is in use; the optimised tree is visited with fewer node fetches; mesh
    updates fall back to the LBVH and the scene is optimised again once it has been left alone; $MCRT_BVH_CACHE keeps the tree."""
    from mcray_tracing_b200 import assets
    A = assets.stress_scene_arrays(shells=4, nu=128, nv=64)                   # 65 536 triangles (the optimiser starts at 32 768)
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]])
    poses = np.repeat(pose[None, :], 4, axis=0)
    kw = dict(elements=64, samples=4)

    def visits(sim):
        sim.set_option("count_traversal", 1)
        sim.simulate(poses, seed=3, first_frame=0)
        v = sim.stats().bvh_node_visits
        sim.set_option("count_traversal", 0)
        return v

    monkeypatch.setenv("MCRT_BVH_OPTIMISE", "0")
    with api.Simulator(A, api.default_params(**kw)) as plain:
        ref = plain.simulate(poses, seed=3, first_frame=0)
        plain.set_option("bvh_wait", 1)                                       # nothing to wait for
        assert plain.get_info().bvh_optimised == 0
        v_plain = visits(plain)
        plain.set_option("bvh_optimise", 1)                                   # switched on later
        plain.set_option("bvh_wait", 1)
        assert plain.get_info().bvh_optimised == 1
        assert np.array_equal(plain.simulate(poses, seed=3, first_frame=0), ref)
    monkeypatch.delenv("MCRT_BVH_OPTIMISE")
    with api.Simulator(A, api.default_params(**kw)) as sim:
        assert np.array_equal(sim.simulate(poses, seed=3, first_frame=0), ref)       # whichever tree is in use by now
        sim.set_option("bvh_wait", 1)
        assert sim.get_info().bvh_optimised == 1
        assert np.array_equal(sim.simulate(poses, seed=3, first_frame=0), ref)
        segs, nseg = sim.cast_rays(pose, seed=3, frame=0)
        v_opt = visits(sim)
        assert v_opt < 0.97 * v_plain, (v_opt, v_plain)
        # a mesh update: the LBVH of the new scene at once ...
        sim.set_mesh_origin(1, np.array([0.5, -0.3, 0.2], np.float32))
        moved = sim.simulate(poses, seed=3, first_frame=0)
        assert sim.get_info().bvh_optimised == 0 and not np.array_equal(moved, ref)
        # ... the optimiser again after eight calls on the unchanged scene
        for _ in range(9):
            assert np.array_equal(sim.simulate(poses, seed=3, first_frame=0), moved)
        sim.set_option("bvh_wait", 1)
        assert sim.get_info().bvh_optimised == 1
        assert np.array_equal(sim.simulate(poses, seed=3, first_frame=0), moved)
        sim.set_option("bvh_optimise", 0)                                     # and back to the plain LBVH
        assert sim.get_info().bvh_optimised == 0
        assert np.array_equal(sim.simulate(poses, seed=3, first_frame=0), moved)
    monkeypatch.setenv("MCRT_BVH_CACHE", str(tmp_path))
    with api.Simulator(A, api.default_params(**kw)) as first:
        first.set_option("bvh_wait", 1)
        assert first.get_info().bvh_optimised == 1 and len(list(tmp_path.glob("sah_*.bvh"))) == 1
    with api.Simulator(A, api.default_params(**kw)) as second:               # adopts the validated cached tree, no build
        second.set_option("bvh_wait", 1)
        assert second.get_info().bvh_optimised == 1
        assert np.array_equal(second.simulate(poses, seed=3, first_frame=0), ref)
        s2, n2 = second.cast_rays(pose, seed=3, frame=0)
        assert np.array_equal(n2, nseg) and np.array_equal(s2["tri_id"], segs["tri_id"])


def test_log_compression_option(api, O, assets_dirs):
    """The log compression the reference keeps commented out (rfimage.h:127-136), as an option: applied to
    the envelope image, so both rf_out and the scan-converted image change; bit-exact to the oracle."""
    path = assets_dirs["sphere"] / "sphere.scene"
    with api.Simulator(path, api.default_params(elements=512, samples=2)) as sim:
        poses = np.repeat(sim.start_pose[None, :], 3, axis=0)
        lin = sim.simulate(poses, seed=4, first_frame=0)
        sim.set_option("log_compress", 1)
        rf, scan = sim.simulate(poses, seed=4, first_frame=0, scan=True)
        sim.set_option("log_compress", 0)
        assert np.array_equal(sim.simulate(poses, seed=4, first_frame=0), lin)
    mx, my = O.create_mapping(O.default_params(samples=2))
    for i in range(len(poses)):
        ref = O.log_compress(lin[i].T)                       # oracle layout [rows][cols]
        assert np.array_equal(rf[i].T, ref, equal_nan=True)
        assert np.array_equal(scan[i], O.scan_convert(ref, mx, my), equal_nan=True)
    finite = rf[np.isfinite(rf)]
    assert finite.max() == 1.0 and np.isfinite(rf).mean() > 0.99


def test_streaming_driver_matches_batched_call(api, assets_dirs):
    """stream.FrameStreamer (pose stream in, frames out through pinned double buffers on one CUDA
    stream) returns exactly the frames of one batched mcrt_simulate call, in order."""
    from mcray_tracing_b200 import assets, stream
    path = assets_dirs["ircad11"] / "santi-liver.scene"
    poses = assets.sweep_poses(14)
    with api.Simulator(path, api.default_params(elements=128, samples=4)) as sim:
        ref, ref_scan = sim.simulate(poses, seed=3, first_frame=0, scan=True)
        got, got_scan, tickets = [], [], []
        fs = stream.FrameStreamer(sim, depth=3, frames_per_submit=2, scan=True, seed=3)
        fs.run(stream.sweep_pose_stream(poses, 2), lambda t, rf, sc: (tickets.append(t), got.append(rf.copy()), got_scan.append(sc.copy())))
        # the same with every submission of 7 poses simulated as 3 consecutive calls (sizes 2, 2, 3) whose copies start as soon as each
        # part is computed: the split must not show in the result
        got3, got3_scan = [], []
        fs3 = stream.FrameStreamer(sim, depth=2, frames_per_submit=7, scan=True, seed=3, sub_batches=3)
        fs3.run(stream.sweep_pose_stream(poses, 7), lambda t, rf, sc: (got3.append(rf.copy()), got3_scan.append(sc.copy())))
        with pytest.raises(ValueError):
            stream.FrameStreamer(sim, depth=2, frames_per_submit=2, sub_batches=3)
    assert tickets == list(range(1, 8))
    assert np.array_equal(np.concatenate(got), ref)
    assert np.array_equal(np.concatenate(got_scan), ref_scan)
    assert np.array_equal(np.concatenate(got3), ref) and np.array_equal(np.concatenate(got3_scan), ref_scan)


def test_entry_points_on_different_streams_are_ordered(api, assets_dirs):
    """One context = one workspace: an mcrt_simulate_async call still in flight on a user stream, followed at once by blocking
    entry points on the library's stream (simulate, cast_rays, transducer elements) and by an async call on a SECOND user
    stream, must not be corrupted by them, nor they by it (every entry point waits for the previous call's event)."""
    import torch
    from mcray_tracing_b200 import assets
    path = assets_dirs["ircad11"] / "santi-liver.scene"
    poses = assets.sweep_poses(48)
    with api.Simulator(path, api.default_params(elements=128, samples=8)) as sim:
        want_a = sim.simulate(poses[:32], seed=9, first_frame=0)
        want_b = sim.simulate(poses[32:], seed=9, first_frame=500)
        want_segs, want_n = sim.cast_rays(poses[40], seed=9, frame=77)
        dev = torch.device("cuda", sim.info.device)
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        for rep in range(3):
            a = torch.zeros((32, sim.cols, sim.rows), dtype=torch.float32, device=dev)
            c = torch.zeros((32, sim.cols, sim.rows), dtype=torch.float32, device=dev)
            sim.simulate_device(poses[:32], a.data_ptr(), seed=9, first_frame=0, stream=s1.cuda_stream, sync=False)      # in flight on s1
            got_b = sim.simulate(poses[32:], seed=9, first_frame=500)                                                  # library stream, blocking
            sim.simulate_device(poses[:32], c.data_ptr(), seed=9, first_frame=0, stream=s2.cuda_stream, sync=False)      # in flight on s2
            segs, n = sim.cast_rays(poses[40], seed=9, frame=77)
            sim.transducer_elements(poses[3])
            s1.synchronize(); s2.synchronize()
            assert np.array_equal(got_b, want_b)
            assert np.array_equal(a.cpu().numpy(), want_a) and np.array_equal(c.cpu().numpy(), want_a)
            assert np.array_equal(n, want_n) and np.array_equal(segs["tri_id"], want_segs["tri_id"])


@pytest.mark.parametrize("kw", [
    dict(elements=96, samples=4),                                                     # fused PSF+envelope kernel (465 rows)
    dict(elements=70, samples=2, axial_scale=5.0, psf_axial=15, psf_lateral=9),       # long-scanline kernels (2 k+ rows)
], ids=["short-fused", "long-split"])
def test_scanline_block_partition_equals_whole_frame(api, assets_dirs, kw):
    """SURVEY 8(e) secondary partition: ONE frame split into scanline blocks (what each of G ranks would run,
    mcrt_simulate_scanlines) and concatenated is bit-identical to mcrt_simulate of the whole frame, for every
    G incl. ragged splits and blocks narrower than the PSF halo.  The halo is re-traced, never exchanged."""
    from mcray_tracing_b200 import sweep
    path = assets_dirs["ircad11"] / "santi-liver-rough.scene"
    with api.Simulator(path, api.default_params(**kw)) as sim:
        pose = sim.start_pose
        whole = sim.simulate(pose[None, :], seed=11, first_frame=5)[0]
        E = sim.cols
        for world in (1, 2, 3, 8, 16):
            parts = []
            for r in range(world):
                b, e = sweep.shard_bounds(E, world, r)
                parts.append(sim.simulate_scanlines(pose, b, e - b, seed=11, frame=5))
            assert np.array_equal(np.concatenate(parts), whole), f"world={world}"
        # the batched entry point afterwards still gives the same frame (no state leaks from the block runs)
        assert np.array_equal(sim.simulate(pose[None, :], seed=11, first_frame=5)[0], whole)
        with pytest.raises(api.McrtError):
            sim.simulate_scanlines(pose, E - 2, 3)


@pytest.mark.parametrize("kw", [
    dict(elements=256, samples=16), dict(elements=64, samples=5), dict(elements=40, samples=1), dict(elements=33, samples=3),
    dict(elements=24, samples=128), dict(elements=32, samples=8, axial_scale=17.6, psf_axial=63, psf_lateral=31),
], ids=["256x16", "64x5", "40x1", "33x3", "24x128", "long-scanlines"])
def test_windowed_accumulate_equals_column_accumulate(api, assets_dirs, kw):
    """The row-window synchronous accumulate kernel (columns in shared memory, in-kernel sample reduction, no HBM
    columns) and the k_accumulate + k_reduce_samples pair give bit-identical frames and step counts; no echo ever
    arrives for an already finished window (late_echoes == 0)."""
    path = assets_dirs["ircad11"] / "santi-liver-rough.scene"
    from mcray_tracing_b200 import assets
    poses = assets.sweep_poses(5)
    with api.Simulator(path, api.default_params(**kw)) as sim:
        win = sim.simulate(poses, seed=17, first_frame=3)
        st_w = sim.stats()
        sim.set_option("accumulate_windowed", 0)
        col = sim.simulate(poses, seed=17, first_frame=3)
        st_c = sim.stats()
        sim.set_option("accumulate_windowed", 1)
        again = sim.simulate(poses, seed=17, first_frame=3)
    assert st_w.kernel_launches == st_c.kernel_launches - 1          # really two different kernels
    assert st_w.march_steps == st_c.march_steps and st_w.late_echoes == 0
    assert np.array_equal(win, col) and np.array_equal(again, win)
    assert np.count_nonzero(win) > 0.2 * win.size


def _read_png8(path):
    """Decode an 8-bit grayscale, non-interlaced, filter-0 PNG (what the CLI writes) with zlib; checks CRCs."""
    import struct
    import zlib
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    off, idat, shape = 8, b"", None
    while off < len(b):
        n, typ = struct.unpack(">I4s", b[off:off + 8])
        data = b[off + 8:off + 8 + n]
        assert struct.unpack(">I", b[off + 8 + n:off + 12 + n])[0] == (zlib.crc32(typ + data) & 0xffffffff)
        if typ == b"IHDR":
            w, h, depth, ctype, comp, flt, inter = struct.unpack(">IIBBBBB", data)
            assert (depth, ctype, comp, flt, inter) == (8, 0, 0, 0, 0)
            shape = (h, w)
        elif typ == b"IDAT":
            idat += data
        off += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(shape[0], shape[1] + 1)
    assert np.all(raw[:, 0] == 0)
    return raw[:, 1:]


def test_bmode_display_chain_and_cli_png(api, O, assets_dirs, tmp_path):
    """TGC + log compression to a dynamic range + scan conversion + 8-bit (mcrt_bmode), bit-exact to
    the oracle restatement (shared numerics); the CLI writes the same image as PNG (and prelog.png like rfimage.h:147)."""
    import subprocess
    path = assets_dirs["sphere"] / "sphere.scene"
    gp = api.default_params(elements=512, samples=2)
    op = O.default_params(elements=512, samples=2)
    mx, my = O.create_mapping(op)
    with api.Simulator(path, gp) as sim:
        poses = np.repeat(sim.start_pose[None, :], 2, axis=0)
        env, scan = sim.simulate(poses, seed=0, first_frame=0, scan=True)
        for gain, tgc, dr in ((0.0, 0.0, 60.0), (6.0, 0.7, 45.0), (-3.0, 2.0, 80.0)):
            cmp_, img8 = sim.bmode(env, gain_db=gain, tgc_db_per_cm=tgc, dynamic_range_db=dr)
            for i in range(len(poses)):
                ref = O.bmode(env[i].T, gp.depth_cm, gain, tgc, dr)                  # oracle layout [rows][cols]
                assert np.array_equal(cmp_[i].T, ref)
                ref8 = np.clip(np.rint(O.scan_convert(ref, mx, my) * np.float32(255.0)), 0, 255).astype(np.uint8)
                assert np.array_equal(img8[i], ref8)
            assert img8.max() > 200 and 0.02 < (img8 > 0).mean() < 1.0
        with pytest.raises(api.McrtError):
            sim.bmode(env, dynamic_range_db=0.0)
        _, img8 = sim.bmode(env[:1], gain_db=6.0, tgc_db_per_cm=0.7, dynamic_range_db=45.0)
    from pathlib import Path
    exe = Path(api.__file__).resolve().parent / "mattausch"
    out = subprocess.run([str(exe), str(path), "--samples", "2", "--seed", "0", "--out", str(tmp_path), "--png", "--bmode", "45", "--gain", "6",
                          "--tgc", "0.7"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "rf_image: 465, 512" in out.stdout, out.stdout + out.stderr
    assert np.array_equal(_read_png8(tmp_path / "bmode_0000.png"), img8[0])
    pre = _read_png8(tmp_path / "prelog.png")
    assert np.array_equal(pre, np.clip(np.rint(scan[0] * np.float32(255.0)), 0, 255).astype(np.uint8))


def test_moving_and_deforming_meshes_and_sah_cache(api, O, tmp_path, monkeypatch):
    """staged mesh updates (rigid move, vertex deformation) are applied by one device BVH rebuild and
    give exactly the frames / segments of a context created from the modified scene (and of the oracle); the host SAH
    tree is cached on disk under $MCRT_BVH_CACHE and a second build is a cache hit with identical results."""
    from mcray_tracing_b200 import assets
    A = assets.stress_scene_arrays(shells=4, nu=48, nv=24)
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]])
    poses = np.repeat(pose[None, :], 3, axis=0)
    scaling = np.float32(A["scaling"])
    A2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in A.items()}
    A2["mesh_deltas"][1] += np.array([0.6, -0.35, 0.25], np.float32) / scaling
    b, e = int(A["tri_offsets"][2]), int(A["tri_offsets"][3])
    v = A2["tri_vertices"][b:e].reshape(-1, 3)
    v *= (1.0 + 0.04 * np.sin(3.0 * v[:, :1] * scaling)).astype(np.float32)          # bumpy radial deformation of mesh 2
    kw = dict(elements=64, samples=4)
    with api.Simulator(A, api.default_params(**kw)) as sim:
        before = sim.simulate(poses, seed=8, first_frame=0)
        sim.set_mesh_origin(1, A2["mesh_deltas"][1] * scaling)
        sim.set_mesh_vertices(2, A2["tri_vertices"][b:e] * scaling)
        after = sim.simulate(poses, seed=8, first_frame=0)
        segs, nseg = sim.cast_rays(pose, seed=8, frame=0)
        with pytest.raises(api.McrtError):
            sim.set_mesh_vertices(2, A2["tri_vertices"][b:e - 1] * scaling)
        # SAH tree + disk cache
        monkeypatch.setenv("MCRT_BVH_CACHE", str(tmp_path))
        sim.set_option("bvh_builder", 1)
        assert sim.get_info().bvh_cache_hit == 0 and len(list(tmp_path.glob("sah_*.bvh"))) == 1
        sah = sim.simulate(poses, seed=8, first_frame=0)
        sim.set_option("bvh_builder", 1)
        assert sim.get_info().bvh_cache_hit == 1
        sah2 = sim.simulate(poses, seed=8, first_frame=0)
    with api.Simulator(A2, api.default_params(**kw)) as fresh:
        ref = fresh.simulate(poses, seed=8, first_frame=0)
        # the DEFAULT tree (device LBVH) is cached too: written by the first context created with a cache directory ...
        assert fresh.get_info().bvh_cache_hit == 0 and len(list(tmp_path.glob("lbvh_*.bvh"))) == 1
    with api.Simulator(A2, api.default_params(**kw)) as cached:                    # ... and loaded (after validation) by the next one
        assert cached.get_info().bvh_cache_hit == 1
        assert np.array_equal(cached.simulate(poses, seed=8, first_frame=0), ref)
    # a damaged cache file is detected and ignored: child reference out of range, then a swapped triangle
    f = next(tmp_path.glob("lbvh_*.bvh"))
    good = f.read_bytes()
    for damage in ("child", "slot", "truncated"):
        raw = bytearray(good)
        if damage == "child":
            raw[40 + 48:40 + 52] = (2 ** 30).to_bytes(4, "little")                # node 0, child[0]
        elif damage == "slot":
            n_nodes = int.from_bytes(good[8:16], "little")
            off = 40 + 64 * n_nodes
            raw[off:off + 4] = np.float32(123.5).tobytes()                         # first vertex coordinate of slot 0
        else:
            raw = raw[: len(raw) // 2]
        f.write_bytes(bytes(raw))
        with api.Simulator(A2, api.default_params(**kw)) as rebuilt:
            assert rebuilt.get_info().bvh_cache_hit == 0, damage
            assert np.array_equal(rebuilt.simulate(poses, seed=8, first_frame=0), ref), damage
    assert not np.array_equal(before, after)
    assert np.array_equal(after, ref) and np.array_equal(sah, ref) and np.array_equal(sah2, ref)
    osc = O.OracleScene(A2)
    os_, on, _ = osc.cast_rays(O.default_params(**kw), pose[:3], pose[3:], seed=8, frame=0, use_bvh=True)
    _segments_equal(segs, nseg, os_, on)


def test_elevational_psf_and_ray_fans(api, O, assets_dirs, tmp_path):
    """SURVEY 8(f) item 2, the elevation kernel the reference declares and never fills (psf.h:42,77): n_planes ray fans per
    frame offset along the elevation axis, combined with the elevation taps before the axial / lateral passes.  Taps and fan
    poses bit-equal to the oracle's restatement, frames within the RF tolerance, batching irrelevant, n_planes = 1 = off."""
    path = assets_dirs["ircad11"] / "santi-liver-rough.scene"
    A = O.load_scene_py(path)
    osc = O.OracleScene(A)
    kw = dict(elements=64, samples=3)
    op = O.default_params(**kw)
    with api.Simulator(path, api.default_params(**kw)) as sim:
        pose = sim.start_pose
        plain = sim.simulate(np.repeat(pose[None, :], 2, axis=0), seed=6, first_frame=4)
        taps, z = sim.set_elevation(5, 0.1)
        otaps, oz = O.elevation_taps(op, 5, 0.1)
        assert np.array_equal(taps, otaps) and np.array_equal(z, oz) and taps.max() <= 1.0 and len(set(taps.tolist())) > 2
        for j in range(5):
            pj = sim.elevation_pose(pose, j)
            assert np.array_equal(pj[:3], O.elevation_position(pose[:3], pose[3:], float(z[j]))) and np.array_equal(pj[3:], pose[3:])
        poses = np.stack([pose, sim.elevation_pose(pose, 0), pose])
        rf = sim.simulate(poses, seed=6, first_frame=4)
        assert sim.stats().poses == 15
        sim.set_option("max_batch_poses", 7)                        # one frame (5 fans) per batch
        assert np.array_equal(sim.simulate(poses, seed=6, first_frame=4), rf)
        sim.set_option("max_batch_poses", 256)
        for i in (0, 2):
            ref = O.simulate_frame_elevation(osc, op, poses[i][:3], poses[i][3:], seed=6, frame=4 + i, n_planes=5, var_z=0.1).T
            assert np.all(np.abs(rf[i] - ref) <= _tol(ref)), np.abs(rf[i] - ref).max()
        assert not np.array_equal(rf[0], plain[0])
        # the same chain (+ the depth-dependent lateral PSF) from the headless CLI: --elevation N VAR, --psf-depth FOCUS SPREAD
        sim.set_psf_depth_profile(4.0, 0.5)
        rf_cli_ref = sim.simulate(pose[None, :], seed=6, first_frame=0)[0]
        sim.set_psf_depth_profile(4.0, 0.0)
        sim.set_elevation(1)
        assert np.array_equal(sim.simulate(np.repeat(pose[None, :], 2, axis=0), seed=6, first_frame=4), plain)
        with pytest.raises(api.McrtError):
            sim.set_elevation(4)
    import subprocess
    exe = Path(api.__file__).resolve().parent / "mattausch"
    out = subprocess.run([str(exe), str(path), "--elements", "64", "--samples", "3", "--seed", "6", "--elevation", "5", "0.1", "--psf-depth", "4", "0.5",
                          "--out", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "rf_image: 465, 64" in out.stdout, out.stdout + out.stderr
    assert np.array_equal(np.fromfile(tmp_path / "rf_0000.f32", np.float32).reshape(64, 465), rf_cli_ref)
    bad = subprocess.run([str(exe), str(path), "--elements", "64", "--elevation", "4", "0.1", "--out", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert "The program found an error and will terminate." in bad.stdout and "elevation" in bad.stdout


@pytest.mark.parametrize("scene_name,det", [("santi-liver-rough.scene", 0), ("santi-liver.scene", 1)])
def test_ray_tree_mode_matches_oracle(api, O, assets_dirs, scene_name, det):
    """SURVEY 8(f) item 4 / the north star's ray *tree*: with option ray_tree both children of every boundary hit are
    followed (level-by-level wavefront kept in (path, node) order by warp-local compaction: no sort, no library kernel).  All
    segments of a frame, by (path, node), are bit-identical to the oracle's tree; the RF frame (a lane per segment through the
    windowed accumulate kernel, 32 segments per round in (level, path, node) order) is within the 1e-4 tolerance, identical
    from run to run and independent of the batch size; the tree contains the single-path segments' root and is strictly larger."""
    path = assets_dirs["ircad11"] / scene_name
    A = O.load_scene_py(path)
    osc = O.OracleScene(A)
    kw = dict(elements=48, samples=3, deterministic=det)
    gp, op = api.default_params(**kw), O.default_params(**kw)
    with api.Simulator(path, gp) as sim:
        pose = sim.start_pose
        single = sim.simulate(pose[None, :], seed=6, first_frame=2)[0]
        n_single = sim.stats().segments
        sim.set_option("ray_tree", 256)
        gs, gpath, gnode = sim.cast_rays_tree(pose, seed=6, frame=2)
        rf = sim.simulate(np.repeat(pose[None, :], 2, axis=0), seed=6, first_frame=2)
        st = sim.stats()
        rf_again = sim.simulate(np.repeat(pose[None, :], 5, axis=0), seed=6, first_frame=2)
        assert np.array_equal(rf_again[0], rf[0]) and np.array_equal(rf_again[1], rf[1])      # deterministic, batch-independent
        assert np.array_equal(rf_again[4], sim.simulate(pose[None, :], seed=6, first_frame=6)[0])
        sim.set_option("ray_tree", 2)                                  # far too small a budget: an error, not silent truncation
        with pytest.raises(api.McrtError):
            sim.simulate(pose[None, :], seed=6, first_frame=2)
        sim.set_option("ray_tree", 0)
        assert np.array_equal(sim.simulate(pose[None, :], seed=6, first_frame=2)[0], single)
    os_, opath, onode = osc.cast_rays_tree(op, pose[:3], pose[3:], seed=6, frame=2)
    assert len(gs) == len(os_) and np.array_equal(gpath, opath) and np.array_equal(gnode, onode)
    for name in gs.dtype.names:
        if name in ("mesh_id", "hit_fraction"):
            continue                                                   # not kept by the tree debug hook
        a, b = gs[name], os_[name]
        same = (a == b) | (np.isnan(a) & np.isnan(b)) if a.dtype.kind == "f" else (a == b)
        assert np.all(same), f"{name}: {np.count_nonzero(~same)} of {same.size} differ"
    assert len(gs) > n_single and np.count_nonzero(gnode == 1) == 48 * 3 and gnode.max() >= 8
    ref = osc.simulate_frame_tree(op, pose[:3], pose[3:], seed=6, frame=2)
    assert np.all(np.abs(rf[0] - ref["rf"].T) <= _tol(ref["rf"].T)), np.abs(rf[0] - ref["rf"].T).max()
    assert st.segments == len(gs) + len(osc.cast_rays_tree(op, pose[:3], pose[3:], seed=6, frame=3)[0])
    assert not np.array_equal(rf[0], single)


def test_cli_pose_sweep(api, assets_dirs, tmp_path):
    """`mattausch <scene> --poses FILE [--gpus G]`: the C++ host's probe sweep (one context + thread per GPU, contiguous
    pose blocks) writes exactly the frames of one batched mcrt_simulate call."""
    import subprocess
    from pathlib import Path
    import torch
    from mcray_tracing_b200 import assets
    path = assets_dirs["ircad11"] / "santi-liver.scene"
    poses = assets.sweep_poses(11)
    pf = tmp_path / "poses.txt"
    pf.write_text("# x y z ax ay az\n" + "\n".join(" ".join(repr(float(v)) for v in q) for q in poses) + "\n")
    with api.Simulator(path, api.default_params(elements=64, samples=4)) as sim:
        ref = sim.simulate(poses, seed=5, first_frame=0)
    exe = Path(api.__file__).resolve().parent / "mattausch"
    for gpus in sorted({1, min(2, torch.cuda.device_count())}):
        out = subprocess.run([str(exe), str(path), "--elements", "64", "--samples", "4", "--seed", "5", "--out", str(tmp_path), "--poses", str(pf),
                              "--gpus", str(gpus), "--batch", "4"], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "rf_image: 465, 64" in out.stdout, out.stdout + out.stderr
        got = np.fromfile(tmp_path / "sweep_rf.f32", np.float32).reshape(ref.shape)
        assert np.array_equal(got, ref), f"gpus={gpus}"


def test_depth_dependent_lateral_psf(api, O, assets_dirs):
    """the lateral PSF widens away from the focus (per-row lateral taps).  The tap table equals the
    oracle's bit for bit, the frame equals accumulate -> orc_convolve_depth -> envelope within the RF tolerance (the
    PSF/envelope arithmetic itself is bit-exact on identical input, checked through the scanline-block path), spread = 0
    restores the reference PSF exactly."""
    from mcray_tracing_b200 import sweep
    path = assets_dirs["ircad11"] / "santi-liver-rough.scene"
    A = O.load_scene_py(path)
    osc = O.OracleScene(A)
    kw = dict(elements=96, samples=4)
    gp, op = api.default_params(**kw), O.default_params(**kw)
    with api.Simulator(path, gp) as sim:
        pose = sim.start_pose
        plain = sim.simulate(pose[None, :], seed=3, first_frame=1)[0]
        tab = sim.set_psf_depth_profile(6.0, 1.5)
        deep = sim.simulate(np.repeat(pose[None, :], 2, axis=0), seed=3, first_frame=1)
        parts = [sim.simulate_scanlines(pose, *sweep.shard_bounds(sim.cols, 3, r)[:1], sweep.shard_bounds(sim.cols, 3, r)[1] - sweep.shard_bounds(sim.cols, 3, r)[0],
                                        seed=3, frame=1) for r in range(3)]
        assert sim.set_psf_depth_profile(6.0, 0.0) is None
        back = sim.simulate(pose[None, :], seed=3, first_frame=1)[0]
        with pytest.raises(api.McrtError):
            sim.set_psf_depth_profile(0.0, 1.0)
    assert np.array_equal(tab, O.psf_depth_table(op, 6.0, 1.5))
    assert np.array_equal(back, plain) and not np.array_equal(deep[0], plain)
    assert np.array_equal(np.concatenate(parts), deep[0])                      # scanline blocks use global rows / columns too
    os_, on, _ = osc.cast_rays(op, pose[:3], pose[3:], seed=3, frame=1)
    acc, _ = osc.accumulate(op, os_, on)
    ax, lat = O.psf_taps(op)
    ref = O.envelope(O.convolve_depth(acc, ax, tab)).T
    assert np.all(np.abs(deep[0] - ref) <= _tol(ref)), np.abs(deep[0] - ref).max()
    # at the focus row the table is the reference lateral kernel
    focus_row = int(round(6.0 / gp.depth_cm * tab.shape[1]))
    assert np.allclose(tab[:, focus_row], lat, rtol=0, atol=2e-3)
