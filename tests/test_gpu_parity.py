"""GPU parity tests: every stage of the CUDA path, called through the C ABI (libmcrt.so), against
the CPU oracle on the same seeded inputs.  Bars: bit-exact for hit ids / bounce counts / every
segment field / PSF / envelope / scan conversion; RF after accumulation within
|gpu - oracle| <= 1e-4 * max(|oracle|, 1e-3 * max|oracle image|)."""
import os
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _tol(ref):
    return 1e-4 * np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())


def _assert_segments_equal(gs, gn, os_, on):
    assert np.array_equal(gn, on), "bounce counts differ"
    D = gs.shape[-1]
    valid = np.arange(D)[None, None, :] < on[:, :, None]
    for name in gs.dtype.names:
        a, b = gs[name][valid], os_[name][valid]
        same = (a == b) | (np.isnan(a) & np.isnan(b)) if a.dtype.kind == "f" else (a == b)
        assert np.all(same), f"segment field {name!r} differs in {np.count_nonzero(~same)} of {same.size} entries"


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle_py
    return oracle_py


@pytest.fixture(scope="module")
def api(built):
    from mcray_tracing_b200 import api as a
    return a


@pytest.fixture(scope="module")
def sphere(api, O, assets_dirs):
    path = assets_dirs["sphere"] / "sphere.scene"
    A = O.load_scene_py(path)
    return path, A, O.OracleScene(A)


@pytest.fixture(scope="module")
def ircad(api, O, assets_dirs):
    path = assets_dirs["ircad11"] / "santi-liver.scene"
    A = O.load_scene_py(path)
    return path, A, O.OracleScene(A)


@pytest.fixture(scope="module")
def ircad_rough(api, O, assets_dirs):
    path = assets_dirs["ircad11"] / "santi-liver-rough.scene"
    A = O.load_scene_py(path)
    return path, A, O.OracleScene(A)


# ---------------------------------------------------------------------------------------------------
def test_numerics_contract_device_equals_host(api, O):
    """The shared transcendentals / Philox must be bit-identical on the device and on the host."""
    rng = np.random.default_rng(1)
    n = 200000
    cases = [
        (0, rng.uniform(-40, 10, n), None),
        (1, np.exp(rng.uniform(-25, 5, n)), None),
        (2, rng.uniform(-1, 1, n), rng.choice([0.001, 0.2, 0.5, 1.0, 2.0, 3.0], n)),
        (3, rng.uniform(0, 2 * np.pi, n), None),
        (4, rng.uniform(0, 2 * np.pi, n), None),
        (5, rng.integers(0, 2**31, n).astype(np.float64), rng.integers(0, 2**20, n).astype(np.float64)),
        (6, rng.uniform(1e-12, 1, n), rng.uniform(1e-9, 1, n)),
        (7, rng.uniform(-700, 700, n), None),
        (8, np.exp(rng.uniform(-700, 700, n)), None),
    ]
    for op, a, b in cases:
        dev = api.numerics_probe(op, a, b)
        host = O.numerics(op, a, b)
        same = (dev == host) | (np.isnan(dev) & np.isnan(host))
        assert np.all(same), f"op {op}: {np.count_nonzero(~same)} mismatches"


def test_scene_loader_matches_python_loader(api, O, ircad, sphere):
    for path, A, osc in (sphere, ircad):
        with api.Simulator(path, api.default_params(elements=64, samples=1)) as sim:
            tri, tm, org, mats = sim.scene_arrays()
        L = O.oracle()
        ref_tri = np.empty((osc.n_tri, 9), np.float32)
        ref_org = np.empty((len(A["mesh_deltas"]), 3), np.float32)
        L.orc_scene_get_local_vertices(osc.h, ref_tri.ctypes.data)
        L.orc_scene_get_mesh_origins(osc.h, ref_org.ctypes.data)
        assert np.array_equal(tri, ref_tri)
        assert np.array_equal(org, ref_org)
        assert np.array_equal(mats, A["materials"])
        assert np.array_equal(tm, np.repeat(np.arange(len(A["mesh_deltas"])), np.diff(A["tri_offsets"])).astype(np.int32))


def test_volume_upload_is_the_reference_volume(api, O, sphere):
    with api.Simulator(sphere[0], api.default_params(elements=64, samples=1)) as sim:
        v = sim.volume()
    assert np.array_equal(v, O.volume_raw())


@pytest.mark.parametrize("pose", [(-13.5, 0, 0, 0, 0, -90), (-17.5, 1, 5, 120, 0, -90), (-16, 3, 14, 45, 45, -90), (1.25, -4.5, 7.75, 33.3, -71.2, 190.0)])
def test_transducer_elements_bit_exact(api, O, sphere, pose):
    for E in (512, 256):
        with api.Simulator(sphere[0], api.default_params(elements=E, samples=1)) as sim:
            pos, d = sim.transducer_elements(pose)
        op, od = O.transducer_elements(O.default_params(elements=E, samples=1), pose[:3], pose[3:])
        assert np.array_equal(pos, op) and np.array_equal(d, od)


def _random_rays(rng, n, center, radius, length):
    o = center + rng.normal(size=(n, 3)) * radius
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), (o + d * length).astype(np.float32)


def test_closest_hit_vs_brute_force_sphere(api, O, sphere):
    """Device LBVH traversal vs the oracle's brute force over all 20 492 triangles."""
    path, A, osc = sphere
    rng = np.random.default_rng(3)
    f, t = _random_rays(rng, 3000, np.zeros(3), 6.0, 40.0)
    # plus astronomically long rays (GEL: `to` ~ 1e9 world units)
    f2, t2 = _random_rays(rng, 500, np.array([-13.5, 0, 0]), 0.5, 1.0e9)
    f, t = np.concatenate([f, f2]), np.concatenate([t, t2])
    with api.Simulator(path, api.default_params(elements=64, samples=1)) as sim:
        tri, mesh, frac, pt, nr = sim.closest_hit(f, t)
    for i in range(len(f)):
        otri, omesh, o7 = osc.closest_hit(f[i], t[i], use_bvh=False)
        assert tri[i] == otri and mesh[i] == omesh, f"ray {i}: gpu tri {tri[i]} oracle {otri}"
        assert frac[i] == o7[0]
        assert np.array_equal(pt[i], o7[1:4])
        if otri >= 0:
            assert np.array_equal(nr[i], o7[4:7])
    assert np.count_nonzero(tri >= 0) > 1000


def test_closest_hit_vs_oracle_bvh_ircad(api, O, ircad):
    path, A, osc = ircad
    rng = np.random.default_rng(4)
    f, t = _random_rays(rng, 20000, np.array([-2.0, -1.0, 7.0]), 8.0, 60.0)
    with api.Simulator(path, api.default_params(elements=64, samples=1)) as sim:
        tri, mesh, frac, pt, nr = sim.closest_hit(f, t)
    # the oracle BVH is itself validated against its brute force on a subset
    for i in range(0, 200):
        a = osc.closest_hit(f[i], t[i], use_bvh=False)
        b = osc.closest_hit(f[i], t[i], use_bvh=True)
        assert a[0] == b[0] and np.array_equal(a[2], b[2])
    bad = 0
    for i in range(len(f)):
        otri, omesh, o7 = osc.closest_hit(f[i], t[i], use_bvh=True)
        if not (tri[i] == otri and frac[i] == o7[0] and np.array_equal(pt[i], o7[1:4])):
            bad += 1
    assert bad == 0
    assert np.count_nonzero(tri >= 0) > 10000


def test_cast_rays_deterministic_sphere_c1(api, O, sphere):
    """BASELINE config 1: sphere scene, default transducer, deterministic mode: hit triangle ids,
    bounce counts and every segment field bit-exact (oracle closest hit = brute force)."""
    path, A, osc = sphere
    gp = api.default_params(samples=1, deterministic=1)
    op = O.default_params(samples=1, deterministic=1)
    with api.Simulator(path, gp) as sim:
        pose = sim.start_pose
        gs, gn = sim.cast_rays(pose, seed=11, frame=0)
    os_, on, tests = osc.cast_rays(op, pose[:3], pose[3:], seed=11, frame=0, use_bvh=False)
    _assert_segments_equal(gs, gn, os_, on)
    assert gn.sum() == tests and gn.max() >= 4


@pytest.mark.parametrize("scene_name,det", [("ircad", 1), ("ircad", 0), ("ircad_rough", 0)])
def test_cast_rays_ircad_path_for_path(api, O, request, scene_name, det):
    """ircad11 256 x 16: with the shared Philox keying the GPU wavefront is path-for-path identical to
    the oracle in deterministic AND stochastic mode (the stronger check of SURVEY.md section 8d)."""
    path, A, osc = request.getfixturevalue(scene_name)
    S = 1 if det else 16
    gp = api.default_params(elements=256, samples=S, deterministic=det)
    op = O.default_params(elements=256, samples=S, deterministic=det)
    with api.Simulator(path, gp) as sim:
        pose = sim.start_pose
        for frame in (0, 5):
            gs, gn = sim.cast_rays(pose, seed=1234, frame=frame)
            os_, on, _ = osc.cast_rays(op, pose[:3], pose[3:], seed=1234, frame=frame, use_bvh=True)
            _assert_segments_equal(gs, gn, os_, on)
    assert on.max() >= 5


def test_accumulate_matches_oracle(api, O, ircad_rough):
    path, A, osc = ircad_rough
    op = O.default_params(elements=256, samples=16)
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]])
    os_, on, _ = osc.cast_rays(op, pose[:3], pose[3:], seed=5, frame=1)
    ref, steps = osc.accumulate(op, os_, on)
    with api.Simulator(path, api.default_params(elements=256, samples=16)) as sim:
        rf = sim.accumulate(os_, on)
        assert sim.stats().march_steps == steps
    ref = ref.T
    assert np.all(np.abs(rf - ref) <= _tol(ref)), f"max abs err {np.abs(rf - ref).max()}"
    assert np.count_nonzero(ref) > 10000


@pytest.mark.parametrize("cols,rows,ka,kl", [(512, 465, 7, 13), (256, 465, 7, 13), (40, 64, 7, 13), (64, 300, 31, 15), (33, 31, 3, 5),
                                              # rows > 2048: the long-scanline kernels (register-blocked PSF, mask-based envelope)
                                              (40, 2500, 7, 13), (9, 2100, 3, 5), (70, 4099, 63, 31), (17, 2049, 9, 1),
                                              # tap counts on and around the 8-tap (axial) / 16-tap (lateral) groups of the compile-time kernels,
                                              # scanline counts that leave interior AND border groups of 16, the largest supported taps
                                              (150, 2100, 40, 17), (130, 2300, 64, 32), (50, 2200, 16, 16), (97, 2070, 8, 33), (35, 2060, 1, 48)])
def test_convolve_and_envelope_bit_exact(api, O, sphere, cols, rows, ka, kl):
    rng = np.random.default_rng(cols * 1000 + rows)
    img = rng.normal(size=(rows, cols)).astype(np.float32)          # oracle layout [rows][cols]
    img[rng.random(img.shape) < 0.3] = 0.0                          # plateaus / exact ties
    ax = rng.normal(size=ka).astype(np.float32)
    lat = rng.random(kl).astype(np.float32)
    with api.Simulator(sphere[0], api.default_params(elements=64, samples=1)) as sim:
        g_conv = sim.postprocess(img.T, ax, lat, convolve=True, envelope=False)
        g_env = sim.postprocess(img.T, ax, lat, convolve=False, envelope=True)
        g_both = sim.postprocess(img.T, ax, lat, convolve=True, envelope=True)
    o_conv = O.convolve(img, ax, lat)
    assert np.array_equal(g_conv.T, o_conv)
    assert np.array_equal(g_env.T, O.envelope(img))
    assert np.array_equal(g_both.T, O.envelope(o_conv))
    if rows > 2048:
        # the round-1 long-scanline kernels (run-time tap loops, global-memory envelope) give the same bits
        with api.Simulator(sphere[0], api.default_params(elements=64, samples=1)) as sim:
            sim.set_option("long_ct", 0)
            try:
                assert np.array_equal(sim.postprocess(img.T, ax, lat, convolve=True, envelope=True), g_both)
            finally:
                sim.set_option("long_ct", 1)


def test_envelope_long_peak_gaps(api, O, sphere):
    """Long scanlines whose peaks lie more than 2048 rows apart (the envelope's reciprocal-table division covers gaps up to 2048 rows
    and falls back to the IEEE division beyond), a scanline without any peak, and one with a peak in every other row."""
    rows, cols = 9000, 6
    r = np.arange(rows, dtype=np.float32)
    img = np.zeros((rows, cols), np.float32)
    img[:, 0] = r * 0.25 - 7.0                                           # monotone: no peak at all
    img[:, 1] = -np.abs(r - 4500.0)                                      # one peak in the middle: gaps of 4500 rows
    img[:, 2] = np.where(r < 3000, r, 6000.0 - r) + np.where(r > 8000, (r - 8000) * 5, 0)   # peak at 3000, then a valley: gap > 2048
    img[:, 3] = (np.arange(rows) % 2).astype(np.float32) * (1.0 + r / 100)               # a peak in every other row
    rng = np.random.default_rng(5)
    img[:, 4] = rng.normal(size=rows).astype(np.float32)
    img[::2500, 5] = 3.0                                                  # isolated spikes 2500 rows apart
    with api.Simulator(sphere[0], api.default_params(elements=64, samples=1)) as sim:
        g = sim.postprocess(img.T, np.ones(1, np.float32), np.ones(1, np.float32), convolve=False, envelope=True)
    assert np.array_equal(g.T, O.envelope(img))


def test_cast_rays_bit_exact_to_the_reference_loop_golden(api, O):
    """mcrt_trace_debug (k_first_hit / k_bounce / k_compact) against the committed segments of the REFERENCE'S OWN
    scene::cast_rays<5,512> (scene.cpp:50-183 compiled from the reference, tests/golden/reference_cast_rays.npz), on the
    impedance-matched scene where the reference's random draws have no effect: every field of every segment bit for bit."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_cast_rays", GOLD / "make_golden_cast_rays.py")
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    gold = np.load(GOLD / "reference_cast_rays.npz")
    A = mg.matched_scene()
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]]).astype(np.float32)
    with api.Simulator(A, api.default_params(elements=512, samples=5)) as sim:
        segs, nseg = sim.cast_rays(pose, seed=mg.SEED, frame=mg.FRAME)
    segs = segs.reshape(512, 5, 10); nseg = nseg.reshape(512, 5)
    assert np.array_equal(nseg, gold["nseg"])
    live = np.arange(10)[None, None, :] < nseg[..., None]
    for k, (lo, hi) in {"from": (0, 3), "to": (3, 6), "dir": (6, 9)}.items():
        assert np.array_equal(segs[k][live], gold["seg12"][..., lo:hi][live]), k
    for k, i in {"reflected_intensity": 9, "initial_intensity": 10, "attenuation": 11}.items():
        assert np.array_equal(segs[k][live], gold["seg12"][..., i][live]), k
    assert np.array_equal(segs["distance_traveled"][live], gold["dist_mm"][live])


def test_accumulate_against_the_reference_loop_golden(api, O, sphere):
    """mcrt_accumulate (k_accumulate_win) on the fixed segments against the committed image of the REFERENCE'S OWN accumulation
    loop (main.cpp:106-144 compiled from the reference, tests/golden/reference_accumulate_loop.npz): 1e-4 relative (the GPU
    sums the samples of a scanline per row, the reference path by path)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_accumulate", GOLD / "make_golden_accumulate.py")
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    gold = np.load(GOLD / "reference_accumulate_loop.npz")["rf"]
    segs, nseg, mats, p, osc = mg.fixed_segments()
    with api.Simulator(sphere[0], api.default_params(elements=512, samples=5)) as sim:
        rf = sim.accumulate(segs, nseg).T
        assert sim.stats().late_echoes == 0
    assert np.array_equal(rf != 0, gold != 0)
    assert np.all(np.abs(rf - gold) <= _tol(gold)), np.abs(rf - gold).max()


@pytest.mark.parametrize("cols,rows", [(256, 465), (37, 465), (50, 466), (8, 64), (45, 639), (300, 17)])
def test_tma_staged_post_kernel_equals_round1_kernel(api, O, sphere, cols, rows):
    """The TMA-staged fused post kernel (k_post_tma: one bulk copy per tile, register-blocked taps) against the oracle AND
    against round 1's k_post_fused (option post_tma = 0), incl. ragged tiles, pitch == rows and rows at the kernel's limits."""
    rng = np.random.default_rng(cols * 7 + rows)
    img = rng.normal(size=(rows, cols)).astype(np.float32)
    img[rng.random(img.shape) < 0.3] = 0.0
    with api.Simulator(sphere[0], api.default_params(elements=64, samples=1)) as sim:
        ax, lat = sim.psf_taps()                                     # the reference's 7 x 13 taps
        new = sim.postprocess(img.T, ax, lat)
        sim.set_option("post_tma", 0)
        old = sim.postprocess(img.T, ax, lat)
    assert np.array_equal(new, old)
    assert np.array_equal(new.T, O.envelope(O.convolve(img, ax, lat)))


def test_scan_convert_bit_exact(api, O, sphere):
    rng = np.random.default_rng(9)
    with api.Simulator(sphere[0], api.default_params(elements=512, samples=1)) as sim:
        img = rng.random((sim.rows, sim.cols)).astype(np.float32)
        g = sim.scan_convert(img.T)
    mx, my = O.create_mapping(O.default_params(samples=1))
    assert np.array_equal(g, O.scan_convert(img, mx, my))


@pytest.mark.parametrize("scene_name,E,S,det", [("sphere", 512, 1, 1), ("sphere", 512, 5, 0), ("ircad", 256, 16, 0), ("ircad_rough", 256, 16, 0)])
def test_full_frame_rf_within_tolerance(api, O, request, scene_name, E, S, det):
    path, A, osc = request.getfixturevalue(scene_name)
    gp = api.default_params(elements=E, samples=S, deterministic=det)
    op = O.default_params(elements=E, samples=S, deterministic=det)
    with api.Simulator(path, gp) as sim:
        pose = sim.start_pose
        rf, scan = sim.simulate(pose[None, :], seed=77, first_frame=3, scan=True)
        st = sim.stats()
    o = osc.simulate_frame(op, pose[:3], pose[3:], seed=77, frame=3, scan=True)
    ref = o["rf"].T
    assert st.segments == o["tests"] and st.march_steps == o["steps"]
    assert np.all(np.abs(rf[0] - ref) <= _tol(ref)), f"max abs err {np.abs(rf[0] - ref).max()} (max |ref| {np.abs(ref).max()})"
    sref = o["scan"]
    assert np.all(np.abs(scan[0] - sref) <= 1e-4 * np.maximum(np.abs(sref), 1e-3 * np.abs(sref).max()) + 1e-7)


def test_batched_poses_equal_single_pose_calls(api, ircad, O):
    """Sharding invariance: frames are pure functions of (scene, pose, seed, frame index), so any
    batching / split of a sweep is bit-identical (what makes N-GPU == 1-GPU)."""
    from mcray_tracing_b200 import assets
    path, A, osc = ircad
    poses = assets.sweep_poses(12)
    with api.Simulator(path, api.default_params(elements=128, samples=4)) as sim:
        all_rf = sim.simulate(poses, seed=9, first_frame=100)
        sim.set_option("max_batch_poses", 5)
        chunked = sim.simulate(poses, seed=9, first_frame=100)
        singles = np.stack([sim.simulate(poses[i:i + 1], seed=9, first_frame=100 + i)[0] for i in range(len(poses))])
        tail = sim.simulate(poses[7:], seed=9, first_frame=107)
        sim.set_option("use_graph", 0)
        nograph = sim.simulate(poses, seed=9, first_frame=100)
    assert np.array_equal(all_rf, chunked)
    assert np.array_equal(all_rf, singles)
    assert np.array_equal(all_rf[7:], tail)
    assert np.array_equal(all_rf, nograph)
    assert len({all_rf[i].tobytes() for i in range(len(poses))}) == len(poses)


def test_device_output_direct_and_strided(api, ircad):
    """Device-buffer calls: the post kernel writes the frames straight into the caller's buffer (option direct_out, default on) --
    same bits as the internal-image + copy path and as the host call, across batch splits, with and without the CUDA graph, with the
    elevational PSF; option rf_out_frame_stride = G puts frame i at slot i * G (the interleaved slots of a round-robin sweep) and leaves
    the slots in between untouched."""
    import torch
    from mcray_tracing_b200 import assets
    path, A, osc = ircad
    poses = assets.sweep_poses(11)
    with api.Simulator(path, api.default_params(elements=128, samples=4)) as sim:
        cols, rows = sim.cols, sim.rows
        host = sim.simulate(poses, seed=9, first_frame=40)

        def dev_call(n_slots=len(poses), fill=0.0):
            out = torch.full((n_slots, cols, rows), fill, dtype=torch.float32, device="cuda")
            sim.simulate_device(poses, out.data_ptr(), seed=9, first_frame=40)
            return out.cpu().numpy()

        assert np.array_equal(dev_call(), host)
        for opt, val in (("direct_out", 0), ("max_batch_poses", 4), ("use_graph", 0)):
            sim.set_option(opt, val)
            assert np.array_equal(dev_call(), host), opt
            sim.set_option(opt, {"direct_out": 1, "max_batch_poses": 256, "use_graph": 1}[opt])
        for direct in (1, 0):
            sim.set_option("direct_out", direct)
            sim.set_option("rf_out_frame_stride", 3)
            sim.set_option("max_batch_poses", 4)
            strided = dev_call(3 * len(poses), fill=-7.0)
            assert np.array_equal(strided[0::3], host)
            assert np.all(strided[1::3] == -7.0) and np.all(strided[2::3] == -7.0)
            sim.set_option("rf_out_frame_stride", 1)
            sim.set_option("max_batch_poses", 256)
        sim.set_option("direct_out", 1)
        with pytest.raises(api.McrtError):                                  # a stride needs a device buffer
            sim.set_option("rf_out_frame_stride", 2)
            sim.simulate(poses[:2], seed=9, first_frame=40)
        sim.set_option("rf_out_frame_stride", 1)
        sim.set_elevation(3, 0.1)
        e_host = sim.simulate(poses[:3], seed=9, first_frame=40)
        out = torch.zeros((3, cols, rows), dtype=torch.float32, device="cuda")
        sim.simulate_device(poses[:3], out.data_ptr(), seed=9, first_frame=40)
        assert np.array_equal(out.cpu().numpy(), e_host)


def test_rf_layout_cv_mat(api, sphere):
    gp0 = api.default_params(elements=128, samples=2)
    gp1 = api.default_params(elements=128, samples=2, rf_layout=1)
    with api.Simulator(sphere[0], gp0) as s0, api.Simulator(sphere[0], gp1) as s1:
        a = s0.simulate(s0.start_pose[None, :], seed=1)
        b = s1.simulate(s1.start_pose[None, :], seed=1)
    assert b.shape == (1, s1.rows, s1.cols)
    assert np.array_equal(a[0].T, b[0])


def test_ircad11_scene_without_shininess_loads(api, assets_dirs):
    """examples/ircad11/ircad11.scene lacks shininess/thickness and fails in the reference
    (scene.cpp:217-218); the drop-in defaults them (documented extension)."""
    with api.Simulator(assets_dirs["ircad11"] / "ircad11.scene", api.default_params(elements=64, samples=2)) as sim:
        rf = sim.simulate(sim.start_pose[None, :], seed=2)
    assert np.isfinite(rf).all() and np.abs(rf).max() > 0


def test_errors_are_codes_not_crashes(api, assets_dirs, tmp_path):
    with pytest.raises(api.McrtError) as e:
        api.Simulator(tmp_path / "missing.scene")
    assert e.value.code == api.MCRT_ERR_SCENE and "Error while loading scene" in e.value.message
    bad = tmp_path / "bad.scene"
    bad.write_text('{"materials": 3}')
    with pytest.raises(api.McrtError) as e:
        api.Simulator(bad)
    assert e.value.code == api.MCRT_ERR_SCENE
    with pytest.raises(api.McrtError) as e:
        api.Simulator(assets_dirs["sphere"] / "sphere.scene", api.default_params(psf_axial=8))
    assert e.value.code == api.MCRT_ERR_INVALID


@pytest.mark.parametrize("elements,samples", [(32, 4), (256, 16)], ids=["32x4", "256x16-baseline-config-2"])
def test_stochastic_mode_statistics_over_independent_seeds(api, O, ircad_rough, elements, samples):
    """North-star stochastic criterion, made independent of the shared Philox keying: the GPU simulates N = 256
    frames with seeds 0..255, the oracle N frames with DIFFERENT seeds (independent draws), rough ircad11 scene
    (thickness + shininess jitter active).  Per pixel with non-zero variance:
      * means agree within 4 standard errors, |m_g - m_o| <= 4 sqrt((v_g + v_o)/N), for >= 99 % of pixels
        (achieved oracle-vs-oracle on this scene: 100 %);
      * the variance ratio lies in [0.7, 1.4] for >= 85 % and in [1/3, 3] for >= 94 % of pixels.  SURVEY 8(d)
        asked for 99 % in [0.7, 1.4]; that is unreachable for ANY correct implementation because RF pixels are
        heavy-tailed (rare specular paths): the oracle against itself with two seed sets reaches 88.9 % / 96.2 %.
        So the bar that matters is the control: the GPU's fractions are within 2 points of the oracle-vs-oracle
        fractions computed in this same test.
    Run at a reduced acquisition (32 x 4) and at BASELINE.json's configuration 2 itself (256 scanlines x 16 samples, 256 seeds:
    measured fractions are recorded in DESIGN.md section 2)."""
    path, A, osc = ircad_rough
    N = 256
    kw = dict(elements=elements, samples=samples)
    op = O.default_params(**kw)
    O.oracle().orc_set_threads(min(16, os.cpu_count() or 1))
    with api.Simulator(path, api.default_params(**kw)) as sim:
        pose = sim.start_pose
        g = np.stack([sim.simulate(pose[None, :], seed=s, first_frame=0)[0] for s in range(N)])           # [N][cols][rows]
    o_same = np.stack([osc.simulate_frame(op, pose[:3], pose[3:], seed=s, frame=0)["rf"].T for s in range(0, N, 32)])
    o_b = np.stack([osc.simulate_frame(op, pose[:3], pose[3:], seed=100000 + s, frame=0)["rf"].T for s in range(N)])
    o_c = np.stack([osc.simulate_frame(op, pose[:3], pose[3:], seed=200000 + s, frame=0)["rf"].T for s in range(N)])
    O.oracle().orc_set_threads(1)
    # same seeds: path-for-path (the stronger check, kept alongside)
    for i, s in enumerate(range(0, N, 32)):
        assert np.all(np.abs(g[s] - o_same[i]) <= _tol(o_same[i]))

    def fractions(a, b):
        m1, m2, v1, v2 = a.mean(0), b.mean(0), a.var(0, ddof=1), b.var(0, ddof=1)
        nz = (v1 > 0) & (v2 > 0)
        z = np.abs(m1 - m2)[nz] / np.sqrt((v1 + v2)[nz] / N)
        r = v1[nz] / v2[nz]
        return float((z <= 4).mean()), float(((r >= 0.7) & (r <= 1.4)).mean()), float(((r >= 1 / 3) & (r <= 3)).mean()), int(nz.sum())

    gm, gv, gw, npx = fractions(g, o_b)
    cm, cv, cw, _ = fractions(o_c, o_b)
    print(f"stochastic statistics {elements}x{samples} over {N} seeds, {npx} pixels: GPU-vs-oracle mean {gm:.4f} var[0.7,1.4] {gv:.4f} var[1/3,3] {gw:.4f}; "
          f"oracle-vs-oracle control {cm:.4f} {cv:.4f} {cw:.4f}")
    assert npx > 0.9 * g[0].size
    assert gm >= 0.99 and gv >= 0.85 and gw >= 0.94
    assert gm >= cm - 0.02 and gv >= cv - 0.02 and gw >= cw - 0.02


@pytest.mark.parametrize("resolution_um", [145, 100, 333, 7])
def test_voxel_index_fma_division_is_exact(api, O, ircad_rough, resolution_um):
    """The accumulate kernel's 3-instruction voxel index (FMA division, image.cu) is enabled only after an exhaustive
    device check over all 2^32 float bit patterns for the context's resolution.  With it on and off the RF frame is
    bit-identical, and both match the oracle's IEEE divisions (volume.h:49-51), also for non-default resolutions."""
    path, A, osc = ircad_rough
    kw = dict(elements=64, samples=4, resolution_um=resolution_um)
    op = O.default_params(**kw)
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]])
    os_, on, _ = osc.cast_rays(op, pose[:3], pose[3:], seed=5, frame=1)
    ref, steps = osc.accumulate(op, os_, on)
    with api.Simulator(path, api.default_params(**kw)) as sim:
        validated = sim.get_info().voxel_fma_division
        on_rf = sim.accumulate(os_, on)
        assert sim.stats().march_steps == steps
        sim.set_option("voxel_fma_division", 0)
        assert sim.get_info().voxel_fma_division == 0
        off_rf = sim.accumulate(os_, on)
    if resolution_um == 145:
        assert validated == 1            # the reference's resolution (main.cpp:33) must take the fast path
    assert np.array_equal(on_rf, off_rf)
    assert np.all(np.abs(on_rf - ref.T) <= _tol(ref.T))


@pytest.mark.parametrize("kw", [
    dict(elements=1, samples=1), dict(elements=2, samples=1, max_depth=1), dict(elements=3, samples=7, max_depth=2),
    dict(elements=512, samples=5), dict(elements=17, samples=33), dict(elements=5, samples=16, max_depth=16),
], ids=["1x1", "2x1-depth1", "3x7-depth2", "reference-512x5", "17x33", "depth16"])
def test_edge_sizes_full_frame(api, O, ircad_rough, kw):
    """Smallest and odd acquisition sizes (one element, one sample, one bounce, sample counts that do not divide a warp,
    the maximum depth) and the reference's own 512 x 5: bounce counts and hit ids exact, RF within tolerance, through
    the default path (windowed accumulate with G = 32 / S scanlines per warp, or the CTA-wide variant for S > 32)."""
    path, A, osc = ircad_rough
    gp, op = api.default_params(**kw), O.default_params(**kw)
    with api.Simulator(path, gp) as sim:
        pose = sim.start_pose
        rf = sim.simulate(pose[None, :], seed=13, first_frame=4)[0]
        st = sim.stats()
        segs, nseg = sim.cast_rays(pose, seed=13, frame=4)
    o = osc.simulate_frame(op, pose[:3], pose[3:], seed=13, frame=4)
    os_, on, _ = osc.cast_rays(op, pose[:3], pose[3:], seed=13, frame=4)
    assert np.array_equal(nseg, on) and st.segments == o["tests"] and st.march_steps == o["steps"] and st.late_echoes == 0
    valid = np.arange(segs.shape[-1])[None, None, :] < on[:, :, None]
    assert np.array_equal(segs["tri_id"][valid], os_["tri_id"][valid])
    ref = o["rf"].T
    assert rf.shape == ref.shape and np.all(np.abs(rf - ref) <= _tol(ref)), np.abs(rf - ref).max()


def test_empty_and_extreme_calls(api, ircad):
    """Zero poses is a no-op; a frame index near 2^32 and a large seed work (the Philox counter takes frame modulo 2^32);
    ragged batches smaller than max_batch_poses reuse the workspace; bad arguments are error codes."""
    path, A, osc = ircad
    with api.Simulator(path, api.default_params(elements=32, samples=2)) as sim:
        pose = sim.start_pose
        empty = sim.simulate(np.zeros((0, 6), np.float32), seed=1, first_frame=0)
        assert empty.shape == (0, sim.cols, sim.rows)
        a = sim.simulate(np.repeat(pose[None, :], 3, axis=0), seed=2**63 + 5, first_frame=2**32 - 2)
        b = sim.simulate(pose[None, :], seed=2**63 + 5, first_frame=2**32 - 1)[0]
        c = sim.simulate(pose[None, :], seed=2**63 + 5, first_frame=2**32)[0]
        assert np.array_equal(a[1], b) and np.array_equal(a[2], c) and not np.array_equal(b, c)
        for n in (7, 2, 5, 1):
            assert sim.simulate(np.repeat(pose[None, :], n, axis=0), seed=3, first_frame=10).shape[0] == n
        with pytest.raises(api.McrtError):
            sim.set_option("no_such_option", 1)
        with pytest.raises(api.McrtError):
            sim.simulate_scanlines(pose, -1, 4)
    for bad in (dict(elements=0), dict(samples=0), dict(max_depth=0), dict(max_depth=17), dict(psf_lateral=4)):
        with pytest.raises(api.McrtError):
            api.Simulator(path, api.default_params(**bad))
