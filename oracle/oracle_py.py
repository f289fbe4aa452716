"""ctypes bindings of the CPU oracle (oracle/liboracle.so) and of the reference probe
(oracle/_ref/libref_probe.so), plus an independent pure-Python scene/OBJ loader.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under mcray_tracing_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFERENCE_ROOT = Path("/root/reference")

MAT_KEYS = ("impedance", "attenuation", "mu0", "mu1", "sigma", "specularity", "shininess", "thickness")


class OrcParams(C.Structure):
    _fields_ = [("elements", C.c_int32), ("samples", C.c_int32), ("max_depth", C.c_int32), ("frequency_mhz", C.c_float),
                ("radius_cm", C.c_double), ("fov_deg", C.c_double), ("depth_cm", C.c_double), ("speed_of_sound", C.c_uint32),
                ("resolution_um", C.c_uint32), ("psf_axial", C.c_int32), ("psf_lateral", C.c_int32), ("psf_var_x", C.c_float),
                ("psf_var_y", C.c_float), ("deterministic", C.c_int32), ("scan_rows", C.c_int32), ("scan_cols", C.c_int32),
                ("axial_scale", C.c_float), ("reserved", C.c_int32)]


class OrcDerived(C.Structure):
    _fields_ = [("axial_resolution_mm", C.c_double), ("axial_resolution_f", C.c_float), ("max_travel_time_us", C.c_double),
                ("max_travel_time_u", C.c_uint32), ("rf_axial_um", C.c_uint32), ("rows", C.c_int32), ("cols", C.c_int32),
                ("element_separation_mm", C.c_double), ("time_step_us", C.c_double), ("row_period_us", C.c_double)]


SEGMENT_DTYPE = np.dtype([("from", np.float32, 3), ("to", np.float32, 3), ("dir", np.float32, 3),
                          ("reflected_intensity", np.float32), ("initial_intensity", np.float32), ("attenuation", np.float32),
                          ("distance_traveled", np.float64), ("media_id", np.int32), ("tri_id", np.int32), ("mesh_id", np.int32),
                          ("hit_fraction", np.float32)], align=True)
assert SEGMENT_DTYPE.itemsize == 72


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def build_oracle(force: bool = False) -> Path:
    so = HERE / "liboracle.so"
    srcs = [HERE / "mcrt_oracle.cpp", HERE / "mcrt_oracle.h", HERE.parent / "mcray_tracing_b200/csrc/common/mcrt_numerics.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.check_call(["make", "-C", str(HERE), "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_ref_probe() -> Path | None:
    """Build oracle/_ref/libref_probe.so from the reference sources where they lie (only possible
    where /root/reference is mounted); returns None when neither the tree nor a prebuilt .so exists."""
    so = HERE / "_ref" / "libref_probe.so"
    if REFERENCE_ROOT.exists():
        subprocess.check_call(["make", "-C", str(HERE), "_ref"], stdout=subprocess.DEVNULL)
    return so if so.exists() else None


_ORACLE = None


def oracle():
    global _ORACLE
    if _ORACLE is None:
        L = C.CDLL(str(build_oracle()))
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_num_triangles.restype = C.c_int64
        L.orc_volume_get.restype = C.c_void_p
        L.orc_volume_raw.restype = C.c_void_p
        L.orc_volume_get_scattering.restype = C.c_float
        L.orc_volume_get_scattering.argtypes = [C.c_void_p] + [C.c_float] * 6
        L.orc_max_ray_length.restype = C.c_float
        L.orc_max_ray_length.argtypes = [C.c_float] * 3
        L.orc_travel.argtypes = [C.c_float, C.c_float, C.c_float, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_reflection_intensity.restype = C.c_float
        L.orc_reflection_intensity.argtypes = [C.c_float] * 5
        L.orc_reflected_intensity_eq8.restype = C.c_float
        L.orc_reflected_intensity_eq8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        L.orc_snells_law.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_random_unit_vector.argtypes = [C.c_void_p, C.c_float, C.c_double, C.c_double, C.c_void_p]
        L.orc_distance_in_mm.restype = C.c_double
        L.orc_distance_in_mm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_hit_boundary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_closest_hit.restype = C.c_int32
        L.orc_closest_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_cast_rays.restype = C.c_int64
        L.orc_cast_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_accumulate.restype = C.c_int64
        L.orc_accumulate.argtypes = [C.c_void_p] * 6
        L.orc_convolve.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.orc_envelope.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_log_compress.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_bmode.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_float, C.c_float, C.c_float]
        L.orc_create_mapping.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_scan_convert.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_simulate_frame.restype = C.c_int64
        L.orc_simulate_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_get_local_vertices.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_scene_get_mesh_origins.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_transducer_elements.argtypes = [C.c_void_p] * 5
        L.orc_psf_taps.argtypes = [C.c_void_p] * 3
        L.orc_get_max_threads.restype = C.c_int32
        L.orc_numerics.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _ORACLE = L
    return _ORACLE


_REF = None


def ref_probe():
    """The reference's own code (psf/volume/transducer/ray_physics/rf_image/objloader) or None."""
    global _REF
    if _REF is None:
        so = build_ref_probe()
        if so is None:
            return None
        L = C.CDLL(str(so))
        L.ref_volume_raw.restype = C.c_void_p
        L.ref_volume_get_scattering.restype = C.c_float
        L.ref_volume_get_scattering.argtypes = [C.c_float] * 6
        L.ref_max_ray_length.restype = C.c_float
        L.ref_max_ray_length.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.ref_travel.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.ref_reflection_intensity.restype = C.c_float
        L.ref_reflection_intensity.argtypes = [C.c_float] * 5
        L.ref_reflected_intensity_eq8.restype = C.c_float
        L.ref_reflected_intensity_eq8.argtypes = [C.c_void_p] * 4
        L.ref_snells_law.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.ref_random_unit_vector.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.ref_power_cosine_variate.restype = C.c_float
        L.ref_power_cosine_variate.argtypes = [C.c_int]
        L.ref_world_create.restype = C.c_void_p
        L.ref_world_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_world_destroy.argtypes = [C.c_void_p]
        L.ref_world_start.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.ref_world_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_world_set_intensity.argtypes = [C.c_void_p, C.c_float]
        L.ref_rf_add_echo.argtypes = [C.c_uint, C.c_float, C.c_double]
        L.ref_rf_set.argtypes = [C.c_void_p]
        L.ref_rf_get.argtypes = [C.c_void_p]
        L.ref_rf_mapping.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_load_obj.restype = C.c_int
        L.ref_load_obj.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.ref_transducer_elements.argtypes = [C.c_void_p] * 4
        L.ref_psf_taps.argtypes = [C.c_void_p] * 2
        L.ref_constants.argtypes = [C.c_void_p]
        _REF = L
    return _REF


# -------------------------------------------------------------------------------------------------
# independent Python scene loading (json + a small OBJ reader with tinyobj 0.9.5 semantics:
# tiny_obj_loader.cpp:97-109 fixIndex, :272-285 polygon -> triangle fan, faces kept in file order,
# objloader.h:23-139 un-welds to a triangle soup)
# -------------------------------------------------------------------------------------------------
def load_obj_py(path) -> np.ndarray:
    verts: list[tuple[float, float, float]] = []
    tris: list[int] = []
    with open(path, "r") as fh:
        for line in fh:
            tok = line.lstrip(" \t")
            if tok.startswith("v ") or tok.startswith("v\t"):
                p = tok[2:].split()
                verts.append((float(p[0]), float(p[1]), float(p[2])))
            elif tok.startswith("f ") or tok.startswith("f\t"):
                idx = []
                for t in tok[2:].split():
                    i = int(t.split("/")[0])
                    n = len(verts)
                    idx.append(i - 1 if i > 0 else (0 if i == 0 else n + i))
                for k in range(2, len(idx)):
                    tris.extend((idx[0], idx[k - 1], idx[k]))
    v = np.asarray(verts, dtype=np.float64).astype(np.float32).reshape(-1, 3)
    t = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
    return v[t].reshape(-1, 9).copy()


def load_scene_py(scene_path) -> dict:
    """Scene JSON (scene.cpp:185-247, main.cpp:65-69) -> arrays.  Extensions as documented in
    DESIGN.md: shininess / thickness default to 1e6 / 0; a non-existent workingDirectory falls back
    to the scene file's directory."""
    scene_path = Path(scene_path)
    cfg = json.loads(scene_path.read_text())
    wd = cfg.get("workingDirectory", "")
    if not wd or not os.path.isdir(wd):
        wd = str(scene_path.parent) + "/"
    names = [m["name"] for m in cfg["materials"]]
    mats = np.array([[m["impedance"], m["attenuation"], m["mu0"], m["mu1"], m["sigma"], m["specularity"],
                      m.get("shininess", 1000000), m.get("thickness", 0.0)] for m in cfg["materials"]], dtype=np.float32)
    soups, offs = [], [0]
    for me in cfg["meshes"]:
        s = load_obj_py(wd + me["file"])
        soups.append(s)
        offs.append(offs[-1] + len(s))
    return dict(
        materials=mats, material_names=names, starting_material=names.index(cfg["startingMaterial"]),
        mesh_material_inside=np.array([names.index(m["material"]) for m in cfg["meshes"]], np.int32),
        mesh_material_outside=np.array([names.index(m["outsideMaterial"]) for m in cfg["meshes"]], np.int32),
        mesh_vascular=np.array([int(bool(m["vascular"])) for m in cfg["meshes"]], np.int32),
        mesh_deltas=np.array([m["deltas"] for m in cfg["meshes"]], np.float32).reshape(-1, 3),
        tri_offsets=np.array(offs, np.int64),
        tri_vertices=np.concatenate(soups, axis=0) if soups else np.zeros((0, 9), np.float32),
        scaling=float(cfg["scaling"]), origin=np.array(cfg["origin"], np.float32), spacing=np.array(cfg["spacing"], np.float32),
        transducer_position=np.array(cfg["transducerPosition"], np.float32), transducer_angles=np.array(cfg["transducerAngles"], np.float32))


class OracleScene:
    """Owns an orc_scene built from scene arrays (as returned by load_scene_py / assets.stress_scene_arrays)."""

    def __init__(self, arrays: dict):
        self.a = arrays
        L = oracle()
        n_mat, n_mesh = len(arrays["materials"]), len(arrays["mesh_material_inside"])
        self._keep = [np.ascontiguousarray(arrays[k]) for k in ("materials", "mesh_material_inside", "mesh_material_outside", "mesh_vascular",
                                                                  "mesh_deltas", "tri_offsets", "tri_vertices", "origin", "spacing")]
        m, mi, mo, mv, md, to, tv, org, sp = self._keep
        L.orc_scene_create.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        self.h = L.orc_scene_create(n_mat, _p(m), int(arrays["starting_material"]), n_mesh, _p(mi), _p(mo), _p(mv), _p(md), _p(to), _p(tv),
                                    float(arrays["scaling"]), _p(org), _p(sp))
        self.n_tri = int(L.orc_scene_num_triangles(C.c_void_p(self.h)))

    def __del__(self):
        try:
            if self.h:
                oracle().orc_scene_destroy(C.c_void_p(self.h))
                self.h = None
        except Exception:
            pass

    def closest_hit(self, frm, to, use_bvh=True):
        f = np.ascontiguousarray(frm, np.float32)
        t = np.ascontiguousarray(to, np.float32)
        out = np.zeros(7, np.float32)
        mesh = C.c_int32(-1)
        tri = oracle().orc_closest_hit(self.h, _p(f), _p(t), int(use_bvh), _p(out), C.byref(mesh))
        return tri, mesh.value, out

    def cast_rays(self, params: OrcParams, pos, angles, seed=0, frame=0, use_bvh=True):
        E, S, D = params.elements, params.samples, params.max_depth
        segs = np.zeros((E, S, D), dtype=SEGMENT_DTYPE)
        nseg = np.zeros((E, S), dtype=np.int32)
        pos = np.ascontiguousarray(pos, np.float32)
        ang = np.ascontiguousarray(angles, np.float32)
        tests = oracle().orc_cast_rays(self.h, C.byref(params), _p(pos), _p(ang), int(seed), int(frame), int(use_bvh), _p(segs), _p(nseg))
        return segs, nseg, int(tests)

    def accumulate(self, params: OrcParams, segs, nseg):
        d = derive(params)
        rf = np.zeros((d.rows, d.cols), np.float32)
        steps = oracle().orc_accumulate(self.h, C.byref(params), oracle().orc_volume_get(), _p(segs), _p(nseg), _p(rf))
        return rf, int(steps)

    def cast_rays_tree(self, params: OrcParams, pos, angles, seed=0, frame=0, use_bvh=True, capacity=None):
        """Ray-tree mode: (segments[n], path[n], node[n]) in (path, node) order."""
        cap = int(capacity or params.elements * params.samples * 256)
        segs = np.zeros(cap, SEGMENT_DTYPE); path = np.zeros(cap, np.int32); node = np.zeros(cap, np.int32)
        pos = np.ascontiguousarray(pos, np.float32); ang = np.ascontiguousarray(angles, np.float32)
        L = oracle()
        L.orc_cast_rays_tree.restype = C.c_int64
        L.orc_cast_rays_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        n = L.orc_cast_rays_tree(self.h, C.byref(params), _p(pos), _p(ang), int(seed), int(frame), int(use_bvh), cap, _p(segs), _p(path), _p(node))
        if n < 0:
            raise RuntimeError("cast_rays_tree: capacity too small")
        return segs[:n].copy(), path[:n].copy(), node[:n].copy()

    def accumulate_flat(self, params: OrcParams, segs, path):
        d = derive(params)
        rf = np.zeros((d.rows, d.cols), np.float32)
        L = oracle()
        L.orc_accumulate_flat.restype = C.c_int64
        L.orc_accumulate_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        segs = np.ascontiguousarray(segs); path = np.ascontiguousarray(path, np.int32)
        steps = L.orc_accumulate_flat(self.h, C.byref(params), L.orc_volume_get(), _p(segs), _p(path), len(segs), _p(rf))
        return rf, int(steps)

    def simulate_frame_tree(self, params: OrcParams, pos, angles, seed=0, frame=0):
        """Ray-tree frame: tree cast, accumulation in (path, node) order, convolve, envelope -> rf [rows][cols]."""
        segs, path, node = self.cast_rays_tree(params, pos, angles, seed, frame)
        rf, steps = self.accumulate_flat(params, segs, path)
        ax, lat = psf_taps(params)
        return dict(rf=envelope(convolve(rf, ax, lat)), segments=len(segs), steps=steps)

    def simulate_frame(self, params: OrcParams, pos, angles, seed=0, frame=0, scan=False):
        d = derive(params)
        rf = np.zeros((d.rows, d.cols), np.float32)
        sc = np.zeros((params.scan_rows, params.scan_cols), np.float32) if scan else None
        st = np.zeros(4, np.float64)
        steps = C.c_int64(0)
        pos = np.ascontiguousarray(pos, np.float32)
        ang = np.ascontiguousarray(angles, np.float32)
        tests = oracle().orc_simulate_frame(self.h, C.byref(params), _p(pos), _p(ang), int(seed), int(frame), _p(rf),
                                            _p(sc) if scan else None, _p(st), C.byref(steps))
        return dict(rf=rf, scan=sc, tests=int(tests), steps=int(steps.value), stage_seconds=st)


def default_params(**kw) -> OrcParams:
    p = OrcParams()
    oracle().orc_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def derive(p: OrcParams) -> OrcDerived:
    d = OrcDerived()
    oracle().orc_derive(C.byref(p), C.byref(d))
    return d


def psf_taps(p: OrcParams):
    ax = np.zeros(p.psf_axial, np.float32)
    lat = np.zeros(p.psf_lateral, np.float32)
    oracle().orc_psf_taps(C.byref(p), _p(ax), _p(lat))
    return ax, lat


def transducer_elements(p: OrcParams, pos, angles):
    pos = np.ascontiguousarray(pos, np.float32)
    ang = np.ascontiguousarray(angles, np.float32)
    op = np.zeros((p.elements, 3), np.float32)
    od = np.zeros((p.elements, 3), np.float32)
    oracle().orc_transducer_elements(C.byref(p), _p(pos), _p(ang), _p(op), _p(od))
    return op, od


def convolve(rf, ax, lat):
    out = np.ascontiguousarray(rf, np.float32).copy()
    ax = np.ascontiguousarray(ax, np.float32)
    lat = np.ascontiguousarray(lat, np.float32)
    oracle().orc_convolve(_p(out), out.shape[0], out.shape[1], _p(ax), len(ax), _p(lat), len(lat))
    return out


def psf_depth_table(p: OrcParams, focus_cm, spread):
    rows = derive(p).rows
    tab = np.zeros((p.psf_lateral, rows), np.float32)
    L = oracle()
    L.orc_psf_depth_table.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p]
    L.orc_psf_depth_table(C.byref(p), rows, float(focus_cm), float(spread), _p(tab))
    return tab


def convolve_depth(rf, ax, lat_by_row):
    out = np.ascontiguousarray(rf, np.float32).copy()
    ax = np.ascontiguousarray(ax, np.float32); tab = np.ascontiguousarray(lat_by_row, np.float32)
    L = oracle()
    L.orc_convolve_depth.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    L.orc_convolve_depth(_p(out), out.shape[0], out.shape[1], _p(ax), len(ax), _p(tab), tab.shape[0])
    return out


def envelope(rf):
    out = np.ascontiguousarray(rf, np.float32).copy()
    oracle().orc_envelope(_p(out), out.shape[0], out.shape[1])
    return out


def log_compress(rf):
    out = np.ascontiguousarray(rf, np.float32).copy()
    oracle().orc_log_compress(_p(out), out.shape[0], out.shape[1])
    return out


def bmode(rf, depth_cm, gain_db, tgc_db_per_cm, dynamic_range_db):
    """B-mode display chain on a [rows][cols] envelope image -> [0, 1] (the 8-bit image is np.rint(scan_convert(.) * 255))."""
    out = np.ascontiguousarray(rf, np.float32).copy()
    oracle().orc_bmode(_p(out), out.shape[0], out.shape[1], float(depth_cm), float(gain_db), float(tgc_db_per_cm), float(dynamic_range_db))
    return out


def create_mapping(p: OrcParams):
    mx = np.zeros((p.scan_rows, p.scan_cols), np.float32)
    my = np.zeros((p.scan_rows, p.scan_cols), np.float32)
    oracle().orc_create_mapping(C.byref(p), _p(mx), _p(my))
    return mx, my


def scan_convert(rf, mx, my):
    rf = np.ascontiguousarray(rf, np.float32)
    out = np.zeros(mx.shape, np.float32)
    oracle().orc_scan_convert(_p(rf), rf.shape[0], rf.shape[1], _p(mx), _p(my), mx.shape[0], mx.shape[1], _p(out))
    return out


def volume_raw() -> np.ndarray:
    L = oracle()
    ptr = L.orc_volume_raw(C.c_void_p(L.orc_volume_get()))
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(256, 256, 256, 2))


def elevation_taps(params, n: int, var_z: float):
    taps = np.zeros(n, np.float32); z = np.zeros(n, np.float32)
    L = oracle()
    L.orc_elevation_taps.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]
    L.orc_elevation_taps(C.byref(params), int(n), float(var_z), _p(taps), _p(z))
    return taps, z


def elevation_position(pos, angles, z_mm: float) -> np.ndarray:
    pos = np.ascontiguousarray(pos, np.float32); ang = np.ascontiguousarray(angles, np.float32)
    out = np.zeros(3, np.float32)
    L = oracle()
    L.orc_elevation_position.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    L.orc_elevation_position(_p(pos), _p(ang), float(z_mm), _p(out))
    return out


def simulate_frame_elevation(osc, params, pos, angles, seed: int, frame: int, n_planes: int, var_z: float) -> np.ndarray:
    """Frame with the elevational PSF: n_planes ray fans offset along the elevation axis, Philox frame counter frame * n + j,
    raw RF images combined with the elevation taps (j ascending, separate multiply and add), then convolve + envelope.
    -> rf [rows][cols]."""
    taps, z = elevation_taps(params, n_planes, var_z)
    acc = None
    for j in range(n_planes):
        pj = elevation_position(pos, angles, float(z[j]))
        segs, nseg, _ = osc.cast_rays(params, pj, angles, seed=seed, frame=frame * n_planes + j)
        rf, _ = osc.accumulate(params, segs, nseg)
        if acc is None:
            acc = np.zeros_like(rf)
        acc = (acc + (rf * np.float32(taps[j])).astype(np.float32)).astype(np.float32)
    ax, lat = psf_taps(params)
    return envelope(convolve(acc, ax, lat))


def ref_accumulate_loop(R, segs, nseg, materials) -> np.ndarray:
    """The reference's own accumulation loop (main.cpp:106-144, compiled into the reference probe) on oracle-layout segments
    [512][5][D] -> raw RF image [465][512].  The segment's medium is copied at emission (SURVEY B-1)."""
    E, S, D = segs.shape
    assert (E, S) == (512, 5), "the reference's sizes are compile-time constants (main.cpp:26-27)"
    f = np.zeros((E, S, D, 12), np.float32)
    f[..., 0:3] = segs["from"]; f[..., 3:6] = segs["to"]; f[..., 6:9] = segs["dir"]
    f[..., 9] = segs["reflected_intensity"]; f[..., 10] = segs["initial_intensity"]; f[..., 11] = segs["attenuation"]
    dist = np.ascontiguousarray(segs["distance_traveled"], np.float64)
    mats = np.asarray(materials, np.float32).reshape(-1, 8)
    mid = np.clip(segs["media_id"], 0, len(mats) - 1)
    media3 = np.ascontiguousarray(mats[mid][..., [2, 3, 4]], np.float32)         # mu0, mu1, sigma
    n = np.ascontiguousarray(nseg, np.int32)
    out = np.zeros((465, 512), np.float32)
    R.ref_accumulate_loop(_p(f), _p(dist), _p(media3), _p(n), C.c_int(D), _p(out))
    return out


def numerics(op: int, a, b=None) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b if b is not None else np.zeros_like(a), np.float64)
    out = np.empty_like(a)
    oracle().orc_numerics(int(op), a.size, _p(a), _p(b), _p(out))
    return out
