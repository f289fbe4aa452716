"""ctypes binding of libmcrt.so (include/mcrt.h) -- the thin Python mirror used by the parity
tests, bench.py and the torch.distributed sweep driver.  All compute happens inside the library's
CUDA kernels; this module only marshals pointers.  It fails loudly if the library is missing."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("MCRT_LIB_PATH", PKG_DIR / "libmcrt.so"))   # override: A/B builds during development

MCRT_OK = 0
MCRT_ERR_INVALID, MCRT_ERR_SCENE, MCRT_ERR_CUDA, MCRT_ERR_NOMEM = -1, -2, -3, -4


class McrtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[mcrt {code}] {message}")
        self.code = code
        self.message = message


class Params(C.Structure):
    """mcrt_params (main.cpp:23-37,54 made runtime)."""
    _fields_ = [("elements", C.c_int32), ("samples", C.c_int32), ("max_depth", C.c_int32), ("frequency_mhz", C.c_float),
                ("radius_cm", C.c_double), ("fov_deg", C.c_double), ("depth_cm", C.c_double), ("speed_of_sound", C.c_uint32),
                ("resolution_um", C.c_uint32), ("psf_axial", C.c_int32), ("psf_lateral", C.c_int32), ("psf_var_x", C.c_float),
                ("psf_var_y", C.c_float), ("deterministic", C.c_int32), ("scan_rows", C.c_int32), ("scan_cols", C.c_int32),
                ("axial_scale", C.c_float), ("rf_layout", C.c_int32)]


class Pose(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("angles_deg", C.c_float * 3)]


class SceneArrays(C.Structure):
    _fields_ = [("n_materials", C.c_int32), ("materials8", C.c_void_p), ("starting_material", C.c_int32), ("n_meshes", C.c_int32),
                ("mesh_material_inside", C.c_void_p), ("mesh_material_outside", C.c_void_p), ("mesh_vascular", C.c_void_p),
                ("mesh_deltas", C.c_void_p), ("tri_offsets", C.c_void_p), ("tri_vertices", C.c_void_p), ("scaling", C.c_float),
                ("origin", C.c_float * 3), ("spacing", C.c_float * 3)]


class Info(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("scan_rows", C.c_int32), ("scan_cols", C.c_int32), ("n_materials", C.c_int32),
                ("n_meshes", C.c_int32), ("n_triangles", C.c_int64), ("n_bvh_nodes", C.c_int64), ("device", C.c_int32),
                ("sm_count", C.c_int32), ("start_pose", C.c_float * 6), ("axial_resolution_mm", C.c_double), ("time_step_us", C.c_double),
                ("row_period_us", C.c_double), ("max_travel_time_us", C.c_double), ("voxel_fma_division", C.c_int32), ("bvh_cache_hit", C.c_int32),
                ("bvh_optimised", C.c_int32), ("reserved0", C.c_int32)]


class BmodeParams(C.Structure):
    _fields_ = [("gain_db", C.c_float), ("tgc_db_per_cm", C.c_float), ("dynamic_range_db", C.c_float), ("reserved0", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("poses", C.c_int64), ("segments", C.c_int64), ("march_steps", C.c_int64), ("kernel_launches", C.c_int64),
                ("ms_total", C.c_float), ("ms_trace", C.c_float), ("ms_accumulate", C.c_float), ("ms_post", C.c_float),
                ("bvh_node_visits", C.c_int64), ("bvh_triangle_tests", C.c_int64), ("late_echoes", C.c_int64)]


SEGMENT_DTYPE = np.dtype([("from", np.float32, 3), ("to", np.float32, 3), ("dir", np.float32, 3),
                          ("reflected_intensity", np.float32), ("initial_intensity", np.float32), ("attenuation", np.float32),
                          ("distance_traveled", np.float64), ("media_id", np.int32), ("tri_id", np.int32), ("mesh_id", np.int32),
                          ("hit_fraction", np.float32)], align=True)
assert SEGMENT_DTYPE.itemsize == 72

EXPORTS = ["mcrt_default_params", "mcrt_create", "mcrt_create_from_arrays", "mcrt_destroy", "mcrt_last_error", "mcrt_get_info",
           "mcrt_get_stats", "mcrt_set_option", "mcrt_simulate", "mcrt_simulate_async", "mcrt_trace_debug", "mcrt_closest_hit",
           "mcrt_transducer_elements", "mcrt_accumulate", "mcrt_postprocess", "mcrt_scan_convert", "mcrt_get_psf_taps",
           "mcrt_get_scene", "mcrt_get_volume", "mcrt_numerics_probe", "mcrt_load_obj", "mcrt_scene_probe", "mcrt_host_tables", "mcrt_simulate_scanlines", "mcrt_bmode", "mcrt_set_mesh_origin", "mcrt_set_mesh_vertices", "mcrt_trace_tree_debug", "mcrt_device_alloc", "mcrt_set_psf_depth_profile",
           "mcrt_device_free", "mcrt_ipc_export", "mcrt_ipc_open", "mcrt_ipc_close", "mcrt_copy_async", "mcrt_copy2d_async", "mcrt_set_elevation", "mcrt_elevation_pose",
           "mcrt_host_build_sah"]


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile libmcrt.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = PKG_DIR / "csrc"
    cmd = ["make", "-C", str(csrc), "all"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        L.mcrt_last_error.restype = C.c_char_p
        vp = C.c_void_p
        L.mcrt_default_params.argtypes = [vp]
        L.mcrt_create.argtypes = [C.c_char_p, vp, C.c_int, vp]
        L.mcrt_create_from_arrays.argtypes = [vp, vp, C.c_int, vp]
        L.mcrt_destroy.argtypes = [vp]
        L.mcrt_destroy.restype = None
        L.mcrt_get_info.argtypes = [vp, vp]
        L.mcrt_get_stats.argtypes = [vp, vp]
        L.mcrt_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
        L.mcrt_simulate.argtypes = [vp, vp, C.c_int32, C.c_uint64, C.c_uint64, vp, vp]
        L.mcrt_simulate_async.argtypes = [vp, vp, C.c_int32, C.c_uint64, C.c_uint64, vp, vp, vp]
        L.mcrt_trace_debug.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, vp]
        L.mcrt_simulate_scanlines.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, vp]
        L.mcrt_bmode.argtypes = [vp, vp, C.c_int32, vp, vp, vp]
        L.mcrt_trace_tree_debug.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_int64, vp, vp, vp, vp]
        L.mcrt_set_psf_depth_profile.argtypes = [vp, C.c_float, C.c_float, vp]
        L.mcrt_device_alloc.argtypes = [C.c_int, C.c_size_t, vp]
        L.mcrt_device_free.argtypes = [C.c_int, vp]
        L.mcrt_ipc_export.argtypes = [C.c_int, vp, vp]
        L.mcrt_ipc_open.argtypes = [C.c_int, vp, vp]
        L.mcrt_ipc_close.argtypes = [C.c_int, vp]
        L.mcrt_copy_async.argtypes = [C.c_int, vp, vp, C.c_size_t, vp]
        L.mcrt_copy2d_async.argtypes = [C.c_int, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp]
        L.mcrt_set_mesh_origin.argtypes = [vp, C.c_int32, vp]
        L.mcrt_set_mesh_vertices.argtypes = [vp, C.c_int32, vp, C.c_int64]
        L.mcrt_closest_hit.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp]
        L.mcrt_transducer_elements.argtypes = [vp, vp, vp, vp]
        L.mcrt_accumulate.argtypes = [vp, vp, vp, vp]
        L.mcrt_postprocess.argtypes = [vp, vp, C.c_int32, C.c_int32, vp, C.c_int32, vp, C.c_int32, C.c_int32, vp]
        L.mcrt_scan_convert.argtypes = [vp, vp, vp]
        L.mcrt_get_psf_taps.argtypes = [vp, vp, vp]
        L.mcrt_get_scene.argtypes = [vp, vp, vp, vp, vp]
        L.mcrt_get_volume.argtypes = [vp, vp]
        L.mcrt_numerics_probe.argtypes = [C.c_int, C.c_int32, C.c_int64, vp, vp, vp]
        L.mcrt_load_obj.argtypes = [C.c_char_p, vp, C.c_int64, vp]
        L.mcrt_scene_probe.argtypes = [C.c_char_p, vp, vp, vp, vp]
        L.mcrt_host_tables.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.mcrt_host_build_sah.argtypes = [vp, vp, C.c_int64, vp, C.c_int32, C.c_int32, vp, vp, vp]
        _LIB = L
    return _LIB


def _check(rc: int):
    if rc != MCRT_OK:
        raise McrtError(rc, lib().mcrt_last_error().decode("utf-8", "replace"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---- peer-memory plumbing (mcrt_device_alloc / mcrt_ipc_* / mcrt_copy_async) -----------------------------------------
def device_alloc(device: int, nbytes: int) -> int:
    p = C.c_void_p()
    _check(lib().mcrt_device_alloc(int(device), int(nbytes), C.byref(p)))
    return int(p.value)


def device_free(device: int, ptr: int):
    _check(lib().mcrt_device_free(int(device), C.c_void_p(ptr)))


def ipc_export(device: int, ptr: int) -> bytes:
    h = (C.c_ubyte * 64)()
    _check(lib().mcrt_ipc_export(int(device), C.c_void_p(ptr), h))
    return bytes(h)


def ipc_open(device: int, handle: bytes) -> int:
    h = (C.c_ubyte * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    _check(lib().mcrt_ipc_open(int(device), h, C.byref(p)))
    return int(p.value)


def ipc_close(device: int, ptr: int):
    _check(lib().mcrt_ipc_close(int(device), C.c_void_p(ptr)))


def copy_async(device: int, dst: int, src: int, nbytes: int, stream: int = 0):
    _check(lib().mcrt_copy_async(int(device), C.c_void_p(dst), C.c_void_p(src), int(nbytes), C.c_void_p(stream) if stream else None))


def copy2d_async(device: int, dst: int, dst_pitch: int, src: int, src_pitch: int, width_bytes: int, height: int, stream: int = 0):
    _check(lib().mcrt_copy2d_async(int(device), C.c_void_p(dst), int(dst_pitch), C.c_void_p(src), int(src_pitch), int(width_bytes), int(height),
                                   C.c_void_p(stream) if stream else None))


def default_params(**kw) -> Params:
    p = Params()
    _check(lib().mcrt_default_params(C.byref(p)))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def make_poses(poses) -> np.ndarray:
    """float32 [n, 6] = (pos xyz, angles xyz deg): the memory layout of mcrt_pose[n]."""
    a = np.ascontiguousarray(np.asarray(poses, dtype=np.float32).reshape(-1, 6))
    return a


class Simulator:
    """One scene + acquisition geometry on one GPU: the objects main.cpp:52-81 builds
    (volume, psf, rf_image, transducer, scene) behind one handle."""

    def __init__(self, scene, params: Params | None = None, device: int = 0):
        L = lib()
        self.params = params if params is not None else default_params()
        h = C.c_void_p()
        if isinstance(scene, (str, os.PathLike)):
            rc = L.mcrt_create(str(scene).encode(), C.byref(self.params), int(device), C.byref(h))
        else:
            # explicit dtypes: the C side reads raw memory, and the arrays must outlive the call (pointers are taken from _keep only)
            dt = {"materials": np.float32, "mesh_material_inside": np.int32, "mesh_material_outside": np.int32, "mesh_vascular": np.int32,
                  "mesh_deltas": np.float32, "tri_offsets": np.int64, "tri_vertices": np.float32}
            self._keep = {k: np.ascontiguousarray(scene[k], dtype=t) for k, t in dt.items()}
            k = self._keep
            sa = SceneArrays()
            sa.n_materials = len(k["materials"].reshape(-1, 8)); sa.materials8 = _p(k["materials"])
            sa.starting_material = int(scene["starting_material"]); sa.n_meshes = len(k["mesh_material_inside"])
            sa.mesh_material_inside = _p(k["mesh_material_inside"]); sa.mesh_material_outside = _p(k["mesh_material_outside"])
            sa.mesh_vascular = _p(k["mesh_vascular"]); sa.mesh_deltas = _p(k["mesh_deltas"]); sa.tri_offsets = _p(k["tri_offsets"])
            sa.tri_vertices = _p(k["tri_vertices"]); sa.scaling = float(scene["scaling"])
            for i in range(3):
                sa.origin[i] = float(scene["origin"][i]); sa.spacing[i] = float(scene["spacing"][i])
            rc = L.mcrt_create_from_arrays(C.byref(sa), C.byref(self.params), int(device), C.byref(h))
        _check(rc)
        self.h = h
        self.info = Info()
        _check(L.mcrt_get_info(self.h, C.byref(self.info)))
        self.rows, self.cols = self.info.rows, self.info.cols

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            lib().mcrt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def start_pose(self) -> np.ndarray:
        return np.array(list(self.info.start_pose), dtype=np.float32)

    def set_option(self, name: str, value: int):
        _check(lib().mcrt_set_option(self.h, name.encode(), int(value)))
        self._options = getattr(self, "_options", {})
        self._options[name] = int(value)

    def stats(self) -> Stats:
        s = Stats()
        _check(lib().mcrt_get_stats(self.h, C.byref(s)))
        return s

    # -- the frame loop body (main.cpp:102-148) --------------------------------------------------
    def rf_shape(self, n_poses: int):
        return (n_poses, self.rows, self.cols) if self.params.rf_layout == 1 else (n_poses, self.cols, self.rows)

    def simulate(self, poses, seed: int = 0, first_frame: int = 0, scan: bool = False, rf_out: np.ndarray | None = None,
                 scan_out: np.ndarray | None = None):
        """Host-buffer call: poses -> RF (+ scan-converted image).  Copies are inside the call."""
        P = make_poses(poses)
        n = len(P)
        rf = rf_out if rf_out is not None else np.empty(self.rf_shape(n), np.float32)
        sc = scan_out if scan_out is not None else (np.empty((n, self.info.scan_rows, self.info.scan_cols), np.float32) if scan else None)
        _check(lib().mcrt_simulate(self.h, _p(P), n, int(seed), int(first_frame), _p(rf), _p(sc)))
        return (rf, sc) if (scan or scan_out is not None) else rf

    def simulate_device(self, poses, rf_ptr: int, seed: int = 0, first_frame: int = 0, scan_ptr: int | None = None,
                        stream: int | None = None, sync: bool = True):
        """Device-buffer call: rf_ptr / scan_ptr are raw device addresses (e.g. torch.Tensor.data_ptr())."""
        P = make_poses(poses)
        n = len(P)
        if sync and stream is None:
            _check(lib().mcrt_simulate(self.h, _p(P), n, int(seed), int(first_frame), C.c_void_p(rf_ptr),
                                       C.c_void_p(scan_ptr) if scan_ptr else None))
        else:
            _check(lib().mcrt_simulate_async(self.h, _p(P), n, int(seed), int(first_frame), C.c_void_p(rf_ptr),
                                             C.c_void_p(scan_ptr) if scan_ptr else None, C.c_void_p(stream) if stream else None))

    def simulate_scanlines(self, pose, first_element: int, n_elements: int, seed: int = 0, frame: int = 0, rf_ptr: int | None = None):
        """One pose, scanline block [first_element, first_element + n_elements) -> [n_elements, rows] (host array, or
        written to the device address rf_ptr)."""
        P = make_poses(pose)
        if rf_ptr is not None:
            _check(lib().mcrt_simulate_scanlines(self.h, _p(P), int(seed), int(frame), int(first_element), int(n_elements), C.c_void_p(rf_ptr)))
            return None
        out = np.empty((n_elements, self.rows), np.float32)
        _check(lib().mcrt_simulate_scanlines(self.h, _p(P), int(seed), int(frame), int(first_element), int(n_elements), _p(out)))
        return out

    def bmode(self, env, gain_db: float = 0.0, tgc_db_per_cm: float = 0.0, dynamic_range_db: float = 60.0):
        """B-mode display chain (mcrt_bmode) on envelope images [n][cols][rows] (rf_layout 0):
        returns (compressed float [n][cols][rows] in [0,1], 8-bit scan-converted [n][scan_rows][scan_cols])."""
        e = np.ascontiguousarray(env, np.float32)
        if e.ndim == 2:
            e = e[None]
        n = e.shape[0]
        bp = BmodeParams(gain_db, tgc_db_per_cm, dynamic_range_db, 0.0)
        cmp_ = np.empty_like(e)
        img8 = np.empty((n, self.info.scan_rows, self.info.scan_cols), np.uint8)
        _check(lib().mcrt_bmode(self.h, _p(e), n, C.byref(bp), _p(cmp_), _p(img8)))
        return cmp_, img8

    def cast_rays_tree(self, pose, seed: int = 0, frame: int = 0, capacity: int | None = None):
        """Ray-tree mode parity hook (option ray_tree > 0): (segments[n], path[n], node[n]) sorted by (path, node)."""
        P = make_poses(pose)
        budget = getattr(self, "_options", {}).get("ray_tree", 0)
        cap = int(capacity or self.params.elements * self.params.samples * max(budget, 1))
        segs = np.zeros(cap, SEGMENT_DTYPE); path = np.zeros(cap, np.int32); node = np.zeros(cap, np.int32)
        n = C.c_int64(0)
        _check(lib().mcrt_trace_tree_debug(self.h, _p(P), int(seed), int(frame), cap, _p(segs), _p(path), _p(node), C.byref(n)))
        return segs[: n.value].copy(), path[: n.value].copy(), node[: n.value].copy()

    def set_elevation(self, n_planes: int, var_z: float = 0.1):
        """Elevational PSF (mcrt_set_elevation): n_planes ray fans per frame (odd; 1 = off).  Returns (taps, z_mm)."""
        taps = np.zeros(max(n_planes, 1), np.float32); z = np.zeros(max(n_planes, 1), np.float32)
        _check(lib().mcrt_set_elevation(self.h, int(n_planes), C.c_float(var_z), _p(taps), _p(z)))
        return taps, z

    def elevation_pose(self, pose, plane: int) -> np.ndarray:
        P = make_poses(pose)
        out = make_poses(np.zeros(6, np.float32))
        _check(lib().mcrt_elevation_pose(self.h, _p(P), int(plane), _p(out)))
        return out[0].copy()

    def set_psf_depth_profile(self, focus_cm: float, spread: float):
        """Depth-dependent lateral PSF (spread = 0: off); returns the taps [psf_lateral][rows] (None when off)."""
        tab = np.zeros((self.params.psf_lateral, self.rows), np.float32) if spread > 0 else None
        _check(lib().mcrt_set_psf_depth_profile(self.h, float(focus_cm), float(spread), _p(tab)))
        return tab

    def set_mesh_origin(self, mesh: int, origin3):
        """Move a mesh: new body origin in world cm; applied (one BVH rebuild) at the next compute call."""
        o = np.ascontiguousarray(origin3, np.float32).reshape(3)
        _check(lib().mcrt_set_mesh_origin(self.h, int(mesh), _p(o)))

    def set_mesh_vertices(self, mesh: int, tri_local9):
        """Deform a mesh: its triangles in the local frame, [n][9], same count and order as loaded."""
        v = np.ascontiguousarray(tri_local9, np.float32).reshape(-1, 9)
        _check(lib().mcrt_set_mesh_vertices(self.h, int(mesh), _p(v), len(v)))

    def get_info(self) -> Info:
        """mcrt_get_info, re-read (options can change what it reports)."""
        info = Info()
        _check(lib().mcrt_get_info(self.h, C.byref(info)))
        return info

    # -- parity hooks ----------------------------------------------------------------------------
    def cast_rays(self, pose, seed: int = 0, frame: int = 0):
        """scene::cast_rays (scene.cpp:50-183): segments[E][S][D], n_segments[E][S]."""
        P = make_poses(pose)
        E, S, D = self.params.elements, self.params.samples, self.params.max_depth
        segs = np.zeros((E, S, D), dtype=SEGMENT_DTYPE)
        nseg = np.zeros((E, S), dtype=np.int32)
        _check(lib().mcrt_trace_debug(self.h, _p(P), int(seed), int(frame), _p(segs), _p(nseg)))
        return segs, nseg

    def closest_hit(self, frm, to):
        f = np.ascontiguousarray(np.asarray(frm, np.float32).reshape(-1, 3))
        t = np.ascontiguousarray(np.asarray(to, np.float32).reshape(-1, 3))
        n = len(f)
        tri = np.empty(n, np.int32); mesh = np.empty(n, np.int32); frac = np.empty(n, np.float32)
        pt = np.empty((n, 3), np.float32); nr = np.empty((n, 3), np.float32)
        _check(lib().mcrt_closest_hit(self.h, n, _p(f), _p(t), _p(tri), _p(mesh), _p(frac), _p(pt), _p(nr)))
        return tri, mesh, frac, pt, nr

    def transducer_elements(self, pose):
        P = make_poses(pose)
        pos = np.empty((self.params.elements, 3), np.float32); d = np.empty((self.params.elements, 3), np.float32)
        _check(lib().mcrt_transducer_elements(self.h, _p(P), _p(pos), _p(d)))
        return pos, d

    def accumulate(self, segs, nseg):
        """main.cpp:106-144 -> raw RF, scanline-major [cols][rows]."""
        segs = np.ascontiguousarray(segs); nseg = np.ascontiguousarray(nseg, np.int32)
        rf = np.empty((self.cols, self.rows), np.float32)
        _check(lib().mcrt_accumulate(self.h, _p(segs), _p(nseg), _p(rf)))
        return rf

    def postprocess(self, rf_cols_rows, axial=None, lateral=None, convolve=True, envelope=True):
        rf = np.ascontiguousarray(rf_cols_rows, np.float32)
        cols, rows = rf.shape
        if axial is None or lateral is None:
            axial, lateral = self.psf_taps()
        ax = np.ascontiguousarray(axial, np.float32); lat = np.ascontiguousarray(lateral, np.float32)
        out = np.empty_like(rf)
        _check(lib().mcrt_postprocess(self.h, _p(rf), cols, rows, _p(ax), len(ax), _p(lat), len(lat), (1 if convolve else 0) | (2 if envelope else 0), _p(out)))
        return out

    def scan_convert(self, rf_cols_rows):
        rf = np.ascontiguousarray(rf_cols_rows, np.float32)
        out = np.empty((self.info.scan_rows, self.info.scan_cols), np.float32)
        _check(lib().mcrt_scan_convert(self.h, _p(rf), _p(out)))
        return out

    def psf_taps(self):
        ax = np.empty(self.params.psf_axial, np.float32); lat = np.empty(self.params.psf_lateral, np.float32)
        _check(lib().mcrt_get_psf_taps(self.h, _p(ax), _p(lat)))
        return ax, lat

    def scene_arrays(self):
        n = int(self.info.n_triangles)
        tri = np.empty((n, 9), np.float32); tm = np.empty(n, np.int32)
        org = np.empty((self.info.n_meshes, 3), np.float32); mats = np.empty((self.info.n_materials, 8), np.float32)
        _check(lib().mcrt_get_scene(self.h, _p(tri), _p(tm), _p(org), _p(mats)))
        return tri, tm, org, mats

    def volume(self):
        v = np.empty((256, 256, 256, 2), np.float32)
        _check(lib().mcrt_get_volume(self.h, _p(v)))
        return v


def numerics_probe(op: int, a, b=None, device: int = 0) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b if b is not None else np.zeros_like(a), np.float64)
    out = np.empty_like(a)
    _check(lib().mcrt_numerics_probe(int(device), int(op), a.size, _p(a), _p(b), _p(out)))
    return out


# ---- host-only entry points (no GPU) --------------------------------------------------------------
def load_obj(path) -> np.ndarray:
    """objloader.h:154-161: un-welded triangle soup float32 [n, 9]."""
    n = C.c_int64(0)
    _check(lib().mcrt_load_obj(str(path).encode(), None, 0, C.byref(n)))
    out = np.empty((n.value, 9), np.float32)
    _check(lib().mcrt_load_obj(str(path).encode(), _p(out), n.value, C.byref(n)))
    return out


def host_build_sah(tri_local, tri_mesh, mesh_origins, threads: int = 0):
    """The host binned-SAH tree of the background optimisation (no GPU needed): (nodes [n - 1, 16] as float32 whose words 12, 13 are the
    int32 child references, slot_triangle [n], max_depth)."""
    tri_local = np.ascontiguousarray(tri_local, np.float32).reshape(-1, 9)
    tri_mesh = np.ascontiguousarray(tri_mesh, np.int32)
    mesh_origins = np.ascontiguousarray(mesh_origins, np.float32).reshape(-1, 3)
    n = len(tri_mesh)
    nodes = np.zeros((max(n - 1, 0), 16), np.float32)
    slots = np.zeros(n, np.int32)
    depth = C.c_int32(0)
    _check(lib().mcrt_host_build_sah(_p(tri_local), _p(tri_mesh), n, _p(mesh_origins), len(mesh_origins), threads, _p(nodes), _p(slots),
                                     C.byref(depth)))
    return nodes, slots, depth.value


def scene_probe(path) -> dict:
    nt, nm, nmat = C.c_int64(0), C.c_int32(0), C.c_int32(0)
    pose = np.zeros(6, np.float32)
    _check(lib().mcrt_scene_probe(str(path).encode(), C.byref(nt), C.byref(nm), C.byref(nmat), _p(pose)))
    return dict(n_triangles=nt.value, n_meshes=nm.value, n_materials=nmat.value, start_pose=pose)


def host_tables(params: Params) -> dict:
    info = Info()
    _check(lib().mcrt_host_tables(C.byref(params), C.byref(info), None, None, None, None, None))
    sc = np.empty((params.elements, 2), np.float32)
    ax = np.empty(params.psf_axial, np.float32); lat = np.empty(params.psf_lateral, np.float32)
    mx = np.empty((params.scan_rows, params.scan_cols), np.float32); my = np.empty_like(mx)
    _check(lib().mcrt_host_tables(C.byref(params), C.byref(info), _p(sc), _p(ax), _p(lat), _p(mx), _p(my)))
    return dict(info=info, elem_sincos=sc, axial=ax, lateral=lat, map_x=mx, map_y=my)
