#!/usr/bin/env python
"""Golden segments of the REFERENCE'S OWN scene::cast_rays<5,512> (scene.cpp:50-183, extracted verbatim into the reference probe,
oracle/ref_probe.cpp::ref_cast_rays) on a scene where every random draw of the reference is without effect.

The reference draws from a fresh random_device-seeded mt19937 at every hit, so its cast_rays is irreproducible in general.
On the scene built here it is not: thickness 0 (penetration q = 0, scene.cpp:132-139), shininess 2147483520 (cos(theta') rounds
to exactly 1.0f and random_unit_vector returns the normal bit-exactly, SURVEY C-6), and ALL IMPEDANCES EQUAL, so the reflected
intensity is exactly 0 and `reflection_probability > x` (ray.cpp:89) is false for every x: every hit refracts straight on.
What remains is the deterministic skeleton of the loop: ray set-up from the transducer, max_ray_length / enlarge, the 1 mm
start offset of the ray test, travel / distance_in_mm, the medium state machine, segment emission, termination on a miss, on
depth 10 and on the intensity epsilon.  The ray test itself is served by the oracle's closest hit (Bullet is not available).

    python tests/golden/make_golden_cast_rays.py
"""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import oracle_py as O  # noqa: E402

SEED, FRAME = 99, 1


def matched_scene():
    """the generated sphere scene (BOX + SPHERE) with the materials made impedance-matched, smooth and hard"""
    from mcray_tracing_b200 import assets
    d = assets.ensure_all()
    A = dict(O.load_scene_py(d["sphere"] / "sphere.scene"))
    m = np.asarray(A["materials"], np.float32).reshape(-1, 8).copy()
    m[:, 0] = 1.5                      # impedance: all equal -> reflected intensity exactly 0 at every boundary
    m[:, 1] = np.linspace(0.3, 1.1, len(m)).astype(np.float32)   # distinct attenuations identify the medium of a segment
    m[:, 5] = 1.0                      # specularity 1: pow(x, 1) is exact in glibc and in the contract
    m[:, 6] = 2147483520.0             # shininess: cos(theta') == 1.0f exactly
    m[:, 7] = 0.0                      # thickness
    A["materials"] = m
    return A


def oracle_segments(A):
    osc = O.OracleScene(A)
    p = O.default_params(elements=512, samples=5)
    pos = np.asarray(A["transducer_position"], np.float32); ang = np.asarray(A["transducer_angles"], np.float32)
    segs, nseg, tests = osc.cast_rays(p, pos, ang, seed=SEED, frame=FRAME)
    return segs, nseg, tests, osc, p


def reference_segments(R, A, osc):
    """scene::cast_rays<5,512> of the reference, its ray test answered by orc_closest_hit"""
    L = O.oracle()
    m = np.ascontiguousarray(A["materials"], np.float32).reshape(-1, 8)
    mi = np.ascontiguousarray(A["mesh_material_inside"], np.int32); mo = np.ascontiguousarray(A["mesh_material_outside"], np.int32)
    mv = np.ascontiguousarray(A["mesh_vascular"], np.int32)
    sp = np.ascontiguousarray(A["spacing"], np.float32)
    pos = np.ascontiguousarray(A["transducer_position"], np.float32); ang = np.ascontiguousarray(A["transducer_angles"], np.float32)
    seg12 = np.zeros((512, 5, 10, 12), np.float32); dist = np.zeros((512, 5, 10), np.float64); nseg = np.zeros((512, 5), np.int32)
    R.ref_cast_rays.restype = C.c_int64
    R.ref_cast_rays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)
    total = R.ref_cast_rays(p_(m), len(m), p_(mi), p_(mo), p_(mv), len(mi), int(A["starting_material"]), p_(sp), p_(pos), p_(ang),
                            C.cast(L.orc_closest_hit, C.c_void_p), C.c_void_p(osc.h), p_(seg12), p_(dist), p_(nseg))
    return seg12, dist, nseg, int(total)


def main():
    R = O.ref_probe()
    if R is None or not hasattr(R, "ref_cast_rays"):
        raise SystemExit("the reference tree (/root/reference) is not available: cannot regenerate the golden segments")
    A = matched_scene()
    segs, nseg, tests, osc, p = oracle_segments(A)
    seg12, dist, rnseg, total = reference_segments(R, A, osc)
    print("reference: segments", total, "| oracle: segments", int(nseg.sum()), "tests", tests, "| bounce counts equal:", bool(np.array_equal(nseg, rnseg)))
    mine = np.zeros_like(seg12)
    mine[..., 0:3] = segs["from"]; mine[..., 3:6] = segs["to"]; mine[..., 6:9] = segs["dir"]
    mine[..., 9] = segs["reflected_intensity"]; mine[..., 10] = segs["initial_intensity"]; mine[..., 11] = segs["attenuation"]
    live = np.arange(10)[None, None, :] < rnseg[..., None]
    print("bit-equal fraction of segment floats:", float(np.mean((mine == seg12)[live])), "max abs diff",
          float(np.abs(mine - seg12)[live].max()), "distance max abs diff", float(np.abs(segs["distance_traveled"] - dist)[live].max()))
    np.savez_compressed(HERE / "reference_cast_rays.npz", seg12=seg12, dist_mm=dist, nseg=rnseg, seed=np.array([SEED, FRAME], np.int64))
    print("wrote", HERE / "reference_cast_rays.npz", (HERE / "reference_cast_rays.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
