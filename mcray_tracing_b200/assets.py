"""Synthetic scene / mesh generator.

The reference ships scene files (examples/sphere/*.scene, examples/ircad11/*.scene) that name OBJ
meshes which are NOT in its tree (BOX.obj, SPHERE.obj and the 11 IRCAD
3D-IRCADb-01 patient-11 organs).  There is no network, so this module synthesises stand-in meshes
under the same file names, and writes scene files with the reference's JSON schema
(scene.cpp:185-247, main.cpp:65-69) and the reference's material tables / mesh placement values.

Everything is deterministic (fixed formulas, no RNG) so that every machine regenerates identical
bytes.  Output goes to assets/_gen/ (git-ignored, travels with gpurun).
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np

REPO_ROOT = Path(__file__).resolve().parent.parent
GEN_DIR = REPO_ROOT / "assets" / "_gen"

# ---------------------------------------------------------------------------------------------
# material tables (values of examples/sphere/sphere.scene:5-126, examples/ircad11/santi-liver.scene)
# columns: impedance, attenuation, mu0, mu1, sigma, specularity, shininess, thickness
# ---------------------------------------------------------------------------------------------
_MAT_KEYS = ("impedance", "attenuation", "mu0", "mu1", "sigma", "specularity", "shininess", "thickness")


def _materials(gel_impedance: float, bone_thickness: float = 0.0, kidney_shininess: float = 1000000,
               with_surface_keys: bool = True, overrides: dict | None = None):
    rows = [
        ("GEL", gel_impedance, 1e-8, 0.0, 0.0, 0.0, 1.0, 1000000, 0.0),
        ("AIR", 0.0004, 1.64, 0.78, 0.56, 0.1, 1.0, 1000000, 0.0),
        ("FAT", 1.38, 0.63, 0.5, 0.5, 0.0, 1.0, 1000000, 0.0),
        ("LIVER", 1.65, 0.7, 0.19, 1.0, 0.24, 1.0, 1000000, 0.0),
        ("BONE", 7.8, 5.0, 0.78, 0.56, 0.1, 1.0, 1000000, bone_thickness),
        ("BLOOD", 1.61, 0.18, 0.001, 0.0, 0.01, 1.0, 1000000, 0.0),
        ("VESSEL", 1.99, 1.09, 0.2, 0.1, 0.2, 1.0, 1000000, 0.0),
        ("KIDNEY", 1.62, 1.0, 0.4, 0.6, 0.3, 1.0, kidney_shininess, 0.0),
        ("SUPRARRENAL", 1.62, 1.0, 0.4, 0.6, 0.3, 1.0, 1000000, 0.0),
        ("GALLBLADDER", 1.62, 1.0, 0.4, 0.6, 0.3, 1.0, 1000000, 0.0),
        ("SKIN", 1.99, 1.0, 0.4, 0.6, 0.3, 1.0, 1000000, 0.0),
    ]
    out = []
    for r in rows:
        m = {"name": r[0]}
        for k, v in zip(_MAT_KEYS, r[1:]):
            m[k] = v
        if overrides and r[0] in overrides:
            m.update(overrides[r[0]])
        if not with_surface_keys:     # examples/ircad11/ircad11.scene:7-15 has no shininess/thickness
            m.pop("shininess")
            m.pop("thickness")
        out.append(m)
    return out


def _mesh(file, material, outside, deltas=(0.0, 0.0, 0.0), vascular=False):
    return {"file": file, "rigid": True, "vascular": bool(vascular), "deltas": [float(d) for d in deltas],
            "material": material, "outsideMaterial": outside, "outsideNormals": True}


# organ -> (deltas [mm, IRCAD frame], inside, outside, vascular)   santi-liver.scene:128-228
IRCAD_ORGANS = [
    ("aorta", (152.533512115, 174.472991943, 105.106495678), "BLOOD", "FAT", True),
    ("bones", (188.265544891, 202.440551758, 105.599998474), "BONE", "FAT", False),
    ("liver", (141.238292694, 176.429901123, 130.10585022), "LIVER", "FAT", False),
    ("cava", (206.332504272, 192.29649353, 104.897496045), "BLOOD", "FAT", True),
    ("right_kidney", (118.23374939, 218.907501221, 53.6022927761), "KIDNEY", "SKIN", False),
    ("left_kidney", (251.052993774, 227.63949585, 64.8468027115), "KIDNEY", "SKIN", False),
    ("right_suprarrenal", (152.25050354, 213.971496582, 115.338005066), "SUPRARRENAL", "FAT", False),
    ("left_suprarrenal", (217.128997803, 209.525497437, 102.477149963), "SUPRARRENAL", "FAT", False),
    ("gallbladder", (128.70715332, 146.592498779, 112.361503601), "GALLBLADDER", "FAT", False),
    ("skin", (188.597551346, 199.367202759, 105.622316509), "FAT", "GEL", False),
    ("porta", (182.364089966, 177.214996338, 93.0034988523), "BLOOD", "FAT", True),
]

# synthetic organ shapes: semi-axes [mm] and icosphere subdivision level (our assumption; the
# original meshes are lost).  20*4^level triangles each.
IRCAD_SHAPES = {
    "aorta": ((12.0, 12.0, 140.0), 5),
    "bones": ((25.0, 30.0, 150.0), 6),
    "liver": ((75.0, 65.0, 70.0), 6),
    "cava": ((14.0, 14.0, 130.0), 5),
    "right_kidney": ((28.0, 22.0, 50.0), 5),
    "left_kidney": ((28.0, 22.0, 50.0), 5),
    "right_suprarrenal": ((12.0, 8.0, 15.0), 4),
    "left_suprarrenal": ((12.0, 8.0, 15.0), 4),
    "gallbladder": ((18.0, 15.0, 30.0), 5),
    "skin": ((168.0, 125.0, 160.0), 7),
    "porta": ((10.0, 35.0, 10.0), 5),
}


# ---------------------------------------------------------------------------------------------
# meshes
# ---------------------------------------------------------------------------------------------
def icosphere(level: int):
    """Unit icosphere: (vertices float64 [n,3], faces int64 [m,3]), m = 20 * 4**level."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        n = len(v)
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * n + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // n, uniq % n
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        m = inv + n
        nf = len(f)
        m01, m12, m20 = m[:nf], m[nf:2 * nf], m[2 * nf:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    return v, f


def blob(center, radii, level, phase=0.0, bump=0.06):
    """Smoothly displaced ellipsoid: unit icosphere scaled by `radii`, radius modulated by a few
    low-frequency sinusoids (deterministic), centred at `center`."""
    v, f = icosphere(level)
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    mod = 1.0 + bump * (np.sin(3.0 * x + phase) * np.cos(2.0 * y - phase) + 0.5 * np.sin(5.0 * z + 2.0 * phase) * np.cos(4.0 * x))
    p = v * mod[:, None] * np.asarray(radii, dtype=np.float64)[None, :] + np.asarray(center, dtype=np.float64)[None, :]
    return p, f


def uv_sphere_soup(center, radius, nu, nv, phase=0.0, bump=0.0):
    """UV sphere as a triangle soup float32 [2*nu*nv, 9] (degenerate pole triangles included)."""
    u = np.linspace(0.0, 2.0 * np.pi, nu + 1)
    w = np.linspace(0.0, np.pi, nv + 1)
    uu, ww = np.meshgrid(u, w, indexing="ij")
    r = radius * (1.0 + bump * np.sin(7.0 * uu + phase) * np.sin(5.0 * ww - phase))
    p = np.stack([r * np.sin(ww) * np.cos(uu), r * np.sin(ww) * np.sin(uu), r * np.cos(ww)], axis=-1) + np.asarray(center)[None, None, :]
    p00, p10, p01, p11 = p[:-1, :-1], p[1:, :-1], p[:-1, 1:], p[1:, 1:]
    t0 = np.concatenate([p00, p10, p11], axis=-1).reshape(-1, 9)
    t1 = np.concatenate([p00, p11, p01], axis=-1).reshape(-1, 9)
    return np.concatenate([t0, t1], axis=0).astype(np.float32)


def write_obj(path: Path, verts: np.ndarray, faces: np.ndarray, *, quads: np.ndarray | None = None,
              with_normals: bool = False, header: str = "") -> None:
    """Write a Wavefront OBJ.  `faces` are 0-based triangles; `quads` optional 0-based 4-gons (written
    as polygon faces so loaders must fan-triangulate, tiny_obj_loader.cpp:272-285)."""
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_suffix(path.suffix + ".tmp")
    with open(tmp, "w") as fh:
        fh.write(f"# {header}\n")
        fh.write("o " + path.stem + "\n")
        np.savetxt(fh, verts, fmt="v %.6f %.6f %.6f")
        if with_normals:
            c = verts.mean(axis=0, keepdims=True)
            nrm = verts - c
            nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-12)
            np.savetxt(fh, nrm, fmt="vn %.6f %.6f %.6f")
        if quads is not None and len(quads):
            np.savetxt(fh, quads + 1, fmt="f %d %d %d %d")
        if len(faces):
            if with_normals:
                idx = np.repeat(faces + 1, 2, axis=1)
                np.savetxt(fh, idx, fmt="f %d//%d %d//%d %d//%d")
            else:
                np.savetxt(fh, faces + 1, fmt="f %d %d %d")
    os.replace(tmp, path)


def _box(half: float):
    v = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64) * half
    q = np.array([[0, 3, 2, 1], [4, 5, 6, 7], [0, 1, 5, 4], [2, 3, 7, 6], [1, 2, 6, 5], [0, 4, 7, 3]], dtype=np.int64)
    return v, q


# ---------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------
def _write_scene(path: Path, scene: dict) -> None:
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_suffix(path.suffix + ".tmp")
    with open(tmp, "w") as fh:
        json.dump(scene, fh, indent=2)
    os.replace(tmp, path)


def ensure_sphere(root: Path = GEN_DIR) -> Path:
    """examples/sphere: BOX.obj (cube half-edge 10 cm, 6 quads -> 12 triangles), SPHERE.obj (icosphere
    radius 4 cm, level 5 = 20480 triangles, written with vn + `f a//a`), sphere/box/simple.scene."""
    d = root / "sphere"
    if not (d / "BOX.obj").exists():
        v, q = _box(10.0)
        write_obj(d / "BOX.obj", v, np.zeros((0, 3), np.int64), quads=q, header="synthetic BOX: axis-aligned cube, half edge 10")
    if not (d / "SPHERE.obj").exists():
        v, f = icosphere(5)
        write_obj(d / "SPHERE.obj", v * 4.0, f, with_normals=True, header="synthetic SPHERE: icosphere radius 4, level 5")
    common = dict(transducerPosition=[-13.5, 0.0, 0.0], transducerAngles=[0.0, 0.0, -90.0], origin=[0.0, 0.0, 0.0],
                  spacing=[1.0, 1.0, 1.0], scaling=1.0)
    if not (d / "sphere.scene").exists():
        _write_scene(d / "sphere.scene", dict(workingDirectory="/home/santiago/Proyectos/MCRay-Tracing/examples/sphere/", **common,
                                              materials=_materials(1.99),
                                              meshes=[_mesh("BOX.obj", "LIVER", "GEL"), _mesh("SPHERE.obj", "BONE", "LIVER")],
                                              startingMaterial="GEL"))
    if not (d / "box.scene").exists():
        _write_scene(d / "box.scene", dict(**common, materials=_materials(1.99), meshes=[_mesh("BOX.obj", "BONE", "FAT")], startingMaterial="FAT"))
    if not (d / "simple.scene").exists():
        _write_scene(d / "simple.scene", dict(**common, materials=_materials(1.99), meshes=[_mesh("SPHERE.obj", "BONE", "FAT")], startingMaterial="FAT"))
    return d


def ensure_ircad11(root: Path = GEN_DIR) -> Path:
    """examples/ircad11: 11 synthetic organs in absolute-mm IRCAD coordinates with centroid = the
    scene's `deltas`, and the four scene files."""
    d = root / "ircad11"
    for i, (name, deltas, _, _, _) in enumerate(IRCAD_ORGANS):
        p = d / f"{name}.obj"
        if p.exists():
            continue
        radii, level = IRCAD_SHAPES[name]
        v, f = blob(deltas, radii, level, phase=0.7 * i, bump=0.02 if name == "skin" else 0.06)
        write_obj(p, v, f, header=f"synthetic {name}: displaced ellipsoid radii {radii} mm, icosphere level {level}")
    meshes = [_mesh(f"{n}.obj", mi, mo, dl, vas) for (n, dl, mi, mo, vas) in IRCAD_ORGANS]
    base = dict(workingDirectory="/home/santiago/Proyectos/MCRay-Tracing/examples/ircad11/", origin=[-18.0, -22.0, -5.0],
                spacing=[1.0, 1.0, 1.0], scaling=0.1, startingMaterial="GEL", meshes=meshes)
    variants = {
        "santi-liver.scene": dict(transducerPosition=[-17.5, 1.0, 5.0], transducerAngles=[120.0, 0.0, -90.0], materials=_materials(1.38, bone_thickness=0.3)),
        "santi-morison1.scene": dict(transducerPosition=[-16.0, 3.0, 14.0], transducerAngles=[45.0, 45.0, -90.0],
                                     materials=_materials(1.38, bone_thickness=0.1, kidney_shininess=10000)),
        "santi-morison2.scene": dict(transducerPosition=[-16.0, 3.0, 2.0], transducerAngles=[90.0, 0.0, -90.0], materials=_materials(1.38, bone_thickness=0.3)),
        # ircad11.scene omits shininess/thickness (does not load in the reference)
        "ircad11.scene": dict(transducerPosition=[-16, 3, 2], transducerAngles=[90, 0, -90],
                              materials=_materials(1.38, with_surface_keys=False,
                                                   overrides={"BLOOD": {"specularity": 0.001}, "KIDNEY": {"specularity": 0.2}})),
        # a "rough" variant for stochastic-mode tests: finite shininess and thickness everywhere
        "santi-liver-rough.scene": dict(transducerPosition=[-17.5, 1.0, 5.0], transducerAngles=[120.0, 0.0, -90.0],
                                        materials=_materials(1.38, bone_thickness=0.3,
                                                             overrides={k: {"shininess": 50, "thickness": 0.2} for k in
                                                                        ("FAT", "LIVER", "BLOOD", "KIDNEY", "SUPRARRENAL", "GALLBLADDER", "SKIN")})),
    }
    for fname, extra in variants.items():
        if not (d / fname).exists():
            sc = dict(base)
            sc.update(extra)
            _write_scene(d / fname, sc)
    return d


def stress_scene_arrays(shells: int = 8, nu: int = 512, nv: int = 256):
    """BASELINE config 4: `shells` nested bumpy UV-sphere shells (2*nu*nv triangles each; default
    8 x 262144 = 2 097 152 triangles), materials alternating LIVER / FAT / BLOOD(vascular), finite
    shininess/thickness so paths scatter.  Returned as in-memory arrays for mcrt_create_from_arrays."""
    mats = _materials(1.38, overrides={k: {"shininess": 2, "thickness": 0.5} for k in ("FAT", "LIVER", "BLOOD", "GEL", "KIDNEY")})
    names = [m["name"] for m in mats]
    cycle = [("LIVER", "FAT", False), ("FAT", "LIVER", False), ("BLOOD", "FAT", True), ("KIDNEY", "FAT", False)]
    tris, offs, m_in, m_out, vas = [], [0], [], [], []
    for s in range(shells):
        radius = 9.0 - s * (8.0 / shells)
        soup = uv_sphere_soup((0.0, 0.0, 0.0), radius, nu, nv, phase=0.9 * s, bump=0.03)
        tris.append(soup)
        offs.append(offs[-1] + len(soup))
        a, b, v = cycle[s % len(cycle)]
        m_in.append(names.index(a)); m_out.append(names.index(b)); vas.append(int(v))
    return dict(
        materials=np.array([[m[k] for k in _MAT_KEYS] for m in mats], dtype=np.float32), material_names=names,
        starting_material=names.index("GEL"), mesh_material_inside=np.array(m_in, np.int32), mesh_material_outside=np.array(m_out, np.int32),
        mesh_vascular=np.array(vas, np.int32), mesh_deltas=np.zeros((shells, 3), np.float32), tri_offsets=np.array(offs, np.int64),
        tri_vertices=np.concatenate(tris, axis=0), scaling=1.0, origin=np.zeros(3, np.float32), spacing=np.ones(3, np.float32),
        transducer_position=np.array([-13.5, 0.0, 0.0], np.float32), transducer_angles=np.array([0.0, 0.0, -90.0], np.float32))


def ensure_all(root: Path = GEN_DIR) -> dict:
    return {"sphere": ensure_sphere(root), "ircad11": ensure_ircad11(root)}


def sweep_poses(n: int = 512) -> np.ndarray:
    """BASELINE config 3: freehand sweep bracketing the three santi-* poses:
    z from 2.0 to 14.0, X angle 90 -> 120 deg.  Returns float32 [n, 6] = (pos xyz, angles xyz deg)."""
    t = np.linspace(0.0, 1.0, n)
    poses = np.zeros((n, 6), dtype=np.float32)
    poses[:, 0] = -17.5
    poses[:, 1] = 1.0 + 2.0 * t
    poses[:, 2] = 2.0 + 12.0 * t
    poses[:, 3] = 90.0 + 30.0 * t
    poses[:, 4] = 0.0
    poses[:, 5] = -90.0
    return poses


if __name__ == "__main__":
    print(ensure_all())
