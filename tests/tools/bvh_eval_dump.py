"""Dump for tests/tools/bvh_eval.cpp: the scene's triangles (mesh-local + body origins) and the closest-hit queries (from_test, to) of one
frame as the oracle casts them.  usage: python tests/tools/bvh_eval_dump.py ircad|stress  ->  /tmp/ev/<name>.bin"""
import os
import sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.makedirs('/tmp/ev', exist_ok=True)
from mcray_tracing_b200 import assets
from oracle import oracle_py as O
which = sys.argv[1]
if which == 'ircad':
    d = assets.ensure_all()
    scene = d["ircad11"] / "santi-liver.scene"
    A = O.load_scene_py(scene)
    E, S = 256, 16
    import json
else:
    A = assets.stress_scene_arrays()
    E, S = 512, 4
osc = O.OracleScene(A)
op = O.default_params(elements=E, samples=S)
if which == 'ircad':
    pos, ang = np.array(A["transducer_position"], np.float32), np.array(A["transducer_angles"], np.float32)
else:
    pos, ang = A["transducer_position"], A["transducer_angles"]
segs, nseg, tests = osc.cast_rays(op, pos, ang, seed=1234, frame=0)
print("segments", int(nseg.sum()), "tests", tests)
# scene triangles in world coordinates
tri = (np.asarray(A["tri_vertices"], np.float32) * np.float32(A["scaling"])).reshape(-1, 3, 3)
offs = np.asarray(A["tri_offsets"])
mesh = np.concatenate([np.full(int(offs[m + 1] - offs[m]), m, np.int32) for m in range(len(offs) - 1)])
org = (np.asarray(A["mesh_deltas"], np.float32) * np.float32(A["scaling"]) * np.float32(A["scaling"]) + np.asarray(A["origin"], np.float32)[None, :])
mats = np.asarray(A["materials"], np.float32)
print("materials cols", mats.shape)
# rays
fr, di, I0, att = [], [], [], []
for e in range(E):
    for s in range(S):
        for k in range(nseg[e, s]):
            g = segs[e, s, k]
            fr.append(g["from"]); di.append(g["dir"]); I0.append(g["initial_intensity"]); att.append(g["attenuation"])
fr = np.array(fr, np.float32); di = np.array(di, np.float32); I0 = np.array(I0, np.float32); att = np.array(att, np.float32)
freq = np.float32(op.frequency_mhz)
sp = np.asarray(A["spacing"], np.float32)
rl = np.float32(10.0) * np.log(np.float32(1e-10) / I0) / -att * freq
to = fr + di * sp[None, :] * (rl / np.float32(100.0))[:, None]
ft = fr + di * np.float32(0.1)
with open(f'/tmp/ev/{which}.bin', 'wb') as f:
    np.array([len(mesh), len(org), len(ft)], np.int64).tofile(f)
    tri.astype(np.float32).tofile(f); mesh.tofile(f); org.astype(np.float32).tofile(f)
    ft.astype(np.float32).tofile(f); to.astype(np.float32).tofile(f)
print("wrote", len(mesh), "tris", len(ft), "rays")
