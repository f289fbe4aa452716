#!/usr/bin/env python
"""Golden RF image of the REFERENCE'S OWN echo-accumulation loop (main.cpp:106-144, extracted verbatim into the reference probe,
oracle/Makefile + oracle/ref_probe.cpp::ref_accumulate_loop) on fixed segments.

Runs only where /root/reference is mounted.  The segments are the oracle's cast_rays output for the generated sphere scene
(512 elements x 5 samples, the reference's compile-time sizes) at a recorded seed / frame / pose -- tests regenerate them
deterministically -- and the raw `intensities` image the reference loop accumulates from them is stored.

    python tests/golden/make_golden_accumulate.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import oracle_py as O  # noqa: E402

SEED, FRAME = 4242, 3


def fixed_segments():
    """(segments [512][5][10], n_segments [512][5], materials [n][8], params) -- deterministic, no reference needed"""
    from mcray_tracing_b200 import assets
    d = assets.ensure_all()
    A = O.load_scene_py(d["sphere"] / "sphere.scene")
    osc = O.OracleScene(A)
    p = O.default_params(elements=512, samples=5)
    pos = np.asarray(A["transducer_position"], np.float32); ang = np.asarray(A["transducer_angles"], np.float32)
    segs, nseg, _ = osc.cast_rays(p, pos, ang, seed=SEED, frame=FRAME)
    return segs, nseg, np.asarray(A["materials"], np.float32).reshape(-1, 8), p, osc


def main():
    R = O.ref_probe()
    if R is None or not hasattr(R, "ref_accumulate_loop"):
        raise SystemExit("the reference tree (/root/reference) is not available: cannot regenerate the golden image")
    segs, nseg, mats, p, osc = fixed_segments()
    img = O.ref_accumulate_loop(R, segs, nseg, mats)
    mine, steps = osc.accumulate(p, segs, nseg)
    tol = 1e-4 * np.maximum(np.abs(img), 1e-3 * np.abs(img).max())
    print("reference loop: nonzero", np.count_nonzero(img), "max", np.abs(img).max(), "| oracle max abs diff", np.abs(mine - img).max(),
          "within tol:", bool(np.all(np.abs(mine - img) <= tol)), "bit-equal fraction", float(np.mean(mine == img)))
    np.savez_compressed(HERE / "reference_accumulate_loop.npz", rf=img, seed=np.array([SEED, FRAME], np.int64),
                        n_segments_total=np.array([int(nseg.sum())], np.int64))
    print("wrote", HERE / "reference_accumulate_loop.npz", (HERE / "reference_accumulate_loop.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
