"""GPU box: primary hits of degenerate scenes (coincident / collinear / nested primitives, zero radii, zero-area
quads, FP32-range coordinates, inverted boxes) against the oracle, for every tree kind, plus a short render of
each to show the integrator terminates.  Run under `timeout`: a hang here would be a traversal bug.

    timeout 60 python scripts/gpu_degenerate.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_binding import OracleCamera  # noqa: E402  (the checker)

from mcp_raytracer_b200 import createCameraFromSceneData  # noqa: E402

rng = np.random.default_rng(0)
m = {"type": "lambert", "color": [0.7, 0.6, 0.5]}
g = {"type": "glass", "ior": 1.5}


def sph(p, r, mat=m):
    return {"type": "sphere", "pos": [float(x) for x in p], "r": float(r), "material": mat}


cases = {
    "coincident": ([sph([0, 0, 0], 1.0) for _ in range(200)], [0, 0, 5]),
    "collinear": ([sph([i - 100, 0, 0], 0.6) for i in range(200)], [0, 3, 40]),
    "nested": ([sph([0, 0, 0], 1 + i * 1e-2, g if i % 2 else m) for i in range(200)], [0, 0, 9]),
    "zero_radius": ([sph(rng.random(3) * 4 - 2, 0.0) for _ in range(200)] + [sph([0, 0, 0], 0.5)], [0, 0, 5]),
    "fp32_large": ([sph((rng.random(3) - 0.5) * 1e30, 1e29) for _ in range(200)], [0, 0, 2e30]),
    "fp32_tiny": ([sph((rng.random(3) - 0.5) * 1e-15, 1e-16) for _ in range(200)], [0, 0, 2e-15]),
    "zero_area_quads": ([{"type": "quad", "pos": (rng.random(3) * 2 - 1).tolist(), "u": [0, 0, 0], "v": [0, 0, 0], "material": m}
                         for _ in range(100)] + [sph([0, 0, 0], 0.5)], [0, 0, 5]),
    "inverted_boxes": ([sph(rng.random(3) * 4 - 2, -0.3, g) for _ in range(100)] + [sph([0, -101, 0], 100)], [0, 1, 7]),
}
bad = 0
for name, (objs, eye) in cases.items():
    # focus = distance to the origin: with the default focus 1 the pixel grid of a camera 1e30 away collapses in FP32
    focus = float(np.linalg.norm(eye))
    sd = {"type": "custom", "camera": {"vfov": 40, "from": eye, "at": [0, 0, 0], "focus": focus}, "objects": objs}
    for bvh in ("auto", "sah", "reference"):
        opts = {"width": 64, "aspect": 16 / 9, "samples": 8, "depth": 12, "aTolerance": 0, "seed": 3, "bvh": bvh}
        t0 = time.perf_counter()
        with createCameraFromSceneData(sd, opts) as cam:
            ids, t, nrm, _ = cam.tracePrimary()
            rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
            st = cam.render(rgb)
        oids, ot, onrm, _ = OracleCamera(sd, opts).trace_primary()
        hit = oids >= 0
        same_ids = float(np.mean(ids == oids))
        rel_t = float(np.max(np.abs(t[hit] - ot[hit]) / np.maximum(np.abs(ot[hit]), 1e-30))) if hit.any() else 0.0
        ok = same_ids == 1.0 and rel_t <= 1e-4 and st.samples["total"] == st.pixels * 8
        # a FORCED SAH tree hits negative-radius spheres the reference's inverted boxes hide (DESIGN.md section 7;
        # AUTO picks the reference topology for such scenes): reported, not counted
        documented = name == "inverted_boxes" and bvh == "sah"
        bad += (not ok) and not documented
        print(f"{name:16s} {bvh:9s} hits {int(hit.sum()):5d}/{hit.size}  ids equal {same_ids:.4f}  max rel t {rel_t:.2e}  "
              f"rays {st.rays:7d}  {1e3 * (time.perf_counter() - t0):7.1f} ms  {'ok' if ok else ('differs (documented: forced SAH)' if documented else 'DIFFERS')}", flush=True)
print("degenerate scenes:", "all ok" if bad == 0 else f"{bad} differ")
sys.exit(1 if bad else 0)
