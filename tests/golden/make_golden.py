"""Generates the committed golden fixtures from the CPU oracle.

The reference is TypeScript and cannot run in this image (no node/tsc), and its own
tests hold no golden images, so the fixtures are outputs of the oracle — which is itself pinned to the
reference's known-answer vectors by tests/test_oracle_reference_vectors.py.  Re-run after an
intentional oracle change:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_binding as ob  # noqa: E402
from mcp_raytracer_b200 import scenes  # noqa: E402

PRIMARY = {
    "C2-cornell": (scenes.generateCornellSceneData, 128),
    "default": (scenes.generateDefaultSceneData, 128),
}
RENDERS = {
    # name: (scene fn, options, seed)
    "C2-cornell": (scenes.generateCornellSceneData, {"width": 32, "samples": 16, "aTolerance": 0}, 0),
    "C1-spheres": (lambda: scenes.generateSpheresSceneData({"count": 100, "seed": 12345}),
                   {"width": 48, "samples": 16, "depth": 10, "aTolerance": 0}, 0),
    "default-adaptive": (scenes.generateDefaultSceneData, {"width": 40, "samples": 30}, 0),
}


def main():
    for name, (fn, width) in PRIMARY.items():
        cam = ob.OracleCamera(fn(), {"width": width, "samples": 1})
        ids, t, nrm, ff = cam.trace_primary()
        np.savez_compressed(os.path.join(HERE, f"primary_{name}.npz"), width=width, ids=ids, t=t.astype(np.float64),
                            normal=nrm, front=ff)
    for name, (fn, opts, seed) in RENDERS.items():
        r = ob.OracleCamera(fn(), opts).render(seed=seed, threads=1)
        st = r["stats"]
        np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), linear=r["linear"], rgb8=r["rgb8"], seed=seed,
                            stats=np.array([st.pixels, st.samples_total, st.samples_min, st.samples_max, st.bounces_total,
                                            st.bounces_min, st.bounces_max, st.rays], np.int64))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
