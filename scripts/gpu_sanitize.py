"""Tiny renders through every render kernel — the command run under compute-sanitizer (memcheck / racecheck).
usage: gpu_sanitize.py [kernels...]   (default: all)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mcp_raytracer_b200 import (createCameraFromSceneData, generateCornellSceneData, generateLayeredMixedSceneData,
                                generateRainSceneData, generateSpheresSceneData)

CASES = {
    "pool_list": (generateCornellSceneData, {}, {"width": 48, "samples": 24, "aTolerance": 0, "integrator": "megakernel"}),
    "sorted_list": (generateLayeredMixedSceneData, {}, {"width": 48, "samples": 24, "aTolerance": 0, "integrator": "sorted"}),
    "pool_sah": (generateSpheresSceneData, {"count": 100, "seed": 12345}, {"width": 64, "samples": 8, "aTolerance": 0, "integrator": "megakernel"}),
    "sorted_sah": (generateSpheresSceneData, {"count": 100, "seed": 12345}, {"width": 64, "samples": 8, "aTolerance": 0, "integrator": "sorted"}),
    "trav": (generateRainSceneData, {"count": 20000, "seed": 1, "sphereRadius": 0.01}, {"width": 64, "samples": 2, "aTolerance": 0}),
    "reference_tree": (generateSpheresSceneData, {"count": 30, "seed": 3}, {"width": 48, "samples": 4, "aTolerance": 0, "bvh": "reference"}),
    "stream_adaptive": (generateCornellSceneData, {}, {"width": 40, "samples": 40, "aTolerance": 0.05}),
    "stream_region": (generateSpheresSceneData, {"count": 50, "seed": 5}, {"width": 70, "samples": 12, "mode": "bounces"}),
    "wavefront": (generateCornellSceneData, {}, {"width": 48, "samples": 8, "aTolerance": 0, "integrator": "wavefront"}),
}
names = sys.argv[1:] or list(CASES)
for name in names:
    gen, sopts, ropts = CASES[name]
    sd = gen(sopts) if sopts else gen()
    with createCameraFromSceneData(sd, ropts) as cam:
        rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
        region = {"x": 5, "y": 3, "width": 41, "height": 22} if name == "stream_region" else None
        st = cam.renderRegion(rgb, region) if region else cam.render(rgb)
        if name == "pool_list":
            cam.tracePrimary()
        print(name, cam.imageWidth, cam.imageHeight, st.pixels, st.samples["total"], st.rays, int(rgb.sum()), flush=True)
print("done")
