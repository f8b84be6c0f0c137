"""LIST vs SAH crossover on the spheres scene (development aid). usage: gpu_list_vs_sah.py counts..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mcp_raytracer_b200 import createCameraFromSceneData, generateSpheresSceneData
for count in [int(a) for a in sys.argv[1:]] or [16, 24, 32, 48, 64]:
    sd = generateSpheresSceneData({"count": count, "seed": 7})
    # a ground sphere so that paths bounce (the stock scene is mostly sky)
    sd["objects"].append({"type": "sphere", "pos": [0, -1000.5, 0], "r": 1000, "material": {"type": "lambert", "color": [0.5, 0.5, 0.5]}})
    out = []
    for bvh in ("list", "sah"):
        for integ in ("megakernel", "sorted"):
            with createCameraFromSceneData(sd, {"width": 1024, "samples": 64, "aTolerance": 0, "bvh": bvh, "integrator": integ}) as cam:
                rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
                best = min(cam.render(rgb).deviceMs for _ in range(3))
                out.append(f"{bvh}/{integ} {best:.2f}")
    print(count + 1, "objects:", "  ".join(out), flush=True)
