"""Whole-query while-while (k_render_pool<SAH>) vs resumable state machine (k_render_trav) by tree size.
Run twice: plain (AUTO rule) and with RT_B200_NO_TRAV=1 / RT_B200_TRAV_ALWAYS=1.  usage: gpu_trav_threshold.py counts..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mcp_raytracer_b200 import createCameraFromSceneData, generateRainSceneData
for count in [int(a) for a in sys.argv[1:]] or [1000, 4000, 16000, 50000]:
    sd = generateRainSceneData({"count": count, "seed": 1, "sphereRadius": 0.01 * (100000 / count) ** (1 / 3)})
    with createCameraFromSceneData(sd, {"width": 1920, "samples": 8, "aTolerance": 0}) as cam:
        rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
        best = min(cam.render(rgb).deviceMs for _ in range(3))
        print(count, "spheres,", cam.info.n_bvh_nodes, "nodes:", f"{best:.2f} ms", flush=True)
