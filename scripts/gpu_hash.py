"""Render workloads at reduced size and print an image hash + stats: run under different RT_B200_* development
switches to check that kernel variants are bit-identical.
usage: gpu_hash.py [width] [spp] [workloads...]"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from mcp_raytracer_b200 import createCameraFromSceneData
width = int(sys.argv[1]) if len(sys.argv) > 1 else 200
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wls = sys.argv[3:] or ["C1", "C2", "C5"]
for wl in wls:
    label, kind, sopts, ropts = bench.WORKLOADS[wl]
    sd = bench.make_scene(kind, sopts)
    for region in (None, {"x": 13, "y": 7, "width": width // 2 + 3, "height": width // 3 + 1}):
        with createCameraFromSceneData(sd, dict(ropts, width=width, samples=spp)) as cam:
            rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
            st = cam.render(rgb) if region is None else cam.renderRegion(rgb, region)
            print(wl, "region" if region else "full", hashlib.sha1(rgb.tobytes()).hexdigest()[:16], st.pixels, st.samples["total"],
                  st.bounces["total"], st.bounces["min"], st.bounces["max"], st.rays, f"{st.deviceMs:.3f} ms")
