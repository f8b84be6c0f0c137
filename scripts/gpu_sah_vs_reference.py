"""Same-seed renders through the 4-wide SAH tree and through the reference-topology tree (true min/max boxes, the reference's
visiting order): every closest hit is decided by the same primitive tests, so the two images are equal unless a box test
rejected a box the ray really enters.  Prints the number of pixels that differ.
usage: gpu_sah_vs_reference.py [workload:width:spp ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from mcp_raytracer_b200 import createCameraFromSceneData
specs = sys.argv[1:] or ["C3:480:16", "C1:400:16", "C4:480:4"]
for spec in specs:
    wl, width, spp = spec.split(":")
    label, kind, sopts, ropts = bench.WORKLOADS[wl]
    sd = bench.make_scene(kind, sopts)
    imgs = {}
    for bvh in ("sah", "reference"):
        o = dict(ropts, width=int(width), samples=int(spp), bvh=bvh, aTolerance=0)
        with createCameraFromSceneData(sd, o) as cam:
            rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
            lin = np.zeros((cam.imageHeight, cam.imageWidth, 3), np.float32)
            st = cam.render(rgb, lin)
            imgs[bvh] = (lin.copy(), st.rays, st.bounces["total"])
    a, b = imgs["sah"][0], imgs["reference"][0]
    diff = np.any(a != b, axis=2)
    print(spec, "pixels differing:", int(diff.sum()), "of", diff.size, "rays", imgs["sah"][1], imgs["reference"][1],
          "max |diff|", float(np.abs(a - b).max()))
