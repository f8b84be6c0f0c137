"""One render of a workload at reduced spp — the command profiled under ncu.
usage: prof_render.py <C1..C5> [spp] [reps] [key=value render options ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from mcp_raytracer_b200 import createCameraFromSceneData
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
extra = dict(kv.split("=", 1) for kv in sys.argv[4:])
label, kind, sopts, ropts = bench.WORKLOADS[wl]
sd = bench.make_scene(kind, sopts)
with createCameraFromSceneData(sd, dict(ropts, samples=spp, **extra)) as cam:
    rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
    for _ in range(reps):
        st = cam.render(rgb)
    print(f"{wl} {cam.imageWidth}x{cam.imageHeight}@{spp} {extra}: {st.deviceMs:.3f} ms, {st.samples['total']/st.deviceMs/1e3:.1f} Mpaths/s, {st.rays/st.deviceMs/1e6:.3f} Grays/s")
