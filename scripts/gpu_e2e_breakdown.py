"""Where the end-to-end time of one create + render + close goes (host clock, after warm-up)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from mcp_raytracer_b200 import createCameraFromSceneData
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else None
label, kind, sopts, ropts = bench.WORKLOADS[wl]
if spp: ropts = dict(ropts, samples=spp)
sd = bench.make_scene(kind, sopts)
for it in range(4):
    t0 = time.perf_counter()
    cam = createCameraFromSceneData(sd, ropts)
    t1 = time.perf_counter()
    rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
    t2 = time.perf_counter()
    st = cam.render(rgb)
    t3 = time.perf_counter()
    cam.close()
    t4 = time.perf_counter()
    print(f"{wl} it{it}: create {1e3*(t1-t0):.2f} ms, alloc host {1e3*(t2-t1):.2f}, render call {1e3*(t3-t2):.2f} (device {st.deviceMs:.2f}), close {1e3*(t4-t3):.2f}, total {1e3*(t4-t0):.2f}")
