import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from mcp_raytracer_b200 import createCameraFromSceneData
def run(wl, spp, integ, width=None):
    label, kind, sopts, ropts = bench.WORKLOADS[wl]
    sd = bench.make_scene(kind, sopts)
    o = dict(ropts, samples=spp, integrator=integ)
    if width: o["width"] = width
    with createCameraFromSceneData(sd, o) as cam:
        rgb = np.zeros((cam.imageHeight, cam.imageWidth, 3), np.uint8); lin = np.zeros((cam.imageHeight, cam.imageWidth, 3), np.float32)
        st = cam.render(rgb.reshape(-1), lin)
        t0 = time.perf_counter(); st = cam.render(rgb.reshape(-1), lin); wall = time.perf_counter() - t0
    print(f"{wl} {integ:10s} {cam.imageWidth}x{cam.imageHeight}@{spp}: dev {st.deviceMs:.2f} ms wall {wall*1e3:.1f} ms  {st.samples['total']/st.deviceMs/1e3:.1f} Mpaths/s  launches {st.kernelLaunches} paths {st.samples['total']} bounces {st.bounces['total']} rays {st.rays} px {st.pixels}", flush=True)
    return rgb, lin, st
for wl, spp, w in (("C2", 16, 256), ("C1", 16, None), ("C2", 64, None), ("C3", 16, None), ("C5", 16, None), ("C4", 4, None)):
    a = run(wl, spp, "megakernel", w); b = run(wl, spp, "wavefront", w)
    d = np.abs(a[1] - b[1])
    print("   max |lin diff|", float(d.max()), "rgb8 equal frac", float((a[0] == b[0]).mean()), "stats equal", (a[2].samples, a[2].bounces['total'], a[2].rays) == (b[2].samples, b[2].bounces['total'], b[2].rays))
