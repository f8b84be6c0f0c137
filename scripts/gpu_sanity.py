"""Quick GPU sanity + timing sweep (development aid, not a test)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mcp_raytracer_b200 import *
from oracle_binding import OracleCamera

def primary(name, sd, opts):
    t0 = time.time()
    with createCameraFromSceneData(sd, opts) as cam:
        tb = time.time() - t0
        ids, t, nrm, ff = cam.tracePrimary()
        info = cam.info
    oc = OracleCamera(sd, opts)
    oids, ot, onrm, off = oc.trace_primary()
    mism = int((ids != oids).sum())
    hit = (oids >= 0) & (ids == oids)
    rel_t = float(np.max(np.abs(t[hit] - ot[hit]) / np.abs(ot[hit]))) if hit.any() else 0
    dn = float(np.max(np.abs(nrm[hit] - onrm[hit]))) if hit.any() else 0
    print(f"[primary] {name}: {ids.shape} bvh={info.bvh_kind} nodes={info.n_bvh_nodes} build={tb*1e3:.1f}ms id_mismatch={mism} max_rel_t={rel_t:.2e} max_dn={dn:.2e} ff_mismatch={int((ff[hit]!=off[hit]).sum())}", flush=True)

def timing(name, sd, opts, reps=3):
    with createCameraFromSceneData(sd, opts) as cam:
        W, H = cam.imageWidth, cam.imageHeight
        rgb = np.zeros(W*H*3, np.uint8)
        best = None
        for r in range(reps):
            st = cam.render(rgb)
            if best is None or st.deviceMs < best.deviceMs: best = st
        st = best
        print(f"[time] {name}: {W}x{H}@{opts['samples']} {st.deviceMs:.2f} ms  {st.samples['total']/st.deviceMs/1e3:.1f} Mpaths/s  {st.rays/st.deviceMs/1e6:.3f} Grays/s  bounces avg {st.bounces['avg']:.2f} max {st.bounces['max']}", flush=True)
        return rgb.reshape(H, W, 3)

import __graft_entry__ as g
g.smoke()
print("fp32 peak", measureFp32Peak())
cornell = generateCornellSceneData()
spheres = generateSpheresSceneData({"count": 100, "seed": 12345})
weekend = generateWeekendFinalSceneData()
rain = generateRainSceneData({"count": 100000, "seed": 1, "sphereRadius": 0.01})
layered = generateLayeredMixedSceneData()
default = generateDefaultSceneData()
for bvh in ("auto", "reference", "sah"):
    primary("cornell/"+bvh, cornell, {"width": 512, "samples": 1, "bvh": bvh})
    primary("spheres100/"+bvh, spheres, {"width": 400, "samples": 1, "bvh": bvh})
    primary("weekend/"+bvh, weekend, {"width": 480, "samples": 1, "bvh": bvh})
    primary("layered/"+bvh, layered, {"width": 512, "samples": 1, "bvh": bvh})
    primary("default/"+bvh, default, {"width": 400, "samples": 1, "bvh": bvh})
    primary("rain100k/"+bvh, rain, {"width": 480, "samples": 1, "bvh": bvh})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
from PIL import Image
img = timing("cornell", cornell, {"width": 1024, "samples": 64, "aTolerance": 0})
Image.fromarray(img).save(os.path.join(ROOT, "gpurun_out", "cornell_gpu.png"))
timing("cornell/ref-bvh", cornell, {"width": 1024, "samples": 64, "aTolerance": 0, "bvh": "reference"})
timing("spheres100 C1", spheres, {"width": 400, "samples": 16, "depth": 10, "aTolerance": 0})
img = timing("weekend C3", weekend, {"width": 1920, "samples": 16, "aTolerance": 0})
Image.fromarray(img).save(os.path.join(ROOT, "gpurun_out", "weekend_gpu.png"))
img = timing("rain100k C4", rain, {"width": 3840, "samples": 4, "aTolerance": 0})
Image.fromarray(img[::4, ::4].copy()).save(os.path.join(ROOT, "gpurun_out", "rain_gpu.png"))
timing("rain100k C4 refbvh", rain, {"width": 3840, "samples": 4, "aTolerance": 0, "bvh": "reference"})
img = timing("layered C5", layered, {"width": 2048, "samples": 16, "aTolerance": 0})
Image.fromarray(img[::2, ::2].copy()).save(os.path.join(ROOT, "gpurun_out", "layered_gpu.png"))
img = timing("default adaptive", default, {"width": 800, "samples": 100})
Image.fromarray(img).save(os.path.join(ROOT, "gpurun_out", "default_gpu.png"))
