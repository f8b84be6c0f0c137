import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mcp_raytracer_b200 import *
from oracle_binding import OracleCamera
np.set_printoptions(linewidth=200)
sd = generateCornellSceneData()
for bvh in ("list", "reference", "sah"):
    opts = {"width": 32, "samples": 1, "bvh": bvh}
    with createCameraFromSceneData(sd, opts) as cam:
        ids, t, nrm, ff = cam.tracePrimary()
        print(bvh, cam.info.bvh_kind, cam.info.n_bvh_nodes)
    oc = OracleCamera(sd, opts)
    oids, ot, onrm, off = oc.trace_primary()
    print("mismatch", (ids != oids).sum())
    print(ids[::4, ::2]); print(oids[::4, ::2])
    print(t[16, ::4]); print(ot[16, ::4])
