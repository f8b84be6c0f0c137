"""Times library variants side by side (csrc/ab_<name>.so from scripts/build_variants.sh; 'base' = the shipped library).
Each (variant, workload) runs in its own process with RT_B200_LIB set; prints ms (best of reps), Mpaths/s and an image hash.
usage: gpu_ab.py <workload:spp[:width]>[,...] variant [variant ...] [-- key=value render options]"""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mcp_raytracer_b200", "csrc")

def child(wl, spp, width, reps, extra):
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    from mcp_raytracer_b200 import createCameraFromSceneData
    label, kind, sopts, ropts = bench.WORKLOADS[wl]
    sd = bench.make_scene(kind, sopts)
    o = dict(ropts, samples=spp, **extra)
    if width: o["width"] = width
    with createCameraFromSceneData(sd, o) as cam:
        rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
        best = None
        for _ in range(reps):
            st = cam.render(rgb)
            if best is None or st.deviceMs < best.deviceMs: best = st
        print(json.dumps({"ms": best.deviceMs, "mpaths": best.samples["total"] / best.deviceMs / 1e3, "rays": best.rays,
                          "hash": hashlib.sha1(rgb.tobytes()).hexdigest()[:12], "image": f"{cam.imageWidth}x{cam.imageHeight}"}))

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        extra = dict(kv.split("=", 1) for kv in sys.argv[6:])
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), extra)
        sys.exit(0)
    args = sys.argv[1:]
    extra = []
    if "--" in args:
        k = args.index("--"); extra = args[k + 1:]; args = args[:k]
    wls = []
    for spec in args[0].split(","):
        p = spec.split(":")
        wls.append((p[0], int(p[1]), int(p[2]) if len(p) > 2 else 0))
    print(f"{'variant':24s} " + " ".join(f"{w}@{s}{'/' + str(x) if x else '':>6s} ms   Mpaths/s  hash        " for w, s, x in wls))
    for v in args[1:]:
        env = dict(os.environ)
        if v != "base":
            env["RT_B200_LIB"] = os.path.join(CSRC, f"ab_{v}.so")
        row = []
        for w, s, x in wls:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", w, str(s), str(x), "4"] + extra, env=env, capture_output=True, text=True)
            try:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                row.append(f"{j['ms']:14.2f} {j['mpaths']:9.0f}  {j['hash']}")
            except Exception:
                row.append(f"FAILED: {(r.stderr or r.stdout).strip().splitlines()[-1][:60] if (r.stderr or r.stdout).strip() else '?'}")
        print(f"{v:24s} " + " ".join(row), flush=True)
