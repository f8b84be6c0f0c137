"""All five BASELINE configs at full size on one GPU, every integrator layout, CPU oracle beside them.
Writes gpurun_out/<tag>_configs.json and a markdown table on stdout (development/report aid)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from mcp_raytracer_b200 import createCameraFromSceneData, measureFp32Peak

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
cpu_seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
peak, mhz = measureFp32Peak()
threads = max(1, (os.cpu_count() or 2) - 1)
KIND = {1: "megakernel", 2: "wavefront", 3: "sorted"}
rows = []
for wl in ("C1", "C2", "C3", "C4", "C5"):
    label, kind, sopts, ropts = bench.WORKLOADS[wl]
    t0 = time.perf_counter(); sd = bench.make_scene(kind, sopts); t_scene = time.perf_counter() - t0
    row = {"config": wl, "workload": label, "scene_gen_s": t_scene}
    for integ in ("auto", "megakernel", "sorted", "wavefront"):
        t0 = time.perf_counter()
        with createCameraFromSceneData(sd, dict(ropts, integrator=integ)) as cam:
            t_build = time.perf_counter() - t0
            rgb = np.zeros(cam.imageWidth * cam.imageHeight * 3, np.uint8)
            best = None
            for _ in range(3):
                st = cam.render(rgb)
                if best is None or st.deviceMs < best.deviceMs: best = st
            row.update({f"{integ}_ms": best.deviceMs, f"{integ}_mpaths": best.samples["total"] / best.deviceMs / 1e3,
                        f"{integ}_grays": best.rays / best.deviceMs / 1e6, f"{integ}_launches": best.kernelLaunches})
            if integ == "auto":
                row.update({"auto_choice": KIND.get(cam.info.integrator_kind, "?"),
                            "image": f"{cam.imageWidth}x{cam.imageHeight}", "spp": ropts["samples"], "paths": best.samples["total"], "rays": best.rays,
                            "bvh": {1: "reference", 2: "sah", 3: "list"}[cam.info.bvh_kind], "bvh_nodes": cam.info.n_bvh_nodes,
                            "build_ms": cam.info.build_ms, "create_s": t_build})
    c = bench.cpu_reference_run(sd, ropts, target_seconds=cpu_seconds, threads=threads)
    fpp = bench.algorithmic_flops(c["counters"], c["n_lights"], float(sd["camera"].get("aperture", 0))) / max(1, c["counters"]["paths"])
    row.update({"cpu_mpaths": c["mpaths_per_s"], "cpu_threads": threads, "cpu_sample_spp": c["spp"], "flops_per_path": fpp,
                "achieved_tflops": fpp * row["paths"] / (row["auto_ms"] * 1e-3) / 1e12, "fp32_peak_tflops": peak})
    row["roofline_frac"] = row["achieved_tflops"] / peak
    row["speedup_vs_cpu"] = row["auto_mpaths"] / row["cpu_mpaths"]
    rows.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"{tag}_configs.json"), "w"), indent=1)
print("| config | image @ spp | BVH (nodes) | AUTO = | ms | Mpaths/s | Grays/s | megakernel ms | sorted ms | wavefront ms | create s | CPU port Mpaths/s (threads) | GPU/CPU | kflop/path | TFLOP/s (frac of %.1f) |" % peak)
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r['config']} | {r['image']} @ {r['spp']} | {r['bvh']} ({r['bvh_nodes']}) | {r['auto_choice']} | {r['auto_ms']:.1f} | {r['auto_mpaths']:.0f} | {r['auto_grays']:.2f} | "
          f"{r['megakernel_ms']:.1f} | {r['sorted_ms']:.1f} | {r['wavefront_ms']:.1f} | {r['create_s']:.3f} | {r['cpu_mpaths']:.2f} ({r['cpu_threads']}) | {r['speedup_vs_cpu']:.0f}x | "
          f"{r['flops_per_path']/1e3:.2f} | {r['achieved_tflops']:.1f} ({r['roofline_frac']:.3f}) |")
