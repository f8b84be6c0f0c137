"""generateImageBuffer wall time for a stream of requests, PNG encode on and off the critical path.
usage: gpu_png_overlap.py [requests] [width] [spp] [encoders]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcp_raytracer_b200 import ImagePipeline, generateImageBuffer, renderScene, encodePng
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
width = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 64
enc = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cfgs = [{"type": "cornell", "render": {"width": width, "samples": spp, "aTolerance": 0, "seed": k}} for k in range(n)]
generateImageBuffer(cfgs[0])  # warm-up: CUDA context, module load, device cache
t0 = time.perf_counter(); rgb, _ = renderScene(cfgs[0]); t1 = time.perf_counter(); png = encodePng(rgb); t2 = time.perf_counter()
print(f"one request: render {1e3*(t1-t0):.1f} ms, PNG encode {1e3*(t2-t1):.1f} ms ({len(png)/1e6:.2f} MB)")
t0 = time.perf_counter(); a = [generateImageBuffer(c) for c in cfgs]; seq = time.perf_counter() - t0
with ImagePipeline(encoders=enc) as pipe:
    t0 = time.perf_counter(); b = [f.result() for f in [pipe.submit(c) for c in cfgs]]; ovl = time.perf_counter() - t0
assert a == b, "pipelined PNGs differ from the sequential ones"
print(f"{n} requests Cornell {width}x{width} @{spp}spp: sequential generateImageBuffer {1e3*seq/n:.1f} ms/image, "
      f"ImagePipeline({enc} encoder threads) {1e3*ovl/n:.1f} ms/image, x{seq/ovl:.2f}; identical bytes")
