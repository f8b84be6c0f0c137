"""CLI benchmark — mirror of the reference's `src/benchmark.ts` flags (`node dist/src/index.js --benchmark …`).

    python -m mcp_raytracer_b200.benchmark --cornell spheres -w 1024 -s 1024 --at 0 -i 3 -o out.png

Same options, same stderr report, same timing scope as the reference (`Date.now()` around the whole
`generateImageBuffer`: scene generation + build + render + PNG encode, src/benchmark.ts:116-186).
"""
from __future__ import annotations

import sys
import time

from .raytracer import generateImageBuffer

HELP = """
Raytracer Performance Benchmark

Options:
  --width, -w <number>     Image width (default: 400)
  --samples, -s <number>   Samples per pixel (default: 100)
  --output, -o <file>      Output PNG file
  --iterations, -i <num>   Number of iterations to run (default: 1)
  --spheres <number>       Number of random spheres to generate (spheres scene)
  --rain <number>          Number of metallic raindrops to generate (rain scene)
  --cornell <variant>      Cornell box scene ('spheres' or 'empty')
  --seed <number>          Random seed for deterministic scene generation
  --adaptive-tolerance, --at <n>  Convergence tolerance for adaptive sampling (default: 0.05)
  --adaptive-batch, --ab <n>      Number of samples to process in one batch (default: 10)
  --parallel, -p           Use all visible GPUs (the reference: worker threads)
  --threads, -t <number>   Number of GPUs to use
  --mode, -m <mode>        default | bounces | samples
  --help, -h               Show this help
"""


def runRaytracerBenchmark(argv=None) -> int:
    args = list(sys.argv[1:] if argv is None else argv)
    render, spheres, rain, cornell = {}, {}, {}, {}
    gen = {"verbose": True, "output": None, "iterations": 1, "parallel": False, "threads": None, "sceneType": "default"}
    i = 0
    while i < len(args):
        a = args[i]

        def nxt():
            nonlocal i
            i += 1
            return args[i]

        if a in ("--width", "-w"): render["width"] = int(nxt())
        elif a in ("--samples", "-s"): render["samples"] = int(nxt())
        elif a in ("--output", "-o"): gen["output"] = nxt()
        elif a in ("--iterations", "-i"): gen["iterations"] = int(nxt())
        elif a == "--spheres": spheres["count"] = int(nxt()); gen["sceneType"] = "spheres"
        elif a == "--seed":
            v = int(nxt()); spheres["seed"] = v; rain["seed"] = v
        elif a in ("--adaptive-tolerance", "--at"): render["aTolerance"] = float(nxt())
        elif a in ("--adaptive-batch", "--ab"): render["aBatch"] = int(nxt())
        elif a in ("--parallel", "-p"): gen["parallel"] = True
        elif a in ("--threads", "-t"): gen["threads"] = int(nxt())
        elif a in ("--mode", "-m"):
            mode = nxt()
            if mode not in ("default", "bounces", "samples"):
                print(f"Invalid render mode: {mode}", file=sys.stderr)
                return 1
            render["mode"] = mode
        elif a == "--rain": rain["count"] = int(nxt()); gen["sceneType"] = "rain"
        elif a == "--cornell":
            v = nxt()
            if v not in ("spheres", "empty"):
                print(f"Invalid Cornell variant: {v}. Use 'spheres' or 'empty'.", file=sys.stderr)
                return 1
            cornell["variant"] = v; gen["sceneType"] = "cornell"
        elif a in ("--help", "-h"):
            print(HELP)
            return 0
        i += 1
    e = lambda *x: print(*x, file=sys.stderr)  # noqa: E731
    e("Running raytracer with options:")
    e(f"  Dimensions: {render.get('width')}")
    e(f"  Samples: {render.get('samples')}")
    e(f"  Scene type: {gen['sceneType']}")
    e(f"  Output: {gen['output'] or 'none (image discarded)'}")
    cfg = {"type": gen["sceneType"], "render": render}
    if gen["sceneType"] == "spheres": cfg["options"] = spheres
    elif gen["sceneType"] == "rain": cfg["options"] = rain
    elif gen["sceneType"] == "cornell": cfg["options"] = cornell
    total = 0.0
    try:
        for it in range(1, gen["iterations"] + 1):
            t0 = time.time()
            png = generateImageBuffer(cfg, {"parallel": gen["parallel"], "threads": gen["threads"], "verbose": True})
            ms = (time.time() - t0) * 1e3
            total += ms
            e(f"Render time: {ms:.0f}ms" if gen["iterations"] == 1 else f"Iteration {it} render time: {ms:.0f}ms")
            if gen["output"] and it == gen["iterations"]:
                with open(gen["output"], "wb") as f:
                    f.write(png)
                e(f"Image saved to {gen['output']}")
    except Exception as ex:  # src/benchmark.ts:173-176
        e(f"Error generating image: {ex}")
        return 1
    if gen["iterations"] > 1:
        e(f"Average render time: {total / gen['iterations']:.2f}ms")
    return 0


if __name__ == "__main__":
    sys.exit(runRaytracerBenchmark())
