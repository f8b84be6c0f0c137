"""Render orchestration — mirror of `src/raytracer.ts` with the parallel renderer replaced
by native dispatch (the seam SURVEY.md §8b names).

* `generateImageBuffer(sceneConfig, options)`  <- src/raytracer.ts:39-113
    options.parallel=False : `camera.render(pixelData)`            (raytracer.ts:56-59)
    options.parallel=True  : the image is split over GPUs instead of worker threads
                             (raytracer.ts:60-90): one Camera per visible device, each renders
                             its interleaved 16x16-tile set into the shared buffer, stats merged
                             with RenderStats.merge — the roles `divideIntoRegions` + workers +
                             SharedArrayBuffer play in the reference.
  Returns PNG bytes like the reference (PIL instead of sharp; PNG encoding is outside the
  hot path) or, with `options.raw=True`, the RGB8 array itself.
* `divideIntoRegions`                           <- src/raytracer.ts:185-205 (kept for callers
  that want the reference's row-strip partition, e.g. the CPU-baseline harness)
"""
from __future__ import annotations

import io
import math
import sys
import threading
from typing import Any, Dict, List, Optional

import numpy as np

from . import _native
from .camera import Camera, RenderStats, createCameraFromSceneData
from .scenes import generateSceneData


def divideIntoRegions(imageWidth: int, imageHeight: int, count: int) -> List[Dict[str, int]]:
    regionHeight = math.ceil(imageHeight / count)
    regions = []
    for i in range(count):
        startY = i * regionHeight
        height = min(regionHeight, imageHeight - startY)
        if height <= 0:
            break
        regions.append({"x": 0, "y": startY, "width": imageWidth, "height": height})
    return regions


def renderScene(sceneConfig: Dict[str, Any], options: Optional[Dict[str, Any]] = None):
    """Everything `generateImageBuffer` does before PNG encoding: returns (rgb8[H,W,3], RenderStats)."""
    options = options or {}
    parallel = options.get("parallel", False)
    verbose = options.get("verbose", False)
    sceneData = generateSceneData(sceneConfig)
    render = sceneConfig.get("render")
    ndev = _native.lib().rt_device_count()
    threads = options.get("threads") or ndev  # "threads" = number of GPUs here
    if not parallel or threads <= 1:
        with createCameraFromSceneData(sceneData, render) as camera:
            pixelData = np.zeros(camera.imageWidth * camera.imageHeight * camera.channels, np.uint8)
            stats = camera.render(pixelData)
            W, H = camera.imageWidth, camera.imageHeight
    else:
        threads = min(threads, max(ndev, 1))
        if verbose:
            print(f"Starting parallel render on {threads} GPUs", file=sys.stderr)
        cams = [createCameraFromSceneData(sceneData, {**(render or {}), "device": d, "partIndex": d, "partCount": threads})
                for d in range(threads)]
        W, H = cams[0].imageWidth, cams[0].imageHeight
        pixelData = np.zeros(W * H * 3, np.uint8)  # the SharedArrayBuffer of raytracer.ts:71-72
        results: List[Optional[RenderStats]] = [None] * threads
        errors: List[BaseException] = []

        def work(k: int) -> None:
            try:
                results[k] = cams[k].render(pixelData)  # disjoint tiles, no locks (renderWorker.ts:23-26)
            except BaseException as e:  # noqa: BLE001
                errors.append(e)

        ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for c in cams:
            c.close()
        if errors:
            raise errors[0]
        stats = RenderStats.merge([r for r in results if r is not None])
    if verbose:
        print(f"Adaptive sampling stats: avg={stats.samples['avg']:.2f}, min={stats.samples['min']}, max={stats.samples['max']}", file=sys.stderr)
        print(f"Ray bounce stats: avg={stats.bounces['avg']:.2f}, min={stats.bounces['min']}, max={stats.bounces['max']}", file=sys.stderr)
    return pixelData.reshape(H, W, 3), stats


def generateImageBuffer(sceneConfig: Optional[Dict[str, Any]] = None, options: Optional[Dict[str, Any]] = None):
    sceneConfig = sceneConfig or {"type": "default"}
    options = options or {}
    rgb, _stats = renderScene(sceneConfig, options)
    if rgb.size == 0:
        raise RuntimeError("Generated pixelData buffer is empty before calling sharp.")  # raytracer.ts:97-99
    if options.get("raw"):
        return rgb
    from PIL import Image  # PNG encode: sharp/libvips in the reference, outside the hot path

    buf = io.BytesIO()
    Image.fromarray(rgb, "RGB").save(buf, format="PNG")
    return buf.getvalue()
