"""Scene generators — host-side mirror of the reference's `src/scenes/` API.

Same function names, option names, defaults and draw order as the reference so that the
same seed yields the same `SceneData` (plain dicts shaped like `src/scenes/sceneData.ts:8-110`):

* `SeededRandom`                 <- src/scenes/scenes-utils.ts:8-59 (mulberry32)
* `generateDefaultSceneData`     <- src/scenes/scenes-default.ts:8-82
* `generateSpheresSceneData`     <- src/scenes/scenes-spheres.ts:27-119
* `generateRainSceneData`        <- src/scenes/scenes-rain.ts:19-145
* `generateCornellSceneData`     <- src/scenes/scenes-cornell.ts:19-135
* `generateSceneData`            <- src/scenes/scenes.ts:42-50

plus the two BASELINE.json configs that are not reference generators and are therefore
expressed as `type: 'custom'` SceneData: `generateWeekendFinalSceneData`
(C3) and `generateLayeredMixedSceneData` (C5).

JS numbers are IEEE doubles, Python floats are too; `Math.pow` / `Math.log10` are libm's here
(V8 ships fdlibm ports) so a last-ulp difference is possible in principle — the SceneData
dict, not the seed, is the shared input of the oracle and the GPU path.
"""
from __future__ import annotations

import math
import random as _pyrandom
from typing import Any, Dict, List, Optional

import numpy as np

from .scene_data import RT_MAT_LAMBERT, RT_MAT_METAL, SoAMaterials, SoAObjects

SceneData = Dict[str, Any]
# generators emit array-backed objects / materials (SoAObjects, SoAMaterials) from this many objects on: 100 000 dicts cost
# ~0.4 s to build and ~0.2 s to flatten again, the arrays ~10 ms; smaller scenes stay plain lists of records
SOA_THRESHOLD = 4096

_M32 = 0xFFFFFFFF


def _imul(a: int, b: int) -> int:
    return (a * b) & _M32


def _f32(x: float) -> float:
    return float(np.float32(x))


class SeededRandom:
    """mulberry32 exactly as src/scenes/scenes-utils.ts:18-23 computes it.

    The reference keeps `seed` as an unwrapped JS double and lets `^`, `>>>` and
    `Math.imul` coerce to 32 bits; integer arithmetic mod 2**32 is the same function while
    the double stays exact (< 2**53, i.e. ~4.9 M draws — SURVEY.md App. A.9).
    """

    def __init__(self, seed: Optional[int] = None):
        self.seed = int(seed) if seed is not None else int(_pyrandom.random() * 2147483647)

    def next(self) -> float:
        self.seed += 0x6D2B79F5
        t = self.seed & _M32
        t = _imul(t ^ (t >> 15), t | 1)
        t ^= (t + _imul(t ^ (t >> 7), t | 61)) & _M32
        return ((t ^ (t >> 14)) & _M32) / 4294967296

    def next_batch(self, n: int) -> np.ndarray:
        """`n` consecutive `next()` values, vectorised (the state is a pure counter)."""
        k = np.arange(1, n + 1, dtype=np.uint64)
        s = (np.uint64(self.seed & _M32) + k * np.uint64(0x6D2B79F5)) & np.uint64(_M32)
        self.seed += 0x6D2B79F5 * n
        m = np.uint64(_M32)
        t = s
        t = ((t ^ (t >> np.uint64(15))) * (t | np.uint64(1))) & m
        t = t ^ ((t + (((t ^ (t >> np.uint64(7))) * (t | np.uint64(61))) & m)) & m)
        return ((t ^ (t >> np.uint64(14))) & m).astype(np.float64) / 4294967296.0

    def nextInRange(self, lo: float, hi: float) -> float:
        return lo + (hi - lo) * self.next()

    def randomInUnitSphere(self):
        # scenes-utils.ts:35-50: goes through Vec3.create => components are FP32-rounded
        while True:
            x = _f32(self.nextInRange(-1, 1))
            y = _f32(self.nextInRange(-1, 1))
            z = _f32(self.nextInRange(-1, 1))
            if x * x + y * y + z * z < 1:
                return (x, y, z)

    def randomColor(self):
        return [self.next(), self.next(), self.next()]


# --------------------------------------------------------------------------------------
# default scene — src/scenes/scenes-default.ts:8-82
# --------------------------------------------------------------------------------------
def generateDefaultSceneData() -> SceneData:
    materials = [
        {"id": "ground", "material": {"type": "lambert", "color": [0.4, 0.4, 0.0]}},
        {"id": "blue", "material": {"type": "lambert", "color": [0.1, 0.1, 0.9]}},
        {"id": "glass", "material": {"type": "glass", "ior": 1.5}},
        {"id": "silver", "material": {"type": "metal", "color": [0.8, 0.8, 0.8], "fuzz": 0.0}},
        {"id": "gold", "material": {"type": "metal", "color": [0.8, 0.6, 0.2], "fuzz": 0.5}},
        {
            "id": "layered-paint",
            "material": {
                "type": "layered",
                "outer": {"type": "glass", "ior": 1.5},
                "inner": {"type": "lambert", "color": [0.7, 0.3, 0.3]},
            },
        },
        {"id": "sun-light", "material": {"type": "light", "emit": [15.0, 14.0, 13.0]}},
    ]
    objects: List[Dict[str, Any]] = [
        {"type": "plane", "pos": [0, 0, 0], "u": [1, 0, 0], "v": [0, 0, 1], "material": "ground"},
        {"type": "sphere", "pos": [0, 0.5, -1], "r": 0.5, "material": "layered-paint"},
        {"type": "sphere", "pos": [-1, 0.5, -1], "r": 0.5, "material": "silver"},
        {"type": "sphere", "pos": [1, 0.5, -1], "r": 0.5, "material": "gold"},
        {"type": "sphere", "pos": [0.5, 0.25, -0.5], "r": 0.25, "material": "glass"},
    ]
    for v in ({"r": 0.25, "material": "glass"}, {"r": -0.24, "material": "glass"}, {"r": 0.20, "material": "blue"}):
        objects.append({"type": "sphere", "pos": [-0.5, 0.25, -0.5], **v})
    objects.append(
        {"type": "quad", "pos": [-2, 3, 0], "u": [1, 0, 0], "v": [0, -0.707, -0.707], "material": "sun-light", "light": True}
    )
    objects.append({"type": "sphere", "pos": [30, 30.5, 15], "r": 10, "material": "sun-light", "light": True})
    return {
        "camera": {
            "vfov": 40, "aperture": 0.05, "focus": 2.8,
            "from": [0, 0.75, 2], "at": [0, 0.5, -1], "up": [0, 1, 0],
            "background": {"type": "gradient", "top": [1, 1, 1], "bottom": [0.5, 0.7, 1.0]},
        },
        "materials": materials,
        "objects": objects,
        "metadata": {
            "name": "Default Scene",
            "description": "A scene with various spheres demonstrating different materials including layered, mixed, and basic materials",
            "version": "2.0",
        },
    }


# --------------------------------------------------------------------------------------
# spheres scene — src/scenes/scenes-spheres.ts
# --------------------------------------------------------------------------------------
def _checkOverlap(center, radius, placed) -> bool:  # scenes-spheres.ts:124-137
    for c, r in placed:
        dx = center[0] - c[0]
        dy = center[1] - c[1]
        dz = center[2] - c[2]
        if math.sqrt(dx * dx + dy * dy + dz * dz) < (radius + r):
            return True
    return False


def _randomPointInSphere(center, radius, rnd: SeededRandom):  # scenes-spheres.ts:142-156
    ux, uy, uz = rnd.randomInUnitSphere()
    distanceFactor = math.pow(rnd.next(), 1 / 3) * radius
    return [center[0] + ux * distanceFactor, center[1] + uy * distanceFactor, center[2] + uz * distanceFactor]


def _generateRandomMaterialData(rnd: SeededRandom):  # scenes-spheres.ts:161-182
    materialType = rnd.next()
    if materialType < 0.6:
        return {"type": "lambert", "color": [rnd.next(), rnd.next(), rnd.next()]}
    elif materialType < 0.9:
        fuzz = rnd.next() * 0.5
        return {"type": "metal", "color": [rnd.next(), rnd.next(), rnd.next()], "fuzz": fuzz}
    else:
        ior = 1.3 + rnd.next() * 1.2
        return {"type": "glass", "ior": ior}


def generateSpheresSceneData(sceneOpts: Optional[Dict[str, Any]] = None) -> SceneData:
    opts = {
        "count": 10,
        "centerPoint": [0, 0, -2],
        "radius": 1.25,
        "minSphereRadius": 0.1,
        "maxSphereRadius": 0.2,
        "seed": int(_pyrandom.random() * 2147483647),
    }
    opts.update({k: v for k, v in (sceneOpts or {}).items() if v is not None})
    rnd = SeededRandom(opts["seed"])
    scaleFactor = max(opts["minSphereRadius"], 1 - math.log10(opts["count"] + 1) / 4)
    adjustedMaxRadius = opts["maxSphereRadius"] * scaleFactor
    materials, objects, placed = [], [], []
    attempts = 0
    maxAttempts = opts["count"] * 100
    created = 0
    while created < opts["count"] and attempts < maxAttempts:
        attempts += 1
        radius = adjustedMaxRadius
        center = _randomPointInSphere(opts["centerPoint"], opts["radius"], rnd)
        if _checkOverlap(center, radius, placed):
            continue
        materialId = f"sphere-{created}"
        materials.append({"id": materialId, "material": _generateRandomMaterialData(rnd)})
        objects.append({"type": "sphere", "pos": center, "r": radius, "material": materialId})
        placed.append((center, radius))
        created += 1
    return {
        "camera": {
            "vfov": 40, "aperture": 0.0, "focus": 1.0,
            "from": [0, 0, 2], "at": list(opts["centerPoint"]), "up": [0, 1, 0],
            "background": {"type": "gradient", "top": [0.5, 0.7, 1.0], "bottom": [1.0, 1.0, 1.0]},
        },
        "materials": materials,
        "objects": objects,
        "metadata": {
            "name": "Random Spheres",
            "description": f"Scene with {created} randomly placed spheres (seed: {opts['seed']})",
            "version": "2.0",
        },
    }


# --------------------------------------------------------------------------------------
# rain scene — src/scenes/scenes-rain.ts
# --------------------------------------------------------------------------------------
def generateRainSceneData(sceneOpts: Optional[Dict[str, Any]] = None) -> SceneData:
    o = {
        "count": 50, "sphereRadius": 0.05, "width": 4, "height": 3, "depth": 2,
        "centerPoint": [0, 0, -2], "metalFuzz": 0.1, "groundSphere": True,
        "groundY": -100.5, "groundRadius": 100, "seed": int(_pyrandom.random() * 2147483647),
    }
    o.update({k: v for k, v in (sceneOpts or {}).items() if v is not None})
    rnd = SeededRandom(o["seed"])
    materials: List[Dict[str, Any]] = []
    objects: List[Dict[str, Any]] = []
    if o["groundSphere"]:
        materials.append({"id": "ground", "material": {"type": "lambert", "color": [0.1, 0.1, 0.1]}})
        objects.append({"type": "sphere", "pos": [0, o["groundY"], 0], "r": o["groundRadius"], "material": "ground"})
    n = int(math.ceil(math.pow(o["count"], 1 / 3)))
    xs, ys, zs = o["width"] / n, o["height"] / n, o["depth"] / n
    startX = o["centerPoint"][0] - (n * xs) / 2 + xs / 2
    startY = o["centerPoint"][1] - (n * ys) / 2 + ys / 2
    startZ = o["centerPoint"][2] - (n * zs) / 2 + zs / 2
    # scenes-rain.ts:80-93 — x outer, y, z inner; three draws per cell in x,y,z order
    cells = n * n * n
    u = rnd.next_batch(3 * cells).reshape(cells, 3)
    ix, iy, iz = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    px = (startX + ix.reshape(-1) * xs) + (u[:, 0] - 0.5) * xs * 0.3
    py = (startY + iy.reshape(-1) * ys) + (u[:, 1] - 0.5) * ys * 0.3
    pz = (startZ + iz.reshape(-1) * zs) + (u[:, 2] - 0.5) * zs * 0.3
    positions = np.stack([px, py, pz], axis=1)
    # Fisher-Yates, scenes-rain.ts:152-158
    order = list(range(cells))
    if cells > 1:
        r = rnd.next_batch(cells - 1)
        for k, i in enumerate(range(cells - 1, 0, -1)):
            j = int(math.floor(r[k] * (i + 1)))
            order[i], order[j] = order[j], order[i]
    selected = positions[order[: o["count"]]]
    m = rnd.next_batch(2 * len(selected)).reshape(-1, 2)
    brightness = 0.7 + m[:, 0] * 0.3
    fuzz = o["metalFuzz"] * m[:, 1]
    if len(selected) >= SOA_THRESHOLD:
        # large scene: straight into arrays (SoAObjects / SoAMaterials read like the reference's lists of records, and
        # FlatScene hands them to the C ABI without touching 100 000 dicts): same values, same order
        k = len(selected)
        g = 1 if o["groundSphere"] else 0
        mat_type = np.concatenate([np.full(g, RT_MAT_LAMBERT, np.uint8), np.full(k, RT_MAT_METAL, np.uint8)])
        mat_color = np.concatenate([np.full((g, 3), 0.1), np.repeat(brightness[:, None], 3, axis=1)])
        mat_param = np.concatenate([np.zeros(g), fuzz])
        ids = (lambda i: "ground" if i == 0 else f"rain-{i - 1}") if g else (lambda i: f"rain-{i}")
        materials = SoAMaterials(mat_type, mat_color, mat_param, np.full((g + k, 2), -1, np.int32), np.arange(g + k, dtype=np.int32), ids)
        pos = np.concatenate([np.array([[0.0, o["groundY"], 0.0]])[:g], selected])
        radius = np.concatenate([np.full(g, float(o["groundRadius"])), np.full(k, float(o["sphereRadius"]))])
        objects = SoAObjects(np.zeros(g + k, np.uint8), pos, np.zeros((g + k, 3)), np.zeros((g + k, 3)), radius,
                             np.arange(g + k, dtype=np.int32), np.zeros(g + k, np.uint8), materials)
    else:
        for i, pos in enumerate(selected):
            materialId = f"rain-{i}"
            b = float(brightness[i])
            materials.append({"id": materialId, "material": {"type": "metal", "color": [b, b, b], "fuzz": float(fuzz[i])}})
            objects.append({"type": "sphere", "pos": [float(pos[0]), float(pos[1]), float(pos[2])], "r": o["sphereRadius"], "material": materialId})
    return {
        "camera": {
            "vfov": 40, "aperture": 0.0, "focus": 1.0,
            "from": [0, 0, 2], "at": list(o["centerPoint"]), "up": [0, 1, 0],
            "background": {"type": "gradient", "top": [0.5, 0.7, 1.0], "bottom": [1.0, 1.0, 1.0]},
        },
        "materials": materials,
        "objects": objects,
        "metadata": {
            "name": "Rain Scene",
            "description": f"Scene with {len(selected)} metallic rain spheres (seed: {o['seed']})",
            "version": "2.0",
        },
    }


# --------------------------------------------------------------------------------------
# cornell scene — src/scenes/scenes-cornell.ts
# --------------------------------------------------------------------------------------
def generateCornellSceneData(sceneOpts: Optional[Dict[str, Any]] = None) -> SceneData:
    options = {"variant": "spheres"}
    options.update({k: v for k, v in (sceneOpts or {}).items() if v is not None})
    boxSize = 2.0
    h = boxSize / 2
    materials = [
        {"id": "red", "material": {"type": "lambert", "color": [0.65, 0.05, 0.05]}},
        {"id": "green", "material": {"type": "lambert", "color": [0.12, 0.45, 0.15]}},
        {"id": "white", "material": {"type": "lambert", "color": [0.73, 0.73, 0.73]}},
        {"id": "light", "material": {"type": "light", "emit": [15, 15, 15]}},
    ]
    if options["variant"] == "spheres":
        materials.append({"id": "sphere-white", "material": {"type": "lambert", "color": [0.6, 0.6, 0.6]}})
        materials.append({"id": "sphere-glass", "material": {"type": "glass", "ior": 1.5}})
    lightSize = boxSize * 0.3
    objects: List[Dict[str, Any]] = [
        {"type": "quad", "pos": [-h, -h, -h], "u": [0, boxSize, 0], "v": [0, 0, boxSize], "material": "red"},
        {"type": "quad", "pos": [h, -h, h], "u": [0, boxSize, 0], "v": [0, 0, -boxSize], "material": "green"},
        {"type": "quad", "pos": [-h, -h, -h], "u": [boxSize, 0, 0], "v": [0, boxSize, 0], "material": "white"},
        {"type": "quad", "pos": [-h, -h, -h], "u": [boxSize, 0, 0], "v": [0, 0, boxSize], "material": "white"},
        {"type": "quad", "pos": [-h, h, h], "u": [boxSize, 0, 0], "v": [0, 0, -boxSize], "material": "white"},
        {
            "type": "quad", "pos": [-lightSize / 2, h - 0.01, -lightSize / 2],
            "u": [lightSize, 0, 0], "v": [0, 0, lightSize], "material": "light", "light": True,
        },
    ]
    if options["variant"] == "spheres":
        r = 0.3
        objects.append({"type": "sphere", "pos": [-h * 0.4, -h + r, -h * 0.3], "r": r, "material": "sphere-white"})
        objects.append({"type": "sphere", "pos": [h * 0.4, -h + r, h * 0.3], "r": r, "material": "sphere-glass"})
    return {
        "camera": {
            "vfov": 40, "aperture": 0.0, "focus": 1.0,
            "from": [0, 0, h * 4], "at": [0, 0, 0], "up": [0, 1, 0],
            "background": {"type": "gradient", "top": [0, 0, 0], "bottom": [0, 0, 0]},
        },
        "render": {"aspect": 1.0, "roulette": True, "rouletteDepth": 5},
        "materials": materials,
        "objects": objects,
        "metadata": {
            "name": f"Cornell Box ({options['variant']})",
            "description": "Cornell box scene with "
            + ("two spheres inside" if options["variant"] == "spheres" else "empty interior"),
            "version": "2.0",
        },
    }


# --------------------------------------------------------------------------------------
# BASELINE.json configs that are not reference generators (custom SceneData)
# --------------------------------------------------------------------------------------
def generateWeekendFinalSceneData(sceneOpts: Optional[Dict[str, Any]] = None) -> SceneData:
    """C3: "Ray Tracing in One Weekend" final scene as `type:'custom'` SceneData.

    All randoms come from `SeededRandom(seed)` in a fixed order per grid cell:
    choose, ox, oz, then the material's own draws.
    """
    o = {"seed": 1}
    o.update(sceneOpts or {})
    rnd = SeededRandom(o["seed"])
    materials = [{"id": "ground", "material": {"type": "lambert", "color": [0.5, 0.5, 0.5]}}]
    objects: List[Dict[str, Any]] = [{"type": "sphere", "pos": [0, -1000, 0], "r": 1000, "material": "ground"}]
    k = 0
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose = rnd.next()
            cx = a + 0.9 * rnd.next()
            cz = b + 0.9 * rnd.next()
            if math.sqrt((cx - 4) ** 2 + (0.2 - 0.2) ** 2 + cz**2) <= 0.9:
                continue
            mid = f"small-{k}"
            k += 1
            if choose < 0.8:
                c = [rnd.next() * rnd.next(), rnd.next() * rnd.next(), rnd.next() * rnd.next()]
                mat = {"type": "lambert", "color": c}
            elif choose < 0.95:
                c = [0.5 + 0.5 * rnd.next(), 0.5 + 0.5 * rnd.next(), 0.5 + 0.5 * rnd.next()]
                mat = {"type": "metal", "color": c, "fuzz": 0.5 * rnd.next()}
            else:
                mat = {"type": "glass", "ior": 1.5}
            materials.append({"id": mid, "material": mat})
            objects.append({"type": "sphere", "pos": [cx, 0.2, cz], "r": 0.2, "material": mid})
    materials += [
        {"id": "big-glass", "material": {"type": "glass", "ior": 1.5}},
        {"id": "big-lambert", "material": {"type": "lambert", "color": [0.4, 0.2, 0.1]}},
        {"id": "big-metal", "material": {"type": "metal", "color": [0.7, 0.6, 0.5], "fuzz": 0.0}},
    ]
    objects += [
        {"type": "sphere", "pos": [0, 1, 0], "r": 1.0, "material": "big-glass"},
        {"type": "sphere", "pos": [-4, 1, 0], "r": 1.0, "material": "big-lambert"},
        {"type": "sphere", "pos": [4, 1, 0], "r": 1.0, "material": "big-metal"},
    ]
    return {
        "camera": {
            "vfov": 20, "aperture": 0.1, "focus": 10.0,
            "from": [13, 2, 3], "at": [0, 0, 0], "up": [0, 1, 0],
            # names are inverted w.r.t. the screen in the reference:
            # `top` is what a down-pointing ray sees.
            "background": {"type": "gradient", "top": [1, 1, 1], "bottom": [0.5, 0.7, 1.0]},
        },
        "render": {"aspect": 16 / 9, "depth": 50},
        "materials": materials,
        "objects": objects,
        "metadata": {"name": "Weekend final", "description": f"{len(objects)} spheres (seed: {o['seed']})", "version": "2.0"},
    }


def generateLayeredMixedSceneData(sceneOpts: Optional[Dict[str, Any]] = None) -> SceneData:
    """C5: Cornell shell + quad light + 7x7 grid of r=0.1 spheres cycling five composite
    materials that together exercise every Layered/Mixed nesting the reference's material
    tests build (tests/materials/layeredMaterial.test.ts:34-42, mixedMaterial.test.ts)."""
    base = generateCornellSceneData({"variant": "empty"})
    materials = base["materials"]
    glass = {"type": "glass", "ior": 1.5}
    materials += [
        {"id": "paint", "material": {"type": "layered", "outer": glass, "inner": {"type": "lambert", "color": [0.7, 0.3, 0.3]}}},
        {"id": "coated-metal", "material": {"type": "layered", "outer": glass, "inner": {"type": "metal", "color": [0.8, 0.8, 0.8], "fuzz": 0.1}}},
        {"id": "plastic", "material": {"type": "mixed", "diff": {"type": "lambert", "color": [0.2, 0.4, 0.8]}, "spec": {"type": "metal", "color": [0.9, 0.9, 0.9], "fuzz": 0.05}, "weight": 0.7}},
        {"id": "frosted", "material": {"type": "mixed", "diff": glass, "spec": {"type": "lambert", "color": [0.8, 0.8, 0.3]}, "weight": 0.4}},
        {"id": "coated-plastic", "material": {"type": "layered", "outer": {"type": "glass", "ior": 1.33}, "inner": "plastic"}},
    ]
    cycle = ["paint", "coated-metal", "plastic", "frosted", "coated-plastic"]
    objects = base["objects"]
    k = 0
    for a in range(7):
        for b in range(7):
            x = -0.75 + 0.25 * a
            z = -0.75 + 0.25 * b
            objects.append({"type": "sphere", "pos": [x, -0.9, z], "r": 0.1, "material": cycle[k % len(cycle)]})
            k += 1
    base["metadata"] = {"name": "Layered/mixed grid", "description": "Cornell shell + 49 composite-material spheres", "version": "2.0"}
    return base


# --------------------------------------------------------------------------------------
# src/scenes/scenes.ts:42-50
# --------------------------------------------------------------------------------------
def generateSceneData(sceneConfig: Dict[str, Any]) -> SceneData:
    t = sceneConfig.get("type")
    if t == "default":
        return generateDefaultSceneData()
    if t == "spheres":
        return generateSpheresSceneData(sceneConfig.get("options"))
    if t == "rain":
        return generateRainSceneData(sceneConfig.get("options"))
    if t == "cornell":
        return generateCornellSceneData(sceneConfig.get("options"))
    if t == "custom":
        return sceneConfig["data"]
    raise ValueError(f"Unknown scene type: {t}")
