"""Per-function GPU parity: ONE device function of the render path at a time (rt_debug_* hooks of the C ABI,
the same code the render kernels inline) against the CPU oracle's counterpart, both driven with the SAME explicit
`Math.random()` sequence — on the reference's own Jest vectors first, then on seeded batches that reach the rare
branches whole-image tests cannot see (total internal reflection, metal absorption, layered-over-metal, mixed
fallbacks, sphere light seen from inside, lights behind the surface).

Run on the B200 box with `-m gpu`.  Nothing here reads /root/reference; every vector cites the reference test it
was transcribed from.
"""
import math
import zlib

import numpy as np
import pytest

import oracle_binding as ob
from mcp_raytracer_b200 import createCameraFromSceneData, generateCornellSceneData, generateDefaultSceneData, generateWeekendFinalSceneData
from mcp_raytracer_b200.scene_data import FlatScene

pytestmark = pytest.mark.gpu

U24 = 1.0 / 16777216.0


def u24(rng, shape):
    """Uniforms exactly as the device stream produces them: k / 2^24 (so float and double see the same number)."""
    return rng.integers(0, 1 << 24, size=shape).astype(np.float64) * U24


def unit(v):
    v = np.asarray(v, np.float64)
    return v / np.linalg.norm(v)


def f32(v):
    return np.asarray(v, np.float32).astype(np.float64)


def mat_scene(material, extra_objects=()):
    return {"camera": {"vfov": 40, "from": [0, 0, 5], "at": [0, 0, 0]}, "materials": [{"id": "m", "material": material}],
            "objects": [{"type": "sphere", "pos": [0, 0, 0], "r": 1, "material": "m"}, *extra_objects]}


class Pair:
    """The same material on both sides: a GPU camera (device functions) and the oracle's flattened scene."""

    def __init__(self, material):
        self.sd = mat_scene(material)
        self.cam = createCameraFromSceneData(self.sd, {"width": 16, "samples": 1})
        self.flat = FlatScene(self.sd)
        self.root = int(self.flat.obj_material[0])

    def close(self):
        self.cam.close()

    def both(self, ro, rd, p, n, front, uniforms):
        """Runs n records on the GPU in one call and one by one on the oracle; returns (gpu dict, list of oracle tuples)."""
        ro, rd, p, n = (np.atleast_2d(np.asarray(x, np.float64)) for x in (ro, rd, p, n))
        cnt = max(len(ro), len(rd), len(p), len(n), len(np.atleast_2d(uniforms)))
        bc = lambda a: np.broadcast_to(a, (cnt, 3))  # noqa: E731
        ro, rd, p, n = bc(ro), bc(rd), bc(p), bc(n)
        uniforms = np.broadcast_to(np.atleast_2d(np.asarray(uniforms, np.float64)), (cnt, np.atleast_2d(uniforms).shape[1]))
        front = np.broadcast_to(np.asarray(front, np.int32), (cnt,))
        g = self.cam.debugScatter(0, ro, rd, p, n, front, uniforms)
        o = [ob.material_scatter_u(self.flat, self.root, f32(ro[k]), f32(rd[k]), f32(p[k]), f32(n[k]), front[k], list(uniforms[k]))
             for k in range(cnt)]
        return g, o


def assert_scatter_agrees(g, o, dir_tol=2e-5, min_agree=1.0):
    """kind / uniforms used / attenuation identical, scattered direction within FP32 resolution.  A record whose
    stochastic decision (reflectance > u, |p| < 1 rejection, fuzz below the surface) sits within FP32 rounding of
    its uniform may legitimately take the other branch: `min_agree` < 1 admits that many per batch."""
    n = len(o)
    agree = 0
    for k in range(n):
        kind, att, d, em, used, _refl = o[k]
        assert np.allclose(g["emitted"][k], em, rtol=1e-6, atol=0)
        if g["kind"][k] != kind or g["used"][k] != used:
            continue
        agree += 1
        if kind:
            assert np.allclose(g["attenuation"][k], att, rtol=1e-6, atol=0), (k, g["attenuation"][k], att)
        if kind == 1:
            scale = max(1.0, float(np.abs(d).max()))
            assert np.abs(g["dir"][k] - d).max() <= dir_tol * scale, (k, g["dir"][k], d)
    assert agree >= min_agree * n, f"only {agree} of {n} records took the oracle's branch"
    return agree


# ------------------------------------------------------------------------------------------
# tests/materials/metal.test.ts:12-74, :127-156
# ------------------------------------------------------------------------------------------
def test_metal_jest_vectors(gpu):
    albedo = [0.8, 0.6, 0.2]
    pr = Pair({"type": "metal", "color": albedo, "fuzz": 0.0})
    try:  # metal.test.ts:30-74: 45-degree ray on an up-facing surface, fuzz 0
        g, o = pr.both([-1, 1, 0], unit([1, -1, 0]), [0, 0, 0], [0, 1, 0], 1, [[0.5] * 4])
        assert g["kind"][0] == 1 and o[0][0] == 1
        assert np.allclose(g["attenuation"][0], np.float32(albedo))
        assert np.allclose(g["dir"][0], unit([1, 1, 0]), atol=1e-5)  # toBeCloseTo(..., 5)
        assert g["used"][0] == 0 == o[0][4]                             # no random number without fuzz
        assert_scatter_agrees(g, o)
    finally:
        pr.close()
    pr = Pair({"type": "metal", "color": albedo, "fuzz": 0.5})
    try:  # metal.test.ts:76-124: straight down, fuzz 0.5, 100 draws: always outward, (nearly) never the mirror direction
        rng = np.random.default_rng(1)
        u = u24(rng, (100, 15))
        g, o = pr.both([0, 1, 0], [0, -1, 0], [0, 0, 0], [0, 1, 0], 1, u)
        assert np.all(g["kind"] == 1) and np.all(g["dir"][:, 1] > 0)
        perfect = (np.abs(g["dir"][:, 0]) < 1e-3) & (np.abs(g["dir"][:, 1] - 1) < 1e-3) & (np.abs(g["dir"][:, 2]) < 1e-3)
        assert perfect.sum() < 10
        assert_scatter_agrees(g, o, min_agree=0.98)
    finally:
        pr.close()
    pr = Pair({"type": "metal", "color": albedo, "fuzz": 0.8})
    try:  # metal.test.ts:127-156: grazing ray, randomInUnitSphere forced to (0,-1,0) => absorbed (null)
        # (0,-1,0) is not strictly inside the unit sphere for the rejection loop; the nearest accepted draw is
        # y = -1 + 2^-23: uniforms (0.5, 2^-24, 0.5) -> (0, -1 + 2^-23, 0)
        g, o = pr.both([0, 0.1, 0], unit([1, -0.01, 0]), [0, 0, 0], [0, 1, 0], 1, [[0.5, U24, 0.5]])
        assert g["kind"][0] == 0 == o[0][0] and g["used"][0] == 3 == o[0][4]
    finally:
        pr.close()


def test_metal_fuzz_clamp_on_device(gpu):  # metal.test.ts:12-28 (metal.ts:20): fuzz 1.5 -> 1, -0.5 -> 0
    u = [[0.5, 0.5, 1 - U24]]  # randomInUnitSphere -> (0, 0, 1 - 2^-23)
    for fuzz, eff in ((1.5, 1.0), (-0.5, 0.0), (0.25, 0.25)):
        pr = Pair({"type": "metal", "color": [1, 1, 1], "fuzz": fuzz})
        try:
            g, o = pr.both([0, 1, 0], [0, -1, 0], [0, 0, 0], [0, 1, 0], 1, u)
            assert np.allclose(g["dir"][0], [0, 1, eff * (1 - 2 * U24)], atol=1e-6)
            assert_scatter_agrees(g, o)
        finally:
            pr.close()


# ------------------------------------------------------------------------------------------
# tests/materials/dielectric.test.ts:30-134
# ------------------------------------------------------------------------------------------
def test_dielectric_jest_vectors(gpu):
    pr = Pair({"type": "glass", "ior": 1.5})
    try:
        # :109-134 Schlick: r0 ~ 0.04 at normal incidence, > 0.5 at cos = 0.1.  The device has no separate
        # reflectance function; the decision `reflectance > Math.random()` (dielectric.ts:69) exposes it.
        down = [0, -1, 0]
        g, o = pr.both([0, 1, 0], down, [0, 0, 0], [0, 1, 0], 1, [[0.039], [0.041]])
        assert list(g["kind"]) == [1, 1] and np.allclose(g["attenuation"], 1.0)       # :30-60 white attenuation, always a ray
        assert np.allclose(g["dir"][0], [0, 1, 0], atol=1e-6)                          # u < r0: reflected
        assert np.allclose(g["dir"][1], [0, -1, 0], atol=1e-6)                         # u > r0: refracted straight through
        assert_scatter_agrees(g, o)
        graze = unit([math.sqrt(1 - 0.01), -0.1, 0])                                   # cos(theta) = 0.1
        g, o = pr.both([0, 1, 0], graze, [0, 0, 0], [0, 1, 0], 1, [[0.5]])
        assert g["dir"][0][1] > 0                                                      # reflectance(0.1) > 0.5: reflected
        assert_scatter_agrees(g, o)
        # :80-107 total internal reflection: from inside (frontFace false, ratio = ior) at a grazing angle
        g, o = pr.both([0, 1, 0], unit([1, -0.1, 0]), [0, 0, 0], [0, 1, 0], 0, u24(np.random.default_rng(2), (20, 2)))
        assert np.all(g["dir"][:, 1] > 0) and np.all(g["used"] == 0)                    # cannot refract: no random number drawn
        assert_scatter_agrees(g, o)
    finally:
        pr.close()


# ------------------------------------------------------------------------------------------
# tests/materials/layeredMaterial.test.ts:57-123, :188-200
# ------------------------------------------------------------------------------------------
def test_layered_jest_vectors(gpu):
    glass = {"type": "glass", "ior": 1.5}
    red = {"type": "lambert", "color": [0.8, 0.2, 0.2]}
    silver = {"type": "metal", "color": [0.9, 0.9, 0.9], "fuzz": 0.1}
    rng = np.random.default_rng(3)
    pr = Pair({"type": "layered", "outer": glass, "inner": red})
    try:  # :57-99 ray (1,0,0) onto normal (-1,0,0): reflected -> white + ray; transmitted -> inner Lambertian's pdf
        g, o = pr.both([0, 0, 0], [1, 0, 0], [1, 0, 0], [-1, 0, 0], 1, u24(rng, (100, 8)))
        refl, inner = g["kind"] == 1, g["kind"] == 2
        assert refl.any() and inner.any() and not (g["kind"] == 0).any()
        assert np.allclose(g["attenuation"][refl], 1.0) and np.allclose(g["attenuation"][inner], np.float32([0.8, 0.2, 0.2]))
        assert_scatter_agrees(g, o)
    finally:
        pr.close()
    pr = Pair({"type": "layered", "outer": glass, "inner": silver})
    try:  # :101-123 ray (1,-1,0) onto normal (0,1,0): transmission reaches the metal (silver attenuation + ray)
        g, o = pr.both([0, 0, 0], [1, -1, 0], [1, 0, 0], [0, 1, 0], 1, u24(rng, (100, 12)))
        metal = np.all(np.isclose(g["attenuation"], np.float32(0.9)), axis=1) & (g["kind"] == 1)
        assert metal.any()
        assert_scatter_agrees(g, o, min_agree=0.98)
    finally:
        pr.close()
    pr = Pair({"type": "layered", "outer": glass, "inner": {"type": "light", "emit": [2, 3, 4]}})
    try:  # :188-200 emission comes from the inner material
        g, o = pr.both([0, 0, 0], [0, -1, 0], [0, 0, 0], [0, 1, 0], 1, [[0.9]])
        assert list(g["emitted"][0]) == [2, 3, 4] and g["kind"][0] == 0  # refracted into a light: scatter null
        assert_scatter_agrees(g, o)
    finally:
        pr.close()


# ------------------------------------------------------------------------------------------
# tests/materials/mixedMaterial.test.ts:49-128, :214-240
# ------------------------------------------------------------------------------------------
def test_mixed_jest_vectors(gpu):
    red = {"type": "lambert", "color": [0.8, 0.2, 0.2]}
    silver = {"type": "metal", "color": [0.9, 0.9, 0.9], "fuzz": 0.1}
    rng = np.random.default_rng(4)
    for weight, want in ((1.0, 2), (0.0, 1)):  # :63-97 weight 1 -> always material1 (pdf), weight 0 -> always material2 (ray)
        pr = Pair({"type": "mixed", "diff": red, "spec": silver, "weight": weight})
        try:
            g, o = pr.both([0, 0, 0], [1, -1, 0], [1, 0, 0], [0, 1, 0], 1, u24(rng, (10, 12)))
            assert np.all(g["kind"] == want)
            assert_scatter_agrees(g, o)
        finally:
            pr.close()
    pr = Pair({"type": "mixed", "diff": red, "spec": silver, "weight": 0.3})
    try:  # :99-128 30 / 70 split within 5 %
        g, o = pr.both([0, 0, 0], [1, -1, 0], [1, 0, 0], [0, 1, 0], 1, u24(rng, (1000, 12)))
        assert abs(float(np.mean(g["kind"] == 2)) - 0.3) < 0.05 and abs(float(np.mean(g["kind"] == 1)) - 0.7) < 0.05
        assert_scatter_agrees(g, o, min_agree=0.995)
    finally:
        pr.close()
    pr = Pair({"type": "mixed", "diff": {"type": "light", "emit": [1, 2, 3]}, "spec": {"type": "light", "emit": [4, 5, 6]}, "weight": 0.3})
    try:  # :214-240 emitted = w e1 + (1 - w) e2
        g, o = pr.both([0, 0, 0], [0, -1, 0], [0, 0, 0], [0, 1, 0], 1, [[0.5]])
        assert np.allclose(g["emitted"][0], [0.3 * 1 + 0.7 * 4, 0.3 * 2 + 0.7 * 5, 0.3 * 3 + 0.7 * 6], atol=1e-5)
        assert_scatter_agrees(g, o)
    finally:
        pr.close()


# ------------------------------------------------------------------------------------------
# every material kind on random hits: the rare branches
# ------------------------------------------------------------------------------------------
MATERIALS = {
    "lambert": {"type": "lambert", "color": [0.1, 0.2, 0.3]},
    "light": {"type": "light", "emit": [15, 14, 13]},
    "metal_fuzz": {"type": "metal", "color": [0.7, 0.6, 0.5], "fuzz": 0.9},
    "glass": {"type": "glass", "ior": 1.5},
    "glass_low_ior": {"type": "glass", "ior": 0.8},
    "layered_over_metal": {"type": "layered", "outer": {"type": "glass", "ior": 1.5}, "inner": {"type": "metal", "color": [0.8, 0.8, 0.8], "fuzz": 0.3}},
    "layered_over_mixed": {"type": "layered", "outer": {"type": "glass", "ior": 1.3},
                           "inner": {"type": "mixed", "diff": {"type": "lambert", "color": [0.6, 0.4, 0.2]},
                                     "spec": {"type": "metal", "color": [0.8, 0.6, 0.4], "fuzz": 0.1}, "weight": 0.7}},
    "mixed_glass_lambert": {"type": "mixed", "diff": {"type": "glass", "ior": 1.5}, "spec": {"type": "lambert", "color": [0.3, 0.3, 0.9]}, "weight": 0.4},
    "mixed_of_mixed": {"type": "mixed", "weight": 0.5,
                       "diff": {"type": "mixed", "diff": {"type": "lambert", "color": [0.9, 0.1, 0.1]}, "spec": {"type": "metal", "color": [1, 1, 1], "fuzz": 0}, "weight": 0.5},
                       "spec": {"type": "layered", "outer": {"type": "glass", "ior": 1.5}, "inner": {"type": "lambert", "color": [0.1, 0.9, 0.1]}}},
}


@pytest.mark.parametrize("name", list(MATERIALS))
def test_scatter_random_hits_match_oracle(gpu, name):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    n = 1500
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.float32(nrm).astype(np.float64)
    # incoming directions against the (face-forwarded) normal, from head-on to grazing, not normalised
    tang = np.cross(nrm, rng.normal(size=(n, 3)))
    tang /= np.linalg.norm(tang, axis=1, keepdims=True)
    cos_in = rng.uniform(0.0, 1.0, size=(n, 1)) ** 2
    rd = (-cos_in * nrm + np.sqrt(1 - cos_in**2) * tang) * rng.uniform(0.2, 5.0, size=(n, 1))
    front = rng.integers(0, 2, size=n)
    p = rng.uniform(-2, 2, size=(n, 3))
    u = u24(rng, (n, 16))
    pr = Pair(MATERIALS[name])
    try:
        g, o = pr.both(p - rd, rd, p, nrm, front, u)
        agree = assert_scatter_agrees(g, o, dir_tol=1e-4, min_agree=0.99)
        kinds = {int(k) for k in g["kind"]}
        if name in ("metal_fuzz", "layered_over_metal"):
            assert 0 in kinds and 1 in kinds  # absorption below the surface really happens
        if name == "glass":
            tir = sum(1 for k in range(n) if o[k][4] == 0)  # no uniform drawn = total internal reflection
            assert tir > 20
        assert agree > 0
    finally:
        pr.close()


# ------------------------------------------------------------------------------------------
# light pdfs: tests/entities/quad.test.ts:205-261, tests/entities/sphere.test.ts:100-137
# ------------------------------------------------------------------------------------------
def light_scene(light_obj):
    return {"camera": {"vfov": 40, "from": [0, 0, -5], "at": [0, 0, 0]},
            "materials": [{"id": "l", "material": {"type": "light", "emit": [1, 1, 1]}}, {"id": "m", "material": {"type": "lambert", "color": [0.5, 0.5, 0.5]}}],
            "objects": [{**light_obj, "material": "l", "light": True}, {"type": "sphere", "pos": [0, -100, 0], "r": 1, "material": "m"}]}


def test_quad_light_pdf_jest_vectors(gpu):
    sd = light_scene({"type": "quad", "pos": [0, 0, 5], "u": [1, 0, 0], "v": [0, 1, 0]})
    oc = ob.OracleCamera(sd, {"width": 16, "samples": 1})
    with createCameraFromSceneData(sd, {"width": 16, "samples": 1}) as cam:
        assert cam.nLights == 1 == oc.n_lights
        d_hit = unit([0.5, 0.5, 5])
        v = cam.debugLightPdf(0, [[0, 0, 0], [0, 0, 0]], [d_hit, [0, 1, 0]])
        expected = (0.5**2 + 0.5**2 + 25) / (1.0 * abs(d_hit[2]))  # quad.test.ts:240-261: distance^2 / (area * cosine)
        assert v[0] == pytest.approx(expected, abs=1e-4) and v[1] == 0   # :207-222 a direction that misses -> 0
        assert v[0] == pytest.approx(ob.light_pdf_value(oc, 0, [0, 0, 0], f32(d_hit)), rel=1e-5)
        # quad.test.ts:264-300: random vectors are unit and point at the quad
        rng = np.random.default_rng(5)
        u = u24(rng, (200, 2))
        org = rng.uniform(-3, 3, size=(200, 3)) * [1, 1, 0.5]
        dirs = cam.debugLightRandomVec(0, org, u)
        assert np.allclose(np.linalg.norm(dirs, axis=1), 1, atol=1e-5)
        vals = cam.debugLightPdf(0, org, dirs)
        for k in range(200):
            od, used = ob.light_random_vec_u(oc, 0, f32(org[k]), list(u[k]))
            assert used == 2 and np.abs(dirs[k] - od).max() <= 2e-6
            ov = ob.light_pdf_value(oc, 0, f32(org[k]), dirs[k].astype(np.float64))
            # a sampled point on the quad's very edge may fall either side of the inclusive test in FP32
            edge = min(u[k][0], 1 - u[k][0], u[k][1], 1 - u[k][1]) < 1e-5
            assert vals[k] == pytest.approx(ov, rel=2e-4) or edge


def test_sphere_light_pdf_jest_vectors(gpu):
    sd = light_scene({"type": "sphere", "pos": [0, 0, -1], "r": 0.5})
    oc = ob.OracleCamera(sd, {"width": 16, "samples": 1})
    with createCameraFromSceneData(sd, {"width": 16, "samples": 1}) as cam:
        v = cam.debugLightPdf(0, [[0, 0, 0], [0, 0, 0]], [[0, 0, -1], [0, 1, 0]])
        assert v[0] == pytest.approx(1 / (2 * math.pi * (1 - math.sqrt(0.75))), abs=1e-5)  # sphere.test.ts:100-119
        assert v[1] == 0                                                                     # :121-137 miss -> 0
        rng = np.random.default_rng(6)
        u = u24(rng, (300, 2))
        org = rng.uniform(-3, 3, size=(300, 3))
        org[:20] = [0, 0, -1] + rng.uniform(-0.2, 0.2, size=(20, 3))   # origins INSIDE the light: 1/(4 pi) and NaN directions
        dirs = cam.debugLightRandomVec(0, org, u)
        vals = cam.debugLightPdf(0, org, np.nan_to_num(dirs.astype(np.float64), nan=0.3))
        for k in range(300):
            od, used = ob.light_random_vec_u(oc, 0, f32(org[k]), list(u[k]))
            assert used == 2
            if np.any(np.isnan(od)):
                assert np.any(np.isnan(dirs[k]))   # sphere.ts:140-147 from inside: sqrt of a negative number on both sides
                continue
            assert np.abs(dirs[k] - od).max() <= 5e-5, (k, dirs[k], od)
            ov = ob.light_pdf_value(oc, 0, f32(org[k]), np.nan_to_num(dirs[k].astype(np.float64), nan=0.3))
            assert vals[k] == pytest.approx(ov, rel=5e-4, abs=1e-6), (k, vals[k], ov)
        # sphere.test.ts:189-211: sampled directions stay inside the cone subtended by the sphere
        o = np.array([0.0, 0.0, 2.0])
        d = cam.debugLightRandomVec(0, np.tile(o, (200, 1)), u[:200])
        cos_max = math.sqrt(1 - 0.25 / 9.0)
        assert np.all(d @ unit([0, 0, -3]) >= cos_max - 1e-5)


# ------------------------------------------------------------------------------------------
# the diffuse branch of rayColor: mixture pdf with 0, 1 (quad) and 2 (quad + sphere) lights
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scene", ["weekend-no-light", "cornell-quad", "default-quad+sphere"])
def test_diffuse_bounce_matches_oracle(gpu, scene):
    sd = {"weekend-no-light": generateWeekendFinalSceneData, "cornell-quad": generateCornellSceneData, "default-quad+sphere": generateDefaultSceneData}[scene]()
    opts = {"width": 32, "samples": 4}
    oc = ob.OracleCamera(sd, opts)
    rng = np.random.default_rng(7)
    n = 2000
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[:6] = [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]  # both ONB branches (|w.x| > 0.9)
    nrm = np.float32(nrm).astype(np.float64)
    p = np.float32(rng.uniform(-0.9, 0.9, size=(n, 3))).astype(np.float64)
    u = u24(rng, (n, 3))
    with createCameraFromSceneData(sd, opts) as cam:
        assert cam.nLights == oc.n_lights == {"weekend-no-light": 0, "cornell-quad": 1, "default-quad+sphere": 2}[scene]
        g = cam.debugDiffuseBounce(p, nrm, u)
    picked_light = same_decision = 0
    for k in range(n):
        o, used = ob.diffuse_bounce_u(oc, p[k], nrm[k], list(u[k]))
        assert used == 3
        same_decision += g[k][5] == o[5]
        if np.any(np.isnan(o[:3])):  # sphere light sampled from inside it (sphere.ts:140-147): NaN on both sides, path ends
            assert np.any(np.isnan(g[k][:3])) and g[k][5] == 0 == o[5]
            continue
        assert np.abs(g[k][:3] - o[:3]).max() <= 5e-5, (k, g[k], o)
        # the light pdf of a direction that grazes the light's edge may differ (inclusive edge in FP32 vs FP64)
        if g[k][5] == o[5]:
            assert g[k][3] == pytest.approx(o[3], rel=1e-3, abs=2e-6), (k, g[k], o)
        assert g[k][4] == pytest.approx(o[4], rel=1e-4, abs=2e-6), (k, g[k], o)
        picked_light += u[k][0] * (0.5 + 0.5 * (oc.n_lights > 0)) >= 0.5
    assert same_decision >= 0.995 * n
    if oc.n_lights:
        assert picked_light > n // 3   # the light-sampling half of the mixture was exercised


# ------------------------------------------------------------------------------------------
# Camera.getRay: tests/camera.test.ts:202-222 (no jitter), :225-330 (defocus), :537-587
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,samples", [("cornell", 1), ("cornell", 64), ("weekend-aperture", 64), ("weekend-aperture", 1)])
def test_get_ray_matches_oracle_bit_for_bit(gpu, name, samples):
    sd = generateCornellSceneData() if name == "cornell" else generateWeekendFinalSceneData()
    opts = {"width": 200, "samples": samples}
    oc = ob.OracleCamera(sd, opts)
    rng = np.random.default_rng(8)
    n = 500
    with createCameraFromSceneData(sd, opts) as cam:
        ij = np.stack([rng.integers(0, cam.imageWidth, n), rng.integers(0, cam.imageHeight, n)], axis=1).astype(np.int32)
        u = u24(rng, (n, 16))
        org, dirs, used = cam.debugGetRay(ij, u)
        aperture = cam.aperture
    for k in range(n):
        oo, od, oused = ob.get_ray_u(oc, ij[k][0], ij[k][1], list(u[k]))
        assert used[k] == oused
        # camera_ray keeps the reference's FP32 rounding sequence (scale, then add, each rounded): bit-identical
        assert np.array_equal(org[k], oo) and np.array_equal(dirs[k], od), (k, org[k], oo, dirs[k], od)
    if samples == 1 and aperture == 0:
        assert np.all(used == 0)          # camera.test.ts:202-222: identical rays when samples = 1 and aperture = 0
    if samples > 1:
        assert np.all(used >= 2)          # pixel jitter draws two numbers (camera.ts:184-191)
    if aperture > 0:
        assert np.all(used >= (2 if samples > 1 else 0) + 2) and np.any(used > (2 if samples > 1 else 0) + 2)  # disk rejection loop
        assert np.any(org != org[0])      # camera.test.ts:225-330: the origin moves on the defocus disk
