"""Pins the CPU oracle against the known-answer vectors held by the reference's own Jest
tests.  Every test cites the reference test it transcribes
(`/root/reference/tests/...`).  Runs on CPU (`-m "not gpu"`)."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_binding as ob
from oracle_binding import d3, lib
from mcp_raytracer_b200 import scenes
from mcp_raytracer_b200.scene_data import RT_OBJ_PLANE, RT_OBJ_QUAD, RT_OBJ_SPHERE, FlatScene

INF = float("inf")


def sphere_hit(c, r, o, d, tmin, tmax):
    out = (C.c_double * 8)()
    h = lib().orc_sphere_hit(d3(c), r, d3(o), d3(d), tmin, tmax, out)
    return list(out) if h else None


def planar_hit(kind, q, u, v, o, d, tmin=0.0, tmax=INF):
    out = (C.c_double * 8)()
    h = lib().orc_planar_hit(kind, d3(q), d3(u), d3(v), d3(o), d3(d), tmin, tmax, out)
    return list(out) if h else None


def unit(v):
    out = (C.c_float * 3)()
    lib().orc_unit(d3(v), out)
    return list(out)


# ---------------------------------------------------------------- Philox (Random123 KAT)
def test_philox4x32_known_answers():
    """Random123's kat_vectors for philox4x32 at 10 rounds (its default) and at 7 (the fewest that are crush-resistant,
    Salmon et al., SC'11 table 2)."""
    def ph(ctr, key, rounds):
        out = (C.c_uint32 * 4)()
        lib().orc_philox4x32((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), rounds, out)
        return list(out)

    pi = ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])
    assert ph([0] * 4, [0] * 2, 10) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert ph([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, 10) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert ph(*pi, 10) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    assert ph([0] * 4, [0] * 2, 7) == [0x5F6FB709, 0x0D893F64, 0x4F121F81, 0x4F730A48]
    assert ph([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, 7) == [0x5207DDC2, 0x45165E59, 0x4D8EE751, 0x8C52F662]
    assert ph(*pi, 7) == [0x4DFCCABA, 0x190A87F0, 0xC47362BA, 0xB6B5242A]
    assert lib().orc_philox_rounds() in (7, 10)


# ---------------------------------------------------------------- tests/entities/sphere.test.ts
def test_sphere_hit_vectors():  # sphere.test.ts:16-90
    c, r = (0, 0, -1), 0.5
    h = sphere_hit(c, r, (0, 0, 0), (0, 0, -1), 0, INF)
    assert h and h[0] == pytest.approx(0.5) and h[3] == pytest.approx(-0.5) and h[7] == 1.0
    assert (h[4], h[5], h[6]) == pytest.approx((0, 0, 1))
    assert sphere_hit(c, r, (0, 1, 0), (0, 0, -1), 0, INF) is None
    h = sphere_hit(c, r, (0, 0, -1), (0, 0, -1), 0.001, INF)  # from inside
    assert h and h[0] == pytest.approx(0.5) and h[3] == pytest.approx(-1.5) and h[7] == 0.0
    assert (h[4], h[5], h[6]) == pytest.approx((0, 0, 1))  # flipped inward
    h = sphere_hit(c, r, (0, 0.5, 0), (0, 0, -1), 0, INF)  # tangent
    assert h and h[0] == pytest.approx(1.0) and h[2] == pytest.approx(0.5) and h[3] == pytest.approx(-1.0) and h[7] == 1.0
    assert sphere_hit(c, r, (0, 0, 0), (0, 0, -1), 0.6, 1.0) is None
    assert sphere_hit(c, r, (0, 0, 0), (0, 0, -1), 0.0, 0.4) is None


def test_sphere_pdf_value():  # sphere.test.ts:100-137
    c, r = (0, 0, -1), 0.5
    assert lib().orc_sphere_pdf_value(d3(c), r, d3((0, 0, 0)), d3(unit((0, 1, 0)))) == 0
    v = lib().orc_sphere_pdf_value(d3(c), r, d3((0, 0, 0)), d3(unit((0, 0, -1))))
    assert v == pytest.approx(1 / (2 * math.pi * (1 - math.sqrt(0.75))), abs=1e-5)
    far = lib().orc_sphere_pdf_value(d3(c), r, d3((0, 0, 4)), d3(unit((0, 0, -1))))
    assert far > v  # smaller solid angle, higher pdf (sphere.test.ts:139-160)


def test_sphere_pdf_samples_stay_in_cone():  # sphere.test.ts:189-211
    c, r, o = (0, 0, -1), 0.5, (0, 0, 0)
    n = 200
    out = (C.c_float * (3 * n))()
    lib().orc_sphere_pdf_random(d3(c), r, d3(o), 42, n, out)
    v = np.array(out, np.float64).reshape(n, 3)
    assert np.allclose(np.linalg.norm(v, axis=1), 1, atol=1e-4)
    cos_max = math.sqrt(1 - r * r / 1.0)
    assert np.all(v @ np.array([0, 0, -1.0]) >= cos_max - 1e-4)
    for k in range(n):  # every sampled direction hits the sphere => pdf > 0
        assert lib().orc_sphere_pdf_value(d3(c), r, d3(o), d3(v[k])) > 0 or (v[k] @ np.array([0, 0, -1.0])) < cos_max + 1e-4


# ---------------------------------------------------------------- tests/entities/quad.test.ts
Q, U, V = (0, 0, 5), (1, 0, 0), (0, 1, 0)


def test_quad_hits():  # quad.test.ts:50-118, :136-161, :356-370
    h = planar_hit(RT_OBJ_QUAD, Q, U, V, (0.5, 0.5, 0), (0, 0, 1))
    assert h and h[0] == pytest.approx(5) and (h[1], h[2], h[3]) == pytest.approx((0.5, 0.5, 5))
    assert planar_hit(RT_OBJ_QUAD, Q, U, V, (1.5, 0.5, 0), (0, 0, 1)) is None
    for corner in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0)]:  # inclusive edges
        h = planar_hit(RT_OBJ_QUAD, Q, U, V, corner, (0, 0, 1))
        assert h and h[0] == pytest.approx(5) and h[1] == pytest.approx(corner[0]) and h[2] == pytest.approx(corner[1])
    assert planar_hit(RT_OBJ_QUAD, Q, U, V, (0.5, 0.5, 0), (1, 0, 0)) is None  # parallel
    assert planar_hit(RT_OBJ_QUAD, Q, U, V, (1.0001, 0.5, 0), (0, 0, 1)) is None
    h = planar_hit(RT_OBJ_QUAD, Q, U, V, (0.5, 0.5, 0), (0, 0, 1))  # from -z side
    assert h[7] == 0.0 and h[6] == pytest.approx(-1)
    h = planar_hit(RT_OBJ_QUAD, Q, U, V, (0.5, 0.5, 10), (0, 0, -1))  # from +z side
    assert h[7] == 1.0 and h[6] == pytest.approx(1)


def test_quad_bbox_padding():  # quad.test.ts:164-203
    out = (C.c_double * 6)()
    lib().orc_object_bbox(RT_OBJ_QUAD, d3(Q), d3(U), d3(V), 0.0, out)
    b = list(out)
    assert b[0] == pytest.approx(-1e-4) and b[1] == pytest.approx(-1e-4) and b[2] == pytest.approx(5 - 1e-4)
    assert b[3] == pytest.approx(1 + 1e-4) and b[4] == pytest.approx(1 + 1e-4) and b[5] == pytest.approx(5 + 1e-4)


def test_quad_pdf_value():  # quad.test.ts:225-261
    d = unit((0.5, 0.5, 5))
    v = lib().orc_quad_pdf_value(d3(Q), d3(U), d3(V), d3((0, 0, 0)), d3(d))
    dist2 = 0.5**2 + 0.5**2 + 25
    cosine = abs(d[2])
    assert v == pytest.approx(dist2 / (1.0 * cosine), abs=1e-4)
    assert lib().orc_quad_pdf_value(d3(Q), d3(U), d3(V), d3((0, 0, 0)), d3(unit((5, 5, 1)))) == 0


def test_quad_pdf_random_points_at_quad():  # quad.test.ts:264-300
    n = 100
    out = (C.c_float * (3 * n))()
    lib().orc_quad_pdf_random(d3(Q), d3(U), d3(V), d3((0, 0, 0)), 7, n, out)
    v = np.array(out, np.float64).reshape(n, 3)
    assert np.allclose(np.linalg.norm(v, axis=1), 1, atol=1e-4)
    for k in range(n):
        assert lib().orc_quad_pdf_value(d3(Q), d3(U), d3(V), d3((0, 0, 0)), d3(v[k])) > 0


# ---------------------------------------------------------------- tests/entities/plane.test.ts
def plane_intersect(q, u, v, o, d, tmin=0.0, tmax=INF):
    out = (C.c_double * 11)()
    lib().orc_plane_intersect(d3(q), d3(u), d3(v), d3(o), d3(d), tmin, tmax, out)
    return list(out)


def test_plane_ctor_and_intersect():  # plane.test.ts:17-136
    r = plane_intersect(Q, U, V, (0, 0, 0), (0, 0, 1))
    assert r[0:3] == pytest.approx([0, 0, 1]) and r[3] == pytest.approx(5) and r[4:7] == pytest.approx([0, 0, 1])
    assert r[7] == 1.0 and r[8] == pytest.approx(5) and r[9] == pytest.approx(0) and r[10] == pytest.approx(0)
    # tilted plane: normal = unit(u x v), d = n.q
    r = plane_intersect((1, 1, 1), (1, 1, 0), (0, 1, 1), (0, 0, 0), (0, 0, 1))
    n = np.cross([1, 1, 0], [0, 1, 1]) / np.linalg.norm(np.cross([1, 1, 0], [0, 1, 1]))
    assert r[0:3] == pytest.approx(list(n), abs=1e-6) and r[3] == pytest.approx(float(n @ [1, 1, 1]), abs=1e-6)
    assert plane_intersect(Q, U, V, (0, 0, 0), (1, 1, 0))[7] == 0.0  # parallel
    assert plane_intersect(Q, U, V, (0, 0, 0), (0, 0, 1), 0, 4)[7] == 0.0  # interval ends before plane
    r = plane_intersect((0, 0, 0), (2, 0, 0), (0, 3, 0), (1, 1.5, -1), (0, 0, 1))
    assert r[7] == 1.0 and r[8] == pytest.approx(1) and r[9] == pytest.approx(0.5) and r[10] == pytest.approx(0.5)


def test_plane_hit_faces():  # plane.test.ts:140-192
    h = planar_hit(RT_OBJ_PLANE, Q, U, V, (0, 0, 0), (0, 0, 1))
    assert h and h[0] == pytest.approx(5) and h[6] == pytest.approx(-1) and h[7] == 0.0
    h = planar_hit(RT_OBJ_PLANE, Q, U, V, (0, 0, 10), (0, 0, -1))
    assert h and h[0] == pytest.approx(5) and h[6] == pytest.approx(1) and h[7] == 1.0
    assert planar_hit(RT_OBJ_PLANE, Q, U, V, (0, 0, 0), (1, 0, 0)) is None
    assert planar_hit(RT_OBJ_PLANE, Q, U, V, (0, 0, 0), unit((0, 1e-6, 1))) is not None  # plane.test.ts:303-317
    assert planar_hit(RT_OBJ_PLANE, Q, U, V, (0, 0, 10), (0, 0, 1)) is None  # negative t


def test_plane_bboxes():  # plane.test.ts:212-286
    out = (C.c_double * 6)()
    lib().orc_object_bbox(RT_OBJ_PLANE, d3((0, 0, 5)), d3((1, 0, 0)), d3((0, 1, 0)), 0.0, out)
    assert out[0] == -INF and out[1] == -INF and out[2] == pytest.approx(5 - 1e-4) and out[3] == INF and out[5] == pytest.approx(5 + 1e-4)
    lib().orc_object_bbox(RT_OBJ_PLANE, d3((0, 3, 0)), d3((1, 0, 0)), d3((0, 0, 1)), 0.0, out)
    assert out[0] == -INF and out[1] == pytest.approx(3 - 1e-4) and out[2] == -INF and out[4] == pytest.approx(3 + 1e-4)
    lib().orc_object_bbox(RT_OBJ_PLANE, d3((-2, 0, 0)), d3((0, 1, 0)), d3((0, 0, 1)), 0.0, out)
    assert out[0] == pytest.approx(-2 - 1e-4) and out[1] == -INF and out[3] == pytest.approx(-2 + 1e-4) and out[4] == INF
    lib().orc_object_bbox(RT_OBJ_PLANE, d3((0, 0, 0)), d3((1, 1, 0)), d3((0, 1, 1)), 0.0, out)
    assert list(out) == [-INF, -INF, -INF, INF, INF, INF]


def test_negative_radius_sphere_box_is_inverted():  # SURVEY.md App. A.4 (sphere.ts:25-30)
    out = (C.c_double * 6)()
    lib().orc_object_bbox(RT_OBJ_SPHERE, d3((-0.5, 0.25, -0.5)), d3((0, 0, 0)), d3((0, 0, 0)), -0.24, out)
    assert out[0] > out[3] and out[1] > out[4] and out[2] > out[5]


# ---------------------------------------------------------------- tests/geometry/aabb.test.ts
def test_aabb_vectors():  # aabb.test.ts:8-110
    mn, mx = d3((-1, -1, -1)), d3((1, 1, 1))
    assert lib().orc_aabb_hit(mn, mx, d3((0, 0, -5)), d3((0, 0, 1)), 0.1, 100) == 1
    assert lib().orc_aabb_hit(mn, mx, d3((5, 0, 0)), d3((0, 0, 1)), 0.1, 100) == 0
    assert lib().orc_aabb_hit(mn, mx, d3((0, 0, -5)), d3((0, 0, 1)), 0.1, 3) == 0
    out = (C.c_double * 6)()
    lib().orc_surrounding_box(mn, mx, d3((0, 0, 0)), d3((2, 2, 2)), out)
    assert list(out) == [-1, -1, -1, 2, 2, 2]
    lib().orc_surrounding_box(d3((INF,) * 3), d3((-INF,) * 3), mn, mx, out)  # empty is the identity
    assert list(out) == [-1, -1, -1, 1, 1, 1]


# ---------------------------------------------------------------- tests/geometry/bvh.test.ts, hittableList.test.ts
def _sphere_scene(spheres):
    return {
        "camera": {"vfov": 90, "from": [0, 0, 0], "at": [0, 0, -1], "up": [0, 1, 0], "aperture": 0, "focus": 1,
                   "background": {"type": "gradient", "top": [1, 1, 1], "bottom": [0.5, 0.7, 1.0]}},
        "materials": [{"id": "m", "material": {"type": "lambert", "color": [0.5, 0.5, 0.5]}}],
        "objects": [{"type": "sphere", "pos": list(c), "r": r, "material": "m"} for c, r in spheres],
    }


def test_bvh_equals_linear_list():  # bvh.test.ts:52-159
    four = [((0, 0, -1), 0.5), ((-1, 0, -1), 0.5), ((1, 0, -1), 0.5), ((0, -100.5, -1), 100)]
    cam = ob.OracleCamera(_sphere_scene(four), {"width": 8, "samples": 1})
    ids, t, _, _ = cam.trace_rays([[0, 0, 0]], [[0, 0, -1]], 0.1, 100)
    assert ids[0] == 0 and t[0] == pytest.approx(0.5)
    ids, t, _, _ = cam.trace_rays([[0, 5, 0]], [[0, 1, 0]], 0.1, 100)
    assert ids[0] == -1
    d = unit((0.5, -0.5, -1))
    a = cam.trace_rays([[0, 0, 0]], [d], 0.1, 100, use_bvh=True)
    b = cam.trace_rays([[0, 0, 0]], [d], 0.1, 100, use_bvh=False)
    assert a[0][0] == b[0][0] and a[1][0] == pytest.approx(b[1][0])
    ten = [((i - 5, 0, -5), 0.3) for i in range(10)]
    cam = ob.OracleCamera(_sphere_scene(ten), {"width": 8, "samples": 1})
    a = cam.trace_rays([[0, 0, 0]], [[0, 0, -1]], 0.1, 100, use_bvh=True)
    b = cam.trace_rays([[0, 0, 0]], [[0, 0, -1]], 0.1, 100, use_bvh=False)
    assert a[0][0] == b[0][0] == 5 and a[1][0] == pytest.approx(b[1][0]) == pytest.approx(4.7)


def test_hittable_list_closest_hit():  # hittableList.test.ts:55-123
    three = [((0, 0, -1), 0.5), ((0, 0, -2), 0.5), ((0, 0, -3), 0.5)]
    cam = ob.OracleCamera(_sphere_scene(three), {"width": 8, "samples": 1})
    for use_bvh in (True, False):
        ids, t, _, _ = cam.trace_rays([[0, 0, 0]], [[0, 0, -1]], 0, INF, use_bvh=use_bvh)
        assert ids[0] == 0 and t[0] == pytest.approx(0.5)
        ids, t, _, _ = cam.trace_rays([[0, 0, 0]], [[0, 0, -1]], 1.0, INF, use_bvh=use_bvh)
        assert t[0] == pytest.approx(1.5)
        ids, t, _, _ = cam.trace_rays([[0, 0, 0]], [[0, 0, -1]], 2.0, INF, use_bvh=use_bvh)
        assert t[0] == pytest.approx(2.5)
        ids, _, _, _ = cam.trace_rays([[0, 0, 0]], [[0, 0, -1]], 0.6, 0.9, use_bvh=use_bvh)
        assert ids[0] == -1


def test_bvh_matches_brute_force_on_random_rays():
    """Not a reference vector: cross-check of the BVH restatement against the no-BVH list on
    every generator scene (SURVEY.md §8c 'cross-check it with an independent brute-force mode')."""
    rng = np.random.default_rng(3)
    for sd in (scenes.generateDefaultSceneData(), scenes.generateCornellSceneData(),
               scenes.generateSpheresSceneData({"count": 60, "seed": 5}), scenes.generateRainSceneData({"count": 300, "seed": 2})):
        cam = ob.OracleCamera(sd, {"width": 8, "samples": 1})
        n = 400
        o = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32) + np.array([0, 0.5, 0.5], np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        a = cam.trace_rays(o, d, use_bvh=True)
        b = cam.trace_rays(o, d, use_bvh=False)
        # the negative-radius sphere of the default scene is only reachable without the BVH
        # (inverted box): exclude rays whose list hit is that object
        neg = [i for i, ob_ in enumerate(sd["objects"]) if ob_.get("r", 1) < 0]
        keep = ~np.isin(b[0], neg)
        same_t = np.isclose(a[1][keep], b[1][keep], rtol=1e-12, atol=0) | (np.isinf(a[1][keep]) & np.isinf(b[1][keep]))
        assert same_t.all()


# ---------------------------------------------------------------- tests/geometry/pdf.test.ts, onbasis.test.ts
def test_cosine_pdf_values():  # pdf.test.ts:55-77
    n = (0, 1, 0)
    f = lib().orc_cosine_pdf_value
    assert f(d3(n), d3((0, 1, 0))) == pytest.approx(1 / math.pi)
    assert f(d3(n), d3(unit((1, 1, 0)))) == pytest.approx(0.7071 / math.pi, abs=1e-4)
    assert f(d3(n), d3((1, 0, 0))) == pytest.approx(0)
    assert f(d3(n), d3((0, -1, 0))) == 0


def test_cosine_generate_statistics():  # pdf.test.ts:28-52, :159-179
    cnt = 4000
    out = (C.c_float * (3 * cnt))()
    lib().orc_cosine_pdf_generate(d3((0, 1, 0)), 11, cnt, out)
    v = np.array(out, np.float64).reshape(cnt, 3)
    assert np.allclose(np.linalg.norm(v, axis=1), 1, atol=1e-3)
    assert np.all(v[:, 1] >= -1e-6)
    lib().orc_random_cosine_direction(12, cnt, out)
    z = np.array(out, np.float64).reshape(cnt, 3)[:, 2]
    assert z.mean() == pytest.approx(2 / 3, abs=0.02)


def test_mixture_pdf_value_and_selection():  # pdf.test.ts:80-156
    mv = lib().orc_mixture_value
    arr = lambda a: (C.c_double * len(a))(*a)  # noqa: E731
    assert mv(2, arr([1.0, 1.0]), arr([0.5, 0.5])) == pytest.approx(1.0)
    assert mv(2, arr([1.0, 3.0]), arr([0.5, 0.5])) == pytest.approx(2.0)
    assert mv(2, arr([1.0, 3.0]), arr([1.0, 3.0])) == pytest.approx(2.5)
    rng = np.random.default_rng(0)
    picks = [lib().orc_mixture_select(2, arr([1.0, 3.0]), float(u)) for u in rng.random(4000)]
    assert np.mean(np.array(picks) == 0) == pytest.approx(0.25, abs=0.03)  # 1:3 ratio (pdf.test.ts:94-124)


def test_onb_orthonormal():  # onbasis.test.ts
    for n in [(0, 1, 0), (1, 0, 0), (0.95, 0.1, 0.2), (1, 2, 3), (0, 0, -4)]:
        out = (C.c_float * 9)()
        lib().orc_onb(d3(n), out)
        b = np.array(out, np.float64).reshape(3, 3)
        assert np.allclose(b @ b.T, np.eye(3), atol=1e-5)
        assert np.allclose(b[2], np.array(n) / np.linalg.norm(n), atol=1e-6)


# ---------------------------------------------------------------- tests/materials/*.test.ts
def _mat_scene(material):
    return FlatScene({"camera": {}, "materials": [{"id": "m", "material": material}],
                      "objects": [{"type": "sphere", "pos": [0, 0, 0], "r": 1, "material": "m"}]})


def scatter(material, rd, n=(0, 1, 0), front=True, seed=1, sample=0, p=(0, 0, 0)):
    fs = _mat_scene(material)
    out = (C.c_double * 9)()
    em = (C.c_float * 3)()
    rc = lib().orc_material_scatter(C.byref(fs.desc), int(fs.obj_material[0]), d3((0, 1, 0)), d3(rd), d3(p), d3(n),
                                    1 if front else 0, seed, sample, out, em)
    assert rc >= 0
    return (list(out) if rc == 1 else None), list(em)


def test_metal_reflection_and_fuzz():  # metal.test.ts:12-74, :127-175
    r, _ = scatter({"type": "metal", "color": [0.8, 0.8, 0.8], "fuzz": 0.0}, (1, -1, 0))
    assert r[0] == 1.0 and r[2:5] == pytest.approx([0.8, 0.8, 0.8])
    assert r[5:8] == pytest.approx([1 / math.sqrt(2), 1 / math.sqrt(2), 0], abs=1e-5)
    assert lib().orc_metal_fuzz_clamp(-0.5) == 0 and lib().orc_metal_fuzz_clamp(1.5) == 1 and lib().orc_metal_fuzz_clamp(0.3) == 0.3
    out = (C.c_float * 3)()
    lib().orc_reflect(d3((1, -1, 0)), d3((0, 1, 0)), out)
    assert list(out) == pytest.approx([1, 1, 0])
    # grazing ray + full fuzz is absorbed sometimes (reflection pushed below the surface)
    res = [scatter({"type": "metal", "color": [1, 1, 1], "fuzz": 1.0}, (1, -0.01, 0), seed=s)[0] for s in range(200)]
    assert any(x is None for x in res) and any(x is not None for x in res)


def test_dielectric_vectors():  # dielectric.test.ts:30-134
    r, _ = scatter({"type": "glass", "ior": 1.5}, (0, -1, 0))
    assert r[0] == 1.0 and r[2:5] == [1, 1, 1]
    f = lib().orc_dielectric_reflectance
    assert f(1.0, 1 / 1.5) == pytest.approx(0.04, abs=1e-3)
    assert f(0.1, 1 / 1.5) > 0.5
    # total internal reflection: from inside at a grazing angle the ray must reflect
    d = unit((1, -0.1, 0))
    for s in range(20):
        r, _ = scatter({"type": "glass", "ior": 1.5}, d, front=False, seed=s)
        assert r[8] == 1.0 and r[6] > 0


def test_lambertian_and_light():  # lambertian.test.ts, diffuseLight.test.ts, defaultMaterial.test.ts
    r, em = scatter({"type": "lambert", "color": [0.1, 0.2, 0.3]}, (0, -1, 0))
    assert r[0] == 0.0 and r[1] == 1.0 and r[2:5] == pytest.approx([0.1, 0.2, 0.3]) and em == [0, 0, 0]
    r, em = scatter({"type": "light", "emit": [15, 14, 13]}, (0, -1, 0))
    assert r is None and em == [15, 14, 13]


def test_mixed_material():  # mixedMaterial.test.ts:49-128, :214-240
    assert lib().orc_mixed_weight_clamp(1.5) == 1 and lib().orc_mixed_weight_clamp(-1) == 0 and lib().orc_mixed_weight_clamp(0.3) == 0.3
    lam = {"type": "lambert", "color": [0.5, 0.5, 0.5]}
    met = {"type": "metal", "color": [0.9, 0.9, 0.9], "fuzz": 0.0}
    for s in range(20):
        assert scatter({"type": "mixed", "diff": lam, "spec": met, "weight": 1.0}, (1, -1, 0), seed=s)[0][1] == 1.0
        assert scatter({"type": "mixed", "diff": lam, "spec": met, "weight": 0.0}, (1, -1, 0), seed=s)[0][0] == 1.0
    picks = [scatter({"type": "mixed", "diff": lam, "spec": met, "weight": 0.3}, (1, -1, 0), seed=s)[0][1] for s in range(2000)]
    assert np.mean(picks) == pytest.approx(0.3, abs=0.05)
    _, em = scatter({"type": "mixed", "diff": {"type": "light", "emit": [1, 2, 3]}, "spec": {"type": "light", "emit": [4, 5, 6]}, "weight": 0.3}, (0, -1, 0))
    assert em == pytest.approx([0.3 * 1 + 0.7 * 4, 0.3 * 2 + 0.7 * 5, 0.3 * 3 + 0.7 * 6], abs=1e-5)


def test_layered_material():  # layeredMaterial.test.ts:57-123, :188-200
    glass = {"type": "glass", "ior": 1.5}
    paint = {"type": "layered", "outer": glass, "inner": {"type": "lambert", "color": [0.7, 0.3, 0.3]}}
    kinds = set()
    for s in range(300):
        r, _ = scatter(paint, unit((1, -0.3, 0)), seed=s)
        if r[0] == 1.0:  # reflected off the coat: white attenuation + a ray
            assert r[2:5] == [1, 1, 1] and r[8] == 1.0
            kinds.add("reflect")
        else:            # refracted: the inner Lambertian's pdf result
            assert r[1] == 1.0 and r[2:5] == pytest.approx([0.7, 0.3, 0.3])
            kinds.add("inner")
    assert kinds == {"reflect", "inner"}
    coated = {"type": "layered", "outer": glass, "inner": {"type": "metal", "color": [0.8, 0.8, 0.8], "fuzz": 0.0}}
    got_inner = False
    for s in range(100):
        r, _ = scatter(coated, (0, -1, 0), seed=s)
        if r is not None and r[8] == 0.0:
            assert r[0] == 1.0 and r[2:5] == pytest.approx([0.8, 0.8, 0.8])
            got_inner = True
    assert got_inner
    _, em = scatter({"type": "layered", "outer": glass, "inner": {"type": "light", "emit": [2, 3, 4]}}, (0, -1, 0))
    assert em == [2, 3, 4]


# ---------------------------------------------------------------- tests/camera.test.ts
def _empty_world(camera=None):
    sd = _sphere_scene([((0, 0, 1e6), 1e-3)])  # nothing any test ray can reach ("emptyWorld")
    if camera:
        sd["camera"].update(camera)
    return sd


def test_camera_dimensions():  # camera.test.ts:161-167, :820-880
    assert (lambda c: (c.imageWidth, c.imageHeight))(ob.OracleCamera(_empty_world(), {})) == (400, 225)
    for aspect, h in ((1.0, 400), (16 / 9, 225), (4 / 3, 300)):
        c = ob.OracleCamera(_empty_world(), {"width": 400, "aspect": aspect})
        assert (c.imageWidth, c.imageHeight) == (400, h)


def test_camera_rays_identical_without_jitter():  # camera.test.ts:202-222
    c = ob.OracleCamera(_empty_world(), {"width": 20, "samples": 1})
    o1, d1 = c.get_ray(3, 4, sample=0)
    o2, d2 = c.get_ray(3, 4, sample=5)
    assert np.array_equal(o1, o2) and np.array_equal(d1, d2)
    c = ob.OracleCamera(_empty_world(), {"width": 20, "samples": 4})
    _, d1 = c.get_ray(3, 4, sample=0)
    _, d2 = c.get_ray(3, 4, sample=1)
    assert not np.array_equal(d1, d2)


def test_camera_defocus_offsets_origin():  # camera.test.ts:225-330
    c = ob.OracleCamera(_empty_world({"aperture": 2.0, "focus": 10.0}), {"width": 20, "samples": 4})
    origins = np.array([c.get_ray(5, 5, sample=s)[0] for s in range(50)])
    assert np.any(np.abs(origins) > 1e-3)
    assert np.all(np.linalg.norm(origins, axis=1) <= 1.0 + 1e-5)  # inside the aperture/2 disk
    c0 = ob.OracleCamera(_empty_world({"aperture": 0.0}), {"width": 20, "samples": 4})
    assert np.array_equal(c0.get_ray(5, 5, sample=3)[0], np.zeros(3, np.float32))


def test_background_gradient():  # camera.test.ts:725-817
    c = ob.OracleCamera(_empty_world(), {})
    up, _ = c.ray_color((0, 0, 0), (0, 1, 0))
    down, _ = c.ray_color((0, 0, 0), (0, -1, 0))
    assert down[0] == pytest.approx(1.0, abs=0.05) and up[0] == pytest.approx(0.5, abs=0.05) and down.sum() > up.sum()
    c = ob.OracleCamera(_empty_world({"background": {"type": "gradient", "top": [1, 0, 0], "bottom": [0, 1, 0]}}), {})
    up, _ = c.ray_color((0, 0, 0), (0, 1, 0))
    down, _ = c.ray_color((0, 0, 0), (0, -1, 0))
    assert up[1] > up[0] and down[0] > down[1]
    c = ob.OracleCamera(_empty_world({"background": {"type": "gradient", "top": [0.5, 0, 0.5], "bottom": [0.5, 0, 0.5]}}), {})
    for d in ((0, 1, 0), (0, -1, 0)):
        col, _ = c.ray_color((0, 0, 0), d)
        assert list(col) == pytest.approx([0.5, 0, 0.5], abs=1e-5)


def test_render_stats_counts():  # camera.test.ts:332-357, :381-396
    r = ob.OracleCamera(_empty_world(), {"width": 10, "aspect": 1.0, "samples": 1}).render()
    st = r["stats"]
    assert (st.pixels, st.samples_total, st.samples_min, st.samples_max) == (100, 100, 1, 1)
    c = ob.OracleCamera(_empty_world(), {"width": 20, "aspect": 1.0, "samples": 1})
    st = c.render(region=ob.rt_region(5, 5, 10, 10))["stats"]
    assert st.pixels == 100
    st = c.render(region=ob.rt_region(15, 15, 10, 10))["stats"]  # clipped at the image edge (camera.ts:390-391)
    assert st.pixels == 25


def test_russian_roulette_bounds_bounces():  # camera.test.ts:591-655
    sd = scenes.generateCornellSceneData()
    a = ob.OracleCamera(sd, {"width": 24, "samples": 8, "aTolerance": 0, "roulette": True}).render(threads=4)["stats"]
    b = ob.OracleCamera(sd, {"width": 24, "samples": 8, "aTolerance": 0, "roulette": False, "depth": 30}).render(threads=4)["stats"]
    assert a.bounces_total < b.bounces_total
    assert b.bounces_max <= 30


# ---------------------------------------------------------------- tests/scenes/*.test.ts
def test_scene_generators_structure():
    c = scenes.generateCornellSceneData()  # scenes.cornell.test.ts:6-24
    assert len(c["objects"]) == 8 and sum(1 for o in c["objects"] if o.get("light")) == 1 and c["render"]["aspect"] == 1.0
    assert len(scenes.generateCornellSceneData({"variant": "empty"})["objects"]) == 6
    assert len(scenes.generateRainSceneData({"count": 20, "seed": 1})["objects"]) == 21  # scenes.rain.test.ts:8-13
    for n in (5, 50, 200):  # scenes.spheres.test.ts:8-63
        assert len(scenes.generateSpheresSceneData({"count": n, "seed": 9})["objects"]) == n
    r10 = scenes.generateSpheresSceneData({"count": 10, "seed": 1})["objects"][0]["r"]
    r1000 = scenes.generateSpheresSceneData({"count": 1000, "seed": 1})["objects"][0]["r"]
    assert r1000 < r10
    d = scenes.generateDefaultSceneData()  # scenes.default.test.ts
    assert len(d["objects"]) == 10 and sum(1 for o in d["objects"] if o.get("light")) == 2
    a = scenes.generateSpheresSceneData({"count": 30, "seed": 77})
    b = scenes.generateSpheresSceneData({"count": 30, "seed": 77})
    assert a == b  # deterministic for a seed


def test_seeded_random_batch_equals_scalar():
    a = scenes.SeededRandom(12345)
    b = scenes.SeededRandom(12345)
    xs = [a.next() for _ in range(1000)]
    ys = b.next_batch(1000)
    assert np.array_equal(np.array(xs), ys)
    assert a.next() == b.next()
    assert all(0 <= x < 1 for x in xs)


def test_mulberry32_known_values():
    """mulberry32(seed=1) first outputs, computed independently with exact uint32 arithmetic."""
    def mb(seed, n):
        out, s = [], seed
        for _ in range(n):
            s = (s + 0x6D2B79F5) & 0xFFFFFFFF
            t = s
            t = ((t ^ (t >> 15)) * (t | 1)) & 0xFFFFFFFF
            t ^= (t + (((t ^ (t >> 7)) * (t | 61)) & 0xFFFFFFFF)) & 0xFFFFFFFF
            out.append(((t ^ (t >> 14)) & 0xFFFFFFFF) / 4294967296)
        return out
    r = scenes.SeededRandom(1)
    assert [r.next() for _ in range(8)] == mb(1, 8)


# ---------------------------------------------------------------- output quantisation / adaptive rule
def test_write_color_quantisation():  # camera.ts:455-472 (asserted by no reference test: oracle-pinned)
    out = (C.c_uint8 * 3)()
    lib().orc_write_color(d3((0.0, 0.25, 1.0)), out)
    assert list(out) == [0, 127, 255]
    lib().orc_write_color(d3((4.0, float("nan"), -1.0)), out)
    assert list(out) == [255, 0, 0]


def test_pixel_converged_rule():  # camera.ts:348-368
    c = ob.OracleCamera(_empty_world(), {"samples": 100, "aTolerance": 0.05, "aBatch": 10})
    pc = lib().orc_pixel_converged
    assert pc(c.h, 1, 1.0, 1.0) == 0            # < 2 samples
    assert pc(c.h, 7, 7.0, 7.0) == 0            # not on a batch boundary
    assert pc(c.h, 10, 10.0, 10.0) == 1         # zero variance => converged
    # mean 1, sample variance 1 at n=10: CI = 1.96/sqrt(10) = 0.62 > 0.05
    assert pc(c.h, 10, 10.0, 19.0) == 0
    c0 = ob.OracleCamera(_empty_world(), {"samples": 100, "aTolerance": 0.0})
    assert pc(c0.h, 10, 10.0, 10.0) == 0        # adaptive off


# ---------------------------------------------------------------- Vec3 / Ray / Interval on their own
def _v3(op, a, b=None, s=0.0):
    out = (C.c_float * 3)()
    val = lib().orc_vec3_op(op, d3(a), d3(b) if b is not None else None, float(s), out)
    return list(out), val


NEG, ADD, SUB, MUL, MULV, DIV, CROSS, UNIT, LEN2, LEN, DOT, NEARZERO, ILLUM = range(13)


def test_vec3_operator_methods():  # tests/geometry/vec3.test.ts:34-58 (exact on small integers)
    v1, v2 = (1, 2, 3), (4, 5, 6)
    assert _v3(NEG, v1)[0] == [-1, -2, -3]
    assert _v3(ADD, v1, v2)[0] == [5, 7, 9]
    assert _v3(SUB, v1, v2)[0] == [-3, -3, -3] and _v3(SUB, v2, v1)[0] == [3, 3, 3]
    assert _v3(MUL, v1, s=2)[0] == [2, 4, 6] and _v3(MUL, v1, s=0)[0] == [0, 0, 0]
    assert _v3(MULV, v1, v2)[0] == [4, 10, 18]
    assert _v3(DIV, (2, 4, 6), s=2)[0] == [1, 2, 3]
    # vec3.ts:60-70: negate never yields -0 (toEqual distinguishes +0 / -0 in Jest)
    assert all(math.copysign(1.0, x) == 1.0 for x in _v3(NEG, (0, 0, 0))[0])


def test_vec3_magnitude_dot_cross_unit():  # tests/geometry/vec3.test.ts:81-111, :183-218
    v = (3, 4, 0)
    assert _v3(LEN2, v)[1] == 25 and _v3(LEN2, (0, 0, 0))[1] == 0
    assert _v3(LEN, v)[1] == pytest.approx(5, abs=5e-3) and _v3(LEN, (0, 0, 0))[1] == 0
    assert _v3(LEN, (1, 1, 1))[1] == pytest.approx(math.sqrt(3), abs=5e-3)
    assert _v3(DOT, v, (4, -5, 6))[1] == -8
    assert _v3(CROSS, v, (0, 0, 1))[0] == [4, -3, 0]
    u = _v3(UNIT, v)[0]
    assert u == pytest.approx([0.6, 0.8, 0.0], abs=5e-3) and _v3(LEN, u)[1] == pytest.approx(1, abs=5e-3)
    v1, v2 = (1, 2, 3), (4, -5, 6)
    i, j, k = (1, 0, 0), (0, 1, 0), (0, 0, 1)
    assert _v3(DOT, v1, v2)[1] == 12 and _v3(DOT, v1, (0, 0, 0))[1] == 0
    assert _v3(DOT, i, j)[1] == 0 and _v3(DOT, i, i)[1] == 1
    assert _v3(CROSS, v1, v2)[0] == [27, 6, -13]
    assert _v3(CROSS, i, j)[0] == list(k) and _v3(CROSS, j, i)[0] == [0, 0, -1]
    assert _v3(CROSS, v1, v1)[0] == [0, 0, 0] and _v3(CROSS, v1, (0, 0, 0))[0] == [0, 0, 0]
    assert _v3(UNIT, i)[0] == [1, 0, 0]


def test_vec3_near_zero():  # tests/geometry/vec3.test.ts:162-171
    assert _v3(NEARZERO, (1e-9, -1e-9, 1e-9))[1] == 1
    assert _v3(NEARZERO, (0.001, 0.001, 0.001))[1] == 0
    assert _v3(NEARZERO, (0, 0, 0))[1] == 1


def test_vec3_results_are_stored_in_fp32():
    """gl-matrix ARRAY_TYPE = Float32Array: every vector result is rounded to FP32 on store, scalars stay FP64
    (SURVEY.md App. A.1).  Asserted by NO reference test ("parity unpinned" for this aspect): this pins the
    oracle to the documented model so it cannot drift."""
    a, b = (0.1, 0.2, 0.3), (0.7, 0.11, 1e-9)
    fa, fb = np.float32(a), np.float32(b)
    got = _v3(ADD, a, b)[0]
    want = (fa.astype(np.float64) + fb.astype(np.float64)).astype(np.float32)  # FP64 sum of the FP32 inputs, one rounding
    assert got == [float(x) for x in want]
    assert _v3(DOT, a, b)[1] == float(np.sum(fa.astype(np.float64) * fb.astype(np.float64)[[0, 1, 2]]))  # FP64, not rounded
    d = _v3(DOT, a, b)[1]
    assert d != float(np.float32(d))  # i.e. really more than FP32 precision
    u = _v3(UNIT, a)[0]
    inv = 1.0 / math.sqrt(float(np.sum(fa.astype(np.float64) ** 2)))  # gl-matrix normalize: multiply by 1/sqrt(len^2)
    assert u == [float(np.float32(float(x) * inv)) for x in fa]
    assert _v3(UNIT, (0, 0, 0))[0] == [0, 0, 0]  # gl-matrix leaves the zero vector at zero


def test_vec3_random_generators():  # tests/geometry/vec3.test.ts:115-159 (the two used on the render path) + unit disk
    n = 2000
    out = (C.c_float * (3 * n))()
    lib().orc_sample_vec3(0, 5, -1.0, 1.0, n, out)
    v = np.array(out, np.float64).reshape(n, 3)
    assert v.min() >= -1 and v.max() <= 1 and abs(v.mean()) < 0.05
    lib().orc_sample_vec3(1, 6, 0.0, 0.0, n, out)
    v = np.array(out, np.float64).reshape(n, 3)
    assert np.all((v ** 2).sum(1) <= 1) and np.abs(v.mean(0)).max() < 0.05
    # uniform in the ball: P(|p| < 0.5) = 1/8
    assert abs(np.mean((v ** 2).sum(1) < 0.25) - 0.125) < 0.03
    lib().orc_sample_vec3(2, 7, 0.0, 0.0, n, out)  # vec3.ts:357-364, used by the defocus disk (camera.ts:197-207)
    v = np.array(out, np.float64).reshape(n, 3)
    assert np.all(v[:, 2] == 0) and np.all((v ** 2).sum(1) < 1)


def test_ray_at():  # tests/geometry/ray.test.ts:41-87
    def at(t):
        out = (C.c_float * 3)()
        lib().orc_ray_at(d3((1, 2, 3)), d3((4, 5, 6)), t, out)
        return list(out)
    assert at(0) == [1, 2, 3]
    assert at(1) == [5, 7, 9]
    assert at(0.5) == [3, 4.5, 6]
    assert at(-1) == [-3, -3, -3]


def test_interval():  # tests/geometry/interval.test.ts:27-99
    iv = lib().orc_interval_op
    SIZE, CONTAINS, SURROUNDS, CLAMP = range(4)
    assert iv(SIZE, 1, 5, 0) == 4 and iv(SIZE, -2, 3, 0) == 5
    assert iv(SIZE, INF, -INF, 0) < 0 and iv(SIZE, -INF, INF, 0) == INF          # EMPTY / UNIVERSE
    assert all(iv(CONTAINS, 1, 5, x) == 1 for x in (1, 3, 5)) and all(iv(CONTAINS, 1, 5, x) == 0 for x in (0.9, 5.1))
    assert all(iv(CONTAINS, -INF, INF, x) == 1 for x in (0, -1e10, 1e10)) and iv(CONTAINS, INF, -INF, 0) == 0
    assert all(iv(SURROUNDS, 1, 5, x) == 1 for x in (1.1, 3, 4.9))
    assert all(iv(SURROUNDS, 1, 5, x) == 0 for x in (1, 5, 0.9, 5.1))             # strict: what every hit test relies on
    assert all(iv(SURROUNDS, -INF, INF, x) == 1 for x in (0, -1e10, 1e10)) and iv(SURROUNDS, INF, -INF, 0) == 0
    assert [iv(CLAMP, 1, 5, x) for x in (3, 1, 5, 0, -10, 6, 100)] == [3, 1, 5, 1, 1, 5, 5]
    assert iv(CLAMP, -INF, INF, 100) == 100 and iv(CLAMP, -INF, INF, -100) == -100
    assert iv(CLAMP, INF, -INF, 0) == INF


def test_ray_color_edge_cases():  # camera.test.ts:441-487 (PDF materials), :518-590 (roulette), :658-720 (attenuation 0 / 2)
    """The reference asserts "does not throw, components >= 0"; the restatement adds what follows from the code:
    a black Lambertian returns exactly zero radiance, an albedo of 2 stays finite (continuation probability is
    capped at 0.95, camera.ts:236), and every path ends within `depth` bounces."""
    def cam(albedo, **render):
        sd = _sphere_scene([((0, 0, -1), 0.5)])
        sd["materials"][0]["material"]["color"] = [albedo] * 3
        return ob.OracleCamera(sd, {"width": 10, "aspect": 1.0, "samples": 1, **render})
    for s in range(64):
        rgb, bounces = cam(0.0, roulette=True, rouletteDepth=1).ray_color((0, 0, 0), (0, 0, -1), sample=s)
        assert np.all(rgb == 0) and 1 <= bounces <= 100
        rgb, bounces = cam(2.0, roulette=True, rouletteDepth=1, depth=40).ray_color((0, 0, 0), (0, 0, -1), sample=s)
        assert np.all(np.isfinite(rgb)) and np.all(rgb >= 0) and bounces <= 40
        rgb, bounces = cam(0.5, roulette=True, rouletteDepth=5).ray_color((0, 0, 0), (0, 0, -1), sample=s)
        assert np.all(np.isfinite(rgb)) and np.all(rgb >= 0) and bounces >= 1
    # roulette off: only the depth cut ends a path that keeps hitting (camera.ts:228), and it returns black
    inside = _sphere_scene([((0, 0, 0), 5.0)])  # the camera sits inside a closed grey sphere: no ray ever escapes
    c = ob.OracleCamera(inside, {"width": 10, "aspect": 1.0, "samples": 1, "roulette": False, "depth": 7})
    rgb, bounces = c.ray_color((0, 0, 0), (0, 0, -1))
    assert bounces == 7 and np.all(rgb == 0)
