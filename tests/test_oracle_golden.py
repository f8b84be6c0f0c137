"""The oracle must keep reproducing its committed fixtures bit for bit (guards the checker),
and the host-side logic (flattening, option merge, error behaviour) is checked on CPU."""
import os
import sys

import numpy as np
import pytest

import oracle_binding as ob
from mcp_raytracer_b200 import scenes
from mcp_raytracer_b200.scene_data import (
    DEFAULT_RENDER_DATA, FlatScene, RaytracerError, merge_render_options, render_opts_struct,
    RT_MAT_GLASS, RT_MAT_LAMBERT, RT_MAT_LAYERED, RT_MAT_LIGHT, RT_MAT_METAL, RT_MAT_MIXED,
)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import make_golden  # noqa: E402


@pytest.mark.parametrize("name", list(make_golden.PRIMARY))
def test_oracle_primary_fixture(name):
    fn, width = make_golden.PRIMARY[name]
    f = np.load(os.path.join(GOLDEN, f"primary_{name}.npz"))
    ids, t, nrm, ff = ob.OracleCamera(fn(), {"width": width, "samples": 1}).trace_primary()
    assert np.array_equal(ids, f["ids"]) and np.array_equal(t, f["t"]) and np.array_equal(nrm, f["normal"]) and np.array_equal(ff, f["front"])


@pytest.mark.parametrize("name", list(make_golden.RENDERS))
def test_oracle_render_fixture(name):
    fn, opts, seed = make_golden.RENDERS[name]
    f = np.load(os.path.join(GOLDEN, f"render_{name}.npz"))
    r = ob.OracleCamera(fn(), opts).render(seed=seed, threads=3)  # strips must not change the image
    st = r["stats"]
    assert np.array_equal(r["linear"], f["linear"]) and np.array_equal(r["rgb8"], f["rgb8"])
    assert [st.pixels, st.samples_total, st.samples_min, st.samples_max, st.bounces_total, st.bounces_min, st.bounces_max, st.rays] == list(f["stats"])


def test_oracle_sequential_rng_is_statistically_the_same_image():
    """mode 1 (one sequential stream per strip, like one Math.random per worker) vs the path-keyed
    Philox streams: same estimator, so the two images agree within Monte-Carlo noise."""
    sd = scenes.generateCornellSceneData()
    opts = {"width": 24, "samples": 256, "aTolerance": 0}
    a = ob.OracleCamera(sd, opts).render(rng_mode=0, seed=1, threads=8)["linear"].astype(np.float64)
    b = ob.OracleCamera(sd, opts).render(rng_mode=1, seed=1, threads=8)["linear"].astype(np.float64)
    assert abs(a.mean() - b.mean()) < 0.03 * a.mean()


def test_divide_into_regions():  # raytracer.ts:185-205
    from mcp_raytracer_b200.raytracer import divideIntoRegions

    r = divideIntoRegions(100, 10, 3)
    assert [(x["y"], x["height"]) for x in r] == [(0, 4), (4, 4), (8, 2)]
    assert len(divideIntoRegions(50, 3, 8)) == 3  # stops when rows are exhausted
    assert sum(x["height"] for x in divideIntoRegions(7, 225, 7)) == 225


def test_flatten_materials_and_lights():
    fs = FlatScene(scenes.generateDefaultSceneData())
    assert fs.desc.n_objects == 10
    types = list(fs.mat_type_a)
    assert types.count(RT_MAT_LAYERED) == 1 and types.count(RT_MAT_LIGHT) == 1
    li = types.index(RT_MAT_LAYERED)
    inner, outer = fs.mat_child_a[li]
    assert fs.mat_type_a[inner] == RT_MAT_LAMBERT and fs.mat_type_a[outer] == RT_MAT_GLASS
    assert inner < li and outer < li  # children precede parents (the library relies on it)
    assert list(fs.obj_light) == [0] * 8 + [1, 1]
    assert fs.obj_r[6] == -0.24
    # string references are shared, inline definitions are not
    assert fs.obj_material[8] == fs.obj_material[9]
    fs5 = FlatScene(scenes.generateLayeredMixedSceneData())
    t5 = list(fs5.mat_type_a)
    assert RT_MAT_MIXED in t5 and RT_MAT_METAL in t5 and t5.count(RT_MAT_LAYERED) == 3


def test_flatten_error_behaviour():  # scenes.ts:137,154,178,191,195
    base = scenes.generateCornellSceneData()
    bad = dict(base, objects=[dict(base["objects"][0], type="torus")])
    with pytest.raises(RaytracerError, match="Unknown object type: torus"):
        FlatScene(bad)
    bad = dict(base, objects=[dict(base["objects"][0], material="nope")])
    with pytest.raises(RaytracerError, match="Material not found: nope"):
        FlatScene(bad)
    bad = dict(base, objects=[dict(base["objects"][0], material={"type": "velvet"})])
    with pytest.raises(RaytracerError, match="Unknown material type: velvet"):
        FlatScene(bad)
    bad = dict(base, objects=[dict(base["objects"][0], material={"type": "layered", "outer": "red", "inner": "white"})])
    with pytest.raises(RaytracerError, match="Material is not a dielectric: red"):
        FlatScene(bad)


def test_option_merge_order():  # camera.ts:73-83,116 ; scenes.ts:97-100
    o = merge_render_options({"aspect": 1.0, "rouletteDepth": 5}, {"width": 64, "samples": None})
    assert o["aspect"] == 1.0 and o["rouletteDepth"] == 5 and o["width"] == 64
    assert o["samples"] == DEFAULT_RENDER_DATA["samples"] == 100 and o["depth"] == 100 and o["aTolerance"] == 0.05
    s = render_opts_struct(o)
    assert (s.width, s.samples, s.depth, s.a_batch, s.roulette, s.roulette_depth, s.mode) == (64, 100, 100, 10, 1, 5, 0)
    with pytest.raises(RaytracerError):
        render_opts_struct(dict(o, mode="heatmap"))
