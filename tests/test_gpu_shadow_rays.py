"""RT_LIGHTS_SHADOW_RAYS (`lightSampling: "shadowRays"`, include/rt_b200.h): next-event estimation in the pixel-stream
kernels.  It is NOT the reference's estimator (src/camera.ts:285-315 draws one ray from a cosine/light mixture), but it
has the same expectation per pixel, so the bar is the converged-image one of the north star: the shadow-ray image at
1024 spp against the oracle's MIXTURE estimator at 2048 spp with an independent seed, through the same statistics and
bars as the BASELINE-size parity tests (tests/parity_stats.py — nothing subtracted).  Scenes: the Cornell box (quad
light, glass), the default scene (quad AND sphere light, unbounded plane), a Cornell box with Mixed and Layered
spheres added, every tree kind.
"""
import copy
import json

import numpy as np
import pytest

import oracle_binding as ob
import parity_stats as ps
from mcp_raytracer_b200 import createCameraFromSceneData, generateCornellSceneData, generateDefaultSceneData

pytestmark = pytest.mark.gpu


def cornell_composites():
    """Cornell box whose two spheres are a Mixed (Lambert | Metal) and a Layered (glass over Lambert) material."""
    sd = copy.deepcopy(generateCornellSceneData())
    sd["materials"] += [
        {"id": "mix", "material": {"type": "mixed", "diff": {"type": "lambert", "color": [0.2, 0.4, 0.8]},
                                   "spec": {"type": "metal", "color": [0.9, 0.9, 0.9], "fuzz": 0.1}, "weight": 0.6}},
        {"id": "coat", "material": {"type": "layered", "outer": {"type": "glass", "ior": 1.5}, "inner": {"type": "lambert", "color": [0.8, 0.3, 0.2]}}},
    ]
    spheres = [o for o in sd["objects"] if o["type"] == "sphere"]
    spheres[0]["material"], spheres[1]["material"] = "mix", "coat"
    return sd


SCENES = {
    "cornell": (generateCornellSceneData, {"width": 96}),
    "default": (generateDefaultSceneData, {"width": 128}),
    "cornell-composites": (cornell_composites, {"width": 80}),
}


def render_moments(sd, opts):
    with createCameraFromSceneData(sd, opts) as cam:
        W, H = cam.imageWidth, cam.imageHeight
        rgb = np.zeros((H, W, 3), np.uint8)
        lin = np.zeros((H, W, 3), np.float32)
        mom = np.zeros((H, W, 8), np.float32)
        st = cam.renderRegion(rgb, None, lin, mom)
        kind = cam.info.bvh_kind
    return rgb, lin, mom, st, kind


@pytest.mark.parametrize("bvh", ["auto", "sah", "reference"])
@pytest.mark.parametrize("name", list(SCENES))
def test_shadow_ray_image_converges_to_the_reference_estimator(gpu, name, bvh):
    """Against the ORACLE (the reference's mixture estimator, 2048 spp, independent seed).  The two images come from different
    estimators: the shadow-ray image has 3-10x less variance, so sigma of the difference is the oracle's — a sample variance
    that underestimates wherever no firefly landed — and the z-score bars of the same-estimator tests (mean z^2 <= 1.15,
    99.95 % within 4 sigma) do not transfer; the RMSE, 1 % and mean-radiance bars do, unchanged.  The sharp bias test is the
    next one (GPU against GPU at 4x the samples)."""
    make, ropts = SCENES[name]
    if bvh != "auto" and name != "cornell":
        pytest.skip("forced tree kinds are covered on the Cornell box (AUTO = LIST there; the default scene's AUTO is the reference topology)")
    sd = make()
    n_g, n_o = 1024, 2048
    opts = {**ropts, "aTolerance": 0, "bvh": bvh}
    rgb, lin, mom, st, kind = render_moments(sd, {**opts, "samples": n_g, "seed": 5, "lightSampling": "shadowRays"})
    g_var = mom[..., 3:6].astype(np.float64) / (n_g - 1.0)
    o = ob.OracleCamera(sd, {**ropts, "aTolerance": 0, "samples": n_o}).render(seed=6, threads=8, want_moments=True)
    o_mean, o_var = ps.moments_to_mean_var(o["moments"], float(n_o))
    r = ps.compare_converged(lin, g_var, n_g, o_mean, o_var, n_o)
    r.update({"case": f"shadow-rays {name} bvh={bvh} (kind {kind})", "rays_per_path": st.rays / st.samples["total"]})
    print(json.dumps({k: v for k, v in r.items() if k != "worst_z"}))
    assert st.samples["total"] == n_g * lin.shape[0] * lin.shape[1]
    assert r["rel_rmse_raw"] <= 1.10 * r["rel_rmse_noise_floor"] + 1e-6, r      # raw per-pixel RMSE: all Monte-Carlo noise
    assert r["rel_rmse_block"]["8"] <= 0.01 and r["rel_rmse_block"]["16"] <= 0.01, r   # the 1 % bar on box-filtered images
    assert r["frac_within_3sigma"] >= 0.99 and r["z2_mean"] <= 1.35, r
    for a, b in zip(r["mean_gpu"], r["mean_oracle"]):
        assert abs(a - b) <= 0.005 * abs(b) + 1e-6, r
    assert r["quiet_max_excess"] <= 0.0 and r["gpu_nonfinite_pixels"] == 0, r


@pytest.mark.parametrize("name", list(SCENES))
def test_shadow_ray_estimator_has_no_bias_against_the_mixture_estimator(gpu, name):
    """Both estimators on the GPU at 4096 / 8192 spp (the mixture kernel is the one pinned to the oracle): 16x16-block means
    carry ~0.1 % noise, so a bias of a few tenths of a percent anywhere in the image shows."""
    make, ropts = SCENES[name]
    sd = make()
    opts = {**ropts, "aTolerance": 0}
    n_s, n_m = 4096, 8192
    _, lin_s, mom_s, _, _ = render_moments(sd, {**opts, "samples": n_s, "seed": 11, "lightSampling": "shadowRays"})
    _, lin_m, mom_m, _, _ = render_moments(sd, {**opts, "samples": n_m, "seed": 12})
    H, W = lin_s.shape[:2]
    b = 16
    Hb, Wb = H // b * b, W // b * b

    def blocks(a):
        return a[:Hb, :Wb].reshape(Hb // b, b, Wb // b, b, 3).astype(np.float64)

    d = blocks(lin_s).mean(axis=(1, 3)) - blocks(lin_m).mean(axis=(1, 3))
    var = (blocks(mom_s[..., 3:6]) / (n_s - 1.0) / n_s + blocks(mom_m[..., 3:6]) / (n_m - 1.0) / n_m).sum(axis=(1, 3)) / b**4
    z = d / np.sqrt(var + 1e-14)
    ref = float(np.sqrt(np.mean(blocks(lin_m).mean(axis=(1, 3)) ** 2)))
    rel = float(np.sqrt(np.mean(d**2))) / ref
    print(f"{name}: block-mean rel RMSE {rel:.5f}, z^2 mean {float(np.mean(z**2)):.3f}, max |z| {float(np.abs(z).max()):.2f}, "
          f"image means {lin_s.mean():.6f} vs {lin_m.mean():.6f}")
    assert rel <= 0.003
    assert float(np.mean(z**2)) <= 1.6 and float(np.abs(z).max()) <= 5.0
    assert abs(float(lin_s.mean()) - float(lin_m.mean())) <= 0.002 * float(lin_m.mean())


def test_shadow_rays_lower_the_variance_on_the_cornell_box(gpu):
    """The point of the estimator: per-sample variance of the directly lit walls under a small light."""
    sd = generateCornellSceneData()
    base = {"width": 96, "samples": 256, "aTolerance": 0, "seed": 3}
    _, lin_m, mom_m, st_m, _ = render_moments(sd, base)
    _, lin_s, mom_s, st_s, _ = render_moments(sd, {**base, "lightSampling": "shadowRays"})
    var_m, var_s = float(mom_m[..., 3:6].mean()), float(mom_s[..., 3:6].mean())
    print(f"mean per-sample variance: mixture {var_m / 255:.4f}, shadow rays {var_s / 255:.4f}; rays/path {st_m.rays / st_m.samples['total']:.2f} vs {st_s.rays / st_s.samples['total']:.2f}")
    assert var_s < 0.7 * var_m
    assert st_s.rays > st_m.rays                                        # the shadow rays are counted (Grays/s)
    assert abs(float(lin_s.mean()) - float(lin_m.mean())) < 0.02 * float(lin_m.mean())


def test_shadow_rays_are_deterministic_and_partition_independent(gpu):
    sd = generateCornellSceneData()
    opts = {"width": 64, "samples": 32, "aTolerance": 0, "seed": 9, "lightSampling": "shadowRays"}
    rgb0, lin0, _, st0, _ = render_moments(sd, opts)
    rgb1, lin1, _, _, _ = render_moments(sd, opts)
    assert np.array_equal(lin0, lin1) and np.array_equal(rgb0, rgb1)
    merged = np.zeros_like(rgb0)
    paths = 0
    for part in range(3):
        with createCameraFromSceneData(sd, {**opts, "partIndex": part, "partCount": 3}) as cam:
            buf = np.zeros_like(rgb0)
            st = cam.render(buf)
            paths += st.samples["total"]
            merged |= buf
    assert np.array_equal(merged, rgb0) and paths == st0.samples["total"]


def test_shadow_rays_with_the_reference_defaults_adaptive_and_modes(gpu):
    """Adaptive sampling (the reference's default) and the render modes run through the same kernel."""
    sd = generateDefaultSceneData()
    with createCameraFromSceneData(sd, {"width": 160, "samples": 200, "lightSampling": "shadowRays"}) as cam:   # aTolerance 0.05, aBatch 10
        rgb = np.zeros((cam.imageHeight, cam.imageWidth, 3), np.uint8)
        st = cam.render(rgb)
    assert st.pixels == rgb.shape[0] * rgb.shape[1]
    assert st.samples["min"] >= 10 and st.samples["min"] % 10 == 0 and st.samples["max"] <= 200
    assert st.samples["total"] < 200 * st.pixels                        # some pixels converged early
    with createCameraFromSceneData(generateCornellSceneData(), {"width": 48, "samples": 16, "mode": "bounces", "lightSampling": "shadowRays"}) as cam:
        rgb = np.zeros((cam.imageHeight, cam.imageWidth, 3), np.uint8)
        cam.render(rgb)
    assert rgb[..., 2].max() > 0 and rgb[..., 0].max() == 0
