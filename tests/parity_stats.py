"""Statistics of the converged-image bar (north star: ">= 1024 spp within 1% RMSE and within 3 sigma per pixel
of the reference's high-spp render") — TEST INFRASTRUCTURE shared by tests/test_gpu_full_parity.py and
scripts/gpu_full_parity.py.

Two renders with INDEPENDENT random streams are compared: the GPU image (n_g samples per pixel) and the CPU
oracle's (n_o >= 2 n_g).  Both deliver per-pixel moments sum(c), sum(c^2), so each pixel's Monte-Carlo
standard error is measured, not assumed.  Nothing is subtracted from any asserted number: the raw RMSE is
reported beside the noise floor those standard errors predict for two unbiased estimates of the same image.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict

import numpy as np


def moments_to_mean_var(mom: np.ndarray, n: float):
    """[H,W,8] (sum rgb, sum rgb^2, samples, bounces) -> per-channel mean and UNBIASED sample variance."""
    m = mom.astype(np.float64)
    mean = m[..., 0:3] / n
    var = np.maximum(m[..., 3:6] / n - mean**2, 0.0) * n / (n - 1.0)
    return mean, var


def block_mean(a: np.ndarray, b: int) -> np.ndarray:
    """Means over b x b pixel blocks (ragged edges dropped)."""
    H, W = a.shape[:2]
    h, w = H // b * b, W // b * b
    return a[:h, :w].reshape(h // b, b, w // b, b, -1).mean(axis=(1, 3))


def compare_converged(g_mean: np.ndarray, g_var: np.ndarray, n_g: int, o_mean: np.ndarray, o_var: np.ndarray, n_o: int) -> Dict[str, Any]:
    """g_mean / o_mean: [H,W,3] mean radiance of each render; g_var / o_var: per-sample variance of each.

    Pixels whose ORACLE mean is not finite are counted and left out: one NaN sample poisons a pixel in the reference
    (PixelStats.add sums it, writeColorToBuffer stores NaN as 0 = a black pixel, src/camera.ts:455-472); the device
    adds such a sample as 0 instead, so those pixels are not comparable by construction."""
    g_mean, g_var, o_mean, o_var = (np.asarray(a, np.float64) for a in (g_mean, g_var, o_mean, o_var))
    ok = np.isfinite(o_mean).all(axis=-1) & np.isfinite(o_var).all(axis=-1)
    n_bad = int((~ok).sum())
    H, W = g_mean.shape[:2]
    if n_bad:  # neutralise them: difference 0, variance 0 (they are reported, and bounded by check_converged)
        g_mean, g_var, o_mean, o_var = g_mean.copy(), g_var.copy(), o_mean.copy(), o_var.copy()
        o_mean[~ok] = g_mean[~ok]
        o_var[~ok] = 0.0
        g_var[~ok] = 0.0
    diff = g_mean - o_mean
    sig2 = g_var / n_g + o_var / n_o           # variance of the difference of two independent estimates
    ref_rms = float(np.sqrt(np.mean(o_mean**2)))
    out: Dict[str, Any] = {
        "n_gpu": n_g, "n_oracle": n_o, "pixels": int(H * W), "oracle_nonfinite_pixels": n_bad,
        "gpu_nonfinite_pixels": int((~np.isfinite(g_mean).all(axis=-1)).sum()), "ref_rms": ref_rms,
        # RAW relative RMSE of the two images, per pixel, and what pure Monte-Carlo noise predicts for it
        "rel_rmse_raw": float(np.sqrt(np.mean(diff**2))) / ref_rms,
        "rel_rmse_noise_floor": float(np.sqrt(np.mean(sig2))) / ref_rms,
        # the same RAW figure on box-filtered images: noise falls with the block size, a bias would not
        "rel_rmse_block": {},
        "rel_rmse_block_noise_floor": {},
        # whole-image mean radiance per channel
        "mean_gpu": [float(x) for x in g_mean.mean(axis=(0, 1))],
        "mean_oracle": [float(x) for x in o_mean.mean(axis=(0, 1))],
    }
    for b in (4, 8, 16):
        if min(diff.shape[:2]) >= 2 * b:
            out["rel_rmse_block"][str(b)] = float(np.sqrt(np.mean(block_mean(diff, b) ** 2))) / ref_rms
            out["rel_rmse_block_noise_floor"][str(b)] = float(np.sqrt(np.mean(block_mean(sig2, b)) / (b * b))) / ref_rms
    sigma = np.sqrt(sig2)
    slack = 2e-6 + 1e-5 * np.abs(o_mean)      # FP32 vs FP64-scalar rounding of a colour both sides agree on (no variance)
    lit = sigma > slack                       # channels whose Monte-Carlo error is above that rounding level
    z = np.zeros_like(diff)
    z[lit] = diff[lit] / sigma[lit]
    out["frac_within_3sigma"] = float(np.mean(np.abs(diff) <= 3.0 * sigma + slack))
    out["frac_within_4sigma"] = float(np.mean(np.abs(diff) <= 4.0 * sigma + slack))
    out["z2_mean"] = float(np.mean(z[lit] ** 2)) if lit.any() else 0.0   # 1.0 for unbiased estimates with honest variances
    out["z_mean"] = float(np.mean(z[lit])) if lit.any() else 0.0         # 0.0 +- 1/sqrt(N)
    out["quiet_channels"] = int((~lit).sum())                            # (near-)zero variance in both renders: must agree to rounding
    out["quiet_max_abs_diff"] = float(np.abs(diff[~lit]).max()) if (~lit).any() else 0.0
    out["quiet_max_excess"] = float((np.abs(diff[~lit]) - 4.0 * sigma[~lit] - slack[~lit]).max()) if (~lit).any() else 0.0
    worst = np.argsort(-np.abs(z), axis=None)[:8]
    out["worst_z"] = [{"y": int(i // (W * 3)), "x": int(i // 3 % W), "c": int(i % 3), "gpu": float(g_mean.flat[i]), "oracle": float(o_mean.flat[i]),
                       "sigma": float(sigma.flat[i]), "z": float(z.flat[i])} for i in worst]
    return out


def check_converged(r: Dict[str, Any], firefly_allowance: float = 0.002, one_percent_bar: bool = True) -> None:
    """The asserted bars (every number RAW, nothing subtracted):
    * per pixel |gpu - oracle| <= 3 sigma for >= 99.73 % - `firefly_allowance` of the channels (a Gaussian leaves
      0.27 % outside; heavy-tailed radiance makes the sample variance an underestimate on a few pixels), and
      >= 99.95 % within 4 sigma;
    * the raw per-pixel RMSE is what Monte-Carlo noise predicts: <= 1.10 x the noise floor, and the mean squared
      z-score lies in [0.85, 1.15] (a bias of 0.4 sigma anywhere near the image's energy would break it);
    * the 1 % bar: raw relative RMSE <= 1 % on the 8x8 and 16x16 box-filtered images (per-pixel noise at
      ~1000 spp is 1-4 % in EITHER implementation, so a per-pixel 1 % is not a statement about parity; averaging
      64 pixels divides noise by 8 and leaves any bias in place); `one_percent_bar=False` (renders of ~100 spp, where
      even the filtered noise is above 1 %) asserts instead that the filtered RMSE is explained by noise;
    * whole-image mean radiance within 0.5 % per channel;
    * channels without Monte-Carlo variance (flat background, unlit pixels) agree to FP32 rounding."""
    assert r["frac_within_3sigma"] >= 0.9973 - firefly_allowance, r
    assert r["frac_within_4sigma"] >= 0.9995, r
    assert r["rel_rmse_raw"] <= 1.10 * r["rel_rmse_noise_floor"] + 1e-6, r
    assert 0.85 <= r["z2_mean"] <= 1.15, r
    for b in ("8", "16"):
        if b in r["rel_rmse_block"]:
            if one_percent_bar:
                assert r["rel_rmse_block"][b] <= 0.01, r
            else:  # too few samples for noise to drop under 1 % even after filtering: the filtered RMSE is still all noise
                assert r["rel_rmse_block"][b] <= 1.15 * r["rel_rmse_block_noise_floor"][b] + 1e-6, r
    for a, b in zip(r["mean_gpu"], r["mean_oracle"]):
        assert abs(a - b) <= 0.005 * abs(b) + 1e-6, r
    assert r["quiet_max_excess"] <= 0.0, r                                # flat-colour channels agree to FP32 rounding
    assert r["gpu_nonfinite_pixels"] == 0, r
    # pixels the reference itself would poison with a NaN sample: a handful per billion paths at most
    assert r["oracle_nonfinite_pixels"] <= max(2, int(2e-5 * r["pixels"])), r


def oracle_trace_primary_parallel(oc, threads: int):
    """orc_trace_primary is single-threaded; row strips on Python threads (ctypes releases the GIL)."""
    import oracle_binding as ob

    H, W = oc.imageHeight, oc.imageWidth
    threads = max(1, min(threads, H))
    rows = -(-H // threads)
    strips = [(y, min(rows, H - y)) for y in range(0, H, rows)]

    def work(s):
        y, h = s
        return s, oc.trace_primary(region=ob.rt_region(0, y, W, h))

    ids = np.full((H, W), -2, np.int32)
    t = np.zeros((H, W), np.float64)
    nrm = np.zeros((H, W, 3), np.float32)
    ff = np.zeros((H, W), np.uint8)
    with ThreadPoolExecutor(threads) as ex:
        for (y, h), (i, tt, n, f) in ex.map(work, strips):
            ids[y:y + h], t[y:y + h], nrm[y:y + h], ff[y:y + h] = i[y:y + h], tt[y:y + h], n[y:y + h], f[y:y + h]
    return ids, t, nrm, ff
