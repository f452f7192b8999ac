"""Parity bookkeeping shared by tests/ and bench.py's checker leg — TEST INFRASTRUCTURE ONLY (see oracle.py).

Bars (BASELINE.json north_star): segment ids agree on >= 99.9 % of rays, relative hit t <= 1e-4, image PSNR >= 45 dB.
Both sides execute one fp32 operation sequence, so the tests additionally demand bit-identical records."""
import numpy as np

SEGMENT_AGREEMENT_MIN = 0.999
REL_T_MAX = 1e-4
PSNR_MIN_DB = 45.0


def stratified_pixels(width, height, n, seed=0x5EED):
    """One seeded random pixel from each of n equal strata of the row-major frame -> sorted uint64[n]."""
    total = width * height
    n = min(n, total)
    rng = np.random.default_rng(seed)
    edges = np.linspace(0, total, n + 1).astype(np.int64)
    return (edges[:-1] + (rng.random(n) * np.maximum(1, np.diff(edges))).astype(np.int64)).astype(np.uint64)


def psnr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    mse = np.mean((a - b) ** 2)
    return 99.0 if mse == 0 else float(10.0 * np.log10(255.0 ** 2 / mse))


def parity_metrics(hits_gpu, hits_oracle, rgba_gpu=None, rgba_oracle=None):
    """-> dict for the bench line / test assertions.  hits_*: HIT_DTYPE arrays over the same rays."""
    hg, ho = np.asarray(hits_gpu), np.asarray(hits_oracle)
    assert hg.shape == ho.shape
    same_bytes = hg.view(np.uint8).reshape(-1, 32) == ho.view(np.uint8).reshape(-1, 32)
    n_diff = int((~same_bytes.all(axis=1)).sum())
    hit_g, hit_o = (hg["flags"] & 1).astype(bool), (ho["flags"] & 1).astype(bool)
    agree = float(np.mean(hg["segment"] == ho["segment"])) if hg.size else 1.0
    both = hit_g & hit_o & (hg["segment"] == ho["segment"])
    max_rel_t = float((np.abs(hg["t"][both] - ho["t"][both]) / np.abs(ho["t"][both])).max()) if both.any() else 0.0
    out = {"rays_checked": int(hg.size), "bit_identical": n_diff == 0, "records_differing": n_diff,
           "segment_agreement": agree, "max_rel_t": max_rel_t, "hit_fraction": float(hit_o.mean()) if ho.size else 0.0}
    if rgba_gpu is not None and rgba_oracle is not None:
        out["psnr"] = psnr(rgba_gpu, rgba_oracle)
        out["rgba_identical"] = bool(np.array_equal(rgba_gpu, rgba_oracle))
    out["within_tolerance"] = bool(agree >= SEGMENT_AGREEMENT_MIN and max_rel_t <= REL_T_MAX and out.get("psnr", 99.0) >= PSNR_MIN_DB)
    return out
