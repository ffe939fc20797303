"""Synthetic diagonal-Gaussian posteriors shaped like PLN latents (SURVEY.md section 8d recipe,
probe-calibrated to Kodak-like rates: mean ~0.76 bits/dim, p99 ~5.5 bits).  Used by bench.py and tests."""
import numpy as np

# latent sizes of the two PLN levels (transforms.py strides; miracle_arguments.py:219-222)
def pln_level_sizes(height, width, latent_channels1=128, latent_channels2=24):
    h1, w1 = -(-height // 16), -(-width // 16)
    h2, w2 = -(-h1 // 4), -(-w1 // 4)
    return h1 * w1 * latent_channels1, h2 * w2 * latent_channels2


def synthetic_posterior(n, image_index=0, nontrivial_prior=False):
    """-> (q_loc, q_scale, p_loc, p_scale) float32[n]."""
    rng = np.random.default_rng(20261018 + int(image_index))
    q_scale = np.exp(-np.abs(rng.normal(0.0, 0.7, n)))
    q_loc = 1.2 * rng.normal(size=n) * (rng.random(n) < 0.35)
    if nontrivial_prior:
        p_loc = rng.normal(size=n)
        p_scale = np.log1p(np.exp(rng.normal(size=n)))
        q_loc = p_loc + p_scale * q_loc
        q_scale = p_scale * q_scale
    else:
        p_loc = np.zeros(n)
        p_scale = np.ones(n)
    f = np.float32
    return q_loc.astype(f), q_scale.astype(f), p_loc.astype(f), p_scale.astype(f)


class Normal:
    """Minimal stand-in for tfd.Normal at the reference boundary (.loc, .scale, .dtype)."""

    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale
        self.dtype = getattr(loc, "dtype", np.float32)
