"""The error bound behind the contraction kernels' arithmetic, checked on the host with emulated roundings (no GPU).

The kernels form w*g = w_hi*g_hi + (w_lo*g_hi + w_hi*g_lo) with TF32 hi parts; the bracket goes through ONE bfloat16 MMA
(synference_b200/csrc/synth3_kernel.cuh).  Claims tested here: the dropped w_lo*g_lo term is below 2^-22 of
the product; a bfloat16 factor is good to 2^-8, so each small term (2^-11 of the product) is within 2^-18 of the product and
both together within 2^-17 = 7.6e-6 AT WORST -- the bound a single wavelength of a single-bin galaxy could reach, which is
why launches that output spectra keep three TF32 passes; on sums over bins (weights and spectra are non-negative) the
typical error is 4e-7, and band integrals average the grid-side roundings further (GPU tests: 1.2e-6 at most between the
two arithmetics over 400 000 fluxes)."""
import numpy as np


def tf32_rna(x):
    """cvt.rna.tf32.f32: keep 10 mantissa bits, round to nearest, ties away from zero."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def bf16_rn(x):
    """cvt.rn.bf16.f32: keep 7 mantissa bits, round to nearest even (returned as float32)."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split(x64):
    hi = tf32_rna(x64.astype(np.float32))
    lo = (x64 - hi.astype(np.float64)).astype(np.float32)
    return hi, lo


def test_roundings_are_what_the_kernels_use():
    x = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -10, 3.14159274, 1e-20, 6.5e4], dtype=np.float32)
    t = tf32_rna(x)
    assert np.all(np.abs(t.astype(np.float64) - x) <= np.abs(x) * 2.0 ** -11)
    assert t[1] == np.float32(1.0 + 2.0 ** -10)              # tie: away from zero
    b = bf16_rn(x)
    assert np.all(np.abs(b.astype(np.float64) - x) <= np.abs(x) * 2.0 ** -8)
    assert bf16_rn(np.float32(1.0 + 2.0 ** -8)) == np.float32(1.0)      # tie: to even
    assert np.all((b.view(np.uint32) & 0xFFFF) == 0) and np.all((t.view(np.uint32) & 0x1FFF) == 0)


def test_split_product_error_bounds():
    rng = np.random.default_rng(7)
    n, k, m = 400, 102, 96
    # weights: a normalised star-formation history over many decades; spectra: positive, many decades
    w = rng.lognormal(0.0, 3.0, (n, k))
    w /= w.sum(1, keepdims=True)
    g = rng.lognormal(0.0, 2.0, (k, m))
    w_hi, w_lo = split(w)
    g_hi, g_lo = split(g)
    f = lambda a: a.astype(np.float64)   # noqa: E731
    exact = w @ g
    # per product: what is dropped, and what the bfloat16 rounding of the small terms costs
    prod = w[:, :, None] * g[None, :, :]
    dropped = np.abs(f(w_lo)[:, :, None] * f(g_lo)[None, :, :])
    assert np.max(dropped / prod) <= 2.0 ** -22
    small_tf32 = f(w_lo)[:, :, None] * f(g_hi)[None] + f(w_hi)[:, :, None] * f(g_lo)[None]
    small_bf16 = f(bf16_rn(w_lo))[:, :, None] * f(bf16_rn(g_hi))[None] + f(bf16_rn(w_hi))[:, :, None] * f(bf16_rn(g_lo))[None]
    assert np.max(np.abs(small_bf16 - small_tf32) / prod) <= 2.0 ** -17      # two terms of <= 2^-18 each
    # sums over the bins (accumulated exactly here: the tensor core's own accumulation error is what the GPU tests measure)
    hh = f(w_hi) @ f(g_hi)
    three_tf32 = hh + f(tf32_rna(w_lo)) @ f(g_hi) + f(w_hi) @ f(tf32_rna(g_lo))
    with_bf16 = hh + small_bf16.sum(1)
    e3 = np.max(np.abs(three_tf32 - exact) / exact)
    eb = np.max(np.abs(with_bf16 - exact) / exact)
    assert e3 <= 2.0 ** -21
    assert eb <= 2.0 ** -17
    assert np.sqrt(np.mean(((with_bf16 - exact) / exact) ** 2)) <= 5e-7      # typical error: far inside the bound
