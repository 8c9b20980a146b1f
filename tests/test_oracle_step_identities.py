"""CPU checks (numpy, fp64 / fp16) of the algebra the one-pass step kernels rely on -- the identities are what
csrc/head_step2.cuh and csrc/step_pair.cu compute with, so they are pinned here where no GPU is needed:

* a heatmap split in two halves, each summed relative to its OWN maximum, merges into the softmax statistics of the
  whole (online-softmax merge) and the second moments about each half's own mean merge by the parallel-variance formula
  (Chan et al.) -- step_pair.cu, one exchange per heatmap;
* the same formula over many parts (the 32 warps of a CTA);
* JS: log2(1 + (G + 2 eps) / P) = log2(2 M) - log2 P with M = (P + G)/2 + eps -- the backward of the Gaussian window
  reuses the forward's log2 M (head_step2.cuh);
* fp16 of e * 2^15 keeps 11 significant bits for e >= 2e-9 -- the bf16 kernels' stash of e between the two sweeps.
"""

import numpy as np

EPS = 1e-24


def grid(n):
    return (2.0 * np.arange(n) + 1.0) / n - 1.0          # src/dsnt/nn.py:30-37


def direct_stats(z):
    h, w = z.shape
    e = np.exp2((z - z.max()) * np.log2(np.e))
    s = e.sum()
    p = e / s
    xs, ys = grid(w), grid(h)
    mux, muy = (p.sum(0) * xs).sum(), (p.sum(1) * ys).sum()
    vx = (p.sum(0) * (xs - mux) ** 2).sum()
    vy = (p.sum(1) * (ys - muy) ** 2).sum()
    return z.max(), s, mux, muy, vx, vy


def part_stats(z, rows, xs, ys):
    """What one CTA of the pair (or one warp) holds: sums relative to its own maximum, moments about its own mean."""
    m = z.max()
    e = np.exp2((z - m) * np.log2(np.e))
    s = e.sum()
    sx, sy = (e.sum(0) * xs).sum(), (e.sum(1) * ys[rows]).sum()
    ax = (e.sum(0) * (xs - sx / s) ** 2).sum()
    ay = (e.sum(1) * (ys[rows] - sy / s) ** 2).sum()
    return m, s, sx, sy, ax, ay


def merge(parts):
    m = max(p[0] for p in parts)
    sc = [np.exp2((p[0] - m) * np.log2(np.e)) for p in parts]
    s = sum(p[1] * c for p, c in zip(parts, sc))
    sx = sum(p[2] * c for p, c in zip(parts, sc))
    sy = sum(p[3] * c for p, c in zip(parts, sc))
    mux, muy = sx / s, sy / s
    m2x = sum(c * (p[4] + p[1] * (p[2] / p[1] - mux) ** 2) for p, c in zip(parts, sc))
    m2y = sum(c * (p[5] + p[1] * (p[3] / p[1] - muy) ** 2) for p, c in zip(parts, sc))
    return m, s, mux, muy, m2x / s, m2y / s


def test_two_halves_merge_like_an_online_softmax_and_chan_variance():
    rng = np.random.default_rng(0)
    for scale, spike in ((1.0, 0.0), (5.0, 0.0), (1.0, 12.0), (1.0, -30.0)):
        z = rng.standard_normal((64, 48)) * scale
        z[40:44, 10:14] += spike                              # mass (or a hole) in the lower half only
        xs, ys = grid(48), grid(64)
        halves = [part_stats(z[:32], slice(0, 32), xs, ys), part_stats(z[32:], slice(32, 64), xs, ys)]
        got, want = merge(halves), direct_stats(z)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-14), (scale, spike, got, want)
        # the two-part form of the cross term used by step_pair.cu: (mu_0 - mu_1)^2 S_0 S_1 / S
        (m0, s0, sx0, _, ax0, _), (m1, s1, sx1, _, ax1, _) = halves
        m = max(m0, m1)
        c0, c1 = np.exp(m0 - m), np.exp(m1 - m)
        s0, s1 = s0 * c0, s1 * c1
        vx = (ax0 * c0 + ax1 * c1 + (sx0 * c0 / s0 - sx1 * c1 / s1) ** 2 * s0 * s1 / (s0 + s1)) / (s0 + s1)
        assert abs(vx - want[4]) < 1e-12 * max(want[4], 1e-30) + 1e-15


def test_many_parts_merge_by_the_same_formula():
    rng = np.random.default_rng(1)
    z = rng.standard_normal((64, 64)) * 3.0
    xs, ys = grid(64), grid(64)
    parts = [part_stats(z[r:r + 2], slice(r, r + 2), xs, ys) for r in range(0, 64, 2)]       # 32 "warps"
    assert np.allclose(merge(parts), direct_stats(z), rtol=1e-12, atol=1e-14)


def test_js_window_backward_reuses_log2_m():
    rng = np.random.default_rng(2)
    p = rng.random(1000) ** 8 + 1e-12                    # P >> eps: the kernels drop the eps next to P
    g = rng.random(1000) ** 8
    m = 0.5 * (p + g) + EPS
    lhs = np.log2(1.0 + (g + 2 * EPS) / p)
    rhs = (np.log2(m) + 1.0) - np.log2(p)
    assert np.allclose(lhs, rhs, rtol=1e-12, atol=1e-12)
    # and the gradient term of SURVEY.md Appendix A.2: r = 1/2 [ln(P+eps) - ln(M+eps) + P/(P+eps) - M/(M+eps)] ~ 1/2 ln 2 - 1/2 ln2 log2(1 + G/P)
    r = 0.5 * (np.log(p) - np.log(0.5 * (p + g) + EPS))
    assert np.allclose(r, 0.5 * np.log(2.0) - 0.5 * np.log(2.0) * lhs, rtol=1e-9, atol=1e-9)


def test_fp16_stash_of_e_keeps_eleven_bits():
    rng = np.random.default_rng(3)
    t = -rng.random(200000) * 28.9                       # e = 2^t down to 2e-9
    e = np.exp2(t)
    stash = (e * 2.0 ** 15).astype(np.float16).astype(np.float64) * 2.0 ** -15
    rel = np.abs(stash - e) / e
    assert rel.max() <= 2.0 ** -11 * 1.0001
    assert (np.exp2(np.array([-30.5, -35.0, -39.0])) * 2.0 ** 15).astype(np.float16).min() > 0      # below that: gradual
