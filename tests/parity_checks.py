"""Comparison helpers shared by the golden (CPU) and live (GPU) parity tests, with the tolerances written down.

who = "gpu": the library's EXACT arithmetic (the default) against the reference's kernels or golden files made by them on a
B200 -- equality, no tolerance. who = "gpu_fast": the fast arithmetic, who = "cpu": the plain-C oracle; for those two:

Why not bit-exact everywhere: the path is fp32 with --use_fast_math in the reference build (ex2/rcp/rsqrt/sin/cos
approximations) and goes through the texture unit's 8-bit-weight bilinear filter, so a one-ulp change in a warped
coordinate can move an interpolated intensity by gradient/256 and an NCC cost by ~1e-3. Decisions (arg-min, accept
tests) then flip for near-ties and trajectories diverge ("stochastic parity"). What IS exact:
the XORWOW streams (integer), the number of draws per pixel per stage (control flow), the pixels a launch may touch,
and everything produced by pure RNG paths on the same hardware.

Tolerances (T_*): measured on B200 against the reference build and then fixed with margin; the measured values are
recorded in DESIGN.md section "Parity".
"""
import numpy as np

T_COST = 2e-3            # |delta NCC cost| per sample
T_COST_FRAC = {"plane3": 0.995, "dtu5": 0.995, "room6": 0.93}     # fraction of samples within T_COST (room6: weak texture)
T_GEOM = 1e-3            # |delta geometric-consistency cost| (pixels), fraction >= 0.999
T_SAME_PLANE = {"gpu": {"plane3": 1.0, "dtu5": 1.0, "room6": 1.0},        # exact arithmetic vs reference on the same GPU: bit-identical
                "gpu_fast": {"plane3": 0.93, "dtu5": 0.93, "room6": 0.78},   # fast arithmetic vs reference on the same GPU
                "cpu": {"plane3": 0.25, "dtu5": 0.25, "room6": 0.25}}   # CPU oracle (libm, no fast-math): bits differ, values close
T_CLOSE_PLANE = {"plane3": 0.85, "dtu5": 0.9, "room6": 0.75}           # planes equal to 1e-4 relative among updated pixels


def colour_mask(h, w, red):
    yy, xx = np.mgrid[0:h, 0:w]
    return ((xx + yy) & 1) == red


def frac_within(a, b, tol):
    return float((np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) <= tol).mean())


def check_cost_map(name, got, want, frac=None, exact=False):
    if exact:
        np.testing.assert_array_equal(got, want)
        return
    f = frac_within(got, want, T_COST)
    assert f >= (frac if frac is not None else T_COST_FRAC[name]), (name, f)
    assert np.abs(got - want).mean() < 1e-3, (name, float(np.abs(got - want).mean()))


def check_state(name, got, want, who, upd=None, rng_digest=None, planes_exact=False):
    """got/want: dicts planes, costs, views, rng (full state or digest). upd: mask of pixels the stage may have changed."""
    h, w = want["costs"].shape
    m = np.ones((h, w), bool) if upd is None else upd
    if upd is not None:      # pixels of the other colour must be bit-untouched
        np.testing.assert_array_equal(got["planes"][~upd], want["planes"][~upd])
        np.testing.assert_array_equal(got["costs"][~upd], want["costs"][~upd])
    g_rng = rng_digest(got["rng"]) if (rng_digest and got["rng"].ndim == 3) else got["rng"]
    w_rng = rng_digest(want["rng"]) if (rng_digest and want["rng"].ndim == 3) else want["rng"]
    np.testing.assert_array_equal(g_rng, w_rng)          # identical draw counts everywhere
    same = np.all(got["planes"] == want["planes"], -1)
    close = np.all(np.abs(got["planes"] - want["planes"]) <= 1e-4 * (1 + np.abs(want["planes"])), -1)
    if who == "gpu":          # the exact arithmetic: every plane, cost and view mask is the reference's
        assert same[m].all(), (name, who, float(same[m].mean()))
        np.testing.assert_array_equal(got["costs"][m], want["costs"][m])
        np.testing.assert_array_equal(got["views"][m], want["views"][m])
        return 1.0, 1.0
    who = "gpu" if who == "gpu_fast" else who
    if planes_exact:
        assert close[m].mean() >= 0.9999, (name, who, float(close[m].mean()))
        if who == "gpu":
            assert same[m].mean() >= 0.999, (name, who, float(same[m].mean()))
    else:
        assert same[m].mean() >= T_SAME_PLANE["gpu_fast" if who == "gpu" else who][name], (name, who, float(same[m].mean()))
        assert close[m].mean() >= T_CLOSE_PLANE[name], (name, who, float(close[m].mean()))
    sel = m & close
    assert frac_within(got["costs"][sel], want["costs"][sel], 5e-3) >= 0.9, name
    assert (got["views"] == want["views"])[sel].mean() >= 0.97, name
    return float(same[m].mean()), float(close[m].mean())
