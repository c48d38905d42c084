"""CPU: the several-lanes-per-stream formulation of the adaptive predictor (alac/net_b200/csrc/k2_lpc.cuh,
lpc_lanes<L, T>, L = 4 or 8) restated in Python and checked against the independent model's PredictorDecompressFirAdapt.

What it pins, without a GPU: the reference's early-exit loop over the taps (AlacFile.cs:322-331) as a
suffix scan over unconditionally computed, clamped steps; taps split j = r*T + t over the lanes;
surplus taps (j >= order) with weight 0 and an unreachable threshold; warm-up samples reading a base that
is not there yet (modelled as garbage) without touching the coefficients; 32-bit wraparound everywhere."""
import random

import pytest

from pymodel import alac_model as M

U32 = 0xFFFFFFFF
CLAMP = 1 << 25


def _i32(v):
    return M.i32(v)


def quad_predict(e, n, rss, coef, order, q, T, rng, L=4):
    """L lanes per stream (the kernel uses L = 4); lane r owns taps j = r*T + t."""
    c = [[0] * T for _ in range(L)]
    H = [[0] * T for _ in range(L)]
    wgt = [[0] * T for _ in range(L)]
    thr = [[0] * T for _ in range(L)]
    for r in range(L):
        for t in range(T):
            j = r * T + t
            valid = j < order
            c[r][t] = coef[j] if valid else 0
            wgt[r][t] = order - j if valid else 0
            thr[r][t] = 0 if valid else 0x7FFFFFFF
    rnd = _i32(1 << ((q - 1) & 31))
    rneg = (1 << q) - 1
    o = [0] * n
    ring = [rng.randint(-(1 << 31), (1 << 31) - 1) for _ in range(32)]      # uninitialised shared memory
    prev = e[0]
    H[0][0] = prev
    ring[0] = prev
    o[0] = prev
    for i in range(1, n):
        ee = e[i]
        base = ring[(i - 1 - order) & 31]
        main = i > order
        nsg = 1 if ee < 0 else -1
        sgbase = _i32(-base) if ee < 0 else base
        E0 = (_i32(-ee) if ee < 0 else ee) if main else 0
        rr = rneg if ee < 0 else 0
        acc = [0] * L
        dp = [[0] * T for _ in range(L)]
        st = [[0] * T for _ in range(L)]
        mine = [0] * L
        for r in range(L):
            for t in range(T - 1, -1, -1):
                dp[r][t] = _i32(H[r][t] * nsg + sgbase)
                acc[r] = (acc[r] + c[r][t] * dp[r][t]) & U32
                mag = (abs(dp[r][t]) + rr) & U32
                st[r][t] = min(((mag >> q) * wgt[r][t]) & U32, CLAMP)        # unsigned min
                mine[r] = _i32(mine[r] + st[r][t])
        # inclusive suffix sums over the lanes (shuffle-down scan: log2(L) stages), minus the lane's own total
        suf = list(mine)
        d = 1
        while d < L:
            suf = [_i32(suf[r] + (suf[r + d] if r + d < L else 0)) for r in range(L)]
            d *= 2
        for r in range(L):
            rem = _i32(E0 - _i32(suf[r] - mine[r])) if main else -1
            for t in range(T - 1, -1, -1):
                sg = max(min(dp[r][t], 1), -1)
                if rem > thr[r][t]:
                    c[r][t] = _i32(c[r][t] - sg)
                rem = _i32(rem - st[r][t])
        s = _i32((sum(acc) & U32) * nsg)
        v = M.sar(_i32(rnd + s), q)
        v = _i32(_i32(v + base) + ee)
        w = _i32(prev + ee)
        oo = M.sext(v if main else w, rss)
        below = [H[r - 1][T - 1] if r > 0 else 0 for r in range(L)]
        for r in range(L):
            for t in range(T - 1, 0, -1):
                H[r][t] = H[r][t - 1]
            H[r][0] = oo if r == 0 else below[r]
        ring[i & 31] = oo
        prev = oo
        o[i] = oo
    return o


def _taps_per_lane(order):          # lpc_role's choice of the lpc_lanes<4, T> instantiation
    for limit, T in ((8, 2), (12, 3), (16, 4), (20, 5), (24, 6), (28, 7)):
        if order <= limit:
            return T
    return 8


@pytest.mark.parametrize("seed", range(4))
def test_quad_formulation_equals_the_reference_loop(seed):
    rng = random.Random(seed)
    for _ in range(60):
        order = rng.randint(1, 30)
        q = rng.randint(0, 15)
        rss = rng.choice([16, 17, 24, 25])
        n = rng.randint(order + 2, 160)
        # a warp's instantiation is chosen by its LARGEST order: smaller orders also run with more taps per lane
        T = rng.choice([t for t in range(_taps_per_lane(order), 9)])
        lim = 1 << (rss - 2)
        kind = rng.random()
        e = [rng.randint(-lim, lim) if rng.random() < (0.3 if kind < 0.7 else 0.9) else rng.randint(-50, 50) for _ in range(n)]
        coef = [rng.randint(-32768, 32767) if kind > 0.85 else rng.randint(-2000, 2000) for _ in range(order)]
        want = M.predict(list(e), n, rss, list(coef), order, q)
        got = quad_predict(e, n, rss, list(coef), order, q, T, rng)
        assert got == want, (order, q, rss, T)


@pytest.mark.parametrize("lanes", [2, 8])
def test_other_lane_counts_use_the_same_scan(lanes):
    """eight lanes per stream (lpc_lanes<8, T>: tiny mono batches, T = ceil(order / 8) taps per lane) and two (not in
    the kernels) are the same formulation"""
    rng = random.Random(100 + lanes)
    for _ in range(40):
        order = rng.randint(1, 30)
        q = rng.randint(0, 15)
        rss = rng.choice([16, 17, 24, 25])
        n = rng.randint(order + 2, 120)
        T = (order + lanes - 1) // lanes + rng.randint(0, 1)
        lim = 1 << (rss - 2)
        e = [rng.randint(-lim, lim) if rng.random() < 0.4 else rng.randint(-50, 50) for _ in range(n)]
        coef = [rng.randint(-2000, 2000) for _ in range(order)]
        want = M.predict(list(e), n, rss, list(coef), order, q)
        assert quad_predict(e, n, rss, list(coef), order, q, T, rng, L=lanes) == want, (order, q, rss, T, lanes)
