"""CPU model of the two rigorous margins of the tensor-core exact scan (leann_rs_b200/csrc/exact_scan_tc.cu). No product code
runs here: numpy restates (i) the bf16 rounding of both operands, (ii) the squared-L2 augmentation (|x|^2/2 as three exact
bf16 pieces scaled by a power of two), (iii) the candidate threshold `tc_l2_threshold` / `tc_eps` and (iv) the re-rank's
approximate cut, with the kernels' constants, and checks on adversarial inputs that no row of the exact f32 top-k can be
lost. The GPU tests (tests/test_exact_gpu.py) check the same property end to end; this file documents WHY it holds and keeps
the constants honest when someone edits them (update both places together)."""
import numpy as np

f32 = np.float32


def bf16(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16            # round to nearest even
    return r.astype(np.uint32).view(np.float32)


def prep(x):
    ss = (x.astype(f32) ** 2).sum(1, dtype=f32)
    xb = bf16(x)
    res = (np.sqrt(((xb - x).astype(f32) ** 2).sum(1, dtype=f32)) * f32(1.0001)).astype(f32)   # |x^ - x|, rounded up
    nrm = (np.sqrt(ss) * f32(1.0001)).astype(f32)                                              # |x|, rounded up
    return xb, ss, nrm, res


def l2_scale(xmax):
    e = np.array([xmax], dtype=f32).view(np.uint32)[0] & 0x7F800000
    return f32(1) if e in (0, 0x7F800000) else np.array([e], dtype=np.uint32).view(f32)[0]


def model_l2(x, q, k):
    """Returns (must-survive mask, survivor mask, tensor-core scores, eps) of one round with threshold = exact k-th best."""
    n, d = x.shape
    d4 = (d + 3) // 4
    dp8 = (d4 * 4 + 4 + 7) // 8 * 8
    xb, xss, xn, xr = prep(x)
    qb, qss, qn, qr = prep(q)
    xmax, xres = xn.max(), xr.max()
    c = l2_scale(xmax)
    h = (f32(0.5) * xss / c).astype(f32)
    hi = bf16(h); r1 = (h - hi).astype(f32); mid = bf16(r1); r2 = (r1 - mid).astype(f32); lo = bf16(r2)
    assert np.all(lo == r2), "third piece must be exact"
    assert np.all(hi.astype(np.float64) + mid + lo == h.astype(np.float64)), "pieces must sum to h exactly"
    xa = np.concatenate([xb, -hi[:, None], -mid[:, None], -lo[:, None]], 1).astype(np.float64)
    qa = np.concatenate([qb, np.full((len(q), 3), c, dtype=f32)], 1).astype(np.float64)
    sp = (qa @ xa.T).astype(f32)                                   # f32 accumulate modelled by f64 + one rounding
    D = ((q[:, None, :].astype(f32) - x[None, :, :].astype(f32)) ** 2).sum(2, dtype=f32)
    dk = np.sort(D, axis=1)[:, k - 1]
    # tc_l2_threshold
    hmax = f32(0.5) * xmax * xmax * f32(1.000001); hc = hmax / c
    qhat = qn * f32(1.00390625); xhat = xmax * f32(1.00390625)
    qaN = np.sqrt(qhat * qhat + f32(3) * c * c) * f32(1.000001)
    xaN = np.sqrt(xhat * xhat + f32(1.01) * hc * hc) * f32(1.000001)
    acc = f32(dp8) * f32(2.3841858e-7) * qaN * xaN
    eps = (qr * xmax + qhat * xres + acc) * f32(1.0001) + f32(5e-5) * hmax
    L = qn * qn * f32(0.99975)
    T = f32(0.5) * (L - dk * f32(1.0001)) - eps
    T = T - np.abs(T) * f32(1e-6) - (L + dk) * f32(1e-6)
    return D <= dk[:, None], sp > T[:, None], sp, eps, D, (qn, xmax)


def lowdim(n, d, seed):
    r = np.random.default_rng(seed)
    W = r.standard_normal((32, d), dtype=f32)
    return (r.standard_normal((n, 32), dtype=f32) @ W + 0.3 * r.standard_normal((n, d), dtype=f32)).astype(f32)


def datasets():
    rng = np.random.default_rng(0)
    x, q = lowdim(6000, 96, 1), lowdim(24, 96, 2)
    yield "raw", x, q, 10
    yield "normalised", x / np.linalg.norm(x, axis=1, keepdims=True), q / np.linalg.norm(q, axis=1, keepdims=True), 10
    xs = (np.abs(rng.standard_normal((6000, 128))) * 50).astype(f32); qs = (np.abs(rng.standard_normal((24, 128))) * 50).astype(f32)
    yield "all-positive (SIFT-like)", xs, qs, 10
    yield "tiny unit", (xs * 1e-3).astype(f32), (qs * 1e-3).astype(f32), 10
    yield "varying norms", (x * np.linspace(0.1, 30, len(x), dtype=f32)[:, None]).astype(f32), q, 20
    cc = f32(1 + 2.0 ** -8)                                         # bf16 tie point: both operands lose 2^-8 at once
    xt = (rng.integers(-8, 9, size=(6000, 128)) / 64.0).astype(f32); xt[:5] = f32(1.002); xt[[700, 2345, 5999]] = cc
    yield "tie points", xt, np.full((3, 128), cc, dtype=f32), 3


def test_l2_threshold_never_drops_a_top_k_row():
    worst = 0.0
    for name, x, q, k in datasets():
        must, surv, sp, eps, D, _ = model_l2(x, q, k)
        assert not (must & ~surv).any(), name
        sstar = 0.5 * ((q.astype(np.float64) ** 2).sum(1)[:, None] - ((q[:, None, :].astype(np.float64) - x[None, :, :]) ** 2).sum(2))
        ratio = float((np.abs(sp - sstar) / eps[:, None]).max())
        assert ratio < 1.0, (name, ratio)                           # the bound holds ...
        worst = max(worst, ratio)
        assert surv.sum(1).mean() < 0.12 * len(x), (name, surv.sum(1).mean())   # ... and still filters
    assert worst > 0.9                                              # ... and is tight: the tie-point case uses > 90 % of it


def test_rerank_cut_keeps_every_row_that_can_enter_the_top_k():
    """rerank_kernel: with a = the k-th largest tensor-core score among a round's survivors, rows below a - 2 (eps + delta) are
    dropped. No row of the exact top-k of the survivors may be among them (L2 form; delta = 5e-5 (|q| + max|x|)^2)."""
    for name, x, q, k in datasets():
        must, surv, sp, eps, D, (qn, xmax) = model_l2(x, q, k)
        for i in range(len(q)):
            idx = np.nonzero(surv[i])[0]
            if len(idx) <= 2 * k:
                continue
            a = np.sort(sp[i, idx])[::-1][k - 1]
            slack = f32(2) * (eps[i] * f32(1.001) + f32(5e-5) * (qn[i] + xmax) ** 2) * f32(1.001)
            kept = idx[sp[i, idx] >= a - slack]
            topk = idx[np.argsort(D[i, idx], kind="stable")[:k]]    # exact f32 top-k of the survivors, ties by row
            assert set(topk.tolist()) <= set(kept.tolist()), (name, i)
            assert len(kept) >= k


def test_dot_cut_on_near_ties():
    """Dot form of the same cut on scores packed closer than the bf16 error: everything within 2 eps of the k-th stays."""
    rng = np.random.default_rng(3)
    d, n, k = 128, 4000, 10
    qv = rng.standard_normal(d).astype(f32); qv /= np.linalg.norm(qv)
    x = (qv[None, :] * (1 - 1e-4 * rng.random((n, 1))) + 1e-3 * rng.standard_normal((n, d))).astype(f32)   # scores within ~1e-3
    xb, _, xn, xr = prep(x); qb, _, qn, qr = prep(qv[None, :])
    sp = (qb.astype(np.float64) @ xb.astype(np.float64).T).astype(f32)[0]
    s = (x.astype(f32) @ qv.astype(f32)).astype(f32)
    qhat = qn[0] * f32(1.00390625)
    eps = (qr[0] * xn.max() + qhat * xr.max() + f32(128) * f32(2.3841858e-7) * qhat * xn.max() * f32(1.00390625)) * f32(1.0001)
    assert np.abs(sp - s).max() <= eps
    a = np.sort(sp)[::-1][k - 1]
    kept = np.nonzero(sp >= a - f32(2) * (eps + f32(5e-5) * qn[0] * xn.max()) * f32(1.001))[0]
    topk = np.argsort(-s, kind="stable")[:k]
    assert set(topk.tolist()) <= set(kept.tolist())
