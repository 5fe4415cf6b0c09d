"""The error band of precision "f16r" (k_fused_tc.cuh, k_sample_thr), checked on the CPU with a numpy restatement of the
arithmetic: tables scaled by powers of two, rounded to fp16 (round-to-nearest, subnormals kept), products exact, FP32
accumulation in two different orders, threshold and bias carried as fp16 pairs.  The GPU tests prove the end result
(bit-identical to the FP32 path); this one pins the bound the candidate pass relies on:

    eps = 1.25 [(2^-10 + (2.5 d + 8) 2^-22) ||u|| N_max + 2^-22 B_max + ((2.5 d + 8) 2^-22 + 2^-21) B_max
                + 2^-25 sqrt(d) (N_max / s_u + ||u|| / s_i)]          (the B_max terms of the second kind with a bias only)
"""
import numpy as np

TARGET, THR_SHIFT = 9, 12


def scale_exp(amax):
    """f16_scale_exp: e with amax * 2^e in [2^9, 2^10)"""
    if not (amax > 0 and np.isfinite(amax)):
        return 0
    ex = int(np.floor(np.log2(float(amax))))
    return int(np.clip(TARGET - ex, -60, 60))


def f16(x):
    return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)


def candidate_scores(U, I, b, order):
    """what the tensor core leaves in the accumulator, brought back to true units: sum_k fp16(u s_u) fp16(i s_i) (+ g beta),
    FP32 accumulation in the given k order, divided by s_u s_i"""
    ei = scale_exp(np.abs(I).max())
    si = 2.0 ** ei
    Ih = f16(I * si)
    out = np.zeros((U.shape[0], I.shape[0]), np.float64)
    beta_h = beta_l = None
    if b is not None:
        bmax = float(np.abs(b).max())
        m = 13 - int(np.floor(np.log2(bmax))) - ei if bmax > 0 else 0
        beta = b.astype(np.float64) * 2.0 ** (ei + m)
        beta_h = f16(beta)
        beta_l = f16(beta - beta_h)
    eus = []
    for r in range(U.shape[0]):
        eu = scale_exp(np.abs(U[r]).max())
        if b is not None:
            eu = min(eu, m + 15)
        eus.append(eu)
        uh = f16(U[r] * 2.0 ** eu)
        acc = np.zeros(I.shape[0], np.float32)
        if b is not None and eu - m >= -24:
            g = 2.0 ** (eu - m)
            acc = (acc + np.float32(g) * beta_h.astype(np.float32)).astype(np.float32)   # exact products, FP32 adds
            acc = (acc + np.float32(g) * beta_l.astype(np.float32)).astype(np.float32)
        for k in order:
            acc = (acc.astype(np.float64) + uh[k] * Ih[:, k]).astype(np.float32)          # product exact, one FP32 rounding per add
        out[r] = acc.astype(np.float64) / 2.0 ** (eu + ei)
    return out, np.array(eus), ei


def band(U, I, b, eus, ei):
    d = U.shape[1]
    un = np.linalg.norm(U.astype(np.float64), axis=1)
    nmax = np.linalg.norm(I.astype(np.float64), axis=1).max()
    bmax = float(np.abs(b).max()) if b is not None else 0.0
    eps = (2.0 ** -10 + (2.5 * d + 8) * 2.0 ** -22) * un * nmax + 2.0 ** -22 * bmax
    if b is not None:
        eps = eps + ((2.5 * d + 8) * 2.0 ** -22 + 2.0 ** -21) * bmax
        m = 13 - int(np.floor(np.log2(bmax))) - ei if bmax > 0 else 0
        eps = eps + np.where(eus - m < -24, bmax, 0.0)  # rows whose g = s_u 2^-m leaves fp16: the candidate pass drops their bias
    eps = eps + 2.0 ** -25 * np.sqrt(d) * (nmax / 2.0 ** eus + un / 2.0 ** ei)
    return 1.25 * eps


def _case(seed, d, su, si, bias_scale):
    g = np.random.default_rng(seed)
    U = (g.standard_normal((24, d)) * 0.1 * su).astype(np.float32)
    U[::3] *= 1e-4
    U[1, : d // 2] *= 1e-7                      # far below the row's largest element: fp16 subnormals after scaling
    I = (g.standard_normal((700, d)) * 0.1 * si).astype(np.float32)
    I[::11] *= 30.0
    I[5] *= 1e-9
    b = (g.standard_normal(700) * bias_scale * su * si).astype(np.float32) if bias_scale > 0 else None
    return U, I, b


def test_fp16_candidate_scores_stay_inside_the_band():
    for seed, (d, su, si, bs) in enumerate([(64, 1.0, 1.0, 0.0), (128, 1.0, 1.0, 0.05), (128, 1e-6, 1e-3, 0.0), (64, 3e3, 2e2, 1.0),
                                            (128, 1e-12, 1e6, 30.0), (32, 1.0, 1.0, 0.5), (100, 0.3, 7.0, 0.0)]):
        U, I, b = _case(seed, d, su, si, bs)
        exact = U.astype(np.float64) @ I.astype(np.float64).T + (0.0 if b is None else b.astype(np.float64))
        for order in (range(d), reversed(range(d)), np.random.default_rng(seed).permutation(d)):
            got, eus, ei = candidate_scores(U, I, b, list(order))
            eps = band(U, I, b, eus, ei)
            err = np.abs(got - exact).max(axis=1)
            assert np.all(err <= 0.5 * eps), (d, su, si, bs, float((err / eps).max()))
    # the band is needed: the candidate scores are NOT FP32-exact
    U, I, b = _case(0, 64, 1.0, 1.0, 0.0)
    got, _, _ = candidate_scores(U, I, None, list(range(64)))
    exact = U.astype(np.float64) @ I.astype(np.float64).T
    assert np.abs(got - exact).max() > 2.0 ** -20 * np.abs(exact).max()


def test_threshold_operand_is_exact_and_never_above_the_threshold():
    """T0 s_u s_i / C = hi + lo with fp16 hi (nearest) and lo (rounded down): hi + lo <= t, spans <= 22 bits, so the
    threshold the kernel really applies, (hi + lo) C / (s_u s_i), is an FP32 number."""
    g = np.random.default_rng(3)
    for t in np.concatenate([g.standard_normal(2000) * 3000.0, g.standard_normal(2000) * 0.01, [0.0, 31999.0, -31999.0, 2.0 ** -20]]):
        t = float(np.float32(t))
        hi = float(np.float16(t))
        rem = np.float32(t - hi)
        lo16 = np.float16(rem)
        if float(lo16) > float(rem):                     # round towards -inf on the fp16 grid
            lo16 = np.nextafter(lo16, np.float16(-np.inf))
        lo = float(lo16)
        s = hi + lo
        assert s <= t and float(np.float32(s)) == s      # collect a superset; exact in FP32
        assert t - s <= max(2.0 ** -24, abs(t) * 2.0 ** -20)


def test_rows_dwarfed_by_the_bias_and_rows_that_dwarf_it():
    """g = s_u 2^-m has to be an fp16 power of two: a row 2^-20 of the bias is scaled less (its largest element then sits far
    below 2^9 and the absolute 2^-25 term of the band carries it), a row 2^35 times the bias loses the bias in the candidate
    pass (B_max joins its band).  Both stay inside the band."""
    g = np.random.default_rng(11)
    d = 64
    I = (g.standard_normal((500, d)) * 0.1).astype(np.float32)
    b = (g.standard_normal(500) * 0.5).astype(np.float32)
    U = (g.standard_normal((6, d)) * 0.1).astype(np.float32)
    U[0] *= 2.0 ** -22      # all bias
    U[1] *= 2.0 ** 38       # no bias to speak of
    exact = U.astype(np.float64) @ I.astype(np.float64).T + b.astype(np.float64)
    got, eus, ei = candidate_scores(U, I, b, list(range(d)))
    bmax = float(np.abs(b).max())
    m = 13 - int(np.floor(np.log2(bmax))) - ei
    assert eus[0] == m + 15 and eus[0] < scale_exp(np.abs(U[0]).max())   # scaled less than its own magnitude asks for
    assert eus[1] - m < -24                                                # g underflows: bias dropped
    eps = band(U, I, b, eus, ei)
    err = np.abs(got - exact).max(axis=1)
    assert np.all(err <= 0.9 * eps), (err / eps).tolist()
    assert err[1] > 0.5 * bmax                                             # the bias really is missing from row 1's candidate scores
