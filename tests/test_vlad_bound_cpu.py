"""CPU property test of the coarse-score error bound the v3 VLAD assignment kernel relies on
(anyloc_b200/csrc/vlad_tc.cu, tile epilogue): the tensor core multiplies tf32-TRUNCATED features with tf32-ROUNDED
centres; a row's label is taken from the coarse scores alone only if every other score is more than 2*eps below the
maximum, with

    eps = 1.01 * (|d| * max|c^| + |x| * (max|e| + 1e-4 * max|c^|)),   d = x - trunc_tf32(x),  e = c^ - rna_tf32(c^).

This test restates that arithmetic with numpy (bit masks for tf32, fp32 accumulation in a truncating and in the
natural order) and checks, on random, adversarial (heavy-tailed, tiny, huge, sign-structured) inputs, that
|coarse - exact_fp64| <= eps always holds -- i.e. that "one candidate" really implies "that candidate is the exact
argmax" -- and that the bound is not vacuous (typically well below the a-priori 2^-9 |x||c^| of the v2 pipeline)."""
import numpy as np
import pytest


def trunc_tf32(a):
    return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rna_tf32(a):
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x1000)) & np.uint64(0xFFFFE000)          # cvt.rna.tf32.f32: round half away (magnitude)
    return u.astype(np.uint32).view(np.float32)


def coarse_scores(x, chat_t, order="natural"):
    """fp32 accumulation of tf32 x tf32 products (exact in fp32: 11 x 11 significant bits), k-steps of 8 like the UMMA;
    'trunc' emulates a round-toward-zero accumulator."""
    xt = trunc_tf32(x)
    D = x.shape[1]
    acc = np.zeros((x.shape[0], chat_t.shape[0]), np.float32)
    for k0 in range(0, D, 8):
        part = (xt[:, None, k0:k0 + 8].astype(np.float64) * chat_t[None, :, k0:k0 + 8].astype(np.float64)).sum(-1)
        s = acc.astype(np.float64) + part
        if order == "trunc":          # truncate the running sum to fp32 toward zero
            f = s.astype(np.float32)
            over = np.abs(f.astype(np.float64)) > np.abs(s)
            f = np.where(over, np.nextafter(f, np.float32(0)), f)
            acc = f.astype(np.float32)
        else:
            acc = s.astype(np.float32)
    return acc


def eps_bound(x, chat, chat_t):
    d = x - trunc_tf32(x)
    dn = np.sqrt((d.astype(np.float32) ** 2).sum(1, dtype=np.float32))
    xn = np.sqrt((x.astype(np.float32) ** 2).sum(1, dtype=np.float32))
    cmax = np.sqrt((chat.astype(np.float64) ** 2).sum(1)).max()
    dcmax = np.sqrt(((chat - chat_t).astype(np.float64) ** 2).sum(1)).max()
    return (1.01 * (dn * cmax + xn * (dcmax + 1e-4 * cmax))).astype(np.float64), xn, cmax


CASES = {
    "unit_rows": lambda g, n, D: g.standard_normal((n, D)) / np.sqrt(D),
    "heavy_tail": lambda g, n, D: g.standard_cauchy((n, D)) * 1e-2,
    "tiny": lambda g, n, D: g.standard_normal((n, D)) * 1e-20,
    "huge": lambda g, n, D: g.standard_normal((n, D)) * 1e15,
    "all_just_below_tf32_ulp": lambda g, n, D: (1.0 + (2.0 ** -10) * (1 - 2.0 ** -13)) * np.sign(g.standard_normal((n, D))),
    "sparse": lambda g, n, D: g.standard_normal((n, D)) * (g.random((n, D)) < 0.05),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("D,K", [(384, 8), (1536, 32), (2048, 128)])
@pytest.mark.parametrize("order", ["natural", "trunc"])
def test_coarse_bound_holds(name, D, K, order):
    g = np.random.default_rng(hash((name, D, K)) % (2 ** 32))
    x = CASES[name](g, 48, D).astype(np.float32)
    c = (g.standard_normal((K, D)) * g.uniform(0.2, 3.0, (K, 1))).astype(np.float32)
    if name == "all_just_below_tf32_ulp":      # worst case for Cauchy-Schwarz: centres parallel to the dropped part
        c[0] = np.sign(x[0])
    chat = (c / (np.sqrt((c.astype(np.float64) ** 2).sum(1, keepdims=True)) + 1e-8)).astype(np.float32)   # fpk cos_sim
    chat_t = rna_tf32(chat)
    exact = x.astype(np.float64) @ chat.astype(np.float64).T
    coarse = coarse_scores(x, chat_t, order).astype(np.float64)
    eps, xn, cmax = eps_bound(x, chat, chat_t)
    err = np.abs(coarse - exact).max(1)
    assert np.all(err <= eps), (name, float((err / np.maximum(eps, 1e-300)).max()))
    # single-candidate rows: the coarse argmax is the exact argmax
    smax = coarse.max(1)
    cand = coarse >= (smax - 2 * eps)[:, None]
    single = cand.sum(1) == 1
    assert np.array_equal(coarse.argmax(1)[single], exact.argmax(1)[single])
    if name == "unit_rows":                     # not vacuous: well below v2's a-priori 2^-9 |x||c^|
        assert np.median(eps / (2.0 ** -9 * xn * cmax)) < 0.5
        assert single.mean() > 0.3 or K >= 128


def test_tf32_helpers():
    a = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -10, -3.1415927, 0.0, 65504.0], np.float32)
    t, r = trunc_tf32(a), rna_tf32(a)
    assert np.all(np.abs(t) <= np.abs(a)) and np.all((t.view(np.uint32) & 0x1FFF) == 0)
    assert np.all(np.abs(r - a) <= np.abs(a) * 2.0 ** -11 + 1e-45) and np.all((r.view(np.uint32) & 0x1FFF) == 0)
    assert r[1] == np.float32(1.0 + 2.0 ** -10)          # tie rounds away from zero (cvt.rna)
