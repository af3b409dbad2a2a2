import numpy as np

from oracle.gae_oracle import gae_delta_form, gae_targets, standardize


def _case(T=37, N=11, seed=0, cuts=True):
    rng = np.random.default_rng(seed)
    r = rng.normal(-1.5, 1.0, (T, N)).astype(np.float32)
    v = rng.normal(-30, 5.0, (T + 1, N)).astype(np.float32)
    cut = np.zeros(T, bool)
    cv = rng.normal(-30, 5.0, (T, N)).astype(np.float32)
    if cuts:
        cut[[9, 19, 29]] = True
    return r, v, cut, cv


def test_rllib_form_equals_textbook_gae():
    for cuts in (False, True):
        r, v, cut, cv = _case(cuts=cuts)
        for gamma, lam in ((0.99, 0.95), (0.95, 0.9), (1.0, 1.0), (0.995, 0.0)):
            a1, t1 = gae_targets(r, v, gamma, lam, cut, cv)
            a2, t2 = gae_delta_form(r, v, gamma, lam, cut, cv)
            np.testing.assert_allclose(a1, a2, rtol=2e-4, atol=2e-4)
            np.testing.assert_allclose(t1, t2, rtol=2e-5, atol=2e-4)


def test_termination_bootstraps_from_zero():
    r, v, cut, _ = _case()
    a, tg = gae_targets(r, v, 0.99, 0.95, cut, None)
    np.testing.assert_allclose(tg[9], r[9], rtol=1e-6)       # last step of an episode that terminated


def test_lambda_one_is_discounted_return():
    r, v, _, _ = _case(cuts=False)
    _, tg = gae_targets(r, v, 0.9, 1.0)
    ret = v[-1].astype(np.float64)
    for t in range(r.shape[0] - 1, -1, -1):
        ret = r[t] + 0.9 * ret
    np.testing.assert_allclose(tg[0], ret, rtol=1e-4)


def test_standardize():
    x = np.random.default_rng(1).normal(3, 2, 1000).astype(np.float32)
    y = standardize(x)
    assert abs(y.mean()) < 1e-5 and abs(y.std() - 1) < 1e-4
    assert np.all(standardize(np.full(8, 2.0, np.float32)) == 0)
