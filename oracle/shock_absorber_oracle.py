"""TEST INFRASTRUCTURE ONLY: numpy restatement of the exact log-density of the reference's shock-absorber example
(BASELINE.json configs[0]), needed at test time to drive the Metropolis-Hastings step that follows tt_irt1
(reference python/test_shock_absorber_tt.py:160-171) on a box where /root/reference does not exist.

  logf_prior    reference python/test_shock_absorber_tt.py:15-37   (Gaussian prior on beta with precision theta2, Gamma on theta2)
  logL_weibull  reference python/test_shock_absorber_tt.py:41-66   (Weibull accelerated-failure-time likelihood, censored data)
  quantile      reference python/test_shock_absorber_tt.py:176-179 (the quantity of interest of the example)
  draw_seeds / draw_uniforms  reference :146-148, :167 (np.random.random, legacy generator)

Pinned: tests/golden/shock_mh_D2.npz holds values of the REFERENCE's own functions (executed from the reference source by
tests/golden/make_golden_mh.py); tests/test_shock_mh.py compares this restatement with them entry by entry.
"""
import numpy as np

ALPHA, BETA = 6.8757, 2.2932          # :19-20


def logf_prior(theta, beta_mean, beta_var):
    """theta: (I, d+2) = (beta_0..beta_d, lambda-parameter theta2).  Same operation order as :30-36."""
    theta = np.asarray(theta, dtype=np.float64)
    d = theta.shape[1] - 2
    theta2 = theta[:, d + 1:d + 2]
    F = -((theta[:, :d + 1] - np.asarray(beta_mean)[None, :]) ** 2) * 0.5 * theta2 / np.asarray(beta_var)[None, :]   # :32
    F = np.sum(F, axis=1)[:, None]                                                                                   # :33
    with np.errstate(divide="ignore"):
        return F + (ALPHA - 0.5) * np.log(theta2) - BETA * theta2                                                    # :35


def logL_weibull(theta, x, y, c):
    """x: (d, m) covariates, y: (m,) times, c: (m,) censoring flags.  Same accumulation order over the data as :53-65."""
    theta = np.asarray(theta, dtype=np.float64)
    d = theta.shape[1] - 2
    I = theta.shape[0]
    beta = theta[:, :d + 1]
    lam = theta[:, d + 1:d + 2]
    x = np.reshape(np.asarray(x, dtype=np.float64), [d, np.size(y)], order="F")
    F = np.zeros((I, 1))
    beta0 = beta[:, 0:1]
    for i in range(np.size(y)):
        logeta = beta[:, 1:d + 1].dot(x[:, i])[:, None] + beta0                                                      # :54-56
        eta = np.exp(logeta)
        yeta = y[i] / eta
        with np.errstate(divide="ignore", invalid="ignore"):
            if c[i] == 1:                                                                                            # :59 censored
                f = -(yeta ** lam)
            else:
                f = np.log(lam) - logeta + (lam - 1.0) * (np.log(y[i]) - logeta) - (yeta ** lam)                     # :62
                f = f + np.log(30000.0)                                                                              # :63
        F = F + f
    return F


def log_posterior(Z, x, y, c, beta_mean, beta_var):
    """lPex of reference :161 as a flat vector."""
    return (logL_weibull(Z, x, y, c) + logf_prior(Z, beta_mean, beta_var)).ravel()


def quantile_of_interest(Z, d_cov, q=0.05):
    """reference :176-179."""
    theta1 = np.exp(Z[:, 0])
    theta2 = Z[:, d_cov + 1]
    return theta1 * ((-np.log(q)) ** (1.0 / theta2))


def draw_seeds(M, D, seed):
    """q exactly as the reference draws it (:146-148), legacy np.random seeded."""
    np.random.seed(int(seed))
    q = np.random.random([M, D])
    return np.asfortranarray(np.reshape(q, [M, D], order="F"))


def draw_uniforms(count, seed):
    """The uniforms the reference's MH loop draws one by one (:167)."""
    np.random.seed(int(seed))
    return np.random.random(count)
