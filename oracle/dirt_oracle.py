"""TEST INFRASTRUCTURE ONLY: numpy restatement of the DIRT sampler loop, /root/reference/matlab/samplers/tt_dirt_sample.m:1-82
(spline interpolation, TT cross variants: the branch that calls tt_irt_sqr at :46 and :71), on top of
oracle/tt_irt_sqr_oracle.py.  Pinned against the reference's own tt_dirt_sample.m / tt_dirt_inverse.m executed from source by
oracle/mlite.py on a synthetic two-level DIRT (uniform and truncated-normal reference; tests/golden/matlab_dirt_*.npz,
tests/test_matlab_pins.py) and by the properties in tests/test_dirt.py.  Nothing under tt-irt_b200/ imports this file.
"""
import math
import re

import numpy as np
from scipy.special import erf, erfinv

from .tt_irt_sqr_oracle import tt_irt_sqr_oracle, tt_rt_sqr_oracle


def parse_reference(reference):
    """tt_dirt_sample.m:21-30: None for a uniform reference, else the half-width sigma of the truncated normal
    ('Normal' -> 4, 'Normal S' -> S: digits and dots of the string)."""
    if str(reference)[0].lower() == "u":
        return None
    digits = "".join(ch for ch in str(reference) if ch == "." or ch.isdigit())
    try:
        sigma = float(digits)
    except ValueError:
        sigma = float("nan")
    return 4.0 if math.isnan(sigma) else sigma


def tt_dirt_sample_oracle(levels, q, reference="uni"):
    """[z, lFapp] = tt_dirt_sample(IRTstruct, q).  levels[0] = (n, x0, ranks, cores) of IRTstruct.F0 on IRTstruct.x0,
    levels[j] (j = 1..nlvl) = (n, x, ranks, cores) of IRTstruct.F{j} on IRTstruct.x."""
    sigma = parse_reference(reference)
    z = np.array(q, dtype=np.float64, order="F", copy=True)
    lF = np.zeros(z.shape[0])
    if sigma is not None:
        cdf_factor = 0.5 / erf(sigma / math.sqrt(2.0))                      # :30
    for j in range(len(levels) - 1, 0, -1):                                  # :34
        if sigma is not None:
            z = erf(z / math.sqrt(2.0)) * cdf_factor + 0.5                  # :36
        n, xs, rk, c = levels[j]
        z, dl = tt_irt_sqr_oracle(n, xs, rk, c, z)                           # :46
        lF = lF + dl                                                         # :51
        if sigma is not None:
            lF = lF + np.sum(z ** 2, axis=1) / 2 - math.log(2 * cdf_factor ** 2 / math.pi) * z.shape[1] / 2   # :54
    if sigma is not None:
        z = erf(z / math.sqrt(2.0)) * cdf_factor + 0.5                      # :60
    n, xs, rk, c = levels[0]
    z, dl = tt_irt_sqr_oracle(n, xs, rk, c, z)                               # :71
    lF = lF + dl                                                             # :73
    return z, lF


def tt_dirt_inverse_oracle(levels, x, reference="uni"):
    """[q, lFapp] = tt_dirt_inverse(IRTstruct, x), /root/reference/matlab/samplers/tt_dirt_inverse.m:1-60: the levels walked
    from level 0 upwards with the forward transform tt_rt_sqr."""
    sigma = parse_reference(reference)
    if sigma is not None:
        cdf_factor = 0.5 / erf(sigma / math.sqrt(2.0))                      # :33
    n, xs, rk, c = levels[0]
    q, dl = tt_rt_sqr_oracle(n, xs, rk, c, np.asarray(x, dtype=np.float64))  # :39
    if sigma is not None:
        q = erfinv((q - 0.5) / cdf_factor) * math.sqrt(2.0)                 # :42
    lF = np.zeros(q.shape[0]) + dl                                           # :44
    for j in range(1, len(levels)):                                          # :47
        if sigma is not None:
            lF = lF + np.sum(q ** 2, axis=1) / 2                             # :50
        n, xs, rk, c = levels[j]
        q, dl = tt_rt_sqr_oracle(n, xs, rk, c, q)                            # :52
        if sigma is not None:
            q = erfinv((q - 0.5) / cdf_factor) * math.sqrt(2.0)             # :55
        lF = lF + dl                                                         # :57
    return q, lF
