"""Generates tests/golden/special_functions.json: reference values of the SpecialFunctions-extension operators
(/root/reference/ext/functionlist.jl) from mpmath at 40 digits -- the reference's own values come from SpecialFunctions.jl /
openspecfun, which is not vendored and cannot run here.  Run from the repo root: python tests/golden/make_special_golden.py"""
import json
import os

import mpmath as mp

mp.mp.dps = 40
F1 = {
    "erf": mp.erf, "erfc": mp.erfc, "erfi": mp.erfi, "erfcx": lambda x: mp.exp(x * x) * mp.erfc(x),
    "digamma": mp.digamma, "trigamma": lambda x: mp.polygamma(1, x),
    "invdigamma": lambda y: mp.findroot(lambda t: mp.digamma(t) - y, mp.exp(y) + 0.5 if y >= -2.22 else -1 / (y - mp.digamma(1))),
    "gamma": mp.gamma, "airyai": mp.airyai, "airybi": mp.airybi,
    "airyaiprime": lambda x: mp.airyai(x, derivative=1), "airybiprime": lambda x: mp.airybi(x, derivative=1),
    "besselj0": lambda x: mp.besselj(0, x), "bessely0": lambda x: mp.bessely(0, x),
    "besselj1": lambda x: mp.besselj(1, x), "bessely1": lambda x: mp.bessely(1, x),
    "dawson": lambda x: mp.sqrt(mp.pi) / 2 * mp.exp(-x * x) * mp.erfi(x),
    "erfinv": mp.erfinv, "erfcinv": lambda y: mp.erfinv(1 - y),
}
PTS = {
    "default": [-3.7, -1.25, -0.3, 0.05, 0.7, 1.3, 2.9, 6.1, 11.5],
    "erfi": [-5.5, -1.25, -0.3, 0.05, 0.7, 1.3, 2.9, 6.1, 7.3],
    "erfcx": [-2.5, -0.3, 0.05, 0.7, 2.9, 11.5, 40.0],
    "digamma": [-3.7, -1.25, -0.3, 0.05, 0.7, 1.3, 2.9, 6.1, 11.5, 150.0],
    "trigamma": [-3.7, -1.25, -0.3, 0.05, 0.7, 1.3, 2.9, 6.1, 11.5, 150.0],
    "invdigamma": [-6.0, -2.5, -0.3, 0.05, 0.7, 2.9, 5.0],
    "gamma": [-3.7, -1.25, -0.3, 0.05, 0.7, 1.3, 2.9, 6.1, 11.5, 25.5],
    "airyai": [-35.2, -20.4999, -11.37, -3.7, -0.3, 0.0, 0.7, 2.9, 6.1, 11.5, 20.4999, 33.1],
    "bessely0": [0.05, 0.7, 1.3, 2.9, 6.1, 11.5, 40.3], "bessely1": [0.05, 0.7, 1.3, 2.9, 6.1, 11.5, 40.3],
    "besselj0": [-3.7, 0.0, 0.05, 0.7, 2.9, 6.1, 11.5, 40.3], "besselj1": [-3.7, 0.0, 0.05, 0.7, 2.9, 6.1, 11.5, 40.3],
    "dawson": [-7.3, -3.7, -0.3, 0.05, 0.7, 1.3, 2.9, 6.1, 6.49, 6.51, 11.5, 80.0],
    "erfinv": [-0.999, -0.7, -0.05, 0.0, 0.3, 0.7, 0.95, 0.999999], "erfcinv": [1e-12, 0.001, 0.3, 0.7, 1.0, 1.3, 1.95, 1.999],
}
for k in ("airybi", "airyaiprime", "airybiprime"):
    PTS[k] = PTS["airyai"]
def main():
    out = {"univariate": {}, "bivariate": {}}
    for name, f in F1.items():
        xs = PTS.get(name, PTS["default"])
        out["univariate"][name] = [[x, float(f(mp.mpf(x)))] for x in xs]
    # higher polygammas (used by the derivative entries of digamma / trigamma / invdigamma / gamma)
    out["polygamma2"] = [[x, float(mp.polygamma(2, mp.mpf(x)))] for x in PTS["digamma"]]
    out["polygamma3"] = [[x, float(mp.polygamma(3, mp.mpf(x)))] for x in PTS["digamma"]]
    B = [(0.7, 1.3), (2.5, 3.0), (0.05, 4.2), (11.5, 6.1), (30.0, 45.0)]
    out["bivariate"]["beta"] = [[a, b, float(mp.beta(a, b))] for a, b in B]
    out["bivariate"]["logbeta"] = [[a, b, float(mp.log(mp.beta(a, b)))] for a, b in B]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "special_functions.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", path)


if __name__ == "__main__":
    main()
