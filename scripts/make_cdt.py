"""Cumulative tables of the rounded Gaussians used by the counter-based generators (csrc/kernels.cuh: CLUE_CDT, csrc/keygen.cuh:
KG_CDT_*, and their twins in oracle/omr_oracle.hpp):  P(|e| <= k) * 2^32 for k = 0, 1, ... until the tail is below 2^-32, with
e = round(N(0, sigma^2)), i.e. P(|e| <= k) = erf((k + 1/2) / (sigma sqrt 2)).  Sigmas: parameters/mod.rs:45,54,80,88; the
key-switching sigma 2.0329 * 2^10 (:58-66) is realised as 512 x + U[-256, 256) with Var = 512^2 (sigma_x^2 + 1/12)."""
import math


def cdt(sigma):
    out, k = [], 0
    while True:
        v = round(math.erf((k + 0.5) / (sigma * math.sqrt(2))) * 2 ** 32)
        if v >= 2 ** 32:
            return out
        out.append(v); k += 1


if __name__ == "__main__":
    ks = 2.0329 * 2 ** 10
    for name, s in (("CLUE_CDT (0.8293)", 0.8293), ("KG_CDT_L1 (3.1859)", 3.1859), ("KG_CDT_L2 (0.3908)", 0.3908),
                    ("KG_CDT_KS (x of sigma %.4f for 512 x + U)" % math.sqrt((ks ** 2 - 512 ** 2 / 12) / 512 ** 2), math.sqrt((ks ** 2 - 512 ** 2 / 12) / 512 ** 2))):
        t = cdt(s)
        print(f"{name}: {len(t)} entries\n  " + ", ".join(f"{v}u" for v in t))
