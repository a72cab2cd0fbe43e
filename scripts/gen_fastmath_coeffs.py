"""Generate the polynomial coefficients of pyhillfit_b200/csrc/phf_fastmath.cuh (developer tool).

Each kernel function is a near-minimax polynomial obtained by Chebyshev interpolation in 60-digit arithmetic
(mpmath) on the reduced interval, converted to the monomial basis and rounded to double.  The script prints the
max relative / absolute error of the ROUNDED polynomial evaluated in high precision, then writes
pyhillfit_b200/csrc/phf_fastmath_coeffs.inc.

    python scripts/gen_fastmath_coeffs.py
"""
import os
import sys

import mpmath as mp

mp.mp.dps = 60
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cheb_fit(f, a, b, deg):
    """Monomial coefficients (in x) of the degree-`deg` Chebyshev interpolant of f on [a,b]."""
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    xs = [(a + b) / 2 + (b - a) / 2 * t for t in nodes]
    ys = [f(x) for x in xs]
    # Chebyshev coefficients
    c = []
    for j in range(n):
        s = mp.fsum(ys[k] * mp.cos(mp.pi * j * (2 * k + 1) / (2 * n)) for k in range(n))
        c.append(2 * s / n)
    c[0] /= 2
    # T_j(t) in monomial basis of t, then substitute t = (2x - a - b)/(b - a)
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        prev, prev2 = T[j - 1], T[j - 2]
        cur = [mp.mpf(0)] + [2 * v for v in prev]
        for i, v in enumerate(prev2):
            cur[i] -= v
        T.append(cur)
    pt = [mp.mpf(0)] * n
    for j in range(n):
        for i, v in enumerate(T[j]):
            pt[i] += c[j] * v
    # substitute t = alpha x + beta
    alpha = 2 / (b - a)
    beta = -(a + b) / (b - a)
    px = [mp.mpf(0)] * n
    # (alpha x + beta)^i expansion
    lin_pow = [mp.mpf(1)]
    for i in range(n):
        for k, v in enumerate(lin_pow):
            px[k] += pt[i] * v
        nxt = [mp.mpf(0)] * (len(lin_pow) + 1)
        for k, v in enumerate(lin_pow):
            nxt[k] += v * beta
            nxt[k + 1] += v * alpha
        lin_pow = nxt
    return px


def as_double(c):
    return [float(v) for v in c]


def horner(c, x):
    r = mp.mpf(0)
    for v in reversed(c):
        r = r * x + mp.mpf(v)
    return r


def max_err(approx, exact, a, b, npts=4001, rel=True):
    worst = mp.mpf(0)
    for k in range(npts):
        x = a + (b - a) * mp.mpf(k) / (npts - 1)
        e = exact(x)
        d = abs(approx(x) - e)
        if rel and e != 0:
            d /= abs(e)
        worst = max(worst, d)
    return float(worst)


LOG_DEG, SIN_DEG, COS_DEG, EXP_DEG = 6, 5, 5, 9
ERFCX_KS = [mp.mpf(k) for k in (3, 4, 5, 6)]
ERFCX_DEGS = (16, 18, 20, 22)
ERFCX_TOL = 1.5e-15


def main():
    out = []

    # ---- exp(r) = 1 + r + r^2 Q(r),  |r| <= ln2/2 ----
    h = mp.log(2) / 2 * mp.mpf("1.002")   # n is picked with a 21-bit log2(e): |r| can exceed ln2/2 by 3e-4
    fq = lambda r: (mp.exp(r) - 1 - r) / r ** 2 if abs(r) > mp.mpf('1e-12') else mp.mpf(1) / 2 + r / 6 + r * r / 24
    q = as_double(cheb_fit(fq, -h, h, EXP_DEG))
    err = max_err(lambda r: 1 + r + r * r * horner(q, r), mp.exp, -h, h)
    print("exp   : deg(Q)=%d  max rel err %.3e" % (EXP_DEG, err))
    out.append(("kExpQ", q))

    # ---- log(m) = 2 f + f^3 R(f^2), f = (m-1)/(m+1), m in [sqrt(1/2), sqrt 2] ----
    fmax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
    smax = fmax ** 2 * mp.mpf("1.001")
    fr = lambda s: (2 * mp.atanh(mp.sqrt(s)) - 2 * mp.sqrt(s)) / mp.sqrt(s) ** 3 if s != 0 else mp.mpf(2) / 3
    r = as_double(cheb_fit(fr, mp.mpf(0), smax, LOG_DEG))
    def log_ap(f):
        s = f * f
        return 2 * f + f * s * horner(r, s)
    err = max_err(log_ap, lambda f: 2 * mp.atanh(f), -fmax, fmax)
    print("log   : deg(R)=%d   max rel err of log(m) %.3e" % (LOG_DEG, err))
    out.append(("kLogR", r))

    # ---- sin(r) = r + r^3 S(r^2), cos(r) = 1 - r^2/2 + r^4 C(r^2), |r| <= pi/4 ----
    smax = (mp.pi / 4) ** 2 * mp.mpf("1.0001")
    fs = lambda s: (mp.sin(mp.sqrt(s)) - mp.sqrt(s)) / mp.sqrt(s) ** 3 if s != 0 else -mp.mpf(1) / 6
    fc = lambda s: (mp.cos(mp.sqrt(s)) - 1 + s / 2) / s ** 2 if s != 0 else mp.mpf(1) / 24
    sc = as_double(cheb_fit(fs, mp.mpf(0), smax, SIN_DEG))
    cc = as_double(cheb_fit(fc, mp.mpf(0), smax, COS_DEG))
    es = max_err(lambda x: x + x ** 3 * horner(sc, x * x), mp.sin, -mp.pi / 4, mp.pi / 4, rel=False)
    ec = max_err(lambda x: 1 - x * x / 2 + x ** 4 * horner(cc, x * x), mp.cos, -mp.pi / 4, mp.pi / 4, rel=False)
    print("sin   : deg(S)=%d   max abs err %.3e ; cos: deg(C)=%d max abs err %.3e" % (SIN_DEG, es, COS_DEG, ec))
    out.append(("kSinS", sc))
    out.append(("kCosC", cc))

    # ---- erfcx(x) (1 + 2x) = P(q), q = (x - K)/(x + K), x in [0, inf) <-> q in [-1, 1) ----
    best = None
    for K in ERFCX_KS:
        for deg in ERFCX_DEGS:
            def fp(qv, K=K):
                if qv >= 1:
                    return 2 / mp.sqrt(mp.pi)
                x = K * (1 + qv) / (1 - qv)
                if x > 40:   # asymptotic series, erfc underflows in mp.erfc * exp(x^2) only at huge x; mp handles it
                    pass
                return mp.exp(x * x) * mp.erfc(x) * (1 + 2 * x)
            p = as_double(cheb_fit(fp, mp.mpf(-1), mp.mpf(1), deg))
            err = max_err(lambda qv: horner(p, qv), fp, mp.mpf(-1), mp.mpf("0.9999"), npts=3001)
            print("erfcx : K=%s deg=%d max rel err %.3e" % (K, deg, err))
            if best is None or (err < ERFCX_TOL and (deg < best[1] or best[3] > ERFCX_TOL or (deg == best[1] and err < best[3]))) or (best[3] > ERFCX_TOL and err < best[3]):
                best = (K, deg, p, err)
    K, deg, p, err = best
    print("erfcx : using K=%s deg=%d err %.3e" % (K, deg, err))
    out.append(("kErfcxP", p))

    path = os.path.join(ROOT, "pyhillfit_b200", "csrc", "phf_fastmath_coeffs.inc")
    misc = [("kLn2Lo", mp.log(2) - mp.mpf(float.fromhex("0x1.62e43p-1"))), ("kPiOver2p31", mp.pi / 2 ** 31),
            ("kSqrtHalfC", mp.sqrt(mp.mpf(1) / 2)), ("kLn10HiC", None), ("kLn10LoC", None)]
    ln10hi = float(mp.log(10))
    vals_misc = [float(misc[0][1]), float(misc[1][1]), float(misc[2][1]), ln10hi, float(mp.log(10) - mp.mpf(ln10hi)), 0.0]
    with open(path, "w") as f:
        f.write("// generated by scripts/gen_fastmath_coeffs.py -- do not edit\n")
        f.write("// one table, every polynomial starts on a 16-byte boundary (pairs are fetched with one 128-bit load)\n")
        f.write("#define PHF_ERFCX_K %s\n" % repr(float(K)))
        off = 0
        body = []
        for name, c in out + [("kMisc", vals_misc)]:
            vals = list(c) + ([0.0] if len(c) & 1 else [])
            f.write("#define PHF_FM_%s %d\n" % (name.upper(), off))
            for v in vals:
                body.append("    %s,  // %s[%d]" % (float(v).hex(), name, len(body) - off))
            off += len(vals)
        f.write("#define PHF_FM_TABLE_SIZE %d\n" % off)
        f.write("PHF_COEFF_TABLE double kFmTable[PHF_FM_TABLE_SIZE] = {\n")
        f.write("\n".join(body))
        f.write("\n};\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
