"""Generates the polynomial coefficients of mcrt_numerics.h (exp / log / sin / cos cores) with mpmath:
Chebyshev-interpolation polynomials (near-minimax) of the reduced-range kernels, rounded to double.
Prints C initialisers and the approximation error of the ROUNDED coefficients.

    python scripts/gen_numerics_coeffs.py
"""
import mpmath as mp

mp.mp.dps = 60


def fit(f, a, b, n):
    """degree n-1 polynomial (n coefficients, ascending) near-minimax for f on [a, b]"""
    c = mp.chebyfit(f, [a, b], n)          # highest degree first
    return [mp.mpf(x) for x in reversed(c)]


def as_double(c):
    return [float(x) for x in c]


def max_err(f, coeffs, a, b, rel=False, n=4001):
    worst = mp.mpf(0)
    for i in range(n):
        x = a + (b - a) * mp.mpf(i) / (n - 1)
        p = mp.mpf(0)
        for ck in reversed(coeffs):
            p = p * x + mp.mpf(ck)
        e = abs(p - f(x))
        if rel:
            e = e / abs(f(x))
        worst = max(worst, e)
    return worst


def show(name, coeffs):
    print(f"// {name}")
    for k, ck in enumerate(coeffs):
        print(f"    const double {name}{k} = {float(ck).hex()};   // {float(ck)!r}")


# exp(r), |r| <= ln2/2 (+ slack): p(r) ~ exp(r), degree 11
LN2H = mp.log(2) / 2 * mp.mpf("1.0001")
ce = as_double(fit(mp.exp, -LN2H, LN2H, 12))
print("exp degree 11 max rel err", mp.nstr(max_err(mp.exp, ce, -LN2H, LN2H, rel=True), 5))
show("E", ce)

# log: log(m) = 2 s (1 + z Q(z)), s = (m-1)/(m+1), z = s^2 <= 0.0294373: H(z) = atanh(sqrt z)/sqrt z = 1 + z/3 + z^2/5 ...
ZMAX = ((mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)) ** 2 * mp.mpf("1.0001")


def H(z):
    if z == 0:
        return mp.mpf(1)
    s = mp.sqrt(z)
    return mp.atanh(s) / s


def Q(z):
    if z == 0:
        return mp.mpf(1) / 3
    return (H(z) - 1) / z


for n in (7, 8, 9):
    cq = as_double(fit(Q, 0, ZMAX, n))
    # error of log(m)/(2s) = 1 + z Q(z)
    worst = mp.mpf(0)
    for i in range(2001):
        z = ZMAX * mp.mpf(i) / 2000
        p = mp.mpf(0)
        for ck in reversed(cq):
            p = p * z + mp.mpf(ck)
        worst = max(worst, abs(1 + z * p - H(z)))
    print(f"log Q with {n} coeffs: max err of 1+zQ(z) vs atanh(s)/s", mp.nstr(worst, 5))
cq = as_double(fit(Q, 0, ZMAX, 8))
show("L", cq)

# sin(r) = r (1 + z S(z)), cos(r) = 1 - z/2 + z^2 C(z), z = r^2, |r| <= pi/4 (+ slack)
ZS = (mp.pi / 4 * mp.mpf("1.0001")) ** 2


def S(z):
    if z == 0:
        return -mp.mpf(1) / 6
    r = mp.sqrt(z)
    return (mp.sin(r) / r - 1) / z


def Cc(z):
    if z == 0:
        return mp.mpf(1) / 24
    r = mp.sqrt(z)
    return (mp.cos(r) - 1 + z / 2) / (z * z)


for n in (6, 7):
    cs = as_double(fit(S, 0, ZS, n))
    cc = as_double(fit(Cc, 0, ZS, n))
    ws = wc = mp.mpf(0)
    for i in range(2001):
        z = ZS * mp.mpf(i) / 2000
        ps = pc = mp.mpf(0)
        for ck in reversed(cs):
            ps = ps * z + mp.mpf(ck)
        for ck in reversed(cc):
            pc = pc * z + mp.mpf(ck)
        r = mp.sqrt(z)
        ws = max(ws, abs(r * (1 + z * ps) - mp.sin(r)))
        wc = max(wc, abs(1 - z / 2 + z * z * pc - mp.cos(r)))
    print(f"sin/cos with {n} coeffs: abs err sin {mp.nstr(ws, 5)} cos {mp.nstr(wc, 5)}")
cs = as_double(fit(S, 0, ZS, 6))
cc = as_double(fit(Cc, 0, ZS, 6))
show("S", cs)
show("C", cc)

# exp(r) = 1 + r + r^2 G(r): the two leading terms exact, G of degree 9
def G(r):
    if r == 0:
        return mp.mpf(1) / 2
    return (mp.exp(r) - 1 - r) / (r * r)


for n in (9, 10, 11):
    cg = as_double(fit(G, -LN2H, LN2H, n))
    worst = mp.mpf(0)
    for i in range(4001):
        r = -LN2H + 2 * LN2H * mp.mpf(i) / 4000
        p = mp.mpf(0)
        for ck in reversed(cg):
            p = p * r + mp.mpf(ck)
        worst = max(worst, abs((1 + r + r * r * p) / mp.exp(r) - 1))
    print(f"exp via G with {n} coeffs: max rel err", mp.nstr(worst, 5))
cg = as_double(fit(G, -LN2H, LN2H, 10))
show("G", cg)
cq = as_double(fit(Q, 0, ZMAX, 7))
show("L", cq)
