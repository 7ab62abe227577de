"""Depth table of the Legendre continued fraction of Gamma(a, z) used beyond the series limit (csrc/special.cuh kCfDepthZ):
smallest depth d of the forward (Wallis) recurrence such that |Q_d(a,z) - Q(a,z)| <= 1e-16 (Q = Gamma(a,z)/Gamma(a), i.e. the
error is relative to the LOWER function the kernels need), for every a in [ai, ai+1) and every z >= limit[ai] + 4 b.
40-digit mpmath evaluation; prints C initialisers.   python tools/gen_cf_depth.py"""
import mpmath as mp
mp.mp.dps = 40
LIMIT = [18.0, 18.0, 19.0, 19.0, 20.0, 20.0, 21.0, 21.0, 22.0, 22.0, 23.0, 23.0, 24.0, 25.0, 25.0, 26.0, 26.0, 26.0]
NB, STEP, MAXD = 16, 4.0, 16


def cf_q(a, z, d):
    a, z = mp.mpf(a), mp.mpf(z)
    b = z + 1 - a
    Pm, Pc, Qm, Qc = mp.mpf(1), b, mp.mpf(0), mp.mpf(1)
    for n in range(1, d + 1):
        an = n * (a - n)
        b += 2
        Pm, Pc = Pc, b * Pc + an * Pm
        Qm, Qc = Qc, b * Qc + an * Qm
    return mp.exp(a * mp.log(z) - z - mp.loggamma(a)) * Qc / Pc


rows = []
for ai in range(18):
    row = []
    for b in range(NB):
        zlo = LIMIT[ai] + STEP * b
        need = 0
        for a in [ai + f for f in (0.0, 0.25, 0.5, 0.75, 0.999)]:
            if a <= 0:
                continue
            for z in (zlo, zlo + 1.0, zlo + 2.0, zlo + 3.999) if b < NB - 1 else (zlo, zlo + 10, zlo + 40, zlo + 150):
                exact = mp.gammainc(a, z, mp.inf, regularized=True)
                d = 0
                while d < MAXD + 8 and abs(cf_q(a, z, d) - exact) > mp.mpf("1e-16"):
                    d += 1
                need = max(need, d)
        row.append(need)
    rows.append(row)
print("// rows: floor(a_top) = 0..17; columns: z in [limit + 4b, limit + 4b + 4), last column: everything beyond")
for r in rows:
    print("    {" + ", ".join(f"{v:2d}" for v in r) + "},")
print("max", max(max(r) for r in rows))
