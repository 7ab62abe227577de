"""Series and Taylor degrees of the FixedThreshold node loop for a given truncation target (40-digit mpmath search).
   python tools/gen_degrees.py 1e-13
Series: S(z) = sum_n z^n/(a)_{n+1}; degree N(z_bin, a_bin) = smallest N with tail/S <= tol at the bin's largest z and smallest a.
Taylor: S(z) = sum_m t_m r^m about X_c (t from the three-term recurrence), r = z/X_c - 1 in [-rmax, 0.15 rmax]; degree K(rmax) =
smallest K with |tail|/S(z) <= tol for all a in [1, 18], X_c in (0, 25.5]."""
import sys
import mpmath as mp
mp.mp.dps = 40
tol = mp.mpf(sys.argv[1]) if len(sys.argv) > 1 else mp.mpf("1e-13")
KSERZ, KSERA = 26, 18
LIMIT = [18.0, 18.0, 19.0, 19.0, 20.0, 20.0, 21.0, 21.0, 22.0, 22.0, 23.0, 23.0, 24.0, 25.0, 25.0, 26.0, 26.0, 26.0]


def S_exact(a, z):
    # S = gamma(a,z) z^-a e^z = sum z^n/(a)_{n+1}
    s, t, n = mp.mpf(0), 1 / mp.mpf(a), 0
    while True:
        s += t
        n += 1
        t *= z / (a + n)
        if t < s * mp.mpf("1e-38"):
            return s


def series_degree(a, z):
    S = S_exact(a, z)
    s, t, n = mp.mpf(0), 1 / mp.mpf(a), 0
    while True:
        s += t
        if (S - s) <= tol * S:
            return n
        n += 1
        t *= z / (a + n)


print("series degrees (rows: floor(z) = 0..25 evaluated at z+1, columns floor(a) = 0..17 evaluated at max(a, 1)); 0 beyond the series limit")
for zi in range(KSERZ):
    row = []
    for ai in range(KSERA):
        z = mp.mpf(zi + 1)
        a = mp.mpf(max(ai, 1))
        row.append(series_degree(a, z) if z <= LIMIT[ai] + 0.5 else 0)
    print("    {" + ", ".join(f"{v:2d}" for v in row) + "},")

rho_thr = [1e-5, 1e-4, 1e-3, 3e-3, 1e-2, 2e-2, 3e-2, 5e-2, 7e-2, 0.1]
print("Taylor degrees by class (|r| <= 1.15 rho):")
out = []
for rho in rho_thr + [0.115 / 1.15]:
    rmax = mp.mpf(1.15 * rho) if rho != rho_thr[-1] or True else rho
    need = 0
    for a in (1, 1.5, 2, 3, 4.5, 6, 8, 11, 14, 17.99):
        for Xc in (0.05, 0.3, 1, 2, 4, 6, 9, 12, 15, 18, 21, 23.5, 25.5):
            if Xc > LIMIT[min(int(a), 17)] - 0.5 + 1e-9:
                continue
            a_, X = mp.mpf(a), mp.mpf(Xc)
            t = [S_exact(a_, X)]
            t.append((X - a_) * t[0] + 1)
            for m in range(1, 60):
                t.append(((X - a_ - m) * t[m] + X * t[m - 1]) / (m + 1))
            for r in (-rmax, mp.mpf("0.15") * rmax):
                z = X * (1 + r)
                if z <= 0:
                    continue
                Sz = S_exact(a_, z)
                s = mp.mpf(0)
                K = None
                for m in range(60):
                    s += t[m] * r ** m
                    if abs(Sz - s) <= tol * Sz:
                        K = m
                        break
                need = max(need, K if K is not None else 99)
    out.append(need)
print("rho_thr", rho_thr + ["(capped: 0.115)"])
print("K     ", out)
