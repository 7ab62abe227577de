"""Generate tests/golden/*.npz: oracle (scipy restatement) tendencies for seeded inputs of every workload family.

The reference is Julia and cannot run in the build container, so these are RESTATEMENT golden vectors (the oracle is
itself pinned to the reference's own golden values, tests/test_oracle_goldens.py).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cloudy_b200 import workloads as W  # noqa: E402  (host-side generators only; no GPU needed)
from oracle import cloudy_oracle as O  # noqa: E402
from tests.oracle_bridge import oracle_params  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "c1_smoluchowski": (W.c1_smoluchowski, {}, 1),
    "c2_gamma_exp": (W.c2_gamma_exp, {"n_parcels": 96}, 96),
    "c2_gamma_gamma": (W.c2_gamma_gamma, {"n_parcels": 48}, 48),
    "mono_gamma": (W.mono_gamma, {"n_parcels": 32}, 32),
    "long_kernel_two_modes": (W.long_kernel_two_modes, {"n_parcels": 32}, 32),
    "c4_three_modes": (W.c4_three_modes, {"n_parcels": 24}, 24),
    "moving_four_modes": (W.moving_four_modes, {"n_parcels": 24}, 24),
    "moving_gamma_exp": (W.moving_gamma_exp, {"n_parcels": 32}, 32),
}


def main():
    for name, (gen, kw, n) in CASES.items():
        par, state = gen(**kw)
        opar = oracle_params(par)
        ref = np.zeros((n, state.shape[1]))
        scale = np.zeros_like(ref)
        for i in range(n):
            ref[i], scale[i] = O.rhs_coal(state[i], opar, return_scale=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), state=state[:n], tendency=ref, scale=scale)
        print(name, ref.shape)
    # one rainshaft column: RHS and 10 SSPRK33 steps
    par, cols = W.c3_rainshaft(n_columns=1, nz=20)
    rng = np.random.default_rng(123)
    st = cols[0].copy()
    frac = rng.uniform(0.0, 2e-3, st.shape[0]) * (st[:, 1] > 0)
    th = np.exp(rng.uniform(np.log(1.0), np.log(8.0), st.shape[0])) * 1e-9
    k = rng.uniform(0.8, 3.0, st.shape[0])
    m1 = st[:, 1] * frac
    st[:, 3], st[:, 4], st[:, 5] = m1 / (th * k), m1, m1 * th * (k + 1)
    st[3, 2] = -1e-20  # a negative entry: clipped in place by the RHS
    opar = oracle_params(par)
    m = st.copy()
    rhs, scale = O.rainshaft_rhs(m, opar, return_scale=True)
    after = O.ssprk33(lambda mm: O.rainshaft_rhs(mm, opar), st.copy(), par.dt, 10)
    after[after < 0] = 0
    np.savez_compressed(os.path.join(HERE, "c3_rainshaft_column.npz"), state=st, clipped=m, rhs=rhs, scale=scale, after10=after)
    print("c3_rainshaft_column", rhs.shape)


if __name__ == "__main__":
    main()
