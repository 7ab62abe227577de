"""Full-size and full-length parity cases: the reference's 1000-step rainshaft run against the C oracle,
C3 at 4096 x 256 cells and C5 at 67,108,864 parcels through size-independent properties plus oracle spot checks."""
import numpy as np
import pytest

from oracle import c_oracle, cloudy_oracle as O
from tests.oracle_bridge import oracle_params, tendency_close

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def cb():
    import cloudy_b200
    return cloudy_b200


def _cfg(cb, par, nz=1):
    kinds = tuple(d.kind for d in par.pdists)
    return cb.build_config(kinds, par.coal_data, norms=par.norms, vel=tuple(getattr(par, "vel", ())), dz=getattr(par, "dz", 1.0), nz=nz)


def test_rainshaft_reference_run_1000_steps(cb):
    """rainshaft_gamma_mixture.jl:13-49 in full: 20 levels over 3000 m, dt = 1 s, tspan = 1000 s -> 1000 SSPRK33 steps.  Column 0 is
    the script's own initial condition, column 1 a scaled one.  Integrated moments against the C oracle (same scheme, same
    right-hand side, oracle/cloudy_oracle.c) at 1e-7 of the column's largest value per moment."""
    from cloudy_b200 import workloads as W
    nz, nsteps = 20, 1000
    par, cols = W.c3_rainshaft(n_columns=2, nz=nz)
    z = np.arange(par.dz / 2, 3000.0, par.dz)[:nz]
    cols[0] = O.initial_condition(z, [1e7, 1e-3, 2e-13, 0.0, 0.0, 0.0])      # the script's ic, amplitude factor 1
    st = np.ascontiguousarray(cols)
    model = cb.CoalescenceModel(par, nz=nz)
    u = model.ensemble(2 * nz).upload(st.reshape(-1, 6))
    model.ssprk33_steps(u, par.dt, nsteps, cb.MODEL_RAINSHAFT)
    got = u.download().reshape(st.shape)
    ref = c_oracle.ssprk33(_cfg(cb, par, nz), st, par.dt, nsteps, model=1)
    ref[ref < 0] = 0  # the stage kernels clip what they write (the reference clips at the next evaluation, rainshaft_helpers.jl:52)
    assert np.all(np.isfinite(ref)) and np.all(np.isfinite(got))
    assert ref[0, :, 3].max() > 0  # rain has formed and fallen
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True) * 1e-6)
    err = np.abs(got - ref) / scale
    assert err.max() <= 1e-7, err.max()
    assert model.ctx.error_count() == 0


def _rain_state(cols, seed):
    rng = np.random.default_rng(seed)
    st = cols.copy()
    shp = st.shape[:-1]
    frac = rng.uniform(0.0, 2e-3, shp) * (st[..., 1] > 0)
    th = np.exp(rng.uniform(np.log(1.0), np.log(8.0), shp)) * 1e-9
    k = rng.uniform(0.8, 3.0, shp)
    m1 = st[..., 1] * frac
    st[..., 3] = m1 / (th * k)
    st[..., 4] = m1
    st[..., 5] = m1 * th * (k + 1)
    # negative entries (what a sedimentation step leaves behind just outside the cloud): only in cells that are otherwise empty,
    # so that the active cells keep consistent moment sets and the mass budget below is sharp
    empty = (st[..., 1] == 0)[..., None] & (rng.random(st.shape) < 0.02)
    st[empty] = -1e-9 * np.array([1e7, 1e-3, 2e-13, 1e7, 1e-3, 2e-13])[np.nonzero(empty)[-1]]
    return np.ascontiguousarray(st)


def test_c3_full_size_properties(cb):
    """BASELINE configs[2] at full size (4096 columns x 256 levels): the in-place clip, the column mass budget (coalescence
    conserves mass in every cell, the upwind divergence telescopes to the flux through the column bottom), invariance under a
    permutation of the columns (bit-exact), determinism, and oracle spot checks of whole columns."""
    from cloudy_b200 import workloads as W
    ncol, nz = 4096, 256
    par, cols = W.c3_rainshaft(n_columns=ncol, nz=nz)
    st = _rain_state(cols, seed=21)
    flat = st.reshape(-1, 6)
    model = cb.CoalescenceModel(par, nz=nz)
    u = model.ensemble(flat.shape[0]).upload(flat); du = model.ensemble(flat.shape[0]); fl = model.ensemble(flat.shape[0])
    model.rainshaft_rhs(u, du)
    got = du.download().reshape(st.shape)
    clipped = np.maximum(st, 0.0)
    assert np.array_equal(u.download().reshape(st.shape), clipped)          # rainshaft_helpers.jl:52
    assert np.all(np.isfinite(got))
    model.sedimentation_flux(u, fl)
    flux = fl.download().reshape(st.shape)
    # column mass budget: sum over levels and modes of dM1/dt = (flux of M1 through the bottom face)/dz  (flux <= 0: downwards)
    dmass = (got[:, :, 1] + got[:, :, 4]).sum(axis=1)
    bottom = (flux[:, 0, 1] + flux[:, 0, 4]) / par.dz
    scale = (np.abs(got[:, :, 1]) + np.abs(got[:, :, 4])).sum(axis=1) + np.abs(bottom) + 1e-300
    assert np.all(np.abs(dmass - bottom) <= 1e-9 * scale), np.max(np.abs(dmass - bottom) / scale)
    assert np.all(flux <= 0.0)
    # columns are independent: permuting them permutes the result bit for bit
    perm = np.random.default_rng(2).permutation(ncol)
    up = model.ensemble(flat.shape[0]).upload(st[perm].reshape(-1, 6)); dup = model.ensemble(flat.shape[0])
    model.rainshaft_rhs(up, dup)
    assert np.array_equal(dup.download().reshape(st.shape), got[perm])
    model.rainshaft_rhs(u, du)
    assert np.array_equal(du.download().reshape(st.shape), got)
    # whole columns against the scipy oracle (cancellation-aware tolerance) and, with the same tolerance, the C oracle
    opar = oracle_params(par)
    sel = np.array([0, 1777, 4095])
    ref_c = c_oracle.rainshaft_rhs(_cfg(cb, par, nz), np.ascontiguousarray(st[sel]))
    for q, c in enumerate(sel):
        ref, sc = O.rainshaft_rhs(st[c].copy(), opar, return_scale=True)
        ok, worst = tendency_close(got[c], ref, sc, RTOL)
        assert ok, (c, worst)
        ok, worst = tendency_close(got[c], ref_c[q], sc, RTOL)
        assert ok, (c, worst)


def test_c5_full_size_properties(cb):
    """BASELINE configs[4] on one GPU: 67,108,864 parcels (the generator's 8 Mi-parcel block repeated 8 times, as bench.py
    builds it).  Mass tendency cancels, number tendency is non-positive, a parcel's tendency does not depend on its position
    (the 8 copies agree bit for bit, although the regime sort sends them to different warps), determinism, oracle spot checks."""
    from cloudy_b200 import workloads as W
    block, reps = 1 << 23, 8
    par, base = W.c2_gamma_exp(n_parcels=block, seed=W.SEED0 + 5)
    state = np.ascontiguousarray(np.tile(base, (reps, 1)))
    n = state.shape[0]
    assert n == 1 << 26
    model = cb.CoalescenceModel(par)
    u = model.ensemble(n).upload(state); du = model.ensemble(n)
    model.coal_tendency(u, du)
    assert u.order() is not None                       # the ensemble is resident in regime order
    got = du.download()
    first = got[:block]
    for r in range(1, reps):
        assert np.array_equal(got[r * block:(r + 1) * block], first), r
    del got
    mass = first[:, 1] + first[:, 4]
    scale = np.abs(first[:, 1]) + np.abs(first[:, 4]) + 1e-300
    assert np.all(np.abs(mass) <= 1e-9 * scale + 1e-9 * np.abs(base[:, 1] + base[:, 4]))
    assert np.all(first[:, 0] + first[:, 3] <= 0)
    assert np.all(np.isfinite(first))
    model.coal_tendency(u, du)
    assert np.array_equal(du.download()[:block], first)
    assert np.array_equal(u.download()[:block], base)  # the input is logically untouched by the sort
    sums = model.moment_sums(u)
    assert np.allclose(sums, base.sum(axis=0) * reps, rtol=1e-11, atol=0)
    opar = oracle_params(par)
    for i in range(0, block, block // 24):
        ref, sc = O.rhs_coal(base[i], opar, return_scale=True)
        ok, worst = tendency_close(first[i], ref, sc, RTOL)
        assert ok, (i, worst)
    assert model.ctx.error_count() == 0
