"""Output / restart formats: the reference's NetCDF schema and a raw checkpoint."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params

pytestmark = pytest.mark.gpu


def test_box_output_and_checkpoint(tmp_path):
    import cloudy_b200 as cb
    from cloudy_b200 import output, workloads as W
    par, state = W.c2_gamma_gamma(n_parcels=1)
    model = cb.CoalescenceModel(par)
    u = model.ensemble(1).upload(state)
    traj, times = [state[0].copy()], [0.0]
    for s in range(4):
        model.ssprk33_steps(u, par.dt, 1, cb.MODEL_BOX)
        traj.append(u.download()[0].copy())
        times.append((s + 1) * par.dt)
    fn = str(tmp_path / "box.nc")
    output.box_output(times, np.array(traj), par, fn)
    with netcdf_file(fn, "r", mmap=False) as ds:
        assert set(ds.variables) >= {"time", "moments", "total_moments", "params"}
        assert ds.variables["moments"].shape == (5, 2, 3)
        m = ds.variables["moments"][:].copy()
        tot = ds.variables["total_moments"][:].copy()
        prm = ds.variables["params"][:].copy()
    assert np.array_equal(m.reshape(5, 6), np.array(traj))
    assert np.allclose(tot, np.array(traj)[:, :3] + np.array(traj)[:, 3:], rtol=1e-15)
    assert np.allclose(tot[:, 1], tot[0, 1], rtol=1e-12)  # mass conserved
    d = O.update_dist_from_moments(O.Dist(O.GAMMA, 0.0, 1.0, 1.0), traj[2][:3])
    assert np.allclose(prm[2, 0], (d.n, d.theta, d.k), rtol=1e-13)
    ck = str(tmp_path / "restart.npz")
    output.save_checkpoint(ck, u, 4, times[-1], par)
    u2, step, t = output.load_checkpoint(ck, model)
    assert step == 4 and t == times[-1] and np.array_equal(u2.download(), u.download())
    model.ssprk33_steps(u, par.dt, 2, cb.MODEL_BOX)
    model.ssprk33_steps(u2, par.dt, 2, cb.MODEL_BOX)
    assert np.array_equal(u2.download(), u.download())  # restart reproduces the run bit for bit


def test_rainshaft_output(tmp_path):
    import cloudy_b200 as cb
    from cloudy_b200 import output, workloads as W
    par, cols = W.c3_rainshaft(n_columns=1, nz=20)
    model = cb.CoalescenceModel(par, nz=20)
    u = model.ensemble(20).upload(cols[0])
    states = [cols[0].copy()]
    for _ in range(2):
        model.ssprk33_steps(u, par.dt, 20, cb.MODEL_RAINSHAFT)
        states.append(u.download().copy())
    z = (np.arange(20) + 0.5) * par.dz
    fn = str(tmp_path / "rain.nc")
    output.rainshaft_output(z, [0.0, 20.0, 40.0], np.array(states), par, fn, model=model)
    with netcdf_file(fn, "r", mmap=False) as ds:
        assert ds.variables["moments"].shape == (3, 20, 2, 3)
        nc, nr, mc, mr = (ds.variables[k][:].copy() for k in ("Nc", "Nr", "Mc", "Mr"))
    st = np.array(states)
    # a mode whose RAW first moment is below eps is rebuilt as empty (ParticleDistributions.jl:461 on unnormalised moments,
    # as netcdf_helpers.jl:108-112 does), so compare where both modes are either absent or well resolved
    ok = ((st[..., 1] == 0) | (st[..., 1] > 1e-12)) & ((st[..., 4] == 0) | (st[..., 4] > 1e-12))
    assert ok.mean() > 0.5
    assert np.allclose((nc + nr)[ok], (st[..., 0] + st[..., 3])[ok], rtol=1e-12, atol=1e-300)
    assert np.allclose((mc + mr)[ok], (st[..., 1] + st[..., 4])[ok], rtol=1e-12, atol=1e-300)
    iz = int(np.argmax(st[2, :, 3]))
    pd = [O.update_dist_from_moments(O.Dist(O.GAMMA, 0.0, 1.0, 1.0), st[2, iz, 3 * j:3 * j + 3]) for j in range(2)]
    ref = O.get_standard_N_q(pd, 5.236e-10)
    assert np.allclose((nc[2, iz], nr[2, iz], mc[2, iz], mr[2, iz]), ref, rtol=1e-10)
