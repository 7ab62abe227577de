"""Output / restart formats either side of the path.

NetCDF files with the reference's schema (test/examples/utils/netcdf_helpers.jl:10-61 box, :63-125 rainshaft):
same dimension and variable names, written in NetCDF-3 classic format with ``scipy.io.netcdf_file`` (the reference
writes NetCDF-4 through NCDatasets.jl; readers address variables by name either way).  Derived quantities come from
the device: distribution parameters through ``update_dist_from_moments``, the cloud/rain split through the batched
``get_standard_N_q`` kernel.  A raw checkpoint (moments + step counter) supports restart."""
import json

import numpy as np
from scipy.io import netcdf_file

from .distributions import nparams, update_dist_from_moments
from .helpers import get_dist_moment_ind, get_dist_moments_ind_range


def _def_common(ds, times, par):
    Ndist = len(par.pdists)
    ds.createDimension("t", len(times))
    ds.createDimension("dist", Ndist)
    ds.createDimension("order", max(par.NProgMoms))
    t = ds.createVariable("time", "d", ("t",))
    t[:] = np.asarray(times, dtype=np.float64)
    return Ndist


def box_output(times, states, par, filename):
    """box_output(sol, p, filename, FT): ``states`` is (nt, n_moments) — sol.u of one box."""
    states = np.asarray(states, dtype=np.float64)
    nt = len(times)
    with netcdf_file(filename, "w") as ds:
        Ndist = _def_common(ds, times, par)
        nmax, nmin = max(par.NProgMoms), min(par.NProgMoms)
        M = ds.createVariable("moments", "d", ("t", "dist", "order"))
        Mtot = ds.createVariable("total_moments", "d", ("t", "order"))
        pp = ds.createVariable("params", "d", ("t", "dist", "order"))
        mom = np.zeros((nt, Ndist, nmax))
        tot = np.zeros((nt, nmax))
        params = np.zeros((nt, Ndist, nmax))
        for i in range(1, Ndist + 1):
            for j in range(1, par.NProgMoms[i - 1] + 1):
                ind = get_dist_moment_ind(par.NProgMoms, i, j) - 1
                mom[:, i - 1, j - 1] = states[:, ind]
                if j <= nmin:
                    tot[:, j - 1] += states[:, ind]
        for it in range(nt):
            for j in range(1, Ndist + 1):
                rng = get_dist_moments_ind_range(par.NProgMoms, j)
                d = update_dist_from_moments(par.pdists[j - 1], tuple(states[it, r - 1] for r in rng))
                p = d.params()
                params[it, j - 1, : len(p)] = p
        M[:] = mom
        Mtot[:] = tot
        pp[:] = params


def rainshaft_output(z, times, states, par, filename, model=None, size_cutoff=5.236e-10):
    """rainshaft_output(z, sol, p, filename, FT): ``states`` is (nt, nz, n_moments)."""
    from .ensemble import CoalescenceModel
    states = np.asarray(states, dtype=np.float64)
    nt, nz, nm = states.shape
    model = model or CoalescenceModel(par, nz=nz)
    with netcdf_file(filename, "w") as ds:
        Ndist = _def_common(ds, times, par)
        ds.createDimension("z", nz)
        zz = ds.createVariable("altitude", "d", ("z",))
        zz[:] = np.asarray(z, dtype=np.float64)
        nmax, nmin = max(par.NProgMoms), min(par.NProgMoms)
        M = ds.createVariable("moments", "d", ("t", "z", "dist", "order"))
        Mtot = ds.createVariable("total_moments", "d", ("t", "z", "order"))
        mom = np.zeros((nt, nz, Ndist, nmax))
        tot = np.zeros((nt, nz, nmax))
        for i in range(1, Ndist + 1):
            for j in range(1, par.NProgMoms[i - 1] + 1):
                ind = get_dist_moment_ind(par.NProgMoms, i, j) - 1
                mom[:, :, i - 1, j - 1] = states[:, :, ind]
                if j <= nmin:
                    tot[:, :, j - 1] += states[:, :, ind]
        M[:] = mom
        Mtot[:] = tot
        # cloud / rain number and mass of every (t, z) cell in one batched kernel launch (netcdf_helpers.jl:106-121)
        u = model.ensemble(nt * nz).upload(states.reshape(-1, nm))
        nq = model.standard_N_q(u, size_cutoff, normalized=False).reshape(4, nt, nz)
        for name, k in (("Nc", 0), ("Nr", 1), ("Mc", 2), ("Mr", 3)):
            v = ds.createVariable(name, "d", ("t", "z"))
            v[:] = nq[k]


def save_checkpoint(path, ensemble, step, time, par=None):
    """Raw restart file: the ensemble's moments (host layout) + step counter."""
    meta = {"step": int(step), "time": float(time), "n": int(ensemble.n), "n_slots": int(ensemble.n_slots),
            "NProgMoms": list(par.NProgMoms) if par is not None else None}
    np.savez(path, moments=ensemble.download(), meta=json.dumps(meta))


def load_checkpoint(path, model):
    """→ (ensemble on the model's device, step, time)"""
    with np.load(path if str(path).endswith(".npz") else str(path) + ".npz") as f:
        meta = json.loads(str(f["meta"]))
        mom = f["moments"]
    if meta["n_slots"] != model.n_slots:
        raise ValueError("checkpoint does not match the model's moment layout")
    return model.ensemble(mom.shape[0]).upload(mom), meta["step"], meta["time"]
