"""CPU-only: host logic of the mirror, the C-ABI library's exported symbols, and workload generators."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import cloudy_b200
    from cloudy_b200 import _lib
    header = open(os.path.join(ROOT, "include", "cloudy_b200.h")).read()
    declared = set(re.findall(r"\b(cloudy_[A-Za-z0-9_]+)\s*\(", header))
    declared -= {"cloudy_ctx", "cloudy_state", "cloudy_config"}
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/cloudy_b200.h but not exported"
    assert declared - {"cloudy_last_error"} == set(_lib.SIGNATURES), "ctypes binding and header disagree"


def test_config_struct_layout_matches_header():
    """sizeof(cloudy_config) as the C compiler sees it == the ctypes mirror"""
    import subprocess, tempfile
    from cloudy_b200 import _lib
    src = '#include <stdio.h>\n#include "cloudy_b200.h"\nint main(){printf("%zu %zu %zu", sizeof(cloudy_config), ' \
          '__builtin_offsetof(cloudy_config, c), __builtin_offsetof(cloudy_config, dz));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_lib.cloudy_config)
    assert int(out[1]) == _lib.cloudy_config.c.offset
    assert int(out[2]) == _lib.cloudy_config.dz.offset


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import cloudy_b200 as cb
    with pytest.raises(cb.CloudyError):
        cb.Context(0)


def test_layout_helpers():
    import cloudy_b200 as cb
    npm = (2, 2, 3)
    assert cb.get_dist_moment_ind(npm, 1, 2) == 2 and cb.get_dist_moment_ind(npm, 2, 1) == 3 and cb.get_dist_moment_ind(npm, 3, 2) == 6
    for a in ((4, 2), (2, 0), (3, 4)):
        with pytest.raises(Exception):
            cb.get_dist_moment_ind(npm, *a)
    assert cb.get_dist_moments_ind_range(npm, 1) == range(1, 3) and cb.get_dist_moments_ind_range(npm, 3) == range(5, 8)
    nf = cb.get_moments_normalizing_factors(npm, (10.0, 0.1))
    assert np.allclose(nf, (10.0, 1.0, 10.0, 1.0, 10.0, 1.0, 0.1), atol=1e-12)
    assert cb.rflatten(((1, 2), (3.2, (1.2, 1.0)), (1,))) == (1, 2, 3.2, 1.2, 1.0, 1)


def test_kernel_tensors_host():
    """test_KernelTensors_correctness.jl:11-52"""
    import cloudy_b200 as cb
    ker = cb.CoalescenceTensor(np.array([[0.1, 0.0], [0.0, 0.2]]))
    assert np.array_equal(ker.c, [[0.1, 0.0], [0.0, 0.2]])
    ker = cb.CoalescenceTensor(lambda x, y: 0.02 + x + y, 1, 10.0)
    assert np.allclose(ker.c, [[0.02, 1.0], [1.0, 0.0]], rtol=1e-5, atol=1e-9)
    cb.check_symmetry(np.array([[1.0, -0.2, 0.1], [-0.2, -1.0, 1.1], [0.1, 1.1, 3.0]]))
    with pytest.raises(Exception):
        cb.check_symmetry(np.array([[1.0, 0.2, 0.1], [-0.2, -1.0, 1.1], [0.1, 1.1, 3.0]]))
    cb.check_symmetry(lambda x, y: x + y)
    with pytest.raises(Exception):
        cb.check_symmetry(lambda x, y: x - y)
    assert np.allclose(cb.polyfit(lambda x, y: 0.1 + 0.2 * x * y, 1, 10.0), [[0.1, 0.0], [0.0, 0.2]], rtol=1e-5, atol=1e-9)
    f = lambda x, y: 0.1 - 0.23 * x - 0.23 * y + 0.2 * x * y
    for lim in (10.0, 100.0, 1000.0):
        assert np.allclose(cb.polyfit(f, 1, lim), [[0.1, -0.23], [-0.23, 0.2]], rtol=1e-5)
    kn = cb.get_normalized_kernel_tensor(cb.CoalescenceTensor(np.array([[1.0, 2.0], [2.0, 3.0]])), (10.0, 0.2))
    assert np.allclose(kn.c, [[10.0, 4.0], [4.0, 1.2]], atol=1e-12)
    kt = cb.CoalescenceTensor(cb.LinearKernelFunction(5.0), 1, 5e-10)
    assert kt.c.shape == (2, 2) and abs(kt.c[0, 1] - 5.0) < 1e-8 and kt.c[0, 1] == kt.c[1, 0]
    with pytest.raises(Exception):
        cb.polyfit(lambda x, y: x + y, 1, 1.0, 2.0)


def test_coalescence_data_constructor():
    """Coalescence.jl:55-104"""
    import cloudy_b200 as cb
    ker = cb.CoalescenceTensor(np.array([[0.0, 5.0], [5.0, 0.0]]))
    cd = cb.CoalescenceData(ker, (3, 2), (5e-10, math.inf), (1e6, 1e-9))
    assert cd.N_mom_max == 4 and cd.N_2d_ints == (4, 3)
    assert cd.dist_thresholds[0] == 5e-10 / 1e-9 and math.isinf(cd.dist_thresholds[1])
    assert np.allclose(cd.kernels[0][1].c, [[0.0, 5e-3], [5e-3, 0.0]], rtol=1e-15)
    cdm = cb.CoalescenceData(ker, (3, 3), (0.99, 1.0), (1e6, 1e-9), cb.MovingThreshold())
    assert cdm.dist_thresholds == (0.99, 1.0)
    cfg = cb.build_config((cb.GAMMA, cb.EXPONENTIAL), cd)
    assert cfg.n_bins[0] == 75 and cfg.n_bins[1] == 0
    # the n_bins floor hazard: always 75 for thresholds <= 1
    rng = np.random.default_rng(0)
    for t in 10 ** rng.uniform(-12, 0, 2000):
        assert cb.log_grid(float(t))[0] == 75


def test_workload_generators_are_seeded():
    from cloudy_b200 import workloads as W
    a = W.c2_gamma_exp(1000)[1]
    b = W.c2_gamma_exp(1000)[1]
    assert np.array_equal(a, b) and a.shape == (1000, 5)
    assert abs((a == 0).all(axis=1).mean() - 0.02) < 0.02
    par, cols = W.c3_rainshaft(4, 32)
    assert cols.shape == (4, 32, 6) and par.dz == 3000.0 / 32


def test_julia_config_struct_mirrors_the_library_layout():
    """julia/CloudyB200.jl cannot run here, but its `struct CloudyConfig` can be read: field names, order and C layout (Int32 /
    Float64 scalars and NTuples) must equal what the library reports (cloudy_config_sizeof / cloudy_config_offsets) and what the
    ctypes mirror declares — the two bindings cannot drift apart."""
    from cloudy_b200 import _lib
    src = open(os.path.join(ROOT, "julia", "CloudyB200.jl")).read()
    consts = {"MAX_MODES": _lib.MAX_MODES, "MAX_P": _lib.MAX_P, "MAX_VEL": _lib.MAX_VEL}
    m = re.search(r"const MAX_MODES, MAX_P, MAX_VEL = (\d+), (\d+), (\d+)", src)
    assert m and tuple(int(g) for g in m.groups()) == (_lib.MAX_MODES, _lib.MAX_P, _lib.MAX_VEL)
    body = re.search(r"struct CloudyConfig\n(.*?)\nend", src, re.S).group(1)
    fields = []
    for line in body.splitlines():
        line = line.split("#")[0].strip()
        if not line:
            continue
        name, typ = line.split("::")
        mt = re.fullmatch(r"NTuple\{(.+),(Int32|Float64)\}", typ)
        count, base = (eval(mt.group(1), {}, consts), mt.group(2)) if mt else (1, typ)
        fields.append((name, count, {"Int32": 4, "Float64": 8}[base]))
    assert [f[0] for f in fields] == [n for n, _ in _lib.cloudy_config._fields_]
    offs, off = [], 0
    for _, count, size in fields:          # C layout rule: align every field to its element size
        off = (off + size - 1) // size * size
        offs.append(off)
        off += count * size
    total = (off + 7) // 8 * 8
    lib = _lib.load()
    got = (ctypes.c_int64 * 64)()
    n = lib.cloudy_config_offsets(got, 64)
    assert n == len(fields) and list(got[:n]) == offs
    assert lib.cloudy_config_sizeof() == total == ctypes.sizeof(_lib.cloudy_config)
