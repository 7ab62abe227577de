"""Build recipe for libcloudy_b200.so (nvcc, sm_100a, in-tree; translation units compiled in parallel)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
UNITS = ["cloudy_b200.cu"] + [f"tpp_inst_{g}.cu" for g in "ABCDEFGH"]
HEADERS = [os.path.join(CSRC, h) for h in ("special.cuh", "common.cuh", "tpp_kernel.cuh", "tpp_instances.inc")] + \
          [os.path.join(os.path.dirname(HERE), "include", "cloudy_b200.h")]
# development variants: CLOUDY_DEV=<tag> builds libcloudy_b200_<tag>.so with the BASELINE shapes only (-DTPP_DEV_SHAPES) plus
# the flags in CLOUDY_DEV_FLAGS; load it with CLOUDY_LIB=<path> (cloudy.jl_b200/_lib.py)
DEV = os.environ.get("CLOUDY_DEV", "")
LIB = os.path.join(HERE, f"libcloudy_b200_{DEV}.so" if DEV else "libcloudy_b200.so")
OBJDIR = os.path.join(CSRC, f"build_{DEV}" if DEV else "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _obj(unit):
    return os.path.join(OBJDIR, unit.replace(".cu", ".o"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(unit):
    src = os.path.join(CSRC, unit)
    dev = (["-DTPP_DEV_SHAPES"] + os.environ.get("CLOUDY_DEV_FLAGS", "").split()) if DEV else []
    cmd = [_nvcc()] + NVCC_FLAGS + dev + ["-c", "-o", _obj(unit), src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return unit, " ".join(cmd), res.returncode, res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [u for u in UNITS if force or _stale(_obj(u), [os.path.join(CSRC, u)] + HEADERS)]
    logs = []
    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            for unit, cmd, rc, log in ex.map(_compile, todo):
                logs.append(cmd + "\n" + log)
                with open(os.path.join(OBJDIR, unit + ".log"), "w") as f:
                    f.write(cmd + "\n" + log)
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {unit}:\n" + log[-6000:])
    if todo or _stale(LIB, [_obj(u) for u in UNITS]):
        cmd = [_nvcc(), "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [_obj(u) for u in UNITS]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
