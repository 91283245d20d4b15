"""Builds rails_b200/lib/libmol_b200.so (sm_100a) with nvcc, in-tree.

    python -m rails_b200.build          # incremental
    python -m rails_b200.build --force

No torch headers are involved: the library is a plain C-ABI CUDA shared object (include/mol_b200.h)
that the Python side loads with ctypes.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(PKG, "build")
LIB = os.path.join(OUT_DIR, "libmol_b200.so")
PROBE = os.path.join(OUT_DIR, "libmol_probe.so")

SOURCES = ["mol_api.cu", "mol_prologue.cu", "mol_exact.cu", "mol_select.cu", "mol_coarse_sm100.cu", "mol_extras.cu",
           "mol_dotfilter_sm100.cu", "mol_linear_x3_sm100.cu"]
PROBE_SOURCES = ["probe_sm100.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "mol_b200.h"))
    return hs


def _compile(src: str, force: bool) -> str:
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    if force or _stale(obj, [path] + _headers()):
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def _link(objs, target: str) -> None:
    cmd = [_nvcc(), "-shared", "-o", target] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    all_src = SOURCES + [s for s in PROBE_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with ThreadPoolExecutor(max_workers=min(8, len(all_src))) as ex:
        objs = dict(zip(all_src, ex.map(lambda s: _compile(s, force), all_src)))
    main_objs = [objs[s] for s in SOURCES]
    if force or _stale(LIB, main_objs):
        _link(main_objs, LIB)
    probe_objs = [objs[s] for s in PROBE_SOURCES if s in objs]
    if probe_objs and (force or _stale(PROBE, probe_objs)):
        _link(probe_objs, PROBE)
    if verbose:
        print("built", LIB)
    return LIB


def build_variant(name: str, defs) -> str:
    """Tuning build: compiles every source with extra -D flags into lib/libmol_b200_<name>.so (select it with
    MOL_B200_LIB=<path>).  Objects go to build/<name>/ so the default build is untouched."""
    os.makedirs(OUT_DIR, exist_ok=True)
    odir = os.path.join(OBJ_DIR, name)
    os.makedirs(odir, exist_ok=True)
    objs = []

    def one(src):
        obj = os.path.join(odir, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defs] + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(one, SOURCES))
    target = os.path.join(OUT_DIR, f"libmol_b200_{name}.so")
    _link(objs, target)
    shutil.rmtree(odir, ignore_errors=True)  # the objects are not needed again (and would travel with every gpurun snapshot)
    return target


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print("built", build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        build(force="--force" in sys.argv, verbose=True)
