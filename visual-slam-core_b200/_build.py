"""Build the CUDA library in-tree with nvcc for sm_100a (no torch extension, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libb200ba.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
SOURCES = ["ba_local.cu", "ba_tiled.cu", "ba_stream.cu", "ba_mma.cu", "ba_panel.cu", "ba_sparse.cu", "ba_band.cu", "ba_bcr.cu", "ba_pose.cu", "ba_ransac.cu", "ba_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps: List[str] = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "b200ba.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, out: str = "", extra: List[str] = ()) -> str:
    """Compile every .cu under csrc/ and link libb200ba.so next to this file.  `out` / `extra` build a diagnostic
    variant (e.g. extra=["-DBA_PANEL_TRACE"]) under another name; `BA_LIB_PATH` makes `_lib.load()` pick it up."""
    if out:
        return _build_to(out, os.path.join(HERE, "build", "variant_" + os.path.basename(out)), list(extra), verbose)
    if not force and not _stale():
        return LIB_PATH
    return _build_to(LIB_PATH, os.path.join(HERE, "build"), list(extra), verbose)


def _build_to(lib_path: str, objdir: str, extra: List[str], verbose: bool) -> str:
    nvcc = _nvcc()
    os.makedirs(objdir, exist_ok=True)
    objs = []
    log = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("BA_NVCC_EXTRA", "").split(), *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}")
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(objdir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return lib_path


# ---------------------------------------------------------------------------------------------
# host-side helper: the packer's observation loop as a CPython extension (no CUDA; optional)
# ---------------------------------------------------------------------------------------------
FASTPACK_SRC = os.path.join(CSRC, "fastpack.c")


def fastpack_path() -> str:
    import sysconfig
    return os.path.join(HERE, "_fastpack" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_fastpack(force: bool = False) -> str:
    """gcc -shared of csrc/fastpack.c next to this file.  Raises when there is no compiler or no Python.h;
    packing.py treats the extension as optional and keeps its Python loop."""
    import sysconfig
    out = fastpack_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(FASTPACK_SRC):
        return out
    cc = os.environ.get("CC") or shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("no C compiler for the packer extension")
    inc = sysconfig.get_paths()["include"]
    cmd = [cc, "-O2", "-fPIC", "-shared", "-Wall", "-I", inc, FASTPACK_SRC, "-o", out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"packer extension failed to build:\n{r.stdout}")
    return out
