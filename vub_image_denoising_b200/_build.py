"""In-tree build of libb200dn.so (sm_100a only) with plain nvcc.

The shared library is the product: a C-ABI (include/b200dn.h) with no torch types in any signature.
It is built next to this file so that it travels with the repo snapshot to the GPU box; nothing is
JIT-compiled at import time and there is no fallback if it is missing.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libb200dn.so"
BUILD_DIR = PKG_DIR / "build"
STAMP = BUILD_DIR / "sources.sha256"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libb200dn.so cannot be built")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [PKG_DIR.parent / "include" / "b200dn.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    return LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == _digest()


def build_lib(force: bool = False, verbose: bool = False, variant: str = "", defines: tuple = ()) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libb200dn.so. Returns the library path.

    variant / defines: a diagnostics build next to the product library (`libb200dn_<variant>.so`, objects under
    build/<variant>/), e.g. variant="timeline", defines=("B200DN_TIMELINE",) for tools/launch_timeline.py; it is
    loaded only when B200DN_LIB points at it."""
    if variant:
        return _build(PKG_DIR / f"libb200dn_{variant}.so", BUILD_DIR / variant, [f"-D{d}" for d in defines], None, verbose)
    if not force and is_fresh():
        return LIB_PATH
    return _build(LIB_PATH, BUILD_DIR, [], STAMP, verbose)


def _build(lib_path: Path, build_dir: Path, extra: list, stamp, verbose: bool) -> Path:
    nvcc = _nvcc()
    build_dir.mkdir(parents=True, exist_ok=True)
    logs: dict[str, str] = {}

    def compile_one(src: Path) -> Path:
        obj = build_dir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src.name] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib_path), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (build_dir / "ptxas.log").write_text("\n".join(f"== {k}\n{v}" for k, v in sorted(logs.items())))
    if stamp is not None:
        stamp.write_text(_digest())
    if verbose:
        print((build_dir / "ptxas.log").read_text(), file=sys.stderr)
    return lib_path


if __name__ == "__main__":
    if "--timeline" in sys.argv:
        print(build_lib(variant="timeline", defines=("B200DN_TIMELINE",), verbose="-v" in sys.argv))
    else:
        print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
