"""In-tree build of libupnerf_b200.so (hand-written sm_100a CUDA behind a C ABI).

`python -m upnerf_b200.build` or `__graft_entry__.build()` compiles every .cu under
`upnerf_b200/csrc/` with nvcc for sm_100a only and links one shared library next to the
package, so the built artefact travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = ROOT / "build" / "obj"
LIB = PKG / "lib" / "libupnerf_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA path cannot be built")


def _newer(a: Path, b: Path) -> bool:
    return (not b.exists()) or a.stat().st_mtime > b.stat().st_mtime


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(parents=True, exist_ok=True)
    LIB.parent.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))
    hdr_mtime = max(h.stat().st_mtime for h in headers)

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_mtime):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log, file=sys.stderr)
    if force or any(_newer(o, LIB) for o in objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
