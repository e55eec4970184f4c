"""In-tree build of libcphnsw_b200.so (sm_100a only) with plain nvcc.

    python rabitq-ann-search_b200/build.py [--force]

-fmad=false: every fused multiply-add in the kernels is written explicitly (the float sequences
must match the reference's AVX2 / GCC-contracted code bit for bit); -lineinfo for ncu's source
page.  The library links the CUDA runtime statically and has no other dependency (no torch).
"""
from __future__ import annotations

import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "cphnsw_b200" / "libcphnsw_b200.so"
OBJ = HERE / "build"
SOURCES = ["capi.cu", "query_prep.cu", "fastscan_blocks.cu", "search.cu", "relayout.cu", "exhaustive.cu", "exhaustive_tc.cu", "exhaustive_tc16.cu", "postprocess.cu", "neighbor_codes.cu", "calibration.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-Xptxas", "-v",
]


def _newest_dep() -> float:
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "cphnsw_b200.h", Path(__file__)]
    return max(p.stat().st_mtime for p in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines: tuple = ()) -> Path:
    """variant / defines: an A/B build of the same library with extra -D flags, written to
    cphnsw_b200/variants/libcphnsw_b200_<variant>.so (load it with CPHNSW_B200_LIB=...)."""
    global OUT, OBJ
    if variant:
        OUT = HERE / "cphnsw_b200" / "variants" / f"libcphnsw_b200_{variant}.so"
        OBJ = HERE / "build" / f"variant_{variant}"
        OUT.parent.mkdir(exist_ok=True)
        force = True
    newest = _newest_dep()
    if not force and OUT.exists() and OUT.stat().st_mtime >= newest:
        return OUT
    OBJ.mkdir(exist_ok=True, parents=True)

    def compile_one(src: str):
        obj = OBJ / (src + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(OUT), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    if "--variant" in sys.argv:   # build.py --variant NAME [DEFINE[=VALUE] ...]
        i = sys.argv.index("--variant")
        print(build(variant=sys.argv[i + 1], defines=tuple(sys.argv[i + 2:])))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
