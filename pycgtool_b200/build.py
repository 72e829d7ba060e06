"""Build ``libpycgtool_b200.so`` in-tree with nvcc for sm_100a.

The shared library is the product: there is no CPU fallback.  The build is a
plain ``nvcc -shared`` over the three ``.cu`` files (nvcc cross-compiles on a
machine without a GPU); the result stays next to the sources so that it
travels with the repository snapshot to the GPU box.
"""

import hashlib
import os
import pathlib
import shutil
import subprocess

CSRC = pathlib.Path(__file__).resolve().parent / "csrc"
SOURCES = ["cgb_map.cu", "cgb_measure.cu", "cgb_measure_fused.cu", "cgb_map_measure.cu", "cgb_boltzmann.cu",
           "cgb_xtc.cu", "cgb_comm.cu"]
HEADERS = [CSRC / "cgb_common.cuh", CSRC / "cgb_terms.cuh", CSRC / "cgb_termwarp.cuh", CSRC / "cgb_mapwarp.cuh",
           CSRC / "cgb_measure_plan.cpp.inc",
           CSRC.parent.parent / "include" / "pycgtool_b200.h"]
LIBRARY = CSRC / "libpycgtool_b200.so"
STAMP = CSRC / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-ldl", "--threads", "0",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libpycgtool_b200.so cannot be built")
    return exe


def _digest():
    h = hashlib.sha256()
    for path in [CSRC / s for s in SOURCES] + HEADERS:
        h.update(path.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """Compile the library if the sources changed since the last build."""
    digest = _digest()
    if not force and LIBRARY.exists() and STAMP.exists() and STAMP.read_text() == digest:
        return LIBRARY
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIBRARY)] + SOURCES
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    STAMP.write_text(digest)
    return LIBRARY


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
