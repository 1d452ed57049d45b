"""In-tree build of libtrikb200.so (CUDA kernels + C ABI) and the five alias libraries.

nvcc cross-compiles sm_100a without a GPU; the resulting .so files sit next to this file so that
they travel to the GPU box with the repository snapshot (they are git-ignored).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libtrikb200.so")
STAMP = os.path.join(HERE, ".libtrikb200.stamp")

SOURCES = ["trik_kernels.cu", "trik_kernels_line.cu", "trik_kernels_lut.cu", "trik_kernels_anneal.cu", "trik_kernels_grid.cu", "trik_kernels_omtab.cu", "trik_kernels_ommaj.cu", "trik_kernels_ingest.cu", "trik_kernels_edge.cu", "trik_kernels_detect.cu", "trik_kernels_preview.cu", "trik_capi.cu", "trik_host.cpp"]
HEADERS = ["trik_kernels.cuh", "trik_pixel.cuh", "trik_line.cuh", "trik_lut.cuh", "trik_host.hpp"]
KINDS = ("wo", "wl", "oo", "ol", "om")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    for name in ("trik_b200.h", "trik_xdm.h"):
        with open(os.path.join(INCLUDE, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def alias_path(kind):
    return os.path.join(HERE, "libtrik_vidtranscode_cv_%s.so" % kind)


def build(force=False, verbose=False):
    """Compile libtrikb200.so if the sources changed.  Returns the library path."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest and all(os.path.exists(alias_path(k)) for k in KINDS):
                return LIB
    # one object per source, compiled in parallel and only when that source (or a header / flag) changed, then one link
    objdir = os.path.join(HERE, ".obj")
    os.makedirs(objdir, exist_ok=True)
    hdr = hashlib.sha256()
    for name in HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            hdr.update(f.read())
    for name in ("trik_b200.h", "trik_xdm.h"):
        with open(os.path.join(INCLUDE, name), "rb") as f:
            hdr.update(f.read())
    hdr.update(" ".join(NVCC_FLAGS).encode())

    def compile_one(src):
        with open(os.path.join(CSRC, src), "rb") as f:
            key = hashlib.sha256(hdr.digest() + f.read()).hexdigest()
        obj = os.path.join(objdir, src + ".o")
        keyfile = obj + ".key"
        if not force and os.path.exists(obj) and os.path.exists(keyfile):
            with open(keyfile) as f:
                if f.read().strip() == key:
                    return obj
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", "-I", INCLUDE, "-I", CSRC, "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        with open(keyfile, "w") as f:
            f.write(key)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.run([_nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB] + objs, check=True)
    # alias libraries: re-export one sensor's tables under the reference's symbol names
    # (TRIK_VIDTRANSCODE_CV_FXNS / TRIK_VIDTRANSCODE_CV_IALG, <sensor>/trik_vidtranscode_cv.h:17-18)
    for kind in KINDS:
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-I", INCLUDE,
                        "-DTRIKB200_ALIAS_KIND=%s" % kind.upper(),
                        "-o", alias_path(kind), os.path.join(CSRC, "trik_alias.c"),
                        "-L", HERE, "-ltrikb200", "-Wl,-rpath,$ORIGIN"], check=True)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
