"""Builds libgsa_cuda.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and by hand:
    python globalsensitivity.jl_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgsa_cuda.so")
TABLE = os.path.join(HERE, "data", "joe_kuo_21201.bin")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOST_CXX = "/usr/bin/g++"


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    cu = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp"))]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "gsa_cuda.h"), TABLE]
    return cu, deps


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in sources()[1])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cu, _ = sources()
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    common = [nvcc(), "-ccbin", HOST_CXX, "-std=c++17", "-O3", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "-Xptxas", "-v" if verbose else "-O3", "--fmad=true"]
    objs = []
    procs = []
    for src in cu:
        o = os.path.join(obj_dir, os.path.basename(src) + ".o")
        objs.append(o)
        cmd = common + ["-x", "cu", "-c", src, "-o", o]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    blob = os.path.join(obj_dir, "directions_blob.o")
    subprocess.check_call(["/usr/bin/gcc", "-c", os.path.join(CSRC, "directions_blob.S"), f'-DGSA_TABLE_PATH="{TABLE}"', "-o", blob])
    # the fused kernels with an UNRESOLVED user model, as relocatable cubins that gsa_model_register_* link at run time:
    # generic (gsa_user_model), product form (gsa_user_factor), additive (gsa_user_term)
    user_procs, user_blobs = [], []
    for tag, src_in, defs in (("kernel", "user_kernel.cu.in", []), ("product", "user_kernel_separable.cu.in", ["-DGSA_USER_KIND=1"]),
                              ("additive", "user_kernel_separable.cu.in", ["-DGSA_USER_KIND=2"])):
        ucu = os.path.join(obj_dir, f"user_{tag}.cu")
        shutil.copyfile(os.path.join(CSRC, src_in), ucu)
        ucubin = os.path.join(obj_dir, f"user_{tag}.cubin")
        cmd = [nvcc(), "-ccbin", HOST_CXX, "-std=c++17", "-O3", "-lineinfo", *ARCH, "-rdc=true", "-cubin", "-I", CSRC, *defs, "-o", ucubin, ucu]
        user_procs.append((ucu, ucubin, tag, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for ucu, ucubin, tag, p in user_procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {ucu}")
        ublob = os.path.join(obj_dir, f"user_blob_{tag}.o")
        sym = f"gsa_user_{tag}_blob"
        subprocess.check_call(["/usr/bin/gcc", "-c", os.path.join(CSRC, "user_blob.S"), f"-DGSA_BLOB_SYM={sym}", f"-DGSA_BLOB_END={sym}_end",
                               f'-DGSA_USER_CUBIN_PATH="{ucubin}"', "-o", ublob])
        user_blobs.append(ublob)
    # ... and the separable kernels once more as LTO-IR: when the user's image carries LTO-IR too (code=lto_100a) and uses the
    # default symbol name, nvJitLink inlines the per-coordinate function into the kernel that is actually launched (-kernels-used)
    lto_procs = []
    for tag, kind in (("kernel", 0), ("product", 1), ("additive", 2)):
        ucu = os.path.join(obj_dir, f"user_{tag}.cu")
        ltoir = os.path.join(obj_dir, f"user_{tag}.ltoir")
        cmd = [nvcc(), "-ccbin", HOST_CXX, "-std=c++17", "-O3", "-dlto", "-ltoir", "-arch=sm_100a", "-rdc=true", "-I", CSRC,
               f"-DGSA_USER_KIND={kind}", "-DGSA_USER_LTO_ONLY=1", "-o", ltoir, ucu]
        lto_procs.append((ucu, ltoir, tag, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for ucu, ltoir, tag, p in lto_procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc -ltoir failed on {ucu}")
        ublob = os.path.join(obj_dir, f"user_blob_{tag}_lto.o")
        sym = f"gsa_user_{tag}_lto_blob"
        subprocess.check_call(["/usr/bin/gcc", "-c", os.path.join(CSRC, "user_blob.S"), f"-DGSA_BLOB_SYM={sym}", f"-DGSA_BLOB_END={sym}_end",
                               f'-DGSA_USER_CUBIN_PATH="{ltoir}"', "-o", ublob])
        user_blobs.append(ublob)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc(), "-ccbin", HOST_CXX, "-shared", *ARCH, "-o", OUT, *objs, blob, *user_blobs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.check_call(link)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
