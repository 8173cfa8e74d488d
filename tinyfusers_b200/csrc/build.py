"""Builds libtinyfusers_b200.so in-tree with nvcc for sm_100a (and nothing else).

    python -m tinyfusers_b200.csrc.build [--force] [--verbose]

The .so is written next to the ctypes binding (tinyfusers_b200/native/b200/) so that it travels
with the repository snapshot to the GPU box; it is git-ignored.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT_DIR = os.path.join(os.path.dirname(HERE), "native", "b200")
LIB = os.path.join(OUT_DIR, "libtinyfusers_b200.so")
STAMP = LIB + ".stamp"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--use_fast_math",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
    "-I", HERE,
    "-DTF_BUILDING_LIB",
] + (["-DTF_GEMM_TRACE=1"] if os.environ.get("TF_GEMM_TRACE") == "1" else []) \
  + (["-DTF_ATT_TRACE=1"] if os.environ.get("TF_ATT_TRACE") == "1" else [])


def sources():
    return sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(ROOT, "include", "tinyfusers_b200.h"))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    h.update(b"cudart-shared-1")
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    objs = []
    obj_dir = os.path.join(HERE, "_obj")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        # the fp32 parity kernels need IEEE expf / tanhf / division: no --use_fast_math for that file
        flags = [f for f in FLAGS if f != "--use_fast_math"] if os.path.basename(src) == "tf_fp32.cu" else FLAGS
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("libtinyfusers_b200 build failed")
    # The CUDA runtime is linked DYNAMICALLY: the library then imports only the runtime symbols it calls (a statically linked
    # runtime drags every entry point's name into the artefact) and shares the process's libcudart.so.12 with torch. rpath: the
    # copy torch ships (already loaded whenever torch is) and the toolkit's, for a process that loads the library without torch.
    rpaths = [os.path.join(os.path.dirname(os.__file__), "site-packages", "nvidia", "cuda_runtime", "lib"),
              os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")]
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.cuda_runtime")
        if spec and spec.submodule_search_locations:
            rpaths.insert(0, os.path.join(list(spec.submodule_search_locations)[0], "lib"))
    except Exception:
        pass
    link = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared"]
    for r in rpaths:
        link += ["-Xlinker", "-rpath", "-Xlinker", r]
    subprocess.check_call(link)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
