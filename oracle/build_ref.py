"""Recipe for oracle/_ref: compiles the reference's own softmax kernel (the only hot-path source of the reference that
builds without CuPy / cuDNN bindings) from /root/reference into oracle/_ref/libref_softmax.so.

    python oracle/build_ref.py        # in the authoring container (the GPU box only uses the prebuilt file)

Test infrastructure: tests/test_ref_softmax_gpu.py holds the oracle's softmax_rows, the fp32 parity kernel and the fused
attention kernel to what this library computes on the GPU. oracle/_ref/ is git-ignored and travels with the snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CUDA = "/root/reference/tinyfusers/native/cuda"
OUT = os.path.join(HERE, "_ref", "libref_softmax.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build(force=False):
    src = os.path.join(REF_CUDA, "softmax.cu")
    if not os.path.exists(src):
        return OUT if os.path.exists(OUT) else None      # GPU box: prebuilt file or nothing
    launcher = os.path.join(HERE, "ref_softmax_launcher.cu")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(src), os.path.getmtime(launcher)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--use_fast_math", "-D__CUDA_NO_HALF_CONVERSIONS__",
           f"-I{REF_CUDA}", f'-DREF_SOFTMAX_CU="{src}"', "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
           "-Xlinker", "-rpath", "-Xlinker", os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64"),
           "-o", OUT, launcher]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
