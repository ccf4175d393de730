"""Build libhvc_sm100a.so (hand-written sm_100a CUDA + the C ABI of include/hvc.h) in-tree.

    python -m hybrid_vit_cascade_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libhvc_sm100a.so")
OBJ = os.path.join(PKG, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", os.path.join(PKG, "..", "include")]
# --use_fast_math (approximate division / sqrt / transcendentals, flush-to-zero) only where bf16 operands bound the accuracy anyway.
# The fp32 verification kernels (the 1e-4 bar), the losses (compared with the reference to 2e-5) and the optimizer keep IEEE arithmetic.
PRECISE = {"hvc_fp32.cu", "hvc_loss.cu", "hvc_loss_multiscale.cu", "hvc_optim.cu"}
FLAGS += os.environ.get("HVC_EXTRA_NVCC_FLAGS", "").split()     # bring-up only, e.g. -DHVC_TUNE_FWD_EMU


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/hvc.h"]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update((" ".join(FLAGS) + " precise:" + ",".join(sorted(PRECISE))).encode())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC] + FLAGS + ([] if src in PRECISE else ["--use_fast_math"]) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(OBJ, src[:-3] + ".log"), "w") as f:
        f.write(log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{log}")
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(OUT) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return OUT
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), _sources()))
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
