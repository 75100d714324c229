"""Build recipe for libmvae_b200.so (plain nvcc, sm_100a only, in-tree so the .so travels with gpurun)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmvae_b200.so")
SOURCES = ["cfgb.cu", "umma_gemm.cu", "gru_rec.cu", "gru_rec2.cu", "optim.cu", "moses.cu", "cfga.cu", "binding.cu", "text.cu", "errors.cu", "probe.cu", "decode_persist.cu", "umma_gemm2.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false" if False else "-Xptxas=-v",
]


OBJ_DIR = os.path.join(HERE, "_build")
COMPILE_FLAGS = [f for f in NVCC_FLAGS if f != "-shared"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")]
    return hs + [os.path.join(HERE, "..", "include", "mvae_b200.h"), os.path.abspath(__file__)]


def _stale_objects():
    """Sources whose object file is older than the source or than any header (objects live in _build/, git-ignored)."""
    hdr_t = max(os.path.getmtime(p) for p in _headers())
    stale = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(os.path.join(CSRC, src))):
            stale.append(src)
    return stale


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mvae_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    """One nvcc -c per translation unit (changed ones only, in parallel), then one link into the in-tree .so."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo = list(SOURCES) if force else _stale_objects()
    procs = []
    for src in todo:
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        cmd = [nvcc] + COMPILE_FLAGS + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libmvae_b200.so")
    objs = [os.path.join(OBJ_DIR, s[:-3] + ".o") for s in SOURCES]
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking libmvae_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
