"""Build librrnet_b200.so in-tree with plain nvcc for sm_100a (no torch extension machinery).

    python -m rrnet_b200.build [--force] [--verbose]

Each translation unit is compiled to an object with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 ...
and the objects are linked into rrnet_b200/librrnet_b200.so (git-ignored, travels to the GPU
box with the snapshot).  Kernels whose results must round exactly like the reference's
separate fp32 ops are compiled with --fmad=false (FMA where wanted is written as fmaf()).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "librrnet_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

# translation unit -> extra flags
UNITS = {
    "rr_api.cu": [],
    "rr_decode.cu": ["--fmad=false"],
    "rr_tail.cu": [],
    "rr_nms.cu": ["--fmad=false"],
    "rr_softnms.cu": ["--fmad=false"],
    "rr_roialign.cu": ["--fmad=false"],
    "rr_head.cu": [],
    "rr_head_tc.cu": [],
    "rr_bbox.cu": ["--fmad=false"],
    "rr_render.cu": ["--fmad=false"],
    "rr_focal.cu": [],
    "rr_regl1.cu": ["--fmad=false"],
    "rr_stage2loss.cu": [],
    "rr_apmatch.cu": ["--fmad=false"],
}

BASE = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", INCLUDE]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _host_cc():
    # /opt/gcc in this image lacks some spec files; the system compiler is complete
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None


def _digest(paths, flags):
    h = hashlib.sha1()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "rr_common.cuh"), os.path.join(CSRC, "rr_gauss.cuh"), os.path.join(CSRC, "rr_head.cuh"), os.path.join(CSRC, "rr_decode.cuh"), os.path.join(INCLUDE, "rrnet_b200.h")]
    nvcc = _nvcc()
    ccbin = _host_cc()
    objs, rebuilt = [], False
    procs = []
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        stamp = obj + ".sha1"
        flags = BASE + extra + (["-Xptxas", "-v"] if verbose else [])
        dig = _digest([src] + headers, flags)
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(stamp)
                and open(stamp).read() == dig):
            continue
        cmd = [nvcc] + (["-ccbin", ccbin] if ccbin else []) + flags + ["-c", src, "-o", obj]
        procs.append((unit, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        rebuilt = True
    for unit, stamp, dig, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (unit, out))
        if verbose and out.strip():
            print("== %s\n%s" % (unit, out))
        with open(stamp, "w") as f:
            f.write(dig)
    if rebuilt or force or not os.path.exists(LIB):
        cmd = [nvcc] + (["-ccbin", ccbin] if ccbin else []) + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout.decode())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
