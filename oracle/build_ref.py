"""Recipe that compiles the reference's own Cython CPU NMS into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY -- nothing under ``oracle/`` is on the product path.

The reference ships ``ext/nms/nms/cpu_nms.pyx`` (hard NMS ``cpu_nms`` :122-173 and
Gaussian soft-NMS ``cpu_soft_nms`` :17-120).  It does not compile against numpy 2 /
Cython 3 as it lies (``np.int_t`` :130,:133, ``dtype=np.int`` :134, ``np.float thresh``
:122 were removed from numpy), so the recipe streams the file from
``/root/reference`` through exactly three token substitutions (``np.int_t`` -> ``np.intp_t``,
``dtype=np.int)`` -> ``dtype=np.intp)``, ``np.float thresh`` -> ``double thresh``: the removed
``np.float`` alias was Python's float, i.e. a double) into the git-ignored
``oracle/_ref/`` directory and cythonizes it there.  No reference source is ever
written to a tracked path.

The resulting ``oracle/_ref/cpu_nms*.so`` travels to the GPU box with the snapshot
(``oracle/_ref/`` is git-ignored but not gpurun-ignored), where it is used as
  * the checker for the legacy ("+1") NMS / soft-NMS semantics, and
  * the ``ext/nms`` leg of the CPU baseline in ``bench.py``.

The reference's CUDA hard NMS (``ext/nms/nms/nms_kernel.cu:34-144``: ``nms_kernel`` + the ``_nms`` host driver)
is compiled too, UNCHANGED and straight from ``/root/reference``, for sm_100a (the reference's ``setup.py:132``
pins ``-arch=sm_35``, which nvcc 12.9 rejects) into ``oracle/_ref/libref_gpu_nms.so``.  ``_nms`` is a C++ symbol
(``gpu_nms.hpp:1-2`` has no extern "C"): ``load_gpu_nms()`` binds its mangled name.  It is the "reference CUDA
kernel" baseline that ``bench.py`` times next to ``rr_nms_legacy_host`` (same signature) on the GPU box.

Usage:  python oracle/build_ref.py            (no-op when /root/reference is absent)
"""
import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/ext/nms/nms/cpu_nms.pyx"

_SUBS = (
    (r"np\.int_t", "np.intp_t"),
    (r"dtype=np\.int\)", "dtype=np.intp)"),
    (r"np\.float thresh", "double thresh"),
)

_SETUP = """
import numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
setup(name="cpu_nms",
      ext_modules=cythonize(
          [Extension("cpu_nms", ["cpu_nms.pyx"], include_dirs=[numpy.get_include()],
                     extra_compile_args=["-O2", "-w"])],
          language_level=2, quiet=True),
      script_args=["build_ext", "--inplace", "-q"])
"""


def built():
    return bool(glob.glob(os.path.join(REF_DIR, "cpu_nms*.so")))


def build(force=False):
    """Build oracle/_ref/cpu_nms*.so.  Returns True when the module is available."""
    if built() and not force:
        return True
    if not os.path.exists(REF_SRC):
        return False
    os.makedirs(REF_DIR, exist_ok=True)
    with open(REF_SRC, "r", encoding="utf-8") as f:
        text = f.read()
    for pat, rep in _SUBS:
        text = re.sub(pat, rep, text)
    with open(os.path.join(REF_DIR, "cpu_nms.pyx"), "w", encoding="utf-8") as f:
        f.write(text)
    with open(os.path.join(REF_DIR, "_setup.py"), "w") as f:
        f.write(_SETUP)
    subprocess.run([sys.executable, "_setup.py"], cwd=REF_DIR, check=True,
                   stdout=subprocess.DEVNULL)
    return built()


REF_CU = "/root/reference/ext/nms/nms/nms_kernel.cu"
GPU_NMS_SO = os.path.join(REF_DIR, "libref_gpu_nms.so")
_NMS_MANGLED = "_Z4_nmsPiS_PKfiifi"          # void _nms(int*, int*, const float*, int, int, float, int)


def build_gpu_nms(force=False):
    """nvcc -arch=sm_100a on the reference's nms_kernel.cu where it lies -> oracle/_ref/libref_gpu_nms.so."""
    if os.path.exists(GPU_NMS_SO) and not force:
        return True
    if not os.path.exists(REF_CU):
        return False
    os.makedirs(REF_DIR, exist_ok=True)
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    cmd = [nvcc] + (["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []) + [
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC", "-w",
        "-I", os.path.dirname(REF_CU), REF_CU, "-o", GPU_NMS_SO, "-lcudart"]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(GPU_NMS_SO)


def load_gpu_nms():
    """-> callable(dets_sorted float32 [n,5] numpy, thresh, device_id=0) -> int32 keep indices (into the sorted rows),
    i.e. the reference's `_nms` (nms_kernel.cu:91-144) with its own malloc / H2D / mask D2H / CPU reduce, or None."""
    if not os.path.exists(GPU_NMS_SO):
        return None
    import ctypes
    import numpy as np
    lib = ctypes.CDLL(GPU_NMS_SO)
    fn = getattr(lib, _NMS_MANGLED)
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                   ctypes.c_int]

    def _nms(dets_sorted, thresh, device_id=0):
        d = np.ascontiguousarray(dets_sorted, dtype=np.float32)
        keep = np.zeros(d.shape[0], dtype=np.int32)
        num = np.zeros(1, dtype=np.int32)
        fn(keep.ctypes.data, num.ctypes.data, d.ctypes.data, d.shape[0], d.shape[1], float(thresh), int(device_id))
        return keep[: int(num[0])]
    return _nms


def load():
    """Import the compiled reference module (``cpu_nms``, ``cpu_soft_nms``) or None."""
    if not built():
        return None
    import importlib.util
    path = sorted(glob.glob(os.path.join(REF_DIR, "cpu_nms*.so")))[0]
    spec = importlib.util.spec_from_file_location("cpu_nms", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref cpu_nms:", "built" if ok else "unavailable (no /root/reference)")
    ok = build_gpu_nms(force="--force" in sys.argv)
    print("oracle/_ref libref_gpu_nms.so:", "built" if ok else "unavailable (no /root/reference)")
