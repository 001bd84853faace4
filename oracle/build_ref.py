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
