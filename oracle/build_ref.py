#!/usr/bin/env python
"""Build the UNMODIFIED reference (wwang721/pyafv) into ``oracle/_ref/`` as binaries only.

TEST / BASELINE INFRASTRUCTURE -- never imported by the product package ``pyafv_b200``.

What this does (network-free; the reference's own build backend hatchling/hatch-cython is not
installed in this image):

  1. copy ``/root/reference/pyafv`` to a scratch directory under ``/tmp`` (the mount is read-only),
  2. cythonize ``pyafv/cell_geom.pyx`` -> C++ -> ``cell_geom.*.so`` with the directives the reference
     declares in ``pyproject.toml:72`` (boundscheck=False, wraparound=False, language_level=3),
  3. byte-compile the pure-Python modules to sourceless ``*.pyc``,
  4. install ONLY the compiled outputs to ``oracle/_ref/``: ``pyafv_pyc.zip`` (the ``*.pyc`` files; the gpurun
     snapshot drops bare ``*.pyc`` files, a zip travels) and ``cell_geom.*.so``.  ``import_reference()`` unpacks
     both into a scratch directory and imports ``pyafv`` from there.

No reference source file is written into this repository: ``oracle/_ref/`` is git-ignored, holds
binaries only, and travels to the GPU box with the gpurun snapshot so that ``bench.py --impl
reference`` and ``cpu_baseline`` can time the real reference there.  ``/root/reference`` does not
exist on the GPU box, so this script is a no-op there when ``oracle/_ref`` is already populated.

Usage:  python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import compileall
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

REF = Path(os.environ.get("AFV_REFERENCE_DIR", "/root/reference"))
HERE = Path(__file__).resolve().parent
DEST = HERE / "_ref"

_SETUP = r'''
import numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
ext = Extension("pyafv.cell_geom", ["pyafv/cell_geom.pyx"],
                include_dirs=[numpy.get_include()], language="c++",
                extra_compile_args=["-O2"],
                define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")])
setup(name="pyafv_ref", ext_modules=cythonize(
    [ext], compiler_directives=dict(boundscheck=False, wraparound=False, language_level=3)))
'''


def ref_available() -> bool:
    """True when oracle/_ref holds an importable build of the reference."""
    return (DEST / "pyafv_pyc.zip").exists() and any(DEST.glob("cell_geom*.so"))


def build(force: bool = False) -> bool:
    if ref_available() and not force:
        return True
    if not (REF / "pyafv" / "cell_geom.pyx").exists():
        return False  # GPU box, or reference not mounted: nothing to build from
    with tempfile.TemporaryDirectory(prefix="afv_ref_build_") as tmp:
        tmp = Path(tmp)
        shutil.copytree(REF / "pyafv", tmp / "pyafv")
        (tmp / "setup.py").write_text(_SETUP)
        subprocess.run([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL)
        # sourceless byte-code: module.pyc next to where module.py was
        compileall.compile_dir(str(tmp / "pyafv"), quiet=1, legacy=True, force=True)
        if DEST.exists():
            shutil.rmtree(DEST)
        DEST.mkdir(parents=True)
        import zipfile
        with zipfile.ZipFile(DEST / "pyafv_pyc.zip", "w", zipfile.ZIP_DEFLATED) as z:
            for src in sorted((tmp / "pyafv").rglob("*.pyc")):
                if "__pycache__" not in src.parts:
                    z.write(src, str(src.relative_to(tmp)))
        for src in (tmp / "pyafv").glob("cell_geom*.so"):
            shutil.copy2(src, DEST / src.name)
    return ref_available()


_UNPACKED = None


def _unpack() -> Path:
    """Unpack oracle/_ref into a scratch directory laid out as an importable ``pyafv`` package."""
    global _UNPACKED
    if _UNPACKED is None:
        import atexit
        import zipfile
        d = Path(tempfile.mkdtemp(prefix="afv_ref_"))
        atexit.register(shutil.rmtree, d, ignore_errors=True)
        with zipfile.ZipFile(DEST / "pyafv_pyc.zip") as z:
            z.extractall(d)
        for so in DEST.glob("cell_geom*.so"):
            shutil.copy2(so, d / "pyafv" / so.name)
        _UNPACKED = d
    return _UNPACKED


def import_reference():
    """Import the built reference as module ``pyafv`` (from oracle/_ref) and return it."""
    if not ref_available():
        raise ImportError("oracle/_ref is not built; run `python oracle/build_ref.py` where /root/reference exists")
    p = str(_unpack())
    if p not in sys.path:
        sys.path.insert(0, p)
    import pyafv  # noqa: E402
    import pyafv.backend  # noqa: E402
    if pyafv.backend._BACKEND_NAME != "cython":  # pragma: no cover
        raise ImportError(f"reference built without its Cython backend: {pyafv.backend._IMPORT_ERROR!r}")
    return pyafv


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "NOT built (no reference sources and no previous build)")
    if ok:
        m = import_reference()
        print("reference version", m.__version__, "backend", m.backend._BACKEND_NAME)
    sys.exit(0 if ok else 1)
