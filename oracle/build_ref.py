#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: the UNMODIFIED reference (dgsmith7/nerf-mlp), compiled for the CPU arm.

TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package (nerf_mlp_b200).

The reference's hot path is pure Python over PyTorch (nerfmlp/model.py, nerfmlp/renderer.py; it ships no
setup.py / pyproject, so there is nothing to ``pip install``).  "Building" it therefore means byte-compiling the
package's modules from the sources WHERE THEY LIE under /root/reference into sourceless ``.pyc`` files, packed into
``oracle/_ref/nerfmlp_ref.zip`` (one binary artefact; loose ``*.pyc`` files are filtered out of the snapshot that goes
to the GPU box) -- the Python analogue of compiling a C reference into ``oracle/_ref/*.so``:

    python oracle/build_ref.py            # needs /root/reference (the build container); a no-op message elsewhere

No reference SOURCE is copied into the repository: ``oracle/_ref/`` is git-ignored (it is NOT gpurun-ignored, so the
compiled files travel to the GPU box exactly like the built ``libnerf_b200.so``), and the ``.pyc`` files are produced
by CPython's own compiler from the reference files in place.  The GPU box runs the same image (same interpreter), so
the bytecode loads there; ``load()`` checks the magic number and reports a clean "unavailable" otherwise.

Users: ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg: ``kind: "reference"`` when ``oracle/_ref``
loads, else the bit-identical torch port ``oracle/nerf_oracle_torch.py``, ``kind: "port"``) and
``tests/test_oracle_golden.py::test_ref_build_matches_golden`` (the compiled reference reproduces the committed
golden vectors bit for bit).
"""
import importlib
import importlib.util
import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("NERF_REFERENCE_SRC", "/root/reference")
OUT = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT, "nerfmlp_ref.zip")
MODULES = ("__init__", "model", "renderer", "data")          # /root/reference/nerfmlp/*.py


def build(verbose=True):
    """Compile /root/reference/nerfmlp/*.py -> oracle/_ref/nerfmlp_ref.zip (nerfmlp/*.pyc inside).  Returns True if built."""
    pkg_src = os.path.join(REF_SRC, "nerfmlp")
    if not os.path.isdir(pkg_src):
        if verbose:
            print(f"oracle/build_ref.py: {pkg_src} not present (GPU box?) -- keeping the prebuilt oracle/_ref, if any")
        return False
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="nerfref_")
    try:
        with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_STORED) as z:
            for name in MODULES:
                cfile = os.path.join(tmp, name + ".pyc")
                py_compile.compile(os.path.join(pkg_src, name + ".py"), cfile=cfile, dfile=f"<reference>/nerfmlp/{name}.py",
                                   doraise=True, optimize=0)
                z.write(cfile, f"nerfmlp/{name}.pyc")
            z.writestr("BUILD_INFO.txt", f"compiled from {pkg_src} by oracle/build_ref.py with CPython {sys.version.split()[0]} "
                                         f"(magic {importlib.util.MAGIC_NUMBER.hex()}); modules: {', '.join(MODULES)}\n")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    legacy = os.path.join(OUT, "nerfmlp")                      # loose .pyc files of an earlier layout
    if os.path.isdir(legacy):
        shutil.rmtree(legacy, ignore_errors=True)
    if verbose:
        print(f"oracle/build_ref.py: compiled {len(MODULES)} modules of the reference into {ARCHIVE}")
    return True


def available():
    """(ok, why): can the compiled reference be imported by THIS interpreter?"""
    if not os.path.exists(ARCHIVE):
        return False, "oracle/_ref/nerfmlp_ref.zip missing (run oracle/build_ref.py where /root/reference exists)"
    try:
        with zipfile.ZipFile(ARCHIVE) as z:
            names = set(z.namelist())
            for name in MODULES:
                member = f"nerfmlp/{name}.pyc"
                if member not in names:
                    return False, f"{member} missing from oracle/_ref/nerfmlp_ref.zip"
                if z.read(member)[:4] != importlib.util.MAGIC_NUMBER:
                    return False, "oracle/_ref was compiled by a different CPython (bytecode magic mismatch)"
    except zipfile.BadZipFile:
        return False, "oracle/_ref/nerfmlp_ref.zip is not a zip archive"
    return True, ""


def load():
    """Import the compiled reference package (zipimport of sourceless bytecode); returns the module ``nerfmlp``."""
    ok, why = available()
    if not ok:
        raise ImportError(why)
    if ARCHIVE not in sys.path:
        sys.path.insert(0, ARCHIVE)
    mod = importlib.import_module("nerfmlp")
    if not os.path.abspath(getattr(mod, "__file__", "")).startswith(os.path.abspath(ARCHIVE)):
        raise ImportError(f"`nerfmlp` resolved to {getattr(mod, '__file__', None)}, not to oracle/_ref")
    return mod


if __name__ == "__main__":
    build()
