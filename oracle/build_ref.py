#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: the UNMODIFIED reference (dgsmith7/nerf-mlp), compiled for the CPU arm.

TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package (nerf_mlp_b200).

The reference's hot path is pure Python over PyTorch (nerfmlp/model.py, nerfmlp/renderer.py; it ships no
setup.py / pyproject, so there is nothing to ``pip install``).  "Building" it therefore means byte-compiling the
package's modules from the sources WHERE THEY LIE under /root/reference into sourceless ``.pyc`` files under
``oracle/_ref/nerfmlp/`` -- the Python analogue of compiling a C reference into ``oracle/_ref/*.so``:

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
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("NERF_REFERENCE_SRC", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("__init__", "model", "renderer", "data")          # /root/reference/nerfmlp/*.py


def build(verbose=True):
    """Compile /root/reference/nerfmlp/*.py -> oracle/_ref/nerfmlp/*.pyc.  Returns True if built."""
    pkg_src = os.path.join(REF_SRC, "nerfmlp")
    if not os.path.isdir(pkg_src):
        if verbose:
            print(f"oracle/build_ref.py: {pkg_src} not present (GPU box?) -- keeping the prebuilt oracle/_ref, if any")
        return False
    pkg_out = os.path.join(OUT, "nerfmlp")
    os.makedirs(pkg_out, exist_ok=True)
    for name in MODULES:
        src = os.path.join(pkg_src, name + ".py")
        py_compile.compile(src, cfile=os.path.join(pkg_out, name + ".pyc"), dfile=f"<reference>/nerfmlp/{name}.py",
                           doraise=True, optimize=0)
    with open(os.path.join(OUT, "BUILD_INFO.txt"), "w") as f:
        f.write(f"compiled from {pkg_src} by oracle/build_ref.py with CPython {sys.version.split()[0]} "
                f"(magic {importlib.util.MAGIC_NUMBER.hex()}); modules: {', '.join(MODULES)}\n")
    if verbose:
        print(f"oracle/build_ref.py: compiled {len(MODULES)} modules of the reference into {pkg_out}")
    return True


def available():
    """(ok, why): can the compiled reference be imported by THIS interpreter?"""
    pkg_out = os.path.join(OUT, "nerfmlp")
    for name in MODULES:
        f = os.path.join(pkg_out, name + ".pyc")
        if not os.path.exists(f):
            return False, f"{os.path.relpath(f, os.path.dirname(HERE))} missing (run oracle/build_ref.py where /root/reference exists)"
        with open(f, "rb") as fh:
            if fh.read(4) != importlib.util.MAGIC_NUMBER:
                return False, "oracle/_ref was compiled by a different CPython (bytecode magic mismatch)"
    return True, ""


def load():
    """Import the compiled reference package; returns the module ``nerfmlp`` (NeRFMLP, NeRFRenderer, ...)."""
    ok, why = available()
    if not ok:
        raise ImportError(why)
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    mod = importlib.import_module("nerfmlp")
    origin = os.path.dirname(os.path.abspath(mod.__file__))
    if origin != os.path.join(OUT, "nerfmlp"):
        raise ImportError(f"`nerfmlp` resolved to {origin}, not to oracle/_ref")
    return mod


if __name__ == "__main__":
    build()
