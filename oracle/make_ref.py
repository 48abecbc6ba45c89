#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE: make the UNMODIFIED reference importable beside the repo.

    python oracle/make_ref.py            # in the build container, where /root/reference exists

copies /root/reference/src/skan (pure Python, ~2 900 lines) to oracle/_ref/skan and writes the `_version.py` that
setuptools-scm would generate (skan/__init__.py:2 imports it).  oracle/_ref/ is git-ignored (no reference source
enters the history) but not gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.
`import_reference()` then imports that copy with oracle/refshim/ (stand-ins for skimage's map_array and toolz, which
this image lacks) in front of it.  Used by bench.py (`--impl reference`, `cpu_baseline`, `parity`) and tests/ only.
"""
import importlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/src/skan'
DST = os.path.join(HERE, '_ref')
SHIM = os.path.join(HERE, 'refshim')


def make(force=False):
    """Copy the reference package when the reference tree is present; True when oracle/_ref/skan exists afterwards."""
    pkg = os.path.join(DST, 'skan')
    if os.path.isdir(REF_SRC) and (force or not os.path.isdir(pkg)):
        if os.path.isdir(pkg):
            shutil.rmtree(pkg)
        os.makedirs(DST, exist_ok=True)
        shutil.copytree(REF_SRC, pkg, ignore=shutil.ignore_patterns('__pycache__', 'test', '*.pyc'))
        with open(os.path.join(pkg, '_version.py'), 'w') as f:
            f.write("__version__ = '0.0.0+reference'\n")
    return os.path.isdir(pkg)


def available():
    return os.path.isdir(os.path.join(DST, 'skan'))


def import_reference():
    """The reference's own `skan.csr` module (its code, unmodified; stand-ins only for the absent third-party imports)."""
    if not available():
        raise ImportError('oracle/_ref/skan is missing: run `python oracle/make_ref.py` where /root/reference exists')
    try:
        import skimage.util._map_array  # noqa: F401  (the real one, if the image ever has it)
    except Exception:
        if SHIM not in sys.path:
            sys.path.insert(0, SHIM)
        for m in [k for k in sys.modules if k == 'skimage' or k.startswith('skimage.')]:
            del sys.modules[m]
    try:
        import toolz  # noqa: F401
    except Exception:
        if SHIM not in sys.path:
            sys.path.insert(0, SHIM)
    if DST not in sys.path:
        sys.path.insert(0, DST)
    mod = sys.modules.get('skan')
    if mod is not None and not os.path.abspath(getattr(mod, '__file__', '') or '').startswith(DST):
        for m in [k for k in sys.modules if k == 'skan' or k.startswith('skan.')]:
            del sys.modules[m]
    return importlib.import_module('skan.csr')


def map_array_seconds(reset=False):
    """Wall time spent in the map_array stand-in since the last reset (0 when the real scikit-image is in use)."""
    m = sys.modules.get('skimage.util._map_array')
    sec = getattr(m, 'SECONDS', None)
    if sec is None:
        return 0.0
    v = sec[0]
    if reset:
        sec[0] = 0.0
    return v


if __name__ == '__main__':
    ok = make(force='--force' in sys.argv)
    print('oracle/_ref/skan', 'ready' if ok else 'NOT available (no /root/reference here)')
