"""Import the REAL reference (castorini/dhr retrieval/gip_retrieval.py).  TEST INFRASTRUCTURE ONLY.

Two places it can come from:
  * /root/reference            -- the build container (golden-vector generation, tests/golden/make_golden.py);
  * oracle/_ref/retrieval/...  -- a verbatim copy made by the committed recipe `make -C oracle ref` (git-ignored, NOT
    gpurun-ignored, so it travels to the GPU box like a built .so).  bench.py times THIS file as the CPU baseline
    (`cpu_baseline.kind = "reference"`); when it is absent the torch-op port in gip_oracle.py is timed (`"port"`).

The two shims replace modules the reference imports at module scope but that are not installed here: ``pickle5``
(stdlib pickle reads/writes protocol 4) and ``faiss`` (only PQ_IP_retrieval / faiss_search touch it).
"""
import importlib.util
import os
import pickle
import sys
import types

REFERENCE_ROOT = '/root/reference'
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(_HERE, '_ref', 'retrieval', 'gip_retrieval.py')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'retrieval', 'gip_retrieval.py'))


def copy_available():
    return os.path.isfile(REF_COPY)


def _shims():
    sys.modules.setdefault('pickle5', pickle)
    sys.modules.setdefault('faiss', types.ModuleType('faiss'))


def load():
    """the reference module from /root/reference (build container only)"""
    if not available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    _shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import retrieval.gip_retrieval as ref
    return ref


def load_copy():
    """the reference module from the oracle/_ref copy (travels to the GPU box)"""
    if not copy_available():
        raise RuntimeError('oracle/_ref is missing: run `make -C oracle ref` in the build container')
    _shims()
    spec = importlib.util.spec_from_file_location('dhr_ref_gip_retrieval', REF_COPY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
