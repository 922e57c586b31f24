"""Import the REAL reference (castorini/dhr) from /root/reference.  TEST INFRASTRUCTURE ONLY.

Only usable in the build container (the GPU box has no /root/reference).  The two
shims replace modules the reference imports at module scope but that are not
installed here: ``pickle5`` (stdlib pickle reads/writes protocol 4) and ``faiss``
(only PQ_IP_retrieval / faiss_search touch it; both are out of scope).
"""
import os
import pickle
import sys
import types

REFERENCE_ROOT = '/root/reference'


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'retrieval', 'gip_retrieval.py'))


def load():
    if not available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    sys.modules.setdefault('pickle5', pickle)
    sys.modules.setdefault('faiss', types.ModuleType('faiss'))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import retrieval.gip_retrieval as ref
    return ref
