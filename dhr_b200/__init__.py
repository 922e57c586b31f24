"""dhr_b200: B200-native GIP retrieval hot path of castorini/dhr (retrieval/gip_retrieval.py).

Importing the package never touches the GPU; the CUDA library is loaded on first use and its absence is
a hard error (there is no CPU fallback in the product path).
"""
from .index import GipIndex, topk_merge, merge_keys, pack_keys, unpack_keys  # noqa: F401
from .gip_retrieval import GIP_retrieval, IP_retrieval, shard_bounds, write_trec  # noqa: F401

__all__ = ['GipIndex', 'topk_merge', 'merge_keys', 'pack_keys', 'unpack_keys', 'GIP_retrieval', 'IP_retrieval', 'shard_bounds', 'write_trec']
