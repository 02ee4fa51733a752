"""fbtt_embedding_b200 -- B200-native (sm_100a) TT-EmbeddingBag.

``fbtt_embedding_b200.tt_embeddings``      the 11-op extension interface over libttb.so (C ABI)
``fbtt_embedding_b200.tt_embeddings_ops``  TTEmbeddingBag / TableBatchedTTEmbeddingBag / OptimType / ...
``fbtt_embedding_b200.grouped``            TTEmbeddingBagGroup: differently-shaped tables, one call per phase
``fbtt_embedding_b200.fused``              FusedTTEmbeddingBag: tables sharing q-shapes / ranks, one LAUNCH per phase
``fbtt_embedding_b200.sharded``            table-parallel sharding over the GPUs of one box (one all-to-all)
``fbtt_embedding_b200.replicated``         data-parallel replicas of one table (one all-reduce)

To run code written against the reference unchanged (``import tt_embeddings``,
``from tt_embeddings_ops import TTEmbeddingBag``) put ``fbtt_embedding_b200/dropin`` on
``PYTHONPATH``.
"""
from . import tt_embeddings  # noqa: F401  (fails loudly when libttb.so is missing)
from .tt_embeddings_ops import (  # noqa: F401
    BufferList,
    OptimType,
    TableBatchedTTEmbeddingBag,
    TTEmbeddingBag,
    TTLookupFunction,
    suggested_tt_shapes,
    tt_matrix_to_full,
)

from .grouped import TTEmbeddingBagGroup  # noqa: E402,F401
from .fused import FusedTTEmbeddingBag  # noqa: E402,F401

__all__ = ["tt_embeddings", "OptimType", "TTEmbeddingBag", "TableBatchedTTEmbeddingBag", "TTLookupFunction",
           "BufferList", "suggested_tt_shapes", "tt_matrix_to_full", "TTEmbeddingBagGroup", "FusedTTEmbeddingBag"]
