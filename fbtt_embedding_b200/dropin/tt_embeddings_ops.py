"""`from tt_embeddings_ops import TTEmbeddingBag, ...` shim (see tt_embeddings.py beside it)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from fbtt_embedding_b200.tt_embeddings_ops import (  # noqa: F401,E402
    BufferList, OptimType, TableBatchedTTEmbeddingBag, TTEmbeddingBag, TTLookupFunction,
    suggested_tt_shapes, tt_matrix_to_full)
