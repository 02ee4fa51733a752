"""`import tt_embeddings` shim: put this directory on PYTHONPATH to run reference-era code
(tt_embeddings_ops.py:14) on the B200 library without touching it."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from fbtt_embedding_b200.tt_embeddings import *  # noqa: F401,F403,E402
from fbtt_embedding_b200.tt_embeddings import (  # noqa: F401,E402
    cache_backward_dense, cache_backward_rowwise_adagrad_approx, cache_backward_sgd, cache_forward,
    cache_populate, preprocess_indices_sync, tt_adagrad_backward, tt_dense_backward, tt_forward,
    tt_sgd_backward, update_cache_state)
