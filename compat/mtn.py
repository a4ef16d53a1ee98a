"""Import shim: ``import mtn`` / ``from mtn import *`` (reference train.py:18, generate.py)
resolves to the B200 implementation, and whole-module pickles written by the reference's
``torch.save(model)`` name their classes ``mtn.<Class>`` -- which this module provides.
Put this directory first on PYTHONPATH; see INTEGRATION.md."""
from mtn_b200.mtn import *            # noqa: F401,F403
from mtn_b200.mtn import (EncoderDecoder, Generator, Encoder, LayerNorm, SublayerConnection,  # noqa: F401
                          Decoder, DecoderLayer, MultiHeadedAttention, PositionwiseFeedForward,
                          Embeddings, PositionalEncoding, VideoEncoder, make_model, clones, attention)
