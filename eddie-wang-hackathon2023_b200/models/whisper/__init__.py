from .model import CrossAttn_KV, KVLinearBlock, MLP, ResidualAttentionBlock, WhisperDecoder, WhisperEncoder  # noqa: F401
