"""Model classes under the reference's names (T/tensorrt_llm/models/)."""
from .whisper import CrossAttn_KV, KVLinearBlock, ResidualAttentionBlock, WhisperDecoder, WhisperEncoder  # noqa: F401
