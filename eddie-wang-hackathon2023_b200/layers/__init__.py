"""Layer classes under the reference's names (T/tensorrt_llm/layers/): thin eager modules over b200_whisper.functional."""
from .attention import Attention, AttentionMaskType, PositionEmbeddingType, RaggedTensor  # noqa: F401
from .conv import Conv1d  # noqa: F401
from .normalization import LayerNorm  # noqa: F401
