from .whisper_decoding import WhisperDecoding  # noqa: F401
from .whisper_encoder import WhisperEncoder  # noqa: F401
from .sharding import gather_token_ids, shard_bounds  # noqa: F401
from .checkpoint import ModelDimensions, load_checkpoint, read_kv_scales, save_checkpoint, write_kv_scales  # noqa: F401
from .pipeline import WhisperPipeline  # noqa: F401
from .plugin_step import PluginDecoderStep  # noqa: F401
