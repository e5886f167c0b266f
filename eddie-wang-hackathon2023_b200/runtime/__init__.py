from .whisper_decoding import WhisperDecoding  # noqa: F401
from .whisper_encoder import WhisperEncoder  # noqa: F401
from .sharding import gather_token_ids, shard_bounds  # noqa: F401
