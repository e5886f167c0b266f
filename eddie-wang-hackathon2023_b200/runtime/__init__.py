from .whisper_decoding import WhisperDecoding  # noqa: F401
