"""kotoba_whisper_b200 — B200-native (sm_100a) implementation of kotoba-whisper's batched transcription hot path:
log-mel -> Whisper encoder -> greedy decode, behind the reference's own call surface.

    from kotoba_whisper_b200 import WhisperFeatureExtractorB200, WhisperB200ForConditionalGeneration

The arithmetic lives in hand-written CUDA kernels (csrc/) reached through the C ABI in include/kwb200.h.  There is no
CPU or PyTorch fallback: a missing libkwb200.so or CUDA device raises.
"""
from ._lib import KwError, LIB_PATH  # noqa: F401
from .feature_extraction import BatchFeature, LogMelProducer, WhisperFeatureExtractorB200  # noqa: F401
from .modeling import (GenerateStream, WhisperB200Config, WhisperB200ForConditionalGeneration,  # noqa: F401
                       WhisperB200GenerationConfig)
from .pipeline import AsrPipelineB200, chunk_iter, merge_chunk_tokens, transcribe_longform  # noqa: F401
# the `pipeline(...)` factory lives in kotoba_whisper_b200.pipeline (not re-exported: it would shadow the submodule)

__all__ = ["WhisperFeatureExtractorB200", "WhisperB200ForConditionalGeneration", "WhisperB200Config",
           "WhisperB200GenerationConfig", "BatchFeature", "KwError", "chunk_iter", "merge_chunk_tokens",
           "transcribe_longform", "AsrPipelineB200", "LogMelProducer"]
