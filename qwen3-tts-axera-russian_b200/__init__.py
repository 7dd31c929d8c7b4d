"""B200-native backend for the Qwen3-TTS 12 Hz codec vocoder hot path.

The directory name carries a hyphen (it mirrors the reference repository's name), so import
it with ``importlib.import_module("qwen3-tts-axera-russian_b200")`` or through the
``voc_b200`` alias module at the repository root.
"""
from .config import VocoderConfig
from .weights import MODEL_SUFFIX, init_weights, load_model, save_model, weight_shapes

__all__ = ["VocoderConfig", "MODEL_SUFFIX", "init_weights", "load_model", "save_model",
           "weight_shapes"]
