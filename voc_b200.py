"""Importable alias for the hyphenated package directory ``qwen3-tts-axera-russian_b200``."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
sys.modules[__name__] = _pkg
