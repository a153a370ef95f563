"""Importable name of the ``neural-speech-decoding_b200/`` source tree.

The product directory carries the repository's hyphenated name, which is not a legal Python
identifier; this stub only extends ``__path__`` to it, so every submodule
(``neural_speech_decoding_b200.lstm_eeg_model`` ...) is loaded from
``neural-speech-decoding_b200/``.
"""
from pathlib import Path as _Path

__path__.append(str(_Path(__file__).resolve().parent.parent / "neural-speech-decoding_b200"))

from .version import __version__  # noqa: E402,F401
