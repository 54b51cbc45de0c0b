"""wav2sleep_b200: B200-native (sm_100a) implementation of the wav2sleep model forward hot path.

Public surface mirrors ``wav2sleep`` (reference ``src/wav2sleep/__init__.py:3-19``) for the part that is in scope:
the model classes (``wav2sleep_b200.model``), ``load_model`` and ``predict`` (``wav2sleep_b200.api``),
``predict_on_folder`` for parquet nights (``wav2sleep_b200.folder``).
Importing the package does not load the CUDA library; the first forward does, and raises if it is missing.
"""
from .model import (  # noqa: F401
    COLS_TO_SAMPLES_PER_EPOCH,
    MultiModalAttentionEmbedder,
    SequenceCNN,
    SignalEncoder,
    SignalEncoders,
    Wav2Sleep,
    build_default,
)

__all__ = [
    "COLS_TO_SAMPLES_PER_EPOCH", "MultiModalAttentionEmbedder", "SequenceCNN", "SignalEncoder", "SignalEncoders",
    "Wav2Sleep", "build_default",
]
