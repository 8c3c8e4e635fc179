"""distil-whisper-b200: B200-native (sm_100a) compute path for the Distil-Whisper KD training step.

Drop-in for the Hugging Face Whisper modules that huggingface/distil-whisper's training/run_distillation.py drives
(ref:training/run_distillation.py:1465-1495).  All arithmetic runs in libdwb.so (include/dwb.h); there is no CPU or
eager fallback -- importing the compute modules without the built library raises.
"""
__version__ = "0.1.0"
