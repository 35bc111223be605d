"""B200-native (sm_100a) encoder + serialized-CTC hot path for Multi-talker-ASR-with-LLMs.

Host side mirrors the reference's Python surface (WavLMModel / Separator / CTC / HybridLoss and the greedy-CTC
helpers); every arithmetic op below it is a hand-written CUDA kernel reached through the C ABI in
``include/mtasr.h`` (``libmtasr.so``).  There is no CPU or library fallback: importing the kernels without the
built library, or calling them without a CUDA device, raises.
"""
__version__ = "0.1.0"
