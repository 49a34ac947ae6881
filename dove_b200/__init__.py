"""dove_b200 — B200-native (sm_100a) implementation of DOVE's one-step video super-resolution hot path.

Host side: Python mirror of the diffusers `CogVideoXPipeline` surface that
/root/reference/inference_script.py:394-503 (`process_video`) uses; device side: hand-written CUDA kernels
behind the C ABI in include/dove_b200.h (libdove_b200.so, built by `python -m dove_b200.build`).
"""
__version__ = "0.1.0"
