"""codae -- B200-native drop-in for the hot path of victordeleau/MUI-DeepAutoEncoder (CODAE).

Same import paths and class signatures as the reference (`codae.model`, `codae.tool`, `codae.dataset`);
the device work goes through libcodae_b200.so (hand-written sm_100a CUDA behind a C ABI, include/codae_b200.h).
"""
__version__ = "0.1.0"
