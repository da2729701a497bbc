"""dcb200: B200-native (sm_100a) implementation of DynamiCrafter's denoising hot path.

Drop-in mirrors of the reference's classes live under ``dynamicrafter_b200.lvdm`` with the
reference's module paths; all compute goes through ``libdcb200.so`` (hand-written CUDA, C ABI in
``include/dcb200.h``).  There is no CPU / PyTorch fallback.
"""
__version__ = "0.1.0"
