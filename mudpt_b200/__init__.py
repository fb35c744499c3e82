"""mudpt_b200 -- B200-native (sm_100a) implementation of the MuDPT hot path.

Public surface mirrors the reference: `mudpt_b200.clip` (CLIP containers, tokenize) and
`mudpt_b200.trainers.mudpt` (MuDPTPromptLearner, TextEncoder, CustomCLIP, MuDPT).  The compute
lives in `mudpt_b200/lib/libmudpt_b200.so` (C ABI: include/mudpt_b200.h).
"""
__version__ = "0.1.0"
