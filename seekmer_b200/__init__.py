"""seekmer_b200 — B200-native implementation of Seekmer's bulk-infer hot path.

Drop-in surface (mirrors of the reference modules): `seekmer_b200.common`,
`seekmer_b200.mapper`, `seekmer_b200.infer`, `seekmer_b200.impute`, CLI
`python -m seekmer_b200 infer|impute ...`.
The compute runs in `libseekmer_b200.so` (hand-written sm_100a CUDA behind the C ABI of
`include/seekmer_b200.h`); importing the package never needs a GPU, using it does.
"""
__version__ = '0.1.0'
