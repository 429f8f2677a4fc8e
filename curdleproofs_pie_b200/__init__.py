"""B200-native BLS12-381 G1 hot path of curdleproofs (batched MSM, vector scalar-mul, folds,
(de)compression, Fr vectors) behind the C ABI in include/cpg.h.

    runtime.py   ctypes binding of lib/libcpg.so (nvcc, sm_100a); no CPU fallback
    csrc/        CUDA kernels + the C ABI
The reference-facing ``py_arkworks_bls12381`` surface lives in ``dropin/``.
"""
__all__ = ["runtime"]
