"""The error contract of the engine, in one place (DESIGN.md §1 states the same numbers; every GPU parity test imports
them from here).

  C128_BOUND      complex128 / float64: max-abs error relative to the largest |reference| entry (north star: ~1e-12)
  C64_STEP_BOUND  complex64, ONE pairwise step on the default 3xTF32 path (error-free hi/lo split, TMEM chains cut every
                  128 k and summed round-to-nearest), K <= 8192, O(1) Gaussian data: relative to the largest |C| entry
  C64_PATH_BOUND  complex64, a whole contraction path / slice sum / amplitude (hundreds of steps, FP32 rounding on every
                  intermediate): relative to the largest |reference| value of the comparison
The FP32-SIMT mode (TNB_C64_SIMT) and the oracle run in complex64 obey the same two numbers: the bound is that of FP32
arithmetic, the tensor-core path must not be worse than it by more than the factor checked in test_gpu_tc.py.
"""
C128_BOUND = 1e-12
C64_STEP_BOUND = 2e-5
C64_PATH_BOUND = 5e-5
