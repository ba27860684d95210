"""Parity tests: oracle vs the reference fixtures (CPU), CUDA path vs the oracle through the C ABI (GPU)."""
