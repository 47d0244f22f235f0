"""B200-native detection hot path (SigLIP-2 ViT forward + FreqMLP + fusion head + CORAL).

The directory name is not a Python identifier; import it through the `dfd` alias package at the repo root
(`import dfd`, `from dfd import ops, engine, scoring`), which puts this directory on its `__path__`.
All arithmetic runs in `libdfd.so` (csrc/, C ABI in include/dfd.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
