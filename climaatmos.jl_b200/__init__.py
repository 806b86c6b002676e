"""B200-native dynamical-core time step for ClimaAtmos (dry / tracer-carrying configs).

Host-side mirror of the reference's hook surface (ClimaODEFunction hooks wired at
src/simulation/integrator.jl:215-225) on top of the C-ABI library ``libb200dycore.so``
(declared in include/b200_dycore.h).  No CPU fallback: every compute entry point raises if the
CUDA library is missing.
"""
__version__ = "0.1.0"
