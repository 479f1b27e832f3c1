"""probnumdiffeq.jl_b200 -- B200-native ODE-filter hot path of ProbNumDiffEq.jl.

Host-side mirror of the reference's `solve(prob, EK0()/EK1())` / `EnsembleProblem` surface
(`api.py`), the SymPy tracer that turns a user RHS into CUDA C++ for NVRTC (`tracer.py`), and the
ctypes binding of the C-ABI library `libpnode_cuda.so` (`lib.py`, sources in `csrc/`).
"""
from . import tracer, problems  # noqa: F401
