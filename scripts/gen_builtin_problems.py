"""Regenerate probnumdiffeq.jl_b200/csrc/kernels/pn_builtin_problems.cuh from problems.py (SymPy trace -> CUDA C++)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import tracer
path = os.path.join(ROOT, "probnumdiffeq.jl_b200", "csrc", "kernels", "pn_builtin_problems.cuh")
open(path, "w").write(tracer.builtin_header())
print("wrote", path)
