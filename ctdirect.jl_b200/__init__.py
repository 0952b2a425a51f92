"""B200-native NLP-evaluation engine for CTDirect.jl's discretised optimal control problem.

The product is libctdirect_b200.so (csrc/, C ABI in include/ctdirect_b200.h); this package is the
host-side mirror of the reference's interface for that path (docp.py) plus the kernel generator (gen/)
and the in-tree build (build.py)."""
from .docp import (CONS, GRAD, HESS, JAC, OBJ, SCHEMES, B200NLPModel, Collocation, DirectShooting, DOCP, DOCP_constraints,
                   DOCP_initial_guess, DOCP_objective, Meta, direct_transcription, discretize, nlp_model)
from ._lib import CtdError, LIB_PATH, list_models

__all__ = [
    "CONS", "JAC", "HESS", "GRAD", "OBJ", "SCHEMES", "B200NLPModel", "Collocation", "DirectShooting", "DOCP", "Meta",
    "DOCP_constraints", "DOCP_initial_guess", "DOCP_objective", "direct_transcription", "discretize",
    "nlp_model", "CtdError", "LIB_PATH", "list_models",
]
