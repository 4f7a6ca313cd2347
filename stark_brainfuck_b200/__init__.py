"""stark_brainfuck_b200 -- sm_100a polynomial / FRI / Merkle engine behind the call surface of
aszepieniec/stark-brainfuck's hot path (ntt.ntt/intt, Polynomial.scale/evaluate_domain,
Fri.Domain.*, Fri.commit/query/prove, Merkle).  See DESIGN.md and INTEGRATION.md.

The directory is named with an underscore so that it is importable; it is the package the
task statement calls `stark-brainfuck_b200/`.
"""
from ._lib import B2SError, LeafTemplates, load  # noqa: F401
from .engine import Engine, default_engine, set_default_engine  # noqa: F401

__all__ = ["B2SError", "LeafTemplates", "load", "Engine", "default_engine", "set_default_engine"]
