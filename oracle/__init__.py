"""ORACLE -- test infrastructure only.

A CPU/PyTorch restatement of the reference's algorithm for the hot path:
marigold_dc.py's guided DDIM loop and the diffusers==0.31.0 modules it calls.  PARITY UNPINNED: the
reference cannot be imported here (diffusers is absent) and ships no tests or golden vectors; the
restatement is pinned only by known-answer tests of the formulas, parameter counts and key names.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything from this package.  The product (depth_completion_b200/) never does.
"""
