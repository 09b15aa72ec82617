"""CPU oracle for the BTF Gibbs sweep.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the shipped engine.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it,
and only as the checker / the timed CPU baseline -- never as a fallback for the
CUDA path (``functionalmf_b200`` raises if its CUDA library is missing).
"""
