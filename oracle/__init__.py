"""CPU oracle for the MeshAnything-350M hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package.  The product (`meshanything_b200`, `MeshAnything`) never does.
"""
