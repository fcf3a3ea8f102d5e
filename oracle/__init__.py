"""oracle/ -- CPU restatement of the reference algorithm.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker.  The product (equi_articulated_pose_b200)
never imports it and fails loudly when its CUDA library is missing.

  cops.py        ctypes front-end of oracle_ops.c (ball query, FPS, gather, chamfer)
  so3.py         torch-fp32 restatement of the SO(3) grouping / conv / block math
  ref_harness.py shims to import the reference's own Python in place (this container only)
  build_ref.py   recipe that compiles the reference's own CUDA kernels into oracle/_ref/
"""
