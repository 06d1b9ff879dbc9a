"""Measurement and test support that is NOT part of the product package (gprmax_b200/):

  synthetic.py      solver-ready grids of the benchmark workloads built without the reference (closed-form tables for a
                    homogeneous box + default PML + one dipole), pinned bit-exactly against the reference's own build
  refmodel.py       the same models built by the UNMODIFIED reference front end (baseline/_ref) when it is present
  sharded_bench.py  bench.py --gpus N: the sharded weak-scaling run and its checks
"""
