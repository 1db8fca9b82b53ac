"""TEST INFRASTRUCTURE -- CPU restatement (NumPy + a little C) of the reference's
algorithm for the hot path.  Never imported by cupy_b200/; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it."""
