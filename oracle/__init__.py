"""CPU oracle (TEST INFRASTRUCTURE).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product never does."""
