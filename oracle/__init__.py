"""CPU oracle for the kron / operator-tree `mul!` path. TEST INFRASTRUCTURE ONLY: importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never from the
product package."""
