"""TEST INFRASTRUCTURE — the CPU oracle for the LBAudioDetective fingerprint path. Not part of the product.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this package.
"""
