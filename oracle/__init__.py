"""TEST INFRASTRUCTURE — CPU checkers for the SWTPG hot path. See oracle/README.md.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
