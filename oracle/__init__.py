"""TEST INFRASTRUCTURE ONLY: CPU oracle of the two-view hot path (see two_view.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this package.  The product (openwork_b200/) never does.
"""
