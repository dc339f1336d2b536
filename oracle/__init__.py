"""CPU oracle for the ADS-B decode hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(air_rs_b200) must never import it.
"""
