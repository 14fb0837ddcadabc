"""oracle -- TEST INFRASTRUCTURE ONLY (CPU restatement of dspfun's DCT hot path).

Nothing under dspfun_b200/ may import this package.  Allowed users: tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, and there only as the checker / CPU baseline.
"""
