"""TEST INFRASTRUCTURE ONLY — the correctness oracle for the vtb200 hot path.

Nothing under oracle/ is on the product path: only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import it (as the checker or the timed CPU baseline, never as the
thing shipped).  See oracle/restate.py for the restatement and oracle/ref_loader.py for how the real
reference is imported (in the build container only) to pin it.
"""
