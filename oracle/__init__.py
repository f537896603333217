"""ORACLE — CPU/fp32 restatement of the reference's hot path. Test infrastructure only: importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg; never from mvd_b200/."""
