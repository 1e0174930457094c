"""ORACLE — CPU restatements of the reference's concept-database path. Test infrastructure only:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from
semanticlens_b200/."""
