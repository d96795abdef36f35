"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU / plain-PyTorch restatements of the reference's algorithms for the hot path, used only
as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs. Nothing under iclr2025_3d-mom_b200/ imports this package.
"""
