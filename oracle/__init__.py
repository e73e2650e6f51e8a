"""ORACLE — CPU restatement of the reference's hot-path algorithms.  TEST INFRASTRUCTURE ONLY.

Nothing under vub_image_denoising_b200/ imports this package.  Allowed importers: tests/,
__graft_entry__.smoke(), bench.py (cpu_baseline leg and --impl reference).
See DESIGN.md §"Oracle" for what pins each part.
"""
