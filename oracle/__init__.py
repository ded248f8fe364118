"""CPU oracle for the CRISPResso alignment + quantification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``crispresso_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs do, and there only as the checker / the CPU arm.
"""
