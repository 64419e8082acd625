"""CPU oracle for parllel's batch post-processing path.

TEST INFRASTRUCTURE ONLY: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package.  The product package ``parllel_b200`` never does.
"""
