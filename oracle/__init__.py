"""CPU oracle for the RNA-MPNN structure-to-sequence forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU arm.  The product path (``rna-mpnn_b200/``) never imports it and
fails loudly when the CUDA library is missing.
"""
