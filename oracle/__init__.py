"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's SFNO forward path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker (or as the timed CPU baseline), never as a
fallback for the CUDA path.

Parity status: the reference ships no tests, golden vectors or fixtures for this path
 -> "parity unpinned" by the reference itself.  The restatement is
instead pinned against (i) the reference's own Python classes imported through
``oracle/ref_shim.py`` in the build container (fixtures + generator committed under
``tests/golden``), and (ii) independent analytical identities (Gram orthonormality of the
Legendre tables, scipy spherical harmonics, Legendre-Gauss round trip).
"""
