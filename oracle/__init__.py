"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's SAUNet forward/backward + DualLoss hot path
(sunjesse/shape-attentive-unet).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
anything from this package, and only as the checker -- never as the thing that
is measured as the product or shipped.  The product path (the
``shape-attentive-unet_b200`` package) never imports it and fails loudly when
its CUDA library is missing.

Parity status: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF, generated in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference read-only) and
committed under ``tests/golden/``.
"""
