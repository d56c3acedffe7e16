"""CPU oracle for the FCOS hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or as the timed CPU baseline.
``pytorch_object_detection_b200`` never imports this package.

Parity status: PINNED.  ``tests/golden/*.npz`` hold outputs of the unmodified reference
(``/root/reference/model/modules/head.py``, ``model/loss.py``) and of the installed
torchvision 0.26.0 CPU ``batched_nms`` on seeded inputs, produced by
``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks this oracle
against them, plus the reference's one known-answer value (``model/loss.py:219-221``).
"""
