"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the Synergy-CLIP tri-modal contrastive tail
(reference ``model.py:52-58`` and ``model.py:247-272``).  Nothing under
``synergy_clip_b200/`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and there only as the checker or as the timed CPU
baseline, never as the product path.

Parity pinning: the reference ships no tests or golden vectors for this path
("parity unpinned" by the reference's own tests).  The
oracle is therefore pinned against outputs of the reference's *unmodified*
``model.clip_loss`` run in this container (``oracle/ref_import.py`` +
``tests/golden/make_golden.py``); the resulting vectors are committed under
``tests/golden/`` and ``tests/test_oracle.py`` re-checks the oracle against
them on every run.
"""
