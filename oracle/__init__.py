"""CPU oracle for the two-view augmentation + contrastive-loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``medical_image_segmentation_b200/``
imports this package; it is used by ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker
and as the timed CPU arm -- never as the product path.

Pinning status
--------------
* augmentation  : PINNED.  ``oracle/make_golden.py`` runs the reference's own
  ``BYOLRGBDataTransforms`` (imported from /root/reference with stub modules,
  see ``oracle/ref_import.py``) and commits its outputs under ``tests/golden/``;
  ``tests/test_oracle_aug.py`` checks both restatements against those vectors.
* BYOL loss     : PINNED the same way (``BYOL.cosine_similarity_loss``).
* NT-Xent       : PARITY UNPINNED.  The reference contains no NT-Xent / InfoNCE
  (its SSL loss is BYOL, train/model/byol_pytorch.py:181-198); the oracle is the
  canonical SimCLR formulation checked against torch autograd in fp64 only.
"""
