"""CPU oracle for the laplace graph-propagation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker / the CPU
baseline, never as the thing shipped.  The product package
(``laplace_gnn_recommendation_b200``) never imports this module and fails
loudly when its CUDA library is missing.

Parity status
-------------
* Functions whose arithmetic lives in the reference tree itself
  (``bpr_loss``, ``make_predictions_for_user``, ``difference_1d``,
  ``padded_stack``, ``RecallPrecision_ATk``, ``NDCGatK_r``, ``split``,
  ``both_indexes_from_zero``, ``LightGCN.forward``'s cat/stack/mean/split,
  ``GNNEncoder``/``EdgeDecoder`` control flow) are PINNED: the fixtures under
  ``tests/golden/`` were produced by importing the real reference modules from
  ``/root/reference`` (``tests/golden/make_golden.py``) and the oracle is
  checked against them in ``tests/test_oracle_golden.py``.
* Functions whose arithmetic lives in third-party packages that are absent
  from ``/root/reference`` and not installable here (torch_sparse ``SparseTensor``
  / ``matmul``, PyG ``gcn_norm`` / ``structured_negative_sampling`` /
  ``SAGEConv`` / ``to_hetero``; versions unpinned in ``environment.yml:28-30``)
  are restated from their published algorithm:
  **parity unpinned** for those pieces.  They are cross-checked against
  independent implementations (scipy.sparse, dense matmul, python loops) and
  the hand-computed known answers of SURVEY.md Appendix B.
"""
from . import lightgcn_oracle, sampler_oracle, topk_oracle, hetero_oracle  # noqa: F401
