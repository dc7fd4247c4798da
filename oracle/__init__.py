"""CPU oracle for the transformer-attribution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The shipped path (``transformer_explainability_b200``)
never imports this package and fails loudly when its CUDA library is missing.

Contents
--------
``rules``        closed-form restatement of every LRP rule on the path
                 (``modules/layers_ours.py`` of the reference), any float dtype.
``vit``          ViT forward / attention-gradient / relprop / rollout built from
                 ``rules`` (``baselines/ViT/ViT_LRP.py`` wiring), per-sample reductions.
``bert``         BERT forward / relprop / normalised rollout
                 (``BERT_explainability/modules/BERT/*.py`` wiring).
``ref_harness``  imports the UNMODIFIED reference from ``/root/reference`` with the
                 import shims of SURVEY.md §8c (authoring container only — the
                 reference does not exist on the GPU box).
``make_golden``  script that pins the oracle to the reference and writes the
                 fixtures in ``tests/golden/``.

Parity pinning: the reference holds no runnable golden vectors for this path
(its only known answers need downloaded weights), so the oracle
is pinned against outputs of the reference itself run in the authoring
container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
"""
