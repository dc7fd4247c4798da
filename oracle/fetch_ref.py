"""Mirror the reference files of the hot path into ``oracle/_ref/`` (TEST INFRASTRUCTURE; git-ignored, travels to the GPU box).

``/root/reference`` exists only in the authoring container.  ``bench.py``'s ``cpu_baseline`` leg and ``--impl
reference`` arm must time the reference's OWN code on the GPU box's host cores (``cpu_baseline.kind == "reference"``), so
``__graft_entry__.build()`` runs this recipe whenever the reference is present: it copies — byte for byte, never into
the tracked tree — the files SURVEY.md §8(a) cites plus the modules they import:

    modules/{__init__,layers_ours,layers_lrp}.py
    baselines/ViT/{ViT_LRP,ViT_explanation_generator,ViT_new,ViT_orig_LRP,helpers,weight_init,layer_helpers}.py
    BERT_explainability/modules/{__init__,layers_ours,layers_lrp}.py
    BERT_explainability/modules/BERT/{BERT,BertForSequenceClassification,ExplanationGenerator}.py
    BERT_rationale_benchmark/{__init__.py, models/model_utils.py}

``oracle/ref_harness.py`` imports the reference from ``/root/reference`` when it exists and from ``oracle/_ref``
otherwise, with the same import shims; nothing in the product package reads either.

    python -m oracle.fetch_ref            # -> oracle/_ref/, prints the file list and a sha256 manifest
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("TE_REFERENCE_SRC", "/root/reference")

FILES = [
    "modules/__init__.py", "modules/layers_ours.py", "modules/layers_lrp.py",
    "baselines/ViT/ViT_LRP.py", "baselines/ViT/ViT_explanation_generator.py", "baselines/ViT/ViT_new.py",
    "baselines/ViT/ViT_orig_LRP.py", "baselines/ViT/helpers.py", "baselines/ViT/weight_init.py",
    "baselines/ViT/layer_helpers.py",
    "BERT_explainability/modules/__init__.py", "BERT_explainability/modules/layers_ours.py",
    "BERT_explainability/modules/layers_lrp.py", "BERT_explainability/modules/BERT/BERT.py",
    "BERT_explainability/modules/BERT/BertForSequenceClassification.py",
    "BERT_explainability/modules/BERT/ExplanationGenerator.py",
    "BERT_rationale_benchmark/__init__.py", "BERT_rationale_benchmark/models/model_utils.py",
]


def fetch(verbose=False):
    """Copy the files; returns the number copied (0 when the reference is absent — the existing mirror is kept)."""
    if not os.path.isdir(os.path.join(SRC, "baselines", "ViT")):
        return 0
    manifest = []
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), rel))
        if verbose:
            print(rel)
    with open(os.path.join(DEST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    return len(manifest)


if __name__ == "__main__":
    n = fetch(verbose=True)
    print("%d files -> %s" % (n, DEST) if n else "reference not present at %s; nothing copied" % SRC)
    sys.exit(0)
