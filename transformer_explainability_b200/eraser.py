"""Batched driver for the ERASER rationale files of the reference's BERT pipeline
(``BERT_rationale_benchmark/models/pipeline/bert_pipeline.py:456-582``).

The reference explains one document at a time (``test_batch_size = 1``), clamps the scores at zero, takes the top-k
entries for k = 5, 10, ..., 80 and appends one JSON line per document to ``identifier_results_{k}.json``.  Here the
explanations of a whole batch come from one engine call (any generator of ``ExplanationGenerator.Generator``) and
the top-k selection runs on the GPU; the files have the reference's format, including its quirk that
``hard_rationales`` is never reset between the k's of a document (the k = 10 line holds the top-5 entries followed by
the top-10 entries, ``:563-575``).  Tokenisation, word-piece -> word pooling and the dataset readers of the pipeline
are outside this package: pass ``word_scores`` to pool token scores to words the way
``scores_per_word_from_scores_per_token`` does upstream."""
import json
import os

import torch

KS = tuple(range(5, 85, 5))


def topk_rationales(scores, ks=KS):
    """scores [B,T] (GPU or CPU) -> list (per document) of list (per k) of index lists, ``cam.topk(k)`` order."""
    scores = scores.clamp(min=0)
    kmax = min(max(ks), scores.shape[1])
    idx = scores.topk(kmax, dim=1).indices.cpu()
    return [[idx[b, :min(k, kmax)].tolist() for k in ks] for b in range(scores.shape[0])]


def rationale_lines(doc_ids, scores, ks=KS):
    """One JSON string per (k, document) in the reference's accumulation order: returns {k: [line, ...]}."""
    out = {k: [] for k in ks}
    for doc, per_k in zip(doc_ids, topk_rationales(scores, ks)):
        hard = []
        for k, indices in zip(ks, per_k):
            hard.extend({"start_token": i, "end_token": i + 1} for i in indices)
            out[k].append(json.dumps({"annotation_id": doc,
                                      "rationales": [{"docid": doc, "hard_rationale_predictions": list(hard)}]}))
    return out


def write_identifier_results(generator, input_ids, attention_mask, targets, doc_ids, output_dir, method="generate_LRP",
                             folder="ours", word_scores=None, ks=KS, **kw):
    """Explain a batch for its target classes with ``getattr(generator, method)`` and append to
    ``output_dir/folder/identifier_results_{k}.json``.  Returns the [B,T] scores."""
    scores = getattr(generator, method)(input_ids, attention_mask, index=targets, **kw)
    if word_scores is not None:
        scores = word_scores(scores)
    os.makedirs(os.path.join(output_dir, folder), exist_ok=True)
    for k, lines in rationale_lines(doc_ids, scores, ks).items():
        with open(os.path.join(output_dir, folder, "identifier_results_%d.json" % k), "a") as f:
            f.write("".join(line + "\n" for line in lines))
    return scores
