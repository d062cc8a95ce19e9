"""Generate tests/golden/reference_golden_ranking.pt by running the REAL reference ranking-metric code
(utils/metrics_encoder_decoder.py:29-86 get_metrics_universal, model/encoder_decoder.py:155-164 re-batching via
utils/tensor.py padded_stack) from /root/reference.   python tests/golden/make_golden_ranking.py

Every line of these reference functions is plain torch / numpy: nothing is shimmed on this path except the import of
torch_geometric names that utils/metrics_lightgcn.py pulls in at module import (oracle/ref_shims.py).
"""
import os
import sys

import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
sys.path.insert(0, REF)
for name in [m for m in sys.modules if m == "tests" or m.startswith("tests.")]:
    del sys.modules[name]

from utils.metrics_encoder_decoder import get_metrics_universal  # noqa: E402
from utils.tensor import padded_stack  # noqa: E402


def case(seed, n_users, n_items, cands_lo, cands_hi, n_pos, n_excl, k, n_excl_lists=1):
    g = t.Generator().manual_seed(seed)
    # label edges: every user gets a ragged candidate list (like LinkNeighborLoader batches: ~11 pos + 33 neg per user)
    counts = t.randint(cands_lo, cands_hi + 1, (n_users,), generator=g)
    lab_u = t.repeat_interleave(t.arange(n_users), counts)
    lab_i = t.randint(0, n_items, (int(counts.sum()),), generator=g)
    perm = t.randperm(lab_u.numel(), generator=g)                      # label edges arrive unsorted
    edge_label_index = t.stack([lab_u[perm], lab_i[perm]])
    scores = t.randn(lab_u.numel(), generator=g)                       # decoder logits, one per label edge
    # model.infer's re-batching (model/encoder_decoder.py:160-164)
    users = edge_label_index[0].unique(sorted=True)
    out = padded_stack([scores[edge_label_index[0] == u] for u in users], value=-(1 << 50))
    edge_index = t.stack([t.randint(0, n_users, (n_pos,), generator=g), t.randint(0, out.shape[1], (n_pos,), generator=g)])
    excl = [t.stack([t.randint(0, n_users, (n_excl,), generator=g), t.randint(0, out.shape[1], (n_excl,), generator=g)])
            for _ in range(n_excl_lists)]
    recall, precision, ndcg = get_metrics_universal(out.clone(), edge_index, edge_label_index, [e.clone() for e in excl], k)
    return dict(scores=scores, edge_label_index=edge_label_index, infer_out=out, edge_index=edge_index, exclude=excl, k=k,
                recall=recall, precision=precision, ndcg=ndcg)


def main():
    cases = [case(1, 12, 50, 8, 20, 60, 30, 5), case(2, 40, 200, 30, 44, 400, 100, 12, n_excl_lists=2),
             case(3, 5, 9, 3, 6, 10, 0, 3, n_excl_lists=0), case(4, 64, 500, 44, 44, 1000, 300, 12)]
    # 1-D model output (a single user): the reference unsqueezes it (utils/metrics_encoder_decoder.py:54-57)
    g = t.Generator().manual_seed(9)
    out1 = t.randn(20, generator=g)
    ei = t.stack([t.zeros(6, dtype=t.long), t.randint(0, 20, (6,), generator=g)])
    eli = t.stack([t.zeros(20, dtype=t.long), t.arange(20)])
    ex = [t.stack([t.zeros(4, dtype=t.long), t.randint(0, 20, (4,), generator=g)])]
    r, p, n = get_metrics_universal(out1.clone(), ei, eli, [e.clone() for e in ex], 4)
    cases.append(dict(scores=out1, edge_label_index=eli, infer_out=out1, edge_index=ei, exclude=ex, k=4, recall=r, precision=p, ndcg=n))
    path = os.path.join(HERE, "reference_golden_ranking.pt")
    t.save(dict(universal=cases), path)
    for c in cases:
        print(f"k={c['k']} out={tuple(c['infer_out'].shape)} recall={c['recall']:.6f} precision={c['precision']:.6f} ndcg={c['ndcg']:.6f}")
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
