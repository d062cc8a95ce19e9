"""Golden vectors for the device-side sub-graph batch assembly (SubgraphSampler) produced by the REAL reference code:
`/root/reference/data/dataset.py::GraphDataset.__getitem__` (randomization=False, train=True: the mode the reference's own
tests/test_dataset.py uses), run per root user on the graphs of tests/data_generator.py and on a random graph; the per-user
HeteroData results are stored as plain tensors.  The third-party names the module imports resolve to the package's alias
layer (`install_aliases`: HeteroData is the only one the function touches).

    python tests/golden/make_golden_subgraph.py        # writes tests/golden/reference_golden_subgraph.pt
"""
import os
import sys
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)


def reference_items(edge_index, x_user, x_article, roots, n_hops, num_neighbors, pos_ratio, neg_ratio, k):
    from laplace_gnn_recommendation_b200 import aliases
    aliases.install_aliases(force=True)
    tgd = sys.modules["torch_geometric.data"]
    if not hasattr(tgd, "InMemoryDataset"):
        tgd.InMemoryDataset = type("InMemoryDataset", (), {})
    import importlib
    ds_mod = importlib.import_module("data.dataset")
    from utils.constants import Constants
    graph = aliases.HeteroData()
    graph[Constants.node_user].x = x_user
    graph[Constants.node_item].x = x_article
    graph[Constants.edge_key].edge_index = edge_index
    U, A = x_user.shape[0], x_article.shape[0]
    users = [edge_index[1][edge_index[0] == u].tolist() for u in range(U)]          # create_adj_list ordering
    articles = [edge_index[0][edge_index[1] == a].tolist() for a in range(A)]
    ds = object.__new__(ds_mod.GraphDataset)
    ds.graph, ds.users, ds.articles, ds.matchers, ds.train, ds.randomization = graph, users, articles, None, True, False
    ds.config = types.SimpleNamespace(positive_edges_ratio=pos_ratio, negative_edges_ratio=neg_ratio, k=k, n_hop_neighbors=n_hops,
                                      num_neighbors=num_neighbors)
    out = []
    for r in roots:
        d = ds[int(r)]
        e = d[Constants.edge_key]
        out.append(dict(x_user=d[Constants.node_user].x.clone(), x_article=d[Constants.node_item].x.clone(),
                        edge_index=e.edge_index.clone(), edge_label_index=e.edge_label_index.clone(), edge_label=e.edge_label.clone(),
                        rev_edge_index=d[Constants.rev_edge_key].edge_index.clone()))
    return out


def cases():
    gen = torch.Generator().manual_seed(3)
    out = {}
    # the two manual graphs of the reference's tests/data_generator.py:129-159
    e = torch.tensor([[0, 0, 0, 1, 1, 2, 2], [0, 2, 4, 1, 5, 3, 0]])
    out["manual_random"] = dict(edge_index=e, x_user=torch.arange(6.).view(3, 2), x_article=torch.arange(30.).view(6, 5), roots=[0, 1, 2])
    e = torch.tensor([[0, 0, 0, 0, 1, 2, 3, 4], [0, 1, 2, 3, 0, 1, 2, 3]])
    out["manual_star"] = dict(edge_index=e, x_user=torch.arange(30.).view(5, 6), x_article=torch.arange(16.).view(4, 4), roots=[0, 3, 4, 1])
    U, A, E = 40, 60, 160
    e = torch.stack([torch.randint(0, U, (E,), generator=gen), torch.randint(0, A, (E,), generator=gen)])
    e[0, :U] = torch.arange(U)                                   # every user has an article
    out["random_sparse"] = dict(edge_index=e, x_user=torch.randn(U, 3, generator=gen), x_article=torch.randn(A, 4, generator=gen),
                                roots=[5, 0, 17, 39, 5, 22])
    return out


def main():
    golden = {}
    for name, c in cases().items():
        for n_hops in (1, 2, 3):
            cfg = dict(n_hops=n_hops, num_neighbors=1000, pos_ratio=0.5, neg_ratio=3.0, k=12)
            items = reference_items(c["edge_index"], c["x_user"], c["x_article"], c["roots"], **cfg)
            golden[f"{name}/hops{n_hops}"] = dict(inputs=c, config=cfg, items=items)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_subgraph.pt")
    torch.save(golden, path)
    print(path, {k: len(v["items"]) for k, v in golden.items()})


if __name__ == "__main__":
    main()
