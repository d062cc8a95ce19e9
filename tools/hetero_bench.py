"""Config 4 of BASELINE.json: the hetero encoder-decoder training step on LinkNeighborLoader-sized batches cut to the
H&M shape: sizes S/M/L, input widths 84/76 -> hidden 128 -> out 64, 2 SAGE layers, batch norm, BCE loss.

Reports per size: ms per training step (forward + loss + backward, CUDA events), the share and achieved GB/s of the
neighbour-aggregation kernels against seg_bytes = E*(4 + F*4) + (N_dst+1)*4 + N_dst*F*4 per edge type / layer / direction,
and (size S only) the CPU oracle's step time.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _common  # noqa: E402
import laplace_gnn_recommendation_b200 as lg  # noqa: E402
from laplace_gnn_recommendation_b200 import hetero  # noqa: E402
from laplace_gnn_recommendation_b200.csr import DeviceCSR  # noqa: E402

SIZES = {  # E_sub per edge type, N_customer, N_article, label edges
    "S": (72_000, 3_000, 40_000, 1_100),
    "M": (400_000, 16_000, 90_000, 5_800),
    "L": (3_000_000, 130_000, 105_000, 46_000),
}
FC, FA, HID, OUT = 84, 76, 128, 64


def make_batch(size, dev, seed=0):
    E, Nc, Na, L = SIZES[size]
    gen = torch.Generator().manual_seed(seed)
    x = {"customer": torch.randn(Nc, FC, generator=gen).to(dev), "article": torch.randn(Na, FA, generator=gen).to(dev)}
    e = torch.stack([torch.randint(0, Nc, (E,), generator=gen), torch.randint(0, Na, (E,), generator=gen)]).to(dev)
    ei = {hetero.EDGE_KEY: e, hetero.REV_EDGE_KEY: e.flip(0).contiguous()}
    eli = torch.stack([torch.randint(0, Nc, (L,), generator=gen), torch.randint(0, Na, (L,), generator=gen)]).to(dev)
    y = (torch.rand(L, generator=gen) > 0.75).float().to(dev)
    return x, ei, eli, y


def seg_bytes(E, F, n_dst):
    return E * (4 + F * 4) + (n_dst + 1) * 4 + n_dst * F * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="S,M,L")
    ap.add_argument("--aggr", default="add")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--project-first", action="store_true", help="SAGEConv.project_first: narrow layers aggregate after lin_l's weight")
    a = ap.parse_args()
    dev = _common.device()
    if _common.DRYRUN:
        for k, v in list(SIZES.items()):
            SIZES[k] = tuple(max(x // 200, 16) for x in v)
        a.steps = 1
    metadata = (["customer", "article"], [hetero.EDGE_KEY, hetero.REV_EDGE_KEY])
    for size in a.sizes.split(","):
        x, ei, eli, y = make_batch(size, dev)
        torch.manual_seed(0)
        model = lg.Encoder_Decoder_Model(
            encoder_layers=lg.get_SAGEConv_layers(2, HID, OUT, a.aggr), decoder_layers=lg.get_linear_layers(2, 2 * OUT, HID, 1),
            feature_info={}, metadata=metadata, embedding=False, heterogeneous_prop_agg_type="sum", batch_normalize=True,
            p_dropout_edges=None, p_dropout_features=None).to(dev)
        if a.project_first:
            for m in model.modules():
                if isinstance(m, hetero.SAGEConv):
                    m.project_first = True
        lossf = torch.nn.BCEWithLogitsLoss()

        def step():
            model.zero_grad(set_to_none=True)
            loss = lossf(model(dict(x), ei, eli), y)
            loss.backward()
            return loss
        events = []
        orig = DeviceCSR.spmm

        def timed(self, *args, **kw):
            e0, e1 = _common.Event(), _common.Event()
            e0.record(); out = orig(self, *args, **kw); e1.record()
            events.append((e0, e1, self.nnz, self.n_rows, args[0].shape[1]))
            return out
        for _ in range(5):
            step()
        _common.sync()
        DeviceCSR.spmm = timed
        t0, t1 = _common.Event(), _common.Event()
        t0.record()
        for _ in range(a.steps):
            loss = step()
        t1.record()
        _common.sync()
        DeviceCSR.spmm = orig
        ms = t0.elapsed_time(t1) / a.steps
        agg_ms = sum(e0.elapsed_time(e1) for e0, e1, *_ in events) / a.steps
        agg_gb = sum(seg_bytes(nnz, F, rows) for _, _, nnz, rows, F in events) / a.steps / 1e9
        line = {"size": size, "aggr": a.aggr, "project_first": a.project_first, "E_sub": SIZES[size][0], "ms_per_step": ms, "aggregation_ms": agg_ms,
                "aggregation_share": agg_ms / ms, "aggregation_GBps": agg_gb / (agg_ms * 1e-3), "launches_per_step": len(events) // a.steps,
                "loss": float(loss)}
        if size == "S":
            from oracle import hetero_oracle as ho
            xc = {k: v.cpu() for k, v in x.items()}
            eic = {k: v.cpu() for k, v in ei.items()}
            sd = {k: v.detach().cpu().requires_grad_(v.dtype.is_floating_point) for k, v in model.state_dict().items()}
            layers = [{et: dict(w_l=sd[f"encoder.layers.{li}.{'__'.join(et)}.lin_l.weight"], b_l=sd[f"encoder.layers.{li}.{'__'.join(et)}.lin_l.bias"],
                                w_r=sd[f"encoder.layers.{li}.{'__'.join(et)}.lin_r.weight"]) for et in metadata[1]} for li in range(2)]

            def cpu_step():
                z = ho.hetero_encoder(xc, eic, layers, a.aggr, "sum", metadata[1])
                zu = ho.batch_norm_train(z["customer"], sd["encoder_layer_norm_customer.weight"], sd["encoder_layer_norm_customer.bias"])
                zi = ho.batch_norm_train(z["article"], sd["encoder_layer_norm_article.weight"], sd["encoder_layer_norm_article.bias"])
                lin = [(sd[f"decoder.layers.{i}.weight"], sd[f"decoder.layers.{i}.bias"]) for i in range(2)]
                l = ho.bce_with_logits(ho.edge_decoder_mlp(zu, zi, eli.cpu(), lin), y.cpu())
                l.backward()
                return l
            torch.set_num_threads(os.cpu_count() or 1)
            cpu_step()
            c0 = time.perf_counter()
            for _ in range(3):
                cl = cpu_step()
            line["cpu_oracle_ms_per_step"] = (time.perf_counter() - c0) / 3 * 1e3
            line["cpu_cores"] = torch.get_num_threads()
            line["cpu_loss"] = float(cl)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
