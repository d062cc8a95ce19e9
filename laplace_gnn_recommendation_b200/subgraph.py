"""Sub-graph batch assembly on the device: what the reference does per user in Python --
``GraphDataset.__getitem__`` (data/dataset.py:39-182), ``fetch_n_hop_neighbourhood`` (:258-286),
``remap_edges_to_start_from_zero`` (:236-244) -- followed by PyG's ``Batch`` collation, for a whole batch of root users
at once and without leaving the GPU.

Per root user r (reference line numbers in data/dataset.py):
    positives            (r, a) for every article a of r, adjacency order                                   :41-47
    sampled positives    max(1, floor(len * positive_edges_ratio)) draws with replacement -- or, without
                         randomisation, the smallest and the largest article id                             :50-72
    negatives (train)    int(ratio * n_sampled) articles; ratio = k - 1 when n_sampled <= 1                 :74-96,188-233
    n-hop edges          frontier_0 = {r}; frontier_{i+1} = cut(users(cut(articles(frontier_i))) - explored);
                         every (u, a) of every user u of the frontiers 1 .. n-1                               :258-286
    node buckets         sorted unique user / article ids over positives + negatives + n-hop edges          :135-139
    remap                t.bucketize of both edge lists against the buckets                                  :142-152
    labels               ones for the sampled positives, zeros for the negatives                             :155-161
and the batch = the disjoint union of the per-root graphs (node ids offset by the preceding roots' bucket sizes), plus the
reverse edge type.  With ``randomization=False`` (the mode the reference's own tests/test_dataset.py uses) and fan-outs below
``num_neighbors`` every integer of the result is identical to the reference's; the random choices (sampling with
replacement, fan-out cuts, random negatives) use the device generator -- same distributions, another random stream.

The index-heavy steps are hand-written kernels (csrc/sampler.cu: ``lgb_segment_expand``, ``lgb_bucketize_segmented``); sorting,
unique and prefix sums are library calls.  The result carries the per-batch CSR pair (``graphs``), which
``Encoder_Decoder_Model.forward(..., graphs=...)`` takes instead of building it again.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import check, ptr, stream
from .hetero import EDGE_KEY, NODE_ITEM, NODE_USER, REV_EDGE_KEY


def _csr_from_pairs(src: torch.Tensor, dst: torch.Tensor, n_src: int):
    """Adjacency lists ``{s: [d in edge order]}`` as (ptr [n_src+1], idx [E]) -- create_adj_list's ordering (stable)."""
    order = torch.sort(src, stable=True).indices
    ptr_ = torch.zeros(n_src + 1, dtype=torch.int64, device=src.device)
    ptr_[1:] = torch.cumsum(torch.bincount(src, minlength=n_src), 0)
    return ptr_, dst[order].contiguous()


class SubgraphSampler:
    """Device-resident replacement of ``GraphDataset`` + ``DataLoader`` collation for the training split."""

    def __init__(self, edge_index: torch.Tensor, x_user: torch.Tensor, x_article: torch.Tensor, n_hop_neighbors: int = 3,
                 num_neighbors: int = 64, positive_edges_ratio: float = 0.5, negative_edges_ratio: float = 3.0, k: int = 12,
                 randomization: bool = True, generator: Optional[torch.Generator] = None):
        _lib.require_cuda(edge_index, x_user, x_article)
        self.device = edge_index.device
        self.U, self.A = x_user.shape[0], x_article.shape[0]
        u, a = _lib.i64c(edge_index[0]), _lib.i64c(edge_index[1])
        self.u2a_ptr, self.u2a_idx = _csr_from_pairs(u, a, self.U)          # self.users[idx]    (data/dataset.py:30)
        self.a2u_ptr, self.a2u_idx = _csr_from_pairs(a, u, self.A)          # self.articles[idx] (:29)
        self.x_user, self.x_article = x_user, x_article
        self.id_max = int(a.max()) if a.numel() else 0                        # t.max(all_edges, dim=1)[0][1] (:196)
        self.E = int(u.numel())
        self.n_hops, self.fanout = int(n_hop_neighbors), int(num_neighbors)
        self.pos_ratio, self.neg_ratio, self.k = float(positive_edges_ratio), float(negative_edges_ratio), int(k)
        self.randomization, self.generator = bool(randomization), generator

    # ---- kernels -------------------------------------------------------------------------------------------------------
    def _expand(self, ptr_: torch.Tensor, idx: torch.Tensor, nodes: torch.Tensor):
        """(position in ``nodes``, neighbour) for every adjacency entry of every node, adjacency order kept."""
        deg = ptr_[nodes + 1] - ptr_[nodes]
        off = torch.zeros(nodes.numel() + 1, dtype=torch.int64, device=self.device)
        off[1:] = torch.cumsum(deg, 0)
        total = int(off[-1])
        pos = torch.empty(total, dtype=torch.int64, device=self.device)
        nbr = torch.empty(total, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_segment_expand(ptr(ptr_), ptr(idx), ptr(nodes), ptr(off), nodes.numel(), total, ptr(pos), ptr(nbr),
                                                 stream()), "segment_expand")
        _lib.count_launch()
        return pos, nbr

    def _bucketize(self, values: torch.Tensor, seg: torch.Tensor, buckets: torch.Tensor, bucket_ptr: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(values)
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_bucketize_segmented(ptr(values), ptr(seg), values.numel(), ptr(buckets), ptr(bucket_ptr), ptr(out),
                                                      stream()), "bucketize_segmented")
        _lib.count_launch()
        return out

    # ---- helpers on (segment, value) lists --------------------------------------------------------------------------------
    def _cut(self, seg: torch.Tensor, val: torch.Tensor, n_seg: int):
        """shuffle_and_cut per segment (:289-293): at most ``fanout`` random elements of each segment's list (duplicates are
        separate candidates, as in random.sample over the flattened list); lists within the limit are kept whole."""
        counts = torch.bincount(seg, minlength=n_seg)
        if val.numel() == 0 or int(counts.max()) <= self.fanout:
            return seg, val
        key = torch.rand(val.numel(), device=self.device, generator=self.generator)
        key = torch.where(counts[seg] > self.fanout, key, torch.zeros_like(key))     # short lists keep their order
        order = torch.sort(key, stable=True).indices
        order = order[torch.sort(seg[order], stable=True).indices]                   # grouped by segment, random inside long ones
        s, v = seg[order], val[order]
        start = torch.cumsum(counts, 0) - counts
        keep = torch.arange(s.numel(), device=self.device) - start[s] < self.fanout
        return s[keep], v[keep]

    @staticmethod
    def _unique_pairs(seg: torch.Tensor, val: torch.Tensor, span: int):
        key = torch.unique(seg * span + val, sorted=True)
        return key // span, key % span

    # ---- the batch --------------------------------------------------------------------------------------------------------
    def sample(self, roots: torch.Tensor) -> Dict[str, object]:
        """Batch of the roots' sub-graphs.  Returns a dict with ``x_dict``, ``edge_index_dict``, ``edge_label_index``,
        ``edge_label`` (the four things utils/get_info.py:select_properties hands to the model), ``graphs`` (edge type ->
        DeviceCSR, pass as ``Encoder_Decoder_Model.forward(..., graphs=...)``), ``n_id`` (global ids of the batch nodes per
        type) and ``ptr`` (node offsets of every root's graph per type, like PyG's ``Batch.ptr``)."""
        dev, g = self.device, self.generator
        roots = _lib.i64c(roots.to(dev))
        B = roots.numel()
        span_u, span_a = self.U + 1, self.A + 1
        seg_ids = torch.arange(B, device=dev)

        # positives of every root (:41-47)
        ppos, part = self._expand(self.u2a_ptr, self.u2a_idx, roots)
        n_pos = torch.bincount(ppos, minlength=B)
        if int(n_pos.min()) == 0:
            raise RuntimeError("SubgraphSampler: a root user without any article (the reference indexes an empty list there)")
        pstart = torch.cumsum(n_pos, 0) - n_pos

        # sampled positives (:50-72)
        if self.randomization:
            cut = torch.clamp((n_pos.double() * self.pos_ratio).floor().long(), min=1)
            sseg = torch.repeat_interleave(seg_ids, cut)
            draw = (torch.rand(sseg.numel(), device=dev, generator=g, dtype=torch.float64) * n_pos[sseg]).long()
            draw = torch.minimum(draw, n_pos[sseg] - 1)
            sart = part[pstart[sseg] + draw]
        else:   # indices of the smallest and of the largest article id (first occurrence each)
            big = torch.iinfo(torch.int64).max
            key = ppos * (2 * span_a)
            lo = torch.full((B,), big, dtype=torch.int64, device=dev).scatter_reduce_(0, ppos, key + part, "amin", include_self=True)
            hi = torch.full((B,), -1, dtype=torch.int64, device=dev).scatter_reduce_(0, ppos, key + part, "amax", include_self=True)
            sseg = torch.repeat_interleave(seg_ids, 2)
            sart = torch.stack([lo - seg_ids * (2 * span_a), hi - seg_ids * (2 * span_a)], dim=1).reshape(-1)
            cut = torch.full((B,), 2, dtype=torch.int64, device=dev)

        # negatives, training split (:74-96, get_negative_edges_random :188-233)
        ratio = torch.where(cut <= 1, torch.full_like(cut, self.k - 1).double(), torch.full_like(cut, 0).double() + self.neg_ratio)
        n_neg = (ratio * cut.double()).long()
        if not self.randomization:
            nseg, nart = seg_ids.clone(), torch.full((B,), self.id_max, dtype=torch.int64, device=dev)     # t.tensor([id_max])
        else:
            cheap = self.E / n_neg.clamp(min=1).double() > 100          # :198 -- unlikely to hit a positive: no filtering
            n_cheap = torch.where(cheap, n_neg, torch.zeros_like(n_neg))
            nseg = torch.repeat_interleave(seg_ids, n_cheap)
            nart = (torch.rand(nseg.numel(), device=dev, generator=g, dtype=torch.float64) * self.id_max).long().clamp(max=max(self.id_max - 1, 0))
            for b in (~cheap).nonzero().view(-1).tolist():              # toy graphs only (E / n_neg <= 100): distinct random
                cand = torch.ones(self.id_max + 1, dtype=torch.bool, device=dev)      # articles outside the sampled positives
                cand[sart[sseg == b]] = False                                          # (:211-231)
                cand = cand.nonzero().view(-1)
                pick = cand[torch.randperm(cand.numel(), device=dev, generator=g)][: int(n_neg[b])]
                nseg = torch.cat([nseg, torch.full((pick.numel(),), b, dtype=torch.int64, device=dev)])
                nart = torch.cat([nart, pick])

        # n-hop edges (:258-286): frontier of users per root; explored users as sorted (root, user) keys
        fseg, fusr = seg_ids.clone(), roots.clone()
        explored = torch.empty(0, dtype=torch.int64, device=dev)
        hop_seg, hop_usr, hop_art = [], [], []
        for i in range(self.n_hops):
            if fusr.numel() == 0:
                break
            pos_, art_ = self._expand(self.u2a_ptr, self.u2a_idx, fusr)          # create_neighbouring_article_edges for the queue
            explored = torch.unique(torch.cat([explored, fseg * span_u + fusr]))
            if i != 0:
                hop_seg.append(fseg[pos_]); hop_usr.append(fusr[pos_]); hop_art.append(art_)
            aseg, aart = self._cut(fseg[pos_], art_, B)                            # articles_queue
            pos2, usr2 = self._expand(self.a2u_ptr, self.a2u_idx, aart)
            nseg2, nusr2 = self._unique_pairs(aseg[pos2], usr2, span_u)            # set(flatten(...))
            fresh = ~torch.isin(nseg2 * span_u + nusr2, explored)                  # ... - users_explored
            fseg, fusr = self._cut(nseg2[fresh], nusr2[fresh], B)                  # users_queue
        empty = torch.empty(0, dtype=torch.int64, device=dev)
        hseg = torch.cat(hop_seg) if hop_seg else empty
        husr = torch.cat(hop_usr) if hop_usr else empty
        hart = torch.cat(hop_art) if hop_art else empty

        # all_subgraph_edges = positives + n-hop edges; all_touched adds the negatives (:113-131).  Edges are ordered by root.
        proot = roots[ppos]
        sub_seg = torch.cat([ppos, hseg]); sub_u = torch.cat([proot, husr]); sub_a = torch.cat([part, hart])
        order = torch.sort(sub_seg, stable=True).indices
        sub_seg, sub_u, sub_a = sub_seg[order], sub_u[order], sub_a[order]
        lab_seg = torch.cat([sseg, nseg]); lab_u = roots[lab_seg]; lab_a = torch.cat([sart, nart])
        labels = torch.cat([torch.ones(sseg.numel(), device=dev), torch.zeros(nseg.numel(), device=dev)])
        order = torch.sort(lab_seg, stable=True).indices        # per root: its sampled positives, then its negatives
        lab_seg, lab_u, lab_a, labels = lab_seg[order], lab_u[order], lab_a[order], labels[order]

        # node buckets (:135-139) over all_touched_edges, per root, concatenated; Batch-style offsets
        bu_seg, bu = self._unique_pairs(torch.cat([sub_seg, lab_seg]), torch.cat([sub_u, lab_u]), span_u)
        ba_seg, ba = self._unique_pairs(torch.cat([sub_seg, nseg]), torch.cat([sub_a, nart]), span_a)
        uptr = torch.zeros(B + 1, dtype=torch.int64, device=dev); uptr[1:] = torch.cumsum(torch.bincount(bu_seg, minlength=B), 0)
        aptr = torch.zeros(B + 1, dtype=torch.int64, device=dev); aptr[1:] = torch.cumsum(torch.bincount(ba_seg, minlength=B), 0)

        # remap (:142-152) + collation offsets.  (A sampled positive is one of the root's articles, hence in the buckets; the
        # reference bucketizes the label edges against the same buckets.)
        e_u = self._bucketize(sub_u, sub_seg, bu, uptr) + uptr[sub_seg]
        e_a = self._bucketize(sub_a, sub_seg, ba, aptr) + aptr[sub_seg]
        l_u = self._bucketize(lab_u, lab_seg, bu, uptr) + uptr[lab_seg]
        l_a = self._bucketize(lab_a, lab_seg, ba, aptr) + aptr[lab_seg]
        edge_index = torch.stack([e_u, e_a])
        edge_label_index = torch.stack([l_u, l_a])

        from .hetero import build_edge_csr
        n_u, n_a = int(uptr[-1]), int(aptr[-1])
        g_buys = build_edge_csr(edge_index, n_u, n_a, deferred=True)            # dst = article
        graphs = {EDGE_KEY: g_buys, REV_EDGE_KEY: g_buys.transpose()}
        return {
            "x_dict": {NODE_USER: self.x_user[bu], NODE_ITEM: self.x_article[ba]},
            "edge_index_dict": {EDGE_KEY: edge_index, REV_EDGE_KEY: edge_index.flip(0).contiguous()},
            "edge_label_index": edge_label_index, "edge_label": labels.long(), "graphs": graphs,
            "n_id": {NODE_USER: bu, NODE_ITEM: ba}, "ptr": {NODE_USER: uptr, NODE_ITEM: aptr},
        }
