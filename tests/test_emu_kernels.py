"""Kernel-logic tier (runs WITHOUT a GPU): the parity tests of tests/test_gpu_*.py executed against the CPU emulation
of liblaplace_b200 -- the unmodified csrc/*.cu kernel sources compiled with g++ against a warp-lockstep CUDA emulator
(tests/emu/).  Same test bodies, same oracle, same tolerances; only the ``cuda_dev`` fixture differs.

What this tier proves: index arithmetic, warp-shuffle data flow, split plans, fused epilogues, reductions and the Python
host logic above them are right, and no kernel issues a divergent or partially-exited warp collective (the emulator turns
those into launch errors).  What it cannot prove: anything about memory-model races, performance, or PTX-level behaviour
(cache hints, red.global.add.v4, cp.async are replaced by their plain meaning) -- that is what `pytest -m gpu` is for.
"""
import platform

import pytest
import torch

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's context switch is x86-64 only")

from tests.emu.harness import emulated  # noqa: E402
import tests.test_gpu_hetero as TH  # noqa: E402
import tests.test_gpu_lightgcn as TL  # noqa: E402
import tests.test_gpu_zz_unmeasured as TZ  # noqa: E402


@pytest.fixture
def cuda_dev():
    """Overrides conftest's fixture inside this module: the 'device' is the CPU, served by the emulated library."""
    with emulated() as dev:
        yield dev


# ---- LightGCN path -----------------------------------------------------------------------------------------------
test_csr_build_bit_exact = TL.test_csr_build_bit_exact
test_csr_reference_fixture_and_gcn_norm_bit_exact = TL.test_csr_reference_fixture_and_gcn_norm_bit_exact
test_csr_build_rejects_out_of_range_indices = TL.test_csr_build_rejects_out_of_range_indices
test_spmm_vs_oracle = TL.test_spmm_vs_oracle
test_spmm_wide_slice_variant_vs_oracle = TZ.test_spmm_wide_slice_variant_vs_oracle
test_spmm_64bit_index_family_vs_oracle = TZ.test_spmm_64bit_index_family_vs_oracle
test_spmm_256bit_gathers_d128_vs_oracle = TZ.test_spmm_256bit_gathers_d128_vs_oracle
test_spmm_fused_epilogue_and_degree_order = TL.test_spmm_fused_epilogue_and_degree_order
test_spmm_sweep_order_bit_identical = TL.test_spmm_sweep_order_bit_identical


@pytest.mark.parametrize("d", [64, 32, 48, 20])                 # the three filtered kernel shapes (256-bit, d/4 <= 8, d/4 <= 16)
def test_spmm_rowsparse_matches_dense(cuda_dev, d):
    TL.test_spmm_rowsparse_matches_dense(cuda_dev, d)          # (the GPU tier also runs the dense fallbacks: d = 128, d = 6)



@pytest.mark.parametrize("variant,d", [(0, 64), (23, 64)])
def test_spmm_fused_stage2_bit_identical(cuda_dev, variant, d, monkeypatch):
    TL.test_spmm_fused_stage2_bit_identical(cuda_dev, variant, d, monkeypatch)       # (the GPU tier runs every combination)


test_lightgcn_small_batch_rowsparse_backward = TL.test_lightgcn_small_batch_rowsparse_backward
test_spmm_rowsparse_random_shapes = TL.test_spmm_rowsparse_random_shapes
test_lightgcn_against_reference_golden = TL.test_lightgcn_against_reference_golden
test_lightgcn_against_oracle = TL.test_lightgcn_against_oracle
test_bpr_against_reference_golden = TL.test_bpr_against_reference_golden
test_module_contract = TL.test_module_contract
test_sample_mini_batch_bit_exact_vs_reference_golden = TL.test_sample_mini_batch_bit_exact_vs_reference_golden
test_structured_negative_sampling_bit_exact = TL.test_structured_negative_sampling_bit_exact
test_topk_against_reference_golden = TL.test_topk_against_reference_golden
test_topk_against_oracle = TL.test_topk_against_oracle
test_evaluation_matches_oracle = TZ.test_evaluation_matches_oracle
test_fused_adam_matches_torch_adam = TZ.test_fused_adam_matches_torch_adam
test_training_loop_with_fused_step_and_fused_adam = TZ.test_training_loop_with_fused_step_and_fused_adam

# ---- hetero encoder-decoder path ---------------------------------------------------------------------------------
test_aggregate_fwd_bwd_vs_oracle = TH.test_aggregate_fwd_bwd_vs_oracle
test_encoder_decoder_against_reference_golden = TH.test_encoder_decoder_against_reference_golden
test_infer_rebatch_matches_oracle = TH.test_infer_rebatch_matches_oracle
test_decoder_concat_and_dot = TH.test_decoder_concat_and_dot
test_decoder_node_projection_inference_form = TH.test_decoder_node_projection_inference_form
test_hetero_fan_in_three_edge_types = TH.test_hetero_fan_in_three_edge_types
test_subgraph_sampler_against_reference_golden = TH.test_subgraph_sampler_against_reference_golden
test_subgraph_sampler_random_mode_and_model_step = TH.test_subgraph_sampler_random_mode_and_model_step
test_model_gradients_same_with_and_without_wgrad_kernel = TH.test_model_gradients_same_with_and_without_wgrad_kernel


@pytest.mark.parametrize("N,n_in,n_out", [(1300, 84, 128), (777, 76, 64), (513, 6, 5), (1, 4, 4), (900, 130, 70)])
def test_linear_wgrad_vs_torch(cuda_dev, N, n_in, n_out):
    TH.test_linear_wgrad_vs_torch(cuda_dev, N, n_in, n_out)

# full-size property tests (B = 200 003 BPR triples; a 400 k-edge, 84-wide hetero batch): seconds under the emulator
test_bpr_full_split_size = TL.test_bpr_full_split_size
test_batch_sized_aggregation_properties = TH.test_batch_sized_aggregation_properties


@pytest.mark.parametrize("d", [32, 64, 128])
def test_spmm_kernel_variants_agree(cuda_dev, d):
    TZ.test_spmm_kernel_variants_agree(cuda_dev, d, n=700, nnz=14000)   # smaller than on the GPU: 28 variants x 4 launches


def test_spmm_empty_and_single_heavy_row(cuda_dev):
    """tests/test_gpu_lightgcn.py::test_spmm_empty_and_ragged without its "CPU tensors are refused" half (under the
    emulator CPU tensors ARE the device tensors)."""
    d = 64
    e = torch.empty(0, dtype=torch.long)
    g = TL.DeviceCSR.from_coo(e, e, 10, 10)
    assert torch.count_nonzero(g.spmm(torch.randn(10, d))) == 0
    n, nnz = 40, 5000
    row = torch.full((nnz,), 7); col = torch.randint(0, n, (nnz,), generator=torch.Generator().manual_seed(3))
    g = TL.DeviceCSR.from_coo(row, col, n, n, chunk=128)
    X = torch.randn(n, d, generator=torch.Generator().manual_seed(4))
    TL.spmm_close(g.spmm(X), g.rowptr.long(), g.colidx.long(), None, X)
    with pytest.raises(RuntimeError):
        g.spmm(torch.randn(n + 1, d))


def test_sharded_engine_single_rank(cuda_dev):
    """tests/test_gpu_lightgcn.py::test_sharded_engine_single_rank_matches_oracle with dist.CudaOps on the emulated library
    (multi-rank runs of the same ops object: tests/test_dist_gloo.py::test_sharded_step_with_emulated_kernels)."""
    from laplace_gnn_recommendation_b200.dist import ShardedLightGCN
    from tests.test_dist_gloo import make_emu_ops, make_problem, single_process_reference
    for K, schedule in ((3, "layer"), (1, "chains"), (2, "chains")):
        pb = make_problem(seed=K, U=500, I=120, E=9000, d=64, K=K, B=256)
        eng = ShardedLightGCN(pb["U"], pb["I"], pb["d"], K, pb["users"], pb["items"], cuda_dev, ops=make_emu_ops(),
                              init_tables=(pb["Wu"], pb["Wi"]), rank=0, world=1, schedule=schedule, max_batch=256)
        loss = eng.fused_step(pb["u"], pb["p"], pb["n"], pb["lam"])
        o_loss, o_gu, o_gi, o_uf, o_if = single_process_reference(pb)
        TL.close(loss, o_loss)
        TL.close(eng.E_f[: pb["U"]], o_uf); TL.close(eng.E_f[pb["U"]:], o_if)
        TL.close(eng.grad[: pb["U"]], o_gu, atol=1e-9); TL.close(eng.grad[pb["U"]:], o_gi, atol=1e-9)


def test_emulator_reports_undefined_warp_collectives():
    """The emulator itself: a launch error must surface through the library's own error path."""
    from tests.emu.selftest import run_selftest
    run_selftest()


@pytest.mark.parametrize("world,n_floats", [(8, 105_542 * 64), (3, 1000), (4, 8), (16, 4 * 16 * 3 + 4)])
def test_item_block_exchange_kernels(world, n_floats):
    """csrc/peer.cu (the NCCL-free exchange of the replicated item block): every rank reduces ITS slice over all peers and
    republishes it to all peers.  Peers are host buffers here and the NVSwitch multicast address is an emulated key, so this
    checks the slice arithmetic and the kernels' data flow, not the multimem PTX itself."""
    import ctypes as C
    from tests.emu.harness import load_emu
    lib = load_emu()
    gen = torch.Generator().manual_seed(world)
    for kind in ("peer", "multimem"):
        bufs = [torch.randn(n_floats, generator=gen) for _ in range(world)]
        want = torch.stack(bufs).sum(0) if n_floats else torch.zeros(0)
        if kind == "peer":
            arr = (C.c_uint64 * world)(*[b.data_ptr() for b in bufs])
            for rank in range(world):          # ranks run one after the other: their slices are disjoint
                assert lib.lgb_peer_allreduce_f32(arr, n_floats, rank, world, None) == 0, lib.lgb_last_error()
        else:
            key = torch.empty(max(n_floats, 4))   # stands for the multicast mapping: only its address range is used
            ptrs = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
            lib.emu_multicast_bind(C.c_void_p(key.data_ptr()), C.c_size_t(key.numel() * 4), world, ptrs)
            for rank in range(world):
                assert lib.lgb_multimem_allreduce_f32(key.data_ptr(), n_floats, rank, world, None) == 0, lib.lgb_last_error()
        for b in bufs:
            torch.testing.assert_close(b, want, rtol=1e-6, atol=1e-6)
            assert torch.equal(b, bufs[0])      # one reducer per slice: bit-identical on every rank
    # the entry point the engine ships (lgb_exchange_allreduce_f32: barriers + the same slice loop in one launch), at a byte
    # offset inside a larger arena; the barriers are skipped because the ranks run one after the other here
    from laplace_gnn_recommendation_b200._lib import LgbExchange
    lib.lgb_exchange_allreduce_f32.restype = C.c_int
    lib.lgb_exchange_allreduce_f32.argtypes = [C.POINTER(LgbExchange), C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    n4 = n_floats // 4 * 4
    for kind in ("peer", "multimem"):
        arenas = [torch.randn(64 + n4 + 64, generator=gen) for _ in range(world)]
        before = [a.clone() for a in arenas]
        want = torch.stack([a[64:64 + n4] for a in arenas]).sum(0)
        key = torch.empty(64 + n4 + 64)
        if kind == "multimem":
            ptrs = (C.c_void_p * world)(*[a.data_ptr() for a in arenas])
            lib.emu_multicast_bind(C.c_void_p(key.data_ptr()), C.c_size_t(key.numel() * 4), world, ptrs)
        for rank in range(world):
            x = LgbExchange()
            x.multicast_base = key.data_ptr() if kind == "multimem" else None
            for r in range(world):
                x.peer_base[r] = arenas[r].data_ptr()
            x.rank, x.world, x.n_channels = rank, world, 3
            assert lib.lgb_exchange_allreduce_f32(C.byref(x), 64 * 4, n4, 1, 1, None) == 0, lib.lgb_last_error()
        for a, b in zip(arenas, before):
            torch.testing.assert_close(a[64:64 + n4], want, rtol=1e-6, atol=1e-6)
            assert torch.equal(a[:64], b[:64]) and torch.equal(a[64 + n4:], b[64 + n4:])    # nothing outside the exchanged range moves
            assert torch.equal(a[64:64 + n4], arenas[0][64:64 + n4])


def test_reference_driver_loop_runs_in_both_styles():
    """tools/train_lightgcn.py = the reference's `train()` loop (run_pipeline_lightgcn.py:76-222) on this package, dry-run
    on the emulated kernels: the reference's verbatim call sequence (autograd + torch Adam) and the fused form
    (fused_step + FusedAdam) must walk the same trajectory -- splits, sampler, losses, evaluation metrics, candidate dump."""
    import json
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(repo, "tools", "train_lightgcn.py"), "--style", "both"],
                       env=dict(os.environ, LGB_TOOLS_DRYRUN="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    outs = {j["style"]: j for j in map(json.loads, [ln for ln in r.stdout.splitlines() if ln.startswith("{")])}
    a, b = outs["reference"], outs["fused"]
    assert a["candidates_shape"] == b["candidates_shape"] and len(a["log"]) == len(b["log"]) > 0
    for x, y in zip(a["log"], b["log"]):
        for key in ("train_loss", "val_loss", "recall", "precision", "ndcg"):
            assert x[key] == pytest.approx(y[key], rel=1e-4, abs=1e-6), key
    for key in ("loss", "recall", "precision", "ndcg"):
        assert a["test"][key] == pytest.approx(b["test"][key], rel=1e-4, abs=1e-6)


def test_autotune_picks_a_checked_variant(cuda_dev, monkeypatch):
    """DeviceCSR.autotune: every candidate is result-checked against the default before it may win; the choice sticks to
    the graph; LightGCN.autotune / ShardedLightGCN.autotune drive it.  (Wall-clock stands in for CUDA events here.)"""
    import time
    from laplace_gnn_recommendation_b200 import csr
    from laplace_gnn_recommendation_b200.dist import ShardedLightGCN
    from tests.test_dist_gloo import make_emu_ops, make_problem

    def wall(fn, reps, device):
        t = time.perf_counter(); fn(); return (time.perf_counter() - t) * 1e3
    monkeypatch.setattr(csr, "_time_ms", wall)
    row, col = TL.random_graph(3, 400, 400, 12000, skew=True)
    g = TL.DeviceCSR.from_coo(row, col, 400, 400, chunk=64)
    g = g.with_values(g.gcn_norm()[1])
    best = g.autotune(64, hot=csr.HOT_CANDIDATES)
    assert best in csr.AUTOTUNE_CANDIDATES + (30, 31) and g.variant == best
    assert set(g.autotune_report["ms"]) == {f"v{v}" for v in csr.AUTOTUNE_CANDIDATES} | {f"v{v}h{h}" for v, h in csr.HOT_CANDIDATES} \
        and not g.autotune_report["rejected"]
    # slice size and row order are plan-time dimensions too; the winner's plan is the one left installed
    g.autotune(64, candidates=(0, 16), chunks=(64, 16), degree_orders=(False, True))
    ch = g.autotune_report["chosen"]
    assert len(g.autotune_report["ms"]) == 8 and g.chunk == ch["chunk"] and (g.row_order is not None) == ch["degree_order"]
    X = torch.randn(400, 64, generator=torch.Generator().manual_seed(0))
    TL.close(g.spmm(X), g.spmm(X, variant=0), rtol=1e-5, atol=1e-6)          # the tuned default computes the same operator
    assert g.autotune(128) in (0, 26, 27)                                    # d = 128: default or its 256-bit-gather forms
    assert g.autotune(256) == 0                                              # no alternatives there: default kept
    # a candidate that does not exist for the shape is rejected, not chosen
    assert g.autotune(64, candidates=(0, 99)) in (0, 99)
    TL.close(g.spmm(X), g.with_values(g.val).spmm(X, variant=0), rtol=1e-5, atol=1e-6)
    # module-level drivers
    U, I = 30, 20
    m = lg_module().LightGCN(U, I, 32, 2)
    ei = torch.stack([torch.randint(0, U, (300,)), torch.randint(0, I, (300,))])
    r, c, n = TL.lo.wiring_symmetric(ei[0], ei[1], U, I)
    adj = lg_module().SparseTensor(row=r, col=c, sparse_sizes=(n, n))
    fwd, bwd = m.autotune(adj)
    assert fwd in csr.AUTOTUNE_CANDIDATES + (30, 31) and bwd in csr.AUTOTUNE_CANDIDATES + (30, 31)
    pb = make_problem(seed=1, U=200, I=60, E=3000, d=64, K=2, B=64)
    eng = ShardedLightGCN(pb["U"], pb["I"], pb["d"], 2, pb["users"], pb["items"], cuda_dev, ops=make_emu_ops(), rank=0, world=1)
    assert set(eng.autotune()) == {"users", "items"}


def lg_module():
    import laplace_gnn_recommendation_b200 as lg
    return lg
test_sageconv_project_first_matches_reference_order = TZ.test_sageconv_project_first_matches_reference_order


def test_bench_hetero_workload_pieces(cuda_dev, monkeypatch):
    """bench.py --workload hetero_*: the synthetic batch, the model, one training step on the (emulated) kernels and the
    CPU-oracle arm built from the model's state_dict must describe the same computation (same loss)."""
    import bench
    monkeypatch.setitem(bench.HETERO_SIZES, "hetero_s", (600, 40, 90, 30))
    x, ei, eli, y = bench.hetero_batch("hetero_s", cuda_dev)
    for project_first in (False, True):
        model, metadata = bench.hetero_model("add", project_first)
        loss = torch.nn.BCEWithLogitsLoss()(model(dict(x), ei, eli), y)
        loss.backward()
        cpu = bench.hetero_cpu_step_runner("hetero_s", "add", model.state_dict(), metadata)()
        assert float(loss) == pytest.approx(float(cpu), rel=1e-5)
    sd = bench.materialized_state_dict()
    assert set(sd) == {k for k in model.state_dict() if "num_batches_tracked" not in k and "running_" not in k}
    assert bench.seg_bytes(10, 4, 3) == 10 * (4 + 16) + 4 * 4 + 3 * 16


class _WallEvent:
    def __init__(self):
        self.t = 0.0

    def record(self, stream=None):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


@pytest.mark.parametrize("workload", ["hm", "hetero_s"])
def test_bench_script_logic_dry_run(cuda_dev, monkeypatch, capsys, workload):
    """bench.py's own logic (workload set-up, plan-time tuning, timed loop, roofline / e2e / cpu_baseline assembly, the JSON
    contract) executed end to end at N=1 on a tiny workload: kernels = the emulated library, CUDA events = wall clock.  The
    numbers mean nothing; every key the driver reads must be there and be finite."""
    import argparse
    import json
    import math
    import bench
    from laplace_gnn_recommendation_b200 import csr

    class FakeCuda:
        available = staticmethod(lambda: True)
        event = staticmethod(_WallEvent)
        synchronize = staticmethod(lambda: None)
        empty_cache = staticmethod(lambda: None)
        pin = staticmethod(lambda t: t)
        device = staticmethod(lambda local: cuda_dev)
    monkeypatch.setattr(bench, "CUDA", FakeCuda)
    monkeypatch.setattr(csr, "_time_ms", lambda fn, reps, device: (fn(), 1.0)[1])
    monkeypatch.setattr(torch, "Generator", lambda device=None: torch._C.Generator())     # make_graph asks for a device generator
    monkeypatch.setitem(bench.WORKLOADS, "hm", (300, 120, 4000))
    monkeypatch.setitem(bench.HETERO_SIZES, "hetero_s", (600, 40, 90, 30))
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(gpus=1, steps=2, warmup=3, impl="ours", workload=workload, degree="powerlaw", dim=64, layers=3, batch=64,
                              degree_order=False, no_cpu_baseline=False, no_autotune=False, graph="off", exchange="nccl",
                              schedule="auto", hetero_aggr="add", project_first=False)
    (bench.run_hetero if workload.startswith("hetero") else bench.run_ours)(args)
    line = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["gpu_launches"] > 0 and math.isfinite(line["value"]) and line["value"] > 0
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and line["e2e"]["h2d_bytes_per_step"] > 0
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and line["roofline"]["achieved"] > 0
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] == "port"
    assert "workload" in line["config"] and math.isfinite(line["loss"])
    if workload == "hm":
        tuned = line["config"]["spmm_variant"]
        assert "error" not in tuned and tuned["forward"]["variant"] in csr.AUTOTUNE_CANDIDATES + (30, 31) and not tuned["rejected"], tuned


def _bench_rank(rank, world, port, out_dir):
    import argparse
    import contextlib
    import io
    import os
    import time
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    import bench
    from laplace_gnn_recommendation_b200 import csr
    from tests.test_dist_gloo import make_emu_ops

    class FakeCuda:
        available = staticmethod(lambda: True)
        event = staticmethod(_WallEvent)
        synchronize = staticmethod(lambda: None)
        empty_cache = staticmethod(lambda: None)
        pin = staticmethod(lambda t: t)
        device = staticmethod(lambda local: torch.device("cpu"))
        backend = "gloo"
        sharded_ops = staticmethod(make_emu_ops)
        step_timer = staticmethod(lambda fn: (fn(), 1.0)[1])
    bench.CUDA = FakeCuda
    csr._time_ms = lambda fn, reps, device: (fn(), 1.0)[1]
    torch.Generator = lambda device=None: torch._C.Generator()
    bench.WORKLOADS["hm"] = (300, 120, 4000)
    args = argparse.Namespace(gpus=world, steps=2, warmup=3, impl="ours", workload="hm", degree="powerlaw", dim=64, layers=3, batch=64,
                              degree_order=False, no_cpu_baseline=True, no_autotune=False, graph="off", exchange="nccl",
                              schedule="auto", hetero_aggr="add", project_first=False)
    buf = io.StringIO()
    with emulated(), contextlib.redirect_stdout(buf):
        bench.run_ours(args)
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write(buf.getvalue())


def test_bench_script_logic_dry_run_two_ranks(tmp_path):
    """The multi-GPU branch of bench.py (sharded engine, per-rank plan-time tuning, max-over-ranks timing, rank-0 JSON line) on
    two gloo ranks with the emulated kernels."""
    import json
    import torch.multiprocessing as mp
    from tests.emu import build_emu
    from tests.test_dist_gloo import _free_port
    build_emu.build()
    mp.spawn(_bench_rank, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    out0, out1 = (open(tmp_path / f"rank{r}.txt").read() for r in range(2))
    assert not [ln for ln in out1.splitlines() if ln.startswith("{")]                  # only rank 0 prints
    line = json.loads([ln for ln in out0.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["value"] > 0 and line["cpu_baseline"] is None
    assert "users" in line["config"]["spmm_variant"] and "items" in line["config"]["spmm_variant"], line["config"]
    form = line["config"]["spmm_variant"]["step_form"]
    assert len(form["ms"]) == 2 and not form["rejected"] and form["chosen"]["schedule"] in ("layer", "chains"), form
    par = line["shard_parity"]                                                          # the driver-visible N > 1 parity field
    assert par["eager"]["ok"] and par["eager"]["loss_rel_err"] <= 1e-5 and par["eager"]["grad_max_rel_err"] <= 1e-5, par
    assert line["e2e"]["value"] > 0 and line["roofline"]["launches_timed"] > 0 and line["gpu_launches"] > 0
test_acceptance_lightgcn_learns = TZ.test_acceptance_lightgcn_learns
test_device_sampler_distribution_and_validity = TZ.test_device_sampler_distribution_and_validity
test_acceptance_ranking_model_learns = TZ.test_acceptance_ranking_model_learns
test_reverse_edge_type_shares_the_transposed_csr = TZ.test_reverse_edge_type_shares_the_transposed_csr
test_topk_tiled_scoring_is_bit_identical = TZ.test_topk_tiled_scoring_is_bit_identical
test_bpr_misaligned_tables_take_the_scalar_kernel = TZ.test_bpr_misaligned_tables_take_the_scalar_kernel
