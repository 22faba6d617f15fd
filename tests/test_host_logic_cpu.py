"""CPU: data-parallel host logic -- shard arithmetic and the gradient all-reduce over gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pdc_b200 import data_parallel as DP


def test_shard_range_covers_everything():
    for total in (1, 7, 8, 32, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [DP.shard_range(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            cur = 0
            for s, c in spans:
                assert s == cur
                cur += c
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_flat_view_detection():
    flat = torch.arange(40, dtype=torch.float32)
    views = [flat[0:12].view(3, 4), flat[12:15], flat[16:40].view(2, 12)]     # one 1-element (4-byte) gap: allowed
    fv = DP._flat_view_of(views)
    assert fv is not None and fv.numel() == 40 and fv.data_ptr() == flat.data_ptr()
    assert DP._flat_view_of([flat[0:12], torch.zeros(3)]) is None
    assert DP._flat_view_of([flat[12:15], flat[0:12]]) is None


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, flat_mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = DP.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    shapes = [(4, 3, 3, 3), (4,), (4,), (5, 4, 1, 1), (5,)]
    if flat_mode:
        flat = torch.zeros(sum((torch.Size(s).numel() + 3) // 4 * 4 for s in shapes))
        params, off = [], 0
        for s in shapes:
            n = torch.Size(s).numel()
            params.append(torch.nn.Parameter(flat[off:off + n].view(s)))
            off += (n + 3) // 4 * 4
        gflat = torch.zeros_like(flat)
        off = 0
        for p, s in zip(params, shapes):
            n = torch.Size(s).numel()
            p.grad = gflat[off:off + n].view(s)
            off += (n + 3) // 4 * 4
    else:
        params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
        for p in params:
            p.grad = torch.zeros_like(p)
    for i, p in enumerate(params):
        p.grad.copy_(torch.full(p.shape, float((rank + 1) * (i + 1))))
    red = DP.GradientAllReducer(params, num_buckets=3)
    red()
    expect = [(1 + 2) / 2.0 * (i + 1) for i in range(len(params))]
    ok = all(torch.allclose(p.grad, torch.full(p.shape, e)) for p, e in zip(params, expect))
    ok = ok and red.used_flat_path == flat_mode
    # broadcast_parameters: rank 1 adopts rank 0's weights
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(float(rank + 5))
    DP.broadcast_parameters(lin, src=0)
    ok = ok and bool((lin.weight == 5.0).all())
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def _overlap_worker(rank, world, port, q):
    """Host logic of the overlapped reducer: buckets all-reduced as the (here simulated) backward reports them."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from pdc_b200 import data_parallel as DP
    DP.init_from_env(backend="gloo")

    class FakeBackbone(object):
        _bucket_hook = None
    mod = FakeBackbone()
    params = [torch.nn.Parameter(torch.zeros(4))]
    red = DP.GradientAllReducer(params, module=mod, overlap=True)
    ok = mod._bucket_hook is red and abs(red.cotangent_scale() - 0.5) < 1e-12
    flat = torch.arange(40, dtype=torch.float32) * (rank + 1) * red.cotangent_scale()     # the backward pre-scales by 1/world
    for b, (off, n) in enumerate([(24, 16), (8, 16), (4, 4), (0, 4)]):                   # completion order: last layers first
        red.__call_bucket__(flat, b, off, n)
    red.finish(flat)
    ok = ok and torch.allclose(flat, torch.arange(40, dtype=torch.float32) * 1.5) and red.overlapped_steps == 1
    ok = ok and red.bytes_last == 40 * 4
    red()                                              # explicit call afterwards: nothing left to do, must not reduce twice
    ok = ok and torch.allclose(flat, torch.arange(40, dtype=torch.float32) * 1.5)
    red.detach()
    ok = ok and mod._bucket_hook is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_allreduce_host_logic_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("flat_mode", [True, False])
def test_gradient_allreduce_gloo_world2(flat_mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, flat_mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the arm the driver times next to ours): exactly one JSON line on stdout with the contract's
    keys, produced by the CPU oracle port alone (no CUDA needed); ranks other than 0 print nothing and exit 0."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    small = ["--height", "96", "--width", "128", "--pairs-per-gpu", "2", "--matches", "50", "--non-matches", "100"]   # CPU-suite sized
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"] + small,
                       capture_output=True, text=True, timeout=900, cwd=root, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0 and d["higher_is_better"] is True
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["libddn_b200_mapped"] is False            # the reference arm times the oracle alone: the product library is not even loaded
    env["RANK"] = "1"; env["WORLD_SIZE"] = "2"
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                        capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
