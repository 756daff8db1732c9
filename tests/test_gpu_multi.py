"""Two-GPU test of the document-sharded path over NCCL (skipped with fewer than 2 GPUs):
per-shard fpb_search_shard -> all_gather_into_tensor -> fpb_merge_shards on every rank must
equal the unsharded fpb_search_batch bit for bit."""

from __future__ import annotations

import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, out):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from util import build_oracle_index, make_docs, make_queries, to_index_tensors

    from fast_plaid_b200.engine import DeviceIndex, shard_tensors

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    torch.set_num_threads(4)
    dev = f"cuda:{rank}"
    docs = make_docs(700, 10, 60, seed=31)
    oidx, _ = build_oracle_index(docs)
    t = to_index_tensors(oidx)
    queries = make_queries(5, 32, seed=32, docs=docs).half().to(dev)
    whole = DeviceIndex(t, dev)
    sh, base = shard_tensors(t, rank, world)
    mine = DeviceIndex(sh, dev, doc_id_base=base)
    ok = True
    for n_full, top_k in ((64, 10), (4096, 40)):
        params = DeviceIndex.make_params(top_k, n_full, 8)
        ids, scores, counts = whole.search(queries, params)
        rec = mine.search_records(queries, params)
        gathered = torch.empty((world,) + tuple(rec.shape), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered.view(-1), rec.view(-1))
        i2, s2, c2 = mine.merge_records(gathered, top_k)
        torch.cuda.synchronize()
        ok = ok and torch.equal(i2, ids) and torch.equal(s2, scores) and torch.equal(c2, counts)
        # two-step variant
        keys = mine.shard_approx_keys(queries, params)
        all_keys = torch.empty((world,) + tuple(keys.shape), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_keys.view(-1), keys.view(-1))
        rec2 = mine.shard_exact_records(all_keys, rank, int(queries.shape[1]), params)
        dist.all_gather_into_tensor(gathered.view(-1), rec2.view(-1))
        i3, s3, c3 = mine.merge_records(gathered, top_k)
        torch.cuda.synchronize()
        ok = ok and torch.equal(i3, ids) and torch.equal(s3, scores) and torch.equal(c3, counts)
        # subset search: the shards' centroid bitmaps are all-gathered and OR-ed (search.rs:494-517)
        g = torch.Generator().manual_seed(5)
        subset = [torch.randperm(700, generator=g)[:200].tolist() for _ in range(queries.shape[0])]
        subset[1] = list(range(0, 100))  # lives in shard 0 only
        i4, s4, c4 = whole.search(queries, params, subset=subset)
        ps = DeviceIndex.with_subset_flag(params)
        cb = mine.shard_subset_begin(queries, ps, subset)
        all_cb = torch.empty((world,) + tuple(cb.shape), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_cb.view(-1), cb.view(-1))
        keys = mine.shard_subset_keys(all_cb, int(queries.shape[1]), ps)
        dist.all_gather_into_tensor(all_keys.view(-1), keys.view(-1))
        rec3 = mine.shard_exact_records(all_keys, rank, int(queries.shape[1]), ps)
        dist.all_gather_into_tensor(gathered.view(-1), rec3.view(-1))
        i5, s5, c5 = mine.merge_records(gathered, top_k)
        torch.cuda.synchronize()
        ok = ok and torch.equal(c5, c4)
        for b in range(queries.shape[0]):
            n = int(c4[b])
            ok = ok and torch.equal(i5[b, :n], i4[b, :n]) and torch.equal(s5[b, :n], s4[b, :n])
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_sharded_search_equals_unsharded():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == 1
