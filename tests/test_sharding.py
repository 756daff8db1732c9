"""Document sharding: shard_tensors + "per-shard top-R records, all-gather,
global prune, rank" must reproduce the single-index result exactly.  Host-side logic only:
the per-shard pipeline is played by the oracle, the exchange by gloo (world_size 2)."""

from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import (build_oracle_index, make_docs, make_queries, merge_records_host, shard_records_oracle,
                  to_index_tensors)

from fast_plaid_b200.engine import shard_tensors
from oracle import plaid_oracle as po

N_PROBE, N_FULL, TOP_K = 4, 64, 10


def _fixture():
    docs = make_docs(240, 5, 40, seed=55)
    oidx, _ = build_oracle_index(docs)
    queries = make_queries(3, 16, seed=56, docs=docs)
    return oidx, queries


def _shard_oracle(oidx, rank, world):
    sh, base = shard_tensors(to_index_tensors(oidx), rank, world)
    return po.OracleIndex(sh.nbits, sh.centroids, sh.bucket_weights, sh.ivf.long(), sh.ivf_lengths.long(),
                          sh.doc_codes.long(), sh.doc_residuals, sh.doc_lengths.long()), base


@pytest.mark.parametrize("world", [2, 3, 5])
def test_sharded_merge_equals_single_index(world):
    oidx, queries = _fixture()
    R = N_FULL // 4
    for b in range(queries.shape[0]):
        ref_ids, ref_sc = po.search_one(queries[b], oidx, N_PROBE, 2000, N_FULL, TOP_K, ties="canonical")
        recs = []
        for r in range(world):
            sh, base = _shard_oracle(oidx, r, world)
            recs += shard_records_oracle(sh, base, queries[b], N_PROBE, N_FULL)
        got = merge_records_host(recs, R, TOP_K)
        assert [d for d, _ in got] == ref_ids
        assert [s for _, s in got] == ref_sc


def test_shard_tensors_partitions_everything():
    oidx, _ = _fixture()
    t = to_index_tensors(oidx)
    world = 4
    n_docs = n_tok = n_ivf = 0
    for r in range(world):
        sh, base = shard_tensors(t, r, world)
        assert base == n_docs
        n_docs += sh.num_documents
        n_tok += sh.doc_codes.shape[0]
        n_ivf += int(sh.ivf_lengths.sum())
        assert int(sh.doc_lengths.sum()) == sh.doc_codes.shape[0] == sh.doc_residuals.shape[0]
        assert int(sh.ivf.max()) < sh.num_documents and int(sh.ivf.min()) >= 0
    assert (n_docs, n_tok, n_ivf) == (t.num_documents, t.doc_codes.shape[0], int(t.ivf.shape[0]))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    oidx, queries = _fixture()
    sh, base = _shard_oracle(oidx, rank, world)
    R = N_FULL // 4
    B = queries.shape[0]
    # fixed-size records exactly like fpb_record: (approx, exact, doc id), -inf / -1 padding
    rec = torch.full((B, R, 3), float("-inf"), dtype=torch.float64)
    rec[:, :, 2] = -1
    for b in range(B):
        for i, (a, e, d) in enumerate(shard_records_oracle(sh, base, queries[b], N_PROBE, N_FULL)):
            rec[b, i] = torch.tensor([a, e, d], dtype=torch.float64)
    gathered = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(gathered, rec)
    results = []
    for b in range(B):
        recs = [(float(a), float(e), int(d)) for g in gathered for a, e, d in g[b].tolist() if d >= 0]
        results.append(merge_records_host(recs, R, TOP_K))
    if rank == 0:
        out.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_all_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    oidx, queries = _fixture()
    for b in range(queries.shape[0]):
        ref_ids, ref_sc = po.search_one(queries[b], oidx, N_PROBE, 2000, N_FULL, TOP_K, ties="canonical")
        assert [d for d, _ in results[b]] == ref_ids
        assert [float(s) for _, s in results[b]] == pytest.approx(ref_sc, abs=0)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_subset_search_equals_single_index(world):
    """subset= with sharded documents (search.rs:494-517 restricts the probe to the centroids the subset
    documents touch): each shard derives that set from its own subset documents, the sets are united (the
    engine all-gathers and ORs bitmaps), every shard probes with the union and masks its candidates with its
    local subset; global prune + rank as before.  Must equal the unsharded subset search exactly."""
    oidx, queries = _fixture()
    g = torch.Generator().manual_seed(9)
    n_docs = int(oidx.doc_lengths.shape[0])
    subsets = [torch.randperm(n_docs, generator=g)[:80], torch.arange(0, 50), torch.arange(n_docs - 30, n_docs)]
    R = N_FULL // 4
    shards = [_shard_oracle(oidx, r, world) for r in range(world)]
    for b in range(queries.shape[0]):
        sub = subsets[b]
        ref_ids, ref_sc = po.search_one(queries[b], oidx, N_PROBE, 2000, N_FULL, TOP_K, subset=sub, ties="canonical")
        assert len(ref_ids) > 0
        local = []
        for sh, base in shards:
            n_local = int(sh.doc_lengths.shape[0])
            m = (sub >= base) & (sub < base + n_local)
            local.append(sub[m] - base)
        # step 1a on every shard: the centroids its own subset documents touch
        sets = []
        for (sh, _), ls in zip(shards, local):
            st = po.search_one(queries[b], sh, N_PROBE, 2000, N_FULL, 10**9, subset=ls, ties="canonical", return_stages=True)
            sets.append(st["subset_centroids"])
        union = torch.unique(torch.cat(sets), sorted=True)
        # step 1b..: probe with the union, candidates masked by the local subset, records, global merge
        recs = []
        for (sh, base), ls in zip(shards, local):
            st = po.search_one(queries[b], sh, N_PROBE, 2000, N_FULL, 10**9, subset=ls, ties="canonical",
                               return_stages=True, inject={"subset_centroids": union})
            if "rerank" not in st:
                continue
            approx_of = dict(zip(st["candidates"].tolist(), st["approx"].tolist()))
            recs += [(approx_of[d], float(e), d + base) for d, e in zip(st["rerank"].tolist(), st["exact"].tolist())]
        got = merge_records_host(recs, R, TOP_K)
        assert [d for d, _ in got] == ref_ids
        assert [s for _, s in got] == ref_sc
        assert set(ref_ids) <= set(sub.tolist())


def test_query_group_by_document_shard_grid_covers_every_query_and_document_once():
    """csrc/comm.cu: rank r = document shard r % n_shards of query group r / n_shards; a batch of B queries is cut
    into n_groups contiguous slices of ceil(B / n_groups) (the last may be short or empty), the documents into
    n_shards contiguous ranges.  Every (query, document) pair must belong to exactly one rank."""
    from fast_plaid_b200.engine import shard_grid

    for world in (1, 2, 4, 8):
        for n_groups in [g for g in (1, 2, 4, 8) if world % g == 0]:
            n_shards = world // n_groups
            for B, n_docs in ((64, 1000), (5, 17), (1, 3), (16, 8)):
                b_local = -(-B // n_groups)
                seen = {}
                for r in range(world):
                    g, s, ns = shard_grid(r, world, n_groups)
                    assert ns == n_shards and 0 <= g < n_groups and 0 <= s < n_shards
                    q0 = g * b_local
                    qs = range(q0, min(B, q0 + b_local))
                    ds = range((n_docs * s) // n_shards, (n_docs * (s + 1)) // n_shards)
                    for q in qs:
                        for d in ds:
                            assert (q, d) not in seen, (world, n_groups, q, d)
                            seen[(q, d)] = r
                assert len(seen) == B * n_docs
    import pytest

    with pytest.raises(ValueError):
        shard_grid(0, 6, 4)
