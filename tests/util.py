"""Shared helpers for the parity tests (seeded synthetic indexes, tolerant comparisons)."""

from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import index_oracle, plaid_oracle  # noqa: E402


def make_docs(n_docs: int, min_len: int, max_len: int, dim: int = 128, seed: int = 1234,
              normalize: bool = True) -> list[torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_len, max_len + 1, (n_docs,), generator=g)
    docs = []
    for ln in lens.tolist():
        x = torch.randn(ln, dim, generator=g)
        docs.append(torch.nn.functional.normalize(x, dim=-1) if normalize else x)
    return docs


def make_queries(B: int, Q: int, dim: int = 128, seed: int = 4321, docs: list[torch.Tensor] | None = None,
                 noise: float = 0.2) -> torch.Tensor:
    """Random unit queries, or (with `docs`) noisy copies of document tokens -- the second
    kind gives well separated scores like real retrieval."""
    g = torch.Generator().manual_seed(seed)
    if docs is None:
        return torch.nn.functional.normalize(torch.randn(B, Q, dim, generator=g), dim=-1)
    out = []
    for _ in range(B):
        d = docs[int(torch.randint(0, len(docs), (1,), generator=g))]
        rows = torch.randint(0, d.shape[0], (Q,), generator=g)
        x = d[rows].float() + noise * torch.randn(Q, dim, generator=g)
        out.append(torch.nn.functional.normalize(x, dim=-1))
    return torch.stack(out)


def build_oracle_index(docs, nbits: int = 4, seed: int = 42, kmeans_niters: int = 4):
    cent = index_oracle.compute_centroids(docs, kmeans_niters, seed)
    idx, extra = index_oracle.build_index(docs, cent, nbits=nbits, seed=seed)
    return idx, extra


def to_index_tensors(idx: plaid_oracle.OracleIndex):
    from fast_plaid_b200.engine import IndexTensors

    return IndexTensors(
        nbits=idx.nbits,
        centroids=idx.centroids,
        bucket_weights=idx.bucket_weights,
        doc_lengths=idx.doc_lengths,
        doc_codes=idx.doc_codes,
        doc_residuals=idx.doc_residuals,
        ivf=idx.ivf,
        ivf_lengths=idx.ivf_lengths,
    )


def fp16_ulp_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Distance in representable fp16 steps between two fp16 tensors."""
    def key(x):
        i = x.contiguous().view(torch.int16).to(torch.int32)
        return torch.where(i < 0, -(i & 0x7FFF), i)
    return (key(a.half()) - key(b.half())).abs()


def oracle_exact_scores(oidx, query: torch.Tensor, doc_ids: list[int]) -> torch.Tensor:
    """Reference exact MaxSim score (search.rs:626-656) of arbitrary documents."""
    sel = torch.tensor(doc_ids, dtype=torch.int64)
    codes, lens = plaid_oracle.ragged_lookup(oidx.doc_codes, oidx.doc_offsets, oidx.doc_lengths, sel)
    res, _ = plaid_oracle.ragged_lookup(oidx.doc_residuals, oidx.doc_offsets, oidx.doc_lengths, sel)
    emb = plaid_oracle.decompress_residuals(res, oidx.bucket_weights, oidx.byte_reversed_bits_map,
                                            oidx.bucket_weight_indices_lookup, codes, oidx.centroids, oidx.dim,
                                            oidx.nbits)
    padded, mask = plaid_oracle.direct_pad_sequences(emb, lens, 0.0)
    ts = padded.matmul(query.half().unsqueeze(0).transpose(-2, -1))
    return plaid_oracle.colbert_score_reduce(ts, mask)


def oracle_token_matrix(oidx, query: torch.Tensor, doc_id: int) -> torch.Tensor:
    """Reference token-score matrix (search.rs:651-655 then :668-686) of one document: fp16 [Q, len]."""
    sel = torch.tensor([doc_id], dtype=torch.int64)
    codes, lens = plaid_oracle.ragged_lookup(oidx.doc_codes, oidx.doc_offsets, oidx.doc_lengths, sel)
    res, _ = plaid_oracle.ragged_lookup(oidx.doc_residuals, oidx.doc_offsets, oidx.doc_lengths, sel)
    emb = plaid_oracle.decompress_residuals(res, oidx.bucket_weights, oidx.byte_reversed_bits_map,
                                            oidx.bucket_weight_indices_lookup, codes, oidx.centroids, oidx.dim,
                                            oidx.nbits)
    padded, _ = plaid_oracle.direct_pad_sequences(emb, lens, 0.0)
    ts = padded.matmul(query.half().unsqueeze(0).transpose(-2, -1))  # [1, len, Q]
    return ts[0, : int(lens[0])].transpose(0, 1).contiguous()


def ranking_consistent(gpu_ids, gpu_scores, ref_score_of: dict, tol: float, fallback=None) -> tuple[bool, str]:
    """The GPU ranking must be a valid ranking of the reference scores up to `tol`:
    every returned doc has (nearly) the reference score, and no doc is ranked above another
    whose reference score is larger by more than tol.  A doc the reference run did not
    re-rank (possible only through a tie at a topk boundary) is scored through `fallback`."""
    prev = None
    for i, (d, s) in enumerate(zip(gpu_ids, gpu_scores)):
        if d not in ref_score_of:
            if fallback is None:
                return False, f"rank {i}: doc {d} was not scored by the reference"
            ref_score_of[d] = fallback(d)
        r = ref_score_of[d]
        if abs(r - s) > tol * max(1.0, abs(r)):
            return False, f"rank {i}: doc {d} score {s} vs reference {r}"
        if prev is not None and r > prev + tol * max(1.0, abs(r)):
            return False, f"rank {i}: doc {d} (ref {r}) ranked below a doc with ref {prev}"
        prev = r
    return True, ""


def merge_records_host(records: list[tuple[float, float, int]], R: int, top_k: int) -> list[tuple[int, float]]:
    """Host statement of the sharded merge rule (k6_merge_kernel): keep the R best by
    (approx desc, doc id asc), order them by (exact desc, doc id asc), emit top_k."""
    keep = sorted(records, key=lambda r: (-r[0], r[2]))[:R]
    keep.sort(key=lambda r: (-r[1], r[2]))
    return [(r[2], r[1]) for r in keep[:top_k]]


def shard_records_oracle(shard_index, base: int, query: torch.Tensor, n_probe: int, n_full: int):
    """What fpb_search_shard emits for one query, computed with the oracle on a shard."""
    st = plaid_oracle.search_one(query, shard_index, n_probe, 2000, n_full, 10**9, ties="canonical",
                                 return_stages=True)
    if len(st.get("ids", [])) == 0 and "rerank" not in st:
        return []
    approx_of = dict(zip(st["candidates"].tolist(), st["approx"].tolist()))
    return [(approx_of[d], float(e), d + base) for d, e in zip(st["rerank"].tolist(), st["exact"].tolist())]
