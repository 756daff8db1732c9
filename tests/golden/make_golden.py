"""Mint the golden fixtures under tests/golden/ from the CPU oracle.

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference ships no golden vectors and cannot be built in this image
(Rust), so these vectors come from oracle/plaid_oracle.py -- the op-for-op PyTorch-CPU
restatement of rust/search/search.rs run on torch 2.11.0 (the version the reference's CI
pins).  They freeze the oracle's behaviour so that (a) a torch upgrade that changes an ATen
CPU kernel is noticed, and (b) the GPU tests compare against committed numbers, not only
against a live oracle.
"""

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from util import build_oracle_index, make_docs, make_queries  # noqa: E402

from oracle import plaid_oracle as po  # noqa: E402


def make(name, n_docs, lo, hi, dim, nbits, B, Q, top_k, n_full, n_probe, noisy):
    docs = make_docs(n_docs, lo, hi, dim=dim, seed=2024)
    oidx, extra = build_oracle_index(docs, nbits=nbits, seed=42)
    queries = make_queries(B, Q, dim=dim, seed=99, docs=docs if noisy else None)
    per_query = []
    for b in range(B):
        st = po.search_one(queries[b], oidx, n_probe, 2000, n_full, top_k, ties="canonical", return_stages=True)
        per_query.append({
            "cells": st["cells"], "candidates": st["candidates"], "approx": st["approx"],
            "rerank": st["rerank"], "exact": st["exact"], "ids": st["ids"], "scores": st["scores"],
            "S_checksum": float(st["S"].float().sum()),
        })
    blob = {
        "meta": dict(name=name, torch=torch.__version__, n_docs=n_docs, dim=dim, nbits=nbits, B=B, Q=Q,
                     top_k=top_k, n_full=n_full, n_probe=n_probe),
        "index": dict(nbits=nbits, centroids=oidx.centroids, bucket_weights=oidx.bucket_weights, ivf=oidx.ivf.to(torch.int32),
                      ivf_lengths=oidx.ivf_lengths.to(torch.int32), doc_codes=oidx.doc_codes.to(torch.int32),
                      doc_residuals=oidx.doc_residuals, doc_lengths=oidx.doc_lengths.to(torch.int32)),
        "queries": queries.half(),
        "expected": per_query,
    }
    path = os.path.join(HERE, f"{name}.pt")
    torch.save(blob, path)
    print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    make("small_d128_n4", 150, 10, 60, 128, 4, 4, 32, 10, 64, 8, True)
    make("small_d64_n2", 120, 5, 40, 64, 2, 3, 20, 5, 32, 4, True)
