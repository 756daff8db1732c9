"""Pure-torch layout helpers shared by the index builder and the synthetic generator.

This file imports nothing from the package on purpose: `bench.py --impl reference` loads it (and
synthetic.py) by file path so that the reference arm's process never imports the engine or any of its
shared libraries.
"""

from __future__ import annotations

import math

import torch


def num_partitions_for(n_embeddings: float) -> int:
    """K heuristic (fast_plaid.py:152-154, create.rs:292-294)."""
    return int(2 ** math.floor(math.log2(16 * math.sqrt(max(n_embeddings, 1.0)))))


def build_ivf(codes: torch.Tensor, doc_lengths: torch.Tensor, n_cells: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Per centroid, the sorted unique ids of the documents owning a token with that code
    (create.rs:528-559, optimize_ivf :55-132) -- computed with one sort of (code, doc) keys."""
    n_docs = doc_lengths.shape[0]
    tok2doc = torch.repeat_interleave(torch.arange(n_docs, dtype=torch.int64, device=codes.device),
                                      doc_lengths.to(codes.device))
    key = codes.to(torch.int64) * max(n_docs, 1) + tok2doc
    uniq = torch.unique(key, sorted=True)
    cell = torch.div(uniq, max(n_docs, 1), rounding_mode="floor")
    ivf = uniq - cell * max(n_docs, 1)
    n_cells = max(n_cells, int(cell.max()) + 1 if cell.numel() else 0)
    lengths = torch.bincount(cell, minlength=n_cells)
    return ivf, lengths.to(torch.int32)
