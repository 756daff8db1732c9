"""Direct synthetic-index generator for benchmark-sized indexes.

A 1M-document x 300-token index has 3e8 tokens (76.8 GB of raw fp16 embeddings): running
k-means and the encoder over it just to obtain a benchmark input is not what the metric
measures.  This generator writes the index *layout* directly, chunk by chunk and seeded per
chunk so that any document range (a shard) can be produced independently and the union over
ranks is the same index:

  centroids   L2-normalised randn(K, dim), K = 2^floor(log2(16*sqrt(E)))   (fast_plaid.py:152-154)
  codes       uniform over [0, K)  -- the statistical worst case for the candidate stage: every
              document touches ~len distinct cells, so probing 8 cells x 32 tokens reaches about
              a quarter of a 1M-document index (SURVEY.md 8d "uniform-random model")
  residuals   uniform random bytes = every bucket equally likely, which is what quantile
              cutoffs produce on the data they were trained on (create.rs:352-357)
  weights     the (i+0.5)/2^nbits quantiles of N(0, sigma^2)                 (create.rs:359-364)
  ivf         per centroid the sorted unique ids of the documents using it   (create.rs:528-559)

The decompressed tokens are centroid + per-dimension quantised Gaussian noise, renormalised,
i.e. a valid PLAID index of a clustered corpus.

`topics > 0` switches to the *clustered* variant SURVEY.md 8(d) asks to report next to the
worst case: the K centroids are grouped into `topics` blocks of similar directions (block
direction + 0.5 * noise, renormalised), every document belongs to one topic and draws a
fraction `1 - mix` of its codes from its topic's block and `mix` uniformly.  A query token's
nearest centroids then sit in one block and the candidate set shrinks from ~25 % of the index
to a few per cent, which moves the time from the approximate stage to the fixed-cost stages.
"""

from __future__ import annotations

import math
import os
import types

import torch

if __package__:
    from .layout import build_ivf, num_partitions_for
else:  # loaded by file path (bench.py --impl reference): keep the package and its shared libraries out
    import importlib.util

    _spec = importlib.util.spec_from_file_location(
        "_fpb_layout", os.path.join(os.path.dirname(os.path.abspath(__file__)), "layout.py"))
    _layout = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_layout)
    build_ivf, num_partitions_for = _layout.build_ivf, _layout.num_partitions_for


def _normal_quantiles(n: int, sigma: float) -> torch.Tensor:
    p = (torch.arange(n, dtype=torch.float64) + 0.5) / n
    return (sigma * math.sqrt(2.0) * torch.erfinv(2.0 * p - 1.0)).float()


@torch.inference_mode()
def synthetic_arrays(n_docs: int, doc_len: int, dim: int = 128, nbits: int = 4, device: str = "cuda:0",
                     seed: int = 1234, doc_range: tuple[int, int] | None = None, ragged: bool = False,
                     sigma: float = 0.05, docs_per_chunk: int = 25_000, topics: int = 0,
                     mix: float = 0.05) -> tuple[types.SimpleNamespace, int]:
    """Returns (arrays on `device` with the field names of engine.IndexTensors, doc_id_base), using torch only.
    `doc_range=(lo, hi)` generates only that slice of the global index (document sharding); ids in the IVF are
    then local."""
    dev = torch.device(device)
    lo, hi = (0, n_docs) if doc_range is None else doc_range
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    # document lengths of the WHOLE index (cheap) so that K and the chunk seeds are global
    if ragged:
        gl = torch.Generator().manual_seed(seed + 1)
        lengths = torch.randint(max(1, doc_len // 4), doc_len + 1, (n_docs,), generator=gl)
    else:
        lengths = torch.full((n_docs,), doc_len, dtype=torch.int64)
    total_tokens = int(lengths.sum())
    K = num_partitions_for(float(total_tokens))
    centroids = torch.randn(K, dim, generator=g, device=dev)
    block = 0
    if topics > 0:
        topics = min(topics, K)
        block = K // topics  # K is a power of two; topics is clipped to a divisor below
        while K % topics:
            topics -= 1
            block = K // topics
        directions = torch.nn.functional.normalize(torch.randn(topics, dim, generator=g, device=dev), dim=-1)
        centroids = directions.repeat_interleave(block, dim=0) + 0.5 * centroids / math.sqrt(dim)
    centroids = torch.nn.functional.normalize(centroids, dim=-1).half()
    weights = _normal_quantiles(2**nbits, sigma).to(dev).half()
    pd = dim * nbits // 8
    my_lengths = lengths[lo:hi]
    n_tok = int(my_lengths.sum())
    codes = torch.empty(max(n_tok, 1), dtype=torch.int32, device=dev)
    residuals = torch.empty((max(n_tok, 1), pd), dtype=torch.uint8, device=dev)
    offs = torch.zeros(n_docs + 1, dtype=torch.int64)
    offs[1:] = lengths.cumsum(0)
    t_base = int(offs[lo])
    for c0 in range((lo // docs_per_chunk) * docs_per_chunk, hi, docs_per_chunk):
        c1 = min(c0 + docs_per_chunk, n_docs)
        gc = torch.Generator(device=dev)
        gc.manual_seed(seed * 1_000_003 + c0)
        n = int(offs[c1] - offs[c0])
        cc = torch.randint(0, K, (n,), generator=gc, device=dev, dtype=torch.int32)
        if topics > 0:
            # topic of a document: a fixed hash of its global id; tokens keep the uniform draw with
            # probability `mix`, otherwise they land in the topic's block
            doc_of_tok = torch.repeat_interleave(torch.arange(c0, c1, device=dev), lengths[c0:c1].to(dev))
            topic = (doc_of_tok * 2654435761) % topics
            in_block = topic * block + (cc.long() % block)
            keep = torch.rand(n, generator=gc, device=dev) < mix
            cc = torch.where(keep, cc.long(), in_block).to(torch.int32)
        rr = torch.randint(0, 256, (n, pd), generator=gc, device=dev, dtype=torch.uint8)
        # clip the chunk to [lo, hi)
        a = max(c0, lo)
        b = min(c1, hi)
        s0 = int(offs[a] - offs[c0])
        s1 = int(offs[b] - offs[c0])
        d0 = int(offs[a]) - t_base
        codes[d0 : d0 + (s1 - s0)] = cc[s0:s1]
        residuals[d0 : d0 + (s1 - s0)] = rr[s0:s1]
        del cc, rr
    ivf, ivf_lengths = build_ivf(codes[:n_tok], my_lengths, K)
    data = types.SimpleNamespace(
        nbits=nbits,
        centroids=centroids,
        bucket_weights=weights,
        doc_lengths=my_lengths,
        doc_codes=codes[:n_tok] if n_tok > 0 else codes[:0],
        doc_residuals=residuals[:n_tok] if n_tok > 0 else residuals[:0],
        ivf=ivf.to(torch.int32),
        ivf_lengths=ivf_lengths,
    )
    return data, lo


def synthetic_index(*args, **kwargs):
    """`synthetic_arrays` as engine.IndexTensors (what DeviceIndex takes)."""
    from ..engine import IndexTensors

    a, base = synthetic_arrays(*args, **kwargs)
    return IndexTensors(nbits=a.nbits, centroids=a.centroids, bucket_weights=a.bucket_weights,
                        doc_lengths=a.doc_lengths, doc_codes=a.doc_codes, doc_residuals=a.doc_residuals,
                        ivf=a.ivf, ivf_lengths=a.ivf_lengths), base


class SyntheticDocuments:
    """A lazy, seeded corpus: `len()` documents of `doc_len` L2-normalised random tokens,
    generated on `device` in blocks of `block` documents when they are asked for, so that FastPlaid.create() can
    build an index whose raw embeddings (77 GB at 1M x 300 x 128 fp16) never exist at once.  It behaves like the
    `list[torch.Tensor]` the reference takes: `docs[i]`, `docs[a:b]`, iteration, `len(docs)`; `doc_lengths` lets
    the builder size things without touching the data."""

    def __init__(self, n_docs: int, doc_len: int, dim: int = 128, seed: int = 7, device: str = "cuda:0",
                 block: int = 2048, ragged: bool = False, clusters: int = 0, spread: float = 0.03) -> None:
        """`clusters` > 0 draws every token around one of that many random unit directions (+ `spread` * randn per
        dimension, renormalised) -- embeddings of a real encoder are clustered, which is what k-means centroids and
        the IVF probe rely on; 0 gives unstructured unit vectors."""
        self.n_docs, self.doc_len, self.dim, self.seed = int(n_docs), int(doc_len), int(dim), int(seed)
        self.device = torch.device(device)
        self.block = int(block)
        self.spread = float(spread)
        self.centers = None
        if clusters > 0:
            gc = torch.Generator(device=self.device)
            gc.manual_seed(seed * 7 + 3)
            self.centers = torch.nn.functional.normalize(
                torch.randn(int(clusters), dim, generator=gc, device=self.device), dim=-1)
        if ragged:
            g = torch.Generator().manual_seed(seed + 1)
            self.doc_lengths = torch.randint(max(1, doc_len // 4), doc_len + 1, (n_docs,), generator=g)
        else:
            self.doc_lengths = torch.full((n_docs,), doc_len, dtype=torch.int64)
        self._cache: tuple[int, torch.Tensor] | None = None

    def __len__(self) -> int:
        return self.n_docs

    def _block(self, bi: int) -> torch.Tensor:
        if self._cache is not None and self._cache[0] == bi:
            return self._cache[1]
        g = torch.Generator(device=self.device)
        g.manual_seed(self.seed * 1_000_003 + bi)
        x = torch.randn(self.block, self.doc_len, self.dim, generator=g, device=self.device)
        if self.centers is not None:
            which = torch.randint(0, self.centers.shape[0], (self.block, self.doc_len), generator=g, device=self.device)
            x = self.centers[which] + self.spread * x
        x = torch.nn.functional.normalize(x, dim=-1).half()
        self._cache = (bi, x)
        return x

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self.n_docs))]
        if i < 0:
            i += self.n_docs
        if not 0 <= i < self.n_docs:
            raise IndexError(i)
        return self._block(i // self.block)[i % self.block, : int(self.doc_lengths[i])]

    def __iter__(self):
        for i in range(self.n_docs):
            yield self[i]
