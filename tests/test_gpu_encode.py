"""GPU index-build encode kernels (fpb_encode) against the oracle's create.rs restatement."""

from __future__ import annotations

import pytest
import torch

from util import make_docs

from oracle import index_oracle as io

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbits", [4, 2])
def test_encode_matches_oracle(nbits, cuda_device):
    from fast_plaid_b200.engine import encode_tokens

    g = torch.Generator().manual_seed(17)
    K, n = 1000, 5000  # K not a multiple of 128: the last centroid tile is partial
    cent = torch.nn.functional.normalize(torch.randn(K, 128, generator=g), dim=-1).half()
    x = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=-1).half()
    ref_codes = io.compress_into_codes(x, cent)
    res = (x - cent.index_select(0, ref_codes)).float().flatten()
    n_opt = 2 ** nbits
    cutoffs = torch.cat([io.scalar_quantile_kthvalue(res, i / n_opt) for i in range(1, n_opt)])
    codes, packed = encode_tokens(x.to(cuda_device), cent, cutoffs, nbits)
    torch.cuda.synchronize()
    codes, packed = codes.cpu().long(), packed.cpu()
    same = codes == ref_codes
    assert float(same.float().mean()) > 0.999, f"only {float(same.float().mean()):.5f} of the codes agree"
    # where the argmax differs the two centroids score within one fp16 ulp of each other (accumulation order)
    bad = (~same).nonzero().flatten()
    if bad.numel():
        sc = x[bad].float() @ cent.float().t()
        a = sc.gather(1, codes[bad, None]).half().float()
        b = sc.gather(1, ref_codes[bad, None]).half().float()
        assert float((a - b).abs().max()) <= 2.0 ** -10
    # residual bytes: reference packing of the residuals w.r.t. OUR codes (create.rs:413-427)
    r = x - cent.index_select(0, codes)
    bk = torch.bucketize(r, cutoffs, out_int32=True, right=False)
    bits = bk.unsqueeze(-1).expand(n, 128, nbits).bitwise_right_shift(torch.arange(nbits, dtype=torch.int8)) & 1
    ref_packed = io.packbits(bits.flatten()).reshape(n, 128 * nbits // 8)
    assert torch.equal(packed, ref_packed)


def test_create_on_gpu_uses_the_kernels_and_matches_the_cpu_builder(tmp_path, cuda_device):
    from fast_plaid_b200 import search
    from fast_plaid_b200.index import store

    docs = make_docs(200, 10, 50, seed=41)
    a = search.FastPlaid(str(tmp_path / "gpu"), device=cuda_device)
    a.create(docs, kmeans_niters=2, seed=5)
    da = store.read_index(str(tmp_path / "gpu"))
    # same centroids and cutoffs through the oracle's CPU encoder
    oidx, extra = io.build_index(docs, da.centroids, nbits=4, seed=5)
    agree = (oidx.doc_codes == da.doc_codes).float().mean()
    assert float(agree) > 0.995
    assert da.doc_residuals.shape == oidx.doc_residuals.shape
    a.close()
