"""One-off measurement: BASELINE config 3's corpus shape (1M documents x 300 tokens x 128 dims)
built by FastPlaid.create() on one B200 -- k-means on fpb_kmeans_assign / fpb_kmeans_update, streaming chunk encode on
fpb_encode, the documents a lazy seeded sequence -- then loaded by the direct-to-device loader, whole and as one
shard of eight.  Writes gpurun_out/build_<n_docs>.json after every phase."""
import json, os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_plaid_b200 import search
from fast_plaid_b200.index import store
from fast_plaid_b200.index.synthetic import SyntheticDocuments

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
path = "/tmp/fpb_index_1m"
out = {"n_docs": n_docs, "doc_len": 300}


def dump():
    json.dump(out, open(f"gpurun_out/build_{n_docs}.json", "w"), indent=1)


free = shutil.disk_usage("/tmp").free
out["disk_free_gb"] = round(free / 1e9, 1)
need = n_docs * 300 * 72 * 1.1
if free < need:
    out["error"] = f"not enough disk for the index directory ({need / 1e9:.0f} GB needed)"
    dump(); sys.exit(0)
shutil.rmtree(path, ignore_errors=True)
docs = SyntheticDocuments(n_docs, 300, device="cuda:0", seed=11, clusters=16384)
fp = search.FastPlaid(path, device="cuda:0")
t0 = time.time()
fp.create(docs, kmeans_niters=4, seed=42)
torch.cuda.synchronize()
out["create_s"] = round(time.time() - t0, 1)
dump()
meta = store.read_metadata(path)
out["metadata"] = {k: meta[k] for k in ("num_documents", "num_embeddings", "num_partitions", "num_chunks")}
out["dir_gb"] = round(sum(os.path.getsize(os.path.join(path, f)) for f in os.listdir(path)) / 1e9, 2)
fp.close(); del fp
torch.cuda.empty_cache()
t0 = time.time()
fp = search.FastPlaid(path, device="cuda:0")
torch.cuda.synchronize()
out["load_whole_index_s"] = round(time.time() - t0, 1)
dump()
g = torch.Generator().manual_seed(5)
src = sorted(torch.randint(0, n_docs, (64,), generator=g).tolist())
q = torch.stack([torch.nn.functional.normalize(docs[d].float().cpu()[torch.randint(0, 300, (32,), generator=g)]
                                               + 0.05 * torch.randn(32, 128, generator=g), dim=-1) for d in src])
res = fp.search(q, top_k=100)
t0 = time.time()
for _ in range(5):
    res = fp.search(q, top_k=100)
out["search_64_queries_ms"] = round((time.time() - t0) / 5 * 1e3, 2)
out["source_doc_is_top1"] = sum(int(d == r[0][0]) for d, r in zip(src, res))
dump()
fp.close(); del fp
torch.cuda.empty_cache()
t0 = time.time()
data, base = store.read_index_to_device(path, "cuda:0", ((n_docs * 3) // 8, (n_docs * 4) // 8))
torch.cuda.synchronize()
out["load_one_shard_of_8_s"] = round(time.time() - t0, 1)
out["shard_tokens"] = int(data.doc_codes.shape[0])
shutil.rmtree(path, ignore_errors=True)
dump()
print(json.dumps(out))
