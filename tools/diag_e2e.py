"""Per-step timing of the host-buffer path (diagnosis): search_host vs the Python result conversion."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fast_plaid_b200.engine import DeviceIndex
from fast_plaid_b200.index.synthetic import synthetic_index
from fast_plaid_b200.search.fast_plaid import _results_to_lists

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
data, base = synthetic_index(cfg["n_docs"], cfg["doc_len"], 128, 4, "cuda:0", 1234)
didx = DeviceIndex(data, "cuda:0"); del data
params = DeviceIndex.make_params(cfg["top_k"], 4096, 8)
q_host = bench.make_query_batches(didx, cfg, 4, "cuda:0")
q16 = q_host.to(torch.float16).to("cuda:0")
for w in range(3):
    _results_to_lists(*didx.search_host(q_host[w % 4], params))
ts, tc, td = [], [], []
for s in range(24):
    t0 = time.perf_counter()
    out = didx.search_host(q_host[s % 4], params)
    t1 = time.perf_counter()
    res = _results_to_lists(*out)
    t2 = time.perf_counter()
    ts.append((t1 - t0) * 1e3); tc.append((t2 - t1) * 1e3)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    r = didx.search(q16[s % 4], params); torch.cuda.synchronize()
    td.append((time.perf_counter() - t3) * 1e3)
print("search_host ms:", " ".join(f"{x:.2f}" for x in ts))
print("convert ms    :", " ".join(f"{x:.2f}" for x in tc))
print("device search :", " ".join(f"{x:.2f}" for x in td))
