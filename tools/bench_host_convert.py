"""Host-side cost of turning the [B, top_k] result tensors into list[list[(id, score)]] (micro-benchmark)."""
import time
import torch

B, K = 64, 100
ids = torch.randint(0, 10**6, (B, K)); sc = torch.rand(B, K); cnt = torch.full((B,), K, dtype=torch.int32)
if torch.cuda.is_available():
    ids, sc, cnt = ids.pin_memory(), sc.pin_memory(), cnt.pin_memory()


def t(f, n=300):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3


def per_row():
    ids_l = ids.tolist(); sc_l = sc.tolist(); out = []
    for b, n in enumerate(cnt.tolist()): out.append(list(zip(ids_l[b][:n], sc_l[b][:n])))
    return out


def flat():
    k = ids.shape[1]
    fl = list(zip(ids.reshape(-1).tolist(), sc.reshape(-1).tolist()))
    return [fl[b * k: b * k + n] for b, n in enumerate(cnt.tolist())]


def flat_np():
    k = ids.shape[1]
    fl = list(zip(ids.numpy().ravel().tolist(), sc.numpy().ravel().tolist()))
    return [fl[b * k: b * k + n] for b, n in enumerate(cnt.tolist())]


for r in range(3):
    print("per_row %.3f ms  flat %.3f ms  flat_np %.3f ms  threads %d" % (t(per_row), t(flat), t(flat_np), torch.get_num_threads()))
q = torch.randn(64, 32, 128); pin = torch.empty(64, 32, 128, dtype=torch.float16)
if torch.cuda.is_available(): pin = pin.pin_memory()
print("cast into pinned %.3f ms" % t(lambda: pin.copy_(q)))
