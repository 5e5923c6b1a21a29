"""Problem-granularity sharding of independent image pairs across ranks.

No data-path collective: every rank runs its own slice through the C-ABI; only fixed-size result records are gathered.
The partition is size-balanced (longest-processing-time first on N x expected iterations) and deterministic."""
import numpy as np


def expected_cost(kind, n, max_iterations):
    """Rough relative cost of one problem: correspondences x iterations the loop is expected to run."""
    return float(n) * float(min(max_iterations, 20000 if kind == "relpose" else max_iterations))


def partition(costs, world_size):
    """LPT partition of problem indices over ranks; returns list (per rank) of sorted index arrays."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world_size)
    parts = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i))
        load[r] += costs[i]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def pack_results(results, model_len=9):
    """Fixed-size record per problem: [index, iterations, refinements, num_inliers, model_score, model(9)]."""
    rec = np.zeros((len(results), 5 + model_len))
    for j, (idx, r) in enumerate(results):
        m = np.asarray(r["model"], dtype=np.float64).reshape(-1)
        rec[j, 0] = idx
        rec[j, 1] = r["stats"]["iterations"]
        rec[j, 2] = r["stats"]["refinements"]
        rec[j, 3] = r["stats"]["num_inliers"]
        rec[j, 4] = r["stats"]["model_score"]
        rec[j, 5:5 + len(m)] = m
    return rec


def gather_records(local_rec, dist=None):
    """all_gather of the per-rank record blocks (variable row counts) -> records sorted by problem index."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = local_rec
    else:
        import torch
        ws = dist.get_world_size()
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        cnt = torch.tensor([local_rec.shape[0]], dtype=torch.int64, device=dev)
        cnts = [torch.zeros_like(cnt) for _ in range(ws)]
        dist.all_gather(cnts, cnt)
        mx = int(max(c.item() for c in cnts))
        pad = torch.zeros((mx, local_rec.shape[1]), dtype=torch.float64, device=dev)
        pad[:local_rec.shape[0]] = torch.from_numpy(local_rec).to(dev)
        bufs = [torch.zeros_like(pad) for _ in range(ws)]
        dist.all_gather(bufs, pad)
        out = np.concatenate([b[:int(c.item())].cpu().numpy() for b, c in zip(bufs, cnts)])
    return out[np.argsort(out[:, 0], kind="stable")]
