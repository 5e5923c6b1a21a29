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


def gather_records_equal(local_rec, dist=None):
    """gather_records when every rank holds the same number of records (weak scaling): ONE all_gather into a preallocated
    tensor and one device-to-host copy, no size exchange."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = local_rec
    else:
        import torch
        ws = dist.get_world_size()
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        loc = torch.from_numpy(np.ascontiguousarray(local_rec)).to(dev, non_blocking=True)
        buf = torch.empty((ws * loc.shape[0], loc.shape[1]), dtype=loc.dtype, device=dev)
        dist.all_gather_into_tensor(buf, loc)
        out = buf.cpu().numpy()
    return out[np.argsort(out[:, 0], kind="stable")]


def gather_masks(local_masks, local_idx, dist=None):
    """all_gather of the inlier masks as bits.  local_masks: list of 0/1 arrays (one per local problem), local_idx: their
    global problem indices.  Returns {problem index: mask (uint8 0/1)} on every rank."""
    sizes = np.array([len(m) for m in local_masks], dtype=np.int64)
    bits = [np.packbits(np.asarray(m, dtype=np.uint8)) for m in local_masks]
    blob = np.concatenate(bits) if bits else np.zeros(0, dtype=np.uint8)
    head = np.stack([np.asarray(local_idx, dtype=np.int64), sizes], axis=1) if len(local_masks) else np.zeros((0, 2), np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        heads, blobs = [head], [blob]
    else:
        import torch
        ws = dist.get_world_size()
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        meta = torch.tensor([head.shape[0], blob.shape[0]], dtype=torch.int64, device=dev)
        metas = [torch.zeros_like(meta) for _ in range(ws)]
        dist.all_gather(metas, meta)
        mh = int(max(m[0].item() for m in metas))
        mb = int(max(m[1].item() for m in metas))
        hp = torch.zeros((max(mh, 1), 2), dtype=torch.int64, device=dev)
        hp[:head.shape[0]] = torch.from_numpy(head).to(dev)
        bp = torch.zeros(max(mb, 1), dtype=torch.uint8, device=dev)
        bp[:blob.shape[0]] = torch.from_numpy(blob).to(dev)
        hs = [torch.zeros_like(hp) for _ in range(ws)]
        bs = [torch.zeros_like(bp) for _ in range(ws)]
        dist.all_gather(hs, hp)
        dist.all_gather(bs, bp)
        heads = [h[:int(m[0].item())].cpu().numpy() for h, m in zip(hs, metas)]
        blobs = [b[:int(m[1].item())].cpu().numpy() for b, m in zip(bs, metas)]
    out = {}
    for h, b in zip(heads, blobs):
        off = 0
        for idx, n in h:
            nb = (int(n) + 7) // 8
            out[int(idx)] = np.unpackbits(b[off:off + nb])[:int(n)]
            off += nb
    return out
