"""N>1 host logic on CPU: world_size-2 gloo.  Problems are partitioned across ranks with no data-path collective and
the fixed-size result records are gathered back in problem order.  The per-problem solve is the CPU oracle here (the
GPU path is exercised by the -m gpu tests and by bench.py under torchrun)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import plo_py as P
    from poselib_b200 import problem_generator as G, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    probs = []
    for i in range(6):
        if i % 2 == 0:
            p = G.abspose_problem(100, 0.6, 5, i)
            probs.append(("pnp", p["x"] / G.FOCAL, p["X"], 12.0 / G.FOCAL, dict(max_iterations=200, min_iterations=200, seed=i)))
        else:
            p = G.relpose_problem(300, 0.6, 5, i)
            probs.append(("relpose", p["x1"] / G.FOCAL, p["x2"] / G.FOCAL, 1.0 / G.FOCAL,
                          dict(max_iterations=500, min_iterations=100, seed=i)))
    costs = [sharding.expected_cost(k, len(a), kw["max_iterations"]) for k, a, b, me, kw in probs]
    parts = sharding.partition(costs, world)
    mine = parts[rank]
    res = []
    for i in mine:
        k, a, b, me, kw = probs[int(i)]
        res.append((int(i), P.ransac(k, a, b, P.RansacOpt(**kw), me)))
    rec = sharding.gather_records(sharding.pack_results(res), dist)
    masks = sharding.gather_masks([r["inliers"] for _, r in res], [i for i, _ in res], dist)
    eq = sharding.gather_records_equal(np.array([[10.0 * rank + j, rank] for j in range(3)]), dist)
    assert eq[:, 0].tolist() == [0.0, 1.0, 2.0, 10.0, 11.0, 12.0] and eq[:, 1].tolist() == [0, 0, 0, 1, 1, 1]
    if rank == 0:
        q.put((rec, [p.tolist() for p in parts], {k: v.tolist() for k, v in masks.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_balanced_and_complete():
    sys.path.insert(0, ROOT)
    from poselib_b200 import sharding
    costs = [200.0 * 1000, 1e4 * 2e4] * 8
    for ws in (1, 2, 4, 8):
        parts = sharding.partition(costs, ws)
        allidx = np.sort(np.concatenate(parts))
        assert allidx.tolist() == list(range(16))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) <= 1.3 * (sum(costs) / ws) + max(costs)


def test_two_rank_gloo_shard_and_gather():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import plo_py as P
    from poselib_b200 import problem_generator as G
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rec, parts, masks = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(parts[0] + parts[1]) == list(range(6)) and parts[0] and parts[1]
    assert rec[:, 0].astype(int).tolist() == list(range(6))
    # the gathered records equal a single-process run
    for i in range(6):
        if i % 2 == 0:
            p = G.abspose_problem(100, 0.6, 5, i)
            o = P.ransac("pnp", p["x"] / G.FOCAL, p["X"], P.RansacOpt(max_iterations=200, min_iterations=200, seed=i), 12.0 / G.FOCAL)
        else:
            p = G.relpose_problem(300, 0.6, 5, i)
            o = P.ransac("relpose", p["x1"] / G.FOCAL, p["x2"] / G.FOCAL,
                         P.RansacOpt(max_iterations=500, min_iterations=100, seed=i), 1.0 / G.FOCAL)
        assert rec[i, 1] == o["stats"]["iterations"] and rec[i, 3] == o["stats"]["num_inliers"]
        assert np.allclose(rec[i, 5:12], o["model"], rtol=0, atol=0)
        assert masks[i] == np.asarray(o["inliers"]).astype(int).tolist()  # bit-packed masks gathered from both ranks
