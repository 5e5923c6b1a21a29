#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its headline config (relpose_5pt essential, 10 000 2D-2D corrs, 30 % inliers,
max 100 000 iterations), one JSON line.

A step = one pass of the LO-RANSAC hot path over a batch of `--pairs` independent synthetic image pairs of that
config (own data seed each), per GPU.  Metric = RANSAC hypotheses/s (models passed to score_model) with
scored-correspondences/s alongside.
  value : whole job with the correspondences already resident in HBM (plb_resident_create handles)
  e2e   : the same batch through the reference-facing C-ABI call with HOST buffers: the host->device copy of the
          correspondences / sample tables and the device->host read of records, models and inlier masks are inside
          the timed region (counted from the copies the engine makes)
  --impl reference : the CPU restatement of the reference path (oracle/, the reference itself needs Eigen3, which
          this image lacks) on the host cores, one problem per thread, same config/metric; the reference's own sources
          built on mini-Eigen (oracle/_ref/libplref2.so) are timed beside it as `reference_sources` (informational).
Multi-GPU (torchrun): independent image pairs are sharded across ranks, no data-path collective; weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from poselib_b200 import problem_generator as G  # noqa: E402

BYTES_PER_CORR_FP64 = 32  # the exact-mode kernel reads 4 fp64 per 2D-2D correspondence


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def usable_cores():
    """Host cores this process may actually use: affinity mask and cgroup CPU quota (the box reports 128 logical CPUs
    but the container may be limited to fewer)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def make_batch(pairs, first_idx):
    probs = []
    for i in range(pairs):
        p = G.config_c2(first_idx + i)
        probs.append((np.ascontiguousarray(p["x1"] / G.FOCAL), np.ascontiguousarray(p["x2"] / G.FOCAL)))
    return probs


def totals(results):
    hyp = sum(r["counters"]["hypotheses"] for r in results)
    cor = sum(r["counters"]["scored_corrs"] for r in results)
    smp = sum(r["counters"]["samples"] for r in results)
    return hyp, cor, smp


def run_reference(args, rank, world):
    """CPU arm: oracle restatement, all host threads, one problem per thread (the reference itself is single-threaded)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import plo_py as P
    threads = usable_cores()
    pairs = max(threads, min(args.pairs, 2 * threads))  # bounded sample of the same workload, >= one problem per core
    batch = make_batch(pairs, 0)
    opts = [P.RansacOpt(max_iterations=100000, min_iterations=1000, seed=0) for _ in range(pairs)]
    me = [1.0 / G.FOCAL] * pairs
    x1, x2 = [b[0] for b in batch], [b[1] for b in batch]
    for _ in range(args.warmup):
        P.ransac_relpose_batch_mt(x1[:threads], x2[:threads], opts[:threads], me[:threads], threads)
    t_tot, hyp, cor, smp = 0.0, 0, 0, 0
    for _ in range(args.steps):
        sec, _, stats, cnts = P.ransac_relpose_batch_mt(x1, x2, opts, me, threads)
        t_tot += sec
        hyp += sum(c["hypotheses"] for c in cnts)
        cor += sum(c["scored_corrs"] for c in cnts)
        smp += sum(c["samples"] for c in cnts)
    val = hyp / t_tot
    try:  # the -march=native build of the same port beside the reference-flags build (built on this box)
        nsec, _, nstats, ncnts = P.ransac_relpose_batch_mt(x1, x2, opts, me, threads, native=True)
        same = sum(int(a["iterations"] == b["iterations"] and a["num_inliers"] == b["num_inliers"]) for a, b in zip(nstats, stats))
        native = {"value": sum(c["hypotheses"] for c in ncnts) / nsec, "unit": "hypotheses/s", "cores": threads,
                  "kind": "port", "flags": "g++ -O3 -march=native -ffp-contract=off",
                  "same_trajectory_as_reference_flags_build": f"{same}/{len(x1)}"}
    except Exception as e:
        native = {"unavailable": f"{type(e).__name__}: {e}"}
    try:
        ref_src = reference_sources_leg(P, x1, x2, opts, me, threads, stats, cnts)
    except Exception as e:  # informational leg only: never let it take the arm's line down
        ref_src = {"unavailable": f"{type(e).__name__}: {e}"}
    line = {
        "impl": "reference", "metric": "RANSAC hypotheses/sec (5pt E, 10k corrs)", "value": val, "unit": "hypotheses/s",
        "scored_corrs_per_s": cor / t_tot, "samples_per_s": smp / t_tot, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{pairs} x relpose_5pt C2 (10000 corrs, 30% inliers, max 100000 its), one problem per host thread",
                   "pairs_per_step": pairs},
        "cpu_baseline": {"value": val, "unit": "hypotheses/s", "cores": threads, "host_cpus": os.cpu_count(), "cpu_model": cpu_model(), "kind": "port",
                         "sample": f"{pairs} C2 problems per step x {args.steps} steps; restated PoseLib path (no Eigen), "
                                   "g++ -O3 -ffp-contract=off"},
        "e2e": {"value": val, "unit": "hypotheses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    line["cpu_baseline_native"] = native
    if ref_src:
        line["reference_sources"] = ref_src
    print(json.dumps(line))


def reference_sources_leg(P, x1, x2, opts, me, threads, stats, cnts):
    """Informational: the reference's OWN sources (oracle/_ref/libplref2.so: robust/ransac.cc, estimators, solvers,
    scorers, bundle, compiled unmodified on mini-Eigen) on the same problems, one problem per thread.  That build pays for
    mini-Eigen's heap temporaries and bounds checks, so it is slower than the reference with real Eigen would be; the
    arm's `value` therefore stays the faster oracle port and this number is reported beside it, not instead of it."""
    if not P.ref2_available(build_if_possible=False):
        return None
    from concurrent.futures import ThreadPoolExecutor
    with P.reference_sources():
        P.ransac("relpose", x1[0], x2[0], opts[0], me[0])  # load + warm
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:  # ctypes releases the GIL inside the call
            out = list(ex.map(lambda i: P.ransac("relpose", x1[i], x2[i], opts[i], me[i])["stats"], range(len(x1))))
        sec = time.perf_counter() - t0
    same = sum(int(o["iterations"] == s["iterations"] and o["num_inliers"] == s["num_inliers"]) for o, s in zip(out, stats))
    hyp = sum(c["hypotheses"] for c in cnts)  # same trajectory => same hypotheses; counted by the oracle run
    return {"value": hyp / sec, "unit": "hypotheses/s", "cores": threads, "kind": "reference",
            "problems": len(x1), "same_trajectory_as_port": f"{same}/{len(x1)}",
            "note": "PoseLib sources unmodified on mini-Eigen (no Eigen3 in this image); slower than a real-Eigen build"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=256, help="independent C2 image pairs per GPU per step")
    ap.add_argument("--streams", type=int, default=12, help="lock-step problem groups in flight per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="fast", choices=["exact", "fast"],
                    help="fast: fp32 SMEM screening of every model + fp64 confirmation of candidates (identical results)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    # ranks of one node share the host cores: tell the engine its share (helper threads for the sample tables)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    os.environ.setdefault("PLB_HOST_THREADS", str(max(1, usable_cores() // max(1, local_world))))
    from poselib_b200 import cabi
    if not torch.cuda.is_available() or cabi.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    cabi.set_device(local_rank)
    cabi.set_mode(args.mode)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pairs = args.pairs
    batch = make_batch(pairs, rank * pairs)  # every rank gets its own image pairs (weak scaling)
    n = len(batch[0][0])
    ropt = dict(max_iterations=100000, min_iterations=1000, seed=0)
    host_probs = [dict(kind="relpose", a=a, b=b, ransac=cabi.RansacOpt(**ropt), max_error=1.0 / G.FOCAL) for a, b in batch]
    handles = [cabi.resident_create("relpose", a, b) for a, b in batch]
    res_probs = [dict(kind="relpose", resident=h, n=n, ransac=cabi.RansacOpt(**ropt), max_error=1.0 / G.FOCAL)
                 for h in handles]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    from poselib_b200 import sharding
    my_idx = list(range(rank * pairs, (rank + 1) * pairs))

    def timed(probs, steps, gather=False):
        t_tot, agg, last = 0.0, None, None
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations (untimed)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            last = cabi.ransac_batch(probs, streams=args.streams)  # returns after its own stream syncs
            if gather and dist is not None:
                # the only inter-GPU traffic of the path: fixed-size result records gathered over NCCL
                sharding.gather_records(sharding.pack_results(list(zip(my_idx, last))), dist)
            torch.cuda.synchronize()
            t_tot += time.perf_counter() - t0
            c = {k: sum(r["counters"][k] for r in last) for k in last[0]["counters"]}
            agg = c if agg is None else {k: agg[k] + c[k] for k in c}
        return t_tot, agg, last

    for _ in range(args.warmup):
        cabi.ransac_batch(res_probs, streams=args.streams)
        cabi.ransac_batch(host_probs, streams=args.streams)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_res, c_res, last_value = timed(res_probs, args.steps)
    barrier()
    t_e2e, c_e2e, last = timed(host_probs, args.steps, gather=True)
    barrier()
    # roofline pass: the same workload with ONE lock-step group in flight, so that the CUDA-event duration of the scoring
    # kernel is not inflated by kernels of other groups sharing the GPU (still live, still on the engine's stream)
    streams_saved = args.streams
    args.streams = 1
    t_roof, c_roof, _ = timed(res_probs, max(3, args.steps // 2))
    args.streams = streams_saved
    barrier()
    # cross-check pass: the same resident batch in the OTHER precision mode (fp64 scoring of every model when the headline
    # ran with fp32 screening, and vice versa).  Reported beside the headline together with whether every result
    # (stats, model bits, inlier mask) is identical.  No collective inside the try block: a failure here must not be
    # able to desynchronise the ranks or lose the headline line.
    other_mode = "exact" if args.mode == "fast" else "fast"
    t_other, hyp_other, steps_other, same_other = float("nan"), 0.0, max(2, args.steps // 3), 0.0
    try:
        cabi.set_mode(other_mode)
        cabi.ransac_batch(res_probs, streams=args.streams)  # buffers of the other mode
        t_other, c_other, last_other = timed(res_probs, steps_other)
        hyp_other = float(c_other["hypotheses"])
        same_other = float(all(a["stats"] == b["stats"] and np.array_equal(a["model"], b["model"]) and
                               np.array_equal(a["inliers"], b["inliers"]) for a, b in zip(last_value, last_other)))
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"[bench] cross-check pass failed: {e}\n")
    finally:
        cabi.set_mode(args.mode)
    barrier()
    sampler.stop_flag = True

    def allmax(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if dist is None:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    T_res, T_e2e = allmax(t_res), allmax(t_e2e)
    T_other = allmax(t_other if t_other == t_other else 1e30)
    hyp_o = allsum(hyp_other)
    same_o = allsum(same_other)
    hyp, cor, smp = allsum(c_res["hypotheses"]), allsum(c_res["scored_corrs"]), allsum(c_res["samples"])
    hyp_e, cor_e = allsum(c_e2e["hypotheses"]), allsum(c_e2e["scored_corrs"])
    launches = allsum(c_res["gpu_launches"])
    if rank == 0:
        peak, peak_kind = load_peaks()
        # roofline of the scoring kernel (the streaming map-reduce of SURVEY §8d):
        #   fast mode : k_screen  — fp32 SoA, 16 B / scored correspondence, staged in shared memory by TMA
        #   exact mode: k_score_tiled — fp64 SoA, 32 B / scored correspondence
        # algorithmic bytes = models scored x N x bytes/corr ; duration = CUDA events around the scoring launches on
        # the engine's stream.  The 5-point solver kernels (latency/issue bound, no streaming) are timed beside it.
        bpc = 16 if args.mode == "fast" else BYTES_PER_CORR_FP64
        roof_steps = max(3, args.steps // 2)
        alg_bytes = c_roof["models_evaluated"] * n * bpc
        k_sec = c_roof["gpu_seconds_score"]
        k_sec_all = c_roof["gpu_seconds"]
        ach = alg_bytes / k_sec / 1e9 if k_sec > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_scoring_kernel.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.mode, {}).get("dram_bytes_per_launch")
        line = {
            "metric": "RANSAC hypotheses/sec (5pt E, 10k corrs)", "value": hyp / T_res, "unit": "hypotheses/s",
            "scored_corrs_per_s": cor / T_res, "samples_per_s": smp / T_res,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.mode == "exact" else "f64 (fp32 screening of all models, fp64 confirmation + solvers + LO)",
            "data": "synthetic",
            "config": {"workload": f"{pairs} x relpose_5pt C2 (10000 corrs, 30% inliers, max 100000 its) per GPU per step",
                       "pairs_per_step_per_gpu": pairs, "streams": args.streams, "mode": args.mode,
                       "l2": "flushed (256 MiB write) between timed steps", "timing": "host clock around synchronous "
                       "C-ABI calls, cuda synchronize both sides, max over ranks; kernel time by CUDA events"},
            "e2e": {"value": hyp_e / T_e2e, "unit": "hypotheses/s", "scored_corrs_per_s": cor_e / T_e2e,
                    "includes": "host buffers in, results out" + ("; NCCL all_gather of the result records" if world > 1 else ""),
                    "ms_per_step": 1e3 * T_e2e / args.steps,
                    "h2d_bytes_per_step": c_e2e["h2d_bytes"] // args.steps, "d2h_bytes_per_step": c_e2e["d2h_bytes"] // args.steps},
            "gpu_launches": int(launches),
            "other_mode": {"mode": other_mode, "value": (hyp_o / T_other) if T_other < 1e29 else None,
                           "unit": "hypotheses/s", "steps": steps_other,
                           "ms_per_step": (1e3 * T_other / steps_other) if T_other < 1e29 else None,
                           "results_identical_to_headline_mode": bool(same_o == world),
                           "note": "same resident batch; stats, model bits and inlier masks compared problem by problem"},
            "roofline": {"bound": "hbm", "kernel": "k_screen<relpose> (fp32 MSAC screening, TMA-staged SMEM)" if args.mode == "fast"
                         else "k_score_tiled<relpose> (fp64 MSAC scoring)", "achieved": ach,
                         "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "bytes_per_scored_corr": bpc,
                         "algorithmic_bytes_per_step": alg_bytes / roof_steps, "kernel_seconds_per_step": k_sec / roof_steps,
                         "solver_kernels_seconds_per_step": (k_sec_all - k_sec) / roof_steps,
                         "kernel_share_of_step": k_sec / t_roof if t_roof > 0 else None,
                         "measured": f"{roof_steps} extra steps of the same workload with one lock-step group in flight "
                                     "(kernel durations by CUDA events on the engine's stream, no co-running kernels)",
                         "note": "correspondences are SMEM/L2-resident and reused across thousands of models: DRAM "
                                 "traffic << algorithmic bytes by design (SURVEY H7); frac is the SURVEY §8d figure and "
                                 "can exceed 1 for the fp32 screening kernel: its operand stream is served from shared "
                                 "memory after one TMA stage per CTA, the binding resource is fp32 issue (60 % of the "
                                 "slots, profiles/r01_v7_summary.md)"},
            "clocks": sampler.summary(),
        }
        # CPU baseline on this box's host cores: 1 thread (the reference's execution model), bounded sample
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import plo_py as P
            k = min(8, pairs)
            sec, _, _, cnts = P.ransac_relpose_batch_mt([b[0] for b in batch[:k]], [b[1] for b in batch[:k]],
                                                         [P.RansacOpt(**ropt) for _ in range(k)], [1.0 / G.FOCAL] * k, 1)
            line["cpu_baseline"] = {"value": sum(c["hypotheses"] for c in cnts) / sec, "unit": "hypotheses/s",
                                    "scored_corrs_per_s": sum(c["scored_corrs"] for c in cnts) / sec, "cores": 1,
                                    "kind": "port", "sample": f"first {k} problems of the step, 1 thread, restated "
                                    "PoseLib path (no Eigen), g++ -O3 -ffp-contract=off", "seconds": sec,
                                    "host_cpus": os.cpu_count(), "usable_cores": usable_cores(), "cpu_model": cpu_model()}
        except Exception as e:  # the baseline is a reported number, never part of the product path
            line["cpu_baseline"] = {"value": None, "unit": "hypotheses/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line))
    for h in handles:
        cabi.resident_free(h)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
