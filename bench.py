#!/usr/bin/env python
"""bench.py — registrations/sec of the TEASER++ solve() hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Workload (config.workload): BASELINE config C2 — synthetic N=5000 correspondences, 95 % outliers
("ball" outlier model), fixed scale, GNC-TLS, PMC_EXACT — as a batch of independent problems
per step per GPU (weak scaling: every rank owns its own batch; no collective on the data path).

A "step" = one pass of solve() over one batch of --batch problems per GPU.
  value : registrations/s with the inputs already resident in HBM (tzr_solve_batch_dev), CUDA-event timed.
  e2e   : same metric through the host-pointer C-ABI call (tzr_solve_batch): pinned host inputs are copied
          host->device inside the timed region and the solutions + clique index sets are read back.
Inputs per step are larger than L2 (B*240 KB >= 246 MB at the default batch of 1024), so no L2 flush is needed.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_C2 = 5000
OUTLIER_RATIO = 0.95


def bytes_graph(n):
    """Algorithmic bytes of the graph stage per problem: read src+dst (FP64), write the full
    symmetric bitset (rows padded to 64-bit words) and the degree vector."""
    return 48 * n + 8 * n * ((n + 63) // 64) + 4 * n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median of the upper half = clocks under load (the sampler also sees idle gaps)
        s = sorted(sm)
        return {"sm_mhz": float(np.median(s[len(s) // 2:])), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_batch(B, base_seed, synth):
    src = np.empty((B, N_C2, 3))
    dst = np.empty((B, N_C2, 3))
    inl = []
    nb = None
    for b in range(B):
        pr = synth.make_problem(N_C2, OUTLIER_RATIO, 5000 * 1000 + base_seed + b, "ball")
        src[b], dst[b] = pr["src"], pr["dst"]
        inl.append(pr["inliers"])
        nb = pr["noise_bound"]
    return src, dst, inl, nb


def solver_params(mod, nb):
    return mod.default_params(noise_bound=nb, cbar2=1.0, estimate_scaling=0, rotation_estimation_algorithm=0,
                              rotation_gnc_factor=1.4, rotation_max_iterations=100, rotation_cost_threshold=1e-12,
                              rotation_tim_graph=0, inlier_selection_mode=0)


def cpu_reference_sample(n_problems, seed0, synth, budget_s=25.0):
    """Times the CPU restatement of the reference algorithm (oracle, OpenMP on all host cores) on a bounded
    sample of the same workload.  Returns (regs_per_s, cores, n_done, seconds)."""
    import oracle_lib as orc
    cores = orc.lib().orc_num_threads()
    pr = synth.make_problem(N_C2, OUTLIER_RATIO, 5000 * 1000 + seed0, "ball")
    p = solver_params(orc, pr["noise_bound"])
    orc.solve(pr["src"], pr["dst"], p)  # warm-up (page faults, thread pool)
    done, t_total = 0, 0.0
    for i in range(n_problems):
        pr = synth.make_problem(N_C2, OUTLIER_RATIO, 5000 * 1000 + seed0 + 1 + i, "ball")
        t0 = time.perf_counter()
        out = orc.solve(pr["src"], pr["dst"], p)
        t_total += time.perf_counter() - t0
        done += 1
        assert out["valid"]
        if t_total > budget_s:
            break
    return done / t_total, cores, done, t_total


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the true reference cannot be compiled here:
    Eigen/PMC absent) on the host cores, same workload/metric.  Rank 0 only."""
    if rank != 0:
        return
    synth = importlib.import_module("teaser-plusplus_b200.synth")
    import oracle_lib as orc
    cores = orc.lib().orc_num_threads()
    per_step = max(1, args.ref_problems_per_step)
    pr = synth.make_problem(N_C2, OUTLIER_RATIO, 5000 * 1000, "ball")
    p = solver_params(orc, pr["noise_bound"])
    for w in range(min(args.warmup, 1) or 1):
        orc.solve(pr["src"], pr["dst"], p)
    t_total, done = 0.0, 0
    for k in range(args.steps):
        for i in range(per_step):
            q = synth.make_problem(N_C2, OUTLIER_RATIO, 5000 * 1000 + 1 + k * per_step + i, "ball")
            t0 = time.perf_counter()
            orc.solve(q["src"], q["dst"], p)
            t_total += time.perf_counter() - t0
            done += 1
    v = done / t_total
    line = {
        "impl": "reference", "metric": "registrations/sec", "value": v, "unit": "registrations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2: N={N_C2} correspondences, {int(OUTLIER_RATIO*100)}% outliers (ball), fixed scale, "
                               f"GNC-TLS, PMC_EXACT; {per_step} problem(s) per step on the host cores"},
        "cpu_baseline": {"value": v, "unit": "registrations/s", "cores": cores, "kind": "port",
                         "sample": f"{done} problems of the C2 workload, solved back to back with OpenMP on {cores} "
                                   f"threads (reference restatement; Eigen/PMC unavailable so the true reference "
                                   f"cannot be built)"},
        "e2e": {"value": v, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="problems per step per GPU")
    ap.add_argument("--ref-problems-per-step", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    capi = importlib.import_module("teaser-plusplus_b200.capi")
    synth = importlib.import_module("teaser-plusplus_b200.synth")
    W = max(args.warmup, 3)
    K = args.steps
    B = args.batch

    # ---- synthetic inputs: each rank its own batch (weak scaling), pinned host copies + device copies
    src_h, dst_h, inliers, nb = make_batch(B, rank * 100000, synth)
    src_pin = torch.empty((B, N_C2, 3), dtype=torch.float64, pin_memory=True)
    dst_pin = torch.empty((B, N_C2, 3), dtype=torch.float64, pin_memory=True)
    src_pin.numpy()[...] = src_h
    dst_pin.numpy()[...] = dst_h
    src_d = src_pin.cuda(non_blocking=False)
    dst_d = dst_pin.cuda(non_blocking=False)
    sol_d = torch.zeros(B * capi.SOLUTION_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    clq_d = torch.zeros((B, N_C2), dtype=torch.int32, device="cuda")
    params = solver_params(capi, nb)
    ctx = capi.Context(local_rank)
    if os.environ.get("TZR_FLAGS"):  # debug / A-B switches of the library (e.g. 8 = previous graph kernel)
        ctx.set_flags(int(os.environ["TZR_FLAGS"]))
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    def step_dev():
        ctx.solve_batch_dev(params, B, N_C2, src_d.data_ptr(), dst_d.data_ptr(), sol_d.data_ptr(), clq_d.data_ptr())

    src_np, dst_np = src_pin.numpy(), dst_pin.numpy()  # page-locked host buffers
    h_sols = np.zeros(B, dtype=capi.SOLUTION_DTYPE)
    h_clq = np.zeros((B, N_C2), dtype=np.int32)

    def step_host():
        # the public host-pointer call: H2D of this step's inputs, all kernels, D2H of solutions + clique sets
        return ctx.solve_batch_array(src_np, dst_np, params, cliques_out=h_clq, sols_out=h_sols)

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(W):
        step_dev()
    ctx.synchronize()
    # correctness of the timed path itself: identical inlier sets on the whole batch
    sols = np.frombuffer(sol_d.cpu().numpy().tobytes(), dtype=capi.SOLUTION_DTYPE)
    clq = clq_d.cpu().numpy()
    n_ok = sum(int(np.array_equal(clq[b, :sols[b]["clique_size"]], inliers[b])) for b in range(B))

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region 1: device-resident inputs
    barrier()
    l0 = ctx.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph_ms = []
    stage_acc = {"prep": 0.0, "graph": 0.0, "clique": 0.0, "rot_trans": 0.0}
    ev_begin, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev_begin.record(stream)
        for _ in range(K):
            step_dev()
            ev1.record(stream)
            ev1.synchronize()  # per-step sync only to read this step's stage events (graph-kernel time for the roofline)
            st = ctx.last_stage_ms()
            graph_ms.append(st["graph"])
            for k_ in stage_acc:
                stage_acc[k_] += st[k_]
        ev_end.record(stream)
        ev_end.synchronize()
        t_dev = ev_begin.elapsed_time(ev_end)  # device time of exactly K steps, host gaps between steps included
    barrier()
    launches = ctx.kernel_launches() - l0
    # ---- timed region 2: end to end through the host-pointer C-ABI
    for _ in range(2):
        step_host()
    barrier()
    t_e2e = 0.0
    for _ in range(K):
        t0 = time.perf_counter()
        hsols, hcl = step_host()
        t_e2e += (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop()

    tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev_max, t_e2e_max = float(tt[0]), float(tt[1])
    e2e_ok = sum(int(np.array_equal(hcl[b, :hsols[b]["clique_size"]], inliers[b])) for b in range(B))

    if rank == 0:
        value = world * K * B / (t_dev_max * 1e-3)
        e2e = world * K * B / (t_e2e_max * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        g_ms = float(np.mean(graph_ms))
        achieved = bytes_graph(N_C2) * B / (g_ms * 1e-3) / 1e9
        traffic = None
        issue = None
        try:
            # measured once with `ncu --set full` (profiles/graph_kernel_traffic.json), scaled to this launch's batch
            prof = json.load(open(os.path.join(ROOT, "profiles", "graph_kernel_traffic.json")))
            traffic = prof.get("dram_bytes_per_problem") * B
            # the binding resource of this kernel is the warp-instruction issue rate, not HBM: warp instructions of one
            # launch (ncu count per problem x B) / live kernel time, against 4 issue slots per SM per clock
            winst = prof["inst_executed"] / prof["batch"] * B
            sm_mhz = float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
            issue_peak = 148 * 4 * sm_mhz * 1e6
            issue = {"achieved_warp_inst_per_s": winst / (g_ms * 1e-3), "peak_warp_inst_per_s": issue_peak,
                     "frac": winst / (g_ms * 1e-3) / issue_peak,
                     "warp_inst_per_pair": prof["inst_executed"] / prof["batch"] / (N_C2 * (N_C2 - 1) / 2)}
        except Exception:
            pass
        line = {
            "metric": "registrations/sec", "value": value, "unit": "registrations/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": t_dev_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"C2: N={N_C2} correspondences, {int(OUTLIER_RATIO*100)}% outliers (ball), fixed scale, "
                            f"GNC-TLS (cost_thr 1e-12), PMC_EXACT; batch {B} problems/step/GPU",
                "global_batch": B * world, "parallelism": f"batch sharded over {world} GPU(s), no collective",
                "l2": f"inputs larger than L2: {2 * B * N_C2 * 24 / 1e6:.0f} MB of points + "
                      f"{B * N_C2 * 80 * 8 / 1e9:.2f} GB of adjacency per step (L2 = 126 MB); no flush",
                "dtype_note": "FP64 predicate/GNC/TLS; graph stage classifies pairs with an FP32 interval filter "
                              "and re-checks the ambiguous band in exact FP64 (bit-identical bitset)",
            },
            "e2e": {"value": e2e, "unit": "registrations/s", "h2d_bytes_per_step": int(2 * B * N_C2 * 24),
                    # what tzr_solve_batch copies back: the solution records + the used prefix of every clique row
                    "d2h_bytes_per_step": int(B * capi.SOLUTION_DTYPE.itemsize
                                              + B * int(hsols["clique_size"].max()) * 4),
                    "ms_per_step": t_e2e_max / K},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "graph_strip2_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": bytes_graph(N_C2) * B, "kernel_ms": g_ms, "issue": issue,
                         "note": "issue-bound FP32 stage (12 FP32 ops as 6 packed FP32x2 instructions + 2 MUFU + 2 FSETP "
                                 "per pair, ~60 op/B): the HBM fraction is reported because SURVEY §8d defines the "
                                 "roofline of this stage against HBM; DRAM traffic is within 0.9x of algorithmic"},
            "stage_ms_per_step": {k_: v / K for k_, v in stage_acc.items()},
            "parity": {"timed_batch_clique_equals_planted_inliers": f"{n_ok}/{B}",
                       "e2e_batch_clique_equals_planted_inliers": f"{e2e_ok}/{B}",
                       "note": "ground-truth check; a planted set can be strictly inside the maximum clique when an "
                               "outlier happens to be consistent with every inlier"},
        }
        # rotation / translation error vs the oracle on identical inputs (metric's second half) + CPU baseline
        if world == 1 and not args.no_cpu_baseline:
            import oracle_lib as orc
            errs = []
            for b in range(2):
                o = orc.solve(src_h[b], dst_h[b], solver_params(orc, nb))
                Rg = capi.rotation_from_solution_record(sols[b])
                errs.append((synth.angular_error(o["R"], Rg), float(np.linalg.norm(o["t"] - sols[b]["translation"])),
                             bool(np.array_equal(o["clique"], clq[b, :sols[b]["clique_size"]]))))
            line["parity"]["vs_oracle"] = {"rot_err_rad_max": max(e[0] for e in errs),
                                           "trans_err_m_max": max(e[1] for e in errs),
                                           "clique_identical": all(e[2] for e in errs), "problems": len(errs)}
            v, cores, done, secs = cpu_reference_sample(12, 777, synth)
            line["cpu_baseline"] = {"value": v, "unit": "registrations/s", "cores": cores, "kind": "port",
                                    "sample": f"{done} problems of the same C2 workload ({secs:.1f} s), OpenMP on "
                                              f"{cores} host threads, reference restatement (Eigen/PMC unavailable)"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
