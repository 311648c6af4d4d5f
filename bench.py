#!/usr/bin/env python
"""Headline benchmark: batched C2A_Solve (controlled conservative-advancement CCD) queries/s.

Workload (BASELINE.json configs[2]): torus knot (512x32 quads, 32768
triangles) against itself, --batch (1M) synthetic interpolated-motion pose pairs IN TOTAL, FP64,
tolerance_d = tolerance_t = 1e-4 as C2A_Solve hard-codes them.  A step = the whole batch through the hot
path.  STRONG scaling: with N GPUs (one rank per GPU) rank r solves the queries order[r], order[r+N], ...
of the batch's cost-sorted claim order; no data-path collective.

  python bench.py --gpus N --steps K --warmup W          this repo's CUDA path
  python bench.py --impl reference ...                   the reference's own CPU code on the host cores

One JSON line on stdout (rank 0).
  value     whole-job queries/s, motion records already resident in HBM, CUDA-event timed per step, max over ranks
  e2e       the same through the library's host-buffer entry: c2a_b200_solve_batch at N = 1, the multi-device
            c2a_b200_solve_batch_multi (called by rank 0 over all N devices: host motion set-up, per-device H2D,
            kernels, D2H, results gathered into rank 0's arrays) at N > 1
  roofline  c2a_solve_kernel + c2a_wide_kernel (the two kernels of a step): algorithmic bytes (208*Nbv + 144*Ntri +
            448 per query, SURVEY.md 8d, from the kernels' own counters, which equal the reference's) / their
            measured duration, against the measured HBM copy bandwidth
  cpu_baseline  oracle/_ref (the reference's object code) or the port, all host cores (dynamic chunk queue) and one
            thread, bounded samples of the same batch
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

METRIC = "ccd_queries_per_sec"
UNIT = "queries/s"
KNOT = (512, 32)
TOL = 1e-4
SEED = 20260002


def workload_name(batch):
    return (f"torusknot({KNOT[0]}x{KNOT[1]} quads, {2 * KNOT[0] * KNOT[1]} tris) vs torusknot, "
            f"{batch} synthetic interpolated-motion pose pairs per step in total, sharded over the GPUs (BASELINE configs[2])")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(nbv, ntri, n):
    return 208.0 * float(nbv) + 144.0 * float(ntri) + 448.0 * float(n)


def nominal_flops(nbv, ntri, nca):
    return 370.0 * float(nbv) + 1000.0 * float(ntri) + 300.0 * float(nca)


def cpu_solver():
    """(callable(poses, threads) -> results, kind).  Prefers the reference's own object code."""
    import oracle
    from c2a_b200 import api, meshes
    tris = meshes.torus_knot(*KNOT)[0]
    if oracle.have_ref():
        R = oracle.ref()
        m = R.model(tris)
        return (lambda poses, threads: R.solve_batch(m, m, poses, tol_d=TOL, tol_t=TOL, threads=threads)), "reference"
    oracle.build_oracle()
    bvh = api.build_bvh(tris)
    P = oracle.port()
    return (lambda poses, threads: P.solve_batch(bvh, bvh, poses, tol_d=TOL, tol_t=TOL, threads=threads)), "port"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores, on bounded samples of
    the GPU arm's own batch (step k takes the k-th block of it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from c2a_b200 import workloads
    cores = os.cpu_count() or 1
    solve, kind = cpu_solver()
    n = args.ref_batch if args.ref_batch > 0 else 768 * cores
    batch = workloads.approach_batch(args.batch, SEED, radius=workloads.KNOT_RADIUS)
    nblk = max(1, args.batch // n)
    times = []
    for it in range(args.warmup + args.steps):
        poses = batch[(it % nblk) * n:(it % nblk) * n + n]
        t0 = time.perf_counter()
        solve(poses, cores)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = args.steps * n / total
    n1 = max(64, n // (4 * cores))
    t0 = time.perf_counter()
    solve(batch[:n1], 1)
    single = n1 / (time.perf_counter() - t0)
    sample = (f"{n} pose pairs per step: consecutive blocks of the GPU arm's {args.batch}-pair batch (same seed), {cores} std::threads "
              f"on a dynamic chunk queue")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.batch), "tolerance_d": TOL, "tolerance_t": TOL,
                       "reference_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "single_thread": {"value": single, "unit": UNIT, "sample": f"first {n1} pose pairs of the batch, 1 thread"}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import ctypes as C
    import torch
    from c2a_b200 import api, build as c2a_build, meshes, sharding, workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    cpu_group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")  # host-side barriers that keep the GPUs free (an NCCL barrier spins on them)
    torch.cuda.set_device(local)
    if rank == 0:
        c2a_build.build()
    if dist:
        dist.barrier()

    B = args.batch
    bvh = api.build_bvh(meshes.torus_knot(*KNOT)[0])
    model = api.Model(bvh, local)
    poses = workloads.approach_batch(B, SEED, radius=workloads.KNOT_RADIUS)  # the same batch on every rank
    fields = ("status", "collisionfree", "num_ca", "num_bv_tests", "num_tri_tests", "toc", "distance", "pose_toc")

    # ---- device-resident leg ("value"): this rank's shard of the batch, motion records in HBM, CUDA events on the
    # launching stream.  Shard = every world-th entry of the cost-sorted claim order, so it is itself cost-sorted.
    motions_all = api.motions_from_poses(poses)
    order_all = api.schedule_order(model, model, motions_all)
    shard = sharding.shard_indices(order_all, rank, world)
    Br = len(shard)
    stream = torch.cuda.Stream()
    motions = torch.from_numpy(np.ascontiguousarray(motions_all[shard])).pin_memory().cuda(non_blocking=True)
    del motions_all
    out = {"status": torch.empty(Br, dtype=torch.int32, device="cuda"), "collisionfree": torch.empty(Br, dtype=torch.int32, device="cuda"),
           "num_ca": torch.empty(Br, dtype=torch.int32, device="cuda"), "num_bv_tests": torch.empty(Br, dtype=torch.int32, device="cuda"),
           "num_tri_tests": torch.empty(Br, dtype=torch.int32, device="cuda"), "toc": torch.empty(Br, dtype=torch.float64, device="cuda"),
           "distance": torch.empty(Br, dtype=torch.float64, device="cuda"), "pose_toc": torch.zeros(Br, 24, dtype=torch.float64, device="cuda")}
    ptrs = {k: v.data_ptr() for k, v in out.items()}
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    torch.cuda.synchronize()

    def step():
        api.solve_batch_device(model, model, motions.data_ptr(), Br, ptrs, tol_d=TOL, tol_t=TOL, stream=stream.cuda_stream)

    for _ in range(args.warmup):
        step()
    stream.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = api.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kt = np.zeros(3)
    ktb = (C.c_double * 3)()
    for k in range(args.steps):
        scrub.fill_(k)  # flush L2 between timed iterations (on torch's stream, not timed)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            ev[k][0].record(stream)
            step()
            ev[k][1].record(stream)
        stream.synchronize()
        if api.lib().c2a_b200_kernel_times(ktb) == 0:
            kt += np.array(list(ktb))
    torch.cuda.synchronize()
    launches = api.launch_count() - launches0
    if dist:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None

    nbv = int(out["num_bv_tests"].sum(dtype=torch.int64).item())
    ntri = int(out["num_tri_tests"].sum(dtype=torch.int64).item())
    nca = int(out["num_ca"].sum(dtype=torch.int64).item())
    assert int((out["status"] != 0).sum().item()) == 0
    hits = int((out["collisionfree"] == 0).sum().item())
    dev_res = {k: out[k].cpu().numpy() for k in ("collisionfree", "toc", "distance", "num_ca", "num_bv_tests")}

    # ---- end-to-end leg: the library's host-buffer entry over the WHOLE batch, every step H2D of the inputs and D2H
    # of the results; at N > 1 rank 0 drives all N devices through the multi-device entry, the other ranks stand by
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_s, host_out, e2e_breakdown = 0.0, None, None
    if dist:
        dist.barrier(group=cpu_group)
    if rank == 0:
        if world == 1:
            def call():
                return api.solve_batch(model, model, poses, tol_d=TOL, tol_t=TOL, fields=fields)
        else:
            replicas = [model] + [api.Model(bvh, d) for d in range(1, world)]

            def call():
                return api.solve_batch_multi(replicas, replicas, poses, tol_d=TOL, tol_t=TOL, fields=fields)
        host_out = call()  # warm (allocator, pinned pools)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_out = call()
        e2e_s = time.perf_counter() - t0
        tb = (C.c_double * 8)()
        api.lib().c2a_b200_host_timing(tb)  # wall-clock breakdown of the last e2e call
        names = (("alloc_s", "motion_setup_s", "claim_order_s", "enqueue_s", "kernel_wait_s", "d2h_s", "release_s", "total_s") if world == 1 else
                 ("_", "motion_and_order_s", "_", "_", "slowest_shard_s", "fastest_shard_s", "_", "total_s"))
        e2e_breakdown = {k: round(v, 4) for k, v in zip(names, tb) if k != "_"}
        for k in ("collisionfree", "toc", "distance", "num_ca"):
            assert np.array_equal(host_out[k][shard], dev_res[k]), k
    if dist:
        dist.barrier(group=cpu_group)
    h2d = B * 48 * 8
    d2h = sum(np.dtype(dt).itemsize * int(np.prod(shape, dtype=np.int64)) * B for name, dt, shape in api.RESULT_FIELDS if name in fields)

    # ---- max over ranks ----
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t[0])
    tot = torch.tensor([nbv, ntri, nca, hits], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    nbv_all, ntri_all, nca_all, hits_all = (float(x) for x in tot)
    value = args.steps * B / (dev_ms_max * 1e-3)

    if rank == 0:
        e2e_value = e2e_steps * B / e2e_s
        peak, peak_src = measured_peaks()
        # the two kernels of a step on this rank (CUDA events inside the library, on the launching stream)
        k_solve_s, k_wide_s = kt[0] * 1e-3 / args.steps, kt[1] * 1e-3 / args.steps
        kern_s = (k_solve_s + k_wide_s) if (k_solve_s + k_wide_s) > 0 else dev_ms * 1e-3 / args.steps
        ach = algorithmic_bytes(nbv, ntri, Br) / kern_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        f1, f2 = C.c_double(0), C.c_double(0)
        api.lib().c2a_b200_fp64_peak(C.byref(f1), C.byref(f2))
        fl = nominal_flops(nbv, ntri, nca) / kern_s / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(B), "tolerance_d": TOL, "tolerance_t": TOL, "batch_total": B,
                           "batch_per_gpu": Br,
                           "l2": "inputs per step (%.0f MB motion records per GPU) exceed the 126 MB L2 at N <= 2; a 256 MB scrub buffer is "
                                 "written between timed steps at every N" % (Br * 384 / 1e6),
                           "parallelism": f"{world} shards of one batch (cost-sorted claim order interleaved), no collective; e2e gathers into rank 0"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "entry": "c2a_b200_solve_batch" if world == 1 else "c2a_b200_solve_batch_multi (rank 0 drives all devices)",
                        "breakdown_last_call": e2e_breakdown},
                "gpu_launches": launches,
                "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                             "kernel": "c2a_solve_kernel + c2a_wide_kernel (one step of rank 0)", "peak_source": peak_src,
                             "kernel_ms": {"c2a_solve_kernel": 1e3 * k_solve_s, "c2a_wide_kernel": 1e3 * k_wide_s,
                                           "c2a_translation_kernel": kt[2] / args.steps, "step": dev_ms / args.steps},
                             "note": "BVH working set (16.5 MB) is L2-resident by design: DRAM traffic << algorithmic bytes; "
                                     "the binding resource is the FP64 pipe (see fp64)"},
                "fp64": {"achieved_tflops_nominal": fl, "peak_tflops_fma": f1.value, "peak_tflops_mul_add": f2.value,
                         "frac_of_mul_add_peak": fl / f2.value if f2.value else None},
                "bvtt_pairs_per_sec": args.steps * nbv_all / (dev_ms_max * 1e-3),
                "per_query": {"num_bv_tests": nbv_all / B, "num_tri_tests": ntri_all / B, "num_ca": nca_all / B, "hit_fraction": hits_all / B}}
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            solve, kind = cpu_solver()
            ns = min(B, args.cpu_sample if args.cpu_sample > 0 else 768 * cores)
            t0 = time.perf_counter()
            ref = solve(poses[:ns], cores)
            dt = time.perf_counter() - t0
            n1 = max(64, ns // (4 * cores))
            t0 = time.perf_counter()
            solve(poses[:n1], 1)
            dt1 = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"first {ns} pose pairs of the step's batch, {cores} std::threads on a dynamic chunk queue, {dt:.1f} s",
                                    "single_thread": {"value": n1 / dt1, "unit": UNIT, "sample": f"first {n1} pose pairs, 1 thread, {dt1:.1f} s"}}
            # verification: that sample plus the heaviest queries of the step (the ones at the reference's 150-iteration cap)
            heavy = np.argsort(-host_out["num_bv_tests"])[:args.verify_heavy]
            heavy = heavy[heavy >= ns]
            t0 = time.perf_counter()
            ref_h = solve(poses[heavy], cores) if len(heavy) else ref[:0]
            dth = time.perf_counter() - t0
            idx = np.concatenate([np.arange(ns), heavy])
            refc = np.concatenate([ref, ref_h])
            g = {k: host_out[k][idx] for k in host_out}
            line["verified"] = {
                "n": int(len(idx)), "n_heaviest": int(len(heavy)), "heaviest_cpu_s": round(dth, 1),
                "max_num_ca": int(g["num_ca"].max()), "at_iteration_cap": int((g["num_ca"] >= 151).sum()),
                "verdict_match": float(np.mean(g["collisionfree"] == refc["collisionfree"])),
                "toc_within_tol": float(np.mean(np.abs(g["toc"] - refc["toc"]) <= TOL)),
                "dist_within_1e-9_rel": float(np.mean(np.abs(g["distance"] - refc["distance"]) <= 1e-9 * np.maximum(1.0, np.abs(refc["distance"])))),
                "bit_exact_toc_dist_pose": float(np.mean((g["toc"] == refc["toc"]) & (g["distance"] == refc["distance"])
                                                         & (g["pose_toc"] == refc["pose_toc"]).all(1))),
                "num_ca_equal": float(np.mean(g["num_ca"] == refc["numCA"])),
                "counters_equal": float(np.mean((g["num_bv_tests"] == refc["num_bv_tests"]) & (g["num_tri_tests"] == refc["num_tri_tests"])))}
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="c2a_b200", choices=["c2a_b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("C2A_BENCH_BATCH", "1000000")),
                    help="pose pairs per step in total (config 3 names 1M)")
    ap.add_argument("--verify-heavy", type=int, default=1000, help="heaviest queries of the step added to the verified sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--ref-batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
