#!/usr/bin/env python
"""Headline benchmark: batched C2A_Solve (controlled conservative-advancement CCD) queries/s.

Workload (BASELINE.json configs[2]): torus knot (512x32 quads, 32768
triangles) against itself, synthetic interpolated-motion pose pairs, FP64, tolerance_d = tolerance_t
= 1e-4 as C2A_Solve hard-codes them.  A step = one batch of --batch pose pairs per GPU through the hot
path.  Weak scaling: every rank (one per GPU) solves its own batch; no data-path collective.

  python bench.py --gpus N --steps K --warmup W          this repo's CUDA path
  python bench.py --impl reference ...                   the reference's own CPU code on the host cores

One JSON line on stdout (rank 0).
  value     whole-job queries/s, motion records already resident in HBM, CUDA-event timed per step
  e2e       the same through c2a_b200_solve_batch with HOST buffers: host motion set-up, H2D, kernel, D2H
  roofline  c2a_solve_kernel: algorithmic bytes (208*Nbv + 144*Ntri + 448 per query, SURVEY.md 8d, from the
            kernel's own counters) / measured launch time, against the measured HBM copy bandwidth
  cpu_baseline  oracle/_ref (the reference's object code) or the port, all host cores, bounded sample
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
            f"{batch} synthetic interpolated-motion pose pairs per step per GPU (BASELINE configs[2])")


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
    """--impl reference: the reference's CPU implementation of the path on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from c2a_b200 import workloads
    cores = os.cpu_count() or 1
    solve, kind = cpu_solver()
    n = args.ref_batch if args.ref_batch > 0 else 768 * cores
    times = []
    for it in range(args.warmup + args.steps):
        poses = workloads.approach_batch(n, SEED + 1000 + it, radius=workloads.KNOT_RADIUS)
        t0 = time.perf_counter()
        solve(poses, cores)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = args.steps * n / total
    sample = f"{n} pose pairs per step (a bounded sample of the {args.batch}-pair GPU step), {cores} std::threads"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.batch), "tolerance_d": TOL, "tolerance_t": TOL,
                       "reference_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    from c2a_b200 import api, build as c2a_build, meshes, workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if rank == 0:
        c2a_build.build()
    if dist:
        dist.barrier()

    B = args.batch
    bvh = api.build_bvh(meshes.torus_knot(*KNOT)[0])
    model = api.Model(bvh, local)
    poses = workloads.approach_batch(B, SEED + rank, radius=workloads.KNOT_RADIUS)
    fields = ("status", "collisionfree", "num_ca", "num_bv_tests", "num_tri_tests", "toc", "distance", "pose_toc")

    # ---- device-resident leg ("value"): motion records in HBM, CUDA events on the launching stream ----
    stream = torch.cuda.Stream()
    motions_host = api.motions_from_poses(poses)
    motions = torch.from_numpy(motions_host).pin_memory().cuda(non_blocking=True)
    order = torch.from_numpy(api.schedule_order(model, model, motions_host)).cuda()  # scheduling hint, resident like the inputs
    out = {"status": torch.empty(B, dtype=torch.int32, device="cuda"), "collisionfree": torch.empty(B, dtype=torch.int32, device="cuda"),
           "num_ca": torch.empty(B, dtype=torch.int32, device="cuda"), "num_bv_tests": torch.empty(B, dtype=torch.int32, device="cuda"),
           "num_tri_tests": torch.empty(B, dtype=torch.int32, device="cuda"), "toc": torch.empty(B, dtype=torch.float64, device="cuda"),
           "distance": torch.empty(B, dtype=torch.float64, device="cuda"), "pose_toc": torch.zeros(B, 24, dtype=torch.float64, device="cuda")}
    ptrs = {k: v.data_ptr() for k, v in out.items()}
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    torch.cuda.synchronize()

    def step():
        api.solve_batch_device(model, model, motions.data_ptr(), B, ptrs, tol_d=TOL, tol_t=TOL, stream=stream.cuda_stream,
                               order_ptr=order.data_ptr())

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
    for k in range(args.steps):
        scrub.fill_(k)  # flush L2 between timed iterations (on torch's stream, not timed)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            ev[k][0].record(stream)
            step()
            ev[k][1].record(stream)
    stream.synchronize()
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

    # ---- end-to-end leg: the public host-buffer call, every step H2D of inputs + D2H of results ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host_out = api.solve_batch(model, model, poses, tol_d=TOL, tol_t=TOL, fields=fields)  # warm (allocator, pinned pool)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_out = api.solve_batch(model, model, poses, tol_d=TOL, tol_t=TOL, fields=fields)
    e2e_s = time.perf_counter() - t0
    import ctypes as C
    tb = (C.c_double * 8)()
    api.lib().c2a_b200_host_timing(tb)  # wall-clock breakdown of the last e2e call
    e2e_breakdown = {k: round(v, 4) for k, v in zip(("alloc_s", "motion_setup_s", "claim_order_s", "enqueue_s", "kernel_wait_s",
                                                      "d2h_s", "release_s", "total_s"), tb)}
    h2d = B * 48 * 8
    d2h = sum(host_out[k].nbytes for k in fields)
    for k in ("collisionfree", "toc", "distance", "num_ca"):
        assert np.array_equal(host_out[k], out[k].cpu().numpy()), k

    # ---- max over ranks ----
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max = float(t[0]), float(t[1])
    value = world * args.steps * B / (dev_ms_max * 1e-3)
    e2e_value = world * e2e_steps * B / e2e_s_max

    if rank == 0:
        peak, peak_src = measured_peaks()
        per_launch_s = dev_ms * 1e-3 / args.steps
        ach = algorithmic_bytes(nbv, ntri, B) / per_launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        f1, f2 = C.c_double(0), C.c_double(0)
        api.lib().c2a_b200_fp64_peak(C.byref(f1), C.byref(f2))
        fl = nominal_flops(nbv, ntri, nca) / per_launch_s / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(B), "tolerance_d": TOL, "tolerance_t": TOL, "batch_per_gpu": B,
                           "l2": "inputs per step (%.0f MB motion records) exceed the 126 MB L2 and a 256 MB scrub buffer is "
                                 "written between timed steps" % (B * 384 / 1e6),
                           "parallelism": f"{world} independent shards, no collective"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "breakdown_last_call": e2e_breakdown},
                "gpu_launches": launches,
                "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                             "kernel": "c2a_solve_kernel", "peak_source": peak_src,
                             "note": "BVH working set (16.5 MB) is L2-resident by design: DRAM traffic << algorithmic bytes; "
                                     "the binding resource is the FP64 pipe (see fp64)"},
                "fp64": {"achieved_tflops_nominal": fl, "peak_tflops_fma": f1.value, "peak_tflops_mul_add": f2.value,
                         "frac_of_mul_add_peak": fl / f2.value if f2.value else None},
                "bvtt_pairs_per_sec": world * args.steps * nbv / (dev_ms_max * 1e-3),
                "per_query": {"num_bv_tests": nbv / B, "num_tri_tests": ntri / B, "num_ca": nca / B, "hit_fraction": hits / B}}
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            solve, kind = cpu_solver()
            ns = min(B, args.cpu_sample if args.cpu_sample > 0 else 768 * cores)
            t0 = time.perf_counter()
            ref = solve(poses[:ns], cores)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"first {ns} pose pairs of the step's batch, {cores} std::threads, {dt:.1f} s"}
            g = {k: host_out[k][:ns] for k in host_out}
            line["verified"] = {
                "n": ns,
                "verdict_match": float(np.mean(g["collisionfree"] == ref["collisionfree"])),
                "toc_within_tol": float(np.mean(np.abs(g["toc"] - ref["toc"]) <= TOL)),
                "dist_within_1e-9_rel": float(np.mean(np.abs(g["distance"] - ref["distance"]) <= 1e-9 * np.maximum(1.0, np.abs(ref["distance"])))),
                "bit_exact_toc_dist_pose": float(np.mean((g["toc"] == ref["toc"]) & (g["distance"] == ref["distance"])
                                                         & (g["pose_toc"] == ref["pose_toc"]).all(1))),
                "num_ca_equal": float(np.mean(g["num_ca"] == ref["numCA"]))}
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
                    help="pose pairs per step per GPU (config 3 names 1M)")
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
