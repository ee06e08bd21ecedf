#!/usr/bin/env python
"""Benchmark of the hot path: BASELINE.json configs[1]
"MPS norm/expectation contraction L=200 chi=1024 fp64 on 1xB200".

A *step* is one full contraction <psi|psi> of a synthetic random MPS
(L=200, bond 1024, d=2, fp64): 400 launches of the pairwise contraction
kernel, 1.7 TFLOP of algorithmic work (sum over sites of
2*l*l*d*r + 2*l*d*r*r).  `value` is device-resident throughput, `e2e` is the
same step through the public API starting from pinned HOST buffers (H2D of
all 200 site tensors inside the timed region, result read back).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the reference's own CPU implementation of the path
(the numpy/OpenBLAS restatement in oracle/, all host threads; the reference
is pure Python + numpy and cannot be installed on the GPU box).
N > 1: every rank contracts its own independent MPS (weak scaling, no data
path collective); timing is the max over ranks.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

L_SITES, CHI, PHYS = 200, 1024, 2
METRIC = "contracted-TFLOP/s at chi=1024 (MPS norm L=200, fp64)"


def bond_dims(L, chi, d):
    out = [1]
    for i in range(1, L):
        e = min(i, L - i)
        cap = d ** e if e < 40 else chi
        out.append(int(min(cap, chi)))
    out.append(1)
    return out


def step_flops(L, chi, d):
    b = bond_dims(L, chi, d)
    fl = 0
    for i in range(L):
        l, r = b[i], b[i + 1]
        fl += 2 * l * l * d * r + 2 * l * d * r * r
    return fl


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._th = None

    def start(self):
        def run():
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            while not self._stop.is_set():
                try:
                    out = subprocess.run(
                        ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                         "--format=csv,noheader,nounits"],
                        capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self._th = threading.Thread(target=run, daemon=True)
        self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_steps(steps, warmup, target_seconds=12.0):
    """The reference's CPU path (numpy/OpenBLAS tensordot chain) on a bounded
    sample of the same workload: `nsites` bulk chi=1024 sites per step."""
    from oracle import dmrg_np as dm
    # use every host thread the BLAS will take (torchrun exports
    # OMP_NUM_THREADS=1, which would otherwise pin the baseline to one core)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    rng = np.random.default_rng(0)
    A = rng.standard_normal((CHI, PHYS, CHI))
    A /= np.linalg.norm(A) ** 0.5
    per_site = 4 * PHYS * CHI ** 3
    # calibrate
    t0 = time.perf_counter()
    dm.mps_norm2([A[:1].copy()] + [A] * 2 + [A[:, :, :1].copy()])
    t1 = time.perf_counter() - t0
    est_site = max(t1 / 2.5, 1e-3)
    total = max(steps + warmup, 1)
    nsites = int(max(4, min(L_SITES, target_seconds / est_site / total)))
    sites = [A[:1].copy()] + [A] * (nsites - 2) + [A[:, :, :1].copy()]
    fl = sum(2 * s.shape[0] ** 2 * PHYS * s.shape[2] + 2 * s.shape[0] * PHYS * s.shape[2] ** 2
             for s in sites)
    times = []
    for i in range(total):
        t0 = time.perf_counter()
        dm.mps_norm2(sites)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tm = float(np.mean(times))
    try:
        import threadpoolctl
        nthreads = max((p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()),
                       default=os.cpu_count())
    except Exception:
        nthreads = os.cpu_count()
    return {"value": fl / tm / 1e12, "ms_per_step": tm * 1e3, "cores": int(nthreads),
            "sample": f"{nsites} chi={CHI} d={PHYS} sites of the L={L_SITES} chain per step "
                      f"({fl / 1e9:.1f} GFLOP), numpy tensordot on OpenBLAS",
            "nsites": nsites, "per_site_gflop": per_site / 1e9}


def run_reference(args, rank, world):
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "MPS norm L=200 chi=1024 d=2 fp64 (configs[1])",
                   "sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": "TFLOP/s", "cores": r["cores"],
                         "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import quimb_b200 as qb
    from quimb_b200 import _lib, mps as qmps

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    args.warmup = max(args.warmup, 3)
    bonds = bond_dims(L_SITES, CHI, PHYS)
    flops = step_flops(L_SITES, CHI, PHYS)

    # synthetic MPS, quimb layout (l, r, p); scaled like MPS_rand_state does
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    sites = []
    for i in range(L_SITES):
        x = torch.randn((bonds[i], bonds[i + 1], PHYS), dtype=torch.float64,
                        device=dev, generator=g)
        nd = sum(1 for s in x.shape if s > 1) or 1
        x /= torch.linalg.vector_norm(x) ** (1.5 / nd)
        sites.append(x)
    in_bytes = sum(s.numel() * 8 for s in sites)

    def one_step(record=None):
        n = len(sites)
        E = None
        for i, s in enumerate(sites):
            A = qmps.site_lpr(qb.Array(s), "lrp", i, n)
            if E is None:
                E = qb.ones((1, 1), dtype="float64", device=dev)
            if record is not None:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e2 = torch.cuda.Event(enable_timing=True)
                e0.record()
                T = qb.contract_pair(E.t, [3, 0], A.t, [0, 1, 2], [3, 1, 2])
                e1.record()
                E = qb.Array(qb.contract_pair(A.t, [3, 1, 5], T, [3, 1, 2], [5, 2]))
                e2.record()
                l, d, r = A.shape
                record.append((e0, e1, 2 * l * l * d * r, l, r))
                record.append((e1, e2, 2 * l * d * r * r, l, r))
            else:
                E = qmps.norm_step(E, A)
        return E

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------
    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    records = []
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = one_step(records)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1) / args.steps
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * flops / (ms_max * 1e-3) / 1e12
    norm2 = float(out.reshape(()).item())

    # dominant kernel: the full-size (chi x chi.d x chi) contraction launches
    big = [(a.elapsed_time(b), fl) for a, b, fl, l, r in records if l == CHI and r == CHI]
    k_ms = float(np.mean([t for t, _ in big]))
    k_fl = float(np.mean([fl for _, fl in big]))
    achieved = k_fl / (k_ms * 1e-3) / 1e12

    # ---- end to end: pinned host buffers -> public API -> host scalar ------
    e2e = None
    if not args.no_e2e:
        host = [s.cpu().pin_memory() for s in sites]
        copy_stream = torch.cuda.Stream(device=dev)

        def e2e_step():
            # double-buffered H2D prefetch on a side stream, contraction on
            # the current stream (this is what qb.mps_norm2 does for host input)
            return qb.mps_norm2(host, shape="lrp", copy_stream=copy_stream)

        for _ in range(2):
            e2e_step().item()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = e2e_step().item()     # D2H of the result inside the timed region
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert abs(res - norm2) <= 1e-9 * abs(norm2)
        # what the link alone allows: the same pinned buffers copied with no
        # compute (the e2e step cannot be faster than this)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(copy_stream):
            for h in host:
                h.to(dev, non_blocking=True)
        torch.cuda.synchronize()
        h2d_only = time.perf_counter() - t0
        e2e = {"value": world * flops / dt / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 8,
               "ms_per_step": dt * 1e3,
               "h2d_only_ms_per_step": h2d_only * 1e3,
               "h2d_only_GBps": in_bytes / h2d_only / 1e9}
        del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline denominators --------------------------------------------
    import ctypes
    tf = ctypes.c_double()
    _lib.load().qb_measure_dmma_peak(ctypes.byref(tf), None)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # DRAM traffic per launch of the dominant kernel from the committed ncu
    # --set full capture (profiles/): mean of dram read + write over its launches
    traffic = None
    try:
        rd = wr = None
        for ln in open(os.path.join(ROOT, "profiles", "r01_contract_kernel_ncu_full.txt")):
            if ln.startswith("dram__bytes_read.sum [Mbyte]"):
                rd = [float(x) * 1e6 for x in ln.split(":", 1)[1].split("|")]
            if ln.startswith("dram__bytes_write.sum [Kbyte]"):
                wr = [float(x) * 1e3 for x in ln.split(":", 1)[1].split("|")]
        if rd and wr:
            traffic = float(np.mean(rd) + np.mean(wr))
    except Exception:
        pass
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s",
        "frac": achieved / tf.value if tf.value else None, "traffic": traffic,
        "traffic_source": "profiles/r01_contract_kernel_ncu_full.txt (bytes per launch; "
                          "algorithmic 33.5e6)",
        "kernel": "contract_f64_streamk_kernel<128,128,16> (DMMA fp64, persistent stream-K)",
        "peak_source": "fp64 DMMA issue-rate microbenchmark measured live on this GPU "
                       "(tcgen05 has no f64 kind; MEASURED_PEAKS.json holds bf16 only)",
        "flops_per_launch": k_fl, "ms_per_launch": k_ms,
        "bf16_peak_measured": peaks.get("bf16_tflops"),
        "frac_of_bf16_measured": (achieved / peaks["bf16_tflops"]) if peaks.get("bf16_tflops") else None,
    }
    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference_steps(2, 1, target_seconds=12.0)
        cpu = {"value": r["value"], "unit": "TFLOP/s", "cores": r["cores"], "kind": "port",
               "sample": r["sample"]}

    line = {
        "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "MPS norm <psi|psi>, L=200 chi=1024 d=2 fp64 (BASELINE configs[1])",
                   "flops_per_step": flops, "input_bytes": in_bytes,
                   "l2": "inputs (3.3 GB) are larger than L2; no flush needed",
                   "parallelism": "1 GPU" if world == 1 else
                                  f"{world} independent MPS replicas, one per rank (no data-path collective)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "check": {"norm2": norm2},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
