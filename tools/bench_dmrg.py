"""DMRG2 sweep benchmark (BASELINE configs[2]: Heisenberg chain, chi=1024,
fp64, two-site SVD).  Not the driver's bench.py contract (that is configs[1]);
this produces the 'DMRG sweep time' half of the BASELINE metric.

  python tools/bench_dmrg.py [--L 100] [--chi 1024] [--cpu-L 24] [--no-cpu]

GPU: one full `sweep_right(canonize=True, max_bond=chi, cutoff=0)` of
quimb_b200.DMRG2 on a random MPS (device Lanczos ncv=4 tol=1e-3, Jacobi SVD).
CPU: the numpy/scipy oracle (the reference's algorithm: ARPACK + LAPACK) on a
shorter chain that still reaches chi in the middle; both are reported per
two-site update at full chi so they can be compared like for like.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=100)
    ap.add_argument("--chi", type=int, default=1024)
    ap.add_argument("--cpu-L", type=int, default=24)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--method", default="svd",
                    help="bond_compress_method: svd | svd:eig | svd:rand")
    ap.add_argument("--ncv", type=int, default=None, help="device_eig_ncv override")
    ap.add_argument("--left-sweep", action="store_true",
                    help="also time the following sweep_left(canonize=False)")
    ap.add_argument("--shard", action="store_true",
                    help="under torchrun: row-shard the local eigensolves over the ranks (NCCL)")
    args = ap.parse_args()
    import torch
    import quimb_b200 as qb
    from oracle import dmrg_np as dm
    shard = None
    rank = 0
    if args.shard:
        import torch.distributed as dist
        from quimb_b200.dist import BondShard
        rank = int(os.environ["RANK"])
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl")
        shard = BondShard()
        args.no_cpu = True

    out = {"L": args.L, "chi": args.chi, "dtype": "f64"}
    mpo = qb.mpo_ham_heis(args.L)        # same arrays as the oracle's mpo_heis
    d = qb.DMRG2(mpo, args.chi, cutoffs=0.0, mpo_shape="lrdu", seed=2, shard=shard)
    if args.ncv:
        d.opts["device_eig_ncv"] = args.ncv
    out["opts"] = {k: d.opts[k] for k in ("device_eig_ncv", "device_eig_min_steps",
                                            "local_eig_tol")}
    out["method"] = args.method
    if shard is not None:
        out["shard"] = {"world_size": shard.world_size, "backend": "nccl"}
    torch.cuda.synchronize()
    site_t = []
    orig = d._update_local_state_2site

    def timed(i, direction, **kw):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = orig(i, direction, **kw)
        torch.cuda.synchronize(); site_t.append((i, time.perf_counter() - t0, d.nmatvecs[-1],
                                                 d._k[i].shape[2]))
        return r
    d._update_local_state_2site = timed
    n0 = qb.launch_count()
    t0 = time.perf_counter()
    e = d.sweep_right(canonize=True, max_bond=args.chi, cutoff=0.0, cutoff_mode="sum2",
                      method=args.method)
    torch.cuda.synchronize()
    t_sweep = time.perf_counter() - t0
    if args.left_sweep:
        n_first = len(site_t)
        t1 = time.perf_counter()
        e2 = d.sweep_left(canonize=False, max_bond=args.chi, cutoff=0.0, cutoff_mode="sum2",
                          method=args.method)
        torch.cuda.synchronize()
        second = site_t[n_first:]
        out["gpu_left_sweep"] = {
            "sweep_s": time.perf_counter() - t1, "energy": e2,
            "matvecs_per_site": float(np.mean([nmv for (_, _, nmv, _) in second])),
            "s_per_site_full_chi": float(np.median([t for (_, t, _, k) in second
                                                    if k == args.chi] or [0.0])),
        }
        site_t = site_t[:n_first]
    full = [t for (i, t, nmv, k) in site_t if k == args.chi]
    out["gpu"] = {
        "sweep_s": t_sweep, "energy": e, "launches": qb.launch_count() - n0,
        "sites": len(site_t), "sites_at_full_chi": len(full),
        "s_per_site_full_chi": float(np.median(full)) if full else None,
        "matvecs_per_site": float(np.mean([nmv for (_, _, nmv, _) in site_t])),
        "update_s_total": float(sum(t for (_, t, _, _) in site_t)),
    }
    if shard is not None:
        out["shard"]["bytes_gathered"] = shard.bytes_gathered
        out["shard"]["exchange"] = shard.exchange_name
    if rank == 0:
        print(json.dumps(out), flush=True)
    if not args.no_cpu:
        Lc = args.cpu_L
        o = dm.DMRG2(dm.mpo_heis(Lc), args.chi, cutoffs=0.0, seed=2)
        ts = []
        orig_o = o._update_2site

        def timed_o(i, direction, max_bond, cutoff):
            t0 = time.perf_counter()
            r = orig_o(i, direction, max_bond, cutoff)
            ts.append((i, time.perf_counter() - t0, o.nmatvecs[-1], o.k[i].shape[2]))
            return r
        o._update_2site = timed_o
        t0 = time.perf_counter()
        eo = o.sweep("R", canonize=True, max_bond=args.chi, cutoff=0.0)
        t_cpu = time.perf_counter() - t0
        fullc = [t for (i, t, nmv, k) in ts if k == args.chi]
        out["cpu"] = {
            "L": Lc, "sweep_s": t_cpu, "energy": eo, "cores": os.cpu_count(),
            "sites_at_full_chi": len(fullc),
            "s_per_site_full_chi": float(np.median(fullc)) if fullc else None,
            "matvecs_per_site": float(np.mean([nmv for (_, _, nmv, _) in ts])),
            "kind": "port (numpy tensordot + scipy ARPACK + LAPACK gesdd)",
        }
        if full and fullc:
            out["speedup_per_site_full_chi"] = out["cpu"]["s_per_site_full_chi"] / out["gpu"]["s_per_site_full_chi"]
            out["cpu_sweep_extrapolated_s"] = out["cpu"]["s_per_site_full_chi"] * len(site_t)
    if rank == 0:
        if not args.no_cpu:
            print(json.dumps(out), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        name = "bench_dmrg_shard%d.json" % shard.world_size if shard else "bench_dmrg.json"
        json.dump(out, open(os.path.join("gpurun_out", name), "w"), indent=1)
    if shard is not None:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
