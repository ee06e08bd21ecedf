"""Phase timing of the contraction kernel from its in-kernel %globaltimer stamps
(QB_TRACE=1; GPU box only).  Runs the two launches of one MPS-norm site
(chi=1024, d=2) back to back and prints, per launch, the median over CTAs of
each phase and the gap between consecutive kernels.

phases: 0 entry, 1 tables done, 2 prologue issued, 3 first tile landed,
        4 main loop done, 5 stream-K fix-up done, 6 stores issued
"""
import ctypes
import json
import os
import sys

os.environ.setdefault("QB_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import quimb_b200 as qb
from quimb_b200 import _lib, mps


def read_trace():
    n = 16 * 8192
    buf = (ctypes.c_ulonglong * n)()
    rc = _lib.load().qb_debug_trace_read(buf, n)
    assert rc == 0, rc
    return np.frombuffer(buf, dtype=np.uint64).reshape(16, 8192).copy()


def analyse(slot, G=148):
    t = slot[: G * 32].reshape(G, 4, 8).astype(np.int64)
    out = {}
    seg_used = (t[:, :, 0] > 0)
    t0 = t[:, 0, 0][seg_used[:, 0]]
    out["entry_spread_us"] = float((t0.max() - t0.min()) / 1e3)
    ends = np.where(seg_used, t[:, :, 6], 0).max(axis=1)
    out["start_ns"] = int(t0.min()); out["end_ns"] = int(ends.max())
    out["total_us"] = (out["end_ns"] - out["start_ns"]) / 1e3
    names = ["tables", "prologue_issue", "first_tile_wait", "main_loop", "fixup", "stores"]
    for s in range(4):
        m = seg_used[:, s]
        if not m.any():
            continue
        d = {}
        for ph in range(6):
            a, b = t[m, s, ph], t[m, s, ph + 1]
            ok = (a > 0) & (b > 0)
            if ok.any():
                d[names[ph]] = round(float(np.median((b - a)[ok])) / 1e3, 2)
        d["n_cta"] = int(m.sum())
        out[f"seg{s}"] = d
    return out


def main():
    dev = torch.device("cuda:0")
    chi, d = 1024, 2
    g = torch.Generator(device=dev); g.manual_seed(0)
    A = qb.asarray(torch.randn((chi, d, chi), dtype=torch.float64, device=dev, generator=g))
    E = qb.asarray(torch.randn((chi, chi), dtype=torch.float64, device=dev, generator=g))
    for _ in range(3):
        E2 = mps.norm_step(E, A)
    torch.cuda.synchronize()
    read_trace()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        E2 = mps.norm_step(E, A)
    e1.record(); torch.cuda.synchronize()
    print("4 sites:", e0.elapsed_time(e1) * 1e3, "us")
    tr = read_trace()
    used = [i for i in range(16) if tr[i].any()]
    res = []
    for i in used:
        res.append(analyse(tr[i]))
    res.sort(key=lambda r: r["start_ns"])
    for k, r in enumerate(res):
        gap = (r["start_ns"] - res[k - 1]["end_ns"]) / 1e3 if k else None
        print(json.dumps({"launch": k, "gap_before_us": gap, **{a: b for a, b in r.items()
                                                               if a not in ("start_ns", "end_ns")}}))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/trace_contract.json", "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
