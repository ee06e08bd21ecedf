"""Host-only: tree + slices for BASELINE configs[3] (6x6 qubits, depth 24,
complex128 amplitude) with the built-in finders; writes the numbers the
multi-GPU run is planned against.  python tools/tree_cfg4.py [Lx Ly depth] [widths...]"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.circuit_util import random_grid_circuit_amplitude
from quimb_b200 import tree, treeopt

Lx, Ly, depth = (int(v) for v in (sys.argv[1:4] or (6, 6, 24)))
widths = [int(v) for v in sys.argv[4:]] or [32, 30]
arrays, inputs, output, _ = random_grid_circuit_amplitude(Lx, Ly, depth, seed=3, dense=False)
sd = {ix: 2 for t in inputs for ix in t}
res = {"config": f"{Lx}x{Ly} depth {depth} fSim grid circuit, all-zeros bitstring, seed 3",
       "tensors": len(inputs), "finders": {}, "sliced": {}}
for opt in ("greedy", "random-greedy", "spectral", "auto-hq"):
    t0 = time.time()
    tr = tree.find_tree(inputs, output, sd, opt)
    res["finders"][opt] = {"log2_macs": round(math.log2(tr.contraction_cost()), 2),
                           "log2_width": tr.contraction_width(), "seconds": round(time.time() - t0, 1)}
    print(opt, res["finders"][opt], flush=True)
for w in widths:
    for name, mini in (("flops", "flops"), ("combo16", tree.DEVICE_COMBO)):
        t0 = time.time()
        tr, sl = tree.find_sliced_tree(inputs, output, sd, w, minimize=mini)
        ssa = [(i, j) for i, j, _, _ in tr.steps]
        macs = math.log2(tr.contraction_cost()) + len(sl)
        elems = treeopt.tree_traffic(tr.inputs, tr.output, sd, ssa) + len(sl)
        key = f"{w}:{name}"
        res["sliced"][key] = {
            "n_sliced": len(sl), "log2_macs_total": round(macs, 2),
            "log2_elements_moved_total": round(elems, 2), "log2_width": tr.contraction_width(),
            # complex128 on 8 B200: 8 real flop per multiply-add at 25 TFLOP/s (measured DMMA
            # rate on the bench shape), 16 B per element at 7 TB/s
            "model_seconds_8gpu": {"compute": round(8 * 2 ** macs / (8 * 25e12), 1),
                                   "hbm": round(16 * 2 ** elems / (8 * 7e12), 1)},
            "search_seconds": round(time.time() - t0, 1)}
        print(key, res["sliced"][key], flush=True)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_tree_cfg4.json")
if (Lx, Ly, depth) == (6, 6, 24):
    json.dump(res, open(out, "w"), indent=1)
