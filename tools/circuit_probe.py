"""Developer probe (gpurun): circuit-amplitude tree, eager vs CUDA-graph replay vs numpy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, quimb_b200 as qb
from oracle import contract_np as cn
from tests.circuit_util import random_circuit_amplitude
for nq, depth in [(12, 8), (16, 12)]:
    arrays, inputs, output, amp = random_circuit_amplitude(nq, depth, 1) if nq <= 20 else (None,) * 4
    dev = [qb.asarray(a) for a in arrays]
    t0 = time.perf_counter(); tr = qb.find_tree(inputs, output, {ix: 2 for t in inputs for ix in t}, "greedy"); tf = time.perf_counter() - t0
    from quimb_b200.tree import execute
    execute(tr, dev); torch.cuda.synchronize()
    t0 = time.perf_counter(); 
    for _ in range(3): out = execute(tr, dev)
    torch.cuda.synchronize(); te = (time.perf_counter() - t0) / 3
    g = qb.GraphedContraction(inputs, output, arrays, optimize=tr)
    g(*dev); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): g.graph.replay()
    torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter(); ref = cn.array_contract(arrays, inputs, output, "greedy"); tn = time.perf_counter() - t0
    print(f"nq={nq} depth={depth} tensors={len(arrays)} width={tr.contraction_width():.0f} cost={tr.contraction_cost():.3g} | "
          f"find {tf*1e3:.1f} ms | eager {te*1e3:.2f} ms | graph {tg*1e3:.3f} ms | numpy(incl path) {tn*1e3:.1f} ms | err {abs(complex(out.item())-amp):.1e}", flush=True)
