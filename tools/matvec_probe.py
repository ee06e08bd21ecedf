"""Developer probe (gpurun, 1 GPU): the three contraction steps of the DMRG
two-site matvec at chi = 1024 (L.x, .W12, .R), each under the engines that can
take it -- decides the engine defaults from measurements."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quimb_b200 as qb
from quimb_b200.contract import contract_pair

chi, d, w = 1024, 2, 5
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, dtype=torch.float64, device="cuda", generator=g)
L, R, W12, x = rn(chi, w, chi), rn(chi, w, chi), rn(w, d, d, w, d, d), rn(chi, d, d, chi)
LB, W_, L_, S_, T_, R_, W2_, SB_, TB_, RB_ = range(10)


def t_ms(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
T1 = contract_pair(L, [LB, W_, L_], x, [L_, S_, T_, R_], [LB, W_, S_, T_, R_])
T3 = contract_pair(T1, [LB, W_, S_, T_, R_], W12, [W_, S_, T_, W2_, SB_, TB_], [LB, SB_, TB_, W2_, R_])
for eng, name in ((0, "auto"), (1, "dmma"), (2, "ozaki"), (3, "stream")):
    row = {}
    try:
        row["Lx_ms"] = t_ms(lambda: contract_pair(L, [LB, W_, L_], x, [L_, S_, T_, R_],
                                                  [LB, W_, S_, T_, R_], engine=eng))
        row["W12_ms"] = t_ms(lambda: contract_pair(T1, [LB, W_, S_, T_, R_], W12,
                                                   [W_, S_, T_, W2_, SB_, TB_],
                                                   [LB, SB_, TB_, W2_, R_], engine=eng))
        row["R_ms"] = t_ms(lambda: contract_pair(T3, [LB, SB_, TB_, W2_, R_], R, [RB_, W2_, R_],
                                                 [LB, SB_, TB_, RB_], engine=eng))
    except Exception as e:  # noqa: BLE001
        row["error"] = str(e)[:200]
    out[name] = row
    print(name, json.dumps(row), flush=True)
fl = 2.0 * chi * w * chi * d * d * chi
out["gemm_gflop"] = fl / 1e9
# plain big GEMMs for the engine threshold
for (m, n, k) in [(4096, 4096, 4096), (2048, 2048, 2048), (1024, 2048, 1024), (2048, 2048, 1024)]:
    a, b = rn(m, k), rn(k, n)
    row = {}
    for eng, name in ((1, "dmma"), (2, "ozaki")):
        ms = t_ms(lambda: contract_pair(a, [0, 1], b, [1, 2], [0, 2], engine=eng), 3)
        row[name] = {"ms": ms, "tflops": 2.0 * m * n * k / ms / 1e9}
    out[f"gemm_{m}x{n}x{k}"] = row
    print(m, n, k, json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/matvec_probe.json", "w"), indent=1)
