"""Single-precision / complex64 contractions: the native tcgen05 engine (4 int8
slices, 128x128 tiles) against the widening route it replaced (convert both
operands to double, DMMA kernel, convert back), B200."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quimb_b200 as qb
from quimb_b200.contract import contract_pair, convert

g = torch.Generator(device="cuda").manual_seed(0)


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


rows = []
for dt, (m, n, k) in [(torch.float32, (4096, 4096, 4096)), (torch.complex64, (2048, 2048, 2048)),
                      (torch.complex64, (4096, 2048, 512)), (torch.float32, (1024, 1024, 1024)),
                      (torch.complex64, (16384, 256, 2048))]:
    rdt = torch.float32
    if dt.is_complex:
        a = torch.view_as_complex(torch.randn(m, k, 2, dtype=rdt, device="cuda", generator=g))
        b = torch.view_as_complex(torch.randn(k, n, 2, dtype=rdt, device="cuda", generator=g))
    else:
        a = torch.randn(m, k, dtype=rdt, device="cuda", generator=g)
        b = torch.randn(k, n, dtype=rdt, device="cuda", generator=g)
    wide = torch.complex128 if dt.is_complex else torch.float64
    native = t_ms(lambda: contract_pair(a, [0, 1], b, [1, 2], [0, 2]))
    widened = t_ms(lambda: convert(contract_pair(convert(a, wide), [0, 1], convert(b, wide), [1, 2],
                                                 [0, 2]), dt))
    fl = (8.0 if dt.is_complex else 2.0) * m * n * k
    out = contract_pair(a, [0, 1], b, [1, 2], [0, 2])
    ref = (a.to(wide) @ b.to(wide))
    err = float((out.to(wide) - ref).abs().max() / ref.abs().max())
    row = {"dtype": str(dt), "mnk": [m, n, k], "native_ms": native, "native_tflops": fl / native / 1e9,
           "widened_ms": widened, "widened_tflops": fl / widened / 1e9, "rel_err_vs_exact": err}
    rows.append(row)
    print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/oz_single_prof.json", "w"), indent=1)
