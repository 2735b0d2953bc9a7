"""tcgen05 GEMM engine bring-up: numerics vs fp64 and timing vs the mma.sync engine."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200.engine import Engine
policy = sys.argv[1] if len(sys.argv) > 1 else "split"
eng = Engine(None, device=0, max_batch=1, max_new_tokens=8, policy=policy)
shapes = [(256, 128, 64), (256, 128, 128), (300, 288, 96), (1000, 576, 576), (4096, 3072, 576), (777, 527, 4608),
          (2048, 96, 384), (512, 960, 576), (4096, 576, 1536), (640, 768, 544)]
for (m, n, k) in shapes:
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g); w = torch.randn(n, k, generator=g) / k ** 0.5; b = torch.randn(n, generator=g)
    want = a.double() @ w.double().T + b.double()
    res = {}
    for e in (0, 1):
        eng.set_gemm_engine(e)
        try:
            c = eng.op_gemm(a, w, b); torch.cuda.synchronize()
            res[e] = (c.cpu().double() - want).abs().max().item()
        except Exception as ex:
            res[e] = f"FAILED {ex}"
            print((m, n, k), res, flush=True)
            sys.exit(1)
    print((m, n, k), "err mma %.3g  umma %.3g" % (res[0], res[1]), flush=True)
print("numerics done", flush=True)
