import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import native
from mellow_b200.engine import Engine
eng = Engine(None, device=0, max_batch=1, max_new_tokens=8, policy=sys.argv[1] if len(sys.argv) > 1 else "split")
for (m, n, k) in [(300, 288, 96), (128, 960, 576), (5, 4096, 576), (77, 527, 4608)]:
    g = torch.Generator().manual_seed(1)
    a = torch.randn(m, k, generator=g); w = torch.randn(n, k, generator=g) / k ** 0.5
    try:
        c = eng.op_gemm(a, w)
        torch.cuda.synchronize()
        print((m, n, k), "err", (c.cpu().double() - a.double() @ w.double().T).abs().max().item(), flush=True)
    except Exception as e:
        print((m, n, k), "FAILED", e, flush=True)
        break
