"""ncu workload for the decode-attention kernel at the bench operating point (B=128, mean context 539)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine
B, ctx = 128, 539
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=B, max_new_tokens=300, policy=sys.argv[1] if len(sys.argv) > 1 else "split")
wave = synth.synthetic_waveforms(2 * B).cuda(); ids = synth.synthetic_prompt_ids(B).cuda()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    eng.encode(wave[:B], wave[B:]); eng.prefix(ids); eng.prefill(B, want_logits=False)
    eng.bench_decode_attention(B, ctx, 40)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.bench_decode_attention(B, ctx, 6)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
