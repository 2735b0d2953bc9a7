import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine
B = int(os.environ.get("RB", "2"))
if os.environ.get("RPOISON", "1") == "1":
    # poison the memory the library's cudaMalloc calls will get: every byte 0xFF (NaN as fp32 / bf16)
    junk = torch.full((24 << 30,), 0xFF, dtype=torch.uint8, device="cuda")
    del junk
    torch.cuda.empty_cache()
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=B, max_new_tokens=32, policy=os.environ.get("RPOL", "fast"))
wave = synth.synthetic_waveforms(2 * B).cuda(); ids = synth.synthetic_prompt_ids(B).cuda()
eng.encode(wave[:B], wave[B:]); eng.prefix(ids); eng.prefill(B, want_logits=False)
toks, dump = eng.decode(B, 6, dump_logits=True)
torch.cuda.synchronize()
print("finite per step:", [bool(torch.isfinite(dump[i]).all()) for i in range(dump.shape[0])], "tokens", toks.cpu().tolist())
eng.close()
