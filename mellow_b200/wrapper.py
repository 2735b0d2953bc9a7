"""Drop-in for the reference's ``mellow.MellowWrapper`` (mellow/wrapper.py:25-287) on B200.

Same constructor and ``generate()`` signature, same config names (``config='v0'``, ``model in {'v0','v0_s'}``), same
host-side audio/text preparation rules, same stopping rule and detokenisation; everything between the prepared
tensors and the token ids runs in ``libmellow_b200.so``.  The inner seams the reference's own methods use are kept
attribute-compatible: ``self.model.generate_prefix_inference(d)`` (mellow.py:100-108),
``self.model.caption_decoder.lm(inputs_embeds=x).logits[:, -1]`` and ``...lm.model.embed_tokens(ids)``
(wrapper.py:217,237), ``self._generate_batch(embed=...)`` (wrapper.py:197).

Differences a caller can observe, all deliberate:
  * no hub access is attempted when a local checkpoint / tokenizer is supplied (``checkpoint=``, ``tokenizer=`` or
    ``$MELLOW_CHECKPOINT`` / ``$MELLOW_TOKENIZER``); ``checkpoint='synthetic'`` builds the seeded synthetic
    checkpoint (there are no real weights offline) and defaults to the stand-in tokenizer;
  * ``.model`` is an ``Engine`` (native handle), not an ``nn.Module``; ``lm(...).logits`` holds the LAST position only
    (shape (B,1,V)): the reference materialises all S positions (9.8 GB at B=128) and reads ``[:, -1]``;
  * a CUDA device is mandatory -- ``use_cuda=False`` / ``device='cpu'`` raises instead of silently running on CPU;
  * the engine is sized to the request (up to 128 pairs per pass, the rest in further passes) and re-created when a
    call needs more rows or more new tokens than the current one holds;
  * under ``torchrun`` (``torch.distributed`` initialised, world > 1) ``generate()`` shards the examples over the
    ranks -- contiguous slices, no data-path collective -- and every rank returns the full list;
  * the reference's ``tokens.squeeze()`` collapse for batch > 1 with exactly one generated step (wrapper.py:251-253)
    is not reproduced.
"""
import argparse
import os
import random
from pathlib import Path

import torch
import yaml

from . import native, schema as S, weights
from .audio_io import load_audio_into_tensor, plan_fit, read_audio, resampled_length
from .engine import Engine, MellowNativeError
from .tokenizer import STAND_IN, load_tokenizer, tokenize_prompts

MAX_PASS = 128          # pairs per engine pass: one 128-row tile of the decode GEMMs
EOT = "<|endoftext|>"


class MellowWrapper:
    """A class for interfacing the Mellow model (B200-native engine)."""
    model_repo = "soham97/mellow"
    model_name = {"v0": "v0.ckpt", "v0_s": "v0_s.ckpt"}

    def __init__(self, config, model, device, use_cuda=True, *, checkpoint=None, tokenizer=None, policy="split24",
                 max_batch=None, max_new_tokens=None, shard=None):
        self.supported_versions = self.model_name.keys()
        if model not in self.supported_versions:                                       # wrapper.py:37-39
            raise ValueError(f"The model {model} is not supported. The supported versions are {str(self.supported_versions)}")
        if not use_cuda or device == "cpu" or not torch.cuda.is_available():
            raise MellowNativeError("mellow_b200 has no CPU path: a CUDA device (B200, sm_100a) is required")
        self.parent_path = Path(os.path.realpath(__file__)).parent
        self.config_path = os.path.join(self.parent_path, "config", config + ".yaml")
        self.use_cuda = use_cuda
        self.device = device
        self.policy, self.shard = policy, shard
        self._fixed_batch, self._fixed_new = max_batch, max_new_tokens
        self.model_path = self._resolve_checkpoint(model, checkpoint)
        self.model, self.tokenizer, self.args = self.get_model_and_tokenizer(self.config_path, tokenizer)

    # ------------------------------------------------------------------ construction
    def _resolve_checkpoint(self, model, checkpoint):
        checkpoint = checkpoint or os.environ.get("MELLOW_CHECKPOINT")
        if checkpoint:
            return checkpoint
        from huggingface_hub.file_download import hf_hub_download                      # wrapper.py:41
        return hf_hub_download(self.model_repo, self.model_name[model])

    def read_config_as_args(self, config_path):
        with open(config_path, "r") as f:
            yml_config = yaml.load(f, Loader=yaml.FullLoader)
        return argparse.Namespace(**dict(yml_config.items()))

    def get_model_and_tokenizer(self, config_path, tokenizer=None):
        args = self.read_config_as_args(config_path)
        args.model["decoder"]["prefix_dim"] = args.model["encoder"]["d_proj"]
        if args.model["model_type"] != "Mellow":
            raise NotImplementedError                                                  # model/model.py:7
        if args.model["encoder"]["audioenc_name"] != "HTSAT":
            raise Exception("The audio encoder name {} is incorrect or not supported".format(
                args.model["encoder"]["audioenc_name"]))                               # model/audio.py:7
        if "smollm2" not in args.model["decoder"]["text_decoder"].lower():
            raise ValueError(f"text decoder {args.model['decoder']['text_decoder']} not supported")
        if self.model_path == "synthetic":
            from .synth import synthetic_state_dict
            state = synthetic_state_dict()
            tokenizer = tokenizer or os.environ.get("MELLOW_TOKENIZER") or STAND_IN
        else:
            state = torch.load(self.model_path, map_location=torch.device("cpu"))      # wrapper.py:74
        tok = load_tokenizer(args.model["decoder"]["text_decoder"], tokenizer)         # raises when unavailable
        lib = native.load()
        dev = torch.device("cuda", int(self.device))
        self._arena = weights.pack(state, native.weight_entries(lib), lib.mb_weights_size()).to(dev)
        engine = self._new_engine(self._fixed_batch or 8, self._fixed_new or 300)
        params = S.count_parameters()
        print(f"model {os.path.basename(str(self.model_path))}, {os.path.basename(config_path)}, parameter count: {params}")
        return engine, tok, args

    def _new_engine(self, max_batch, max_new_tokens):
        return Engine(None, device=int(self.device), max_batch=max_batch, max_new_tokens=max_new_tokens,
                      policy=self.policy, arena=self._arena)

    def _ensure_engine(self, batch, max_len):
        """Size the native handle to the request: `batch` rows per pass (<= 128), `max_len` new tokens."""
        eng = self.model
        if self._fixed_batch:
            batch = min(batch, self._fixed_batch)
        if eng.max_batch >= batch and eng.max_new_tokens >= max_len:
            return eng
        if self._fixed_new and max_len > self._fixed_new:
            raise ValueError(f"max_len {max_len} exceeds max_new_tokens={self._fixed_new} given to the constructor")
        if S.PREFIX_LEN + max_len > S.MAX_POSITIONS:
            raise ValueError(f"max_len {max_len}: 389 + max_len exceeds SmolLM2's {S.MAX_POSITIONS} positions")
        new_batch, new_len = max(batch, eng.max_batch), max(max_len, eng.max_new_tokens)
        eng.close()
        self.model = self._new_engine(new_batch, new_len)
        return self.model

    # ------------------------------------------------------------------ host-side preparation
    def load_audio_into_tensor(self, audio_path, audio_duration, resample=True):
        return load_audio_into_tensor(audio_path, audio_duration, self.args.data["sampling_rate"], resample, random)

    def preprocess_audio(self, audio_files, resample, keep=None):
        """-> (n, 320000) float32 DEVICE tensor for the files in `keep` (a range; default all).  Files are decoded on
        the host; resampling to 32 kHz, channel flattening and tile-or-crop run on the GPU (the reference does all of it
        serially on the host and uploads one clip at a time, wrapper.py:170-179).  The `random.randrange` draws are
        made for EVERY file in list order like the reference, whether or not this rank keeps the clip, so a sharded
        run crops exactly where the single-process run does."""
        keep = range(len(audio_files)) if keep is None else keep
        target_sr, clip = self.args.data["sampling_rate"], S.CLIP_SAMPLES
        out = torch.empty(len(keep), clip, dtype=torch.float32, device=self.model.device)
        for i, audio_file in enumerate(audio_files):
            if i in keep:
                pcm, sr = read_audio(audio_file)
                self.model.prepare_clip(pcm, sr, target_sr, resample, random, out=out[i - keep.start])
            else:                                                   # another rank's clip: only its random draw
                ch, n_in, sr = read_audio(audio_file, info_only=True)
                total = ch * (resampled_length(n_in, sr, target_sr) if resample and sr != target_sr else n_in)
                plan_fit(total, clip, random)
        return out

    def preprocess_text(self, prompts):
        ids = tokenize_prompts(self.tokenizer, prompts, self.args.data["text_tokenization_len"])
        return {"input_ids": ids}

    # ------------------------------------------------------------------ generation
    def _detokenize(self, tokens):
        """wrapper.py:251-254: decode every row and keep the text before the first '<|endoftext|>'."""
        out = []
        for row in tokens.tolist():
            text = self.tokenizer.decode(row)
            out.append(text.split(EOT)[0])
        return out

    @staticmethod
    def _global_stop(tokens, stop_id):
        """Columns the reference loop would have produced: it breaks after the first step at which every row has
        emitted `stop_id` at least once (wrapper.py:247-249)."""
        seen = (tokens == stop_id).cumsum(dim=1) > 0
        done = seen.all(dim=0).nonzero()
        return int(done[0]) + 1 if done.numel() else tokens.shape[1]

    def _generate_batch(self, embed, max_len, top_p, temperature, stop_token=EOT):
        """The reference's decode loop over a prepared prefix (wrapper.py:197-256): embed (B,389,576) -> list[str]."""
        stop_id = self.tokenizer.encode(stop_token)[0]
        b = embed.shape[0]
        eng = self._ensure_engine(b, max_len)
        if b > eng.max_batch:
            raise ValueError(f"_generate_batch: {b} rows exceed this engine's {eng.max_batch}")
        eng.set_option("skip_finished", int(stop_token == EOT))
        eng.set_prefix(embed)
        eng.prefill(b, want_logits=False)
        toks = eng.decode(b, max_len, temperature=temperature, top_p=top_p, eos_id=stop_id)
        return self._detokenize(toks.cpu())

    def _world(self):
        import torch.distributed as dist
        if self.shard is False or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return 0, 1
        return dist.get_rank(), dist.get_world_size()

    def generate(self, examples, max_len, top_p, temperature, stop_token=EOT, audio_resample=True):
        r"""Produces text response for the given audio files and text prompts (reference wrapper.py:258-287).
        examples: list of [audio path 1, audio path 2, text prompt]; max_len: maximum number of generated tokens;
        top_p / temperature: accepted for signature parity -- like the reference, the decision is an argmax and is
        independent of both; stop_token: token that ends a row; audio_resample: resample inputs to 32 kHz."""
        n = len(examples)
        if n == 0:
            return []
        from .dist import gather_rows, shard_bounds
        rank, world = self._world()
        lo, hi = shard_bounds(n, rank, world)
        keep = range(lo, hi)
        paths1 = [e[0] for e in examples]
        paths2 = [e[1] for e in examples]
        prompts = [e[2] for e in examples]
        eng = self._ensure_engine(min(max(hi - lo, 1), MAX_PASS), max_len)
        audio1 = self.preprocess_audio(paths1, resample=audio_resample, keep=keep)   # draws random crops for list 1
        audio2 = self.preprocess_audio(paths2, resample=audio_resample, keep=keep)   # first, then list 2 (wrapper.py:277-278)
        ids = self.preprocess_text(prompts[lo:hi])["input_ids"] if hi > lo else torch.zeros(0, S.TEXT_LEN, dtype=torch.int64)
        stop_id = self.tokenizer.encode(stop_token)[0]
        passes = -(-(hi - lo) // eng.max_batch) if hi > lo else 0
        # The engine evaluates the reference's stop rule (all rows have emitted the stop id) per pass.  With the default
        # stop token the text is cut at '<|endoftext|>' anyway (wrapper.py:254), so passes / ranks may stop
        # independently.  With any other stop token the reference returns whatever each row produced up to the GLOBAL
        # stop step: then nothing may stop or be skipped early, and the global step is applied afterwards.
        exact_tail = stop_token != EOT
        split_run = exact_tail and (passes > 1 or world > 1)
        eng.set_option("skip_finished", 0 if exact_tail else 1)
        rows = []
        for s in range(0, hi - lo, eng.max_batch):                  # passes of the handle's capacity
            e = min(hi - lo, s + eng.max_batch)
            toks = eng.generate(audio1[s:e], audio2[s:e], ids[s:e], max_len, temperature=temperature, top_p=top_p,
                                eos_id=-1 if split_run else stop_id)
            full = torch.full((e - s, max_len), stop_id, dtype=torch.int32)   # columns after a pass stopped: stop ids
            full[:, :toks.shape[1]] = toks.cpu()
            rows.append(full)
        local = torch.cat(rows, 0) if rows else torch.zeros(0, max_len, dtype=torch.int32)
        if world > 1:
            import torch.distributed as dist
            dev = self.model.device if dist.get_backend() == "nccl" else torch.device("cpu")
            local = gather_rows(local.to(dev), n, rank, world).cpu()
        if split_run or world > 1 or passes > 1:
            local = local[:, :self._global_stop(local, stop_id)]
        else:
            local = local[:, :toks.shape[1]]
        return self._detokenize(local)
