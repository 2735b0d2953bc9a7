"""Drop-in for the reference's ``mellow.MellowWrapper`` (mellow/wrapper.py:25-287) on one B200.

Same constructor and ``generate()`` signature, same config names (``config='v0'``, ``model in {'v0','v0_s'}``), same
host-side audio/text preparation rules, same stopping rule and detokenisation; everything between the prepared
tensors and the token ids runs in ``libmellow_b200.so`` (``Engine.generate_host``).

Differences a caller can observe, all deliberate:
  * no hub access is attempted when a local checkpoint / tokenizer is supplied (``checkpoint=``, ``tokenizer=`` or
    ``$MELLOW_CHECKPOINT`` / ``$MELLOW_TOKENIZER``); ``checkpoint='synthetic'`` builds the seeded synthetic
    checkpoint (there are no real weights offline);
  * ``.model`` is an ``Engine`` (native handle), not an ``nn.Module``;
  * a CUDA device is mandatory -- ``use_cuda=False`` / ``device='cpu'`` raises instead of silently running on CPU.
"""
import argparse
import math
import os
import random
from pathlib import Path

import torch
import yaml

from . import schema as S
from .audio_io import load_audio_into_tensor
from .engine import Engine, MellowNativeError
from .tokenizer import load_tokenizer, tokenize_prompts


class MellowWrapper:
    """A class for interfacing the Mellow model (B200-native engine)."""
    model_repo = "soham97/mellow"
    model_name = {"v0": "v0.ckpt", "v0_s": "v0_s.ckpt"}

    def __init__(self, config, model, device, use_cuda=True, *, checkpoint=None, tokenizer=None, policy="split",
                 max_batch=8, max_new_tokens=300):
        self.supported_versions = self.model_name.keys()
        if model not in self.supported_versions:                                       # wrapper.py:37-39
            raise ValueError(f"The model {model} is not supported. The supported versions are {str(self.supported_versions)}")
        if not use_cuda or device == "cpu" or not torch.cuda.is_available():
            raise MellowNativeError("mellow_b200 has no CPU path: a CUDA device (B200, sm_100a) is required")
        self.parent_path = Path(os.path.realpath(__file__)).parent
        self.config_path = os.path.join(self.parent_path, "config", config + ".yaml")
        self.use_cuda = use_cuda
        self.device = device
        self.policy, self.max_batch, self.max_new_tokens = policy, max_batch, max_new_tokens
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
        else:
            state = torch.load(self.model_path, map_location=torch.device("cpu"))      # wrapper.py:74
        engine = Engine(state, device=int(self.device), max_batch=self.max_batch, max_new_tokens=self.max_new_tokens,
                        policy=self.policy)
        tok = load_tokenizer(args.model["decoder"]["text_decoder"], tokenizer)
        params = S.count_parameters()
        print(f"model {os.path.basename(str(self.model_path))}, {os.path.basename(config_path)}, parameter count: {params}")
        return engine, tok, args

    # ------------------------------------------------------------------ host-side preparation
    def load_audio_into_tensor(self, audio_path, audio_duration, resample=True):
        return load_audio_into_tensor(audio_path, audio_duration, self.args.data["sampling_rate"], resample, random)

    def preprocess_audio(self, audio_files, resample):
        """-> (B, 320000) float32 DEVICE tensor.  Files are decoded on the host; resampling to 32 kHz, channel
        flattening and tile-or-crop run on the GPU (the reference does all of it serially on the host and uploads
        one clip at a time, wrapper.py:170-179).  The `random.randrange` draws happen in file order like the reference."""
        from .audio_io import read_wav
        n = len(audio_files)
        out = torch.empty(n, S.CLIP_SAMPLES, dtype=torch.float32, device=self.model.device)
        for i, audio_file in enumerate(audio_files):
            pcm, sr = read_wav(audio_file)
            self.model.prepare_clip(pcm, sr, self.args.data["sampling_rate"], resample, random, out=out[i])
        return out

    def preprocess_text(self, prompts):
        ids = tokenize_prompts(self.tokenizer, prompts, self.args.data["text_tokenization_len"])
        return {"input_ids": ids}

    # ------------------------------------------------------------------ generation
    def _detokenize(self, tokens):
        """wrapper.py:251-254: decode every row and keep the text before the first stop token."""
        out = []
        for row in tokens.tolist():
            text = self.tokenizer.decode(row)
            out.append(text.split("<|endoftext|>")[0])
        return out

    def generate(self, examples, max_len, top_p, temperature, stop_token="<|endoftext|>", audio_resample=True):
        r"""Produces text response for the given audio files and text prompts (reference wrapper.py:258-287).
        examples: list of [audio path 1, audio path 2, text prompt]; max_len: maximum number of generated tokens;
        top_p / temperature: accepted for signature parity -- like the reference, the decision is an argmax and is
        independent of both; stop_token: token that ends a row; audio_resample: resample inputs to 32 kHz."""
        if len(examples) == 0:
            return []
        paths1 = [e[0] for e in examples]
        paths2 = [e[1] for e in examples]
        prompts = [e[2] for e in examples]
        audio1 = self.preprocess_audio(paths1, resample=audio_resample)     # draws random crops for list 1 first,
        audio2 = self.preprocess_audio(paths2, resample=audio_resample)     # then list 2, like wrapper.py:277-278
        ids = self.preprocess_text(prompts)["input_ids"]
        stop_id = self.tokenizer.encode(stop_token)[0]
        preds = []
        for s in range(0, len(examples), self.max_batch):                   # micro-batches of the handle's capacity
            e = min(len(examples), s + self.max_batch)
            toks = self.model.generate(audio1[s:e], audio2[s:e], ids[s:e], max_len, temperature=temperature,
                                       top_p=top_p, eos_id=stop_id)
            preds.extend(self._detokenize(toks.cpu()))
        return preds
