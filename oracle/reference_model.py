"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Builds the *unmodified* reference model classes from ``/root/reference`` (when
that tree is present, i.e. in the build container; it does not exist on the GPU
box) so that the standalone restatement in ``oracle/restated.py`` can be pinned
against them and so that golden vectors can be generated
(``tests/golden/make_golden.py``).

Import recipe:
  * ``mellow/model/*.py`` are executed as sub-modules of a synthetic package so
    that ``mellow/__init__.py`` (which pulls in ``wrapper.py`` and its missing
    ``importlib_resources`` / hub downloads, reference ``wrapper.py:11,41-42``)
    never runs;
  * the ``torchlibrosa`` stand-in under ``oracle/standin`` is put on
    ``sys.path`` (reference ``htsat.py:7-8``);
  * ``AutoModelForCausalLM.from_pretrained`` (reference ``decoder.py:25``) is
    patched to build ``LlamaForCausalLM`` from a SmolLM2-135M-shaped config,
    because the hub is unreachable.
"""
import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MELLOW_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = "_mellow_ref_model"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mellow", "model", "htsat.py"))


def smollm2_config():
    """SmolLM2-135M hyper-parameters."""
    from transformers import LlamaConfig
    return LlamaConfig(
        vocab_size=49152, hidden_size=576, intermediate_size=1536, num_hidden_layers=30,
        num_attention_heads=9, num_key_value_heads=3, head_dim=64, hidden_act="silu",
        max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=100000.0,
        tie_word_embeddings=True, attention_bias=False, mlp_bias=False,
        bos_token_id=0, eos_token_id=0, pad_token_id=None,
    )


def _load_reference_modules():
    if _PKG in sys.modules:
        return sys.modules[_PKG]
    standin = os.path.join(_HERE, "standin")
    if standin not in sys.path:
        sys.path.insert(0, standin)
    model_dir = os.path.join(REFERENCE_ROOT, "mellow", "model")
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [model_dir]
    sys.modules[_PKG] = pkg
    for name in ("config", "htsat", "audio", "decoder", "mellow", "model"):
        spec = importlib.util.spec_from_file_location(f"{_PKG}.{name}", os.path.join(model_dir, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"{_PKG}.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
    return pkg


def build_reference_model(state_dict=None):
    """Instantiate reference ``Mellow`` exactly as ``wrapper.py:67-73`` + ``v0.yaml`` do."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    import transformers
    from transformers import LlamaForCausalLM
    pkg = _load_reference_modules()
    orig = transformers.AutoModelForCausalLM.from_pretrained
    decoder_mod = sys.modules[f"{_PKG}.decoder"]
    try:
        patched = lambda *a, **k: LlamaForCausalLM(smollm2_config())
        transformers.AutoModelForCausalLM.from_pretrained = patched
        decoder_mod.AutoModelForCausalLM.from_pretrained = patched
        Model = pkg.model.get_model_class("Mellow")
        with torch.no_grad():
            model = Model(audioenc_name="HTSAT", d_in=768, text_decoder="HuggingFaceTB/SmolLM2-135M",
                          prefix_length=389, d_out=576)
    finally:
        transformers.AutoModelForCausalLM.from_pretrained = orig
        decoder_mod.AutoModelForCausalLM.from_pretrained = orig
    if state_dict is not None:
        model.load_state_dict(state_dict)  # strict, like wrapper.py:75-76
    model.eval()
    return model


@torch.no_grad()
def reference_generate_ids(model, prefix, max_len, top_p, temperature, stop_token_index=0, dump_logits=False):
    """Verbatim restatement of the decode loop ``wrapper.py:197-256`` (minus tqdm and the
    detokenisation), driving the *reference* LM object.  Returns (tokens (B,steps) int64, [logits])."""
    tokens = None
    generated = prefix
    filter_value = -float("Inf")
    dumped = []
    import torch.nn.functional as F
    for _ in range(max_len):
        outputs = model.caption_decoder.lm(inputs_embeds=generated)
        logits = outputs.logits
        logits = logits[:, -1, :] / (temperature if temperature > 0 else 1.0)
        if dump_logits:
            dumped.append(logits.clone())
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        sorted_indices_to_remove = cumulative_probs > top_p
        sorted_indices_to_remove[..., 1:] = sorted_indices_to_remove[..., :-1].clone()
        sorted_indices_to_remove[..., 0] = 0
        for k in range(len(sorted_indices_to_remove)):
            indices_to_remove = sorted_indices[k][sorted_indices_to_remove[k]]
            logits[k, indices_to_remove] = filter_value
        next_token = torch.argmax(logits, -1).unsqueeze(1)
        next_token_embed = model.caption_decoder.lm.model.embed_tokens(next_token)
        tokens = next_token if tokens is None else torch.cat((tokens, next_token), dim=1)
        generated = torch.cat((generated, next_token_embed), dim=1)
        condition = (tokens == stop_token_index).sum(dim=-1)
        if (condition > 0).all():
            break
    return (tokens, dumped) if dump_logits else tokens
