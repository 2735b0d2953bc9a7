"""Tokenizer plumbing of the drop-in wrapper (reference mellow/wrapper.py:84-85,181-195,254).

The reference uses the SmolLM2 BPE tokenizer from the Hugging Face hub.  With a local copy (directory given by
``tokenizer=...`` or ``$MELLOW_TOKENIZER``) the same tokenizer is used.  ``ByteStandInTokenizer`` keeps the pipeline
runnable for synthetic-weight tests and is only used when asked for (``tokenizer='stand-in'``, the default of
``checkpoint='synthetic'``); it is NOT the SmolLM2 vocabulary and its text is meaningless.
"""
import os

import torch


class ByteStandInTokenizer:
    """ids: 0 = <|endoftext|>, 17 = '!' (pad, like SmolLM2), 256 + byte otherwise."""
    eos_token = "<|endoftext|>"
    pad_token = "!"
    is_stand_in = True

    def encode(self, text):
        if text == self.eos_token:
            return [0]
        out = []
        for part in text.split(self.eos_token):
            out.extend(17 if b == 0x21 else 256 + b for b in part.encode("utf-8"))
            out.append(0)
        return out[:-1]

    def decode(self, ids):
        chunks, buf = [], bytearray()
        for i in ids:
            i = int(i)
            if i == 0:
                chunks.append(buf.decode("utf-8", errors="replace") + self.eos_token)
                buf = bytearray()
            elif i == 17:
                buf.append(0x21)
            elif 256 <= i < 512:
                buf.append(i - 256)
            else:
                buf.extend(f"<{i}>".encode())
        chunks.append(buf.decode("utf-8", errors="replace"))
        return "".join(chunks)

    def pad_to(self, ids, length):
        ids = ids[:length]
        return ids + [17] * (length - len(ids))


STAND_IN = "stand-in"


def load_tokenizer(name_or_path, local=None):
    """``local`` / ``$MELLOW_TOKENIZER``: a local SmolLM2 tokenizer directory, or the string ``'stand-in'`` to ask for
    ``ByteStandInTokenizer`` explicitly (synthetic-weight runs).  Anything else goes to the hub like the reference
    (wrapper.py:84) and a failure there RAISES: decoding a real checkpoint through the stand-in vocabulary would return
    meaningless text without an error."""
    path = local or os.environ.get("MELLOW_TOKENIZER")
    if path == STAND_IN:
        return ByteStandInTokenizer()
    from transformers import AutoTokenizer
    try:
        tok = AutoTokenizer.from_pretrained(path or name_or_path)
    except Exception as exc:
        raise RuntimeError(
            f"cannot load the tokenizer {path or name_or_path!r} ({type(exc).__name__}: {exc}); pass tokenizer=<local "
            f"SmolLM2 tokenizer directory> (or $MELLOW_TOKENIZER), or tokenizer='{STAND_IN}' for synthetic-weight runs") from exc
    tok.add_special_tokens({"pad_token": "!"})               # wrapper.py:85
    return tok


def tokenize_prompts(tokenizer, prompts, length):
    """-> (B, length) int64, right-padded / truncated (reference wrapper.py:186-190; ``encode_plus`` with
    ``pad_to_max_length`` no longer exists in transformers 5, the call below is its equivalent)."""
    rows = []
    for text in prompts:
        if getattr(tokenizer, "is_stand_in", False):
            ids = tokenizer.pad_to(tokenizer.encode(text), length)
        else:
            ids = tokenizer(text, add_special_tokens=True, truncation=True, max_length=length,
                            padding="max_length")["input_ids"]
        rows.append(torch.tensor(ids, dtype=torch.int64))
    return torch.stack(rows, 0)
