"""Checkpoint schema and architectural constants of the Mellow inference path.

These are facts about the on-disk format the reference consumes
(``torch.load`` of a flat ``state_dict``, reference ``mellow/wrapper.py:74-82``;
key list in SURVEY.md section 8a') and about the model dimensions fixed by
``mellow/config/v0.yaml`` and ``mellow/model/config.py``.
"""
from collections import OrderedDict

# --- data / front end (reference mellow/config/v0.yaml:1-5, mellow/model/config.py:4-9)
SAMPLE_RATE = 32000
SEGMENT_SECONDS = 10
CLIP_SAMPLES = SAMPLE_RATE * SEGMENT_SECONDS          # 320000
N_FFT = 1024
HOP = 320
N_BINS = N_FFT // 2 + 1                               # 513
N_MELS = 64
FMIN, FMAX = 50, 14000
N_FRAMES = CLIP_SAMPLES // HOP + 1                    # 1001 (center=True)
TEXT_LEN = 129                                        # text_tokenization_len
# --- HTSAT (reference mellow/model/htsat.py:599-606)
SPEC_SIZE = 256
FREQ_RATIO = SPEC_SIZE // N_MELS                      # 4
PATCH = 4
EMBED_DIM = 96
DEPTHS = (2, 2, 6, 2)
HEADS = (4, 8, 16, 32)
WINDOW = 8
NUM_CLASSES = 527
ENC_OUT = 768
# --- projection / prefix (v0.yaml:8-15, reference mellow/model/decoder.py:36-55)
D_PROJ = 576
AUDIO_FRAMES = 32            # unique framewise rows per clip
AUDIO_SLOTS = 129            # 1 latent + 128 pooled frame slots
PREFIX_LEN = 389             # 129 + 1 + 129 + 1 + 129
# --- SmolLM2-135M
VOCAB = 49152
HIDDEN = 576
N_LAYERS = 30
N_HEADS = 9
N_KV_HEADS = 3
HEAD_DIM = 64
INTER = 1536
RMS_EPS = 1e-5
ROPE_THETA = 100000.0
MAX_POSITIONS = 8192         # max_position_embeddings; rope table rows of the library (csrc/common.cuh kMaxPos)
LN_EPS = 1e-5
BN_EPS = 1e-5

TOTAL_PARAMS = 167020951     # README.md:4 "167M"; exact value SURVEY.md Appendix B


def stage_dim(i):
    return EMBED_DIM * (2 ** i)


def stage_res(i):
    return (SPEC_SIZE // PATCH) // (2 ** i)


def checkpoint_schema():
    """Ordered {key: (shape, dtype_str)} exactly as ``Mellow(...).state_dict()`` yields it."""
    s = OrderedDict()
    f32, i64 = "float32", "int64"
    h = "audio_encoder.base.htsat."
    s[h + "spectrogram_extractor.stft.conv_real.weight"] = ((N_BINS, 1, N_FFT), f32)
    s[h + "spectrogram_extractor.stft.conv_imag.weight"] = ((N_BINS, 1, N_FFT), f32)
    s[h + "logmel_extractor.melW"] = ((N_BINS, N_MELS), f32)
    for n in ("weight", "bias", "running_mean", "running_var"):
        s[h + "bn0." + n] = ((N_MELS,), f32)
    s[h + "bn0.num_batches_tracked"] = ((), i64)
    s[h + "patch_embed.proj.weight"] = ((EMBED_DIM, 1, PATCH, PATCH), f32)
    s[h + "patch_embed.proj.bias"] = ((EMBED_DIM,), f32)
    s[h + "patch_embed.norm.weight"] = ((EMBED_DIM,), f32)
    s[h + "patch_embed.norm.bias"] = ((EMBED_DIM,), f32)
    for i, depth in enumerate(DEPTHS):
        C, nH, R = stage_dim(i), HEADS[i], stage_res(i)
        for b in range(depth):
            p = f"{h}layers.{i}.blocks.{b}."
            if b % 2 == 1 and R > WINDOW:
                s[p + "attn_mask"] = (((R // WINDOW) ** 2, WINDOW * WINDOW, WINDOW * WINDOW), f32)
            s[p + "norm1.weight"] = ((C,), f32)
            s[p + "norm1.bias"] = ((C,), f32)
            s[p + "attn.relative_position_bias_table"] = (((2 * WINDOW - 1) ** 2, nH), f32)
            s[p + "attn.relative_position_index"] = ((WINDOW * WINDOW, WINDOW * WINDOW), i64)
            s[p + "attn.qkv.weight"] = ((3 * C, C), f32)
            s[p + "attn.qkv.bias"] = ((3 * C,), f32)
            s[p + "attn.proj.weight"] = ((C, C), f32)
            s[p + "attn.proj.bias"] = ((C,), f32)
            s[p + "norm2.weight"] = ((C,), f32)
            s[p + "norm2.bias"] = ((C,), f32)
            s[p + "mlp.fc1.weight"] = ((4 * C, C), f32)
            s[p + "mlp.fc1.bias"] = ((4 * C,), f32)
            s[p + "mlp.fc2.weight"] = ((C, 4 * C), f32)
            s[p + "mlp.fc2.bias"] = ((C,), f32)
        if i < len(DEPTHS) - 1:
            p = f"{h}layers.{i}.downsample."
            s[p + "reduction.weight"] = ((2 * C, 4 * C), f32)
            s[p + "norm.weight"] = ((4 * C,), f32)
            s[p + "norm.bias"] = ((4 * C,), f32)
    s[h + "norm.weight"] = ((ENC_OUT,), f32)
    s[h + "norm.bias"] = ((ENC_OUT,), f32)
    s[h + "tscam_conv.weight"] = ((NUM_CLASSES, ENC_OUT, 2, 3), f32)
    s[h + "tscam_conv.bias"] = ((NUM_CLASSES,), f32)
    s[h + "head.weight"] = ((NUM_CLASSES, NUM_CLASSES), f32)
    s[h + "head.bias"] = ((NUM_CLASSES,), f32)
    s["audio_encoder.base.c2l.weight"] = ((ENC_OUT, NUM_CLASSES), f32)
    s["audio_encoder.base.c2l.bias"] = ((ENC_OUT,), f32)
    s["audio_encoder.projection.linear1.weight"] = ((D_PROJ, ENC_OUT), f32)
    s["audio_encoder.projection.linear2.weight"] = ((D_PROJ, D_PROJ), f32)
    s["audio_encoder.projection.layer_norm.weight"] = ((D_PROJ,), f32)
    s["audio_encoder.projection.layer_norm.bias"] = ((D_PROJ,), f32)
    lm = "caption_decoder.lm."
    s[lm + "model.embed_tokens.weight"] = ((VOCAB, HIDDEN), f32)
    for l in range(N_LAYERS):
        p = f"{lm}model.layers.{l}."
        s[p + "self_attn.q_proj.weight"] = ((N_HEADS * HEAD_DIM, HIDDEN), f32)
        s[p + "self_attn.k_proj.weight"] = ((N_KV_HEADS * HEAD_DIM, HIDDEN), f32)
        s[p + "self_attn.v_proj.weight"] = ((N_KV_HEADS * HEAD_DIM, HIDDEN), f32)
        s[p + "self_attn.o_proj.weight"] = ((HIDDEN, N_HEADS * HEAD_DIM), f32)
        s[p + "mlp.gate_proj.weight"] = ((INTER, HIDDEN), f32)
        s[p + "mlp.up_proj.weight"] = ((INTER, HIDDEN), f32)
        s[p + "mlp.down_proj.weight"] = ((HIDDEN, INTER), f32)
        s[p + "input_layernorm.weight"] = ((HIDDEN,), f32)
        s[p + "post_attention_layernorm.weight"] = ((HIDDEN,), f32)
    s[lm + "model.norm.weight"] = ((HIDDEN,), f32)
    s[lm + "lm_head.weight"] = ((VOCAB, HIDDEN), f32)      # tied alias of embed_tokens
    return s


BUFFER_SUFFIXES = ("attn_mask", "relative_position_index", "running_mean", "running_var", "num_batches_tracked")


def count_parameters(schema=None):
    """Unique trainable/frozen nn.Parameter elements (buffers and the tied lm_head alias excluded)."""
    schema = schema or checkpoint_schema()
    total = 0
    for k, (shape, _) in schema.items():
        if k.endswith(BUFFER_SUFFIXES) or k.endswith("lm_head.weight"):
            continue
        n = 1
        for d in shape:
            n *= d
        total += n
    return total
