"""Build and load ``libmellow_b200.so`` (the C-ABI in ``include/mellow_b200.h``) through ctypes.

The library is compiled in-tree with plain ``nvcc`` for sm_100a (no torch extension: the boundary carries raw
pointers only).  There is no CPU fallback: if the shared object is missing and cannot be built, importing the engine
raises.
"""
import ctypes
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libmellow_b200.so")
HASH_PATH = os.path.join(CSRC, "libmellow_b200.srchash")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
SOURCES = ["api.cu", "frontend.cu", "encoder.cu", "lm.cu", "attn_mma.cu", "attn_umma.cu", "gemm_umma.cu", "gemm_skinny.cu", "audio.cu"]
LAB_SOURCES = ["gemm_mma.cu"]          # mma.sync cross-check engine: only with MB_BUILD_LAB=1 (-DMB_LAB)
HEADERS = ["common.cuh", "gemm.cuh", "kernels.cuh", "umma.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def lab_build():
    """MB_BUILD_LAB=1: also compile the mma.sync cross-check engine (``-DMB_LAB``); the product library ships without it."""
    return bool(os.environ.get("MB_BUILD_LAB"))


def _sources():
    return SOURCES + (LAB_SOURCES if lab_build() else [])


def _flags():
    return NVCC_FLAGS + (["-DMB_LAB"] if lab_build() else [])


def source_hash():
    h = hashlib.sha256()
    for name in _sources() + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(INCLUDE, "mellow_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def is_stale():
    if not os.path.isfile(LIB_PATH) or not os.path.isfile(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library in-tree.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + _flags() + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and (r.stdout or r.stderr):
            sys.stderr.write(r.stdout + r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(_sources()))) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(HASH_PATH, "w") as f:
        f.write(source_hash())
    return LIB_PATH


# (name, restype, argtypes) for every symbol declared in include/mellow_b200.h
_vp, _i, _ll, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
SYMBOLS = [
    ("mb_version", _i, []),
    ("mb_weight_entry_count", _i, []),
    ("mb_weight_entry_name", ctypes.c_char_p, [_i]),
    ("mb_weight_entry_offset", _ll, [_i]),
    ("mb_weight_entry_bytes", _ll, [_i]),
    ("mb_weights_size", _ll, []),
    ("mb_create", _vp, [_i, _i, _i, _i]),
    ("mb_destroy", None, [_vp]),
    ("mb_last_error", ctypes.c_char_p, [_vp]),
    ("mb_bind_weights", _i, [_vp, _vp, _ll]),
    ("mb_workspace_bytes", _ll, [_vp]),
    ("mb_set_option", _i, [_vp, ctypes.c_char_p, _i]),
    ("mb_set_trace", _i, [_vp, _vp]),
    ("mb_kernel_launches", _ll, [_vp]),
    ("mb_frontend", _i, [_vp, _vp, _i, _vp, _vp, _vp]),
    ("mb_encode", _i, [_vp, _vp, _vp, _i, _vp, _vp]),
    ("mb_encode_heads", _i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    ("mb_encode_tap", _i, [_vp, _vp, _i, _i, _vp, _vp]),
    ("mb_prefix", _i, [_vp, _vp, _i, _vp, _vp]),
    ("mb_set_prefix", _i, [_vp, _vp, _i, _vp]),
    ("mb_prefill", _i, [_vp, _i, _vp, _vp]),
    ("mb_lm_forward_last", _i, [_vp, _vp, _i, _i, _vp, _vp]),
    ("mb_embed_tokens", _i, [_vp, _vp, _i, _vp, _vp]),
    ("mb_decode", _i, [_vp, _i, _i, _f, _f, _i, _vp, ctypes.POINTER(_i), _vp, _vp, _vp]),
    ("mb_generate", _i, [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _vp, ctypes.POINTER(_i), _vp]),
    ("mb_generate_host", _i, [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _vp, ctypes.POINTER(_i), _vp]),
    ("mb_audio_resample", _i, [_vp, _vp, _ll, _i, _i, _vp, _i, _i, _vp, _ll, _vp]),
    ("mb_audio_fit", _i, [_vp, _vp, _ll, _ll, _vp, _vp]),
    ("mb_bench_decode_attention", _i, [_vp, _i, _i, _i, _vp]),
    ("mb_op_gemm", _i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
]

_lib = None


def load(auto_build=True):
    """Load the shared library (building it first when the in-tree copy is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if auto_build and is_stale():
        try:
            build()
        except Exception as exc:
            # A library whose recorded source hash differs from the sources next to it may have another ABI; loading
            # it could corrupt memory silently.  Opt in explicitly (MB_ALLOW_STALE_LIB=1) to use it anyway.
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError("libmellow_b200.so is missing and could not be built; there is no fallback path") from exc
            if not os.environ.get("MB_ALLOW_STALE_LIB"):
                raise RuntimeError("libmellow_b200.so is stale (sources changed) and the rebuild failed; fix the build "
                                   "or set MB_ALLOW_STALE_LIB=1 to load the old library") from exc
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing; run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)                                   # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def weight_entries(lib=None):
    """[(name, offset, bytes)] of the device weight arena, as defined by the library."""
    lib = lib or load()
    out = []
    for i in range(lib.mb_weight_entry_count()):
        out.append((lib.mb_weight_entry_name(i).decode(), lib.mb_weight_entry_offset(i), lib.mb_weight_entry_bytes(i)))
    return out
