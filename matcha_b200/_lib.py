"""ctypes binding of libmatcha_b200.so (the C ABI declared in include/matcha_b200.h) and its build recipe.

The library is built IN-TREE with plain nvcc for sm_100a (no torch headers: the ABI is raw pointers and
sizes).  There is no CPU fallback: if the shared object is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libmatcha_b200.so")
SOURCES = ["engine.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_tcg.cu", "qkg_tiles.cu", "attn_fused.cu", "attn_xform.cu", "chain.cu", "rowwise.cu", "optim.cu", "dp_fused.cu", "sampler.cu", "scorer.cu", "pair_tc.cu", "csr_encoder.cu", "recon_tc.cu", "recon_pipe.cu", "enc_tc.cu", "kmers.cu", "metrics.cu", "denoise.cu", "features.cu"]
NVCC_COMPILE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                      "-Xcompiler", "-fPIC"]
NVCC_LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]
MAX_CHROM = 64

_lock = threading.Lock()
_lib = None


class MatchaError(RuntimeError):
    pass


def _newest_source_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + [os.path.join(ROOT, "include", "matcha_b200.h")]
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into matcha_b200/libmatcha_b200.so (nvcc cross-compiles
    without a GPU).  One object per source under matcha_b200/build/, compiled in parallel and only when
    older than the source tree's headers / its own source; skips everything when the library is current."""
    if (not force) and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdr_mtime = max([os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] +
                    [os.path.getmtime(os.path.join(ROOT, "include", "matcha_b200.h"))])

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if (not force) and os.path.exists(obj) and os.path.getmtime(obj) >= max(hdr_mtime, os.path.getmtime(path)):
            return obj, None
        cmd = [nvcc] + NVCC_COMPILE_FLAGS + os.environ.get("MATCHA_NVCC_EXTRA", "").split() + ["-c", "-o", obj, path]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, (res.stdout + res.stderr) if res.returncode != 0 else None

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    errs = [e for _, e in results if e]
    if errs:
        raise MatchaError("nvcc failed:\n" + "\n".join(errs))
    cmd = [nvcc] + NVCC_LINK_FLAGS + ["-o", LIB_PATH] + [o for o, _ in results]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise MatchaError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


class ModelDesc(C.Structure):
    """Mirror of `matcha_model_desc` (include/matcha_b200.h)."""
    _fields_ = [
        ("d", C.c_int32), ("n_head", C.c_int32), ("n_chrom", C.c_int32), ("attr_dim", C.c_int32),
        ("n_nodes", C.c_int64), ("params", C.c_void_p), ("grads", C.c_void_p),
        ("off_attr_w", C.c_int64), ("off_attr_b", C.c_int64), ("off_next_w", C.c_int64), ("off_next_b", C.c_int64),
        ("off_lnq_g", C.c_int64), ("off_lnq_b", C.c_int64), ("off_lnk_g", C.c_int64), ("off_lnk_b", C.c_int64),
        ("off_lnv_g", C.c_int64), ("off_lnv_b", C.c_int64),
        ("off_wq", C.c_int64), ("off_wk", C.c_int64), ("off_wv", C.c_int64), ("off_fc1_w", C.c_int64), ("off_fc1_b", C.c_int64),
        ("off_pff_w0", C.c_int64), ("off_pff_b0", C.c_int64), ("off_pff_w1", C.c_int64), ("off_pff_b1", C.c_int64),
        ("off_pff_g", C.c_int64), ("off_pff_b", C.c_int64),
        ("off_ln1_g", C.c_int64), ("off_ln1_b", C.c_int64), ("off_ln2_g", C.c_int64), ("off_ln2_b", C.c_int64),
        ("off_cls_w", C.c_int64), ("off_cls_b", C.c_int64),
        ("chrom_start", C.c_int64 * MAX_CHROM), ("chrom_end", C.c_int64 * MAX_CHROM),
        ("off_w0", C.c_int64 * MAX_CHROM), ("off_w1", C.c_int64 * MAX_CHROM),
        ("off_rw", C.c_int64 * MAX_CHROM), ("off_rb", C.c_int64 * MAX_CHROM),
        ("feat", C.c_void_p * MAX_CHROM), ("feat_ld", C.c_int64 * MAX_CHROM),
        ("feat_indptr", C.c_void_p * MAX_CHROM), ("feat_indices", C.c_void_p * MAX_CHROM),
        ("feat_values", C.c_void_p * MAX_CHROM),
        ("attr_table", C.c_void_p), ("inter", C.c_void_p), ("inter_ld", C.c_int64),
        ("derived", C.c_void_p), ("derived_grad", C.c_void_p),
        ("p_feature", C.c_float), ("p_attn", C.c_float), ("p_pff", C.c_float),
    ]


# name -> (restype, argtypes); every symbol include/matcha_b200.h declares
_P, _I32, _I64, _U64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
_MD = C.POINTER(ModelDesc)
SYMBOLS = {
    "matcha_last_error": (C.c_char_p, []),
    "matcha_version": (C.c_int, []),
    "matcha_profile_enable": (None, [_I32]),
    "matcha_profile_labels": (_I32, []),
    "matcha_profile_label_name": (C.c_char_p, [_I32]),
    "matcha_profile_read": (C.c_int, [_P, _P, _P, _I32]),
    "matcha_set_gemm_impl": (None, [_I32]),
    "matcha_set_fused": (None, [_I32]),
    "matcha_set_chain": (None, [_I32]),
    "matcha_set_recon_tc": (None, [_I32]),
    "matcha_set_recon_pipe": (None, [_I32]),
    "matcha_set_gemm_tcg": (None, [_I32]),
    "matcha_set_enc_pipe": (None, [_I32, _I32]),
    "matcha_set_dp_two_shot": (None, [_I32]),
    "matcha_set_enc_tc": (None, [_I32]),
    "matcha_set_xform": (None, [_I32]),
    "matcha_set_mma_passes": (None, [_I32]),
    "matcha_derived_elems": (_I64, [_MD]),
    "matcha_workspace_bytes": (_I64, [_MD, _I64, _I32, _I32]),
    "matcha_prepare": (C.c_int, [_MD, _P]),
    "matcha_forward": (C.c_int, [_MD, _P, _I64, _I32, _I32, _U64, _I32, _P, _P, _P, _I64, _P]),
    "matcha_bce_loss": (C.c_int, [_P, _P, _P, _I64, _F, _F, _P, _P, _P, _P]),
    "matcha_backward": (C.c_int, [_MD, _P, _I64, _I32, _U64, _I32, _P, _F, _P, _P, _I64, _P]),
    "matcha_node_embeddings": (C.c_int, [_MD, _P, _I64, _P, _P, _I64, _P]),
    "matcha_adamw": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _F, _P]),
    "matcha_dp_blocks": (_I32, []),
    "matcha_enable_peer_access": (C.c_int, [_I32]),
    "matcha_ipc_get_handle": (C.c_int, [_P, _P]),
    "matcha_ipc_open": (C.c_int, [_P, _P]),
    "matcha_ipc_close": (C.c_int, [_P]),
    "matcha_dp_barrier_bytes": (_I64, []),
    "matcha_dp_reduce_adamw": (C.c_int, [_I32, _I32, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _I32,
                                         C.c_uint32, _F, _F, _F, _F, _F, _P]),
    "matcha_hashset_insert": (C.c_int, [_P, _I64, _P, _I64, _I32, _P]),
    "matcha_hashset_contains": (C.c_int, [_P, _I64, _P, _I64, _I32, _P, _P]),
    "matcha_neg_sample": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _I32,
                                    _I32, _U64, _U64, _I32, _P, _P, _P, _P]),
    "matcha_pair_tables": (C.c_int, [_MD, _P, _P, _P, _I64, _P]),
    "matcha_pair_score_range": (C.c_int, [_P, _P, _P, _P, _I32, _I64, _I64, _I32, _I64, _I64, _I32, _P, _P]),
    "matcha_pair_count": (_I64, [_I64, _I64, _I32]),
    "matcha_pair_tc_workspace_bytes": (_I64, [_I64, _I64]),
    "matcha_pair_tc_prepare": (C.c_int, [_P, _P, _P, _P, _I32, _I64, _I64, _P, _I64, _P]),
    "matcha_pair_tc_score_range": (C.c_int, [_P, _I64, _I64, _I32, _I64, _I64, _I32, _P, _P]),
    "matcha_kmer_count": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _I32, _P, _I64, _P, _P, _P]),
    "matcha_kmer_collect": (C.c_int, [_P, _I64, _P, _I32, _I32, _P, _P, _I64, _P, _P]),
    "matcha_metrics_workspace_bytes": (_I64, [_I64]),
    "matcha_binary_metrics": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _I64, _P]),
    "matcha_denoise_workspace_bytes": (_I64, [_I64]),
    "matcha_denoise_matrix": (C.c_int, [_P, _P, _I64, _I64, _I32, _P, _P, _I64, _P]),
    "matcha_quantile_uniform": (C.c_int, [_P, _I64, _P, _P, _I32, _P]),
    "matcha_gather_f32": (C.c_int, [_P, _P, _I64, _P, _P]),
    "matcha_pair_gather": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "matcha_corrcoef_workspace_bytes": (_I64, [_I64]),
    "matcha_corrcoef": (C.c_int, [_P, _I64, _I64, _P, _I64, _P, _I64, _P]),
    "matcha_zscore_positive_rows": (C.c_int, [_P, _I64, _I64, _I64, _P]),
    "matcha_adj_from_pixels": (C.c_int, [_P, _P, _P, _I64, _P, _I64, _P, _I64, _P, _P, _P]),
    "matcha_adj_from_clusters": (C.c_int, [_P, _P, _I64, _I64, _P, _P]),
    "matcha_f64_to_f32": (C.c_int, [_P, _P, _I64, _P]),
    "matcha_gemm": (C.c_int, [_I32, _I32, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P, _I64, _P]),
    "matcha_gemm_scratch_floats": (_I64, [_I64]),
}


def load(build_if_missing: bool = True):
    """Load the shared library (building it first if sources are newer).  Raises MatchaError when it
    cannot be produced -- there is deliberately no fallback implementation."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing and os.path.isdir(CSRC):
            try:
                build()
            except (MatchaError, FileNotFoundError, OSError) as e:
                if not os.path.exists(LIB_PATH):
                    raise MatchaError(f"libmatcha_b200.so is missing and could not be built: {e}") from e
        if not os.path.exists(LIB_PATH):
            raise MatchaError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().matcha_last_error()
        raise MatchaError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> int:
    """Raw device (or host) pointer of a torch tensor, 0 for None."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
