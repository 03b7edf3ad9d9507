"""ctypes binding of libofa_b200.so (include/ofa_b200.h).  There is no fallback: if the library is missing or a
call fails, this raises -- the product never routes around its CUDA kernels."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OFA_B200_LIB") or os.path.join(_HERE, "libofa_b200.so")    # override: kernel experiments

_lib = None

c_ll, c_i, c_f, c_p = C.c_longlong, C.c_int, C.c_float, C.c_void_p


class OfaAttnBias(C.Structure):
    _fields_ = [("tok_lut", c_p), ("tok_max", c_i), ("q_text_off", c_i), ("k_text_off", c_i),
                ("img_lut", c_p), ("n_img_rel", c_i), ("ibs", c_i), ("q_pid", c_p), ("k_pid", c_p),
                ("n_img_q", c_i), ("n_img_k", c_i)]


class OfaAttnArgs(C.Structure):
    _fields_ = [("q", c_p), ("pq", c_p), ("k", c_p), ("pk", c_p), ("v", c_p), ("o", c_p), ("lse", c_p),
                ("ldq", c_ll), ("ldpq", c_ll), ("ldk", c_ll), ("ldpk", c_ll), ("ldv", c_ll), ("ldo", c_ll),
                ("bsq", c_ll), ("bspq", c_ll), ("bsk", c_ll), ("bspk", c_ll), ("bsv", c_ll), ("bso", c_ll),
                ("B", c_i), ("H", c_i), ("T", c_i), ("S", c_i), ("causal", c_i), ("q_pos_off", c_i),
                ("kpm", c_p), ("head_scale", c_p), ("p_round_bf16", c_i), ("bias", OfaAttnBias)]


class OfaAttnGrads(C.Structure):
    _fields_ = [("dout", c_p), ("dq", c_p), ("dpq", c_p), ("dk", c_p), ("dpk", c_p), ("dv", c_p),
                ("lddq", c_ll), ("lddpq", c_ll), ("lddk", c_ll), ("lddpk", c_ll), ("lddv", c_ll),
                ("bsdq", c_ll), ("bsdpq", c_ll), ("bsdk", c_ll), ("bsdpk", c_ll), ("bsdv", c_ll),
                ("dtok_lut", c_p), ("dimg_lut", c_p), ("delta", c_p), ("P", c_p), ("dS", c_p), ("dq_scale", c_f), ("acc_pos", c_i)]


class OfaDecodeArgs(C.Structure):
    _fields_ = [("q", c_p), ("pq", c_p), ("ldq", c_ll), ("ldpq", c_ll),
                ("k", c_p), ("v", c_p), ("pk", c_p),
                ("ldk", c_ll), ("bsk", c_ll), ("ldv", c_ll), ("bsv", c_ll), ("ldpk", c_ll), ("bspk", c_ll),
                ("kv_row", c_p), ("pk_row", c_p), ("kpm", c_p), ("kpm_stride", c_ll),
                ("o", c_p), ("ldo", c_ll), ("head_scale", c_p), ("tok_lut", c_p), ("tok_max", c_i), ("q_pos", c_i),
                ("R", c_i), ("G", c_i), ("H", c_i), ("S", c_i),
                ("bias_in", c_p), ("score_out", c_p), ("bias_ld", c_ll),
                ("page_table", c_p), ("page_len", c_i), ("max_pages", c_i), ("page_stride", c_ll)]


class OfaBeamArgs(C.Structure):
    _fields_ = [("logits", c_p), ("ld", c_ll), ("dtype", c_i), ("R", c_i), ("beam", c_i), ("V", c_i), ("K", c_i),
                ("temperature", c_f), ("prev_scores", c_p), ("step0", c_i), ("eos", c_i), ("pad", c_i), ("unk", c_i),
                ("unk_penalty", c_f), ("block_eos", c_i), ("force_eos", c_i), ("eos_one", c_i),
                ("range_lo", c_i), ("range_hi", c_i), ("range_post", c_i),
                ("trie_ptr", c_p), ("trie_tok", c_p), ("node", c_p), ("trie_post", c_i),
                ("tokens", c_p), ("ldtok", c_ll), ("step", c_i), ("ngram", c_i),
                ("prefix_tok", c_p), ("prefix_fill", c_p),
                ("row_val", c_p), ("row_idx", c_p), ("cand_scores", c_p), ("cand_index", c_p)]


# name -> argtypes, exactly the prototypes of include/ofa_b200.h
SIGNATURES = {
    "ofa_abi_version": [],
    "ofa_set_pdl": [c_i],
    "ofa_gemm_bf16": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_ll, c_ll, c_ll, c_ll, c_ll, c_ll, c_i, c_i, c_i, c_p, c_f,
                      c_i, c_p, c_ll, c_ll, c_p, c_ll, c_i, c_p],
    "ofa_gemm_workspace_bytes": [c_i, c_i, c_i, c_i],
    "ofa_gemm_set_pair_mode": [c_i],
    "ofa_gemm_set_tma_store": [c_i],
    "ofa_gemm_set_wgrad_bn256": [c_i],
    "ofa_gemm_set_small64": [c_i],
    "ofa_attn_decode_set_short": [c_i],
    "ofa_attn_decode_set_online": [c_i],
    "ofa_gemm_set_pair_min_tiles": [c_i],
    "ofa_split3_bf16": [c_p, c_ll, c_i, c_i, c_p, c_ll, c_ll, c_i, c_p],
    "ofa_layernorm_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_f, c_i, c_i, c_p],
    "ofa_layernorm_bwd_nparts": [c_i],
    "ofa_layernorm_set_staged": [c_i],
    "ofa_layernorm_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_p],
    "ofa_colsum": [c_p, c_ll, c_i, c_i, c_p, c_p, c_f, c_i, c_i, c_p],
    "ofa_embed_gather": [c_p, c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_p],
    "ofa_embed_scatter_add": [c_p, c_p, c_ll, c_p, c_i, c_i, c_ll, c_i, c_p],
    "ofa_add": [c_p, c_p, c_p, c_ll, c_i, c_p],
    "ofa_mask_rows": [c_p, c_p, c_i, c_i, c_i, c_p],
    "ofa_gelu": [c_p, c_p, c_p, c_ll, c_i, c_i, c_p],
    "ofa_dropout_residual": [c_p, c_p, c_p, c_ll, c_i, c_i, c_f, c_p, c_p, c_i, c_p],
    "ofa_batchnorm_workspace_floats": [c_i],
    "ofa_batchnorm_set_tuning": [c_i, c_i],
    "ofa_batchnorm_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_i, c_f, c_f, c_i, c_i, c_p, c_p, c_i, c_i, c_p],
    "ofa_batchnorm_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_i, c_i, c_i, c_p, c_i, c_i, c_p],
    "ofa_ls_ce_fwd_bwd": [c_p, c_ll, c_p, c_p, c_p, c_i, c_i, c_i, c_ll, c_f, c_i, c_i, c_i, c_f, c_p, c_p, c_p, c_p,
                          c_i, c_p],
    "ofa_conv3x3_bf16": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "ofa_conv3x3_wgrad_workspace_bytes": [c_i, c_i, c_i, c_i, c_i],
    "ofa_conv3x3_wgrad_bf16": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_ll, c_p],
    "ofa_attn_decode": [C.POINTER(OfaDecodeArgs), c_i, c_p],
    "ofa_cache_gather": [c_p, c_p, c_p, c_i, c_i, c_i, c_ll, c_ll, c_i, c_i, c_p],
    "ofa_page_reorder": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "ofa_page_write": [c_p, c_p, c_p, c_p, c_ll, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "ofa_beam_topk": [C.POINTER(OfaBeamArgs), c_p],
    "ofa_beam_topk_width": [c_i],
    "ofa_trie_advance": [c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_p, c_i, c_p],
    "ofa_normalize_u8": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_i, c_p],
    "ofa_beam_advance": [c_p, c_p, c_i, c_p, c_p, c_ll, c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "ofa_trie_score": [c_p, c_ll, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_i, c_p],
    "ofa_maxpool3x3s2_fwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "ofa_maxpool3x3s2_bwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "ofa_subsample2": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "ofa_im2col": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_ll, c_i, c_p],
    "ofa_col2im": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_ll, c_i, c_p],
    "ofa_maxpool3x3s2_any": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "ofa_stem_patches": [c_p, c_p, c_i, c_i, c_i, c_p],
    "ofa_adam_step": [c_p, c_i, c_p, c_p, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_f, c_i, c_p],
    "ofa_scale_rows": [c_p, c_ll, c_i, c_i, c_p, c_p, c_i, c_i, c_p],
    "ofa_attn_fwd_simt": [C.POINTER(OfaAttnArgs), c_i, c_p],
    "ofa_attn_bwd_simt": [C.POINTER(OfaAttnArgs), C.POINTER(OfaAttnGrads), c_i, c_p],
    "ofa_attn_fwd_tc": [C.POINTER(OfaAttnArgs), c_p],
    "ofa_attn_set_fwd_ws": [c_i],
    "ofa_attn_set_bwd_small": [c_i],
    "ofa_attn_bwd_tc": [C.POINTER(OfaAttnArgs), C.POINTER(OfaAttnGrads), c_p, c_p],
}


class OfaKernelError(RuntimeError):
    pass


def load(path=None):
    """Load the shared library (building is `python -m musketeer_b200.build`).  Raises if it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise OfaKernelError(
            "musketeer_b200: %s not found. Build it with `python -m musketeer_b200.build` "
            "(there is no CPU / PyTorch fallback for the OFA hot path)." % path)
    lib = C.CDLL(path)
    lib.ofa_last_error.restype = C.c_char_p
    lib.ofa_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = c_ll if name.endswith(("_bytes", "_floats")) else c_i
    if os.environ.get("OFA_WGRAD_BN256") is not None:
        lib.ofa_gemm_set_wgrad_bn256(int(os.environ["OFA_WGRAD_BN256"]))
    if os.environ.get("OFA_PAIR_MIN_TILES") is not None:
        lib.ofa_gemm_set_pair_min_tiles(int(os.environ["OFA_PAIR_MIN_TILES"]))
    if os.environ.get("OFA_BN_TUNING") is not None:      # "waves,bwd_unroll"
        w, u = os.environ["OFA_BN_TUNING"].split(",")
        lib.ofa_batchnorm_set_tuning(int(w), int(u))
    if os.environ.get("OFA_LN_STAGED") is not None:
        lib.ofa_layernorm_set_staged(int(os.environ["OFA_LN_STAGED"]))
    if os.environ.get("OFA_ATTN_FWD_WS") is not None:   # A/B switch: warp-specialised attention forward (default on)
        lib.ofa_attn_set_fwd_ws(int(os.environ["OFA_ATTN_FWD_WS"]))
    if os.environ.get("OFA_ATTN_BWD_SMALL") is not None:     # A/B switch: short-query attention backward kernel
        lib.ofa_attn_set_bwd_small(int(os.environ["OFA_ATTN_BWD_SMALL"]))
    if os.environ.get("OFA_GEMM_SMALL64") is not None:      # A/B switch: 128 x 64 tiles for small-M GEMMs
        lib.ofa_gemm_set_small64(int(os.environ["OFA_GEMM_SMALL64"]))
    if os.environ.get("OFA_DECODE_ONLINE") is not None:     # A/B switch: one-pass (online softmax) long-key decode attention
        lib.ofa_attn_decode_set_online(int(os.environ["OFA_DECODE_ONLINE"]))
    if os.environ.get("OFA_DECODE_SHORT") is not None:      # A/B switch: warp-per-(row, head) self-attention decode kernel
        lib.ofa_attn_decode_set_short(int(os.environ["OFA_DECODE_SHORT"]))
    if os.environ.get("OFA_PDL") is not None:       # A/B switch for programmatic dependent launch
        lib.ofa_set_pdl(int(os.environ["OFA_PDL"]))
    _lib = lib
    return lib


_SYNC_DEBUG = bool(int(os.environ.get("OFA_SYNC_DEBUG", "0")))
LAUNCHES = 0          # number of C-ABI kernel entry calls (each launches >= 1 kernel of this library)
PROFILE = None        # when a dict: name -> list of (start_event, end_event, work) recorded on the launching stream


# NVTX ranges around every C-ABI launch (the reference wraps its step in record_function / emit_nvtx scopes: ofa_task.py:337-346,
# train.py:537-540): on when OFA_NVTX=1, after set_nvtx(True), or whenever a torch profiler / emit_nvtx context is active.
NVTX = bool(int(os.environ.get("OFA_NVTX", "0")))


def set_nvtx(enabled):
    global NVTX
    NVTX = bool(enabled)


def _nvtx_on():
    if NVTX:
        return True
    try:
        import torch
        return torch.autograd._profiler_enabled()
    except Exception:
        return False


def call(name, *args, work=None):
    global LAUNCHES
    lib = load()
    LAUNCHES += 1
    if _nvtx_on():
        import torch
        torch.cuda.nvtx.range_push(name)
        try:
            rc = getattr(lib, name)(*args)
        finally:
            torch.cuda.nvtx.range_pop()
        if rc != 0:
            raise OfaKernelError("%s: %s" % (name, lib.ofa_last_error().decode()))
        return
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        PROFILE.setdefault(name, []).append((e0, e1, work))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise OfaKernelError("%s: %s" % (name, lib.ofa_last_error().decode()))
    if _SYNC_DEBUG:      # debugging aid: surface asynchronous faults at the call that caused them
        import torch
        try:
            torch.cuda.synchronize()
        except Exception as e:
            raise OfaKernelError("%s faulted asynchronously: %s" % (name, str(e).splitlines()[0]))
