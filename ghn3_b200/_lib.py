"""
ctypes binding of libghn3_b200.so (the C ABI declared in include/ghn3_b200.h).

There is no fallback: if the library is missing it is built with nvcc (ghn3_b200.build); if that fails, or a call
returns a non-zero status, a RuntimeError is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libghn3_b200.so')

BF16, TF32, F32 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
SCATTER_CHUNK = 16384

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class SpdArgs(C.Structure):
    _fields_ = [('n_graphs', i32), ('cutoff', i32), ('node_off', vp), ('edge_off', vp), ('mat_off', vp),
                ('edge_src', vp), ('edge_dst', vp), ('max_nodes', i32), ('total_nodes', i32), ('total_edges', i32),
                ('bits_total', i64), ('mat_total', i64), ('adj_bits', vp), ('bits_off', vp), ('spd', vp)]


class DeriveArgs(C.Structure):
    _fields_ = [('n_graphs', i32), ('vmax', i32), ('node_off', vp), ('mat_off', vp), ('max_nodes', i32),
                ('total_nodes', i32), ('spd', vp), ('pair', vp), ('deg_in', vp), ('deg_out', vp), ('dist0', vp)]


class NodeFeaturesArgs(C.Structure):
    _fields_ = [('total_nodes', i32), ('hid', i32), ('op', vp), ('shape_idx', vp), ('deg_in', vp), ('deg_out', vp),
                ('dist0', vp), ('embed_op', vp), ('embed_ch', vp), ('embed_sp', vp), ('cent_in', vp),
                ('cent_out', vp), ('dist_embed', vp), ('x', vp)]


class EdgeLutArgs(C.Structure):
    _fields_ = [('hid', i32), ('heads', i32), ('vmax', i32), ('edge_embed', vp), ('w1', vp), ('b1', vp), ('w2', vp),
                ('b2', vp), ('workspace', vp), ('lut', vp)]


class LayerNormArgs(C.Structure):
    _fields_ = [('rows', i32), ('hid', i32), ('x', vp), ('gamma', vp), ('beta', vp), ('out', vp), ('out_dtype', i32),
                ('dst_row', vp), ('out_f32', vp)]


class GemmProblem(C.Structure):
    _fields_ = [('a_row0', i32), ('b_row0', i32), ('m', i32), ('n', i32), ('d_off', i64), ('ldd', i32),
                ('bias_off', i32)]


class GemmArgs(C.Structure):
    _fields_ = [('a', vp), ('a_rows', i64), ('lda', i64), ('b', vp), ('b_rows', i64), ('ldb', i64), ('k', i32),
                ('in_dtype', i32), ('d', vp), ('out_dtype', i32), ('bias', vp), ('act', i32), ('accumulate', i32),
                ('single', GemmProblem), ('problems', vp), ('tiles', vp), ('n_tiles', i32), ('block_n', i32),
                ('k_splits', i32), ('tf32_x3', i32), ('b_group', i32), ('b_group_stride', i32), ('bias_rows', i32), ('b_dynamic', i32), ('rowmap', vp), ('swap_ab', i32), ('ln_out', vp),
                ('ln_gamma', vp), ('ln_beta', vp), ('ln_counters', vp), ('ln_out_dtype', i32), ('kb_list', vp),
                ('kb_off', vp), ('persistent_single', i32)]


class GemmSimtArgs(C.Structure):
    _fields_ = [('a', vp), ('sam', i64), ('sak', i64), ('b', vp), ('sbn', i64), ('sbk', i64), ('bias', vp),
                ('d', vp), ('sdm', i64), ('sdn', i64), ('m', i32), ('n', i32), ('k', i32), ('relu_a', i32),
                ('act', i32), ('batch', i32), ('a_bs', i64), ('d_bs', i64)]


class AttentionArgs(C.Structure):
    _fields_ = [('n_graphs', i32), ('hid', i32), ('heads', i32), ('max_nodes', i32), ('lut_size', i32),
                ('node_off', vp), ('mat_off', vp), ('qkv', vp), ('dtype', i32), ('pair', vp), ('lut', vp),
                ('out', vp), ('lse2', vp), ('total_nodes', i32)]


class LayerWeights(C.Structure):
    _fields_ = [('ln1_w', vp), ('ln1_b', vp), ('w_qkv', vp), ('w_out', vp), ('b_out', vp), ('ln2_w', vp),
                ('ln2_b', vp), ('w_ff1', vp), ('b_ff1', vp), ('w_ff2', vp), ('b_ff2', vp)]


class GraphormerArgs(C.Structure):
    _fields_ = [('hid', i32), ('heads', i32), ('layers', i32), ('dtype', i32), ('layers_host', C.POINTER(LayerWeights)),
                ('ln_w', vp), ('ln_b', vp),
                ('n_graphs', i32), ('total_nodes', i32), ('max_nodes', i32), ('lut_size', i32),
                ('node_off', vp), ('mat_off', vp), ('pair', vp), ('lut', vp), ('x', vp),
                ('h', vp), ('qkv', vp), ('ff', vp),
                ('dec_in', vp), ('dec_dtype', i32), ('dst_row', vp), ('emb_f32', vp), ('ln_counters', vp),
                ('h2', vp), ('tf32_x3', i32), ('skip_final_ln', i32)]


class GraphormerFusedArgs(C.Structure):
    _fields_ = [('hid', i32), ('heads', i32), ('layers', i32), ('w_qkv', vp), ('w_out', vp), ('w_ff1', vp),
                ('w_ff2', vp), ('layers_dev', vp), ('n_graphs', i32), ('total_nodes', i32), ('max_nodes', i32),
                ('lut_size', i32), ('node_off', vp), ('mat_off', vp), ('pair', vp), ('lut', vp), ('x', vp),
                ('ao', vp), ('qkv', vp), ('ff', vp), ('sync', vp), ('stop_after', i32), ('max_ctas', i32)]


class ScatterDesc(C.Structure):
    _fields_ = [('dst', vp), ('src', vp), ('numel', i64), ('chunk0', i64), ('t1', i32), ('t2', i32), ('t3', i32),
                ('so', i32), ('si', i32), ('ld', i32), ('ca', i32), ('ra', i32), ('kh_src', i32), ('kw_src', i32),
                ('cy', i32), ('cx', i32), ('scale', f32), ('mode', i32)] + \
               [(n_, C.c_uint32) for n_ in ('m_t1', 's_t1', 'm_t2', 's_t2', 'm_t3', 's_t3', 'm_so', 's_so', 'm_si', 's_si')] + \
               [('norm_slot', i32), ('reserved', i32)]


def fastdiv(d):
    """(mul, shift) with n // d == (n * mul >> 32) >> shift for all 0 <= n < 2**31 (mul == 0 encodes d == 1)."""
    d = int(d)
    if d <= 1:
        return 0, 0
    s = (d - 1).bit_length()
    mul = -(-(1 << (31 + s)) // d)
    assert mul < (1 << 32)
    return mul, s - 1


def fill_fastdiv(desc):
    """Fills the m_*/s_* fields of a ScatterDesc (ctypes) from its t1, t2, t3, so, si."""
    for f in ('t1', 't2', 't3', 'so', 'si'):
        m, s = fastdiv(getattr(desc, f))
        setattr(desc, 'm_' + f, m)
        setattr(desc, 's_' + f, s)
    return desc


class ScatterArgs(C.Structure):
    _fields_ = [('descs', vp), ('n_descs', i32), ('n_chunks', i64), ('chunk_desc', vp), ('norm_out', vp),
                ('n_norm_slots', i32)]


class ReluTransposeArgs(C.Structure):
    _fields_ = [('src', vp), ('ld', i64), ('src_bs', i64), ('dst', vp), ('dst_dtype', i32), ('rows', i32),
                ('cols', i32), ('batch', i32)]


class SeqOp(C.Structure):
    _fields_ = [('op', i32), ('lane', i32), ('args', vp)]


class SumsqArgs(C.Structure):
    _fields_ = [('ptrs', vp), ('numels', vp), ('n', i32), ('out', vp)]


# ---- training path (adjoint kernels) ----
EW_COPY, EW_GELU, EW_GELU_BWD, EW_RELU_BWD, EW_ADD = 0, 1, 2, 3, 4


class TransposeArgs(C.Structure):
    _fields_ = [('src', vp), ('src_dtype', i32), ('ld_src', i64), ('rows', i32), ('cols', i32), ('group', i32),
                ('group_stride', i32), ('dst', vp), ('dst_dtype', i32), ('ld_dst', i64), ('mul_gelu_grad', vp),
                ('mul_dtype', i32), ('copy_out', vp), ('copy_dtype', i32), ('colsum_out', vp)]


class ElementwiseArgs(C.Structure):
    _fields_ = [('op', i32), ('n', i64), ('a', vp), ('a_dtype', i32), ('b', vp), ('b_dtype', i32), ('out', vp),
                ('out_dtype', i32)]


class ColsumArgs(C.Structure):
    _fields_ = [('src', vp), ('src_dtype', i32), ('ld', i64), ('rows', i32), ('cols', i32), ('group', i32),
                ('group_stride', i32), ('dst', vp)]


class LayerNormBwdArgs(C.Structure):
    _fields_ = [('rows', i32), ('hid', i32), ('x', vp), ('gamma', vp), ('dy', vp), ('dy_dtype', i32), ('dy_row', vp),
                ('dx', vp), ('accumulate', i32), ('dgamma', vp), ('dbeta', vp)]


class AttentionBwdArgs(C.Structure):
    _fields_ = [('n_graphs', i32), ('hid', i32), ('heads', i32), ('max_nodes', i32), ('total_nodes', i32),
                ('lut_size', i32), ('node_off', vp), ('mat_off', vp), ('qkv', vp), ('out', vp), ('d_out', vp),
                ('dtype', i32), ('pair', vp), ('lut', vp), ('d_qkv', vp), ('d_lut', vp), ('lse', vp), ('delta', vp),
                ('fwd_lse2', vp), ('ds_total', vp)]


class LutBinArgs(C.Structure):
    _fields_ = [('n_graphs', i32), ('heads', i32), ('max_nodes', i32), ('lut_size', i32), ('node_off', vp),
                ('mat_off', vp), ('pair', vp), ('ds_total', vp), ('d_lut', vp)]


class ScatterBwdArgs(C.Structure):
    _fields_ = [('descs', vp), ('n_descs', i32), ('n_chunks', i64), ('chunk_desc', vp), ('grads', vp), ('d_src', vp)]


class NodeFeaturesBwdArgs(C.Structure):
    _fields_ = [('total_nodes', i32), ('hid', i32), ('op', vp), ('shape_idx', vp), ('deg_in', vp), ('deg_out', vp),
                ('dist0', vp), ('dx', vp), ('d_embed_op', vp), ('d_embed_ch', vp), ('d_embed_sp', vp),
                ('d_cent_in', vp), ('d_cent_out', vp), ('d_dist_embed', vp)]


class EdgeLutBwdArgs(C.Structure):
    _fields_ = [('hid', i32), ('heads', i32), ('vmax', i32), ('edge_embed', vp), ('w1', vp), ('b1', vp), ('w2', vp),
                ('d_lut', vp), ('workspace', vp), ('d_edge_embed', vp), ('d_w1', vp), ('d_b1', vp), ('d_w2', vp),
                ('d_b2', vp)]


class FcBwdArgs(C.Structure):
    _fields_ = [('problems', vp), ('n_problems', i32), ('max_m', i32), ('rowmap', vp), ('dh0', vp), ('dtype', i32),
                ('dec_in', vp), ('fc_w', vp), ('hid', i32), ('n_out', i32), ('grid_positions', i32), ('d_fc_w', vp),
                ('d_fc_b', vp), ('d_dec_in', vp)]


class ReluTransposeBwdArgs(C.Structure):
    _fields_ = [('src', vp), ('d_src', vp), ('ld', i64), ('src_bs', i64), ('d_rt', vp), ('rows', i32), ('cols', i32),
                ('batch', i32)]


class ExpandSeg(C.Structure):
    _fields_ = [('src_off', i64), ('row0', i32), ('rows', i32), ('ld', i32), ('group', i32), ('tile0', i32),
                ('tiles_c', i32)]


class ExpandArgs(C.Structure):
    _fields_ = [('segs', vp), ('n_segs', i32), ('n_tiles', i32), ('src', vp), ('group_stride', i32), ('x', vp),
                ('ld_x', i64), ('xt', vp), ('ld_xt', i64), ('dtype', i32), ('d_bias', vp)]


class AdamWArgs(C.Structure):
    _fields_ = [('params', vp), ('grads', vp), ('exp_avg', vp), ('exp_avg_sq', vp), ('offsets', vp), ('numels', vp),
                ('chunk0', vp), ('chunk_tensor', vp), ('n_chunks', i64), ('total', i64), ('lr', f32), ('beta1', f32),
                ('beta2', f32), ('eps', f32), ('weight_decay', f32), ('bias_correction1', f32),
                ('bias_correction2', f32), ('max_norm', f32), ('sumsq', vp), ('loss', vp), ('skipped', vp),
                ('range_lo', i64), ('range_hi', i64), ('chunk_begin', i64), ('sumsq_ready', i32)]


class SegnormArgs(C.Structure):
    _fields_ = [('src', vp), ('seg_off', vp), ('seg_numel', vp), ('chunk0', vp), ('chunk_seg', vp), ('n_chunks', i64),
                ('n_segs', i32), ('mode', i32), ('sumsq', vp), ('total', vp), ('grad', vp), ('gscale', vp),
                ('coef', f32)]


class MemsetArgs(C.Structure):
    _fields_ = [('ptr', vp), ('bytes', i64)]


class MemcpyArgs(C.Structure):
    _fields_ = [('dst', vp), ('src', vp), ('bytes', i64)]


class GraphormerTrainArgs(C.Structure):
    _fields_ = [('fwd', GraphormerArgs), ('xs', vp), ('xm', vp), ('h1', vp), ('qkv', vp), ('ao', vp), ('h2', vp),
                ('u', vp), ('g', vp), ('lse2', vp)]


class LayerWeightsT(C.Structure):
    _fields_ = [('w_qkv_t', vp), ('w_out_t', vp), ('w_ff1_t', vp), ('w_ff2_t', vp)]


class LayerGrads(C.Structure):
    _fields_ = [('ln1_w', vp), ('ln1_b', vp), ('w_qkv', vp), ('w_out', vp), ('b_out', vp), ('ln2_w', vp),
                ('ln2_b', vp), ('w_ff1', vp), ('b_ff1', vp), ('w_ff2', vp), ('b_ff2', vp)]


class GraphormerBwdArgs(C.Structure):
    _fields_ = [('saved', C.POINTER(GraphormerTrainArgs)), ('layers_t_host', C.POINTER(LayerWeightsT)),
                ('grads_host', C.POINTER(LayerGrads)), ('d_ln_w', vp), ('d_ln_b', vp), ('d_dec_in', vp),
                ('d_dec_dtype', i32), ('d_lut', vp), ('dx', vp), ('dxa', vp), ('dh', vp), ('dhf', vp), ('dqkv', vp),
                ('dff', vp), ('ta', vp), ('tb', vp), ('m_pad', i32), ('lse', vp), ('delta', vp), ('ds_total', vp),
                ('ds_total_bytes', i64)]


assert C.sizeof(ScatterDesc) == 136 and C.sizeof(GemmProblem) == 32

SYMBOLS = ['ghn3_last_error', 'ghn3_abi_version', 'ghn3_launch_count', 'ghn3_spd_bfs', 'ghn3_graph_derive',
           'ghn3_node_features', 'ghn3_edge_lut', 'ghn3_layernorm', 'ghn3_gemm', 'ghn3_gemm_simt', 'ghn3_attention',
           'ghn3_graphormer_stack', 'ghn3_graphormer_fused', 'ghn3_scatter', 'ghn3_sumsq', 'ghn3_relu_transpose', 'ghn3_convert_f32', 'ghn3_debug_gemm_trace', 'ghn3_run_sequence',
           'ghn3_graphormer_fused_sync_ints', 'ghn3_debug_fused_trace',
           'ghn3_set_programmatic_launch', 'ghn3_set_attention_tc_min', 'ghn3_set_persistent_ctas',
           'ghn3_sequence_capture', 'ghn3_sequence_launch', 'ghn3_sequence_destroy']
TRAIN_SYMBOLS = ['ghn3_transpose', 'ghn3_elementwise', 'ghn3_colsum', 'ghn3_layernorm_bwd', 'ghn3_attention_bwd',
                 'ghn3_scatter_bwd', 'ghn3_node_features_bwd', 'ghn3_edge_lut_bwd', 'ghn3_graphormer_train_fwd',
                 'ghn3_graphormer_bwd', 'ghn3_fc_bwd', 'ghn3_relu_transpose_bwd', 'ghn3_expand_cols', 'ghn3_adamw', 'ghn3_lut_bin', 'ghn3_segnorm']
SYMBOLS_ALL = SYMBOLS + TRAIN_SYMBOLS

_lib = None


def load(build_if_missing=True):
    """Loads (building first if necessary) the CUDA library. Raises if it cannot be had -- no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError('ghn3_b200: %s is missing; run `python -m ghn3_b200.build`' % LIB_PATH)
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for s in SYMBOLS_ALL:
        if not hasattr(lib, s):
            raise RuntimeError('ghn3_b200: %s does not export %s' % (LIB_PATH, s))
    lib.ghn3_last_error.restype = C.c_char_p
    lib.ghn3_launch_count.restype = C.c_int64
    for s in SYMBOLS[3:16] + TRAIN_SYMBOLS:
        getattr(lib, s).restype = C.c_int
        getattr(lib, s).argtypes = [C.c_void_p, C.c_void_p]
    lib.ghn3_graphormer_fused_sync_ints.restype = C.c_int64
    lib.ghn3_graphormer_fused_sync_ints.argtypes = [C.c_int32]
    lib.ghn3_run_sequence.restype = C.c_int
    lib.ghn3_run_sequence.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.ghn3_sequence_capture.restype = C.c_int
    lib.ghn3_sequence_capture.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.ghn3_sequence_launch.restype = C.c_int
    lib.ghn3_sequence_launch.argtypes = [C.c_void_p, C.c_void_p]
    lib.ghn3_sequence_destroy.restype = C.c_int
    lib.ghn3_sequence_destroy.argtypes = [C.c_void_p]
    lib.ghn3_convert_f32.restype = C.c_int
    lib.ghn3_convert_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().ghn3_last_error().decode(errors='replace')
        raise RuntimeError('ghn3_b200: %s failed (status %d): %s' % (what, rc, msg))


def call(name, args, stream):
    """Invokes int ghn3_<name>(const Args*, stream) and raises on a non-zero status."""
    lib = load()
    check(getattr(lib, 'ghn3_' + name)(C.byref(args), C.c_void_p(stream)), 'ghn3_' + name)


_pdl_state = [None]


def set_programmatic_launch(enabled):
    """Process-wide programmatic-dependent-launch switch (see include/ghn3_b200.h); no-op if unchanged."""
    enabled = bool(enabled)
    if _pdl_state[0] is not enabled:
        load().ghn3_set_programmatic_launch(C.c_int(int(enabled)))
        _pdl_state[0] = enabled


_cap_state = [None]


def set_persistent_ctas(ctas):
    """Process-wide grid cap of the persistent GEMM launches (see include/ghn3_b200.h); no-op if unchanged."""
    ctas = int(ctas or 0)
    if _cap_state[0] != ctas:
        load().ghn3_set_persistent_ctas(C.c_int(ctas))
        _cap_state[0] = ctas


def launch_count():
    return int(load().ghn3_launch_count())


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else t.data_ptr()
