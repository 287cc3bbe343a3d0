"""ctypes binding of ``libgnndelete_b200.so`` (C ABI in ``include/gnndelete_b200.h``).

There is deliberately no fallback: if the library is missing, or a tensor handed
to a kernel is not a contiguous CUDA tensor of the expected dtype, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgnndelete_b200.so')

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class CsrStruct(C.Structure):
    """``gd_csr_t``"""
    _fields_ = [
        ('num_rows', _i64), ('nnz', _i64), ('rowptr', _vp), ('col', _vp),
        ('seg_len', _i32), ('num_heavy', _i32), ('num_seg', _i32),
        ('heavy_row', _vp), ('heavy_seg_beg', _vp), ('heavy_nseg', _vp),
        ('seg_row', _vp), ('seg_beg', _vp), ('seg_heavy', _vp), ('heavy_ticket', _vp), ('row_perm', _vp),
        ('grp_row', _vp), ('num_grp', _i32),
    ]


_csr_p = C.POINTER(CsrStruct)


class BplanStruct(C.Structure):
    """``gd_spmm_bplan_t``"""
    _fields_ = [
        ('num_rows', _i64), ('num_batches', _i64), ('num_workers', _i32), ('batches_per_worker', _i32),
        ('desc', _vp), ('colp', _vp), ('num_split', _i32), ('num_piece', _i32),
        ('piece_split', _vp), ('split_row', _vp), ('split_piece_beg', _vp), ('split_npiece', _vp), ('split_ticket', _vp),
    ]


_bplan_p = C.POINTER(BplanStruct)

# name -> (restype, argtypes); must list every symbol the header declares
SIGNATURES = {
    'gd_version': (C.c_int, []),
    'gd_last_error': (C.c_char_p, []),
    'gd_launch_count': (C.c_longlong, []),
    'gd_csr_workspace_bytes': (_sz, [_i64, _i64]),
    'gd_csr_from_coo': (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gd_invert_perm': (C.c_int, [_vp, _i64, _vp, _vp]),
    'gd_gcn_dinv': (C.c_int, [_vp, _i64, _vp, _vp]),
    'gd_spmm_plan_build': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'gd_spmm': (C.c_int, [_csr_p, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _i64, _vp, _vp]),
    'gd_spmm_acc': (C.c_int, [_csr_p, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _i64, _vp, _i32, _vp]),
    'gd_spmm_bplan_workspace_bytes': (_sz, [_i64]),
    'gd_spmm_bplan_count': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    'gd_spmm_bplan_fill': (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'gd_spmm_batched_workers': (_i32, [_i32, _i32]),
    'gd_spmm_batched': (C.c_int, [_bplan_p, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _i64, _vp, _i32, _vp]),
    'gd_spmm_batched_tail': (C.c_int, [_bplan_p, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _i64, _vp, _i32, _vp]),
    'gd_gemm_rows_tc_batch': (C.c_int, [_vp, _i64, _i64, _i32, _vp, _i64, _i32, _i32, _vp, _i64, _i64, _i32, _vp]),
    'gd_gat_scores': (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    'gd_gat_scratch_floats': (_sz, [_csr_p, _i32]),
    'gd_gat_fwd': (C.c_int, [_csr_p, _vp, _i64, _i32, _vp, _vp, _vp, _f32, _vp, _i64, _vp, _vp, _vp, _vp]),
    'gd_gat_bwd_dst': (C.c_int, [_csr_p, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _f32,
                                 _vp, _vp, _vp, _vp, _vp]),
    'gd_gat_bwd_src': (C.c_int, [_csr_p, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'gd_rgcn_norm': (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    'gd_rgcn_conv': (C.c_int, [_csr_p, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    'gd_rgcn_edge_tile_rows': (_i32, [_i32, _i32]),
    'gd_rgcn_edge_conv': (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp]),
    'gd_rgcn_edge_reduce': (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp]),
    'gd_permute_f32': (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    'gd_gather_rows': (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i32, _vp, _i64, _vp, _vp]),
    'gd_gemm_rows': (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _i64, _vp]),
    'gd_gemm_rows_tc_supported': (C.c_int, [_i32, _i32, _i64, _i64]),
    'gd_gemm_rows_tc': (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _i64,
                                  _vp, _vp, _vp]),
    'gd_gemm_tn_workspace_bytes': (_sz, [_i64, _i32, _i32]),
    'gd_gemm_tn_rows': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    'gd_gemm_tn_rows_tc_supported': (C.c_int, [_i32, _i32, _i64, _i64]),
    'gd_gemm_tn_tc_workspace_bytes': (_sz, [_i32, _i32]),
    'gd_gemm_tn_rows_tc': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    'gd_gemm_dxdw_tc_supported': (C.c_int, [_i32, _i32, _i32, _i64, _i64]),
    'gd_gemm_dxdw_tc': (C.c_int, [_vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _sz, _vp]),
    'gd_copy_rows': (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    'gd_copy_rows_scaled': (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64, _vp]),
    'gd_relu_fwd': (C.c_int, [_vp, _i64, _vp, _vp]),
    'gd_relu_bwd': (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    'gd_edge_loss_workspace_bytes': (_sz, [_i64]),
    'gd_edge_loss_fwd': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gd_edge_loss_fwd_part': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _i64, _i64,
                                        _i64, _i64, _vp, _sz, _vp]),
    'gd_node_loss_workers': (_i32, [_i32, _i32]),
    'gd_node_loss_workspace_bytes': (_sz, [_i32]),
    'gd_spmm_batched_bf16_workers': (_i32, [_i32, _i32]),
    'gd_spmm_batched_bf16': (C.c_int, [_bplan_p, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _i64, _vp, _i32, _vp]),
    'gd_cast_bf16': (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _i64, _vp]),
    'gd_move_f32': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    'gd_node_loss_fwd_bwd': (C.c_int, [_bplan_p, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _vp, _i32, _i64, _f32,
                                       _vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gd_pair_scatter_add': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    'gd_dec_items_workspace_bytes': (_sz, []),
    'gd_dec_items_fwd': (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _i64, _i64, _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gd_dense_ni_workspace_bytes': (_sz, [_i64]),
    'gd_dense_ni_fwd_bwd': (C.c_int, [_vp, _i64, _i32, _i64, _vp, _i64, _vp, _f32, _vp, _i64, _vp, _vp, _sz, _vp]),
    'gd_row_mse_workspace_bytes': (_sz, [_i64]),
    'gd_row_mse_fwd_bwd': (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i64, _vp, _vp, _f32, _f32, _f32, _f32, _vp, _i64, _vp, _vp, _sz, _vp]),
    'gd_dense_ni_tc_supported': (C.c_int, [_i32]),
    'gd_dense_ni_tc_target_bytes': (_sz, [_i64]),
    'gd_dense_ni_tc_pack_target': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    'gd_dense_ni_tc_workspace_bytes': (_sz, [_i64]),
    'gd_dense_ni_tc_fwd_bwd': (C.c_int, [_vp, _i64, _i64, _vp, _f32, _vp, _i64, _vp, _vp, _sz, _vp]),
    'gd_add_rows': (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    'gd_pair_decode': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    'gd_adam_step': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp]),
    'gd_khop_workspace_bytes': (_sz, [_i64]),
    'gd_khop_masks': (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gd_to_undirected_workspace_bytes': (_sz, [_i64]),
    'gd_to_undirected': (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(nvcc, sm_100a). gnndelete_b200 has no CPU or PyTorch fallback for its kernels.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().gd_last_error().decode()


def call(name: str, *args):
    """Invoke an ``int``-returning entry point; raise with ``gd_last_error`` on failure."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {last_error()}')


_DTYPES = {'f32': torch.float32, 'i32': torch.int32, 'i64': torch.int64, 'u8': torch.uint8}


def ptr(t, kind=None):
    """Device pointer of a contiguous CUDA tensor (``None`` -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('gnndelete_b200 kernels take CUDA tensors only (no CPU fallback)')
    if not t.is_contiguous():
        raise RuntimeError('non-contiguous tensor passed to a gnndelete_b200 kernel')
    if kind is not None and t.dtype != _DTYPES[kind]:
        raise RuntimeError(f'expected {kind} tensor, got {t.dtype}')
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
