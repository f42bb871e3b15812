"""ctypes binding of lib3dfeatnet_b200.so (include/feat3dnet_b200.h).  There is NO fallback: if the CUDA library
is missing or a call fails, this raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib3dfeatnet_b200.so")

PRECISION_IMAGES_CACHED = 0x100  # F3D_PRECISION_IMAGES_CACHED


def precision_sm_limit(n):
    """F3D_PRECISION_SM_LIMIT(n): the persistent tensor-core kernels of a forward call launch at most n CTAs (0 = all SMs)."""
    return (int(n) & 0xff) << 16

_c = ctypes
_vp, _i, _f, _sz, _ll = _c.c_void_p, _c.c_int, _c.c_float, _c.c_size_t, _c.c_longlong

# name -> (restype, argtypes); mirrors include/feat3dnet_b200.h one to one
SIGNATURES = {
    "f3d_version": (_i, []),
    "f3d_last_error_string": (_c.c_char_p, []),
    "f3d_launch_count": (_ll, []),
    "f3d_reset_launch_count": (None, []),
    "f3d_farthest_point_sample": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "f3d_farthest_point_sample_gather": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "f3d_farthest_point_sample_gather_ctas": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "f3d_gather_point": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "f3d_gather_point_grad": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_cumsum": (_i, [_i, _i, _vp, _vp, _vp]),
    "f3d_prob_sample": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "f3d_query_ball_point": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "f3d_query_ball_point_workspace_bytes": (_sz, [_i, _i]),
    "f3d_query_ball_point_ws": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_ball_grid_build": (_i, [_i, _i, _f, _vp, _vp, _sz, _vp]),
    "f3d_ball_grid_query": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_query_ball_point2": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "f3d_selection_sort": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "f3d_knn_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "f3d_knn_point": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_group_point": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "f3d_scatter_workspace_bytes": (_sz, [_ll]),
    "f3d_scatter_add_workspace_bytes": (_sz, [_i, _i, _ll]),
    "f3d_group_point_grad": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_packed_weights_floats": (_sz, [_i]),
    "f3d_packed_weights_num_blocks": (_i, []),
    "f3d_packed_weights_offsets": (_i, [_i, _vp, _vp]),
    "f3d_forward_workspace_bytes": (_sz, [_i, _i, _i]),
    "f3d_detector_forward": (_i, [_i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "f3d_descriptor_forward": (_i, [_i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "f3d_nms_workspace_bytes": (_sz, [_i, _i]),
    "f3d_nms": (_i, [_i, _i, _vp, _vp, _c.c_double, _c.c_double, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_group_local_frames": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _f, _i, _vp, _vp, _vp, _vp]),
    "f3d_group_local_frames_angle_grad": (_i, [_i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "f3d_conv_bn_train_workspace_bytes": (_sz, [_c.c_longlong, _i, _i]),
    "f3d_conv_bn_train_forward": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _i, _vp, _sz,
                                       _vp]),
    "f3d_conv_bn_train_forward_pooled": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _f, _vp, _i, _vp, _vp, _vp, _vp, _i,
                                              _vp, _sz, _vp]),
    "f3d_conv_bn_train_backward": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _i, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "f3d_conv_bn_train_forward_chain": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _f, _vp, _vp, _i, _vp, _vp, _vp,
                                             _vp, _vp, _i, _vp, _sz, _vp]),
    "f3d_conv_bn_train_backward_chain": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _i, _vp, _vp, _vp,
                                              _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "f3d_linear_workspace_bytes": (_sz, [_c.c_longlong, _i, _i]),
    "f3d_linear_forward": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_linear_backward": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_maxpool_samples_forward": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _vp, _vp]),
    "f3d_maxpool_samples_backward": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "f3d_triplet_loss_workspace_bytes": (_sz, [_i, _i]),
    "f3d_triplet_loss": (_i, [_i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_adam_step": (_i, [_i, _vp, _c.c_longlong, _f, _f, _f, _f, _c.c_longlong, _f, _vp, _vp]),
    "f3d_detector_heads_workspace_bytes": (_sz, [_i]),
    "f3d_detector_heads_forward": (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "f3d_detector_heads_backward": (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_pack_rows": (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "f3d_match_descriptors": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "f3d_ransac_workspace_bytes": (_sz, [_i, _i]),
    "f3d_ransac_fit_rt": (_i, [_i, _vp, _vp, _i, _vp, _f, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "f3d_rigid_fit": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "f3d_debug_umma_selftest": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "f3d_debug_umma_bench": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp]),
    "f3d_debug_wgrad_tc": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "f3d_debug_set_fuse_dz": (_i, [_i]),
    "f3d_debug_set_epilogue_pool": (_i, [_i]),
    "f3d_debug_set_row_walk": (_i, [_i]),
    "f3d_debug_set_lin_tc_phases": (_i, [_i]),
    "f3d_debug_set_lin_tc_pipe_min_k": (_i, [_i]),
    "f3d_debug_lin_tc_trace": (_i, [_vp]),
    "f3d_debug_lin_tc_weight_bytes": (_sz, [_i, _i]),
    "f3d_debug_lin_tc": (_i, [_c.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "f3d_detector_tc_weight_bytes": (_sz, []),
    "f3d_debug_set_timeline": (None, [_vp]),
    "f3d_debug_set_timeline_desc": (None, [_vp]),
    "f3d_debug_kernel_timer": (None, [_i]),
    "f3d_debug_kernel_timings": (_i, [_i, _vp, _vp, _vp]),
}

_LIB = None


class F3DError(RuntimeError):
    pass


def lib():
    """Load the C-ABI library (once).  Raises if it has not been built -- there is no CPU path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise F3DError(
                "lib3dfeatnet_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "This package has no CPU or PyTorch fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = lib().f3d_last_error_string().decode("utf-8", "replace")
        if rc in (-1,):
            raise ValueError("%s: %s" % (what, msg))
        raise F3DError("%s failed (code %d): %s" % (what, rc, msg))


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def kernel_timings(max_records=512):
    """[(kernel name, ms, algorithmic units)] recorded since f3d_debug_kernel_timer(1) (measurement aid)."""
    names = ctypes.create_string_buffer(max_records * 48)
    ms = (ctypes.c_float * max_records)()
    units = (ctypes.c_double * max_records)()
    n = lib().f3d_debug_kernel_timings(max_records, ctypes.cast(names, ctypes.c_void_p), ctypes.cast(ms, ctypes.c_void_p),
                                       ctypes.cast(units, ctypes.c_void_p))
    return [(names.raw[i * 48:(i + 1) * 48].split(b"\0", 1)[0].decode(), float(ms[i]), float(units[i])) for i in range(n)]


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise F3DError("3dfeatnet_b200 ops need CUDA tensors (got a %s tensor): there is no CPU path" % t.device)
