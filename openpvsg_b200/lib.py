"""ctypes binding of libpvsg_sm100.so (the C ABI declared in include/pvsg.h).

The library is built in-tree by ``build()`` (``make`` in openpvsg_b200/csrc, nvcc
-gencode arch=compute_100a,code=sm_100a).  There is NO fallback: if the shared object is
missing or a call returns an error code, a ``PvsgError`` is raised.
"""
import ctypes
import os
import subprocess
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libpvsg_sm100.so')
CSRC = os.path.join(HERE, 'csrc')

P = c_void_p
I = c_int
L = c_int64
F = c_float
D = c_double

# name -> (restype, argtypes); mirrors include/pvsg.h one to one
SIGNATURES = {
    'pvsg_version': (I, []),
    'pvsg_error_string': (c_char_p, [I]),
    'pvsg_device_info': (I, [I, P, P, P]),
    'pvsg_create': (I, [I, P]),
    'pvsg_destroy': (I, [P]),
    'pvsg_handle_info': (I, [P, P, P, P, P]),
    'pvsg_workspace': (I, [P, L, P]),
    'pvsg_linear': (I, [P, P, P, P, P, P, L, L, L, L, L, L, L, I, L, L, L, L, P]),
    'pvsg_conv2d_nhwc': (I, [P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, P]),
    'pvsg_split_bf16': (I, [P, P, P, P, L, P]),
    'pvsg_stem7x7s2_pack': (I, [P, P, P, I, I, I, I, P]),
    'pvsg_im2col_split': (I, [P, P, P, I, I, I, I, I, I, I, I, I, P]),
    'pvsg_linear_tc': (I, [P, P, L, P, P, L, P, P, L, P, P, P, P, P, L, L, L, L, I, P, P, P, P]),
    'pvsg_linear_tc_batched': (I, [P, P, L, L, P, P, L, L, P, P, P, L, I, L, L, L, P]),
    'pvsg_conv2d_tc': (I, [P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, P, P, P]),
    'pvsg_maxpool3x3s2_nhwc': (I, [P, P, I, I, I, I, P]),
    'pvsg_nchw_to_nhwc': (I, [P, P, I, I, I, I, P]),
    'pvsg_nhwc_to_nchw': (I, [P, P, I, I, I, I, P]),
    'pvsg_layernorm': (I, [P, P, P, P, L, I, F, P]),
    'pvsg_layernorm_split': (I, [P, P, P, P, P, P, L, I, F, P]),
    'pvsg_layernorm_split2': (I, [P, P, P, P, P, P, P, P, P, L, I, F, P]),
    'pvsg_groupnorm_nhwc': (I, [P, P, P, P, P, I, L, I, I, F, I, P]),
    'pvsg_groupnorm_nhwc_split': (I, [P, P, P, P, P, P, P, I, L, I, I, F, I, P]),
    'pvsg_add_rowvec': (I, [P, P, P, L, I, P]),
    'pvsg_bilinear_resize_nhwc': (I, [P, P, I, I, I, I, I, I, I, P]),
    'pvsg_bilinear_resize_nhwc_ex': (I, [P, L, P, P, P, I, I, I, I, I, I, I, P]),
    'pvsg_bilinear_resize_scaled': (I, [P, P, I, I, I, I, I, I, F, F, P]),
    'pvsg_sine_pe': (I, [P, P, P, P, I, I, I, I, F, F, P]),
    'pvsg_msda_forward': (I, [P, P, P, P, P, P, I, L, L, I, I, I, I, P]),
    'pvsg_msda_fused_forward': (I, [P, P, P, P, P, P, I, L, L, I, I, I, I, P]),
    'pvsg_msda_fused_forward_split': (I, [P, P, P, P, P, P, P, P, I, L, L, I, I, I, I, P]),
    'pvsg_attention_workspace_bytes': (L, [I, I, I, I, I]),
    'pvsg_attention': (I, [P, P, P, P, P, P, P, I, I, I, I, I, L, L, L, L, L, L, L, L, F, P]),
    'pvsg_attention_tc_workspace_bytes': (L, [I, I, I, I, I]),
    'pvsg_attention_tc': (I, [P, P, P, P, P, P, P, P, P, I, I, I, I, I, L, L, L, L, L, L, L, L, F, P]),
    'pvsg_attention_t5_workspace_bytes': (L, [I, I, I, I, I]),
    'pvsg_attention_t5': (I, [P, P, P, P, P, P, P, P, P, I, I, I, I, I, L, L, L, L, L, L, L, L, F, P]),
    'pvsg_attention_t5_lse': (I, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, L, L, L, L, L, L, L, L, F, P]),
    'pvsg_mask_logits': (I, [P, P, P, P, P, I, I, L, I, P]),
    'pvsg_panoptic_fuse': (I, [P, P, I, I, I, I, I, I, I, I, I, I, I, F, D, I, I, P, P, P, P, P, P]),
    'pvsg_instance_masks': (I, [P, P, I, I, I, I, I, I, I, I, I, P, P, P, P]),
    'pvsg_instance_select': (I, [P, I, I, I, P, P, P, P]),
    'pvsg_instance_finalize': (I, [P, P, P, P, P, I, I, I, P, P, P, P, P]),
    'pvsg_panoptic_fuse_batched': (I, [P, P, I, I, I, I, I, I, I, I, I, I, I, I, F, D, I, I, P, P, P, P, P, P]),
    'pvsg_instance_select_batched': (I, [P, I, I, I, I, P, P, P, P]),
    'pvsg_instance_masks_batched': (I, [P, P, I, I, I, I, I, I, I, I, I, I, I, P, P, P, P]),
    'pvsg_instance_finalize_batched': (I, [P, P, P, P, P, I, I, I, I, P, P, P, P, P]),
    'pvsg_rle_events': (I, [P, P, I, I, I, I, P, P, P, P, I, P]),
    'pvsg_rle_strings_host': (L, [P, P, L, I, ctypes.c_uint32, P, L, P]),
    'pvsg_window_attention': (I, [P, P, P, P, P, P, I, I, I, I, I, I, I, P]),
    'pvsg_patch_merge_ln': (I, [P, P, P, P, I, I, I, I, F, P]),
    'pvsg_tube_overlap': (I, [P, P, P, I, I, I, I, I, P, P]),
    'pvsg_reconsdot_workspace_bytes': (L, [I, I, I, I, I]),
    'pvsg_reconsdot': (I, [P, P, P, P, I, I, I, I, I, F, P]),
    'pvsg_lap_assign': (I, [P, I, I, D, P, P, P]),
    'pvsg_layernorm_backward': (I, [P, P, P, P, P, P, L, I, F, P]),
    'pvsg_relu_backward': (I, [P, P, P, L, P]),
    'pvsg_transpose_split': (I, [P, P, P, P, L, I, I, I, L, L, L, L, P]),
    'pvsg_maxpool3x3s2_nhwc_backward': (I, [P, P, P, I, I, I, I, P]),
    'pvsg_groupnorm_nhwc_backward': (I, [P, P, P, P, P, P, P, P, I, L, I, I, F, I, P]),
    'pvsg_bilinear_resize_nhwc_backward': (I, [P, P, I, I, I, I, I, I, P]),
    'pvsg_msda_proj_expand': (I, [P, P, P, P, P, I, L, I, I, I, P]),
    'pvsg_msda_proj_backward': (I, [P, P, P, P, P, I, L, I, I, I, P]),
    'pvsg_colsum': (I, [P, P, L, I, L, P]),
    'pvsg_attention_train_forward': (I, [P, P, P, P, P, P, P, I, I, I, I, I, L, L, L, L, L, L, L, L, F, P]),
    'pvsg_attention_train_backward': (I, [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, L, L, L, L, L, L, L, L, F, P]),
    'pvsg_cosine_chain_cost': (I, [P, P, I, I, I, P]),
    'pvsg_lap_square_batched': (I, [P, I, I, P, P, P]),
    'pvsg_perm_chain': (I, [P, P, I, I, P]),
    'pvsg_point_sample': (I, [P, P, P, I, I, I, I, I, P]),
    'pvsg_point_sample_backward': (I, [P, P, P, I, I, I, I, I, P]),
    'pvsg_mask_point_losses': (I, [P, P, I, I, F, F, F, P, P, P]),
    'pvsg_weighted_ce': (I, [P, P, P, P, I, I, F, P, P, P]),
    'pvsg_mask_match_cost': (I, [P, P, P, P, I, I, I, I, F, F, F, F, P, P]),
    'pvsg_msda_backward': (I, [P, P, P, P, P, P, P, P, P, I, L, L, I, I, I, I, P]),
    'pvsg_max_over_time': (I, [P, P, I, I, I, P]),
    'pvsg_temporal_fir': (I, [P, P, P, I, I, I, I, P]),
    'pvsg_temporal_unfold': (I, [P, P, I, I, I, I, P]),
    'pvsg_pair_proposal': (I, [P, P, P, P, P, I, I, P]),
    'pvsg_top_pairs': (I, [P, I, I, P, P, P]),
    'pvsg_gather_pairs': (I, [P, P, P, P, P, I, I, I, P]),
}


class PvsgError(RuntimeError):
    pass


_lib = None


def build(verbose=False):
    """Compile every CUDA source for sm_100a into openpvsg_b200/libpvsg_sm100.so."""
    res = subprocess.run(['make', '-j8', '-C', CSRC], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise PvsgError('building libpvsg_sm100.so failed')
    return LIB_PATH


def load():
    """dlopen the library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PvsgError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                        '(there is no CPU / PyTorch fallback for the pvsg kernels)')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.pvsg_version() != 100:
        raise PvsgError('libpvsg_sm100.so version mismatch')
    _lib = lib
    return lib


class Handle:
    """pvsg_create / pvsg_destroy (include/pvsg.h): pins a device, configures every kernel's opt-in attributes on it
    before any capture, owns a grow-only scratch block."""

    def __init__(self, device):
        self._h = ctypes.c_void_p()
        code = load().pvsg_create(int(device), ctypes.byref(self._h))
        if code != 0:
            raise PvsgError(f'pvsg_create({device}) failed: {load().pvsg_error_string(code).decode()} ({code})')

    def info(self):
        dev, sms, smem, ws = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
        code = load().pvsg_handle_info(self._h, ctypes.byref(dev), ctypes.byref(sms), ctypes.byref(smem), ctypes.byref(ws))
        if code != 0:
            raise PvsgError(f'pvsg_handle_info failed ({code})')
        return {'device': dev.value, 'sm_count': sms.value, 'smem_optin_bytes': smem.value, 'workspace_bytes': ws.value}

    def workspace(self, nbytes):
        ptr = ctypes.c_void_p()
        code = load().pvsg_workspace(self._h, int(nbytes), ctypes.byref(ptr))
        if code != 0:
            raise PvsgError(f'pvsg_workspace({nbytes}) failed: {load().pvsg_error_string(code).decode()} ({code})')
        return ptr.value

    def close(self):
        if self._h:
            load().pvsg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}


def handle(device):
    """The process-wide handle of a device index (created on first use; engine.FrameRunner and the relation head's
    graph capture call this before capturing)."""
    device = int(device)
    if device not in _handles:
        _handles[device] = Handle(device)
    return _handles[device]


# kernels launched per successful C-ABI call (lower bounds; used for bench.py's gpu_launches)
KERNELS_PER_CALL = {'pvsg_groupnorm_nhwc': 2, 'pvsg_groupnorm_nhwc_split': 2, 'pvsg_panoptic_fuse': 4, 'pvsg_instance_masks': 3,
                    'pvsg_panoptic_fuse_batched': 4, 'pvsg_instance_masks_batched': 3, 'pvsg_rle_events': 3, 'pvsg_tube_overlap': 1, 'pvsg_reconsdot': 7}
ATTN_IMPL = [os.environ.get('PVSG_ATTN_IMPL', 't5')]   # 't5' = tcgen05 / TMEM kernel, 'mma' = mma.sync kernel (debug switch)
launch_count = [0]


def check(code, what):
    launch_count[0] += KERNELS_PER_CALL.get(what, 1)
    if code != 0:
        msg = load().pvsg_error_string(code).decode()
        raise PvsgError(f'{what} failed: {msg} ({code})')
