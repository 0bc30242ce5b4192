"""ctypes binding of libcalciumgan_b200.so (include/calciumgan_b200.h).

There is no CPU fallback: a missing library or a box without a CUDA device raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CG_LIB') or os.path.join(_HERE, 'libcalciumgan_b200.so')   # CG_LIB: A/B builds in tools/

GENERATOR, DISCRIMINATOR = 0, 1
FP32, BF16 = 0, 1
NUM_SCALARS = 16
S_DIS_LOSS, S_GP, S_REAL_LOSS, S_FAKE_LOSS, S_GEN_LOSS = 0, 1, 2, 3, 4
S_MET_MIN, S_MET_MAX, S_MET_MEAN, S_MET_STD = 5, 6, 7, 8
FLAG_NO_UPDATE, FLAG_NO_SYNC, FLAG_SAME_REAL, FLAG_NO_FAKE32, FLAG_GEN_PREFETCHED = 1, 2, 4, 8, 16
DEBUG_NO_PS_FUSE, DEBUG_NO_PS_BWD_FUSE, DEBUG_NO_GHEAD, DEBUG_NO_ADAM_FUSE = 1, 2, 4, 8
BUF_X, BUF_H, BUF_DA, BUF_HG, BUF_AG, BUF_DAG = 0, 1, 2, 3, 4, 5


class CgConfig(C.Structure):
  _fields_ = [
      ('seq_len', C.c_int32), ('channels', C.c_int32), ('noise_dim', C.c_int32),
      ('num_units', C.c_int32), ('kernel_size', C.c_int32), ('strides', C.c_int32),
      ('phase_m', C.c_int32), ('layer_norm', C.c_int32), ('normalize', C.c_int32),
      ('max_batch', C.c_int32), ('n_critic', C.c_int32), ('precision', C.c_int32),
      ('gp_lambda', C.c_float), ('learning_rate', C.c_float), ('signals_min', C.c_float),
      ('signals_max', C.c_float), ('world_size', C.c_int32), ('rank', C.c_int32),
      ('force_simt', C.c_int32), ('debug_flags', C.c_int32), ('reserved', C.c_int32 * 6),
  ]


# every symbol include/calciumgan_b200.h declares: name -> (restype, argtypes)
_P, _I, _I64, _F = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_float)
_I32P = C.POINTER(C.c_int32)
SIGNATURES = {
    'cg_version': (_I, []),
    'cg_last_error': (C.c_char_p, []),
    'cg_create': (_I, [C.POINTER(CgConfig), C.POINTER(_P)]),
    'cg_destroy': (None, [_P]),
    'cg_set_stream': (_I, [_P, _P]),
    'cg_synchronize': (_I, [_P]),
    'cg_num_params': (_I64, [_P, _I]),
    'cg_num_tensors': (_I, [_P, _I]),
    'cg_tensor_info': (_I, [_P, _I, _I, C.POINTER(_I64), C.POINTER(_I), C.POINTER(_I64)]),
    'cg_set_weights': (_I, [_P, _I, _P]),
    'cg_get_weights': (_I, [_P, _I, _P]),
    'cg_init_weights': (_I, [_P, C.c_uint64]),
    'cg_get_grads': (_I, [_P, _I, _P]),
    'cg_grad_ptr': (_P, [_P, _I]),
    'cg_num_buckets': (_I, [_P, _I]),
    'cg_bucket_info': (_I, [_P, _I, _I, C.POINTER(_I64), C.POINTER(_I64)]),
    'cg_stream_wait_bucket': (_I, [_P, _I, _I, _P]),
    'cg_set_grad_buffer': (_I, [_P, _I, _P]),
    'cg_reduce_peer_grads': (_I, [_P, _I, C.POINTER(_P), _I, _P]),
    'cg_set_reduced_buffer': (_I, [_P, _I, _P]),
    'cg_peer_reduce_scatter': (_I, [_P, _I, C.POINTER(_P), _I, _I, _P]),
    'cg_peer_all_gather': (_I, [_P, _I, C.POINTER(_P), _I, _I, _P]),
    'cg_apply_update_reduced': (_I, [_P, _I]),
    'cg_reduced_grad_ptr': (_P, [_P, _I]),
    'cg_set_grads': (_I, [_P, _I, _P]),
    'cg_skipped_updates': (_I64, [_P, _I]),
    'cg_get_opt_state': (_I, [_P, _I, _P, _P, C.POINTER(_I64)]),
    'cg_set_opt_state': (_I, [_P, _I, _P, _P, _I64]),
    'cg_seed': (_I, [_P, C.c_uint64]),
    'cg_critic_step': (_I, [_P, _P, _I, _P, _P, _I32P, _I, _F]),
    'cg_generator_step': (_I, [_P, _P, _I, _P, _I32P, _I, _F]),
    'cg_prefetch_generator': (_I, [_P, _P, _I, _P, _P, _I, _I]),
    'cg_apply_update': (_I, [_P, _I]),
    'cg_train_step': (_I, [_P, _P, _I, _P, _P, _I32P, _F]),
    'cg_validate': (_I, [_P, _P, _I, _P, _P, _I32P, _P, _F]),
    'cg_gather_rows': (_I, [_P, _P, _I64, _P, _I, _I64, _P]),
    'cg_metrics': (_I, [_P, _P, _P, _I, _F]),
    'cg_generate': (_I, [_P, _P, _I, _I, _P]),
    'cg_debug_critic_forward': (_I, [_P, _P, _I, _I32P, _P]),
    'cg_debug_gp': (_I, [_P, _P, _I, _I32P, _P, _P]),
    'cg_gp_gradient': (_I, [_P, _P, _I, _I32P, _I, _F]),
    'cg_debug_layer': (_I, [_P, _I, _I, _I, _P, _P, _I, _P]),
    'cg_debug_phase_shuffle': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'cg_debug_read': (_I, [_P, _I, _I, _I, _P]),
    'cg_debug_buffer_shape': (_I, [_P, _I, _I, C.POINTER(_I64), C.POINTER(_I64)]),
    'cg_debug_last_draws': (_I, [_P, _P, _I64, _P, _I64, _I32P, _I]),
    'cg_debug_dgrad_ps': (_I, [_P, _I, _P, _P, _I, _I, _I32P, _P]),
    'cg_phase_shuffle_index': (_I, [_I, _I, _I32P]),
    'cg_phase_shuffle_scatter_index': (_I, [_I, _I, _I32P, _I32P]),
    'cg_phase_shuffle_adjoint_plan': (_I, [_I, _I, _I32P, _I32P, _I32P, _I32P]),
    'cg_fake_ptr': (_P, [_P]),
    'cg_scores_ptr': (_P, [_P]),
    'cg_scalars_ptr': (_P, [_P]),
    'cg_launch_count': (_I64, [_P]),
    'cg_tc_launch_count': (_I64, [_P]),
    'cg_device_bytes': (_I64, [_P]),
    'cg_profile': (_I, [_P, _I]),
    'cg_profile_report': (_I, [_P, C.POINTER(C.c_double)]),
    'cg_profile_report_text': (_I, [_P, C.c_char_p, _I]),
    'cg_bench_layer': (_I, [_P, _I, _I, _I, _I, _I, _F, C.POINTER(C.c_double)]),
}

_lib = None


def load():
  """Load the shared library (built in-tree by __graft_entry__.build / csrc/build.sh)."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        'calciumgan_b200: %s is missing. Build it with `python -c "import __graft_entry__ as g; '
        'g.build()"` (nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
  lib = C.CDLL(LIB_PATH)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)   # AttributeError if the header and the library disagree
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


class CgError(RuntimeError):
  pass


def check(rc):
  if rc != 0:
    raise CgError(load().cg_last_error().decode('utf-8', 'replace'))
