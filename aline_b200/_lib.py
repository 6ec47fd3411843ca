"""ctypes binding of ``libaline_b200.so`` (the C ABI declared in ``include/aline_b200.h``).

There is no CPU or pure-PyTorch fallback: if the library is missing, or a tensor
is not a contiguous fp32 CUDA tensor, the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALINE_B200_LIB") or os.path.join(_HERE, "lib", "libaline_b200.so")

TASK_LOCATION, TASK_CES, TASK_PSYCHOMETRIC = 0, 1, 2
GP_KERNELS = ("rbf", "matern12", "matern32", "matern52")


class AlineLik(Structure):
    """``struct aline_lik`` (include/aline_b200.h)."""
    _fields_ = [("task", c_int32), ("dim_x", c_int32), ("K", c_int32), ("dim_theta", c_int32),
                ("c0", c_float), ("c1", c_float), ("c2", c_float), ("c3", c_float)]


class AlineError(RuntimeError):
    pass


_lib = None


def _sig(fn, restype, *argtypes):
    fn.restype = restype
    fn.argtypes = list(argtypes)


def lib():
    """Load the shared library once.  Raises if it has not been built (``python -m aline_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AlineError(f"{LIB_PATH} not found: build it with `python -m aline_b200.build` "
                         "(the B200 path has no CPU / PyTorch fallback)")
    L = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    _sig(L.aline_abi_version, c_int32)
    _sig(L.aline_last_error, c_char_p)
    _sig(L.aline_kernel_launches, c_uint64)
    _sig(L.aline_spce_scratch_bytes, c_size_t, c_int32, c_int32)
    _sig(L.aline_spce_pass_len, c_int32, POINTER(AlineLik), c_int32)
    _sig(L.aline_spce_history, c_int32, POINTER(AlineLik), P, P, P, P, c_int64, c_int32, c_int32, c_int32,
         P, P, P, P, P, c_size_t, P)
    _sig(L.aline_spce_history_ex, c_int32, POINTER(AlineLik), P, P, P, P, c_int64, c_int32, c_int32, c_int32,
         P, P, P, P, P, c_size_t, c_int32, P)
    _sig(L.aline_spce_step, c_int32, POINTER(AlineLik), P, P, P, P, c_int64, c_int32, c_int32,
         P, P, P, P, P, c_size_t, P)
    _sig(L.aline_prior_sample, c_int32, P, ctypes.c_uint64, c_int64, c_int64, c_int32, P, P)
    _sig(L.aline_sample_batch, c_int32, POINTER(AlineLik), P, ctypes.c_uint64, c_int64, c_int32, c_int32, c_float,
         c_float, c_float, P, P, P, P)
    _sig(L.aline_spce_device_prior_seq_rows, c_int64, POINTER(AlineLik), c_int64, c_int32)
    _sig(L.aline_spce_history_device_prior, c_int32, POINTER(AlineLik), P, ctypes.c_uint64, c_int64, P, P, P, P, c_int64,
         c_int32, c_int32, P, P, P, P, P, c_size_t, P)
    _sig(L.aline_lse_combine, c_int32, P, P, P, c_int32, c_int64, P, P, P)
    _sig(L.aline_log_likelihood, c_int32, POINTER(AlineLik), P, P, P, P, c_int64, c_int32, P, P, c_size_t, P)
    _sig(L.aline_model_param_count, c_uint64, P)
    _sig(L.aline_set_option, c_int32, c_char_p, c_int32)
    _sig(L.aline_embed_queries, c_int32, P, P, c_int32, c_int32, P, P)
    _sig(L.aline_embed_queries_ex, c_int32, P, P, c_int32, c_int32, P, P, P)
    _sig(L.aline_ctx_stack, c_int32, P, P, P, c_int32, c_int32, c_int32, P, c_int32, P, P, c_int32, P, P, c_int32, P)
    _sig(L.aline_ctx_stack_ex, c_int32, P, P, P, c_int32, c_int32, c_int32, P, c_int32, P, P, c_int32, P, P, P, c_int32,
         P)
    _sig(L.aline_value_head, c_int32, P, c_int32, c_int32, c_int32, c_int32, P, P, P, P, P, P)
    _sig(L.aline_query_stream, c_int32, P, P, P, c_int32, c_int32, P, c_int32, c_int32, c_float, P, P, P)
    _sig(L.aline_select, c_int32, P, P, c_int32, c_int32, P, P, c_int32, c_int32, P, P, c_int32, c_int32, P, c_int32,
         P, c_int32, P, P, P)
    _sig(L.aline_select_sample, c_int32, P, P, c_int32, c_int32, P, P, c_int32, c_int32, P, P, c_int32, c_int32, P,
         c_int32, P, c_int32, P, P, c_uint64, c_int32, P)
    _sig(L.aline_gmm_head, c_int32, P, P, c_int64, P, P, P, P)
    _sig(L.aline_gmm_variance, c_int32, P, P, P, c_int64, c_int32, c_int64, P, P)
    _sig(L.aline_gmm_head_variance, c_int32, P, P, c_int64, P, P)
    _sig(L.aline_gmm_log_likelihood, c_int32, P, P, P, P, c_int64, c_int32, P, P)
    _sig(L.aline_move_selected, c_int32, P, P, P, c_int32, c_int32, c_int32, c_int32, P, P, P)
    _sig(L.aline_rollout, c_int32, P, P, P, P, P, P, P, c_int32, c_int32, c_int32, c_int32, P, c_int32, P, c_int32,
         P, c_int32, P, c_int32, P, P, P, P, P, P)
    _sig(L.aline_tc_weight_bytes, c_uint64, P)
    _sig(L.aline_tc_max_keys, c_int32, P)
    _sig(L.aline_tc_fast_max_keys, c_int32, P)
    _sig(L.aline_tc_kv_bytes, c_uint64, P, c_int32, c_int32)
    _sig(L.aline_query_stream_tc, c_int32, P, P, P, P, c_int32, c_int32, P, c_int32, c_int32, c_float, P, P, P, P)
    _sig(L.aline_query_stream_tc_ex, c_int32, P, P, P, P, P, c_int32, c_int32, P, c_int32, c_int32, c_float, P, P, P, P)
    _sig(L.aline_rollout_ex, c_int32, P, P, P, P, P, P, P, P, c_int32, c_int32, c_int32, c_int32, P, c_int32, P, c_int32,
         P, c_int32, P, c_int32, P, P, P, P, P, P)
    _sig(L.aline_gp_scratch_bytes, c_size_t, c_int32, c_int32)
    _sig(L.aline_gp_sample, c_int32, P, c_int32, c_int32, c_int32, P, P, P, P, P, c_float, c_float, P, P, P, P, P,
         c_size_t, P)
    _sig(L.aline_gp_kernel_matrix, c_int32, P, P, c_int32, c_int32, c_int32, P, P, c_int32, P, P)
    _sig(L.aline_tc_selftest, c_int32, P, P, c_int32, c_int32, P, P, P)
    _sig(L.aline_tc_selftest_tmem_a, c_int32, P, P, c_int32, c_int32, P, P)
    _sig(L.aline_censored_sigmoid_normal_log_prob, c_int32, P, P, P, c_float, c_float, c_int64, P, P, P)
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise AlineError(lib().aline_last_error().decode())


_graph_launches = 0


def graph_captured(n_kernels):
    """The C-side counter ticked for `n_kernels` launches that were only CAPTURED into a CUDA graph, not executed."""
    global _graph_launches
    _graph_launches -= int(n_kernels)


def graph_replayed(n_kernels):
    """A replayed CUDA graph executes kernels the C-side counter does not see."""
    global _graph_launches
    _graph_launches += int(n_kernels)


def set_option(name: str, value: int):
    """Process-wide kernel-selection switch (include/aline_b200.h, aline_set_option)."""
    check(lib().aline_set_option(name.encode(), int(value)))


def kernel_launches() -> int:
    """Kernels of this library launched so far (eager launches counted in C + kernels inside replayed graphs)."""
    return int(lib().aline_kernel_launches()) + _graph_launches


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dptr(t, dtype=torch.float32, name="tensor"):
    """Device pointer of a contiguous CUDA tensor of the expected dtype (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise AlineError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback), got "
                         f"{type(t).__name__}{'' if not isinstance(t, torch.Tensor) else ' on ' + str(t.device)}")
    if t.dtype != dtype:
        raise AlineError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise AlineError(f"{name}: tensor must be contiguous")
    return c_void_p(t.data_ptr())


def f32c(t, device=None):
    """fp32 contiguous copy-if-needed on the CUDA device (host tensors are refused, not silently moved)."""
    if not t.is_cuda:
        raise AlineError("expected a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_scratch = {}


def scratch(nbytes, device):
    """Grow-only scratch buffer (uint8) per (device, CURRENT STREAM): the kernels of one stream use it in stream order;
    calls on different streams of a device (the pipelined ``eval_boed``) get different buffers, and a buffer is
    allocated -- hence later recycled by torch's caching allocator -- on the stream that uses it."""
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(index).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        if len(_scratch) > 64:                 # streams come and go (side streams of finished evaluations)
            _scratch.clear()
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=torch.device("cuda", index))
        _scratch[key] = buf
    return buf
