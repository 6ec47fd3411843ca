"""Pack an ALINE state dict into the fp32 parameter blob the kernels read (layout: include/aline_b200.h,
``aline_model``; offsets: csrc/model.cuh ``make_layout``).  Every ``nn.Linear`` weight is transposed to [in][out];
segments are zero-padded to a multiple of 4 floats where the layout says so."""
from __future__ import annotations

import ctypes

import torch

from .. import _lib


class AlineModelStruct(ctypes.Structure):
    """``struct aline_model`` (include/aline_b200.h)."""
    _fields_ = [("d", ctypes.c_int32), ("ff", ctypes.c_int32), ("n_head", ctypes.c_int32), ("n_layer", ctypes.c_int32),
                ("dim_x", ctypes.c_int32), ("dim_y", ctypes.c_int32), ("n_theta_tok", ctypes.c_int32),
                ("n_comp", ctypes.c_int32), ("emb_hidden", ctypes.c_int32), ("head_hidden", ctypes.c_int32),
                ("time_token", ctypes.c_int32), ("std_min", ctypes.c_float), ("params", ctypes.c_void_p),
                ("n_params", ctypes.c_uint64)]


def _pad4(t):
    n = t.numel()
    r = (-n) % 4
    t = t.reshape(-1)
    return torch.cat([t, t.new_zeros(r)]) if r else t


def pack_state_dict(sd, n_head, std_min=1e-4, device=None):
    """-> (blob fp32 [n_params] on `device`, dims dict).  `sd` uses the reference key names
    (embedder.* / encoder.encoder.layers.* / head.*)."""
    f = lambda k: sd[k].detach().to(torch.float32)   # noqa: E731
    T = lambda k: f(k).t().contiguous().reshape(-1)  # noqa: E731
    d = f("embedder.x_embedder.2.weight").shape[0]
    eh, dx = f("embedder.x_embedder.0.weight").shape
    dy = f("embedder.y_embedder.0.weight").shape[1]
    ntok = f("embedder.theta_tokens").shape[0] if "embedder.theta_tokens" in sd else 0
    n_layer = 0
    while f"encoder.encoder.layers.{n_layer}.linear1.weight" in sd:
        n_layer += 1
    ff = f("encoder.encoder.layers.0.linear1.weight").shape[0]
    hh, d_acq = f("head.acquisition_head.predictor.0.weight").shape
    time_token = int(d_acq == d + 1)
    n_comp = 0
    while f"head.target_head.heads.{n_comp}.0.weight" in sd:
        n_comp += 1
    parts = []
    for e in ("x_embedder", "y_embedder"):
        parts += [_pad4(T(f"embedder.{e}.0.weight")), f(f"embedder.{e}.0.bias"), T(f"embedder.{e}.2.weight"),
                  f(f"embedder.{e}.2.bias")]
    parts.append(_pad4(f("embedder.theta_tokens")) if ntok else torch.zeros(0))
    for l in range(n_layer):
        p = f"encoder.encoder.layers.{l}."
        W, b = f(p + "self_attn.in_proj_weight"), f(p + "self_attn.in_proj_bias")
        parts += [W[:d].t().contiguous().reshape(-1), W[d:2 * d].t().contiguous().reshape(-1),
                  W[2 * d:].t().contiguous().reshape(-1), b[:d], b[d:2 * d], b[2 * d:],
                  T(p + "self_attn.out_proj.weight"), f(p + "self_attn.out_proj.bias"),
                  f(p + "norm1.weight"), f(p + "norm1.bias"),
                  T(p + "linear1.weight"), f(p + "linear1.bias"), T(p + "linear2.weight"), f(p + "linear2.bias"),
                  f(p + "norm2.weight"), f(p + "norm2.bias")]
    parts += [T("head.acquisition_head.predictor.0.weight"), f("head.acquisition_head.predictor.0.bias"),
              f("head.acquisition_head.predictor.2.weight").reshape(-1),
              _pad4(f("head.acquisition_head.predictor.2.bias"))]
    for c in range(n_comp):
        p = f"head.target_head.heads.{c}."
        parts += [T(p + "0.weight"), f(p + "0.bias"), f(p + "2.weight").reshape(-1), _pad4(f(p + "2.bias"))]
    dev = device if device is not None else parts[0].device
    blob = torch.cat([p.reshape(-1).to(dev) for p in parts]).contiguous()
    dims = dict(d=d, ff=ff, n_head=int(n_head), n_layer=n_layer, dim_x=dx, dim_y=dy, n_theta_tok=ntok, n_comp=n_comp,
                emb_hidden=eh, head_hidden=hh, time_token=time_token, std_min=float(std_min))
    return blob, dims


def tile_bf16(W):
    """[R, K] -> bf16 in the core-matrix tiled operand layout of csrc/tc.cuh: K/8 chunks, chunk c = rows x 8 columns."""
    R, K = W.shape
    return W.to(torch.bfloat16).reshape(R, K // 8, 8).permute(1, 0, 2).contiguous().reshape(-1)


def _split_bf16(v):
    """fp32 -> (hi, lo) with hi = bf16(v), lo = bf16(v - hi): hi + lo carries ~16 mantissa bits."""
    hi = v.to(torch.bfloat16).float()
    return hi, (v - hi).to(torch.bfloat16).float()


def _augment(W, b, wt=None, bias_first=False):
    """[N, K] weight + bias -> [N, K + 16]: the 16 extra input columns [b_hi, b_lo, wt, wt, 0 x 12] multiply the
    operand columns [1, 1, t_hi, t_lo, 0 ...] of the fast tensor-core kernel (csrc/query_tc3.cu)."""
    N = W.shape[0]
    extra = W.new_zeros((N, 16))
    extra[:, 0], extra[:, 1] = _split_bf16(b)
    if wt is not None:
        extra[:, 2] = wt
        extra[:, 3] = wt
    return torch.cat([extra, W], 1) if bias_first else torch.cat([W, extra], 1)


def pack_tc_weights(sd, device):
    """bf16 weight blob of the tensor-core query streams (include/aline_b200.h, aline_query_stream_tc): section 1
    for the general kernel (d = 32 only), section 2 (biases folded in) for the fast kernels."""
    f = lambda k: sd[k].detach().to(torch.float32).cpu()   # noqa: E731
    d = f("embedder.x_embedder.2.weight").shape[0]
    c = 1.4426950408889634 / (8.0 ** 0.5)                  # log2(e) / sqrt(head_dim)
    parts, fast, l = [], [], 0
    while f"encoder.encoder.layers.{l}.linear1.weight" in sd:
        p = f"encoder.encoder.layers.{l}."
        Win, b_in = f(p + "self_attn.in_proj_weight"), f(p + "self_attn.in_proj_bias")
        parts += [tile_bf16(Win[:d]), tile_bf16(f(p + "self_attn.out_proj.weight")),
                  tile_bf16(f(p + "linear1.weight")), tile_bf16(f(p + "linear2.weight"))]
        fast += [tile_bf16(_augment(Win[:d] * c, b_in[:d] * c)),
                 tile_bf16(_augment(f(p + "self_attn.out_proj.weight"), f(p + "self_attn.out_proj.bias"))),
                 tile_bf16(_augment(f(p + "linear1.weight"), f(p + "linear1.bias"))),
                 tile_bf16(_augment(f(p + "linear2.weight"), f(p + "linear2.bias"), bias_first=True))]
        l += 1
    Wa, ba = f("head.acquisition_head.predictor.0.weight"), f("head.acquisition_head.predictor.0.bias")
    parts.append(tile_bf16(Wa[:, :d].contiguous()))
    fast.append(tile_bf16(_augment(Wa[:, :d], ba, wt=Wa[:, d] if Wa.shape[1] == d + 1 else None)))
    if d != 32:                                            # no general kernel: the blob is the fast section alone
        parts = []
    return torch.cat(parts + fast).contiguous().to(device)


class PackedModel:
    """Device blob + the C descriptor; validates the blob size against the C layout."""

    def __init__(self, sd, n_head, std_min=1e-4, device=None):
        self.blob, self.dims = pack_state_dict(sd, n_head, std_min, device)
        if not self.blob.is_cuda:
            raise _lib.AlineError("model parameters must live on a CUDA device (the B200 path has no CPU fallback)")
        self.desc = AlineModelStruct(params=self.blob.data_ptr(), n_params=self.blob.numel(), **self.dims)
        need = _lib.lib().aline_model_param_count(ctypes.byref(self.desc))
        if need != self.blob.numel():
            raise _lib.AlineError(f"packed parameter blob has {self.blob.numel()} floats, the kernels expect {need}")
        # tensor-core (bf16) operands, when the model shape has a tcgen05 kernel
        self.tc_max_keys = int(_lib.lib().aline_tc_max_keys(ctypes.byref(self.desc)))
        self.tc_fast_max_keys = 0
        self.tc_blob = None
        self.tc_fast_max_keys = int(_lib.lib().aline_tc_fast_max_keys(ctypes.byref(self.desc)))
        if self.tc_max_keys > 0 or self.tc_fast_max_keys > 0:
            self.tc_blob = pack_tc_weights(sd, self.blob.device)
            want = int(_lib.lib().aline_tc_weight_bytes(ctypes.byref(self.desc)))
            if self.tc_blob.numel() * 2 != want:
                raise _lib.AlineError(f"bf16 weight blob has {self.tc_blob.numel() * 2} bytes, the kernels expect {want}")

    @property
    def ref(self):
        return ctypes.byref(self.desc)

    @property
    def device(self):
        return self.blob.device
