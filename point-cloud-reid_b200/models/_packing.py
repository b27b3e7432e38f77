"""Weight packing shared by the model modules: torch parameters stay in the reference's layout (so
reference checkpoints load with strict=True); kernels consume k-major fp32 copies with eval-mode
BatchNorm folded in.  Copies are cached and rebuilt when any parameter/buffer changes (as seen through the tensors'
version counters; weights must not be mutated through `.data` -- see PackedModule.invalidate_packed)."""
import torch
from torch import nn


def kmajor(w):
    """torch weight (CO, K[,1[,1]]) -> contiguous k-major (K, CO)."""
    return w.detach().reshape(w.shape[0], -1).t().contiguous().float()


def bn_scale_shift(bn):
    s = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    t = bn.bias.detach() - s * bn.running_mean.detach()
    return s.float(), t.float()


def fold_bn(w, b, bn):
    """y = BN(W x + b) -> (W', b') with W' = s*W, b' = s*b + t."""
    s, t = bn_scale_shift(bn)
    w2 = w.detach().reshape(w.shape[0], -1).float() * s[:, None]
    b2 = t if b is None else s * b.detach().float() + t
    return w2, b2.contiguous()


def invalidate_packed(root):
    """drops the cached packed copies of every module below `root`.  The cache key is (data_ptr, _version, device) of every
    parameter / buffer: load_state_dict, .to(), optimizer steps and in-place ops bump it, but writes that bypass the version
    counter (`p.data.copy_()`, `p.data.mul_()`, storage-level edits as some EMA / quantisation utilities do) do NOT -- after
    such a write call this (ReIDNet.invalidate_packed()), or the kernels keep using the old folded weights."""
    for m in root.modules():
        for attr in ("_pk_key", "_pcreid_key", "_pcreid_fin_key"):
            if hasattr(m, attr):
                object.__setattr__(m, attr, None)


class PackedModule(nn.Module):
    """nn.Module whose forward runs on packed copies of its (and its plain children's) tensors."""

    def _pack(self):          # -> dict of device tensors
        raise NotImplementedError

    def _pack_key(self):
        return tuple((t.data_ptr(), t._version, str(t.device)) for t in
                     list(self.parameters()) + list(self.buffers()))

    def packed(self):
        key = self._pack_key()
        if getattr(self, "_pk_key", None) != key:
            with torch.no_grad():
                object.__setattr__(self, "_pk", self._pack())
            object.__setattr__(self, "_pk_key", key)
        return self._pk

    def invalidate_packed(self):
        invalidate_packed(self)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        object.__setattr__(self, "_pk_key", None)        # belt and braces: a checkpoint load always repacks

    def _inference_only(self):
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: pcreid_b200 implements the inference path only; call .eval() "
                               "(BatchNorm uses running statistics)")
