"""The handful of Theano names a reference network file uses (models/gru_net.py:3-7,19,114-119;
models/net.py:5-9,25-29), for running such a file define-by-run on top of lib/layers.py:

    import theano                  ->  from lib import theano_compat as theano
    import theano.tensor as tensor ->  from lib.theano_compat import tensor

`tensor.tensor4()` / `TensorType(...)()` declare an input (models.net.Net feeds the real batch through the
same attribute), `scan` is a host loop over the leading axis whose step outputs stay on the gradient tape,
`zeros_like` makes the initial recurrent state.  Nothing here computes: arithmetic is the layer classes'.
"""
import numpy as np
import torch


class _Config:
    floatX = 'float32'


config = _Config()


class Symbol:
    """An input declaration (`self.x = tensor5()`): only its rank is known until a batch is fed."""

    def __init__(self, ndim):
        self.ndim = ndim


class ScanResult(list):
    """One output of scan over all steps; `r[-1]` is the last step's value (still on the tape)."""

    def stack(self):
        return torch.stack([torch.as_tensor(v) for v in self])


class _Tensor:
    @staticmethod
    def tensor4(name=None):
        return Symbol(4)

    class TensorType:
        def __init__(self, dtype, broadcastable):
            self.ndim = len(broadcastable)

        def __call__(self, name=None):
            return Symbol(self.ndim)

    @staticmethod
    def zeros_like(a, dtype=None):
        from . import layers as ly
        return torch.zeros(tuple(np.shape(a)), dtype=torch.float32, device=ly.DEVICE)


tensor = _Tensor()


def scan(fn, sequences=None, outputs_info=None, non_sequences=None, n_steps=None, **unused):
    """theano.scan(fn, sequences, outputs_info) for the recurrence of the reference nets: fn(*sequence slices,
    *previous outputs) -> new outputs; returns (outputs over time, updates)."""
    seqs = list(sequences or [])
    state = list(outputs_info or [])
    extra = list(non_sequences or [])
    n = int(n_steps) if n_steps is not None else min(len(s) for s in seqs)
    outs = None
    for t in range(n):
        r = fn(*([s[t] for s in seqs] + state + extra))
        r = list(r) if isinstance(r, (tuple, list)) else [r]
        if outs is None:
            outs = [ScanResult() for _ in r]
        for o, v in zip(outs, r):
            o.append(v)
        if state:
            state = r[:len(state)] if len(r) >= len(state) else r
    outs = outs or []
    return (outs if len(outs) != 1 else outs[0]), {}
