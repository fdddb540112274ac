"""Mirror of models/res_gru_net.py: deep residual encoder/decoder around the 3-D conv GRU, and
the 3D-conv-LSTM variant of BASELINE.json config 4 (no reference code exists for it,
lib/layers.py:508-575 is a stub)."""
from .net import Net
from . import layerwise


class ResidualGRUNet(Net):
    NET_NAME = 'ResidualGRUNet'

    def forward_layerwise(self, x, y=None):
        """models/res_gru_net.py:16-204 through the drop-in layer classes (parity aid)."""
        return layerwise.run(self, x, y, residual=True)


class ResidualLSTMNet(Net):
    NET_NAME = 'ResidualLSTMNet'
