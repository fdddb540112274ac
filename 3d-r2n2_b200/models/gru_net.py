"""Mirror of models/gru_net.py: shallow encoder, 3-D conv GRU, 5-conv decoder."""
from .net import Net
from . import layerwise


class GRUNet(Net):
    NET_NAME = 'GRUNet'

    def forward_layerwise(self, x, y=None):
        """The reference's network_definition executed layer by layer through the drop-in layer
        classes (models/gru_net.py:16-144), per-view encoder inside the loop.  Slow; parity aid."""
        return layerwise.run(self, x, y, residual=False)
