import torch


class cnnAutoencoder(torch.nn.Module):
    """Placeholder kept for import parity: the reference class is an empty stub too
    (codae/model/cnn_autoencoder.py:12-14) and is not on the hot path."""
    pass
