"""codae.model -- the two denoising autoencoders (same import surface as the reference package)."""
from .embedding_denoising_autoencoder import EmbeddingDenoisingAutoencoder
from .mixed_variable_denoising_autoencoder import MixedVariableDenoisingAutoencoder
from .cnn_autoencoder import cnnAutoencoder

__all__ = ["EmbeddingDenoisingAutoencoder", "MixedVariableDenoisingAutoencoder", "cnnAutoencoder"]
