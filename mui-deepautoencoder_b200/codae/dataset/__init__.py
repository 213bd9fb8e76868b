"""codae.dataset -- boundary objects that hold the resident [N, io] tensor and its `arch` descriptor.
(The reference also exports an image-dataset converter here; it is outside the hot path and not provided.)"""
from codae.dataset.mixed_variable_dataset import MixedVariableDataset
from codae.dataset.concatenated_embedding_dataset import ConcatenatedEmbeddingDataset

__all__ = ["ConcatenatedEmbeddingDataset", "MixedVariableDataset"]
