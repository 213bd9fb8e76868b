"""EmbeddingDenoisingAutoencoder -- same constructor, attributes, state_dict keys and exceptions as
codae/model/embedding_denoising_autoencoder.py of the reference; compute on libcodae_b200."""
import math

import torch

from ._flat_mlp import FlatMLP


class EmbeddingDenoisingAutoencoder(FlatMLP):

    def __init__(self, io_size, z_size, embedding_size, nb_input_layer=2, nb_output_layer=2, steep_layer_size=True,
                 activation=torch.nn.ReLU):
        """Layer-size rule of the reference (embedding_denoising_autoencoder.py:49-129): floor increments,
        ReLU(inplace) after every Linear but the z-layer and the last, last Linear fed by the previous
        decoder width.  Configurations the reference cannot build raise at the same point."""
        super(EmbeddingDenoisingAutoencoder, self).__init__()
        if io_size % embedding_size != 0:
            raise Exception("Error: io_size must be a multiple of embedding_size")
        if activation is not torch.nn.ReLU:
            raise Exception("Error: only torch.nn.ReLU is implemented by the B200 kernels")
        self.embedding_size = embedding_size
        self.nb_category = io_size / embedding_size
        self.io_size = io_size
        self.z_size = z_size
        self.nb_input_layer = nb_input_layer
        self.nb_output_layer = nb_output_layer
        self.steep_layer_size = steep_layer_size
        self.activation = activation
        self.mode = 0

        inc_in = inc_out = 0
        if not steep_layer_size:
            delta = io_size - z_size
            inc_in = math.floor(delta / nb_input_layer)
            inc_out = math.floor(delta / nb_output_layer)

        # encoder: module order matters -- nn.Linear's default init consumes the torch RNG exactly like the
        # reference, so a seeded construction yields bit-identical xavier weights.
        enc, width = [torch.nn.Linear(io_size, io_size), activation(True)], None
        for i in range(1, nb_input_layer):
            if steep_layer_size:
                enc += [torch.nn.Linear(io_size, io_size), activation(True)]
            else:
                a = max(io_size - (i - 1) * inc_in, z_size)
                width = max(io_size - i * inc_in, z_size)
                enc += [torch.nn.Linear(a, width), activation(True)]
        if steep_layer_size:
            enc.append(torch.nn.Linear(io_size, z_size))
        else:
            if width is None:
                raise UnboundLocalError("cannot access local variable 'next_layer_output_size' where it is not "
                                        "associated with a value")
            enc.append(torch.nn.Linear(width, z_size))
        self.input_layer = torch.nn.Sequential(*enc)
        self.input_layer.apply(self.init_weight_general_rule)
        self.input_layer.apply(self.init_bias_zero)

        dec, width = [], None
        for i in range(nb_output_layer):
            if steep_layer_size:
                dec += [torch.nn.Linear(z_size if i == 0 else io_size, io_size), activation(True)]
            else:
                a = min(z_size + i * inc_out, io_size)
                width = min(z_size + (i + 1) * inc_out, io_size)
                dec += [torch.nn.Linear(a, width), activation(True)]
        if width is None:  # the reference's steep Embedding model dies here too (…:126)
            raise UnboundLocalError("cannot access local variable 'next_layer_output_size' where it is not "
                                    "associated with a value")
        dec.append(torch.nn.Linear(width, io_size))
        self.output_layer = torch.nn.Sequential(*dec)
        self.output_layer.apply(self.init_weight_general_rule)
        self.output_layer.apply(self.init_bias_zero)

        seq = list(self.input_layer) + list(self.output_layer)
        dims, relu = [], []
        for j, m in enumerate(seq):
            if isinstance(m, torch.nn.Linear):
                dims.append((m.in_features, m.out_features))
                relu.append(j + 1 < len(seq) and isinstance(seq[j + 1], torch.nn.ReLU))
        self._finish_init(dims, relu, sum(isinstance(m, torch.nn.Linear) for m in self.input_layer))

    def corrupt(self, input_data, mask):
        """input_data.clone() * mask (embedding_denoising_autoencoder.py:226-239)."""
        return self.corrupt_dense(input_data, mask)
