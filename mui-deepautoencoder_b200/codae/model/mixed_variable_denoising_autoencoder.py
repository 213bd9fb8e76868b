"""MixedVariableDenoisingAutoencoder -- same constructor, attributes and exceptions as
codae/model/mixed_variable_denoising_autoencoder.py of the reference; compute on libcodae_b200."""
import math

import torch

from ._flat_mlp import FlatMLP


class MixedVariableDenoisingAutoencoder(FlatMLP):

    def __init__(self, arch, io_size, z_size, device, nb_input_layer=2, nb_output_layer=2, steep_layer_size=True,
                 activation=torch.nn.ReLU):
        """Layer-size rule of the reference (mixed_variable_denoising_autoencoder.py:45-125): ceil increments
        and a final Linear(io, io)."""
        super(MixedVariableDenoisingAutoencoder, self).__init__()
        if activation is not torch.nn.ReLU:
            raise Exception("Error: only torch.nn.ReLU is implemented by the B200 kernels")
        self.arch = arch
        self.io_size = io_size
        self.z_size = z_size
        self.device = device
        self.nb_input_layer = nb_input_layer
        self.nb_output_layer = nb_output_layer
        self.steep_layer_size = steep_layer_size
        self.activation = activation
        self.mode = 0

        inc_in = inc_out = 0
        if not steep_layer_size:
            delta = io_size - z_size
            inc_in = math.ceil(delta / nb_input_layer)
            inc_out = math.ceil(delta / nb_output_layer)

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

        dec = []
        for i in range(nb_output_layer):
            if steep_layer_size:
                dec += [torch.nn.Linear(z_size if i == 0 else io_size, io_size), activation(True)]
            else:
                a = min(z_size + i * inc_out, io_size)
                width = min(z_size + (i + 1) * inc_out, io_size)
                dec += [torch.nn.Linear(a, width), activation(True)]
        dec.append(torch.nn.Linear(io_size, io_size))
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

    def corrupt(self, input_data, mask, corruption_type="zero_continuous"):
        """(mixed_variable_denoising_autoencoder.py:222-262)"""
        if corruption_type == "zero_continuous":
            return self._corrupt_zero_continuous(input_data=input_data, mask=mask)
        else:
            raise Exception("Error: invalid corruption type requested (zero_continuous).")

    def _corrupt_zero_continuous(self, input_data, mask):
        return self.corrupt_dense(input_data, mask)
