"""Mirror of the reference's ``ResNetRNN`` model class, inference only.

Reference: /root/reference/catfish/models/resnet_class.py (``ResNetRNN`` :7-41,
``residual_block`` :44-82).  ``ResNet`` is the variant the reference obtains by
commenting out resnet_class.py:23 (conv stack straight into the dense layer).
"""

from .rnn_class import RNN


class ResNetRNN(RNN):
    network_type_name = "ResNetRNN"

    def __init__(self, **kwargs):
        self.n_layers_res = kwargs["n_layers_res"]
        self.layer_size_res = kwargs["layer_size_res"]
        self.network_type = "ResNet-RNN"
        RNN.__init__(self, **kwargs)
        self.model_type = self.network_type


class ResNet(ResNetRNN):
    network_type_name = "ResNet"

    def __init__(self, **kwargs):
        ResNetRNN.__init__(self, **kwargs)
        self.network_type = "ResNet"
        self.model_type = self.network_type
