"""Mirror of /root/reference/catfish/neural_network.py: build / load a network.

Same entry points and behaviour as the reference (``build_model`` :8-23, ``load_network`` :26-34,
``retrieve_hyperparams`` :37-67); the returned objects are the CUDA-backed model classes of this
package instead of TensorFlow graphs.
"""

import os

from .resnet_class import ResNet, ResNetRNN
from .rnn_class import RNN

SHIPPED_MODEL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ResNetRNN")

# network_type -> class; "ResNet" is the variant the reference gets by commenting out resnet_class.py:23
_NETWORKS = {"RNN": RNN, "ResNetRNN": ResNetRNN, "ResNet": ResNet}

# line prefix in <model>.txt -> (kwarg, converter), tested in this order like the reference's elif chain
# ("layer_size:" / "n_layers:" carry the colon so that they do not swallow the *_res lines)
_HYPERPARAM_LINES = (
    ("batch_size", "batch_size", int),
    ("optimizer_choice", "optimizer_choice", str),
    ("learning_rate", "learning_rate", float),
    ("layer_size:", "layer_size", int),
    ("n_layers:", "n_layers", int),
    ("keep_prob", "keep_prob", float),
    ("layer_size_res", "layer_size_res", int),
    ("n_layers_res", "n_layers_res", int),
)


def build_model(network_type, saving=False, **kwargs):
    """Network object for ``network_type`` ("RNN", "ResNetRNN", "ResNet"); None for an unknown type,
    as the reference's if/elif falls through."""
    cls = _NETWORKS.get(network_type)
    return cls(save=saving, **kwargs) if cls is not None else None


def load_network(network_type, path_to_network=None, checkpoint=30000, **model_kwargs):
    """Hyper-parameters from ``<path>/ResNetRNN.txt``, weights from ``<path>/checkpoints/ckpnt-<N>``.
    ``path_to_network`` defaults to the shipped ResNetRNN; extra kwargs (``device``, ``engine``) go to
    the model constructor."""
    root = SHIPPED_MODEL_DIR if path_to_network is None else path_to_network
    hyperparams = retrieve_hyperparams(root + "/ResNetRNN.txt")
    hyperparams.update(model_kwargs)
    model = build_model(network_type, **hyperparams)
    model.restore_network("{}/checkpoints".format(root), ckpnt="ckpnt-{}".format(checkpoint))
    return model


def retrieve_hyperparams(model_file, split_on=": "):
    """``{kwarg: value}`` parsed from the ``key: value`` lines a training run wrote (later lines win)."""
    found = {}
    with open(model_file, "r") as source:
        for line in source:
            for prefix, key, convert in _HYPERPARAM_LINES:
                if line.startswith(prefix):
                    found[key] = convert(line.strip().split(split_on)[1])
                    break
    return found
