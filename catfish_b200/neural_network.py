"""Mirror of /root/reference/catfish/neural_network.py: build / load a network.

``build_model`` :8-23, ``load_network`` :26-34, ``retrieve_hyperparams`` :37-67.
"""

import os

from .resnet_class import ResNet, ResNetRNN
from .rnn_class import RNN

SHIPPED_MODEL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ResNetRNN")


def build_model(network_type, saving=False, **kwargs):
    """neural_network.py:8-23 ("ResNet" is the commented-out variant of resnet_class.py:23)."""
    network = None
    if network_type == "RNN":
        network = RNN(save=saving, **kwargs)
    elif network_type == "ResNetRNN":
        network = ResNetRNN(save=saving, **kwargs)
    elif network_type == "ResNet":
        network = ResNet(save=saving, **kwargs)
    return network


def load_network(network_type, path_to_network=None, checkpoint=30000, **model_kwargs):
    """neural_network.py:26-34.  ``path_to_network`` defaults to the shipped ResNetRNN."""
    if path_to_network is None:
        path_to_network = SHIPPED_MODEL_DIR
    hpm_dict = retrieve_hyperparams(path_to_network + "/ResNetRNN.txt")
    hpm_dict.update(model_kwargs)
    model = build_model(network_type, **hpm_dict)
    model.restore_network("{}/checkpoints".format(path_to_network), ckpnt="ckpnt-{}".format(checkpoint))
    return model


def retrieve_hyperparams(model_file, split_on=": "):
    """neural_network.py:37-67: same line prefixes, same types."""
    hpm_dict = {}
    with open(model_file, "r") as source:
        for line in source:
            if line.startswith("batch_size"):
                hpm_dict["batch_size"] = int(line.strip().split(split_on)[1])
            elif line.startswith("optimizer_choice"):
                hpm_dict["optimizer_choice"] = line.strip().split(split_on)[1]
            elif line.startswith("learning_rate"):
                hpm_dict["learning_rate"] = float(line.strip().split(split_on)[1])
            elif line.startswith("layer_size:"):
                hpm_dict["layer_size"] = int(line.strip().split(split_on)[1])
            elif line.startswith("n_layers:"):
                hpm_dict["n_layers"] = int(line.strip().split(split_on)[1])
            elif line.startswith("keep_prob"):
                hpm_dict["keep_prob"] = float(line.strip().split(split_on)[1])
            elif line.startswith("layer_size_res"):
                hpm_dict["layer_size_res"] = int(line.strip().split(split_on)[1])
            elif line.startswith("n_layers_res"):
                hpm_dict["n_layers_res"] = int(line.strip().split(split_on)[1])
    return hpm_dict
