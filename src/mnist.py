"""Compat module for the reference's src/mnist.py: model constants and the
``inference / loss / evaluation / predictions`` functions, expressed on a dict of
parameter tensors instead of TF graph variables."""
import _bootstrap  # noqa: F401

import torch

from distributedmnist_b200.models.lenet import (IMAGE_SIZE, NUM_CHANNELS, NUM_LABELS, SEED,  # noqa: F401
                                                lenet_forward, lenet_spec, loss_and_accuracy)

IMAGE_PIXELS = IMAGE_SIZE * IMAGE_SIZE
PIXEL_DEPTH = 255


def inference(params, images, train=True, keep_mask=None):
    return lenet_forward(params, images, train=train, keep_mask=keep_mask)


def loss(logits, labels):
    return loss_and_accuracy(logits, labels)[0]


def evaluation(logits, labels):
    return loss_and_accuracy(logits, labels)[1]


def predictions(logits):
    return torch.softmax(logits, dim=1)
