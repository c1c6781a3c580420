"""Compat module for the reference's src/mnist_data.py (``DataSet``, ``extract_data``,
``extract_labels``, ``read_data_sets``, ``load_mnist``)."""
import _bootstrap  # noqa: F401

from distributedmnist_b200.data import (DataSet, Datasets, extract_data, extract_labels,  # noqa: F401
                                        load_mnist, make_synthetic_mnist, read_data_sets)
