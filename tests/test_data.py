"""Data decode / normalisation / DataSet semantics (reference mnist_data.py)."""
import numpy as np
import pytest

from distributedmnist_b200 import data as D


def _write_tiny_mnist(tmp_path, n_train=50, n_test=20):
    rng = np.random.RandomState(0)
    tri = rng.randint(0, 256, size=(n_train, 28, 28)).astype(np.uint8)
    trl = rng.randint(0, 10, size=n_train).astype(np.uint8)
    tei = rng.randint(0, 256, size=(n_test, 28, 28)).astype(np.uint8)
    tel = rng.randint(0, 10, size=n_test).astype(np.uint8)
    D.write_idx_images(str(tmp_path / D.TRAIN_IMAGES), tri)
    D.write_idx_labels(str(tmp_path / D.TRAIN_LABELS), trl)
    D.write_idx_images(str(tmp_path / D.TEST_IMAGES), tei)
    D.write_idx_labels(str(tmp_path / D.TEST_LABELS), tel)
    return tri, trl, tei, tel


def test_idx_decode_and_normalisation(tmp_path):
    tri, trl, _, _ = _write_tiny_mnist(tmp_path)
    x = D.extract_data(str(tmp_path / D.TRAIN_IMAGES))
    y = D.extract_labels(str(tmp_path / D.TRAIN_LABELS))
    assert x.shape == (50, 28, 28, 1) and x.dtype == np.float32
    assert y.dtype == np.int64 and (y == trl).all()
    np.testing.assert_allclose(x[..., 0], (tri.astype(np.float32) - 127.5) / 255.0)
    assert x.min() >= -0.5 and x.max() <= 0.5


def test_read_data_sets_validation_is_test_set_and_no_sharding(tmp_path):
    _, _, tei, tel = _write_tiny_mnist(tmp_path)
    ds0 = D.load_mnist(str(tmp_path), worker_id=0, n_workers=4, synthetic=False)
    ds1 = D.load_mnist(str(tmp_path), worker_id=1, n_workers=4, synthetic=False)
    assert ds0.train.num_examples == 50 and ds1.train.num_examples == 50   # every worker gets everything
    assert ds0.validation.num_examples == 20 and ds0.test is None
    assert sorted(ds0.validation.labels.tolist()) == sorted(tel.tolist())
    # independent per-replica shuffles, reproducible
    assert not np.array_equal(ds0.train.labels, ds1.train.labels)
    again = D.load_mnist(str(tmp_path), worker_id=1, n_workers=4, synthetic=False)
    assert np.array_equal(again.train.labels, ds1.train.labels)


def test_missing_files_without_synthetic_raises(tmp_path):
    with pytest.raises(FileNotFoundError):
        D.load_mnist(str(tmp_path), synthetic=False)


def test_next_batch_epoch_wrap():
    x = np.arange(10, dtype=np.float32).reshape(10, 1, 1, 1)
    y = np.arange(10, dtype=np.int64)
    ds = D.DataSet(x, y, seed=3)
    seen = []
    for _ in range(3):
        bx, by = ds.next_batch(3)
        assert (bx[:, 0, 0, 0].astype(np.int64) == by).all()   # images and labels stay paired
        seen += by.tolist()
    assert len(set(seen)) == 9 and ds.epochs_completed == 0
    bx, by = ds.next_batch(3)   # 9+3 > 10 -> reshuffle, restart at 0
    assert ds.epochs_completed == 1 and len(by) == 3


def test_fake_data():
    ds = D.read_data_sets("unused", fake_data=True)
    bx, by = ds.train.next_batch(7)
    assert ds.train.num_examples == 10000
    assert bx.shape == (7, 28, 28, 1) and (bx == 1).all() and (by == 0).all()


def test_synthetic_is_mnist_shaped_and_learnable():
    trx, try_, tex, tey = D.make_synthetic_mnist(2000, 500, seed=1)
    assert trx.shape == (2000, 28, 28, 1) and trx.dtype == np.float32 and try_.dtype == np.int64
    assert trx.min() >= -0.5 and trx.max() <= 0.5 and set(np.unique(try_)) <= set(range(10))
    # nearest-class-mean on raw pixels must beat chance by a wide margin
    means = np.stack([trx[try_ == c].mean(0).ravel() for c in range(10)])
    pred = ((tex.reshape(500, -1)[:, None, :] - means[None]) ** 2).sum(-1).argmin(1)
    assert (pred == tey).mean() > 0.6
