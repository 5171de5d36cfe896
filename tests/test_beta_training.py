"""CPU: the training-side mirror: beta symmetry / unit diagonal / gradients
(reference tests/test_beta_symmetry.py) and a tiny end-to-end training run on a synthetic shard."""
import numpy as np
import pytest
import torch


def test_beta_symmetric_unit_diagonal_and_grads():
    from dl_scl_polar.dlscl.beta import SymmetricBeta
    torch.manual_seed(0)
    m = SymmetricBeta(8)
    b = m.beta_matrix()
    assert torch.allclose(b, b.T) and torch.allclose(torch.diagonal(b), torch.ones(8))
    assert float(b.abs().max()) <= 1.0 and float((b - torch.eye(8)).abs().max()) <= 0.2 + 1e-6
    x1, x2 = torch.rand(8), torch.rand(5, 8)
    assert m(x1).shape == (8,) and m(x2).shape == (5, 8)
    m(x2).sum().backward()
    assert m.off_diag.grad is not None and float(m.off_diag.grad.abs().sum()) > 0
    with pytest.raises(ValueError):
        m(torch.rand(2, 3, 8))
    with pytest.raises(ValueError):
        SymmetricBeta(0)


def test_train_beta_learns_a_planted_metric(tmp_path):
    from dl_scl_polar.train import train_beta as T
    rng = np.random.default_rng(0)
    K, n = 16, 4000
    x = rng.random((n, K)).astype(np.float32) + 0.1
    planted = np.eye(K, dtype=np.float32)
    planted[0, 1] = planted[1, 0] = 0.9                    # position 0 looks worse whenever position 1 is large
    y = np.argmin(x @ planted, axis=1).astype(np.int32)
    np.savez_compressed(tmp_path / "toy_part0.npz", abs_l0=x, flip_idx=y, meta="{}")
    T.main(["--M", "4", "--data", str(tmp_path / "toy_part*.npz"), "--epochs", "12", "--lr", "0.01", "--batch", "256",
            "--lambda_l2", "0.0", "--cpu", "--checkpoint_dir", str(tmp_path / "ck"), "--log_dir", str(tmp_path / "lg")])
    beta = np.load(tmp_path / "ck" / "beta_M4.npy")
    assert beta.shape == (K, K) and beta.dtype == np.float32
    assert np.allclose(beta, beta.T) and np.allclose(np.diag(beta), 1.0)
    rows = (tmp_path / "lg" / "train_M4.csv").read_text().splitlines()
    assert rows[0] == "epoch,train_loss,train_acc,val_loss,val_acc" and len(rows) == 13
    first, last = [float(v) for v in rows[1].split(",")], [float(v) for v in rows[-1].split(",")]
    assert last[3] < first[3] and last[1] < first[1] and last[2] > first[2]   # losses fall, training accuracy rises
    assert beta[0, 1] > 0.3                                  # the planted coupling is recovered in sign and size
    with pytest.raises(FileNotFoundError):
        T._load_dataset([str(tmp_path / "nothing*.npz")])
