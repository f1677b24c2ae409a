"""The trainer on top of the hot path (sngnn_b200/train.py) against R: train.py: flag names / defaults,
factory argument order, the stopping rule (CPU, stubbed steps), and -- on the GPU -- a training trajectory against the CPU
oracle trained with the same Adam settings."""
import pytest
import torch

from sngnn_b200 import train as T


def test_flags_and_defaults_follow_the_reference():
    a = T.parse_args([])
    # R: train.py:23-65 (defaults of the flags that reach the SN models)
    ref = dict(seed=1234, epochs=1000, patience=100, lr=0.01, weight_decay=5e-4, dropout=0.5, hidden_channels=16, num_layers=2,
               part_id=0, top_k=1, thr=0.5, init_beta=0.5, is_remove_self_loops=1)
    for k, v in ref.items():
        assert getattr(a, k) == v, k
    b = T.parse_args("--model SNGNN_Plus --dataset small --top_k 10 --thr 0.9 --init_beta 0.0 --is_remove_self_loops 0".split())
    assert (b.model, b.dataset, b.top_k, b.thr, b.init_beta, b.is_remove_self_loops) == ("SNGNN_Plus", "small", 10, 0.9, 0.0, 0)


def test_factory_argument_order():
    cfg = dict(model="SNGNN_Plus_Plus", hidden_channels=8, num_layers=2, top_k=3, thr=0.25, init_beta=0.125, is_remove_self_loops=0,
               dropout=0.3)
    m = T.build_model(cfg, 12, 4, 50)
    conv = m.lins[0]
    assert (conv.top_k, conv.thr, float(conv.beta), bool(conv.is_remove_self_loops)) == (3, 0.25, 0.125, False)
    assert conv.w.weight.shape == (8, 50) and m.lins[1].lin.out_features == 4 and m.dropout.p == 0.3
    assert type(T.build_model(dict(cfg, model="SNGNN"), 12, 4, 50)).__name__ == "SNGNN"
    assert type(T.build_model(dict(cfg, model="SNGNN_Plus"), 12, 4, 50)).__name__ == "SNGNN_Plus"
    with pytest.raises(ValueError):
        T.build_model(dict(cfg, model="GCN"), 12, 4, 50)


def test_splits_partition_the_nodes():
    tr, va, te = T.make_splits(1000, 3, "cpu")
    assert int(tr.sum()) == 480 and int(va.sum()) == 320 and int(te.sum()) == 200
    assert not (tr & va).any() and not (tr & te).any() and not (va & te).any()
    assert torch.equal(tr, T.make_splits(1000, 3, "cpu")[0]) and not torch.equal(tr, T.make_splits(1000, 4, "cpu")[0])


def test_stopping_rule(monkeypatch):
    """R: train.py:150-158: patience counts epochs without a new smallest validation loss; the reported test accuracy is the one
    at the smallest validation loss."""
    val = iter([1.0, 0.8, 0.9, 0.7, 0.75, 0.74, 0.73, 0.1])
    acc = iter([0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8])
    monkeypatch.setattr(T, "train_step", lambda m, d, o: (torch.tensor(0.5), 0.5))
    monkeypatch.setattr(T, "validate_step", lambda m, d: (torch.tensor(next(val)), 0.0))
    monkeypatch.setattr(T, "test_step", lambda m, d: (torch.tensor(0.0), next(acc)))

    class D:
        x = torch.zeros(1)
    final, hist = T.train(None, D(), None, dict(epochs=100, patience=3, log_every=0))
    assert len(hist) == 7 and final == pytest.approx(0.4)


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["SNGNN", "SNGNN_Plus", "SNGNN_Plus_Plus"])
def test_training_trajectory_matches_the_oracle(model):
    """Five epochs (dropout 0) of the CUDA trainer vs the CPU oracle under the same Adam: losses and accuracies agree."""
    from oracle import sn_ref
    cfg = dict(model=model, hidden_channels=16, num_layers=2, top_k=4, thr=0.0, init_beta=0.3, is_remove_self_loops=1, dropout=0.0,
               epochs=5, patience=100, log_every=0, lr=0.01, weight_decay=5e-4)
    torch.manual_seed(7)
    data, C = T.load_data("small", 0, "cuda")
    n, fd = data.x.shape
    m = T.build_model(cfg, fd, C, n)
    m.dropout.p = 0.0                                     # the SNGNN factory call passes no dropout (R: train.py:304): default 0.5
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to("cuda")
    opt = torch.optim.Adam(m.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
    _, hist = T.train(m, data, opt, cfg)

    # the oracle, trained on the CPU with the same initial parameters
    params = {k: v.clone().contiguous().requires_grad_(True) for k, v in sd0.items()}
    opt_ref = torch.optim.Adam(list(params.values()), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
    x, ei, y = data.x.cpu(), data.edge_index.cpu(), data.y.cpu()
    tr, va = data.train_mask.cpu(), data.val_mask.cpu()
    import torch.nn.functional as F
    for ep in range(cfg["epochs"]):
        opt_ref.zero_grad()
        out = sn_ref.stack_forward(model, sn_ref.params_from_state_dict(params, 2), x, ei, top_k=cfg["top_k"], thr=cfg["thr"],
                                   remove_self_loops=True)
        loss = F.nll_loss(out[tr], y[tr])
        loss.backward()
        opt_ref.step()
        with torch.no_grad():
            out = sn_ref.stack_forward(model, sn_ref.params_from_state_dict(params, 2), x, ei, top_k=cfg["top_k"], thr=cfg["thr"],
                                       remove_self_loops=True)
            vloss = F.nll_loss(out[va], y[va])
        assert hist[ep][0] == pytest.approx(float(loss), rel=2e-4, abs=2e-5), (ep, hist[ep][0], float(loss))
        assert hist[ep][2] == pytest.approx(float(vloss), rel=2e-4, abs=2e-5), (ep, hist[ep][2], float(vloss))


@pytest.mark.gpu
def test_main_runs_the_readme_configuration():
    """R: README.md:63 on the Chameleon-shaped synthetic graph, a few epochs through the command-line entry."""
    lines = []
    final, hist = T.main("--model SNGNN_Plus_Plus --dataset chameleon --num_layers 1 --hidden_channels 32 --top_k 10 --thr 0.9 "
                         "--init_beta 0.0 --is_remove_self_loops 1 --epochs 4 --log-every 1".split(), log=lines.append)
    assert len(hist) == 4 and all(torch.isfinite(torch.tensor(h[0])) for h in hist)
    assert hist[-1][0] < hist[0][0]                       # the training loss goes down
    assert any(l.startswith("Epoch: 3 | Train_loss") for l in lines) and lines[-1].startswith("Part 0 final test acc")


def test_load_data_from_a_graph_file(tmp_path):
    """`--dataset path.pt`: tensors, ten split columns selected by part_id (both layouts), class count."""
    n = 50
    g = torch.Generator().manual_seed(0)
    masks = torch.rand(n, 10, generator=g) < 0.5
    blob = {"x": torch.randn(n, 6, generator=g).double(), "edge_index": torch.randint(0, n, (2, 120), generator=g).int(),
            "y": torch.randint(0, 3, (n,), generator=g), "train_mask": masks, "val_mask": masks.t().contiguous(), "test_mask": masks[:, 0]}
    path = str(tmp_path / "graph.pt")
    torch.save(blob, path)
    data, c = T.load_data(path, part_id=4, device="cpu")
    assert c == 3 and data.x.dtype == torch.float32 and data.edge_index.dtype == torch.int64
    assert torch.equal(data.train_mask, masks[:, 4]) and torch.equal(data.val_mask, masks[:, 4]) and torch.equal(data.test_mask, masks[:, 0])
    del blob["train_mask"], blob["val_mask"], blob["test_mask"]
    torch.save(blob, path)
    data, _ = T.load_data(path, part_id=1, device="cpu")             # no masks in the file: the seeded 48/32/20 split
    assert int(data.train_mask.sum()) == 24 and int(data.val_mask.sum()) == 16 and int(data.test_mask.sum()) == 10
    with pytest.raises(ValueError):
        T.load_data("no-such-shape-or-file", 0, "cpu")
