"""world_size-2 data-parallel distillation step on CPU (gloo): utterances sharded across ranks,
frozen teacher and student replicated, per-rank loss on the local shard, ONE all-reduce of the flat
student-gradient bucket, identical Adam on every rank.  The CUDA kernels are replaced by
the numpy model of the C ABI (tests/cabi_emu.py); NCCL is replaced by gloo."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
CFG_T = dict(kernel_num=[4, 8, 8, 16, 16, 16], rnn_units=16)
CFG_S = dict(kernel_num=[2, 4, 4, 8, 8, 8], rnn_units=8)


def _setup_paths():
    root = os.path.dirname(HERE)
    for p in (root, os.path.join(root, "speech-enhancement-clskd_b200"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)


def _models():
    import clskd_b200
    from oracle import dccrn_oracle as D
    out = []
    for cfg, seed in ((CFG_T, 1), (CFG_S, 2)):
        m = clskd_b200.DCCRN(rnn_units=cfg["rnn_units"], masking_mode="E", use_clstm=True, kernel_num=cfg["kernel_num"])
        m.load_state_dict(D.make_state_dict(cfg["kernel_num"], cfg["rnn_units"], seed=seed))
        out.append(m)
    return out


def _batch():
    g = torch.Generator().manual_seed(0)
    return 0.1 * torch.randn(4, 1500, generator=g), 0.1 * torch.randn(4, 1500, generator=g)


def _worker(rank, world, port, mode, out_dir, unsynced=False):
    _setup_paths()
    torch.set_num_threads(1)
    import cabi_emu
    cabi_emu.install()
    from clskd_b200.distill import DistillTrainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    teacher, student = _models()
    X, y = _batch()
    n = X.shape[0] // world
    Xs, ys = X[rank * n:(rank + 1) * n], y[rank * n:(rank + 1) * n]
    torch.manual_seed(3 + (rank if unsynced else 0))      # unsynced: every rank draws its own ABF weights ...
    if unsynced and rank > 0:                             # ... and starts from a different student
        with torch.no_grad():
            for p in student.parameters():
                p.add_(0.01 * torch.randn(p.shape))
        torch.manual_seed(3 + rank)
    tr = DistillTrainer(teacher, student, mode=mode, lr=1e-3, example_input=Xs)
    losses = [float(tr.train_step(Xs, ys)) for _ in range(2)]
    torch.save({"flat_p": tr.opt.flat_p.clone(), "losses": losses}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode,unsynced", [("spkd_all", False), ("clskd", False), ("clskd", True)])
def test_two_rank_step_equals_averaged_gradients(tmp_path, mode, unsynced, monkeypatch):
    """unsynced = True: the ranks construct different students / ABF blocks; the trainer's initial broadcast of
    rank 0's parameter bucket must make the run identical to the synchronised one."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path), unsynced), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["flat_p"], r1["flat_p"]), "ranks diverged"

    # single-process reference: per-shard gradients averaged by hand, same optimizer
    _setup_paths()
    import cabi_emu
    cabi_emu.install(monkeypatch)
    from clskd_b200.distill import DistillStep, FlatAdam
    teacher, student = _models()
    X, y = _batch()
    torch.manual_seed(3)
    step = DistillStep(teacher, student, mode=mode)
    step.materialize(X[:2])
    student.train()
    opt = FlatAdam(step.trainable_parameters(), lr=1e-3)
    for _ in range(2):
        acc = torch.zeros_like(opt.flat_g)
        for r in range(world):
            opt.zero_grad()
            step(X[2 * r:2 * r + 2], y[2 * r:2 * r + 2]).backward()
            acc += opt.pack_grads()
        opt.flat_g.copy_(acc)
        opt.step(world)
    assert torch.allclose(opt.flat_p, r0["flat_p"], rtol=1e-4, atol=1e-6)
