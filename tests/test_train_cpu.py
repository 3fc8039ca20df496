"""CPU tests of the training-path neighbours: oracle hand vectors for max-pool / softmax / dropout, the oracle
training step itself, and (gloo, world_size 2) the data-parallel gradient bucket + momentum bookkeeping of
tensorium_b200/dataparallel.py checked against a single-process run on the concatenated batch."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_maxpool_hand_vectors(oracle):
    x = np.arange(16, dtype=np.float32).reshape(1, 1, 4, 4)
    y, idx = oracle.maxpool_forward(x, 2, 2, 2)
    assert y.reshape(-1).tolist() == [5, 7, 13, 15] and idx.reshape(-1).tolist() == [5, 7, 13, 15]
    # ties keep the FIRST maximum of the (row, col) walk: strict '>' in nMaxPoolLayer.pas:433
    x = np.ones((1, 1, 2, 2), np.float32)
    y, idx = oracle.maxpool_forward(x, 2, 2, 2)
    assert idx.item() == 0
    # padding=1 shifts the window by -(1 div 2) = 0 and only grows the output (nMaxPoolLayer.pas:106-107, 409)
    x = np.arange(9, dtype=np.float32).reshape(1, 1, 3, 3)
    y, idx = oracle.maxpool_forward(x, 2, 1, 1, padding=1)
    assert y.shape == (1, 1, 3, 3) and y[0, 0, 2, 2] == 8 and y[0, 0, 0, 0] == 4
    up = np.zeros((1, 1, 3, 3), np.float32)
    oracle.maxpool_backward(idx, np.ones_like(y), up)
    assert up.sum() == 9 and up[0, 0, 2, 2] == 4  # element 8 wins four windows


def test_softmax_cross_entropy_hand_vectors(oracle):
    z = np.array([[0.0, np.log(3.0)], [5.0, 5.0]], np.float32)
    p = oracle.softmax_batch(z)
    assert np.allclose(p, [[0.25, 0.75], [0.5, 0.5]], atol=1e-7)
    t = np.array([[0, 1], [1, 0]], np.float32)
    d, e = oracle.softmax_cross_entropy(p, t)
    assert np.allclose(d, t - p) and np.allclose(e, [[0, -np.log(0.75)], [-np.log(0.5), 0]], atol=1e-7)
    d, e = oracle.softmax_cross_entropy(np.array([[0.0, 1.0]], np.float32), np.array([[1.0, 0.0]], np.float32))
    assert np.isclose(e[0, 0], -np.log(1e-6), rtol=1e-6)  # max(p, sEPSILON) nsoftmaxlayer.pas:132


def test_dropout_hash_known_answers(oracle):
    # xorshift64* of cuda_sgemm.cu:1620-1653, first values for seed 12345 (integers kept to 28 bits)
    def ref(seed, i):
        x = (seed + i) & (2 ** 64 - 1)
        x ^= x >> 12
        x ^= (x << 25) & (2 ** 64 - 1)
        x ^= x >> 27
        x = (x * 0x2545F4914F6CDD1D) & (2 ** 64 - 1)
        return np.float32(np.float32(x % 0x10000000) / np.float32(0x10000000))
    r = oracle.dropout_rand(64, 12345)
    assert all(r[i] == ref(12345, i) for i in range(64))
    assert 0 <= r.min() and r.max() < 1
    big = oracle.dropout_rand(1 << 16, 99)
    assert abs(big.mean() - 0.5) < 0.01


def test_oracle_training_step_learns(oracle):
    from oracle import ref_train as rt
    B = 4
    net = rt.build_deepcifar10(B, rt.deepcifar10_params(seed=2), learning_rate=0.01)
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, (B, 3, 32, 32)).astype(np.float32)
    t = np.zeros((B, 10), np.float32)
    t[np.arange(B), [1, 3, 5, 7]] = 1
    costs = []
    for it in range(6):
        for i, l in enumerate(net.layers):
            if isinstance(l, rt.RefDropout):
                l.seed = 31 * it + i
        costs.append(net.propagate(x.copy(), t) / B)
        net.update()
    # iteration 0 leaves the conv stack untouched (unassigned dropout.delta, see ref_train.RefDropout)
    assert costs[-1] < 0.2 * costs[0], costs


def test_oracle_lenet_mnist_learns(oracle):
    """SURVEY config 1 (leNetMNIST, nmodels.pas:50-64): shapes chain 28 -> 14 -> 10 -> 5 -> 1 and one batch is fitted."""
    from oracle import ref_train as rt
    B = 8
    net = rt.build_lenet_mnist(B, rt.lenet_mnist_params(seed=1), learning_rate=0.05)
    assert [l.output.shape[1:] for l in net.layers[:5]] == [(6, 28, 28), (6, 14, 14), (16, 10, 10), (16, 5, 5), (120, 1, 1)]
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32)
    t = np.zeros((B, 10), np.float32)
    t[np.arange(B), rng.integers(0, 10, B)] = 1
    costs = []
    for it in range(30):
        costs.append(net.propagate(x.copy(), t) / B)
        net.update()
    assert costs[-1] < 0.5 * costs[0], costs
    assert np.abs(net.layers[0].bias_updates).max() > 0  # no batch norm: the convolution biases learn


def _small_net(rt, orc, batch, P, replicas=1):
    A = orc.ACT
    L = [rt.RefConvTrain(batch, 3, 8, 8, 8, 3, 1, 1, A["RELU"], False, P[0]["w"], P[0]["b"]),
         rt.RefMaxPool(batch, 8, 8, 8, 2),
         rt.RefConnected(batch, 8 * 16, 16, A["RELU"], P[1]["w"], P[1]["b"]),
         rt.RefConnected(batch, 16, 4, A["LINEAR"], P[2]["w"], P[2]["b"]),
         rt.RefSoftmax(batch, 4)]
    return rt.RefNet(L, batch, learning_rate=0.05, momentum=0.9, decay=0.001, replicas=replicas)


def _small_params():
    rng = np.random.default_rng(5)
    return [dict(w=rng.uniform(-.3, .3, (8, 27)).astype(np.float32), b=np.zeros(8, np.float32)),
            dict(w=rng.uniform(-.2, .2, (16, 128)).astype(np.float32), b=np.zeros(16, np.float32)),
            dict(w=rng.uniform(-.3, .3, (4, 16)).astype(np.float32), b=np.zeros(4, np.float32))]


def _small_data(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (n, 3, 8, 8)).astype(np.float32)
    t = np.zeros((n, 4), np.float32)
    t[np.arange(n), rng.integers(0, 4, n)] = 1
    return x, t


def _dp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc, ref_train as rt
    from tensorium_b200.dataparallel import GradientBucket
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc.set_threads(1)
    B = 4
    net = _small_net(rt, orc, B, _small_params(), replicas=world)
    grads = [(l, n) for l in net.layers if l.trainable for n in ("weight_updates", "bias_updates")]
    bucket = GradientBucket([getattr(l, n).size for l, n in grads], "cpu")
    for i, (l, n) in enumerate(grads):  # alias every update buffer into the flat bucket
        setattr(l, n, bucket.view(i).numpy().reshape(getattr(l, n).shape))
    for it in range(4):
        x, t = _small_data(B * world, 40 + it)
        net.propagate(x[rank * B:(rank + 1) * B].copy(), t[rank * B:(rank + 1) * B])
        bucket.all_reduce()
        net.update()
    out[rank] = {f"{i}.{n}": getattr(l, n).copy() for i, l in enumerate(net.layers) if l.trainable
                 for n in ("weights", "biases", "weight_updates")}
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_bucket_data_parallel_matches_single_process_gloo(oracle):
    import torch.multiprocessing as mp
    from oracle import ref_train as rt
    world, B = 2, 4
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_dp_worker, args=(world, 29633, out), nprocs=world, join=True)
        res = {r: dict(v) for r, v in dict(out).items()}
    single = _small_net(rt, oracle, B * world, _small_params())
    for it in range(4):
        x, t = _small_data(B * world, 40 + it)
        single.propagate(x.copy(), t)
        single.update()
    for i, l in enumerate(single.layers):
        if not l.trainable:
            continue
        for n in ("weights", "biases"):
            a, b = res[0][f"{i}.{n}"], res[1][f"{i}.{n}"]
            assert np.array_equal(a, b), (i, n)  # replicas stay bit-identical
            assert np.allclose(a, getattr(l, n), rtol=2e-5, atol=2e-6), (i, n)  # and follow the big-batch run
        # the carried momentum: each replica holds 1/world of the big-batch buffer
        assert np.allclose(res[0][f"{i}.weight_updates"] * world, l.weight_updates, rtol=2e-4, atol=2e-6), i


def test_bucket_layout():
    from tensorium_b200.dataparallel import GradientBucket
    b = GradientBucket([5, 64, 1, 130], "cpu")
    assert b.offsets == [0, 64, 128, 192] and b.flat.numel() == 192 + 192
    assert all((b.ptr(i) - b.ptr(0)) % 256 == 0 for i in range(4))
    b.view(2)[0] = 7
    assert b.flat[128] == 7 and b.all_reduce() is None  # world 1: no-op


def test_oracle_lstm_matches_float64_formula(oracle):
    """RefLSTM (nlstmlayer.pas:518-611) against the textbook recurrence in float64."""
    from oracle import ref_train as rt
    steps, B, n_in, n_out = 5, 3, 7, 6
    P = rt.lstm_params(n_in, n_out, seed=4)
    x = np.random.default_rng(1).uniform(-1, 1, (steps * B, n_in)).astype(np.float32)
    got = rt.RefLSTM(steps * B, n_in, n_out, steps, P).forward(x).copy()
    sig = lambda v: 1 / (1 + np.exp(-v))
    h = np.zeros((B, n_out)); c = np.zeros((B, n_out))
    want = np.zeros((steps * B, n_out))
    for t in range(steps):
        xt = x[t * B:(t + 1) * B].astype(np.float64)
        z = {g: h @ P["w" + g]["w"].T.astype(np.float64) + P["w" + g]["b"] + xt @ P["u" + g]["w"].T.astype(np.float64)
             + P["u" + g]["b"] for g in "figo"}
        c = c * sig(z["f"]) + sig(z["i"]) * np.tanh(z["g"])
        h = np.tanh(c) * sig(z["o"])
        want[t * B:(t + 1) * B] = h
    assert np.abs(got - want).max() <= 2e-6


def test_oracle_lstm_backward_matches_finite_differences(oracle):
    """RefLSTM.backward (nlstmlayer.pas:621-841 restated sub-layer by sub-layer) with the textbook initial state
    (stale_first_step=False) against central differences of the float64 recurrence; the reference-literal variant
    differs from it only through step 0's stale prevCell / prevState."""
    from oracle import ref_train as rt
    steps, B, n_in, H = 3, 2, 4, 5
    P = rt.lstm_params(n_in, H, seed=4)
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (steps * B, n_in)).astype(np.float32)
    gout = rng.uniform(-0.5, 0.5, (steps * B, H)).astype(np.float32)

    def loss(P64, x64):
        sig = lambda v: 1 / (1 + np.exp(-v))
        h = np.zeros((B, H)); c = np.zeros((B, H)); L = 0.0
        for t in range(steps):
            xt = x64[t * B:(t + 1) * B]
            z = {g: h @ P64["w" + g]["w"].T + P64["w" + g]["b"] + xt @ P64["u" + g]["w"].T + P64["u" + g]["b"] for g in "figo"}
            c = c * sig(z["f"]) + sig(z["i"]) * np.tanh(z["g"])
            h = np.tanh(c) * sig(z["o"])
            L += (gout[t * B:(t + 1) * B].astype(np.float64) * h).sum()
        return L
    P64 = {k: {kk: vv.astype(np.float64) for kk, vv in v.items()} for k, v in P.items()}
    net = rt.RefLSTM(steps * B, n_in, H, steps, P)
    net.forward(x)
    net.delta[...] = gout
    up = np.zeros((steps * B, n_in), np.float32)
    net.backward(up, stale_first_step=False)
    eps = 1e-5
    for kind, upd, bupd in (("w", net.w_updates, net.bw_updates), ("u", net.u_updates, net.bu_updates)):
        for g in "figo":
            for field, got, idx in (("w", upd[g], (1, 2)), ("b", bupd[g], (3,))):
                Pp = {k: {kk: vv.copy() for kk, vv in v.items()} for k, v in P64.items()}
                Pp[kind + g][field][idx] += eps
                a = loss(Pp, x.astype(np.float64))
                Pp[kind + g][field][idx] -= 2 * eps
                b = loss(Pp, x.astype(np.float64))
                assert abs(float(got[idx]) - (a - b) / (2 * eps)) < 2e-6, (kind, g, field)
    x64 = x.astype(np.float64)
    for idx in ((0, 1), (3, 2), (5, 0)):
        xp = x64.copy(); xp[idx] += eps; a = loss(P64, xp)
        xp[idx] -= 2 * eps; b = loss(P64, xp)
        assert abs(float(up[idx]) - (a - b) / (2 * eps)) < 2e-6
    # the literal restatement: only the hidden-side updates see step 0's stale prevState / prevCell
    net2 = rt.RefLSTM(steps * B, n_in, H, steps, P)
    net2.forward(x)
    net2.delta[...] = gout
    up2 = np.zeros_like(up)
    net2.backward(up2)
    assert all(np.allclose(net2.u_updates[g], net.u_updates[g], atol=1e-6) for g in "igo")
    assert any(not np.allclose(net2.w_updates[g], net.w_updates[g], atol=1e-6) for g in "figo")
