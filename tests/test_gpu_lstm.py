"""TLSTMLayer forward on the GPU backend (tensorium_b200/rnn_layers.py; nlstmlayer.pas:876-996) against the oracle's
restatement of the CPU path, in the reference's operator order and in the up-front-projection order."""
import numpy as np
import pytest

from util import TOL_3XTF32, TOL_TF32, rel_max

pytestmark = pytest.mark.gpu
FP32, X3, TF32 = 0, 1, 2


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("mode,tol", [(FP32, 1e-5), (X3, TOL_3XTF32), (TF32, TOL_TF32)])
@pytest.mark.parametrize("steps,B,n_in,n_out", [(6, 4, 24, 32), (16, 32, 96, 256), (3, 5, 10, 12)])
def test_lstm_forward_matches_oracle(cuda_backend, oracle, fused, mode, tol, steps, B, n_in, n_out):
    from oracle import ref_train as rt
    from tensorium_b200.layers import TNNetState
    from tensorium_b200.rnn_layers import TLSTMLayer
    from tensorium_b200.tensor import TSingleTensor
    cu = cuda_backend
    cu.set_math_mode(mode)
    try:
        P = rt.lstm_params(n_in, n_out, seed=steps)
        x = np.random.default_rng(B).uniform(-1, 1, (steps * B, n_in)).astype(np.float32)
        ref = rt.RefLSTM(steps * B, n_in, n_out, steps, P)
        want = ref.forward(x).copy()
        cell = ref.cell.copy()
        l = TLSTMLayer(steps * B, n_in, n_out, steps, fused=fused, params=P)
        st = TNNetState(input=TSingleTensor(x.shape, steps * B, data=x))
        l.forwardGPU(st)
        assert rel_max(l.output.pullFromDevice(), want) <= tol
        assert rel_max(l.Cell.pullFromDevice(), cell) <= tol
        # the state carries over: a second call continues from (c, h), a reset starts over
        l.forwardGPU(st)
        want2 = ref.forward(x).copy()
        assert rel_max(l.output.pullFromDevice(), want2) <= 2 * tol
        l.reset()
        l.forwardGPU(st)
        assert rel_max(l.output.pullFromDevice(), want) <= tol
    finally:
        cu.set_math_mode(X3)


@pytest.mark.parametrize("mode,tol", [(FP32, 2e-5), (X3, TOL_3XTF32)])
@pytest.mark.parametrize("steps,B,n_in,n_out", [(5, 4, 24, 32), (8, 16, 48, 64), (1, 3, 8, 16)])
def test_lstm_training_step_matches_oracle(cuda_backend, oracle, mode, tol, steps, B, n_in, n_out):
    """TLSTMLayer.backward / update (nlstmlayer.pas:621-861) in the stacked B200 ordering against the oracle's literal
    restatement (per-step, per-sub-layer): weight / bias updates of the eight sub-layers, the upstream delta, the
    delta written into earlier steps, then the SGD update."""
    from oracle import ref_train as rt
    from tensorium_b200.layers import TNNetState, TUpdateArgs
    from tensorium_b200.rnn_layers import GATES, TLSTMLayer
    from tensorium_b200.tensor import TSingleTensor
    cu = cuda_backend
    cu.set_math_mode(mode)
    try:
        rng = np.random.default_rng(steps * 7 + B)
        P = rt.lstm_params(n_in, n_out, seed=steps)
        x = rng.uniform(-1, 1, (steps * B, n_in)).astype(np.float32)
        dh = rng.uniform(-0.3, 0.3, (steps * B, n_out)).astype(np.float32)
        up0 = rng.uniform(-0.1, 0.1, (steps * B, n_in)).astype(np.float32)
        ref = rt.RefLSTM(steps * B, n_in, n_out, steps, P)
        ref.forward(x)
        ref.delta[...] = dh
        up_ref = up0.copy()
        ref.backward(up_ref)
        l = TLSTMLayer(steps * B, n_in, n_out, steps, params=P, train=True)
        st = TNNetState(input=TSingleTensor(x.shape, steps * B, data=x), isTraining=True)
        l.forwardGPU(st)
        l.delta.Data[...] = dh
        l.delta.pushToDevice()
        st.delta = TSingleTensor(up0.shape, steps * B, data=up0)
        st.delta.pushToDevice()
        l.backwardGPU(st)
        stack = lambda d: np.concatenate([d[g] for g in GATES])
        scale = lambda a: max(np.abs(a).max(), 1e-6)
        assert rel_max(l.w_all_updates.pullFromDevice(), stack(ref.w_updates)) <= 4 * tol
        assert rel_max(l.u_all_updates.pullFromDevice(), stack(ref.u_updates)) <= 4 * tol
        assert rel_max(l.bw_all_updates.pullFromDevice(), stack(ref.bw_updates)) <= 4 * tol
        assert rel_max(l.bu_all_updates.pullFromDevice(), stack(ref.bu_updates)) <= 4 * tol
        assert rel_max(st.delta.pullFromDevice(), up_ref) <= 4 * tol
        assert rel_max(l.delta.pullFromDevice(), ref.delta) <= 4 * tol    # earlier steps received W^T dz
        for t in (l.w_all_updates, l.u_all_updates, l.bw_all_updates, l.bu_all_updates):
            t.setCUDA()
        l.updateGPU(TUpdateArgs(batch=steps * B, learningRate=0.05, momentum=0.9, decay=0.0005))
        ref.update(0.05, 0.9, 0.0005, steps * B)
        l.unpack()
        for g in GATES:
            assert rel_max(l.w[g].weights.Data, ref.w[g]) <= tol and rel_max(l.u[g].weights.Data, ref.u[g]) <= tol
            assert rel_max(l.w[g].biases.Data, ref.bw[g]) <= 4 * tol and rel_max(l.u[g].biases.Data, ref.bu[g]) <= 4 * tol
    finally:
        cu.set_math_mode(X3)
