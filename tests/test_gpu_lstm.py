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
