"""Host-side mirror of the recurrent layer next to the SGEMM path (SURVEY.md §8f rank 4, config 5 char-LSTM):

  TLSTMLayer   source/nlstmlayer.pas:42-82 (constructor), :876-996 (device path, written against `ocl.*`; the same
               operators exist on `cuda.*`)

Forward only in this round.  `fused=False` issues the reference's own operator sequence (8 connected-layer forwards +
4 addvv + 4 activations + mulvv + fmavv + copy + activation + mulvv + 2 copies per step).  `fused=True` (default) is the
B200-first ordering of the same arithmetic: the four input projections do not depend on the recurrence, so they are
computed for ALL steps up front as four [steps*batch, inputs] x [inputs, outputs] GEMMs — large enough for the TMA-fed
tensor-core SGEMM — and only the four hidden-state GEMMs stay inside the step loop.
"""
from __future__ import annotations

import numpy as np

from . import nncuda as nn
from .layers import TNNetState
from .tensor import TSingleTensor, cuda
from .train_layers import TConnectedLayer

ltLSTM = "lstm"
GATES = ("f", "i", "g", "o")


class TLSTMLayer:
    layerType = ltLSTM
    trainable = True

    def __init__(self, batch_total, inputs, outputs, steps, fused=True, rng=None, params=None):
        assert batch_total % steps == 0
        self.steps, self.inputs, self.outputs = steps, inputs, outputs
        self.batch = batch_total // steps  # nlstmlayer.pas:44
        self.fused = fused
        rng = rng or np.random.default_rng(0)
        B = self.batch
        mk = lambda n_in: TConnectedLayer(B, steps, n_in, outputs, nn.acLINEAR, train=False, rng=rng)
        self.w = {g: mk(outputs) for g in GATES}  # wf wi wg wo: on the hidden state
        self.u = {g: mk(inputs) for g in GATES}   # uf ui ug uo: on the step's input
        if params is not None:
            for kind, d in (("w", self.w), ("u", self.u)):
                for g in GATES:
                    d[g].weights.Data[...] = params[kind + g]["w"].reshape(d[g].weights.FShape)
                    d[g].biases.Data[...] = params[kind + g]["b"]
        dev = lambda rows: TSingleTensor((rows, outputs), groups=rows, host=False)
        self.output, self.Cell = dev(B * steps), dev(B * steps)
        self._f, self._i, self._g, self._o, self._c, self._h, self.temp = (dev(B) for _ in range(7))
        self.outC, self.outH, self.outW = outputs, 1, 1
        self.reset()

    def reset(self):
        """Zero the carried cell and hidden state (fresh tensors are zero in the reference)."""
        cu = cuda()
        for t in (self._c, self._h):
            cu.fill(t.Size(), t.devData, 0, 0.0, 1)
            t.setCUDA()

    # nlstmlayer.pas:876-996
    def forwardGPU(self, state: TNNetState, net=None):
        cu = cuda()
        B, n = self.batch, self.batch * self.outputs
        if not state.input.wasGPU():
            state.input.pushToDevice()
        s = TNNetState(isTraining=state.isTraining, workspace=state.workspace, index=state.index)
        gate = dict(f=self._f, i=self._i, g=self._g, o=self._o)
        if self.fused:
            # all steps of the input projections at once: output rows [t*B, (t+1)*B) are step t's, as in the loop
            for g in GATES:
                l = self.u[g]
                l._push(l.weights, l.biases)
                l.output.setCUDA()
                cu.gemm(False, True, B * self.steps, self.outputs, self.inputs, 1.0, state.input.devData, 0,
                        self.inputs, l.weights.devData, 0, self.inputs, 0.0, l.output.devData, 0, self.outputs)
                cu.forwardBias(n * self.steps, l.output.devData, 0, l.biases.Size(), l.biases.devData, 1,
                               B * self.steps)
        for i in range(self.steps):
            offset = i * n
            s.step, s.inputStep, s.input = i, 0, self._h
            for g in GATES:
                self.w[g].forwardGPU(s)
            if not self.fused:
                s.inputStep, s.input = i, state.input
                for g in GATES:
                    self.u[g].forwardGPU(s)
            for g in GATES:
                cu.addvv(n, self.w[g].output.devData, offset, 1, self.u[g].output.devData, offset, 1,
                         gate[g].devData, 0, 1)
            for g in ("f", "i", "o"):
                cu.ActivateArray(n, gate[g].devData, 0, nn.acLOGISTIC)
            cu.ActivateArray(n, self._g.devData, 0, nn.acTANH)
            cu.mulvv(n, self._i.devData, 0, 1, self._g.devData, 0, 1, self.temp.devData, 0, 1)
            cu.fmavv(n, self._c.devData, 0, 1, self._f.devData, 0, 1, self.temp.devData, 0, 1, self._c.devData, 0, 1)
            cu.copy(n, self._c.devData, 0, 1, self._h.devData, 0, 1)
            cu.ActivateArray(n, self._h.devData, 0, nn.acTANH)
            cu.mulvv(n, self._o.devData, 0, 1, self._h.devData, 0, 1, self._h.devData, 0, 1)
            cu.copy(n, self._c.devData, 0, 1, self.Cell.devData, offset, 1)
            cu.copy(n, self._h.devData, 0, 1, self.output.devData, offset, 1)
        for t in (self.output, self.Cell, self._c, self._h):
            t.setCUDA()
