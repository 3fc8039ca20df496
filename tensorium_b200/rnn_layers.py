"""Host-side mirror of the recurrent layer next to the SGEMM path (SURVEY.md §8f rank 4, config 5 char-LSTM):

  TLSTMLayer   source/nlstmlayer.pas:42-82 (constructor), :876-996 (device path, written against `ocl.*`; the same
               operators exist on `cuda.*`)

Forward only in this round.  `fused=False` issues the reference's own operator sequence (8 connected-layer forwards +
4 addvv + 4 activations + mulvv + fmavv + copy + activation + mulvv + 2 copies per step: 28 launches).  `fused=True`
(default) is the B200-first ordering of the same arithmetic, 2 launches per step:
  * the input projections do not depend on the recurrence: ONE [steps*batch, inputs] x [inputs, 4*outputs] GEMM up
    front over the four gates' stacked weights (large enough for the TMA-fed tensor-core SGEMM) + one bias pass;
  * per step one [batch, outputs] x [outputs, 4*outputs] GEMM on the hidden state and one `lstmGates` launch for
    everything element-wise (`tnn_lstm_gates`).
The eight TConnectedLayer objects stay the owners of the parameters; `pack()` refreshes the stacked copies.
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
        if fused:
            self.w_all = TSingleTensor((4 * outputs, outputs), groups=1)
            self.u_all = TSingleTensor((4 * outputs, inputs), groups=1)
            self.bw_all = TSingleTensor((4 * outputs,), groups=1)
            self.bu_all = TSingleTensor((4 * outputs,), groups=1)
            self.wz = TSingleTensor((B, 4 * outputs), groups=B, host=False)
            self.uz = TSingleTensor((B * steps, 4 * outputs), groups=B * steps, host=False)
            self.pack()
        self.reset()

    def pack(self):
        """Stack the four gates' parameters (f | i | g | o) for the fused path; call again after changing them."""
        host = lambda t: t.pullFromDevice() if t.wasGPU() else t.Data
        for dst, src in ((self.w_all, [self.w[g].weights for g in GATES]), (self.u_all, [self.u[g].weights for g in GATES]),
                         (self.bw_all, [self.w[g].biases for g in GATES]), (self.bu_all, [self.u[g].biases for g in GATES])):
            dst.Data[...] = np.concatenate([host(t).reshape(dst.Data.shape[0] // 4, -1) for t in src]).reshape(dst.Data.shape)
            dst.pushToDevice()

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
            H4 = 4 * self.outputs
            # all steps of the input projections at once: rows [t*B, (t+1)*B) of uz are step t's
            cu.gemm(False, True, B * self.steps, H4, self.inputs, 1.0, state.input.devData, 0, self.inputs,
                    self.u_all.devData, 0, self.inputs, 0.0, self.uz.devData, 0, H4)
            cu.forwardBias(B * self.steps * H4, self.uz.devData, 0, H4, self.bu_all.devData, 1, B * self.steps)
            for i in range(self.steps):
                cu.gemm(False, True, B, H4, self.outputs, 1.0, self._h.devData, 0, self.outputs, self.w_all.devData, 0,
                        self.outputs, 0.0, self.wz.devData, 0, H4)
                cu.lstmGates(B, self.outputs, self.wz.devData, self.bw_all.devData, self.uz.devData, i * B * H4,
                             self._c.devData, self._h.devData, self.Cell.devData, self.output.devData, i * n)
            for t in (self.output, self.Cell, self._c, self._h):
                t.setCUDA()
            return
        for i in range(self.steps):
            offset = i * n
            s.step, s.inputStep, s.input = i, 0, self._h
            for g in GATES:
                self.w[g].forwardGPU(s)
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
