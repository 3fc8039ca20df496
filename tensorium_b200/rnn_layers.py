"""Host-side mirror of the recurrent layer next to the SGEMM path (SURVEY.md §8f rank 4, config 5 char-LSTM):

  TLSTMLayer   source/nlstmlayer.pas:42-82 (constructor), :876-996 (device path, written against `ocl.*`; the same
               operators exist on `cuda.*`)

`fused=False` (forward only) issues the reference's own operator sequence (8 connected-layer forwards +
4 addvv + 4 activations + mulvv + fmavv + copy + activation + mulvv + 2 copies per step: 28 launches).  `fused=True`
(default) is the B200-first ordering of the same arithmetic, 2 launches per step:
  * the input projections do not depend on the recurrence: ONE [steps*batch, inputs] x [inputs, 4*outputs] GEMM up
    front over the four gates' stacked weights (large enough for the TMA-fed tensor-core SGEMM) + one bias pass;
  * per step one [batch, outputs] x [outputs, 4*outputs] GEMM on the hidden state and one `lstmGates` launch for
    everything element-wise (`tnn_lstm_gates`).
The eight TConnectedLayer objects stay the owners of the parameters; `pack()` refreshes the stacked copies.

Training (`train=True`, fused): TLSTMLayer.backward / update (nlstmlayer.pas:621-861) in the same B200-first ordering.
Back-propagation through time needs per step only what depends on the recurrence — one `lstmGatesBackward` launch
(all element-wise work + the sub-layers' delta clamp) and one [batch, 4*outputs] x [4*outputs, outputs] GEMM that
adds W^T dz into the previous step's delta; everything else is batched over ALL steps after the loop: the weight
updates of the four hidden-side layers (dz^T h_prev) and of the four input-side layers (dz^T x) as two stacked
GEMMs with K = steps*batch, the bias updates as one column sum, the upstream delta as one GEMM.  The stacked
tensors are the masters while training (`unpack()` copies them back into the eight sub-layers).  The reference's
stale prevCell / prevState at step 0 (nlstmlayer.pas:664-669 refresh them only `if i <> 0`) is mirrored.
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

    def __init__(self, batch_total, inputs, outputs, steps, fused=True, rng=None, params=None, train=False):
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
        self.train = bool(train)
        if train:
            assert fused, "training runs the stacked (fused) ordering"
            S = steps
            self.wz_all = TSingleTensor((B * S, 4 * outputs), groups=B * S, host=False)   # hidden-side pre-activations
            self.dz_all = TSingleTensor((B * S, 4 * outputs), groups=B * S, host=False)   # gate pre-activation deltas
            self.delta = dev(B * S)                                                        # dL/dh of every step
            self.dc = dev(B)
            zeros = lambda t: TSingleTensor(t.FShape, groups=1)
            self.w_all_updates, self.u_all_updates = zeros(self.w_all), zeros(self.u_all)
            self.bw_all_updates, self.bu_all_updates = zeros(self.bw_all), zeros(self.bu_all)
            for t in (self.w_all_updates, self.u_all_updates, self.bw_all_updates, self.bu_all_updates):
                t.pushToDevice()
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
            self._input = state.input
            for i in range(self.steps):
                # training keeps every step's hidden-side pre-activations for the backward pass
                wz, woff = (self.wz_all.devData, i * B * H4) if self.train else (self.wz.devData, 0)
                cu.gemm(False, True, B, H4, self.outputs, 1.0, self._h.devData, 0, self.outputs, self.w_all.devData, 0,
                        self.outputs, 0.0, wz, woff, H4)
                cu.lstmGates(B, self.outputs, wz + 4 * woff, self.bw_all.devData, self.uz.devData, i * B * H4,
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

    # nlstmlayer.pas:621-841 (fused ordering, see the module docstring)
    def backwardGPU(self, state: TNNetState, net=None):
        assert self.train and self.fused
        cu = cuda()
        B, H, S = self.batch, self.outputs, self.steps
        n, H4 = B * H, 4 * H
        x = state.input if state.input is not None else self._input
        cu.fill(n, self.dc.devData, 0, 0.0, 1)
        for i in range(S - 1, -1, -1):
            prev = (i - 1) * n if i else (0 if S > 1 else -1)   # step 0: the reference's stale prevCell = cell[0]
            cu.lstmGatesBackward(B, H, self.wz_all.devData, self.bw_all.devData, self.uz.devData, i * B * H4,
                                 self.Cell.devData, i * n, prev, self.delta.devData, i * n, self.dc.devData,
                                 self.dz_all.devData, i * B * H4)
            if i:  # wf/wi/wg/wo.backward: delta(t-1) += dz(t) W     (nconnectedlayer.pas:306-313)
                cu.gemm(False, False, B, H, H4, 1.0, self.dz_all.devData, i * B * H4, H4, self.w_all.devData, 0, H, 1.0,
                        self.delta.devData, (i - 1) * n, H)
        dz = self.dz_all.devData
        # hidden-side weight updates: sum_t dz(t)^T h(t-1), with h(-1) := output[0] like the reference's stale prevState
        if S > 1:
            cu.gemm(True, False, H4, H, (S - 1) * B, 1.0, dz, B * H4, H4, self.output.devData, 0, H, 1.0,
                    self.w_all_updates.devData, 0, H)
            cu.gemm(True, False, H4, H, B, 1.0, dz, 0, H4, self.output.devData, 0, H, 1.0, self.w_all_updates.devData, 0, H)
        # input-side weight updates: dz^T x over all steps at once
        cu.gemm(True, False, H4, self.inputs, S * B, 1.0, dz, 0, H4, x.devData, 0, self.inputs, 1.0,
                self.u_all_updates.devData, 0, self.inputs)
        # both bias vectors receive the column sums of dz (addSums, nconnectedlayer.pas:269-270)
        for bu in (self.bw_all_updates, self.bu_all_updates):
            cu.backwardBias(H4, bu.devData, S * B * H4, dz, 0, 1, S * B)
        if state.delta is not None and state.delta.devData:
            cu.gemm(False, False, S * B, self.inputs, H4, 1.0, dz, 0, H4, self.u_all.devData, 0, self.inputs, 1.0,
                    state.delta.devData, 0, self.inputs)
            state.delta.setCUDA()
        for t in (self.w_all_updates, self.u_all_updates, self.bw_all_updates, self.bu_all_updates, self.delta):
            t.setCUDA()

    # nlstmlayer.pas:843-861: the eight sub-layers' update (nconnectedlayer.pas:324-359) on the stacked tensors
    def updateGPU(self, args):
        cu = cuda()
        f32 = np.float32
        rate = float(f32(f32(args.learningRate) / f32(args.batch)))
        decay_term = float(f32(f32(-args.decay) * f32(args.batch)))
        keep = args.momentum / getattr(args, "replicas", 1)
        cu.sgdUpdate(self.bw_all.Size(), self.bw_all.devData, self.bw_all_updates.devData, rate, keep, 0.0)
        cu.sgdUpdate(self.bu_all.Size(), self.bu_all.devData, self.bu_all_updates.devData, rate, keep, 0.0)
        cu.sgdUpdate(self.w_all.Size(), self.w_all.devData, self.w_all_updates.devData, rate, keep, decay_term)
        cu.sgdUpdate(self.u_all.Size(), self.u_all.devData, self.u_all_updates.devData, rate, keep, decay_term)
        for t in (self.w_all, self.u_all, self.bw_all, self.bu_all):
            t.setCUDA()

    def gradients(self):
        return [self.w_all_updates, self.u_all_updates, self.bw_all_updates, self.bu_all_updates]

    def unpack(self):
        """Copy the stacked (master) parameters back into the eight sub-layers' host tensors."""
        H = self.outputs
        for src, dst in ((self.w_all, [self.w[g].weights for g in GATES]), (self.u_all, [self.u[g].weights for g in GATES]),
                         (self.bw_all, [self.w[g].biases for g in GATES]), (self.bu_all, [self.u[g].biases for g in GATES])):
            host = src.pullFromDevice() if src.wasGPU() else src.Data
            src.setCUDA()
            for k, t in enumerate(dst):
                t.Data[...] = host[k * H:(k + 1) * H].reshape(t.Data.shape)
                t.setCPU()
