#!/usr/bin/env python
"""Generate pascal/nnCuda.pas — the FreePascal/Delphi unit that replaces the reference's source/nncuda.pas — from
include/tensorium_b200.h, so the binding cannot drift from the C ABI.

* one `external` declaration per exported symbol of the header (types mapped mechanically);
* class TNNCuda<T> with EVERY public method of the reference class (source/nncuda.pas:100-158), unchanged
  signatures, each body one SAFE_CALL(tnn_xxx(FCtx, ...)) (SAFE_CALL as in source/nncuda.pas:237-274: assert on the
  return code with the library's message), plus `finish()` under USE_TELEMETRY like the reference wrappers
  (nncuda.pas:596-599); the NVRTC entry points (compileToCUBIN / loadCUBIN / compileFile / loadCUBinFile) are no-ops
  because the library carries precompiled sm_100a code;
* the new coarse entry points (fused convolution, batch-norm, SGD, NHWC graph, graphs/events, NCCL exchange) as
  additional methods.

  python tools/gen_pascal.py            # rewrite pascal/nnCuda.pas
  python tools/gen_pascal.py --check    # exit 1 if the committed file is stale

FreePascal is not installed in the build container, so the unit is not compiled here; tests/test_pascal_unit.py checks
it structurally against the header (every symbol bound once, arities equal, every reference method present).
"""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tensorium_b200.h")
OUT = os.path.join(ROOT, "pascal", "nnCuda.pas")

CTYPES = [  # (regex on the C type, Pascal type) — first match wins
    (r"^tnn_context \*\*$", "PPointer"), (r"^tnn_context \*$", "TNNContext"),
    (r"^(const )?tnn_conv_desc \*$", "PNNConvDesc"),
    (r"^tnn_event \*\*$", "PPointer"), (r"^tnn_event \*$", "pointer"),
    (r"^tnn_graph \*\*$", "PPointer"), (r"^tnn_graph \*$", "pointer"),
    (r"^tnn_comm \*\*$", "PPointer"), (r"^tnn_comm \*$", "pointer"),
    (r"^const float \*const \*$", "PPointer"), (r"^float \*const \*$", "PPointer"), (r"^float \*\*$", "PPointer"),
    (r"^(const )?float \*$", "pointer"), (r"^(const )?void \*$", "pointer"), (r"^void \*\*$", "PPointer"),
    (r"^(const )?uint64_t \*$", "PQWord"), (r"^int32_t \*$", "PLongint"), (r"^(const )?char \*$", "PAnsiChar"),
    (r"^int32_t$", "longint"), (r"^int64_t$", "int64"), (r"^uint64_t$", "QWord"), (r"^size_t$", "size_t"),
    (r"^float$", "single"), (r"^void$", ""),
]
RESERVED = {"in", "out", "type", "label", "end", "begin", "to", "set", "file", "string", "array", "var", "const", "function"}


def parse_header(text=None):
    """-> [(ret_ctype, name, [(ctype, argname)])] for every TNN_API declaration."""
    text = text if text is not None else open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for m in re.finditer(r"TNN_API\s+([\w\s\*]+?)\s*\b(tnn_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        ret = re.sub(r"\s*\*", " *", ret).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"^(.*?)(\w+)$", a)
                ctype, an = mm.group(1).strip(), mm.group(2)
                ctype = re.sub(r"\s*\*\s*", " *", ctype).replace("* *", "**").replace("*const", "*const ").strip()
                ctype = re.sub(r"\s+", " ", ctype).replace("*const *", "*const *")
                params.append((ctype, an))
        out.append((ret, name, params))
    return out


def pas_type(ctype):
    c = ctype.replace(" *const *", " *const *").strip()
    c = re.sub(r"\*\s*\*", "**", c)
    for rx, p in CTYPES:
        if re.match(rx, c):
            return p
    raise ValueError(f"no Pascal mapping for C type '{ctype}'")


def pas_name(n):
    return "a" + n[0].upper() + n[1:] if n.lower() in RESERVED else n


def external_decl(ret, name, params):
    ps = "; ".join(f"{pas_name(n)}: {pas_type(t)}" for t, n in params)
    rt = pas_type(ret)
    head = f"function {name}({ps}): {rt}" if rt else f"procedure {name}({ps})"
    return f"{head}; cdecl; external TNNLIB;"


# ---------------------------------------------------------------------------------------------------------------
# Class methods.  (pascal declaration, C symbol, argument expressions after FCtx, telemetry?)  The declarations of
# the first block are the reference's own (source/nncuda.pas:100-158) — the drop-in keeps them verbatim so that
# ntensors.pas and every layer compile unchanged.
M = []


def meth(decl, sym, args, telemetry=True):
    M.append((decl, sym, args, telemetry))


meth("procedure ActivateArray(const N: SizeInt; const x: TCUMem; const offset: SizeInt; const activation: longint)",
     "tnn_activate_array", "N, x, offset, activation")
meth("procedure activateArraySWISH(const N: SizeInt; const x: TCUMem; const offset: SizeInt; const output_sigmoid, output: TCUMem)",
     "tnn_activate_array_swish", "N, x, offset, output_sigmoid, output")
meth("procedure DeriveArray(const N: SizeInt; const x: TCUMem; const offset:SizeInt; const activation: longint; delta: TCUMem)",
     "tnn_derive_array", "N, x, offset, activation, delta")
meth("procedure forwardBias(const dstSize: SizeInt; const dst: TCUMem; const offset:SizeInt; const srcSize: SizeInt; const src: TCUMem; const incb: SizeInt; const batch: SizeInt)",
     "tnn_forward_bias", "dstSize, dst, offset, srcSize, src, incb, batch")
meth("procedure backwardBias(const dstSize: SizeInt; const dst: TCUMem; const srcSize: SizeInt; const src: TCUMem; const srcOffset:SizeInt; const incb: SizeInt ; const batch: SizeInt)",
     "tnn_backward_bias", "dstSize, dst, srcSize, src, srcOffset, incb, batch")
meth("procedure gemm(const transA, transB :boolean; const M, N, K:SizeInt; const ALPHA:T; const A:TCUMem; const aOffset:SizeInt; const lda:SizeInt; const B:TCUMem; const bOffset:SizeInt; const ldb:SizeInt; const BETA: T; const C:TCUMem; const cOffset:SizeInt; const ldc:SizeInt)",
     "tnn_gemm", "ord(transA), ord(transB), M, N, K, ALPHA, A, aOffset, lda, B, bOffset, ldb, BETA, C, cOffset, ldc")
meth("procedure gemmBatched(const transA, transB: boolean; const M, N, K: SizeInt; const ALPHA: T; const A: PCUmem; const aOffset: SizeInt; const lda: SizeInt; const B: PCUMem; const bOffset: SizeInt; const ldb: SizeInt; const BETA: T; const C: PCUMem; const cOffset: SizeInt; const ldc: SizeInt; const batchCount: SizeInt)",
     "tnn_gemm_batched", "ord(transA), ord(transB), M, N, K, ALPHA, PPointer(A), aOffset, lda, PPointer(B), bOffset, ldb, BETA, PPointer(C), cOffset, ldc, batchCount")
meth("procedure gemmStridedBatched(const transA, transB :boolean; const M, N, K:SizeInt; const ALPHA:T; A:TCUMem; const aOffset:SizeInt; const lda:SizeInt; const strideA:SizeInt; B:TCUMem; const bOffset:SizeInt; const ldb:SizeInt; const strideB:SizeInt; const BETA: T; C:TCUMem; const cOffset:SizeInt; const ldc:SizeInt; const strideC:SizeInt; const batchCount : SizeInt)",
     "tnn_gemm_strided_batched", "ord(transA), ord(transB), M, N, K, ALPHA, A, aOffset, lda, strideA, B, bOffset, ldb, strideB, BETA, C, cOffset, ldc, strideC, batchCount")
for nm in ("addvv", "subvv", "mulvv"):
    meth(f"procedure {nm}(const N:SizeInt; const src1:TCUMem; const src1Offset, inca:SizeInt; const src2:TCUMem; const src2Offset, incb:SizeInt; dst:TCUMem; const dstOffset, incc:SizeInt)",
         f"tnn_{nm}", "N, src1, src1Offset, inca, src2, src2Offset, incb, dst, dstOffset, incc")
meth("procedure fmavv(const N: SizeInt; const src1: TCUMem; const src1Offset, inca: SizeInt; const src2: TCUMem; const src2Offset, incb: SizeInt; const src3: TCUMem; const src3Offset, incc: SizeInt; dst: TCUMem; const dstOffset, incd: SizeInt)",
     "tnn_fmavv", "N, src1, src1Offset, inca, src2, src2Offset, incb, src3, src3Offset, incc, dst, dstOffset, incd")
meth("procedure axpy(const N:SizeInt; const a:T; const x:TCUMem; const xOffset:SizeInt; const incx:SizeInt; const y:TCUMem; const yOffset:SizeInt; const incy:sizeInt)",
     "tnn_axpy", "N, a, x, xOffset, incx, y, yOffset, incy")
meth("procedure power(const N:SizeInt; const x:TCUMem; const xOffset:SizeInt; const incx:SizeInt; const a:T; const y:TCUMem; const yOffset:SizeInt; const incy:sizeInt)",
     "tnn_power", "N, x, xOffset, incx, a, y, yOffset, incy")
meth("procedure scale(const N:SizeInt; const a:T; const x:TCUMem; const stride:SizeInt)", "tnn_scale", "N, a, x, stride")
meth("procedure crossEntropyLogistic(const N:SizeInt; const pred, truth: TCUMem; delta, error: TCUMem)",
     "tnn_cross_entropy_logistic", "N, pred, truth, delta, error")
meth("procedure fill(const N:SizeInt; const x: TCUMem; const offset:SizeInt; const val:T; const stride :SizeInt)",
     "tnn_fill", "N, x, offset, val, stride")
meth("procedure copy(const N:SizeInt; const src:TCUMem; const srcOffset, inca:SizeInt; const dst:TCUMem; const dstOffset, incb:SizeInt)",
     "tnn_copy", "N, src, srcOffset, inca, dst, dstOffset, incb")
meth("procedure softmaxBatch(const N: SizeInt; const input: TCUMem; const iOffset: SizeInt; const batch, batch_size, groups, group_size, stride: SizeInt; const temp: T; const output: TCUMem; const oOffset: SizeInt)",
     "tnn_softmax_batch", "N, input, iOffset, batch, batch_size, groups, group_size, stride, temp, output, oOffset")
meth("procedure crossEntropySoftmax(const N:SizeInt; const pred, truth: TCUMem; delta, error: TCUMem)",
     "tnn_cross_entropy_softmax", "N, pred, truth, delta, error")
meth("procedure forwardMaxPool(const aBatch, outC, outH, outW: SizeInt; const input: TCUMem; const c, h, w: SizeInt; const stride_x, stride_y, padding, kernelSize: SizeInt; indexes, output: TCUMem)",
     "tnn_forward_maxpool", "aBatch, outC, outH, outW, input, c, h, w, stride_x, stride_y, padding, kernelSize, PQWord(indexes), output", False)
meth("procedure backwardMaxPool(const aBatch, outC, outH, outW : SizeInt; output:TCUMem; const indexes, delta : TCUMem)",
     "tnn_backward_maxpool", "aBatch, outC, outH, outW, output, PQWord(indexes), delta", False)
meth("procedure im2col(const aChannels, aHeight, aWidth , kernelHeight, kernelWidth, padHeight, padWidth , strideY, strideX, dilationY, dilationX : SizeInt ; const im :TCUMem; const imOffset : SizeInt ; const col:TCUMem; const colOffset:SizeInt)",
     "tnn_im2col", "aChannels, aHeight, aWidth, kernelHeight, kernelWidth, padHeight, padWidth, strideY, strideX, dilationY, dilationX, im, imOffset, col, colOffset")
meth("procedure col2im(const aChannels, aHeight, aWidth, kernelHeight, kernelWidth, padHeight, padWidth, strideY, strideX, dilationY, dilationX: SizeInt; const col: TCUMem; const colOffset: SizeInt; const im: TCUMem; const imOffset: SizeInt)",
     "tnn_col2im", "aChannels, aHeight, aWidth, kernelHeight, kernelWidth, padHeight, padWidth, strideY, strideX, dilationY, dilationX, col, colOffset, im, imOffset")
meth("procedure upSample(const aBatch, aChannels, outHeight, outWidth: SizeInt; const &in: TCUMem;const stride: SizeInt; const isForward: longint; const scale: T; const &out: TCUMem; const zeroIn :integer = 0)",
     "tnn_upsample", "aBatch, aChannels, outHeight, outWidth, &in, stride, isForward, scale, &out, zeroIn")
meth("procedure fmavss(const N: SizeInt; const src: TCUMem; const offset: SizeInt; const scalar, bias: T; dst : TCUMem)",
     "tnn_fmavss", "N, src, offset, scalar, bias, dst")
meth("procedure meansAndVars(const srcSize, dstSize, groups:sizeInt; const src:TCUMem; const offset:sizeInt; means, vars:TCUMem)",
     "tnn_means_and_vars", "srcSize, dstSize, groups, src, offset, means, vars")
meth("procedure means(const srcSize, dstSize, groups:sizeInt; const src:TCUMem; const offset:sizeInt; means:TCUMem)",
     "tnn_means", "srcSize, dstSize, groups, src, offset, means")
meth("procedure variances(const srcSize, dstSize, groups:sizeInt; const src:TCUMem; const offset:sizeInt; means, vars:TCUMem)",
     "tnn_variances", "srcSize, dstSize, groups, src, offset, means, vars")
meth("procedure normalize(const srcSize, dstSize, groups:SizeInt; means:TCUMem; const meansStride:sizeInt; vars:TCUMem; const varsStride:SizeInt; dst:TCUMem; const dstOffset :sizeInt)",
     "tnn_normalize", "srcSize, dstSize, groups, means, meansStride, vars, varsStride, dst, dstOffset")
meth("procedure meansAndVarsDelta(const srcSize, dstSize, groups:SizeInt; delta, x: TCUMem; const offset:SizeInt; mean, variance, mean_delta, variance_delta: TCUMem)",
     "tnn_means_and_vars_delta", "srcSize, dstSize, groups, delta, x, offset, mean, variance, mean_delta, variance_delta")
meth("procedure normalizeDelta(const deltaSize, meanSize, groups: SizeInt; const delta, x: TCUMem; const offset:SizeInt; mean, variance, mean_delta, variance_delta: TCUMem)",
     "tnn_normalize_delta", "deltaSize, meanSize, groups, delta, x, offset, mean, variance, mean_delta, variance_delta")
meth("procedure addDots(const N, dstSize, groups:SizeInt; const src1, src2:TCUMem; const srcOffset:SizeInt; dst:TCUMem)",
     "tnn_add_dots", "N, dstSize, groups, src1, src2, srcOffset, dst")
meth("procedure forwardScale(const dstSize: SizeInt; const dst: TCUMem; const offset :SizeInt; const scaleSize: SizeInt; const scale: TCUMem; const incb: SizeInt ; const batch: SizeInt)",
     "tnn_forward_scale", "dstSize, dst, offset, scaleSize, scale, incb, batch")
meth("procedure forwardScaleAdd(const dstSize: SizeInt; const dst: TCUMem; const offset :SizeInt; const scaleSize: SizeInt; const scales, biases: TCUMem; const incb: SizeInt ; const batch: SizeInt)",
     "tnn_forward_scale_add", "dstSize, dst, offset, scaleSize, scales, biases, incb, batch")
meth("procedure forwardDropout(const N: SizeInt; const src: TCUMem; const probability, scale: T; rnd: TCUMem; dst: TCUMem)",
     "tnn_forward_dropout", "N, src, probability, scale, rnd, dst, QWord(random($10000000))")  # nncuda.pas:1521
meth("procedure backwardDropout(const N: SizeInt; const src: TCUMem; const probability, scale: T; const rnd: TCUMem; dst: TCUMem)",
     "tnn_backward_dropout", "N, src, probability, scale, rnd, dst")
meth("procedure costL2(const N:SizeInt; const pred ,truth, delta, error: TCUMem)", "tnn_cost_l2", "N, pred, truth, delta, error")
meth("procedure clamp(const N:SizeInt; const alpha :T; const src, dst: TCUMem; const stride: SizeInt =1; offset : SizeInt =0)",
     "tnn_clamp", "N, alpha, src, dst, stride, offset")
meth("procedure inverseSqrt(const N:SizeInt; const alpha :T; const src, dst: TCUMem; const stride: SizeInt =1; offset : SizeInt =0)",
     "tnn_inverse_sqrt", "N, alpha, src, dst, stride, offset")
N_REFERENCE_METHODS = len(M)

# ---- new coarse entry points (no counterpart in the reference class; used by the edited layer bodies)
meth("procedure conv2dForward(const d: TNNConvDesc; const inp: TCUMem; const inOffset: SizeInt; const weights, scale, bias: TCUMem; const activation: longint; const outp: TCUMem; const outOffset: SizeInt)",
     "tnn_conv2d_forward", "@d, inp, inOffset, weights, scale, bias, activation, outp, outOffset")
meth("procedure conv2dForwardResidual(const d: TNNConvDesc; const inp: TCUMem; const inOffset: SizeInt; const weights, scale, bias: TCUMem; const activation: longint; const residual: TCUMem; const residualOffset: SizeInt; const outp: TCUMem; const outOffset: SizeInt)",
     "tnn_conv2d_forward_residual", "@d, inp, inOffset, weights, scale, bias, activation, residual, residualOffset, outp, outOffset")
meth("procedure conv2dForwardNhwc(const d: TNNConvDesc; const inp: TCUMem; const inOffset, inPitch: SizeInt; const inLayout: longint; const weights, scale, bias: TCUMem; const activation: longint; const outp: TCUMem; const outOffset, outPitch: SizeInt; const outLayout: longint)",
     "tnn_conv2d_forward_nhwc", "@d, inp, inOffset, inPitch, inLayout, weights, scale, bias, activation, outp, outOffset, outPitch, outLayout")
meth("procedure conv2dBackwardFilter(const d: TNNConvDesc; const inp: TCUMem; const inOffset: SizeInt; const delta: TCUMem; const deltaOffset: SizeInt; const weightUpdates: TCUMem)",
     "tnn_conv2d_backward_filter", "@d, inp, inOffset, delta, deltaOffset, weightUpdates")
meth("procedure conv2dBackwardData(const d: TNNConvDesc; const weights, delta: TCUMem; const deltaOffset: SizeInt; const inDelta: TCUMem; const inDeltaOffset: SizeInt)",
     "tnn_conv2d_backward_data", "@d, weights, delta, deltaOffset, inDelta, inDeltaOffset")
meth("procedure batchNormForwardTraining(const batch, channels, blockSize: SizeInt; const outp, x, x_norm, mean, variance, rolling_mean, rolling_variance: TCUMem; const bnMomentum: T; const scales, biases: TCUMem; const activation: longint)",
     "tnn_batchnorm_forward_training", "batch, channels, blockSize, outp, x, x_norm, mean, variance, rolling_mean, rolling_variance, bnMomentum, scales, biases, activation")
meth("procedure batchNormForwardInference(const batch, channels, blockSize: SizeInt; const outp, rolling_mean, rolling_variance, scales, biases: TCUMem; const activation: longint)",
     "tnn_batchnorm_forward_inference", "batch, channels, blockSize, outp, rolling_mean, rolling_variance, scales, biases, activation")
meth("procedure batchNormBackward(const batch, channels, blockSize: SizeInt; const outp: TCUMem; const activation: longint; const delta, x, x_norm, mean, variance, scales, scale_updates, mean_delta, variance_delta: TCUMem)",
     "tnn_batchnorm_backward", "batch, channels, blockSize, outp, activation, delta, x, x_norm, mean, variance, scales, scale_updates, mean_delta, variance_delta")
meth("procedure sgdUpdate(const N: SizeInt; const params, updates: TCUMem; const rate, momentum, decayTerm: T)",
     "tnn_sgd_update", "N, params, updates, rate, momentum, decayTerm")
meth("procedure sgdUpdateMulti(const count: longint; const params, updates: PPointer; const sizes: PQWord; const rates, decayTerms: pointer; const momentum: T)",
     "tnn_sgd_update_multi", "count, params, updates, sizes, rates, decayTerms, momentum")
meth("procedure yoloActivate(const batch, n, classes, wh: SizeInt; const outp: TCUMem; const scaleXY: T)",
     "tnn_yolo_activate", "batch, n, classes, wh, outp, scaleXY")
meth("procedure yoloCandidates(const batch, n, classes, wh: SizeInt; const outp: TCUMem; const thresh: T; const tag: longint; const maxCand: SizeInt; const cand: TCUMem)",
     "tnn_yolo_candidates", "batch, n, classes, wh, outp, thresh, tag, maxCand, cand")
meth("procedure concatChannels(const batch, c, blockSize: SizeInt; const inp: TCUMem; const outChannels, cOffset: SizeInt; const outp: TCUMem)",
     "tnn_concat_channels", "batch, c, blockSize, inp, outChannels, cOffset, outp")
meth("procedure setWeightsStatic(const weights: TCUMem; const isStatic: boolean)", "tnn_set_weights_static", "weights, ord(isStatic)", False)
meth("procedure setMathMode(const mode: longint)", "tnn_set_math_mode", "mode", False)
meth("procedure conv2dForwardSplit(const d: TNNConvDesc; const inp: TCUMem; const inOffset, inPitch: SizeInt; const inLayout: longint; const inPlane: SizeInt; const weights, scale, bias: TCUMem; const activation: longint; const residual: TCUMem; const residualOffset, residualPitch: SizeInt; const outp: TCUMem; const outOffset, outPitch: SizeInt; const outLayout: longint; const outSplit: TCUMem; const outSplitOffset, outSplitPitch, outSplitPlane: SizeInt)",
     "tnn_conv2d_forward_split", "@d, inp, inOffset, inPitch, inLayout, inPlane, weights, scale, bias, activation, residual, residualOffset, residualPitch, outp, outOffset, outPitch, outLayout, outSplit, outSplitOffset, outSplitPitch, outSplitPlane")
meth("procedure conv2dForwardSplitResidual(const d: TNNConvDesc; const inp: TCUMem; const inOffset, inPitch: SizeInt; const inLayout: longint; const inPlane: SizeInt; const weights, scale, bias: TCUMem; const activation: longint; const residual: TCUMem; const residualOffset, residualPitch: SizeInt; const residualLayout: longint; const residualPlane: SizeInt; const outp: TCUMem; const outOffset, outPitch: SizeInt; const outLayout: longint; const outSplit: TCUMem; const outSplitOffset, outSplitPitch, outSplitPlane: SizeInt)",
     "tnn_conv2d_forward_split_residual", "@d, inp, inOffset, inPitch, inLayout, inPlane, weights, scale, bias, activation, residual, residualOffset, residualPitch, residualLayout, residualPlane, outp, outOffset, outPitch, outLayout, outSplit, outSplitOffset, outSplitPitch, outSplitPlane")
meth("procedure conv2dForwardYoloHead(const d: TNNConvDesc; const inp: TCUMem; const inOffset, inPitch: SizeInt; const inLayout: longint; const inPlane: SizeInt; const weights, scale, bias, outp: TCUMem; const outOffset, classes: SizeInt; const scaleXY: T)",
     "tnn_conv2d_forward_yolo_head", "@d, inp, inOffset, inPitch, inLayout, inPlane, weights, scale, bias, outp, outOffset, classes, scaleXY")
meth("procedure nhwcToSplit(const rows, channels: SizeInt; const inp: TCUMem; const inPitch: SizeInt; const outSplit: TCUMem; const outPitch, outPlane: SizeInt)",
     "tnn_nhwc_to_split", "rows, channels, inp, inPitch, outSplit, outPitch, outPlane")
meth("procedure splitToNhwc(const rows, channels: SizeInt; const inSplit: TCUMem; const inPitch, inPlane: SizeInt; const outp: TCUMem; const outPitch: SizeInt)",
     "tnn_split_to_nhwc", "rows, channels, inSplit, inPitch, inPlane, outp, outPitch")
meth("procedure nchwToNhwc(const batch, channels, blockSize: SizeInt; const inp, outp: TCUMem; const outPitch: SizeInt)",
     "tnn_nchw_to_nhwc", "batch, channels, blockSize, inp, outp, outPitch")
meth("procedure nhwcToNchw(const batch, channels, blockSize: SizeInt; const inp: TCUMem; const inPitch: SizeInt; const outp: TCUMem)",
     "tnn_nhwc_to_nchw", "batch, channels, blockSize, inp, inPitch, outp")
meth("procedure rowsCopy(const rows, channels: SizeInt; const inp: TCUMem; const inPitch: SizeInt; const outp: TCUMem; const outPitch: SizeInt)",
     "tnn_rows_copy", "rows, channels, inp, inPitch, outp, outPitch")
meth("procedure rowsAdd(const rows, channels: SizeInt; const a: TCUMem; const aPitch: SizeInt; const b: TCUMem; const bPitch: SizeInt; const outp: TCUMem; const outPitch: SizeInt)",
     "tnn_rows_add", "rows, channels, a, aPitch, b, bPitch, outp, outPitch")
meth("procedure upSampleNhwc(const batch, channels, inHeight, inWidth, stride: SizeInt; const scale: T; const inp: TCUMem; const inPitch: SizeInt; const outp: TCUMem; const outPitch: SizeInt)",
     "tnn_upsample_nhwc", "batch, channels, inHeight, inWidth, stride, scale, inp, inPitch, outp, outPitch")
meth("procedure lstmGates(const batch, hidden: SizeInt; const wz, bw, uz: TCUMem; const uzOffset: SizeInt; const c, h, cellOut, hOut: TCUMem; const outOffset: SizeInt)",
     "tnn_lstm_gates", "batch, hidden, wz, bw, uz, uzOffset, c, h, cellOut, hOut, outOffset")


def method_name(decl):
    return re.match(r"(?:procedure|function)\s+(\w+)", decl).group(1)


def n_call_args(args):
    depth, n = 0, 1 if args.strip() else 0
    for ch in args:
        depth += ch == "("
        depth -= ch == ")"
        n += ch == "," and depth == 0
    return n


def generate():
    decls = parse_header()
    syms = {name: (ret, params) for ret, name, params in decls}
    L = []
    w = L.append
    w("// GENERATED by tools/gen_pascal.py from include/tensorium_b200.h — do not edit; regenerate instead.")
    w("//")
    w("// unit nnCuda: drop-in replacement for the reference's source/nncuda.pas (class TNNCuda<T>, lines 35-211; the")
    w("// global `cuda : TNNCuda<Single>`, source/ntensors.pas:1053-1056).  Same unit name, class name and public method")
    w("// signatures, so ntensors.pas, nnet.pas and every layer compile unchanged under -dUSE_CUDART.  Every method is one")
    w("// call into libtensorium_b200.so (B200 / sm_100a; there is no NVRTC step and no CPU fallback).")
    w("// Link: -k-ltensorium_b200 (or place the .so next to the executable).")
    w("unit nnCuda;")
    w("{$ifdef FPC}{$mode delphi}{$H+}{$endif}")
    w("interface")
    w("uses SysUtils;")
    w("")
    w("const TNNLIB = 'tensorium_b200';")
    w("  // tnn_math_mode")
    w("  TNN_MATH_FP32 = 0; TNN_MATH_3XTF32 = 1; TNN_MATH_TF32 = 2; TNN_MATH_BF16X3 = 3;")
    w("  // tnn_layout")
    w("  TNN_LAYOUT_NCHW = 0; TNN_LAYOUT_NHWC = 1;")
    w("type")
    w("  TNNContext = pointer;")
    w("  PQWord = ^QWord;")
    w("  TNNConvDesc = packed record   // tnn_conv_desc: 13 x int64")
    w("    batch, c, h, w, filters, kh, kw, pad_h, pad_w, stride_y, stride_x, dilation_y, dilation_x: int64;")
    w("  end;")
    w("  PNNConvDesc = ^TNNConvDesc;")
    w("")
    w("// ---- C ABI: one declaration per symbol of include/tensorium_b200.h (cdecl = SysV x86-64 on Linux)")
    for ret, name, params in decls:
        w(external_decl(ret, name, params))
    w("")
    w("type")
    w("  { TNNCuda }")
    w("  TNNCuda<T> = class")
    w("  const")
    w("    NUM_THREADS = $200;   // kept for source compatibility (nncuda.pas:37); launch shapes live in the library")
    w("  type")
    w("    PCUMem = ^TCUMem;")
    w("    TCUMem = ^T;")
    w("    TNNDeviceProp = record name: ansistring; multiProcessorCount, major, minor: longint; end;")
    w("  private")
    w("    FCtx: TNNContext;")
    w("    procedure SAFE_CALL(const rc: longint); inline;   // nncuda.pas:237-274")
    w("  public")
    w("    stream: pointer;            // the CUstream all operators are enqueued on (nncuda.pas:78)")
    w("    properties: TNNDeviceProp;  // properties.name is printed by the samples (CIFAR10Conv.pas:138)")
    w("    useBLAS: integer;           // kept for source compatibility (CIFAR10Conv.pas:139); see setMathMode")
    w("    class function deviceCount(): longint;")
    w("    constructor Create(deviceIndex: longint = 0);")
    w("    destructor Destroy(); override;")
    w("    function CompileLog: ansistring;")
    w("    function createDeviceBuffer(const N: SizeInt): TCUMem;")
    w("    procedure freeDeviceBuffer(cudaMem: TCUMem);")
    w("    procedure readBuffer(const cudaMem: TCUMem; const bufferSize: size_t; const buffer: pointer);")
    w("    procedure writeBuffer(const cudaMem: TCUMem; const bufferSize: size_t; const buffer: pointer);")
    for decl, _, _, _ in M[:N_REFERENCE_METHODS]:
        w(f"    {decl};")
    w("    procedure finish();")
    w("    // NVRTC entry points of the reference: no-ops here (ntensors.pas:6207 calls compileToCUBIN + loadCUBIN at start-up)")
    w("    function compileToCUBIN(const code, name: ansistring; const headers: TArray<PAnsiChar> = nil; const includeNames: TArray<PAnsiChar> = nil): RawByteString;")
    w("    procedure loadCUBIN(const cubin: RawByteString);")
    w("    function compileFile(const filename: ansistring): RawByteString;")
    w("    procedure loadCUBinFile(const filename: ansistring);")
    w("    // ---- new with this backend: the fused / coarse entry points the edited layer bodies call")
    for decl, _, _, _ in M[N_REFERENCE_METHODS:]:
        w(f"    {decl};")
    w("    property context: TNNContext read FCtx;   // for the entry points not wrapped here (graphs, events, NCCL exchange)")
    w("  end;")
    w("")
    w("implementation")
    w("")
    w("procedure TNNCuda<T>.SAFE_CALL(const rc: longint);")
    w("begin")
    w("  assert(rc = 0, string(tnn_last_error()));")
    w("end;")
    w("")
    w("class function TNNCuda<T>.deviceCount(): longint;")
    w("begin")
    w("  result := 0;")
    w("  tnn_device_count(@result);")
    w("end;")
    w("")
    w("constructor TNNCuda<T>.Create(deviceIndex: longint);")
    w("begin")
    w("  SAFE_CALL(tnn_create(deviceIndex, @FCtx));")
    w("  properties.name := tnn_device_name(FCtx);")
    w("  SAFE_CALL(tnn_device_info(FCtx, @properties.multiProcessorCount, @properties.major, @properties.minor));")
    w("  stream := tnn_get_stream(FCtx);")
    w("  useBLAS := 1;")
    w("end;")
    w("")
    w("destructor TNNCuda<T>.Destroy();")
    w("begin")
    w("  tnn_destroy(FCtx);")
    w("  inherited;")
    w("end;")
    w("")
    w("function TNNCuda<T>.CompileLog: ansistring;")
    w("begin")
    w("  result := '';")
    w("end;")
    w("")
    w("function TNNCuda<T>.createDeviceBuffer(const N: SizeInt): TCUMem;")
    w("begin")
    w("  // plain cudaMalloc memory: TTensor.Free calls cudaFree on it directly (ntensors.pas:6126-6129)")
    w("  SAFE_CALL(tnn_create_device_buffer(FCtx, N, @result));")
    w("end;")
    w("")
    w("procedure TNNCuda<T>.freeDeviceBuffer(cudaMem: TCUMem);")
    w("begin")
    w("  SAFE_CALL(tnn_free_device_buffer(FCtx, cudaMem));")
    w("end;")
    w("")
    w("procedure TNNCuda<T>.readBuffer(const cudaMem: TCUMem; const bufferSize: size_t; const buffer: pointer);")
    w("begin")
    w("  SAFE_CALL(tnn_read_buffer(FCtx, cudaMem, bufferSize, buffer));")
    w("end;")
    w("")
    w("procedure TNNCuda<T>.writeBuffer(const cudaMem: TCUMem; const bufferSize: size_t; const buffer: pointer);")
    w("begin")
    w("  SAFE_CALL(tnn_write_buffer(FCtx, cudaMem, bufferSize, buffer));")
    w("end;")
    w("")
    w("procedure TNNCuda<T>.finish();")
    w("begin")
    w("  SAFE_CALL(tnn_finish(FCtx));")
    w("end;")
    w("")
    w("function TNNCuda<T>.compileToCUBIN(const code, name: ansistring; const headers: TArray<PAnsiChar>; const includeNames: TArray<PAnsiChar>): RawByteString;")
    w("begin")
    w("  result := '';")
    w("end;")
    w("")
    w("procedure TNNCuda<T>.loadCUBIN(const cubin: RawByteString);")
    w("begin")
    w("end;")
    w("")
    w("function TNNCuda<T>.compileFile(const filename: ansistring): RawByteString;")
    w("begin")
    w("  result := '';")
    w("end;")
    w("")
    w("procedure TNNCuda<T>.loadCUBinFile(const filename: ansistring);")
    w("begin")
    w("end;")
    w("")
    for decl, sym, args, tele in M:
        assert sym in syms, f"{sym} is not declared in the header"
        want = len(syms[sym][1]) - 1
        assert n_call_args(args) == want, f"{method_name(decl)} passes {n_call_args(args)} arguments, {sym} takes {want}"
        # strip default values for the implementation header
        impl = re.sub(r"\s*=\s*[\w$]+", "", decl)
        kind, rest = impl.split(" ", 1)
        w(f"{kind} TNNCuda<T>.{rest};")
        w("begin")
        w(f"  SAFE_CALL({sym}(FCtx, {args}));")
        if tele:
            w("  {$ifdef USE_TELEMETRY} finish(); {$endif}")
        w("end;")
        w("")
    w("end.")
    return "\n".join(L) + "\n"


def main():
    text = generate()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != text:
            print("pascal/nnCuda.pas is stale: run python tools/gen_pascal.py")
            sys.exit(1)
        print("pascal/nnCuda.pas is up to date")
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(text)
    print(f"wrote {OUT}: {len(parse_header())} external declarations, {len(M)} operator methods "
          f"({N_REFERENCE_METHODS} with the reference's signatures)")


if __name__ == "__main__":
    main()
