"""Generates the frozen golden fixtures of tests/golden/.

The reference (achinastone/Tensorium) ships no tests or golden vectors for this path and cannot be built here
(FreePascal absent), so the fixtures are produced by the CPU oracle (oracle/tensorium_oracle.c, a restatement of
the cited Pascal routines) with fixed seeds and frozen in-repo; `hand_im2col.json` is derived independently of
the oracle, straight from the definition in source/ntensors.pas:11436-11465 (pure Python below).

Run:  python tests/golden/make_golden.py     (rewrites oracle_fixtures.npz and hand_im2col.json)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as orc  # noqa: E402
from oracle.ref_layers import RefConvLayer  # noqa: E402


def hand_im2col(im, kh, kw, ph, pw, sy, sx, dy, dx):
    """col[c][kr][kc][oh][ow] = im[c][oh*sy-ph+kr*dy][ow*sx-pw+kc*dx] or 0 — from the definition, no oracle."""
    c, h, w = len(im), len(im[0]), len(im[0][0])
    oh = (h + 2 * ph - (dy * (kh - 1) + 1)) // sy + 1
    ow = (w + 2 * pw - (dx * (kw - 1) + 1)) // sx + 1
    rows = []
    for ci in range(c):
        for kr in range(kh):
            for kc in range(kw):
                row = []
                for y in range(oh):
                    for x in range(ow):
                        iy, ix = y * sy - ph + kr * dy, x * sx - pw + kc * dx
                        row.append(im[ci][iy][ix] if 0 <= iy < h and 0 <= ix < w else 0)
                rows.append(row)
    return rows


def main():
    # ---- hand-checkable im2col cases (integers, exact)
    cases = []
    im = [[[float(4 * r + c + 1) for c in range(4)] for r in range(4)]]  # 1x4x4, values 1..16
    for (kh, kw, p, s, d) in [(3, 3, 1, 1, 1), (3, 3, 1, 2, 1), (1, 1, 0, 1, 1), (3, 3, 0, 1, 1), (2, 2, 1, 2, 1),
                              (3, 3, 2, 1, 2)]:
        cases.append(dict(im=im, kh=kh, kw=kw, ph=p, pw=p, sy=s, sx=s, dy=d, dx=d,
                          col=hand_im2col(im, kh, kw, p, p, s, s, d, d)))
    im2 = [[[float(100 * ch + 5 * r + c) for c in range(5)] for r in range(3)] for ch in range(2)]  # 2x3x5 ragged
    cases.append(dict(im=im2, kh=3, kw=2, ph=1, pw=0, sy=1, sx=2, dy=1, dx=1,
                      col=hand_im2col(im2, 3, 2, 1, 0, 1, 2, 1, 1)))
    with open(os.path.join(HERE, "hand_im2col.json"), "w") as f:
        json.dump(cases, f)

    # ---- oracle-generated fixtures (both accumulation modes)
    out = {}
    rng = np.random.default_rng(1234)
    U = lambda *s: rng.uniform(-1, 1, s).astype(np.float32)
    for mode, tag in ((orc.MODE_PAS, "pas"), (orc.MODE_FMA, "fma")):
        orc.set_mode(mode)
        r2 = np.random.default_rng(99)
        A, B, C0 = (r2.uniform(-1, 1, s).astype(np.float32) for s in ((37, 53), (53, 29), (37, 29)))
        for ta, tb in ((0, 0), (0, 1), (1, 0), (1, 1)):
            a = np.ascontiguousarray(A.T) if ta else A
            b = np.ascontiguousarray(B.T) if tb else B
            c = C0.copy()
            orc.sgemm(ta, tb, 37, 29, 53, 0.75, a, a.shape[1], b, b.shape[1], 0.5, c, 29)
            out[f"{tag}_sgemm_{ta}{tb}"] = c
        out[f"{tag}_sgemm_A"], out[f"{tag}_sgemm_B"], out[f"{tag}_sgemm_C0"] = A, B, C0
        # conv layer forward/backward/update, with and without batch-norm
        for bn in (0, 1):
            r3 = np.random.default_rng(7 + bn)
            x = r3.uniform(0, 1, (3, 5, 9, 8)).astype(np.float32)
            w = r3.uniform(-0.3, 0.3, (6, 5 * 9)).astype(np.float32)
            bias = r3.uniform(-0.1, 0.1, 6).astype(np.float32)
            scales = r3.uniform(0.5, 1.5, 6).astype(np.float32)
            rm = r3.uniform(-0.1, 0.1, 6).astype(np.float32)
            rv = r3.uniform(0.5, 1.5, 6).astype(np.float32)
            L = RefConvLayer(3, 5, 9, 8, 6, 3, 2, 1, orc.ACT["LEAKY"], bn, w, bias, scales, rm, rv)
            y_inf = L.forward(x, training=False).copy()
            y_tr = L.forward(x, training=True).copy()
            g = r3.uniform(-1, 1, y_tr.shape).astype(np.float32)
            L.delta[...] = g
            up = np.zeros_like(x)
            L.backward(up)
            pre = f"{tag}_conv_bn{bn}_"
            out.update({pre + "x": x, pre + "w": w, pre + "bias": bias, pre + "scales": scales, pre + "rm": rm,
                        pre + "rv": rv, pre + "y_inf": y_inf, pre + "y_tr": y_tr, pre + "g": g, pre + "dx": up,
                        pre + "dw": L.weight_updates.copy(), pre + "db": L.bias_updates.copy()})
            if bn:
                out.update({pre + "dscale": L.scale_updates.copy(), pre + "mean": L.mean.copy(),
                            pre + "var": L.variance.copy(), pre + "rm_after": L.rolling_mean.copy(),
                            pre + "rv_after": L.rolling_variance.copy()})
            L.update(0.001, 0.9, 0.0001, 3)
            out.update({pre + "w_after": L.weights.copy(), pre + "b_after": L.biases.copy(),
                        pre + "wu_after": L.weight_updates.copy()})
    orc.set_mode(orc.MODE_PAS)
    # activations on a fixed grid (covers every branch point)
    xs = np.concatenate([np.linspace(-8, 8, 161), [-4, 4, 0, 1, -1, 6, 6.5, 0.5, -0.5]]).astype(np.float32)
    out["act_x"] = xs
    for name, a in orc.ACT.items():
        if name in ("SOFTMAX",):
            continue
        y = xs.copy()
        orc.activate_array(y, a)
        out[f"act_{name}"] = y
        if name != "SWISH":
            d = np.ones_like(xs)
            orc.gradient_array(y, a, d)
            out[f"grad_{name}"] = d
    np.savez_compressed(os.path.join(HERE, "oracle_fixtures.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
