"""Oracle restatement of the reference's YOLO post-processing — TEST INFRASTRUCTURE ONLY (never imported by the product).

Follows, statement by statement, the host-side code the reference runs after the yolo layers' outputs are pulled:

  TYoloLayer.entryIndex      source/nyololayer.pas:97-103
  TYoloLayer.getBox          source/nyololayer.pas:105-122      (newCoords = false branch)
  TYoloLayer.getDetections   source/nyololayer.pas:399-441
  TYoloLayer.correctBoxes    source/nyololayer.pas:443-498      (both the letter-box and the plain branch)
  TNNet.Detections           source/nnet.pas:584-677            (yolo layers concatenated in layer order)
  overlap / box_intersection / box_union / box_iou   source/ntypes.pas:1059-1108
  TDetectionsHelper.Comparer source/ntypes.pas:1554-1566
  TTools<T>.QuickSort        source/ntensors.pas:2483-2528      (the UNSTABLE middle-pivot Hoare sort the NMS uses —
                                                                 restated exactly, so ties resolve as in the reference)
  TDetectionsHelper.doNMSSort source/ntypes.pas:1612-1646

Arithmetic: `single` operands are numpy float32; `exp` returns ValReal (double) in FreePascal, so the width / height
products are evaluated in double and rounded once on assignment to the single field, like the Pascal expression.
Parity unpinned by reference outputs (the reference ships no detections fixture and cannot be built here); pinned
by hand-computed cases in tests/test_detect_oracle.py.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


class Detection:
    """TDetection (source/ntypes.pas:278-291), the fields this path touches."""
    __slots__ = ("id", "bbox", "classes", "prob", "objectness", "sort_class")

    def __init__(self, classes):
        self.id = 0
        self.bbox = [f32(0), f32(0), f32(0), f32(0)]  # x, y, w, h
        self.classes = classes
        self.prob = np.zeros(classes, np.float32)
        self.objectness = f32(0)
        self.sort_class = -1


def entry_index(w, h, classes, outputs, abatch, location, entry):
    """nyololayer.pas:97-103"""
    _n = location // (w * h)
    loc = location % (w * h)
    return abatch * outputs + _n * (classes + 4 + 1) * w * h + entry * w * h + loc


def get_box(out, biases, w, h, n, index, col, row, net_w, net_h, stride):
    """nyololayer.pas:105-122 (newCoords = false)"""
    n2 = 2 * n
    x = f32(f32(col) + out[index + 0 * stride]) / f32(w)
    y = f32(f32(row) + out[index + 1 * stride]) / f32(h)
    bw = f32(np.exp(np.float64(out[index + 2 * stride])) * np.float64(biases[n2]) / net_w)
    bh = f32(np.exp(np.float64(out[index + 3 * stride])) * np.float64(biases[n2 + 1]) / net_h)
    return [f32(x), f32(y), bw, bh]


def correct_boxes(dets, w, h, net_w, net_h, relative, letter):
    """nyololayer.pas:443-498"""
    if letter:
        if (net_w / w) < (net_h / h):
            new_w = net_w
            new_h = (h * net_w) // w
        else:
            new_h = net_h
            new_w = (w * net_h) // h
    else:
        new_w, new_h = net_w, net_h
    deltaw = f32(net_w - new_w)
    deltah = f32(net_h - new_h)
    ratiow = f32(new_w / net_w)
    ratioh = f32(new_h / net_h)
    for d in dets:
        x, y, bw, bh = d.bbox
        x = f32((np.float64(x) - np.float64(deltaw) / 2.0 / net_w) / np.float64(ratiow))
        y = f32((np.float64(y) - np.float64(deltah) / 2.0 / net_h) / np.float64(ratioh))
        bw = f32(np.float64(bw) * (1 / np.float64(ratiow)))
        bh = f32(np.float64(bh) * (1 / np.float64(ratioh)))
        if not relative:
            x = f32(x * f32(w)); bw = f32(bw * f32(w))
            y = f32(y * f32(h)); bh = f32(bh * f32(h))
        d.bbox = [x, y, bw, bh]


def get_detections(output, w, h, n, classes, mask, biases, img_w, img_h, net_w, net_h, thresh, relative=True,
                   letter=False, abatch=0):
    """nyololayer.pas:399-441.  `output`: the yolo layer's host tensor [batch, n*(classes+5), h, w] (any shape, flat
    order is what matters)."""
    out = np.ascontiguousarray(output, dtype=np.float32).reshape(-1)
    outputs = h * w * n * (classes + 4 + 1)
    dets = []
    thresh = f32(thresh)
    for i in range(w * h):
        row = i // w
        col = i % w
        for _n in range(n):
            obj_index = entry_index(w, h, classes, outputs, abatch, _n * w * h + i, 4)
            objectness = out[obj_index]
            if objectness > thresh:
                box_index = entry_index(w, h, classes, outputs, abatch, _n * w * h + i, 0)
                d = Detection(classes)
                d.id = len(dets)
                d.bbox = get_box(out, biases, w, h, mask[_n], box_index, col, row, net_w, net_h, w * h)
                d.objectness = objectness
                for j in range(classes):
                    class_index = entry_index(w, h, classes, outputs, abatch, _n * w * h + i, 4 + 1 + j)
                    prob = f32(objectness * out[class_index])
                    d.prob[j] = prob if prob > thresh else f32(0)
                dets.append(d)
    correct_boxes(dets, img_w, img_h, net_w, net_h, relative, letter)
    return dets


def net_detections(yolo_layers, img_w, img_h, net_w, net_h, thresh, relative=True, letter=False, abatch=0):
    """TNNet.Detections nnet.pas:584-677: the yolo layers' detections back to back, in layer order.
    yolo_layers: iterable of dicts {output, w, h, n, classes, mask, biases}."""
    dets = []
    for l in yolo_layers:
        dets += get_detections(l["output"], l["w"], l["h"], l["n"], l["classes"], l["mask"], l["biases"], img_w, img_h,
                               net_w, net_h, thresh, relative, letter, abatch)
    return dets


# ------------------------------------------------------------------ boxes (ntypes.pas:1059-1108)
def overlap(x1, w1, x2, w2):
    l1 = f32(x1 - w1 / f32(2))
    l2 = f32(x2 - w2 / f32(2))
    left = l1 if l1 > l2 else l2
    r1 = f32(x1 + w1 / f32(2))
    r2 = f32(x2 + w2 / f32(2))
    right = r1 if r1 < r2 else r2
    return f32(right - left)


def box_intersection(a, b):
    w = overlap(a[0], a[2], b[0], b[2])
    h = overlap(a[1], a[3], b[1], b[3])
    if w < 0 or h < 0:
        return f32(0)
    return f32(w * h)


def box_union(a, b):
    i = box_intersection(a, b)
    return f32(f32(a[2] * a[3]) + f32(b[2] * b[3]) - i)


def box_iou(a, b):
    i = box_intersection(a, b)
    u = box_union(a, b)
    if i == 0 or u == 0:
        return f32(0)
    return f32(i / u)


# ------------------------------------------------------------------ sort + NMS
def comparer(a: Detection, b: Detection) -> int:
    """ntypes.pas:1554-1566"""
    if b.sort_class >= 0:
        diff = f32(a.prob[b.sort_class] - b.prob[b.sort_class])
    else:
        diff = f32(a.objectness - b.objectness)
    if diff < 0:
        return 1
    if diff > 0:
        return -1
    return 0


def quick_sort(arr, L, R, compare):
    """TTools<T>.QuickSort ntensors.pas:2483-2528 (Descending = false -> neg = 1).  The pivot is a COPY of the middle
    record, as in the reference (P := Arr[(L + R) shr 1]).  Where the reference would evaluate Compare on Arr[R+1] /
    Arr[-1] before its bounds test (short-circuit order `Compare(...) and (I <= R)`), the bounds test is done first."""
    while True:
        I, J = L, R
        P = arr[(L + R) >> 1]
        while True:
            while I <= R and compare(P, arr[I]) > 0:
                I += 1
            while J >= 0 and compare(P, arr[J]) < 0:
                J -= 1
            if I <= J:
                arr[I], arr[J] = arr[J], arr[I]
                I += 1
                J -= 1
            if I > J:
                break
        if J - L < R - I:
            if L < J:
                quick_sort(arr, L, J, compare)
            L = I
        else:
            if I < R:
                quick_sort(arr, I, R, compare)
            R = J
        if L >= R:
            break


def do_nms_sort(dets, classes, thresh):
    """TDetectionsHelper.doNMSSort ntypes.pas:1612-1646 — in place; returns the list (permuted like the reference)."""
    thresh = f32(thresh)
    total = len(dets) - 1
    k = total
    i = 0
    while i <= k:
        if dets[i].objectness == 0:
            dets[i], dets[k] = dets[k], dets[i]
            k -= 1
            i -= 1
        i += 1
    total = k + 1
    for c in range(classes):
        for i in range(total):
            dets[i].sort_class = c
        if total > 1:
            quick_sort(dets, 0, total - 1, comparer)
        for i in range(total):
            if dets[i].prob[c] == 0:
                continue
            a = dets[i].bbox
            for j in range(i + 1, total):
                b = dets[j].bbox
                if box_iou(a, b) > thresh:
                    dets[j].prob[c] = f32(0)
    return dets


def surviving(dets):
    """(class, probability, box) of every class score that survived — the order-free summary tests compare."""
    out = []
    for d in dets:
        for c in np.nonzero(d.prob)[0]:
            out.append((int(c), float(d.prob[c]), [float(v) for v in d.bbox]))
    out.sort(key=lambda t: (t[0], -t[1]))
    return out
