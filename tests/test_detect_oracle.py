"""CPU tests (no GPU): the oracle restatement of the reference's yolo post-processing (oracle/ref_detect.py —
getDetections / correctBoxes / box_iou / QuickSort / doNMSSort, nyololayer.pas:399-498, ntypes.pas:1059-1108,
1554-1566,1612-1646, ntensors.pas:2483-2528) against hand-computed cases, and the product's host-side decode + NMS
(tensorium_b200/nnet.py) against that oracle on random inputs."""
import numpy as np
import pytest

from oracle import ref_detect as rd
from tensorium_b200 import nnet


def D(prob, box=(0.5, 0.5, 0.2, 0.2), obj=0.9):
    d = rd.Detection(len(prob))
    d.prob[:] = prob
    d.bbox = [np.float32(v) for v in box]
    d.objectness = np.float32(obj)
    return d


def test_box_iou_hand_cases():
    a = [0.5, 0.5, 0.2, 0.2]
    assert rd.box_iou(a, a) == pytest.approx(1.0, abs=1e-6)
    # shifted by half a width: intersection 0.1*0.2 = 0.02, union 0.04 + 0.04 - 0.02 = 0.06
    assert rd.box_iou(a, [0.6, 0.5, 0.2, 0.2]) == pytest.approx(1 / 3, rel=1e-5)
    assert rd.box_iou(a, [0.9, 0.9, 0.2, 0.2]) == 0.0          # disjoint (negative overlap -> 0)
    q = [np.float32(v) for v in (0.5, 0.5, 0.25, 0.25)]      # exactly representable: touching edges give exactly 0
    assert rd.box_iou(q, [np.float32(v) for v in (0.75, 0.5, 0.25, 0.25)]) == 0.0
    # contained box: intersection = small area, union = big area
    assert rd.box_iou([0.5, 0.5, 0.4, 0.4], [0.5, 0.5, 0.2, 0.2]) == pytest.approx(0.25, rel=1e-5)


def test_quicksort_traced_by_hand():
    """[0.2, 0.8, 0.5] through the reference's middle-pivot Hoare sort (traced on paper from ntensors.pas:2483-2528):
    pivot 0.8 swaps positions 0/1 -> [0.8, 0.2, 0.5]; the right part [0.2, 0.5] with pivot 0.2 swaps -> [0.8, 0.5, 0.2]."""
    arr = [D([0.2]), D([0.8]), D([0.5])]
    for d in arr:
        d.sort_class = 0
    rd.quick_sort(arr, 0, 2, rd.comparer)
    assert [float(d.prob[0]) for d in arr] == pytest.approx([0.8, 0.5, 0.2])


def test_quicksort_sorts_descending_and_permutes():
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 7, 64, 257):
        vals = rng.integers(0, 20, n).astype(np.float32) / 20  # many ties: the unstable order must still be sorted
        arr = [D([v]) for v in vals]
        for i, d in enumerate(arr):
            d.id, d.sort_class = i, 0
        if n > 1:
            rd.quick_sort(arr, 0, n - 1, rd.comparer)
        got = [float(d.prob[0]) for d in arr]
        assert got == sorted(vals.tolist(), reverse=True)
        assert sorted(d.id for d in arr) == list(range(n))


def test_objectness_comparer_when_sort_class_negative():
    a, b = D([0.1], obj=0.3), D([0.9], obj=0.7)
    a.sort_class = b.sort_class = -1
    assert rd.comparer(a, b) == 1 and rd.comparer(b, a) == -1 and rd.comparer(a, a) == 0


def test_nms_hand_case():
    """Class 0: A(.9) suppresses the overlapping B(.8) (IoU 1/3 > 0.3); C is far away and survives.
    Class 1: only B scores, nothing to suppress.  A zero-objectness entry is moved out of the working set."""
    A = D([0.9, 0.0], (0.5, 0.5, 0.2, 0.2))
    B = D([0.8, 0.6], (0.6, 0.5, 0.2, 0.2))
    C = D([0.7, 0.0], (0.1, 0.1, 0.1, 0.1))
    Z = D([0.99, 0.99], (0.5, 0.5, 0.2, 0.2), obj=0.0)
    dets = rd.do_nms_sort([Z, B, A, C], 2, 0.3)
    sv = rd.surviving([d for d in dets if d.objectness != 0])
    assert [(c, round(p, 4)) for c, p, _ in sv] == [(0, 0.9), (0, 0.7), (1, 0.6)]
    assert dets[-1] is Z  # swapped to the tail, untouched
    # a higher threshold keeps B
    A, B, C = D([0.9, 0.0], (0.5, 0.5, 0.2, 0.2)), D([0.8, 0.6], (0.6, 0.5, 0.2, 0.2)), D([0.7, 0.0], (0.1, 0.1, 0.1, 0.1))
    sv = rd.surviving(rd.do_nms_sort([B, A, C], 2, 0.4))
    assert [(c, round(p, 4)) for c, p, _ in sv] == [(0, 0.9), (0, 0.8), (0, 0.7), (1, 0.6)]


def test_get_detections_hand_case():
    """A 2x1 grid, one anchor, two classes.  Cell 1 passes: x = (1 + .25)/2, y = (0 + .5)/1, w = e^0 * 116 / 416,
    h = e^ln2 * 90 / 416; class scores .9*.8 = .72 (kept) and .9*.3 = .27 (< thresh -> 0)."""
    w, h, n, classes = 2, 1, 1, 2
    out = np.zeros((1, n * (classes + 5), h, w), np.float32)
    out[0, :, 0, 0] = [0.5, 0.5, 0.0, 0.0, 0.4, 0.9, 0.9]                 # objectness .4 < thresh
    out[0, :, 0, 1] = [0.25, 0.5, 0.0, np.log(2.0), 0.9, 0.8, 0.3]
    dets = rd.get_detections(out, w, h, n, classes, [6], np.asarray(nnet.YOLOV3_ANCHORS, np.float32), 416, 416, 416, 416,
                             0.5, True, False, 0)
    assert len(dets) == 1
    assert dets[0].bbox == pytest.approx([0.625, 0.5, 116 / 416, 2 * 90 / 416], rel=1e-6)
    assert dets[0].prob.tolist() == pytest.approx([0.72, 0.0], rel=1e-6)
    assert float(dets[0].objectness) == pytest.approx(0.9)


def test_correct_boxes_letterbox_hand_case():
    """A 200x100 image letter-boxed into 416x416: new_w = 416, new_h = 208, so y is stretched around the centre:
    y' = (y - 104/416) / 0.5, h' = 2h; relative=False then scales to pixels."""
    d = D([1.0], (0.5, 0.5, 0.25, 0.125))
    rd.correct_boxes([d], 200, 100, 416, 416, True, True)
    assert d.bbox == pytest.approx([0.5, 0.5, 0.25, 0.25], rel=1e-6)
    d = D([1.0], (0.5, 0.375, 0.25, 0.125))
    rd.correct_boxes([d], 200, 100, 416, 416, False, True)
    assert d.bbox == pytest.approx([100.0, 25.0, 50.0, 25.0], rel=1e-6)
    d = D([1.0], (0.3, 0.4, 0.1, 0.2))
    rd.correct_boxes([d], 640, 480, 416, 416, True, False)   # no letter box, relative: identity
    assert d.bbox == pytest.approx([0.3, 0.4, 0.1, 0.2], rel=1e-7)


def product_surviving(boxes, probs):
    out = []
    for i in range(len(boxes)):
        for c in np.nonzero(probs[i])[0]:
            out.append((int(c), float(probs[i, c]), [float(v) for v in boxes[i]]))
    out.sort(key=lambda t: (t[0], -t[1]))
    return out


@pytest.mark.parametrize("letter,relative", [(False, True), (True, True), (True, False)])
def test_product_decode_and_nms_match_the_oracle(letter, relative):
    """The product's getDetections / correctBoxes / doNMSSort (stable argsort) vs the oracle's restatement with the
    reference's unstable QuickSort: same surviving (class, score) list, boxes IoU >= 0.99."""
    rng = np.random.default_rng(3)
    anchors = np.asarray(nnet.YOLOV3_ANCHORS, np.float32)
    layers, outs = [], []
    for (h, w, mask) in ((4, 5, [6, 7, 8]), (8, 10, [3, 4, 5])):
        n, classes = 3, 6
        out = rng.uniform(0, 1, (2, n * (classes + 5), h, w)).astype(np.float32)
        out[:, [2, 3, 13, 14, 24, 25]] = rng.uniform(-1, 1, (2, 6, h, w))  # raw w/h logits
        yl = nnet.TYoloLayer.__new__(nnet.TYoloLayer)
        yl.w, yl.h, yl.n, yl.classes, yl.mask, yl.biases = w, h, n, classes, mask, anchors
        layers.append(yl)
        outs.append(out)
    for abatch in (0, 1):
        bs, os_, ps = [], [], []
        for yl, out in zip(layers, outs):
            b, o, p = yl.getDetections(out, 160, 128, 0.55, abatch, 300, 200, relative, letter)
            bs.append(b), os_.append(o), ps.append(p)
        B, O, P = nnet.doNMSSort(np.concatenate(bs), np.concatenate(os_), np.concatenate(ps), 0.45)
        dets = rd.net_detections([dict(output=out, w=yl.w, h=yl.h, n=yl.n, classes=yl.classes, mask=yl.mask,
                                       biases=anchors) for yl, out in zip(layers, outs)], 300, 200, 160, 128, 0.55,
                                 relative, letter, abatch)
        assert len(dets) == len(B) > 20
        rd.do_nms_sort(dets, 6, 0.45)
        want, got = rd.surviving(dets), product_surviving(B, P)
        assert len(want) == len(got) > 0
        for (c1, p1, b1), (c2, p2, b2) in zip(want, got):
            assert c1 == c2 and p1 == pytest.approx(p2, rel=1e-6)
            assert rd.box_iou(b1, b2) >= 0.99


def test_candidate_list_decodes_like_the_full_outputs():
    """TNNet.DetectionsFromCandidates (host side of tnn_yolo_candidates): a candidate list assembled here the way the
    kernel writes it — [count, -, -, -] then records {tag, image, anchor, cell | classes + 5 entries}, in ARBITRARY
    arrival order — must give exactly the boxes / objectness / class scores that the full outputs give; a count beyond
    the list's capacity must raise."""
    rng = np.random.default_rng(11)
    anchors = np.asarray(nnet.YOLOV3_ANCHORS, np.float32)
    net = nnet.TNNet.__new__(nnet.TNNet)
    net.w, net.h = 160, 128
    layers, outs = [], []
    n, classes, batch, thresh = 3, 6, 2, 0.6
    for (h, w, mask) in ((4, 5, [6, 7, 8]), (8, 10, [3, 4, 5])):
        out = rng.uniform(0, 1, (batch, n * (classes + 5), h, w)).astype(np.float32)
        yl = nnet.TYoloLayer.__new__(nnet.TYoloLayer)
        yl.w, yl.h, yl.n, yl.classes, yl.mask, yl.biases, yl.batch = w, h, n, classes, mask, anchors, batch
        layers.append(yl)
        outs.append(out)
    net.yolo_layers = lambda: layers
    rec = classes + 9
    records = []
    for tag, (yl, out) in enumerate(zip(layers, outs)):
        p = out.reshape(batch, n, classes + 5, yl.w * yl.h)
        for b in range(batch):
            for a in range(n):
                for i in range(yl.w * yl.h):
                    if p[b, a, 4, i] > thresh:
                        r = np.empty(rec, np.float32)
                        r[:4].view(np.int32)[:] = (tag, b, a, i)
                        r[4:] = p[b, a, :, i]
                        records.append(r)
    assert len(records) > 40
    order = rng.permutation(len(records))                       # atomics: arrival order is arbitrary
    cap = len(records) + 5
    cand = np.zeros(4 + cap * rec, np.float32)
    cand[:1].view(np.uint32)[0] = len(records)
    cand[4:4 + len(records) * rec] = np.concatenate([records[j] for j in order])
    for b in range(batch):
        full = net.Detections(thresh, 0.45, b, host_outputs=outs, img_w=300, img_h=200, relative=True, letter=True)
        comp = net.DetectionsFromCandidates(cand, thresh, 0.45, b, img_w=300, img_h=200, relative=True, letter=True)
        for u, v in zip(full, comp):
            assert u.shape == v.shape and np.array_equal(u, v)
    cand[:1].view(np.uint32)[0] = cap + 1
    with pytest.raises(OverflowError):
        net.DetectionsFromCandidates(cand, thresh, 0.45, 0)
