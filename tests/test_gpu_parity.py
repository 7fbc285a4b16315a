"""GPU parity tests (run with -m gpu on the B200 box).  Every test calls the CUDA path through the
python mirror of the reference API -> ctypes -> C ABI (libdan_b200.so) and compares with the CPU oracle on
the same seeded inputs and with the committed golden fixtures.

Bar: bit-exact (assert_array_equal) for match indices, labels, scores-as-bits, keep-lists and indices;
encoded/decoded fp32 offsets are ALSO compared bit-exact against the oracle (both sides evaluate the same
Cephes exp/log sequence without fma), which is stricter than the 1e-5 relative tolerance of BASELINE.json."""
import os

import numpy as np
import pytest

from dan_b200 import synthetic

from conftest import to_dev

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PS = [0.1, 0.1, 0.2, 0.2]


def _np(t):
    return t.detach().cpu().numpy()


def _gpu_anchors(cfg, device):
    from dan_b200.utility import anchor_manipulator as am
    enc = am.AnchorEncoder(0.4, 0.4, PS)
    return synthetic.build_anchors(enc, cfg)


# ------------------------------------------------------------------------------------------------
# K1 anchors
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,size,border,clip", [("s3fd", (640, 640), None, False), ("dan", (640, 640), None, False),
                                                   ("s3fd", (1600, 1600), 0., False), ("dan", (800, 1024), 0., True),
                                                   ("s3fd", (333, 517), 8., True)])
def test_anchors_bit_exact(cuda, oracle, kind, size, border, clip):
    cfg = synthetic.pyramid_config(kind, size, border=border, clip=clip)
    ref = synthetic.build_anchors(oracle.AnchorEncoder(0.4, 0.4, PS), cfg)
    got = _gpu_anchors(cfg, cuda)
    for r, g in zip(ref[:4], got[:4]):
        np.testing.assert_array_equal(_np(g), r)
    np.testing.assert_array_equal(_np(got[4]), ref[4])


def test_anchors_multi_depth_and_offsets(cuda, oracle):
    from dan_b200.utility import anchor_manipulator as am
    args = dict(scales=[(16., 24.), (64.,)], extra=[(20.,), ()], ratios=[(1., 2., 0.5), (0.8, 1.3)])
    outs = []
    for enc in (oracle.AnchorEncoder(0.5, 0.5, PS), am.AnchorEncoder(0.5, 0.5, PS)):
        hs, ws, ds = [], [], []
        for i in range(2):
            h, w, d = enc.get_anchors_width_height(args["scales"][i], args["extra"][i], args["ratios"][i])
            hs.append(h), ws.append(w), ds.append(d)
        assert ds == [7, 2]
        outs.append(enc.get_all_anchors([120, 200], hs, ws, ds, [(0.5, 0.25), 0.5], [(30, 50), (8, 13)], [4, 16],
                                        [4., 0.], [False, True]))
    for r, g in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(_np(g), r)
    # generate_anchors_by_offset of one layer, [H*W, depth]
    e_ref, e_gpu = oracle.AnchorEncoder(0.5, 0.5, PS), am.AnchorEncoder(0.5, 0.5, PS)
    h, w, d = e_ref.get_anchors_width_height((32.,), (40.,), (1., 2.))
    r = e_ref.generate_anchors_by_offset(h, w, d, [100, 100], (7, 9), 8, offset=0.5)
    g = e_gpu.generate_anchors_by_offset(h, w, d, [100, 100], (7, 9), 8, offset=0.5)
    for a, b in zip(r, g):
        assert tuple(b.shape) == a.shape == (63, 3)
        np.testing.assert_array_equal(_np(b), a)
    assert e_gpu.get_anchors_count(3, (7, 9)) == (63, 189)


# ------------------------------------------------------------------------------------------------
# a6 IoU matrix
# ------------------------------------------------------------------------------------------------
def test_iou_matrix_bit_exact(cuda, oracle, s3fd_anchors_np):
    from dan_b200.utility import anchor_manipulator as am
    a = np.stack(s3fd_anchors_np[:4], -1)
    for gt in (synthetic.gen_faces(2, 50, min_faces=50), synthetic.gen_faces(4, 30, snap=1.0, min_faces=30),
               synthetic.gen_adversarial("outside"), synthetic.gen_adversarial("anchor_identical")):
        ref = oracle.iou_matrix(a, gt)
        got = am.iou_matrix(to_dev(a, cuda), to_dev(gt, cuda))
        np.testing.assert_array_equal(_np(got), ref)
        np.testing.assert_array_equal(_np(am.intersection(to_dev(a, cuda), to_dev(gt, cuda))), oracle.intersection(a, gt))
    np.testing.assert_array_equal(_np(am.areas(to_dev(a, cuda))), oracle.areas(a))


# ------------------------------------------------------------------------------------------------
# a8 SmallMiningMatch on a dense matrix (the custom-op boundary)
# ------------------------------------------------------------------------------------------------
def test_small_mining_match_known_answers(cuda):
    from dan_b200.utility import custom_op
    g = np.load(os.path.join(GOLD, "smm_known_answers.npz"))
    m, s = custom_op.small_mining_match(to_dev(g["test_op_overlaps"], cuda), 0., 0.6, 0.6, 5, 0.1)
    assert _np(m).tolist() == [4, 2, 1, 3, 3] and m.dtype.is_floating_point is False
    np.testing.assert_array_equal(_np(s), g["test_op_scores"])
    m, s = custom_op.small_mining_match(to_dev(g["tie_overlaps"], cuda), 0., .5, .5, 3, .3)
    np.testing.assert_array_equal(_np(m), g["tie_match"])        # libstdc++ heap order on a straddling tie
    np.testing.assert_array_equal(_np(s), g["tie_scores"])


def test_small_mining_match_reference_goldens(cuda):
    from dan_b200.utility import custom_op
    g = np.load(os.path.join(GOLD, "smm_dense_reference.npz"))
    for k in sorted(k[:-2] for k in g.files if k.endswith("_x")):
        a = g[k + "_attrs"]
        m, s = custom_op.small_mining_match(to_dev(g[k + "_x"], cuda), float(a[0]), float(a[1]), float(a[2]), int(a[3]), float(a[4]))
        np.testing.assert_array_equal(_np(m), g[k + "_match"], err_msg=k)
        np.testing.assert_array_equal(_np(s), g[k + "_scores"], err_msg=k)


def test_small_mining_match_random_vs_oracle(cuda, oracle):
    from dan_b200.utility import custom_op
    rng = np.random.default_rng(11)
    for case in range(40):
        n, m = int(rng.integers(1, 3000)), int(rng.integers(1, 90))
        x = rng.uniform(0, 1, (n, m)).astype(np.float32)
        x[x < rng.uniform(0.2, 0.95)] = 0
        if case % 2:
            lv = int(rng.integers(2, 40))
            x = (np.round(x * lv) / lv).astype(np.float32)
        if case % 7 == 0:
            x[:, int(rng.integers(0, m))] = 0            # all-zero column: ties with every anchor (T7)
        attrs = (0., 0.4, float(rng.choice([0.4, 0.6])), int(rng.integers(1, 9)), float(rng.choice([0.0, 0.05, 0.3])))
        rm, rs = oracle.small_mining_match(x, *attrs)
        gm, gs = custom_op.small_mining_match(to_dev(x, cuda), *attrs)
        np.testing.assert_array_equal(_np(gm), rm, err_msg="case %d" % case)
        np.testing.assert_array_equal(_np(gs), rs, err_msg="case %d" % case)


def test_small_mining_match_bucket_overflow(cuda, oracle):
    """> 64 compensation candidates for one GT: the spill path must agree too (incl. a straddling tie)."""
    from dan_b200.utility import custom_op
    rng = np.random.default_rng(5)
    n = 1000
    x = np.zeros((n, 3), np.float32)
    x[:, 0] = rng.uniform(0.31, 0.39, n)          # nobody reaches pos=0.4, everybody is a candidate of GT 0
    x[:, 1] = np.round(rng.uniform(0.31, 0.39, n) * 50) / 50
    x[5, 2] = 0.9
    for attrs in [(0., 0.4, 0.4, 6, 0.3), (0., 0.4, 0.4, 200, 0.3)]:
        rm, rs = oracle.small_mining_match(x, *attrs)
        gm, gs = custom_op.small_mining_match(to_dev(x, cuda), *attrs)
        np.testing.assert_array_equal(_np(gm), rm)
        np.testing.assert_array_equal(_np(gs), rs)


def test_small_mining_match_errors(cuda):
    import torch
    from dan_b200 import _lib
    from dan_b200.utility import custom_op
    x = torch.zeros((8, 3), device=cuda)
    for bad in [(-0.1, .4, .4, 6, .3), (0., 0., .4, 6, .3), (0., .5, .4, 6, .3), (0., .4, 1., 6, .3), (0., .4, .4, 0, .3),
                (0., .4, .4, 6, 1.)]:
        with pytest.raises(_lib.DanError) as e:
            custom_op.small_mining_match(x, *bad)
        assert e.value.code == -1 and "Need Attr" in str(e.value)
    with pytest.raises(_lib.DanError):
        custom_op.small_mining_match(torch.zeros(8, device=cuda), 0., .4, .4, 6, .3)     # rank != 2 (:311)
    m, s = custom_op.small_mining_match(torch.zeros((5, 0), device=cuda), 0., .4, .4, 6, .3)
    assert _np(m).tolist() == [-2] * 5


# ------------------------------------------------------------------------------------------------
# a7 do_dual_max_match on a dense matrix
# ------------------------------------------------------------------------------------------------
def test_dual_max_match_vs_oracle(cuda, oracle):
    from dan_b200.utility import anchor_manipulator as am
    rng = np.random.default_rng(13)
    for case in range(40):
        n, m = int(rng.integers(1, 3000)), int(rng.integers(1, 90))
        x = rng.uniform(0, 1, (n, m)).astype(np.float32)
        x[x < rng.uniform(0.2, 0.95)] = 0
        if case % 2:
            lv = int(rng.integers(2, 40))
            x = (np.round(x * lv) / lv).astype(np.float32)
        if case % 5 == 0:
            x[:, int(rng.integers(0, m))] = 0
        kw = dict(ignore_between=bool(case % 3), gt_max_first=bool(case % 4))
        low, high = (0.35, 0.35) if case % 2 else (0.3, 0.5)
        rm, rs = oracle.do_dual_max_match(x, low, high, **kw)
        gm, gs = am.do_dual_max_match(to_dev(x, cuda), low, high, **kw)
        assert gm.dtype.itemsize == 8
        np.testing.assert_array_equal(_np(gm), rm, err_msg="case %d %s" % (case, kw))
        np.testing.assert_array_equal(_np(gs), rs, err_msg="case %d %s" % (case, kw))


# ------------------------------------------------------------------------------------------------
# a9 / a10 fused encode
# ------------------------------------------------------------------------------------------------
def _check_encode(ref, got, name=""):
    targets, labels, scores, matched = [_np(t) for t in got[:4]]
    np.testing.assert_array_equal(labels, ref[1], err_msg=name + " labels")
    np.testing.assert_array_equal(scores, ref[2], err_msg=name + " scores")
    np.testing.assert_array_equal(targets, ref[0], err_msg=name + " targets")
    np.testing.assert_array_equal(matched, ref[3], err_msg=name + " matched_gt")
    # bit-exact including the sign of zero: the reference multiplies every target / matched box by float(positive)
    # (anchor_manipulator.py:324,326), which leaves -0.0 where the raw value of a non-positive anchor is negative
    np.testing.assert_array_equal(np.signbit(targets), np.signbit(ref[0]), err_msg=name + " sign of targets")
    np.testing.assert_array_equal(np.signbit(matched), np.signbit(ref[3]), err_msg=name + " sign of matched_gt")
    assert labels.dtype == np.int64 and targets.dtype == np.float32


def test_encode_goldens(cuda, s3fd_anchors_np, dan_anchors_np):
    """fixtures produced with the REFERENCE's own matcher functor (tests/golden/make_golden.py)."""
    from dan_b200.utility import anchor_manipulator as am
    g = np.load(os.path.join(GOLD, "encode_goldens.npz"))
    for name in sorted(k[:-4] for k in g.files if k.endswith("__gt")):
        pos, ign, mining, pa = g[name + "__cfg"]
        anchors_np = dan_anchors_np if name.startswith("dan") else s3fd_anchors_np
        anchors = [to_dev(v, cuda) for v in anchors_np]
        e = am.AnchorEncoder(float(pos), float(ign), PS)
        gt = to_dev(g[name + "__gt"], cuda)
        if pa < 0:
            t, l, s, mg = e.encode_anchors(gt, *anchors, match_mining=bool(mining))
        else:
            t, l, s, mg = e.encode_pa_anchors(gt, *anchors, float(ign), float(pos), match_mining=bool(mining), scale=float(pa))
        labels = np.zeros(l.shape[0], np.int64)
        labels[g[name + "__nonzero_rows"]] = g[name + "__nonzero_labels"]
        np.testing.assert_array_equal(_np(l), labels, err_msg=name)
        np.testing.assert_array_equal(_np(s), g[name + "__scores"], err_msg=name)
        rows = g[name + "__pos_rows"]
        np.testing.assert_array_equal(_np(t)[rows], g[name + "__pos_targets"], err_msg=name)
        np.testing.assert_array_equal(_np(mg)[rows], g[name + "__pos_matched"], err_msg=name)
        assert not _np(t)[labels != 1].any() and not _np(mg)[labels != 1].any()


@pytest.mark.parametrize("mining", [True, False])
def test_encode_per_image_vs_oracle(cuda, oracle, s3fd_anchors_np, mining):
    from dan_b200.utility import anchor_manipulator as am
    thr = 0.4 if mining else 0.5
    e_ref, e_gpu = oracle.AnchorEncoder(thr, thr, PS), am.AnchorEncoder(thr, thr, PS)
    anchors = [to_dev(v, cuda) for v in s3fd_anchors_np]
    gts = [synthetic.gen_faces(i, 50) for i in range(20, 26)]
    gts += [synthetic.gen_faces(i, 50, snap=s) for i, s in ((30, 1.0), (31, 4.0), (32, 4.0), (33, 8.0))]
    gts += [synthetic.gen_adversarial(k) for k in ("empty", "outside", "duplicate", "anchor_identical", "tiny", "huge")]
    for i, gt in enumerate(gts):
        ref = e_ref.encode_anchors(gt, *s3fd_anchors_np, match_mining=mining)
        got = e_gpu.encode_anchors(to_dev(gt, cuda), *anchors, match_mining=mining)
        _check_encode(ref, got, "gt set %d" % i)


def test_encode_inside_mask_and_debug(cuda, oracle):
    """eval-style border 0 -> 2424 anchors masked out; debug=True returns raw anchors."""
    from dan_b200.utility import anchor_manipulator as am
    cfg = synthetic.pyramid_config("s3fd", border=0.)
    a_np = synthetic.build_anchors(oracle.AnchorEncoder(0.4, 0.4, PS), cfg)
    assert int(a_np[4].sum()) == 31701
    a = [to_dev(v, cuda) for v in a_np]
    for mining in (True, False):
        for gt in (synthetic.gen_faces(41, 50), synthetic.gen_adversarial("outside"), synthetic.gen_adversarial("huge")):
            for debug in (False, True):
                ref = oracle.AnchorEncoder(0.4, 0.4, PS).encode_anchors(gt, *a_np, match_mining=mining, debug=debug)
                got = am.AnchorEncoder(0.4, 0.4, PS).encode_anchors(to_dev(gt, cuda), *a, match_mining=mining, debug=debug)
                _check_encode(ref, got)


def test_encode_pa_anchors_vs_oracle(cuda, oracle, s3fd_anchors_np):
    """PyramidBox: face (mining, scale 1), head (dual, scale 2, anchors[N0:]), body (dual, scale 4, anchors[N0+N1:])
    train_pb.py:205-225."""
    from dan_b200.utility import anchor_manipulator as am
    e_ref, e_gpu = oracle.AnchorEncoder(0.4, 0.4, PS), am.AnchorEncoder(0.4, 0.4, PS)
    n0, n1 = 160 * 160, 80 * 80
    for seed in (50, 51):
        gt = synthetic.gen_faces(seed, 40)
        for start, ign, pos, mining, scale in [(0, 0.4, 0.4, True, 1.), (n0, 0.35, 0.35, False, 2.), (n0 + n1, 0.35, 0.35, False, 4.)]:
            sl_np = [v[start:] for v in s3fd_anchors_np]
            sl = [to_dev(v, cuda) for v in sl_np]
            ref = e_ref.encode_pa_anchors(gt, *sl_np, ign, pos, match_mining=mining, scale=scale)
            got = e_gpu.encode_pa_anchors(to_dev(gt, cuda), *sl, ign, pos, match_mining=mining, scale=scale)
            _check_encode(ref, got, "start %d scale %g" % (start, scale))


def test_encode_batch_matches_per_image(cuda, oracle, dan_anchors_np):
    """config-3 style: DAN anchors, dual matcher 0.35, dense tiny faces (up to 1000 / image), CSR batch with an
    empty image in the middle; also checks the int32 match output."""
    from dan_b200.utility import anchor_manipulator as am
    e_ref, e_gpu = oracle.AnchorEncoder(0.35, 0.35, PS), am.AnchorEncoder(0.35, 0.35, PS)
    anchors = [to_dev(v, cuda) for v in dan_anchors_np]
    gts = [synthetic.gen_dense_tiny(0, lo=990, hi=1000), synthetic.gen_adversarial("empty"), synthetic.gen_faces(60, 50),
           synthetic.gen_dense_tiny(1, lo=200, hi=260), synthetic.gen_faces(61, 3, snap=4.0)]
    cat, offs = synthetic.to_csr(gts)
    res = e_gpu.encode_anchors_batch(to_dev(cat, cuda), to_dev(offs, cuda), *anchors, match_mining=False, want_match=True)
    assert tuple(res.targets.shape) == (5, 34125, 4) and tuple(res.labels.shape) == (5, 34125)
    for b, gt in enumerate(gts):
        ref = e_ref.encode_anchors(gt, *dan_anchors_np, match_mining=False, return_match=True)
        _check_encode(ref, [res.targets[b], res.labels[b], res.scores[b], res.matched_gt[b]], "image %d" % b)
        np.testing.assert_array_equal(_np(res.match[b]), ref[4].astype(np.int32))


def test_encode_batch_mining_config2(cuda, oracle, s3fd_anchors_np):
    """config 2: S3FD 640, mining, thresholds 0.4/0.4, batch 32, <= 50 faces / image."""
    from dan_b200.utility import anchor_manipulator as am
    e_ref, e_gpu = oracle.AnchorEncoder(0.4, 0.4, PS), am.AnchorEncoder(0.4, 0.4, PS)
    anchors = [to_dev(v, cuda) for v in s3fd_anchors_np]
    gts = [synthetic.gen_faces(100 + i, 50, snap=(2.0 if i % 8 == 7 else 0.0)) for i in range(32)]
    cat, offs = synthetic.to_csr(gts)
    res = e_gpu.encode_all_anchors(to_dev(cat, cuda), to_dev(offs, cuda), *anchors, match_mining=True, want_match=True)
    n_comp = 0
    for b, gt in enumerate(gts):
        ref = e_ref.encode_anchors(gt, *s3fd_anchors_np, match_mining=True, return_match=True)
        _check_encode(ref, [res.targets[b], res.labels[b], res.scores[b], res.matched_gt[b]], "image %d" % b)
        np.testing.assert_array_equal(_np(res.match[b]), ref[4].astype(np.int32))
        n_comp += int(((ref[1] == 1) & (ref[2] < 0.4)).sum())
    assert n_comp > 0     # the compensation stage really fired in this batch


def test_encode_roundtrip_property_full_size(cuda, s3fd_anchors_np):
    """size-independent property at BASELINE config-3 scale (batch 64): decoding the encoded targets of every
    positive anchor returns its matched GT box; labels/targets/matched are consistent."""
    import torch
    from dan_b200.utility import anchor_manipulator as am
    e = am.AnchorEncoder(0.35, 0.35, PS)
    anchors = [to_dev(v, cuda) for v in s3fd_anchors_np]
    gts = [synthetic.gen_dense_tiny(200 + i) if i % 2 else synthetic.gen_faces(200 + i, 50) for i in range(64)]
    cat, offs = synthetic.to_csr(gts)
    res = e.encode_anchors_batch(to_dev(cat, cuda), to_dev(offs, cuda), *anchors, match_mining=False, want_match=True)
    dec = e.batch_decode_anchors(res.targets, *anchors[:4])
    pos = res.labels == 1
    assert int(pos.sum()) > 64
    assert torch.equal(pos, res.match >= 0) and torch.equal(res.labels == -1, res.match == -2)
    torch.testing.assert_close(dec[pos], res.matched_gt[pos], rtol=1e-5, atol=2e-3)
    assert not res.targets[~pos].any() and not res.matched_gt[~pos].any()
    # matched GT indices stay inside the image's own CSR segment
    offs_l = offs.tolist()
    for b in (0, 1, 63):
        m = res.match[b]
        assert int(m.max()) < offs_l[b + 1] - offs_l[b] and int(m.min()) >= -2
    # gathered matched_gt equals gt[match]
    b = 5
    gt_b = to_dev(gts[b], cuda)
    m = res.match[b]
    assert torch.equal(res.matched_gt[b][m >= 0], gt_b[m[m >= 0].long()])


# ------------------------------------------------------------------------------------------------
# a11 decode
# ------------------------------------------------------------------------------------------------
def test_decode_bit_exact(cuda, oracle, s3fd_anchors_np):
    from dan_b200.utility import anchor_manipulator as am
    rng = np.random.default_rng(17)
    pred = rng.normal(0, 1.5, (3, 34125, 4)).astype(np.float32)
    pred[0, :100] = 0
    pred[1, :50, 2:] = 30.0          # large exp arguments
    e_ref, e_gpu = oracle.AnchorEncoder(None, None, PS), am.AnchorEncoder(None, None, PS)
    anchors = [to_dev(v, cuda) for v in s3fd_anchors_np[:4]]
    np.testing.assert_array_equal(_np(e_gpu.batch_decode_anchors(to_dev(pred, cuda), *anchors)),
                                  e_ref.batch_decode_anchors(pred, *s3fd_anchors_np[:4]))
    np.testing.assert_array_equal(_np(e_gpu.decode_anchors(to_dev(pred[2], cuda), *anchors)),
                                  e_ref.decode_anchors(pred[2], *s3fd_anchors_np[:4]))
    g = np.load(os.path.join(GOLD, "decode_roundtrip.npz"))
    np.testing.assert_array_equal(_np(e_gpu.decode_anchors(to_dev(g["targets"], cuda), *[a[:1] for a in anchors])), g["decoded"])


# ------------------------------------------------------------------------------------------------
# a12-a18 bbox_util
# ------------------------------------------------------------------------------------------------
def test_bbox_util_elementwise(cuda, oracle):
    from dan_b200 import functional as F
    from dan_b200.utility import bbox_util as bu
    rng = np.random.default_rng(19)
    n = 5000
    logits = rng.normal(0, 4, (n, 3)).astype(np.float32)
    boxes = (rng.uniform(-50, 700, (n, 4))).astype(np.float32)
    sm_ref = oracle.softmax(logits)
    sm = F.softmax(to_dev(logits, cuda))
    np.testing.assert_array_equal(_np(sm), sm_ref)
    sb_ref, ss_ref = oracle.select_bboxes(sm_ref, boxes, 3, 0.3)
    sb, ss = bu.select_bboxes(sm, to_dev(boxes, cuda), 3, 0.3)
    for c in (1, 2):
        np.testing.assert_array_equal(_np(sb[c]), sb_ref[c])
        np.testing.assert_array_equal(_np(ss[c]), ss_ref[c])
    cols = [boxes[:, i] for i in range(4)]
    dcols = [to_dev(c, cuda) for c in cols]
    ref = oracle.clip_bboxes(*cols, 640, 480)
    got = bu.clip_bboxes(*dcols, 640, 480)
    for r, g in zip(ref, got):
        np.testing.assert_array_equal(_np(g), r)
    ref = oracle.filter_bboxes(ss_ref[1], *ref, 12.5)
    got = bu.filter_bboxes(ss[1], *got, 12.5)
    for r, g in zip(ref, got):
        np.testing.assert_array_equal(_np(g), r)
    np.testing.assert_array_equal(_np(bu.bbox_point2center(to_dev(boxes, cuda))), oracle.bbox_point2center(boxes))
    np.testing.assert_array_equal(_np(bu.bbox_center2point(to_dev(boxes, cuda))), oracle.bbox_center2point(boxes))


@pytest.mark.parametrize("n,k", [(34125, 5000), (213294, 5000), (3000, 5000), (9000, 100), (1, 4), (8192, 8192),
                                 (20000, 15000), (70000, 70000)])       # beyond the in-shared-memory sort: runs + merges in HBM
def test_sort_bboxes_vs_oracle(cuda, oracle, n, k):
    from dan_b200.utility import bbox_util as bu
    rng = np.random.default_rng(n + k)
    scores = rng.uniform(0, 1, n).astype(np.float32)
    scores[rng.uniform(0, 1, n) < 0.7] = 0                     # many zero rows, like after select/filter
    scores[::97] = np.float32(0.5)                             # exact ties -> lower index first
    cols = [rng.uniform(0, 600, n).astype(np.float32) for _ in range(4)]
    ref = oracle.sort_bboxes(scores, *cols, k)
    got = bu.sort_bboxes(to_dev(scores, cuda), *[to_dev(c, cuda) for c in cols], k)
    for r, g in zip(ref[:5], got):
        assert g.shape[0] == k
        np.testing.assert_array_equal(_np(g), r)


@pytest.mark.parametrize("n,topk,thr", [(5000, 750, 0.3), (2000, 50, 0.5), (300, 750, 0.3), (64, 10, 0.0), (65, 100, 0.7),
                                       (8192, 750, 0.3), (12000, 750, 0.3), (20000, 20000, 0.6)])   # > 8192: no size limit
def test_nms_bboxes_vs_oracle(cuda, oracle, n, topk, thr):
    from dan_b200.utility import bbox_util as bu
    rng = np.random.default_rng(n + topk)
    centres = rng.uniform(0, 640, (max(n // 12, 1), 2))
    c = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 4, (n, 2))
    wh = rng.uniform(8, 60, (n, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    boxes[::50] = boxes[::50][:, [2, 3, 0, 1]]                 # flipped corners are normalised by TF
    boxes[::77, 2] = boxes[::77, 0]                            # zero-area boxes never suppress
    scores = rng.permutation(n).astype(np.float32) / n
    rs, rb, rk = oracle.nms_bboxes_with_padding(scores, boxes, topk, thr)
    gs, gb = bu.nms_bboxes_with_padding(to_dev(scores, cuda), to_dev(boxes, cuda), topk, thr)
    np.testing.assert_array_equal(_np(gs), rs)
    np.testing.assert_array_equal(_np(gb), rb)
    vs, vb = bu.nms_bboxes(to_dev(scores, cuda), to_dev(boxes, cuda), topk, thr)
    np.testing.assert_array_equal(_np(vs), scores[rk])
    np.testing.assert_array_equal(_np(vb), boxes[rk])


def test_nms_dense_cluster_uses_fallback(cuda, oracle):
    """hundreds of mutually overlapping boxes: candidates have more than 64 suppressors, which switches the resolve
    kernel from the suppressor-list relaxation to its round-based fallback; both must give TF's greedy result."""
    from dan_b200.utility import bbox_util as bu
    rng = np.random.default_rng(31)
    for n, topk, thr in ((400, 750, 0.3), (1500, 100, 0.5), (3000, 750, 0.3)):
        c = np.array([[300., 300.]]) + rng.normal(0, 12, (n, 2))
        c[n // 2:] = rng.uniform(50, 600, (n - n // 2, 2))          # half in one dense cluster, half scattered
        wh = rng.uniform(60, 90, (n, 2))
        boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        scores = rng.permutation(n).astype(np.float32) / n
        rs, rb, rk = oracle.nms_bboxes_with_padding(scores, boxes, topk, thr)
        gs, gb = bu.nms_bboxes_with_padding(to_dev(scores, cuda), to_dev(boxes, cuda), topk, thr)
        np.testing.assert_array_equal(_np(gs), rs)
        np.testing.assert_array_equal(_np(gb), rb)


def test_nms_score_ties_are_stable(cuda, oracle):
    """documented tie policy T6: equal scores keep input order (oracle tie='stable')."""
    from dan_b200 import functional as F
    rng = np.random.default_rng(23)
    n = 400
    c = rng.uniform(0, 200, (n, 2))
    boxes = np.concatenate([c, c + rng.uniform(10, 50, (n, 2))], 1).astype(np.float32)
    scores = (rng.integers(0, 6, n) / 8).astype(np.float32)
    ref = oracle.tf_non_max_suppression(boxes, scores, 100, 0.3, tie="stable")
    _, _, keep, cnt = F.nms_boxes(to_dev(scores, cuda), to_dev(boxes, cuda), 100, 0.3)
    assert int(cnt.item()) == len(ref)
    np.testing.assert_array_equal(_np(keep)[:len(ref)], ref)


# ------------------------------------------------------------------------------------------------
# a17 fused parse_by_class
# ------------------------------------------------------------------------------------------------
def _oracle_parse(oracle, size, cls, loc, anchors_np, **kw):
    enc = oracle.AnchorEncoder(None, None, PS)
    boxes = enc.decode_anchors(loc, *anchors_np[:4])
    p = dict(num_classes=cls.shape[1], select_threshold=0.01, min_size=0, keep_topk=5000, nms_topk=750, nms_threshold=0.3)
    p.update(kw)
    return boxes, oracle.parse_by_class(list(size), cls, boxes, p["num_classes"], p["select_threshold"], p["min_size"],
                                        p["keep_topk"], p["nms_topk"], p["nms_threshold"], return_indices=True)


def test_parse_by_class_golden(cuda, oracle):
    from dan_b200.utility import bbox_util as bu
    g = np.load(os.path.join(GOLD, "postprocess_golden.npz"))
    enc = oracle.AnchorEncoder(None, None, PS)
    a_np = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd", border=0.))
    cls, loc, _ = synthetic.gen_predictions(int(g["image_index"]), np.stack(a_np[:4], -1), max_faces=int(g["max_faces"]))
    boxes = enc.decode_anchors(loc, *a_np[:4])
    sb, ss = bu.parse_by_class([640, 640], to_dev(cls, cuda), to_dev(boxes, cuda), 2, 0.01, 0, 5000, 750, 0.3)
    np.testing.assert_array_equal(_np(ss[1]), g["scores"])
    np.testing.assert_array_equal(_np(sb[1]), g["boxes"])


@pytest.mark.parametrize("size,faces", [((640, 640), 300), ((800, 800), 120), ((1024, 1024), 60)])
def test_parse_by_class_batch_vs_oracle(cuda, oracle, size, faces):
    """config 4: decode + threshold 0.01 + top-k 5000 + NMS 0.3 -> 750, anchors regenerated per size (border 0)."""
    from dan_b200.utility import bbox_util as bu
    cfg = synthetic.pyramid_config("s3fd", size, border=0.)
    a_np = synthetic.build_anchors(oracle.AnchorEncoder(None, None, PS), cfg)
    an = np.stack(a_np[:4], -1)
    anchors = [to_dev(v, cuda) for v in a_np[:4]]
    preds = [synthetic.gen_predictions(300 + i, an, size=size, max_faces=faces) for i in range(3)]
    cls = np.stack([p[0] for p in preds])
    loc = np.stack([p[1] for p in preds])
    det = bu.parse_by_class_batch(list(size), to_dev(cls, cuda), 2, 0.01, 0, 5000, 750, 0.3, loc_pred=to_dev(loc, cuda),
                                  anchors=anchors)
    for b in range(3):
        boxes, (sb, ss, idx) = _oracle_parse(oracle, size, cls[b], loc[b], a_np)
        np.testing.assert_array_equal(_np(det.scores[b, 0]), ss[1], err_msg="image %d" % b)
        np.testing.assert_array_equal(_np(det.boxes[b, 0]), sb[1], err_msg="image %d" % b)
        top_idx, keep = idx[1]
        k = int((ss[1] > 0).sum())
        assert int(det.counts[b, 0]) == k
        np.testing.assert_array_equal(_np(det.keep_pos[b, 0])[:len(keep)], keep)        # NMS keep-list, bit exact
        np.testing.assert_array_equal(_np(det.anchor_index[b, 0])[:k], top_idx[keep[:k]])
        assert (_np(det.anchor_index[b, 0])[k:] == -1).all()
    # same result when parse_by_class is given decoded boxes (what the reference function literally takes)
    boxes0, _ = _oracle_parse(oracle, size, cls[0], loc[0], a_np)
    det2 = bu.parse_by_class_batch(list(size), to_dev(cls[:1], cuda), 2, 0.01, 0, 5000, 750, 0.3, bboxes_pred=to_dev(boxes0[None], cuda))
    np.testing.assert_array_equal(_np(det2.scores[0]), _np(det.scores[0]))
    np.testing.assert_array_equal(_np(det2.boxes[0]), _np(det.boxes[0]))


def test_parse_by_class_variants(cuda, oracle):
    """3 classes, min_size, small keep_topk (radix-select path: more survivors than keep_topk), low nms_topk."""
    from dan_b200.utility import bbox_util as bu
    cfg = synthetic.pyramid_config("s3fd", (640, 640), border=0.)
    a_np = synthetic.build_anchors(oracle.AnchorEncoder(None, None, PS), cfg)
    n = a_np[0].shape[0]
    rng = np.random.default_rng(29)
    cls = rng.normal(0, 2.0, (n, 3)).astype(np.float32)       # ~ everything passes the threshold
    loc = rng.normal(0, 0.5, (n, 4)).astype(np.float32)
    anchors = [to_dev(v, cuda) for v in a_np[:4]]
    for kw in (dict(select_threshold=0.2, min_size=10, keep_topk=400, nms_topk=100, nms_threshold=0.45),
               dict(select_threshold=0.0, min_size=0, keep_topk=5000, nms_topk=750, nms_threshold=0.3),
               dict(select_threshold=0.999999, min_size=0, keep_topk=100, nms_topk=20, nms_threshold=0.3)):
        boxes, (sb, ss, idx) = _oracle_parse(oracle, (640, 640), cls, loc, a_np, num_classes=3, **kw)
        det = bu.parse_by_class_batch([640, 640], to_dev(cls[None], cuda), 3, kw["select_threshold"], kw["min_size"],
                                      kw["keep_topk"], kw["nms_topk"], kw["nms_threshold"], loc_pred=to_dev(loc[None], cuda),
                                      anchors=anchors)
        for c in (1, 2):
            np.testing.assert_array_equal(_np(det.scores[0, c - 1]), ss[c], err_msg=str(kw))
            np.testing.assert_array_equal(_np(det.boxes[0, c - 1]), sb[c], err_msg=str(kw))
            keep = idx[c][1]
            np.testing.assert_array_equal(_np(det.keep_pos[0, c - 1])[:len(keep)], keep)


def test_parse_by_class_host_resident_geometry(cuda, oracle):
    """loc_pred / bboxes_pred in PINNED HOST memory (include/dan_b200.h, host-resident geometry): the kernels fetch only the
    rows that pass the threshold, in place; results equal the oracle's and the device-resident run's bit for bit - two
    classes (fused filter), three classes (separate filter kernel), decoded boxes, eagerly and replayed from a CUDA graph.
    Pageable host memory is refused."""
    import ctypes
    import torch
    from dan_b200 import _lib as L, functional as F
    from dan_b200.utility import bbox_util as bu
    size = (640, 640)
    cfg = synthetic.pyramid_config("s3fd", size, border=0.)
    a_np = synthetic.build_anchors(oracle.AnchorEncoder(None, None, PS), cfg)
    an = np.stack(a_np[:4], -1)
    anchors = [to_dev(v, cuda) for v in a_np[:4]]
    preds = [synthetic.gen_predictions(520 + i, an, size=size, max_faces=80) for i in range(3)]
    cls = np.stack([p[0] for p in preds])
    loc = np.stack([p[1] for p in preds])
    loc_host = torch.from_numpy(loc).pin_memory()
    args = (list(size), to_dev(cls, cuda), 2, 0.01, 0, 5000, 750, 0.3)
    ref = bu.parse_by_class_batch(*args, loc_pred=to_dev(loc, cuda), anchors=anchors)
    det = bu.parse_by_class_batch(*args, loc_pred=loc_host, anchors=anchors)
    for name in ("boxes", "scores", "counts", "anchor_index", "keep_pos"):
        assert torch.equal(getattr(ref, name), getattr(det, name)), name
    _, (sb, ss, _) = _oracle_parse(oracle, size, cls[1], loc[1], a_np)
    np.testing.assert_array_equal(_np(det.scores[1, 0]), ss[1])
    np.testing.assert_array_equal(_np(det.boxes[1, 0]), sb[1])
    # decoded boxes on the host
    boxes = np.stack([_oracle_parse(oracle, size, cls[b], loc[b], a_np)[0] for b in range(3)]).astype(np.float32)
    det_b = bu.parse_by_class_batch(*args, bboxes_pred=torch.from_numpy(boxes).pin_memory())
    assert torch.equal(det_b.scores, ref.scores) and torch.equal(det_b.boxes, ref.boxes)
    # CUDA graph: the host pointer is baked into the graph; new rows written by the host between replays are picked up
    params = F.postprocess_params(2, size, 0.01, 0, 5000, 750, 0.3, prior_scaling=PS)
    cls_d = to_dev(cls, cuda)
    ws = L.Workspace()
    out = F.postprocess_batch(params, cls_d, loc_pred=loc_host, anchors=anchors, workspace=ws)      # sizes the workspace
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            out = F.postprocess_batch(params, cls_d, loc_pred=loc_host, anchors=anchors, workspace=ws,
                                      out=(out.boxes, out.scores, out.counts, out.anchor_index, out.keep_pos))
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.boxes, ref.boxes) and torch.equal(out.scores, ref.scores)
    loc_host.copy_(torch.from_numpy(loc[::-1].copy()))                     # other offsets, same buffer
    ref2 = bu.parse_by_class_batch(*args, loc_pred=to_dev(loc[::-1].copy(), cuda), anchors=anchors)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.boxes, ref2.boxes) and torch.equal(out.scores, ref2.scores) and torch.equal(out.counts, ref2.counts)
    # three classes: the separate filter kernel stashes, the NMS kernel reads the stash
    rng = np.random.default_rng(31)
    n = a_np[0].shape[0]
    cls3 = rng.normal(0, 2.0, (1, n, 3)).astype(np.float32)
    loc3 = rng.normal(0, 0.5, (1, n, 4)).astype(np.float32)
    a3 = ([640, 640], to_dev(cls3, cuda), 3, 0.2, 10, 400, 100, 0.45)
    r3 = bu.parse_by_class_batch(*a3, loc_pred=to_dev(loc3, cuda), anchors=anchors)
    d3 = bu.parse_by_class_batch(*a3, loc_pred=torch.from_numpy(loc3).pin_memory(), anchors=anchors)
    assert torch.equal(r3.boxes, d3.boxes) and torch.equal(r3.scores, d3.scores) and torch.equal(r3.keep_pos, d3.keep_pos)
    # pageable host memory: python mirror and the C ABI both refuse
    with pytest.raises(TypeError):
        bu.parse_by_class_batch(*args, loc_pred=torch.from_numpy(loc), anchors=anchors)
    pageable = np.ascontiguousarray(loc)
    nb = L.lib().dan_postprocess_workspace_bytes(n, 3, 2, 5000)
    w = ws.get(nb, cls_d.device)
    rc = L.lib().dan_postprocess_batch(ctypes.byref(params), L.dev_ptr(cls_d), ctypes.c_void_p(pageable.ctypes.data), ctypes.c_void_p(0),
                                       *[L.dev_ptr(a) for a in anchors], n, 3, L.dev_ptr(out.boxes), L.dev_ptr(out.scores),
                                       L.dev_ptr(out.counts), ctypes.c_void_p(0), ctypes.c_void_p(0), L.dev_ptr(w), nb, L.stream_ptr())
    assert rc == -1            # DAN_ERR_INVALID_ARGUMENT
    torch.cuda.synchronize()


def test_postprocess_properties_full_size(cuda):
    """size-independent properties at 1600^2 (213 294 anchors), batch 4: sorted scores, counts consistent, kept boxes
    mutually below the NMS threshold (no +1 IoU), idempotence of NMS on its own output."""
    import torch
    from dan_b200 import functional as F
    from dan_b200.utility import anchor_manipulator as am, bbox_util as bu
    size = (1600, 1600)
    cfg = synthetic.pyramid_config("s3fd", size, border=0.)
    anchors = synthetic.build_anchors(am.AnchorEncoder(None, None, PS), cfg)
    assert anchors[0].numel() == 213294
    an = np.stack([_np(a) for a in anchors[:4]], -1)
    preds = [synthetic.gen_predictions(400 + i, an, size=size, max_faces=300) for i in range(4)]
    cls = to_dev(np.stack([p[0] for p in preds]), cuda)
    loc = to_dev(np.stack([p[1] for p in preds]), cuda)
    det = bu.parse_by_class_batch(list(size), cls, 2, 0.01, 0, 5000, 750, 0.3, loc_pred=loc, anchors=anchors[:4])
    for b in range(4):
        k = int(det.counts[b, 0])
        s, bx = det.scores[b, 0], det.boxes[b, 0]
        assert 0 < k <= 750 and bool((s[:k] > 0.01).all()) and not s[k:].any() and not bx[k:].any()
        assert bool((s[:k - 1] >= s[1:k]).all())
        area = (bx[:k, 2] - bx[:k, 0]) * (bx[:k, 3] - bx[:k, 1])
        ih = (torch.minimum(bx[:k, None, 2], bx[None, :k, 2]) - torch.maximum(bx[:k, None, 0], bx[None, :k, 0])).clamp(min=0)
        iw = (torch.minimum(bx[:k, None, 3], bx[None, :k, 3]) - torch.maximum(bx[:k, None, 1], bx[None, :k, 1])).clamp(min=0)
        iou = ih * iw / (area[:, None] + area[None, :] - ih * iw)
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= 0.3 + 1e-5
        s2, b2, keep2, cnt2 = F.nms_boxes(s[:k].contiguous(), bx[:k].contiguous(), 750, 0.3)
        assert int(cnt2.item()) == k and torch.equal(s2[:k], s[:k])


def test_postprocess_errors(cuda):
    import torch
    from dan_b200 import _lib
    from dan_b200.utility import bbox_util as bu
    cls = torch.zeros((1, 100, 2), device=cuda)
    box = torch.zeros((1, 100, 4), device=cuda)
    with pytest.raises(_lib.DanError):
        bu.parse_by_class_batch([64, 64], cls, 2, -0.5, 0, 100, 10, 0.3, bboxes_pred=box)
    with pytest.raises(_lib.DanError):
        bu.parse_by_class_batch([64, 64], cls, 2, 0.1, 0, 0, 10, 0.3, bboxes_pred=box)           # keep_topk must be >= 1
    det = bu.parse_by_class_batch([64, 64], cls, 2, 0.1, 0, 100000, 10, 0.3, bboxes_pred=box)     # keep_topk > N is fine
    assert int(det.counts.sum()) == 0
    with pytest.raises(TypeError):
        bu.parse_by_class_batch([64, 64], cls.cpu(), 2, 0.1, 0, 100, 10, 0.3, bboxes_pred=box)      # no CPU fallback
    det = bu.parse_by_class_batch([64, 64], cls, 2, 0.9, 0, 100, 10, 0.3, bboxes_pred=box)
    assert int(det.counts.sum()) == 0 and not det.scores.any()


# ------------------------------------------------------------------------------------------------
# hard-negative mining
# ------------------------------------------------------------------------------------------------
def _hnm_inputs(seed, B, N, C=2, quantise=False, pos_rate=0.02):
    rng = np.random.default_rng(seed)
    cls = rng.normal(0., 3., size=(B, N, C)).astype(np.float32)
    if quantise:
        cls = np.round(cls)
    tg = rng.choice([-1, 0, 1], p=[0.05, 0.95 - pos_rate, pos_rate], size=(B, N)).astype(np.int64)
    loc = rng.normal(size=(B * N, 4)).astype(np.float32)
    lt = rng.normal(size=(B, N, 4)).astype(np.float32)
    return cls, loc, tg, lt


def _check_hnm(ref, aux, got, det):
    for name, r, g in zip(("cls_pred", "location_pred", "cls_targets", "loc_targets"), ref, got):
        assert tuple(g.shape) == r.shape, name
        np.testing.assert_array_equal(_np(g), r, err_msg=name)
    np.testing.assert_array_equal(_np(det.final_mask).astype(bool), aux["final_mask"])
    np.testing.assert_array_equal(_np(det.n_neg_select), aux["n_neg_select"])
    np.testing.assert_array_equal(_np(det.score_at_k).view(np.uint32), aux["score_at_k"].view(np.uint32))


@pytest.mark.parametrize("B,N,C,quantise", [(4, 3000, 2, False), (3, 34125, 2, False), (2, 5000, 3, True), (5, 777, 2, True)])
def test_hard_negative_mining_vs_oracle(cuda, oracle, B, N, C, quantise):
    from dan_b200.utility import hard_negative_mining as hnm
    cls, loc, tg, lt = _hnm_inputs(70 + B, B, N, C, quantise)
    tg[-1] = np.where(tg[-1] > 0, 0, tg[-1])                # one image without positives: the max(., 1) clamp
    ref, aux = oracle.mining_hard_neg(B, cls, loc, tg, None, lt, negative_ratio=3., num_classes=C)
    got, det = hnm.mining_hard_neg(B, to_dev(cls.reshape(B * N, C), cuda), to_dev(loc, cuda), to_dev(tg, cuda), None,
                                   to_dev(lt, cuda), negative_ratio=3., num_classes=C, return_details=True)
    _check_hnm(ref, aux, got, det)
    with pytest.raises(ValueError):                         # train_sfd.py rule: no clamp -> position -1
        hnm.mining_hard_neg(B, to_dev(cls, cuda), to_dev(loc, cuda), to_dev(tg, cuda), None, to_dev(lt, cuda),
                            at_least_one=False)


def test_hard_negative_mining_across_batch_vs_oracle(cuda, oracle):
    from dan_b200.utility import hard_negative_mining as hnm
    # the last case is one row of 477 750 keys: slices beyond the shared-memory staging capacity (keys re-read from L2)
    for quantise, B, N in ((False, 3, 4000), (True, 3, 4000), (False, 14, 34125)):
        cls, loc, tg, lt = _hnm_inputs(80, B, N, 2, quantise)
        ref, aux = oracle.mining_hard_neg_across_batch(B, cls, loc, tg, None, lt)
        got, det = hnm.mining_hard_neg_across_batch(B, to_dev(cls, cuda), to_dev(loc, cuda), to_dev(tg, cuda), None,
                                                    to_dev(lt, cuda), return_details=True)
        _check_hnm(ref, aux, got, det)


def test_hard_negative_mining_after_encode_full_size(cuda, s3fd_anchors_np):
    """config 2 shape (batch 32 x 34 125 anchors): labels straight from encode_all_anchors; size-independent properties."""
    import torch
    from dan_b200.utility import anchor_manipulator as am, hard_negative_mining as hnm
    enc = am.AnchorEncoder(0.4, 0.4, PS)
    anchors = [to_dev(v, cuda) for v in s3fd_anchors_np]
    B, N = 32, s3fd_anchors_np[0].shape[0]
    cat, offs = synthetic.to_csr([synthetic.gen_faces(300 + i, 50) for i in range(B)])
    res = enc.encode_all_anchors(to_dev(cat, cuda), to_dev(offs, cuda), *anchors, match_mining=True)
    g = torch.Generator(device=cuda).manual_seed(3)
    cls = torch.randn((B, N, 2), device=cuda, generator=g) * 3.
    loc = torch.randn((B * N, 4), device=cuda, generator=g)
    (c, l, t, lt), det = hnm.mining_hard_neg(B, cls, loc, res.labels, res.scores, res.targets, return_details=True)
    labels = res.labels
    n_pos = (labels > 0).sum(-1)
    n_neg = (labels == 0).sum(-1)
    want = torch.clamp(torch.minimum(3 * n_pos, n_neg), min=1).to(torch.int32)
    assert torch.equal(det.n_neg_select, want)
    fm = det.final_mask.view(B, N).bool()
    assert torch.equal((fm & (labels == 0)).sum(-1).to(torch.int32), want)            # continuous logits: no ties
    assert bool(fm[labels > 0].all()) and not bool(fm[labels < 0].any())
    assert c.shape[0] == int(fm.sum()) and l.shape[0] == int(n_pos.sum())
    assert torch.equal(c, cls.view(-1, 2)[fm.view(-1)]) and torch.equal(t, labels.view(-1)[fm.view(-1)].clamp(0, 2))
    assert torch.equal(l, loc[labels.view(-1) > 0]) and torch.equal(lt, res.targets.view(-1, 4)[labels.view(-1) > 0])


# ------------------------------------------------------------------------------------------------
# DynamicAnchorRouting, evaluation branch
# ------------------------------------------------------------------------------------------------
def test_routing_goldens_from_reference_functor(cuda):
    """CUDA path vs outputs of the reference's own compiled functor (tests/golden/routing_reference.npz)."""
    from dan_b200.utility import custom_op
    from test_oracle import ROUTING_DAN640, ROUTING_SMALL
    g = np.load(os.path.join(GOLD, "routing_reference.npz"))
    for mode in ("plain", "ties"):
        c = ROUTING_SMALL
        a, t, lab, m = synthetic.gen_routing(7, c["feat_heights"], c["feat_widths"], c["depths"], c["strides"], mode)
        mo, do = custom_op.dynamic_anchor_routing_layers(to_dev(a, cuda), to_dev(t, cuda), to_dev(lab, cuda), to_dev(m, cuda),
                                                         c["feat_heights"], c["feat_widths"], c["depths"], c["strides"])
        np.testing.assert_array_equal(_np(mo).astype(np.int8), g["small_%s_mask" % mode])
        np.testing.assert_array_equal(_np(do).view(np.uint32), g["small_%s_decode" % mode].view(np.uint32))
    c = ROUTING_DAN640
    a, t, lab, m = synthetic.gen_routing(11, c["feat_heights"], c["feat_widths"], c["depths"], c["strides"], "plain")
    mo, do = custom_op.dynamic_anchor_routing_layers(to_dev(a, cuda), to_dev(t, cuda), to_dev(lab, cuda), to_dev(m, cuda),
                                                     c["feat_heights"], c["feat_widths"], c["depths"], c["strides"])
    np.testing.assert_array_equal(np.packbits(_np(mo).astype(np.uint8)), g["dan640_mask_bits"])
    np.testing.assert_array_equal(np.bitwise_xor.reduce(_np(do).view(np.uint32), axis=0), g["dan640_decode_xor"])


@pytest.mark.parametrize("mode", ["plain", "ties"])
def test_routing_batch_vs_oracle(cuda, oracle, mode):
    """batch of 3 images x 6 DAN layers in ONE call == the op applied per image and layer (eval_dan.py:386-393)."""
    from dan_b200.utility import custom_op
    from test_oracle import ROUTING_DAN640, routing_port
    c = ROUTING_DAN640
    ins, refs = [], []
    for i in range(3):
        x, mo, do = routing_port(c, 40 + i, mode)
        ins.append(x), refs.append((mo, do))
    stack = [np.stack([x[k] for x in ins]) for k in range(4)]
    mo, do = custom_op.dynamic_anchor_routing_layers(*[to_dev(v, cuda) for v in stack], c["feat_heights"], c["feat_widths"],
                                                     c["depths"], c["strides"])
    for i in range(3):
        np.testing.assert_array_equal(_np(mo[i]), refs[i][0])
        np.testing.assert_array_equal(_np(do[i]).view(np.uint32), refs[i][1].view(np.uint32))
    assert 0 < int(mo.sum()) < mo.numel()


def test_routing_single_layer_op_and_errors(cuda, oracle):
    from oracle import native
    from dan_b200 import _lib
    from dan_b200.utility import custom_op
    a, t, lab, m = synthetic.gen_routing(3, [12], [17], [3], [16], "ties")
    rm, rd = native.dynamic_anchor_routing(a, t, lab, m, 12, 17, 3, 16, 192, 272)
    gm, gd = custom_op.dynamic_anchor_routing(to_dev(a, cuda), to_dev(t, cuda), to_dev(lab, cuda), to_dev(m, cuda), 12, 17, 3, 16,
                                              192, 272, False, 0.03, 0.0)
    np.testing.assert_array_equal(_np(gm), rm)
    np.testing.assert_array_equal(_np(gd).view(np.uint32), rd.view(np.uint32))
    # extreme offsets: exp overflow / underflow follow libm
    t2 = t.copy()
    t2[::5, 2] = 100.
    t2[1::5, 3] = -120.
    rm, rd = native.dynamic_anchor_routing(a, t2, lab, m, 12, 17, 3, 16, 192, 272)
    gm, gd = custom_op.dynamic_anchor_routing(to_dev(a, cuda), to_dev(t2, cuda), to_dev(lab, cuda), to_dev(m, cuda), 12, 17, 3, 16,
                                              192, 272)
    np.testing.assert_array_equal(_np(gd).view(np.uint32), rd.view(np.uint32))
    with pytest.raises(_lib.DanError):                      # dynamic_anchor_routing.cc:527
        custom_op.dynamic_anchor_routing(to_dev(a, cuda), to_dev(t, cuda), to_dev(lab, cuda), to_dev(m, cuda), 12, 17, 3, 16, 192, 272,
                                         False, 1.0, 0.0)
    with pytest.raises(NotImplementedError):
        custom_op.dynamic_anchor_routing(to_dev(a, cuda), to_dev(t, cuda), to_dev(lab, cuda), to_dev(m, cuda), 12, 17, 3, 16, 192, 272,
                                         True, 0.03, 0.0)
    with pytest.raises(ValueError):                         # layer geometry does not match the number of anchors
        custom_op.dynamic_anchor_routing(to_dev(a, cuda), to_dev(t, cuda), to_dev(lab, cuda), to_dev(m, cuda), 12, 16, 3, 16, 192, 272)


# ------------------------------------------------------------------------------------------------
# detect_face top-k + bbox_vote
# ------------------------------------------------------------------------------------------------
def test_bbox_vote_goldens_from_reference_function(cuda):
    """CUDA bbox_vote vs outputs of the reference's own function (tests/golden/vote_reference.npz)."""
    from dan_b200.utility import eval_merge
    from test_oracle import VOTE_CASES
    g = np.load(os.path.join(GOLD, "vote_reference.npz"))
    for image_index, n, faces, bg in VOTE_CASES:
        det = synthetic.gen_vote_dets(image_index, n, faces, bg)
        got = _np(eval_merge.bbox_vote(to_dev(det, cuda)))
        ref = g["vote_%d" % image_index]
        assert got.shape == ref.shape, (image_index, got.shape, ref.shape)
        np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))
    got = _np(eval_merge.bbox_vote(to_dev(synthetic.gen_vote_dets(3, 6750, 150, 0.6), cuda), nms_threshold=0.5, max_per_image=100))
    np.testing.assert_array_equal(got.view(np.uint32), g["vote_3_top100_thr05"].view(np.uint32))


def test_bbox_vote_batch_vs_oracle(cuda, oracle):
    """ragged batch in one launch; also checks the sort order and the head assignment themselves."""
    from dan_b200 import functional as F
    sizes = [0, 1, 2, 17, 700, 3000, 8192]
    cap = max(sizes)
    dets = [synthetic.gen_vote_dets(60 + i, n, max(n // 25, 1), 0.5)[:n] if n else np.zeros((0, 5), np.float32)
            for i, n in enumerate(sizes)]
    batch = np.zeros((len(sizes), cap, 5), np.float32)
    batch[:] = np.nan                                         # rows beyond counts[b] must never be read into a result
    for i, d in enumerate(dets):
        batch[i, :d.shape[0]] = d
    counts = np.array([d.shape[0] for d in dets], np.int32)
    out, cnt, order, assign = F.bbox_vote_batch(to_dev(batch, cuda), to_dev(counts, cuda), 0.3, 750, details=True)
    for i, d in enumerate(dets):
        ref, aux = oracle.bbox_vote(d, return_details=True)
        k = int(cnt[i])
        assert k == ref.shape[0], (i, k, ref.shape)
        np.testing.assert_array_equal(_np(out[i, :k]).view(np.uint32), ref.view(np.uint32))
        assert not _np(out[i, k:]).any()
        n = d.shape[0]
        np.testing.assert_array_equal(_np(order[i, :n]), aux["order"])
        np.testing.assert_array_equal(_np(assign[i, :n]), aux["assign"])


def test_detect_face_select_vs_oracle(cuda, oracle):
    from dan_b200.utility import eval_merge
    rng = np.random.default_rng(12)
    for n, shrink in ((3000, 0.5), (20000, 1.75), (5, 1.0), (1, 2.0), (900, 3.0)):
        b = rng.uniform(0, 640, (n, 4)).astype(np.float32)
        s = rng.uniform(0, 1, n).astype(np.float32)
        s[::10] = s[min(5, n - 1)]                                        # equal scores: higher index first
        ref, _ = oracle.detect_face_select(b, s, shrink)
        got = _np(eval_merge.detect_face_select(to_dev(b, cuda), to_dev(s, cuda), shrink))
        assert got.shape == ref.shape
        np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))


# ------------------------------------------------------------------------------------------------
# input hand-off
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch,use_mirror", [(23, True), (1, False), (1500, True), (6, False)])
def test_input_handoff_vs_oracle(cuda, oracle, batch, use_mirror):
    from dan_b200.utility import input_handoff
    from test_oracle import handoff_case
    gts, hw, mirror = handoff_case(batch, batch)
    if batch == 6:
        gts = [g[:0] for g in gts]                                # nothing at all survives
    cat, offs = synthetic.to_csr(gts)
    ref = oracle.prepare_gt_batch(gts, hw, (640, 640), mirror if use_mirror else None)
    got = input_handoff.prepare_gt_batch(to_dev(cat.reshape(-1, 4), cuda), to_dev(offs, cuda), to_dev(hw, cuda), (640, 640),
                                         to_dev(mirror, cuda) if use_mirror else None)
    np.testing.assert_array_equal(_np(got[0]).view(np.uint32), ref[0].view(np.uint32))
    np.testing.assert_array_equal(_np(got[1]), ref[1])
    np.testing.assert_array_equal(_np(got[2]), ref[2])


def test_input_handoff_feeds_encode(cuda, oracle, s3fd_anchors_np):
    """hand-off -> encode_all_anchors on the device == oracle hand-off -> oracle encode, image by image."""
    from dan_b200.utility import anchor_manipulator as am, input_handoff
    from test_oracle import handoff_case
    gts, hw, mirror = handoff_case(77, 12)
    cat, offs = synthetic.to_csr(gts)
    g_boxes, g_offs, g_idx = input_handoff.prepare_gt_batch(to_dev(cat.reshape(-1, 4), cuda), to_dev(offs, cuda), to_dev(hw, cuda),
                                                            (640, 640), to_dev(mirror, cuda))
    r_boxes, r_offs, r_idx = oracle.prepare_gt_batch(gts, hw, (640, 640), mirror)
    anchors = [to_dev(v, cuda) for v in s3fd_anchors_np]
    res = am.AnchorEncoder(0.4, 0.4, PS).encode_all_anchors(g_boxes, g_offs, *anchors, match_mining=True)
    e_ref = oracle.AnchorEncoder(0.4, 0.4, PS)
    assert res.labels.shape[0] == len(r_idx)
    for k in range(len(r_idx)):
        ref = e_ref.encode_anchors(r_boxes[r_offs[k]:r_offs[k + 1]], *s3fd_anchors_np, match_mining=True)
        _check_encode(ref, [res.targets[k], res.labels[k], res.scores[k], res.matched_gt[k]], "kept image %d" % k)


# ------------------------------------------------------------------------------------------------
# round 2: stress paths of the fused encode, full-size keep-lists, the sharded batch, numerics edges
# ------------------------------------------------------------------------------------------------
def _fused_mining_vs_matcher(cuda, oracle, anchors_np, gts, pos, ign, min_match, stop):
    """dan_encode_batch with arbitrary mining attributes against the matcher run on the materialised overlaps."""
    from dan_b200 import functional as F
    anchors = [to_dev(v, cuda) for v in anchors_np]
    cat, offs = synthetic.to_csr(gts)
    params = F.encode_params(pos, ign, PS, match_mining=True, min_match=min_match, stop_positive_thres=stop)
    res = F.encode_batch(params, *anchors[:4], anchors[4], to_dev(cat, cuda), to_dev(offs, cuda), want_match=True)
    a4 = np.stack(anchors_np[:4], -1)
    n_comp = 0
    for b, gt in enumerate(gts):
        g = gt if len(gt) else np.array([[0., 0., 1., 1.]], np.float32)
        ov = oracle.iou_matrix(a4, g) * anchors_np[4].astype(np.float32)[:, None]
        rm, rs = oracle.small_mining_match(ov, 0., ign, pos, min_match, stop, impl="reference" if _have_ref() else "port")
        np.testing.assert_array_equal(_np(res.match[b]), rm, err_msg="image %d match" % b)
        np.testing.assert_array_equal(_np(res.scores[b]), rs, err_msg="image %d scores" % b)
        np.testing.assert_array_equal(_np(res.labels[b]), (rm > -1).astype(np.int64) - (rm < -1).astype(np.int64))
        s1 = ov.max(1)                                     # anchors that stage 3 turned positive
        n_comp += int(((rm >= 0) & (s1 < pos)).sum())
    return n_comp


def _have_ref():
    from oracle import native
    return native.have_reference()


def test_fused_encode_bucket_overflow_and_large_min_match(cuda, oracle, s3fd_anchors_np):
    """stop_positive_thres 0.01 puts hundreds of candidates into every GT's bucket (> kBucketCap = 64: the rescan path
    of pass 3), min_match 40 / 3000 makes every GT needy and lets neighbouring GTs compete for the same anchors (the
    windows and rounds of pass 3, its selection budget); the reference functor is the judge."""
    gts = [synthetic.gen_faces(70 + i, 12, snap=(4.0 if i == 1 else 0.0)) for i in range(3)]
    gts.append(synthetic.gen_adversarial("duplicate"))
    gts.append(np.concatenate([synthetic.gen_faces(75, 6, smin=60.0, smax=90.0)] * 2) + np.float32(3.0))     # overlapping pairs
    for min_match, stop in ((40, 0.01), (3000, 0.2), (6, 0.05)):
        n = _fused_mining_vs_matcher(cuda, oracle, s3fd_anchors_np, gts, 0.4, 0.4, min_match, stop)
        assert n > 0


@pytest.mark.parametrize("size", [(640, 640), (1024, 1024)])
def test_fused_encode_layout_hint_is_neutral(cuda, oracle, size):
    """dan_encode_params' layout hint (8 x 4 tiles of cells per warp instead of 32 x 1 strips) changes which warp holds
    which anchor and nothing else: every output of the fused encode is bit-identical with and without it - mining
    (with stage-2 patches and stage-3 compensation), dual matcher with both claim orders, encode_pa_anchors - and the
    mining result equals the reference functor's.  Invalid hints are refused."""
    import torch
    from dan_b200 import functional as F, _lib as L
    from dan_b200.utility import anchor_manipulator as am
    cfg = synthetic.pyramid_config("s3fd", size)
    enc = am.AnchorEncoder(0.4, 0.4, PS)
    anchors = synthetic.build_anchors(enc, cfg)
    grids = F.grid_hint(enc.pyramid)
    assert len(grids) >= 3 and grids[0][0] == 0
    gts = [synthetic.gen_faces(40 + i, 50) for i in range(4)] + [synthetic.gen_dense_tiny(5, lo=150, hi=200, snap=2.0),
                                                                  np.zeros((0, 4), np.float32), synthetic.gen_adversarial("duplicate")]
    scale = np.float32(size[0] / 640.0)
    gts = [g * scale for g in gts]
    cat, offs = synthetic.to_csr(gts)
    cat_d, offs_d = to_dev(cat, cuda), to_dev(offs, cuda)
    variants = [dict(match_mining=True), dict(match_mining=True, min_match=40, stop_positive_thres=0.05),
                dict(match_mining=False, gt_max_first=True), dict(match_mining=False, gt_max_first=False, ignore_between=False),
                dict(match_mining=True, pa_scale=0.5)]
    for kw in variants:
        plain = F.encode_batch(F.encode_params(0.4, 0.4, PS, **kw), *anchors[:4], anchors[4], cat_d, offs_d, want_match=True)
        p_hint = F.encode_params(0.4, 0.4, PS, pyramid=enc.pyramid, **kw)
        assert p_hint.num_grids == len(grids)
        hinted = F.encode_batch(p_hint, *anchors[:4], anchors[4], cat_d, offs_d, want_match=True)
        for name in ("targets", "labels", "scores", "matched_gt", "match"):
            a, b = getattr(plain, name), getattr(hinted, name)
            if a.dtype == torch.float32:
                a, b = a.view(torch.int32), b.view(torch.int32)              # bit patterns (sign of zero included)
            assert torch.equal(a, b), "%s differs with the layout hint (%s)" % (name, kw)
    # anchors outside the image masked out (border 0): the mask is read through the same warp -> anchor table
    enc0 = am.AnchorEncoder(0.4, 0.4, PS)
    anchors0 = synthetic.build_anchors(enc0, synthetic.pyramid_config("s3fd", size, border=0.))
    assert 0 < int(anchors0[4].sum()) < anchors0[4].numel()
    for kw in (dict(match_mining=True), dict(match_mining=False)):
        plain = F.encode_batch(F.encode_params(0.4, 0.4, PS, **kw), *anchors0[:4], anchors0[4], cat_d, offs_d, want_match=True)
        hinted = F.encode_batch(F.encode_params(0.4, 0.4, PS, pyramid=enc0.pyramid, **kw), *anchors0[:4], anchors0[4], cat_d, offs_d,
                                want_match=True)
        for name in ("targets", "labels", "scores", "matched_gt", "match"):
            a, b = getattr(plain, name), getattr(hinted, name)
            if a.dtype == torch.float32:
                a, b = a.view(torch.int32), b.view(torch.int32)
            assert torch.equal(a, b), "%s differs with the layout hint and an inside mask (%s)" % (name, kw)
    # the mirror uses the hint for the anchors it generated itself: against the reference functor
    a_np = [_np(a) for a in anchors]
    res = enc.encode_anchors_batch(cat_d, offs_d, *anchors[:4], anchors[4], match_mining=True, want_match=True)
    a4 = np.stack(a_np[:4], -1)
    for b in (0, 4, 5):
        g = gts[b] if len(gts[b]) else np.array([[0., 0., 1., 1.]], np.float32)
        ov = oracle.iou_matrix(a4, g) * a_np[4].astype(np.float32)[:, None]
        rm, rs = oracle.small_mining_match(ov, 0., 0.4, 0.4, 6, 0.3, impl="reference" if _have_ref() else "port")
        np.testing.assert_array_equal(_np(res.match[b]), rm)
        np.testing.assert_array_equal(_np(res.scores[b]), rs)
    # invalid hints
    for start, w, h in ((16, 8, 4), (0, 12, 4), (0, 8, 6), (0, 8, 4 * (anchors[0].numel() // 32 + 1))):
        bad = F.encode_params(0.4, 0.4, PS, match_mining=True)
        bad.num_grids, bad.grid_start[0], bad.grid_w[0], bad.grid_h[0] = 1, start, w, h
        with pytest.raises(L.DanError):
            F.encode_batch(bad, *anchors[:4], anchors[4], cat_d, offs_d)
    bad = F.encode_params(0.4, 0.4, PS, match_mining=True, pyramid=enc.pyramid)
    bad.grid_start[1] = 0                                                  # overlaps grid 0
    with pytest.raises(L.DanError):
        F.encode_batch(bad, *anchors[:4], anchors[4], cat_d, offs_d)


def test_fused_encode_many_needy_gts(cuda, oracle, s3fd_anchors_np):
    """more needy GTs than one window of pass 3 holds (64) and more than one scan range (1024), dense tiny faces that
    share candidate anchors: the serial dependence between the windows."""
    gts = [synthetic.gen_dense_tiny(3, lo=1100, hi=1200), synthetic.gen_dense_tiny(4, lo=150, hi=200, snap=2.0)]
    n = _fused_mining_vs_matcher(cuda, oracle, s3fd_anchors_np, gts, 0.4, 0.4, 6, 0.3)
    assert n > 100


@pytest.mark.parametrize("size,faces", [((1280, 1280), 200), ((1600, 1600), 300)])
def test_parse_by_class_keep_lists_full_size(cuda, oracle, size, faces):
    """config 4 at 1280^2 and 1600^2 (213 294 anchors): the keep-list itself against the oracle, one image each."""
    from dan_b200.utility import bbox_util as bu
    cfg = synthetic.pyramid_config("s3fd", size, border=0.)
    a_np = synthetic.build_anchors(oracle.AnchorEncoder(None, None, PS), cfg)
    an = np.stack(a_np[:4], -1)
    anchors = [to_dev(v, cuda) for v in a_np[:4]]
    cls, loc, _ = synthetic.gen_predictions(500 + size[0], an, size=size, max_faces=faces)
    det = bu.parse_by_class_batch(list(size), to_dev(cls[None], cuda), 2, 0.01, 0, 5000, 750, 0.3, loc_pred=to_dev(loc[None], cuda),
                                  anchors=anchors)
    boxes, (sb, ss, idx) = _oracle_parse(oracle, size, cls, loc, a_np)
    np.testing.assert_array_equal(_np(det.scores[0, 0]), ss[1])
    np.testing.assert_array_equal(_np(det.boxes[0, 0]), sb[1])
    top_idx, keep = idx[1]
    k = int((ss[1] > 0).sum())
    assert int(det.counts[0, 0]) == k and k > 50
    np.testing.assert_array_equal(_np(det.keep_pos[0, 0])[:len(keep)], keep)
    np.testing.assert_array_equal(_np(det.anchor_index[0, 0])[:k], top_idx[keep[:k]])


def test_sharded_batch_256_matches_oracle(cuda, oracle, s3fd_anchors_np):
    """config 5: B = 256 images through shard_range -> HotPath.step -> the detection exchange.  On one GPU the 8 ranks run
    one after the other and their slabs are concatenated the way the all-gather lays them out (on a multi-GPU box the
    driver's scaling run and bench.py's checksum test cover the NCCL path; tests/test_abi_and_host.py the gloo one);
    sampled images of every shard are compared with the oracle."""
    import torch
    from dan_b200 import functional as F, pipeline
    from dan_b200.utility import anchor_manipulator as am
    B, world = 256, 8
    enc = am.AnchorEncoder(0.4, 0.4, PS)
    a_train = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd"))
    a_eval = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd", border=0.))
    ev_np = synthetic.build_anchors(oracle.AnchorEncoder(None, None, PS), synthetic.pyramid_config("s3fd", border=0.))
    an = np.stack(ev_np[:4], -1)
    sample = {r: [pipeline.shard_range(B, r, world).lo, pipeline.shard_range(B, r, world).hi - 1] for r in range(world)}
    gts = {i: synthetic.gen_faces(900 + i, 50) for i in range(B)}
    enc_params = F.encode_params(0.4, 0.4, PS, match_mining=True)
    pp_params = F.postprocess_params(2, (640, 640), 0.01, 0, 5000, 750, 0.3, prior_scaling=PS)
    one = pipeline.DeviceGather(0, 1, cuda)              # world size 1: the C-ABI collective degenerates to a copy
    slabs = []
    e_ref = oracle.AnchorEncoder(0.4, 0.4, PS)
    for r in range(world):
        sh = pipeline.shard_range(B, r, world)
        assert sh.hi - sh.lo == 32
        preds = {i: synthetic.gen_predictions(900 + i, an, max_faces=40) for i in sample[r]}
        cls = np.zeros((32, an.shape[0], 2), np.float32)
        cls[:, :, 0] = 10.0                                   # images that are not sampled: nothing passes the threshold
        loc = np.zeros((32, an.shape[0], 4), np.float32)
        for i in sample[r]:
            cls[i - sh.lo], loc[i - sh.lo] = preds[i][0], preds[i][1]
        cat, offs = synthetic.to_csr([gts[i] for i in range(sh.lo, sh.hi)])
        hp = pipeline.HotPath(a_train[:4], a_train[4], enc_params, pp_params, anchors_eval=a_eval[:4], device_gather=one)
        res, det = hp.step(to_dev(cat, cuda), to_dev(offs, cuda), to_dev(cls, cuda), to_dev(loc, cuda))
        torch.cuda.synchronize()
        (counts, scores, boxes), = hp.gathered()
        assert torch.equal(hp._recv, hp._slab.buf)
        slabs.append((counts.clone(), scores.clone(), boxes.clone()))
        for i in sample[r]:
            ref = e_ref.encode_anchors(gts[i], *s3fd_anchors_np, match_mining=True)
            _check_encode(ref, [res.targets[i - sh.lo], res.labels[i - sh.lo], res.scores[i - sh.lo], res.matched_gt[i - sh.lo]],
                          "image %d" % i)
            _, (sb, ss, _) = _oracle_parse(oracle, (640, 640), preds[i][0], preds[i][1], ev_np)
            np.testing.assert_array_equal(_np(scores[i - sh.lo, 0]), ss[1], err_msg="image %d" % i)
            np.testing.assert_array_equal(_np(boxes[i - sh.lo, 0]), sb[1], err_msg="image %d" % i)
    flat = pipeline.flatten_detections(slabs)
    assert len(flat) == B
    for r in range(world):
        for i in range(B // world):
            n = int(slabs[r][0][i, 0])
            assert flat[r * 32 + i][1][0].shape[0] == n
            assert (n > 0) == ((r * 32 + i) in sample[r])
    one.close()


def test_softmax_logit_gaps_beyond_exp_range(cuda, oracle):
    """logit gaps above ~87.7: exp(x - max) flushes to exactly 0 (2^n is built in one step, like Eigen's pexp), so the
    probability is 0, not a denormal; the filter's quick reject must agree with the exact path there."""
    from dan_b200 import functional as F
    from dan_b200.utility import bbox_util as bu
    gaps = np.array([86.0, 87.0, 87.3, 87.5, 87.7, 88.0, 88.4, 90.0, 120.0, 1e4, 3e5], np.float32)
    logits = np.stack([np.concatenate([gaps, np.zeros_like(gaps)]), np.concatenate([np.zeros_like(gaps), gaps])], 1)
    ref = oracle.softmax(logits)
    np.testing.assert_array_equal(_np(F.softmax(to_dev(logits, cuda))), ref)
    assert ref.min() == 0.0
    n = logits.shape[0]
    boxes = np.tile(np.array([[10., 10., 50., 50.]], np.float32), (n, 1)) + np.arange(n, dtype=np.float32)[:, None] * 60
    sb, ss = bu.parse_by_class([4000, 4000], to_dev(logits, cuda), to_dev(boxes, cuda), 2, 0.0, 0, 100, 100, 0.5)
    rb, rs = oracle.parse_by_class([4000, 4000], logits, boxes, 2, 0.0, 0, 100, 100, 0.5)
    np.testing.assert_array_equal(_np(ss[1]), rs[1])
    np.testing.assert_array_equal(_np(sb[1]), rb[1])


def test_device_gather_single_rank(cuda):
    """dan_gather_detections through the C ABI with a one-rank communicator (ncclAllGather of one slab), eagerly and
    captured in a CUDA graph."""
    import torch
    from dan_b200 import pipeline
    g = pipeline.DeviceGather(0, 1, cuda)
    send = torch.arange(4096, dtype=torch.float32, device=cuda)
    recv = torch.zeros_like(send)
    g.gather(send, recv)
    torch.cuda.synchronize()
    assert torch.equal(send, recv)
    recv.zero_()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            g.gather(send, recv)
    torch.cuda.current_stream().wait_stream(side)
    send.mul_(2)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(send, recv)
    del graph
    with pytest.raises(ValueError):
        g.gather(send, recv[:100])
    g.close()


def test_parse_by_class_beyond_shared_memory_limits(cuda, oracle):
    """keep_topk and nms_topk far above what the kernels hold in shared memory (tf.nn.top_k and
    tf.image.non_max_suppression have no limit): candidates, kept list and stripe index live in the workspace."""
    from dan_b200.utility import bbox_util as bu
    rng = np.random.default_rng(77)
    n = 20000
    cls = rng.normal(0, 2.0, (n, 2)).astype(np.float32)
    c = rng.uniform(0, 2000, (n, 2))
    wh = rng.uniform(4, 40, (n, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    for keep_topk, nms_topk in ((12000, 9500), (9000, 400)):
        rb, rs = oracle.parse_by_class([2000, 2000], cls, boxes, 2, 0.05, 0, keep_topk, nms_topk, 0.5)
        sb, ss = bu.parse_by_class([2000, 2000], to_dev(cls, cuda), to_dev(boxes, cuda), 2, 0.05, 0, keep_topk, nms_topk, 0.5)
        np.testing.assert_array_equal(_np(ss[1]), rs[1])
        np.testing.assert_array_equal(_np(sb[1]), rb[1])
        assert int((rs[1] > 0).sum()) > 0.8 * min(nms_topk, 9000)


def test_peer_exchange_single_rank(cuda, oracle):
    """dan_postprocess_batch_peers with one destination (this rank's own receive buffer, allocated through the C ABI):
    the NMS kernel's peer stores, the arrival flag of the last CTA and dan_wait_detections, eagerly and in a CUDA graph."""
    import torch
    from dan_b200 import functional as F, pipeline
    from dan_b200.utility import anchor_manipulator as am
    enc = am.AnchorEncoder(0.4, 0.4, PS)
    a_train = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd"))
    a_eval = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd", border=0.))
    ev_np = synthetic.build_anchors(oracle.AnchorEncoder(None, None, PS), synthetic.pyramid_config("s3fd", border=0.))
    an = np.stack(ev_np[:4], -1)
    B = 3
    gts = [synthetic.gen_faces(40 + i, 20) for i in range(B)]
    preds = [synthetic.gen_predictions(40 + i, an, max_faces=30) for i in range(B)]
    cat, offs = synthetic.to_csr(gts)
    cls = to_dev(np.stack([p[0] for p in preds]), cuda)
    loc = to_dev(np.stack([p[1] for p in preds]), cuda)
    pp_params = F.postprocess_params(2, (640, 640), 0.01, 0, 5000, 750, 0.3, prior_scaling=PS)
    words = pipeline.DetectionSlab.words_for(B, 1, 750)
    px = pipeline.PeerExchange(0, 1, cuda, words, num_sets=2)
    hp = pipeline.HotPath(a_train[:4], a_train[4], F.encode_params(0.4, 0.4, PS, match_mining=True), pp_params,
                          anchors_eval=a_eval[:4], peer_exchange=px, peer_set=1)
    gt_d, offs_d = to_dev(cat, cuda), to_dev(offs, cuda)
    hp.step(gt_d, offs_d, cls, loc)
    torch.cuda.synchronize()
    assert torch.equal(px.recv(1), hp._slab.buf) and int(px.flags(1)[0]) == 1 and not px.recv(0).any()
    (counts, scores, boxes), = hp.gathered()
    for i in range(B):
        _, (sb, ss, _) = _oracle_parse(oracle, (640, 640), preds[i][0], preds[i][1], ev_np)
        np.testing.assert_array_equal(_np(scores[i, 0]), ss[1])
        np.testing.assert_array_equal(_np(boxes[i, 0]), sb[1])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            hp.step(gt_d, offs_d, cls, loc)
    torch.cuda.current_stream().wait_stream(side)
    px.recv(1).zero_()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(px.recv(1), hp._slab.buf) and int(px.flags(1)[0]) == 3
    del graph
    # the same step with the box offsets left in PINNED HOST memory (host-resident geometry through HotPath.step, with
    # the peer stores): identical slab and identical encode outputs
    ref_det = [t.clone() for t in hp._slab.views()]
    ref_enc = [t.clone() for t in hp._enc_out[:4]]
    for t in hp._slab.views():
        t.fill_(-3)
    loc_host = torch.from_numpy(np.stack([p[1] for p in preds])).pin_memory()
    hp.step(gt_d, offs_d, cls, loc_host)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(ref_det, hp._slab.views()))
    assert all(torch.equal(a, b) for a, b in zip(ref_det, hp.gathered()[0]))
    assert all(torch.equal(a, b) for a, b in zip(ref_enc, hp._enc_out[:4]))
    px.close()
