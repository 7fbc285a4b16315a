"""Generates tests/golden/*.npz.  Run HERE (the build container), where /root/reference
exists: the mining matcher outputs come from the reference's OWN functor compiled verbatim
(oracle/_ref/libsmm_ref.so, `make -C oracle`), everything else from the numpy restatement
oracle/reference_np.py.  The fixtures pin (a) the restatement against the real reference
and (b) the CUDA kernels on the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dan_b200 import synthetic  # noqa: E402
from oracle import native, reference_np as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def pack_encode(res):
    targets, labels, scores, matched, match = res
    nz = np.nonzero(labels != 0)[0].astype(np.int32)
    pos = np.nonzero(labels == 1)[0].astype(np.int32)
    return dict(nonzero_rows=nz, nonzero_labels=labels[nz].astype(np.int8), pos_rows=pos,
                pos_match=match[pos].astype(np.int32), pos_targets=targets[pos], pos_matched=matched[pos],
                scores=scores, match=match.astype(np.int32))


ROUTING_SMALL = dict(feat_heights=[40, 20, 10, 5], feat_widths=[40, 20, 10, 5], depths=[1, 1, 2, 3], strides=[4, 8, 16, 32])
ROUTING_DAN640 = dict(feat_heights=[160, 80, 40, 20, 10, 5], feat_widths=[160, 80, 40, 20, 10, 5], depths=[1] * 6,
                      strides=[4, 8, 16, 32, 64, 128])           # eval_dan.py:386 at 640 x 640


def routing_reference(cfg, image_index, mode):
    """The reference's own DynamicAnchorRouting functor (oracle/_ref/libdar_ref.so), evaluation branch, layer by layer
    like eval_dan.py:387-393."""
    a, t, lab, m = synthetic.gen_routing(image_index, cfg["feat_heights"], cfg["feat_widths"], cfg["depths"], cfg["strides"], mode)
    masks, decs, off = [], [], 0
    for H, W, D, S in zip(cfg["feat_heights"], cfg["feat_widths"], cfg["depths"], cfg["strides"]):
        n = H * W * D
        mo, do = native.dynamic_anchor_routing(a[off:off + n], t[off:off + n], lab[off:off + n], m[off:off + n], H, W, D, S,
                                               640, 640, False, 0.03, 0.0, impl="reference")
        masks.append(mo), decs.append(do)
        off += n
    return np.concatenate(masks), np.concatenate(decs)


def make_routing():
    assert native.have_reference_dar(), "oracle/_ref/libdar_ref.so missing: run `make -C oracle` where /root/reference exists"
    out = {}
    for name, mode in (("plain", "plain"), ("ties", "ties")):
        mo, do = routing_reference(ROUTING_SMALL, 7, mode)
        out["small_%s_mask" % name] = mo.astype(np.int8)
        out["small_%s_decode" % name] = do
    mo, do = routing_reference(ROUTING_DAN640, 11, "plain")
    out["dan640_mask_bits"] = np.packbits(mo.astype(np.uint8))
    out["dan640_decode_xor"] = np.bitwise_xor.reduce(do.view(np.uint32), axis=0)
    out["dan640_decode_sum"] = do.view(np.uint32).astype(np.uint64).sum(axis=0)
    np.savez_compressed(os.path.join(OUT, "routing_reference.npz"), **out)


def reference_bbox_vote(nms_threshold=0.3, max_per_image=750):
    """The reference's OWN bbox_vote (eval_sfd.py:170-210): the function's source is parsed out of the script with `ast`
    and executed here (the script itself imports TensorFlow and cannot be imported); FLAGS as in eval_sfd.py:62-66."""
    import ast
    import warnings
    warnings.simplefilter("ignore")                   # np.row_stack is deprecated in numpy 2
    tree = ast.parse(open("/root/reference/eval_sfd.py").read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "bbox_vote"][0]

    class FLAGS(object):
        pass
    FLAGS.nms_threshold, FLAGS.max_per_image = nms_threshold, max_per_image
    ns = {"np": np, "FLAGS": FLAGS}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "eval_sfd.py", "exec"), ns)
    return ns["bbox_vote"]


VOTE_CASES = [(1, 300, 12, 0.3), (2, 2500, 60, 0.5), (3, 6750, 150, 0.6), (4, 1, 1, 0.0), (5, 40, 1, 0.0), (6, 8000, 1500, 0.05)]


def make_vote():
    vote = reference_bbox_vote()
    out = {}
    for image_index, n, faces, bg in VOTE_CASES:
        det = synthetic.gen_vote_dets(image_index, n, faces, bg)
        out["vote_%d" % image_index] = vote(det.copy())
    # other FLAGS: the cut after max_per_image groups, a stricter threshold
    out["vote_3_top100_thr05"] = reference_bbox_vote(0.5, 100)(synthetic.gen_vote_dets(3, 6750, 150, 0.6))
    np.savez_compressed(os.path.join(OUT, "vote_reference.npz"), **out)


def main():
    native.build()
    if len(sys.argv) > 1 and sys.argv[1] == "routing":
        make_routing()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "vote":
        make_vote()
        return
    assert native.have_reference(), "oracle/_ref/libsmm_ref.so missing: run `make -C oracle` where /root/reference exists"

    # ---- known-answer vectors for the custom op (cpp/ExtraLib/test_op.py:41,56) -------------
    ov = np.array([[0.1, 0.4, 0.6, 0.2, 0.7], [0.5, 0.14, 0.76, 0.32, 0.47], [0.21, 0.94, 0.66, 0.22, 0.57],
                   [0.91, 0.14, 0.26, 0.42, 0.67], [0.11, 0.84, 0.26, 0.42, 0.57]], np.float32)
    m, s = native.small_mining_match(ov, 0., 0.6, 0.6, 5, 0.1, impl="reference")
    tie = np.array([[.4]] * 7 + [[.45]], np.float32)
    mt, st = native.small_mining_match(tie, 0., .5, .5, 3, .3, impl="reference")
    np.savez(os.path.join(OUT, "smm_known_answers.npz"), test_op_overlaps=ov, test_op_match=m, test_op_scores=s,
             test_op_attrs=np.array([0., 0.6, 0.6, 5, 0.1]), tie_overlaps=tie, tie_match=mt, tie_scores=st,
             tie_attrs=np.array([0., .5, .5, 3, .3]))

    # ---- random dense matrices through the reference functor --------------------------------
    rng = np.random.default_rng(synthetic.BASE_SEED)
    dense = {}
    for i, (n, mm, levels) in enumerate([(64, 5, 0), (300, 17, 12), (1000, 40, 30), (33, 70, 8)]):
        x = rng.uniform(0, 1, (n, mm)).astype(np.float32)
        x[x < 0.55] = 0
        if levels:
            x = (np.round(x * levels) / levels).astype(np.float32)    # heavy exact ties
        for attrs in [(0., 0.4, 0.4, 6, 0.3), (0., 0.5, 0.7, 3, 0.1)]:
            mi, sc = native.small_mining_match(x, *attrs[:3], int(attrs[3]), attrs[4], impl="reference")
            key = "d%d_%d" % (i, int(attrs[3]))
            dense[key + "_x"] = x
            dense[key + "_attrs"] = np.array(attrs)
            dense[key + "_match"] = mi
            dense[key + "_scores"] = sc
    np.savez_compressed(os.path.join(OUT, "smm_dense_reference.npz"), **dense)

    # ---- anchors: closed-form facts of SURVEY 8(c)(3) ----------------------------------------
    enc = R.AnchorEncoder(0.4, 0.4, [0.1, 0.1, 0.2, 0.2])
    facts = {}
    for kind, size, border in [("s3fd", (640, 640), None), ("dan", (640, 640), None), ("s3fd", (1600, 1600), 0.),
                               ("s3fd", (640, 640), 0.)]:
        a = synthetic.build_anchors(enc, synthetic.pyramid_config(kind, size, border=border))
        key = "%s_%d_%s" % (kind, size[0], "train" if border is None else "eval")
        facts[key + "_n"] = np.int64(a[0].shape[0])
        facts[key + "_first"] = np.array([v[0] for v in a[:4]], np.float32)
        facts[key + "_last"] = np.array([v[-1] for v in a[:4]], np.float32)
        facts[key + "_inside"] = np.int64(a[4].sum())
        facts[key + "_sum"] = np.array([np.float64(v.astype(np.float64).sum()) for v in a[:4]])
    np.savez(os.path.join(OUT, "anchor_facts.npz"), **facts)

    # ---- encode goldens: BASELINE config 1 (S3FD 640, batch 1) + DAN + adversarial -----------
    s3fd = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd"))
    dan = synthetic.build_anchors(enc, synthetic.pyramid_config("dan"))
    cases = {}

    def add(name, anchors, gt, pos, ign, mining, pa=None):
        e = R.AnchorEncoder(pos, ign, [0.1, 0.1, 0.2, 0.2])
        if pa is None:
            res = e.encode_anchors(gt, *anchors, match_mining=mining, mining_impl="reference", return_match=True)
        else:
            res = e.encode_pa_anchors(gt, *anchors, ign, pos, match_mining=mining, scale=pa, mining_impl="reference",
                                      return_match=True)
        cases[name + "__gt"] = gt
        cases[name + "__cfg"] = np.array([pos, ign, 1.0 if mining else 0.0, -1.0 if pa is None else pa])
        for k, v in pack_encode(res).items():
            cases[name + "__" + k] = v

    add("s3fd_mining_m12", s3fd, synthetic.gen_faces(3, 12, min_faces=12), 0.4, 0.4, True)
    add("s3fd_mining_m50", s3fd, synthetic.gen_faces(7, 50, min_faces=50), 0.4, 0.4, True)
    add("s3fd_mining_snap", s3fd, synthetic.gen_faces(11, 30, snap=4.0, min_faces=30), 0.4, 0.4, True)
    add("s3fd_dual_m12_unittest", s3fd, synthetic.gen_faces(3, 12, min_faces=12), 0.5, 0.5, False)
    add("s3fd_mining_empty", s3fd, synthetic.gen_adversarial("empty"), 0.4, 0.4, True)
    add("dan_dual_m50", dan, synthetic.gen_faces(5, 50, min_faces=50), 0.35, 0.35, False)
    add("dan_dual_snap", dan, synthetic.gen_faces(13, 30, snap=1.0, min_faces=30), 0.35, 0.35, False)
    add("pb_head_dual_scale2", s3fd, synthetic.gen_faces(17, 20, min_faces=20), 0.35, 0.35, False, pa=2.0)
    np.savez_compressed(os.path.join(OUT, "encode_goldens.npz"), **cases)

    # ---- decode round trip of SURVEY 8(c)(4) ----------------------------------------------------
    e = R.AnchorEncoder(0.5, 0.5, [0.1, 0.1, 0.2, 0.2])
    a0 = [v[:1] for v in s3fd[:4]]
    gt = np.array([[3, 2, 20, 17]], np.float32)
    acy, acx, ah, aw = e.point2center(*a0)
    gcy, gcx, gh, gw = e.point2center(*[gt[:, i] for i in range(4)])
    t = np.stack([(gcy - acy) / ah / np.float32(.1), (gcx - acx) / aw / np.float32(.1),
                  R.logf(gh / ah) / np.float32(.2), R.logf(gw / aw) / np.float32(.2)], -1).astype(np.float32)
    np.savez(os.path.join(OUT, "decode_roundtrip.npz"), gt=gt, targets=t, decoded=e.decode_anchors(t, *a0))

    # ---- postprocess golden (640, C=2) -----------------------------------------------------------
    eval_anchors = synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd", border=0.))
    an = np.stack(eval_anchors[:4], -1)
    cls, loc, faces = synthetic.gen_predictions(1, an, max_faces=40)
    boxes = e.decode_anchors(loc, *eval_anchors[:4])
    sb, ss, idx = R.parse_by_class([640, 640], cls, boxes, 2, 0.01, 0, 5000, 750, 0.3, return_indices=True)
    np.savez_compressed(os.path.join(OUT, "postprocess_golden.npz"), image_index=np.int64(1), max_faces=np.int64(40),
                        boxes=sb[1], scores=ss[1], topk_index=idx[1][0][:int((ss[1] > 0).sum()) + 2000],
                        keep=idx[1][1])
    make_routing()
    make_vote()
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print("  %-32s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
