"""Seeded synthetic inputs for the anchor hot path (SURVEY.md 8(d) generators) and the
reference's hard-coded stride pyramids.  numpy only; used by tests/ and bench.py to
produce the SAME inputs for the CUDA path and for the CPU oracle.  Nothing here is
part of the measured path."""
from __future__ import annotations

import math

import numpy as np

BASE_SEED = 20180817   # the reference's tf_random_seed (train_sfd.py:92-93)
f32 = np.float32


# ---------------------------------------------------------------------------------
# pyramids hard-coded in the reference scripts
# ---------------------------------------------------------------------------------
def layer_shapes_for(image_size, strides=(4, 8, 16, 32, 64, 128)):
    """'same'-padded feature map sides: ceil(S / stride) (net/sfd_net.py:127-156)."""
    h, w = image_size
    return [(int(math.ceil(h / s)), int(math.ceil(w / s))) for s in strides]


def pyramid_config(kind="s3fd", image_size=(640, 640), border=None, clip=False):
    """kind 's3fd' (train_sfd.py:179-184, ratio 1) | 'dan' (train_dan.py:181-186, ratio 0.8).
    border=None -> the training setting (border = image size => mask all true);
    eval scripts use border 0 (eval_sfd.py:266-269)."""
    strides = [4, 8, 16, 32, 64, 128]
    ratio = {"s3fd": 1.0, "pyramidbox": 1.0, "dan": 0.8}[kind]
    cfg = dict(
        image_shape=list(image_size),
        anchor_scales=[(16.,), (32.,), (64.,), (128.,), (256.,), (512.,)],
        extra_scales=[(), (), (), (), (), ()],
        anchor_ratios=[(ratio,)] * 6,
        layer_shapes=layer_shapes_for(image_size, strides),
        layer_strides=strides,
        offsets=[0.5] * 6,
        allowed_borders=[float(image_size[0]) if border is None else float(border)] * 6,
        should_clips=[bool(clip)] * 6,
    )
    return cfg


def build_anchors(encoder, cfg):
    """Run get_anchors_width_height + get_all_anchors of ANY AnchorEncoder implementation
    (dan_b200's or the oracle's) on a pyramid config.  -> (ymin, xmin, ymax, xmax, inside_mask)."""
    hs, ws, ds = [], [], []
    for i in range(len(cfg["layer_shapes"])):
        h, w, d = encoder.get_anchors_width_height(cfg["anchor_scales"][i], cfg["extra_scales"][i],
                                                   cfg["anchor_ratios"][i])
        hs.append(h)
        ws.append(w)
        ds.append(d)
    return encoder.get_all_anchors(cfg["image_shape"], hs, ws, ds, cfg["offsets"], cfg["layer_shapes"],
                                   cfg["layer_strides"], cfg["allowed_borders"], cfg["should_clips"])


def anchors_numpy(cfg):
    """Plain numpy anchors for the GENERATORS only (float64 math, not a parity path)."""
    out = []
    for i, (lh, lw) in enumerate(cfg["layer_shapes"]):
        s = cfg["layer_strides"][i]
        hw = []
        for sc in cfg["extra_scales"][i]:
            hw.append((sc, sc))
        for sc in cfg["anchor_scales"][i]:
            for r in cfg["anchor_ratios"][i]:
                hw.append((sc / math.sqrt(r), sc * math.sqrt(r)))
        ys, xs = np.meshgrid(np.arange(lh), np.arange(lw), indexing="ij")
        cy = (ys + 0.5) * s
        cx = (xs + 0.5) * s
        for_layer = []
        for (h, w) in hw:
            for_layer.append(np.stack([cy - (h - 1) / 2, cx - (w - 1) / 2, cy + (h - 1) / 2, cx + (w - 1) / 2], -1))
        out.append(np.stack(for_layer, axis=2).reshape(-1, 4))
    return np.concatenate(out, 0)


# ---------------------------------------------------------------------------------
# ground truth generators
# ---------------------------------------------------------------------------------
def _faces(rng, m, smin, smax, size, snap):
    s = np.exp(rng.uniform(np.log(smin), np.log(smax), m))
    h = s * rng.uniform(1.0, 1.4, m)
    w = s
    h = np.minimum(h, size[0] - 1.0)
    w = np.minimum(w, size[1] - 1.0)
    ymin = rng.uniform(0, (size[0] - 1.0) - h + 1e-6)
    xmin = rng.uniform(0, (size[1] - 1.0) - w + 1e-6)
    boxes = np.stack([ymin, xmin, ymin + h - 1.0, xmin + w - 1.0], axis=1)
    if snap:
        boxes = np.round(boxes / snap) * snap
    boxes = boxes.astype(f32)
    # the reference's small-face filter (preprocessing/sfd_preprocessing.py:544-551): h > 6 & w > 3
    hh = boxes[:, 2] - boxes[:, 0] + 1
    ww = boxes[:, 3] - boxes[:, 1] + 1
    return boxes[(hh > 6) & (ww > 3)]


def gen_faces(image_index, max_faces=50, size=(640, 640), snap=0.0, min_faces=1, smin=8.0, smax=400.0):
    """G1 faces (snap=0) / G1-snap (snap=1 or 4: corners on a grid -> exact IoU ties)."""
    rng = np.random.default_rng(BASE_SEED + image_index)
    m = int(rng.integers(min_faces, max_faces + 1))
    return _faces(rng, m, smin, min(smax, min(size) * 0.625), size, snap)


def gen_dense_tiny(image_index, size=(640, 640), lo=200, hi=1000, snap=0.0):
    """G2 dense tiny faces: M ~ U{200..1000}, side exp(U(ln 7, ln 40))."""
    rng = np.random.default_rng(BASE_SEED + 100000 + image_index)
    m = int(rng.integers(lo, hi + 1))
    return _faces(rng, m, 7.0, 40.0, size, snap)


def gen_adversarial(kind, size=(640, 640)):
    """G-adv: degenerate ground truth sets."""
    h, w = size
    if kind == "empty":
        return np.zeros((0, 4), f32)
    if kind == "outside":      # all-zero IoU column
        return np.asarray([[100, 100, 180, 170], [-500, -500, -400, -420]], f32)
    if kind == "duplicate":
        return np.asarray([[50, 60, 120, 130], [50, 60, 120, 130], [300, 310, 420, 400]], f32)
    if kind == "anchor_identical":   # IoU == 1 with an S3FD stride-8 anchor
        return np.asarray([[4 - 15.5, 4 - 15.5, 4 + 15.5, 4 + 15.5], [100 - 15.5, 204 - 15.5, 100 + 15.5, 204 + 15.5]], f32)
    if kind == "tiny":
        return np.asarray([[10, 10, 17, 14], [300.2, 300.7, 309.1, 306.3], [630, 630, 639, 639]], f32)
    if kind == "huge":
        return np.asarray([[0, 0, h - 1, w - 1]], f32)
    raise ValueError(kind)


def to_csr(list_of_boxes):
    """list of [M_i,4] -> (concat [sum M,4] fp32, offsets int32 [B+1])."""
    offs = np.zeros(len(list_of_boxes) + 1, np.int32)
    for i, b in enumerate(list_of_boxes):
        offs[i + 1] = offs[i] + len(b)
    cat = np.concatenate([np.asarray(b, f32).reshape(-1, 4) for b in list_of_boxes], 0) if list_of_boxes else \
        np.zeros((0, 4), f32)
    return np.ascontiguousarray(cat, f32), offs


# ---------------------------------------------------------------------------------
# G3: predictions of a "trained detector"
# ---------------------------------------------------------------------------------
def _iou_chunked(anchors, faces):
    a = anchors[:, None, :].astype(np.float64)
    g = faces[None, :, :].astype(np.float64)
    ih = np.maximum(np.minimum(a[..., 2], g[..., 2]) - np.maximum(a[..., 0], g[..., 0]) + 1, 0)
    iw = np.maximum(np.minimum(a[..., 3], g[..., 3]) - np.maximum(a[..., 1], g[..., 1]) + 1, 0)
    inter = ih * iw
    aa = (a[..., 2] - a[..., 0] + 1) * (a[..., 3] - a[..., 1] + 1)
    ag = (g[..., 2] - g[..., 0] + 1) * (g[..., 3] - g[..., 1] + 1)
    return inter / (aa + ag - inter)


def gen_predictions(image_index, anchors, size=(640, 640), max_faces=300, prior_scaling=(0.1, 0.1, 0.2, 0.2)):
    """G3: K_true ~ U{1..max_faces} planted faces; anchors with IoU > 0.35 to a face get foreground
    logits (N(0,1), N(5,1)) and loc = encode(face) + N(0, 0.3); all others (N(8,1), N(0,1)) and N(0,0.5).
    -> (cls_pred [N,2] fp32, loc_pred [N,4] fp32, faces [K,4])"""
    rng = np.random.default_rng(BASE_SEED + 200000 + image_index)
    k_true = int(rng.integers(1, max_faces + 1))
    faces = _faces(rng, k_true, 8.0, min(400.0, min(size) * 0.625), size, 0.0)
    n = anchors.shape[0]
    cls = np.stack([rng.normal(8.0, 1.0, n), rng.normal(0.0, 1.0, n)], axis=1)
    loc = rng.normal(0.0, 0.5, (n, 4))
    best_iou = np.zeros(n)
    best_face = np.zeros(n, np.int64)
    for c0 in range(0, n, 16384):
        iou = _iou_chunked(anchors[c0:c0 + 16384], faces)
        best_iou[c0:c0 + 16384] = iou.max(1)
        best_face[c0:c0 + 16384] = iou.argmax(1)
    fg = best_iou > 0.35
    nf = int(fg.sum())
    if nf:
        a = anchors[fg].astype(np.float64)
        g = faces[best_face[fg]].astype(np.float64)
        ah, aw = a[:, 2] - a[:, 0] + 1, a[:, 3] - a[:, 1] + 1
        acy, acx = (a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2
        gh, gw = g[:, 2] - g[:, 0] + 1, g[:, 3] - g[:, 1] + 1
        gcy, gcx = (g[:, 0] + g[:, 2]) / 2, (g[:, 1] + g[:, 3]) / 2
        enc = np.stack([(gcy - acy) / ah / prior_scaling[0], (gcx - acx) / aw / prior_scaling[1],
                        np.log(gh / ah) / prior_scaling[2], np.log(gw / aw) / prior_scaling[3]], 1)
        loc[fg] = enc + rng.normal(0.0, 0.3, (nf, 4))
        cls[fg] = np.stack([rng.normal(0.0, 1.0, nf), rng.normal(5.0, 1.0, nf)], axis=1)
    return cls.astype(f32), loc.astype(f32), faces


def gen_routing(image_index, feat_heights, feat_widths, depths, strides, mode="plain"):
    """Inputs of DynamicAnchorRouting's evaluation branch for one image, all layers concatenated (layer by layer,
    (y, x, depth) inside a layer) -> (anchors [N,4], gt_targets [N,4], labels [N], mask_in [N] int32).

    Decoded stage-1 boxes = a box centred near the anchor's cell (sigma 1.5 cells, so boxes really move to other
    cells and some leave the feature map), stage-2 offsets already divided by 2 x prior scaling (eval_dan.py:384),
    stage-2 probabilities, stage-1 mask (~60 % ones).  mode "ties": box centres snapped to half cells (std::round
    sees exact .5 values), probabilities quantised to 1/8 (equal labels, exact zeros), some degenerate boxes."""
    rng = np.random.default_rng(BASE_SEED + 400000 + image_index)
    out = []
    for H, W, D, S in zip(feat_heights, feat_widths, depths, strides):
        n = H * W * D
        yy, xx, _ = np.meshgrid(np.arange(H), np.arange(W), np.arange(D), indexing="ij")
        cy = (yy.reshape(-1) + 0.5) * S + rng.normal(0, 1.5 * S, n)
        cx = (xx.reshape(-1) + 0.5) * S + rng.normal(0, 1.5 * S, n)
        h = np.abs(rng.normal(4 * S, 2 * S, n))
        w = np.abs(rng.normal(4 * S, 2 * S, n))
        lab = rng.uniform(0, 1, n)
        if mode == "ties":
            cy, cx = np.round(cy / (S / 2)) * (S / 2), np.round(cx / (S / 2)) * (S / 2)
            h, w = np.round(h), np.round(w)
            lab = np.round(lab * 8) / 8
            h[::37] = 0.5                                  # narrower than one pixel: never a source
        a = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], -1)
        t = rng.normal(0, 1, (n, 4)) / np.array([20., 20., 10., 10.])
        m = (rng.uniform(0, 1, n) > 0.4).astype(np.int32)
        out.append((a.astype(f32), t.astype(f32), lab.astype(f32), m))
    return tuple(np.concatenate([o[k] for o in out], 0) for k in range(4))


def gen_vote_dets(image_index, n, faces=40, background=0.5):
    """Input of bbox_vote for one image: the stack of multi-scale / flipped detections, rows (xmin, ymin, xmax, ymax,
    score).  (1 - background) of the rows are jittered copies of `faces` true boxes (they vote for each other), the rest
    are scattered low-score boxes (mostly singletons, which bbox_vote drops).  Scores are continuous and distinct (the
    reference's argsort leaves ties unspecified); every 9th row is degenerate (zero width): it never merges."""
    rng = np.random.default_rng(BASE_SEED + 500000 + image_index)
    k = max(int(faces), 1)
    side = np.exp(rng.uniform(np.log(10.), np.log(300.), k))
    cx, cy = rng.uniform(0, 640, k), rng.uniform(0, 640, k)
    nf = int(round(n * (1. - background)))
    idx = rng.integers(0, k, nf)
    jit = side[idx] * 0.08
    fx, fy = cx[idx] + rng.normal(0, 1, nf) * jit, cy[idx] + rng.normal(0, 1, nf) * jit
    fw, fh = side[idx] * np.exp(rng.normal(0, 0.08, nf)), side[idx] * 1.2 * np.exp(rng.normal(0, 0.08, nf))
    fs = rng.uniform(0.3, 1.0, nf)
    nb = n - nf
    bs_ = np.exp(rng.uniform(np.log(8.), np.log(200.), nb))
    bx, by = rng.uniform(0, 640, nb), rng.uniform(0, 640, nb)
    bsc = rng.uniform(0.0, 0.3, nb)
    x = np.concatenate([fx, bx]); y = np.concatenate([fy, by])
    w = np.concatenate([fw, bs_]); h = np.concatenate([fh, bs_ * 1.2])
    sc = np.concatenate([fs, bsc])
    det = np.stack([x - w / 2, y - h / 2, x + w / 2, y + h / 2, sc], -1).astype(f32)
    det = det[rng.permutation(n)]
    det[::9, 2] = det[::9, 0] - 1                 # zero area under the +1 convention
    _, first = np.unique(det[:, 4], return_index=True)
    return det[np.sort(first)]                    # distinct scores
