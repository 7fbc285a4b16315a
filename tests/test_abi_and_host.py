"""CPU tests of the boundary and the host logic (no GPU, no compute calls):
the C-ABI library loads, exports every symbol include/dan_b200.h declares, validates arguments with the
reference op's own conditions before touching the device, and the product refuses to run without CUDA."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dan_b200 import build, _lib
    build.build()          # nvcc cross-compiles sm_100a without a GPU
    return _lib.lib()


def test_header_symbols_exported(lib):
    from dan_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dan_b200.h")).read()
    declared = set(re.findall(r"\b(dan_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 24
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dan_[a-z0-9_]+)", out))
    assert declared <= exported
    assert lib.dan_version() == 100


def test_sass_is_sm100a(lib):
    from dan_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_header(lib):
    """ctypes mirrors of the PODs have the C sizes (4-byte fields only, no padding)."""
    from dan_b200 import _lib
    assert ctypes.sizeof(_lib.Pyramid) == 4 * (3 + 4 * 16 + 4 * 16 + 2 * 128)
    assert ctypes.sizeof(_lib.EncodeParams) == 4 * (14 + 1 + 3 * 8)        # + layout hint: num_grids, 3 x DAN_MAX_GRIDS
    assert ctypes.sizeof(_lib.PostprocessParams) == 4 * 12


def test_anchor_count_and_pyramid_validation(lib):
    from dan_b200 import functional as F, synthetic, _lib
    from dan_b200.utility.anchor_manipulator import AnchorEncoder
    enc = AnchorEncoder(0.4, 0.4, [0.1, 0.1, 0.2, 0.2])
    for size, n in (((640, 640), 34125), ((1600, 1600), 213294)):
        cfg = synthetic.pyramid_config("s3fd", size)
        hs, ws, ds = zip(*[enc.get_anchors_width_height(cfg["anchor_scales"][i], cfg["extra_scales"][i], cfg["anchor_ratios"][i])
                           for i in range(6)])
        pyr = F.make_pyramid(cfg["image_shape"], hs, ws, ds, cfg["offsets"], cfg["layer_shapes"], cfg["layer_strides"],
                             cfg["allowed_borders"], cfg["should_clips"])
        assert F.anchor_count(pyr) == n
    h, w, d = enc.get_anchors_width_height((16.,), (), (0.8,))
    np.testing.assert_allclose([float(h[0]), float(w[0])], [17.888544, 14.310835], rtol=1e-7)   # train_dan.py:181-183
    assert d == 1 and enc.get_anchors_count(d, (160, 160)) == (25600, 25600)
    with pytest.raises(_lib.DanError):
        F.make_pyramid([640, 640], [h] * 17, [w] * 17, [1] * 17, [0.5] * 17, [(1, 1)] * 17, [4] * 17, [0.] * 17, [False] * 17)
    bad = _lib.Pyramid()
    bad.num_layers = 0
    assert lib.dan_anchor_count(ctypes.byref(bad)) < 0


def test_layout_hint_from_pyramid(lib):
    """functional.grid_hint / encode_params(pyramid=...): only levels with one anchor per cell, a start that is a multiple of
    32, a width that is a multiple of 8 and a height that is a multiple of 4 are described (include/dan_b200.h)."""
    from dan_b200 import functional as F, synthetic
    from dan_b200.utility.anchor_manipulator import AnchorEncoder
    enc = AnchorEncoder(0.4, 0.4, [0.1, 0.1, 0.2, 0.2])

    def pyramid(cfg):
        hs, ws, ds = zip(*[enc.get_anchors_width_height(cfg["anchor_scales"][i], cfg["extra_scales"][i], cfg["anchor_ratios"][i])
                           for i in range(len(cfg["layer_shapes"]))])
        return F.make_pyramid(cfg["image_shape"], hs, ws, ds, cfg["offsets"], cfg["layer_shapes"], cfg["layer_strides"],
                              cfg["allowed_borders"], cfg["should_clips"])
    # 640^2: 160^2, 80^2, 40^2 qualify; 20^2 (width 20), 10^2, 5^2 do not
    assert F.grid_hint(pyramid(synthetic.pyramid_config("s3fd", (640, 640)))) == [(0, 160, 160), (25600, 80, 80), (32000, 40, 40)]
    # 1024^2: all six levels (256 ... 8 cells per side)
    g = F.grid_hint(pyramid(synthetic.pyramid_config("s3fd", (1024, 1024))))
    assert [w for _, w, _ in g] == [256, 128, 64, 32, 16, 8] and g[-1][0] == 256 ** 2 + 128 ** 2 + 64 ** 2 + 32 ** 2 + 16 ** 2
    # two anchors per cell on the first level: that level is skipped, the others keep their (shifted) starts
    cfg = synthetic.pyramid_config("s3fd", (640, 640))
    cfg["extra_scales"] = [(24.,)] + [()] * 5
    assert F.grid_hint(pyramid(cfg)) == [(51200, 80, 80), (57600, 40, 40)]
    p = F.encode_params(0.4, 0.4, [0.1, 0.1, 0.2, 0.2], True, pyramid=pyramid(synthetic.pyramid_config("s3fd", (640, 640))))
    assert p.num_grids == 3 and list(p.grid_start)[:3] == [0, 25600, 32000] and list(p.grid_w)[:3] == [160, 80, 40]
    assert F.encode_params(0.4, 0.4, [0.1, 0.1, 0.2, 0.2], True).num_grids == 0


def test_attr_validation_before_device(lib):
    """the op constructor's InvalidArgument conditions (small_mining_match.cc:292-305) are enforced at the C ABI."""
    null = ctypes.c_void_p(0)
    for bad in [(-0.1, .4, .4, 6, .3), (0., 0., .4, 6, .3), (0., .5, .4, 6, .3), (0., .4, 1., 6, .3), (0., .4, .4, 0, .3),
                (0., .4, .4, 6, 1.)]:
        rc = lib.dan_small_mining_match(null, 10, 3, bad[0], bad[1], bad[2], bad[3], bad[4], null, null, null, 0, null)
        assert rc == -1
        assert b"Need Attr" in lib.dan_last_error()
    assert lib.dan_small_mining_match(null, 10, 3, 0., .4, .4, 6, .3, null, null, null, 0, null) == -1   # NULL outputs
    assert lib.dan_dual_max_match(null, 10, 0, .35, .35, 1, 1, null, null, null, 0, null) == -1           # empty GT axis
    assert lib.dan_match_workspace_bytes(34125, 50) > 0
    assert lib.dan_encode_workspace_bytes(34125, 32, 1600) > 0
    assert lib.dan_postprocess_workspace_bytes(34125, 32, 2, 5000) >= 32 * 34125 * 8
    assert lib.dan_postprocess_workspace_bytes(34125, 32, 1, 5000) == 0


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dan_b200.utility import anchor_manipulator as am, bbox_util as bu, custom_op
    enc = am.AnchorEncoder(0.4, 0.4, [0.1, 0.1, 0.2, 0.2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc.generate_anchors_by_offset(torch.tensor([16.]), torch.tensor([16.]), 1, [64, 64], (16, 16), 4)
    with pytest.raises((RuntimeError, TypeError, AssertionError)):
        custom_op.small_mining_match(torch.zeros(4, 2), 0., .4, .4, 6, .3)
    with pytest.raises((RuntimeError, TypeError, AssertionError)):
        bu.parse_by_class([64, 64], torch.zeros(10, 2), torch.zeros(10, 4), 2, 0.01, 0, 100, 10, 0.3)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under dan_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dan_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), os.path.join(dirpath, f)
                assert "liboracle" not in text and "_ref/" not in text, os.path.join(dirpath, f)


def test_heap_order_transcription_matches_libstdcxx(tmp_path):
    """dan_b200/csrc/heap_order.cuh (used by the stage-3 tie path) vs the real std::priority_queue."""
    exe = str(tmp_path / "heap_check")
    subprocess.run(["g++", "-O2", "-std=c++14", "-I" + os.path.join(ROOT, "dan_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "heap_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "20000"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout


def test_shard_ranges_and_csr():
    from dan_b200 import pipeline, synthetic
    for batch in (256, 32, 7, 1):
        for world in (1, 2, 4, 8):
            shards = [pipeline.shard_range(batch, r, world) for r in range(world)]
            assert shards[0].lo == 0 and shards[-1].hi == batch
            assert all(a.hi == b.lo for a, b in zip(shards, shards[1:]))
            sizes = [s.hi - s.lo for s in shards]
            assert max(sizes) - min(sizes) <= 1
    gts = [synthetic.gen_faces(i, 20) for i in range(9)] + [synthetic.gen_adversarial("empty")]
    cat, offs = synthetic.to_csr(gts)
    cat, offs = torch.from_numpy(cat), torch.from_numpy(offs)
    sh = pipeline.shard_range(10, 1, 3)
    b, o = pipeline.shard_csr(cat, offs, sh)
    assert o[0] == 0 and o.dtype == torch.int32 and o.numel() == sh.hi - sh.lo + 1
    for i in range(sh.lo, sh.hi):
        np.testing.assert_array_equal(b[o[i - sh.lo]:o[i - sh.lo + 1]].numpy(), gts[i])


def test_detection_slab_layout():
    from dan_b200 import pipeline
    slab = pipeline.DetectionSlab(5, 1, 750, "cpu")
    counts, scores, boxes = slab.views()
    assert counts.dtype == torch.int32 and tuple(boxes.shape) == (5, 1, 750, 4) and tuple(scores.shape) == (5, 1, 750)
    assert slab.words == pipeline.DetectionSlab.words_for(5, 1, 750)
    assert (boxes.data_ptr() - slab.buf.data_ptr()) % 16 == 0
    counts[2, 0] = 7
    boxes[2, 0, 6, 3] = 1.5
    scores[4, 0, 749] = 2.5
    c2, s2, b2 = slab.views(slab.buf.clone())
    assert int(c2[2, 0]) == 7 and float(b2[2, 0, 6, 3]) == 1.5 and float(s2[4, 0, 749]) == 2.5


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    from dan_b200 import pipeline
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sh = pipeline.shard_range(7, rank, world)
    cap = (7 + world - 1) // world
    slab = pipeline.DetectionSlab(cap, 1, 8, "cpu")
    counts, scores, boxes = slab.views()
    for i in range(sh.hi - sh.lo):
        g = sh.lo + i
        counts[i, 0] = g % 5
        scores[i, 0, :g % 5] = float(g)
        boxes[i, 0, :g % 5, :] = float(g) + 0.5
    gathered = pipeline.gather_detections(slab, world)
    flat = pipeline.flatten_detections(gathered, [pipeline.shard_range(7, r, world).hi - pipeline.shard_range(7, r, world).lo
                                                  for r in range(world)])
    ok = len(flat) == 7
    for g, det in enumerate(flat):
        b, s = det[1]
        ok = ok and b.shape[0] == g % 5 and bool((s == float(g)).all()) and bool((b == float(g) + 0.5).all())
    # the slabs of several steps in one send buffer -> ONE collective for the group (bench.py's steps in flight)
    import torch
    words = pipeline.DetectionSlab.words_for(cap, 1, 8)
    send = torch.zeros(3 * words)
    slabs = [pipeline.DetectionSlab(cap, 1, 8, "cpu", buf=send[j * words:(j + 1) * words]) for j in range(3)]
    for j, sl in enumerate(slabs):
        sl.views()[0][:] = 100 * rank + j
    recv = pipeline.gather_slab_group(send, torch.empty(world * 3 * words))
    for r in range(world):
        for j, sl in enumerate(slabs):
            c = sl.views(recv[(r * 3 + j) * words:(r * 3 + j + 1) * words])[0]
            ok = ok and bool((c == 100 * r + j).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    """the N>1 path: per-image sharding + ONE all_gather of fixed-capacity slabs (gloo, world_size 2, CPU)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
