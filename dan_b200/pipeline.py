"""Batched hot path = training encode + evaluation postprocess, sharded per image.

Images are independent (the reference itself handles one image at a time:
dataset/dataset_common.py:150, eval_sfd.py:311-327), so a batch is split into
contiguous blocks, one per rank, with NO data-path collective.  The only exchange
is the variable-length detections at the end: every rank's NMS kernel writes its
zero padded keep lists and per-image counts straight into ONE fixed-capacity slab,
and one ``all_gather_into_tensor`` (NCCL on GPUs, gloo in the CPU tests) moves
all slabs -- the counts travel inside the slab, so there is no size exchange.
"""
from __future__ import annotations

import os
from collections import namedtuple

import torch

from . import functional as F

Shard = namedtuple("Shard", ["lo", "hi"])


def shard_range(batch, rank, world_size):
    """Contiguous block [lo, hi) of a batch owned by `rank` (first `batch % world` ranks get one more)."""
    base, rem = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return Shard(lo, lo + base + (1 if rank < rem else 0))


def shard_csr(gt_boxes, gt_offsets, shard):
    """Slice a CSR ground-truth batch to the images [lo, hi) (offsets rebased to 0)."""
    offs = gt_offsets[shard.lo:shard.hi + 1]
    start, end = int(offs[0]), int(offs[-1])
    return gt_boxes[start:end], (offs - offs[0]).to(gt_offsets.dtype)


class DetectionSlab(object):
    """One rank's detections as a single contiguous buffer.

    layout (all 4-byte words): counts int32 [B, L] (padded to 16 B) | boxes f32 [B, L, K, 4] | scores f32 [B, L, K]
    with L = num_classes - 1 and K = nms_topk.  ``views()`` returns typed views that the NMS kernel
    writes in place, so the slab needs no packing step before the collective."""

    def __init__(self, images, lists, nms_topk, device, buf=None):
        """``buf``: optional caller-owned fp32 storage of exactly words_for(...) words (e.g. a slice of a larger send
        buffer, so that the slabs of several steps travel in one collective)."""
        self.images, self.lists, self.k = int(images), int(lists), int(nms_topk)
        self.n_counts = self.images * self.lists
        self.n_scores = self.n_counts * self.k
        self.n_boxes = self.n_scores * 4
        # counts region padded to 4 words so that the box region stays 16-byte aligned
        self.counts_words = (self.n_counts + 3) // 4 * 4
        # (a multiple of 4 words, so that slabs laid end to end keep their box regions 16-byte aligned)
        self.words = (self.counts_words + self.n_scores + self.n_boxes + 3) // 4 * 4
        if buf is None:
            buf = torch.zeros(self.words, dtype=torch.float32, device=device)
        if buf.numel() != self.words or buf.dtype != torch.float32 or not buf.is_contiguous():
            raise ValueError("slab buffer must be %d contiguous fp32 words" % self.words)
        self.buf = buf

    @staticmethod
    def words_for(images, lists, nms_topk):
        nc = images * lists
        return ((nc + 3) // 4 * 4 + nc * nms_topk * 5 + 3) // 4 * 4

    def views(self, buf=None):
        buf = self.buf if buf is None else buf
        o = 0
        counts = buf[o:o + self.n_counts].view(torch.int32).view(self.images, self.lists)
        # boxes first after the counts (16-byte alignment for float4 stores), then scores
        o = self.counts_words
        boxes = buf[o:o + self.n_boxes].view(self.images, self.lists, self.k, 4)
        o += self.n_boxes
        scores = buf[o:o + self.n_scores].view(self.images, self.lists, self.k)
        return counts, scores, boxes


def gather_detections(slab, world_size, group=None):
    """ONE collective for the variable-length detections.  Returns a list (rank order) of
    (counts, scores, boxes) views; every rank must pass a slab of identical capacity."""
    import torch.distributed as dist
    if world_size == 1 or not dist.is_initialized():
        return [slab.views()]
    out = torch.empty(world_size * slab.words, dtype=slab.buf.dtype, device=slab.buf.device)
    dist.all_gather_into_tensor(out, slab.buf, group=group)
    return [slab.views(out[r * slab.words:(r + 1) * slab.words]) for r in range(world_size)]


def gather_slab_group(send, recv, group=None):
    """ONE collective for the slabs of several consecutive steps: ``send`` is a contiguous buffer holding their
    DetectionSlabs back to back, ``recv`` has world_size times its size (rank-major)."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv


class DeviceGather(object):
    """The detection exchange on the DEVICE timeline: one NCCL communicator owned by libdan_b200
    (dan_comm_init) and one ncclAllGather per call, enqueued on torch's current stream by
    dan_gather_detections.  It is captured like any kernel when the step is recorded as a CUDA graph,
    so a replayed step costs no host time for the collective.  torch.distributed is only used once, to
    ship the 128-byte NCCL id from rank 0 to the other ranks."""

    def __init__(self, rank, world_size, device, group=None, max_ctas=None):
        import ctypes
        import torch.distributed as dist
        from . import _lib as L
        self.rank, self.world_size, self.device = int(rank), int(world_size), device
        self._lib = L
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            L.check(L.lib().dan_comm_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if self.world_size > 1:
            backend = dist.get_backend(group)
            t = ident.to(device) if backend == "nccl" else ident
            dist.broadcast(t, src=0, group=group)
            ident = t.cpu()
        raw = (ctypes.c_ubyte * 128)(*[int(v) for v in ident.tolist()])
        comm = ctypes.c_void_p(0)
        if max_ctas is None:
            max_ctas = int(os.environ.get("DAN_NCCL_MAX_CTAS", "0"))      # CTAs of the collective (0 = NCCL's default)
        with torch.cuda.device(device):
            L.check(L.lib().dan_comm_init_ctas(raw, self.rank, self.world_size, int(max_ctas), ctypes.byref(comm)))
        self._comm = comm

    def gather(self, send, recv):
        """recv [world, len(send)] <- send of every rank (rank order), on the current stream."""
        L = self._lib
        if recv.numel() != self.world_size * send.numel() or recv.dtype != send.dtype:
            raise ValueError("recv must hold world_size slabs of the send buffer's size and dtype")
        with torch.cuda.device(self.device):
            L.check(L.lib().dan_gather_detections(self._comm, L.dev_ptr(send), L.dev_ptr(recv), send.numel() * send.element_size(),
                                                  L.stream_ptr()))
        return recv

    def close(self):
        if self._comm is not None and self._comm.value:
            self._lib.lib().dan_comm_destroy(self._comm)
        self._comm = None


class _DeviceMemory(object):
    """A raw device allocation exposed through __cuda_array_interface__ so that torch can wrap it without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerExchange(object):
    """The detection exchange WITHOUT a collective kernel (ranks of one node): every rank allocates ONE receive buffer
    per buffer set `[world slabs | world flags]` through the C ABI (cudaMalloc + CUDA IPC handle), the ranks map each
    other's buffers, and the NMS kernel stores every slab row it writes into all of them over NVLink
    (dan_postprocess_batch_peers).  `wait()` enqueues the one-warp kernel that returns when the slabs of all ranks of
    this step have arrived.  torch.distributed is used once, to exchange the 64-byte IPC handles."""

    def __init__(self, rank, world_size, device, slab_words, num_sets=1, group=None):
        import ctypes
        import torch.distributed as dist
        from . import _lib as L
        self._lib = L
        self.rank, self.world_size, self.device = int(rank), int(world_size), device
        self.slab_words, self.num_sets = int(slab_words), int(num_sets)
        if self.world_size > L.DAN_MAX_PEERS:
            raise ValueError("at most %d ranks" % L.DAN_MAX_PEERS)
        slab_bytes = 4 * self.slab_words
        self.set_bytes = (self.world_size * slab_bytes + 4 * 32 + 255) // 256 * 256      # slabs, then the arrival flags
        self.flag_offset = self.world_size * slab_bytes
        total = self.num_sets * self.set_bytes
        ptr, handle = ctypes.c_void_p(0), (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            L.check(L.lib().dan_peer_alloc(total, ctypes.byref(ptr), handle))
        self._own = int(ptr.value)
        self._mem = torch.as_tensor(_DeviceMemory(self._own, total), device=device)      # uint8 view of the allocation
        mine = torch.tensor(list(handle), dtype=torch.uint8)
        self._bases = [self._own] * self.world_size
        self._opened = []
        if self.world_size > 1:
            backend = dist.get_backend(group)
            t = mine.to(device) if backend == "nccl" else mine
            allh = torch.empty((self.world_size, 64), dtype=torch.uint8, device=t.device)
            dist.all_gather_into_tensor(allh, t, group=group)
            allh = allh.cpu()
            for q in range(self.world_size):
                if q == self.rank:
                    continue
                raw = (ctypes.c_ubyte * 64)(*[int(v) for v in allh[q].tolist()])
                p = ctypes.c_void_p(0)
                with torch.cuda.device(device):
                    L.check(L.lib().dan_peer_open(raw, ctypes.byref(p)))
                self._bases[q] = int(p.value)
                self._opened.append(p)
        self._state = torch.zeros((self.num_sets, 2), dtype=torch.int32, device=device)

    def recv(self, s):
        """fp32 view [world * slab_words] of this rank's receive buffer of set s (rank-major, like an all-gather)."""
        o = s * self.set_bytes
        return self._mem[o:o + self.flag_offset].view(torch.float32)

    def flags(self, s):
        o = s * self.set_bytes + self.flag_offset
        return self._mem[o:o + 4 * 32].view(torch.int32)

    def args(self, s, slab_buf):
        """dan_peer_exchange for a step of set s whose slab lives in `slab_buf` (every rank is a destination, the own
        receive buffer included)."""
        L = self._lib
        a = L.PeerExchangeArgs()
        a.num_destinations = self.world_size
        slab_bytes = 4 * self.slab_words
        for q in range(self.world_size):
            dst = self._bases[q] + s * self.set_bytes
            a.delta_bytes[q] = dst + self.rank * slab_bytes - int(slab_buf.data_ptr())
            a.flag[q] = dst + self.flag_offset + 4 * self.rank
        a.state = int(self._state[s].data_ptr())
        return a

    def wait(self, s):
        """Returns (on the stream) when the slabs of all ranks for the latest step of set s are in recv(s)."""
        L = self._lib
        with torch.cuda.device(self.device):
            L.check(L.lib().dan_wait_detections(L.dev_ptr(self.flags(s)), L.dev_ptr(self._state[s]), self.world_size, L.stream_ptr()))

    def close(self):
        L = self._lib
        for p in self._opened:
            L.lib().dan_peer_close(p)
        self._opened = []
        if self._own:
            self._mem = None
            L.lib().dan_peer_free(ctypes_void(self._own))
            self._own = 0


def ctypes_void(v):
    import ctypes
    return ctypes.c_void_p(v)


def flatten_detections(gathered, image_counts=None):
    """Ragged result for the host: list over images (global order) of dict(class -> (boxes [n,4], scores [n])).
    `image_counts[r]` = number of real images of rank r (ranks may own fewer images than the slab capacity; HotPath
    zeroes the unused tail of its slab, so without image_counts those rows read as images without detections)."""
    result = []
    for r, (counts, scores, boxes) in enumerate(gathered):
        n_img = counts.shape[0] if image_counts is None else image_counts[r]
        c = counts.cpu()
        for i in range(n_img):
            per_class = {}
            for l in range(counts.shape[1]):
                n = int(c[i, l])
                per_class[l + 1] = (boxes[i, l, :n], scores[i, l, :n])
            result.append(per_class)
    return result


class HotPath(object):
    """The whole path for one rank: encode (training side) + postprocess (evaluation side).

    The two halves are independent, so step() forks the encode chain onto a side stream and joins it at the end
    (inside a CUDA-graph capture this becomes two parallel branches); each half has its own workspace."""

    def __init__(self, anchors_train, inside_mask, encode_params, postprocess_params, anchors_eval=None,
                 images_per_rank=None, workspaces=None, overlap=True, slab_buffer=None, device_gather=None, recv_buffer=None,
                 peer_exchange=None, peer_set=0):
        from . import _lib
        self.anchors_train = anchors_train          # (ymin, xmin, ymax, xmax)
        self.inside_mask = inside_mask
        self.anchors_eval = anchors_eval if anchors_eval is not None else anchors_train
        self.enc_params = encode_params
        self.pp_params = postprocess_params
        self.device = anchors_train[0].device
        self.images_per_rank = images_per_rank
        self.ws_enc, self.ws_pp = workspaces if workspaces is not None else (_lib.Workspace(), _lib.Workspace())
        self.overlap = overlap
        self._side = None
        self._slab_buffer = slab_buffer             # optional caller-owned storage of the detection slab
        self._gather = device_gather                # DeviceGather: the all-gather is enqueued right after the NMS kernel
        self._recv = recv_buffer                    # [world, slab words] when device_gather is given
        self._peers = peer_exchange                 # PeerExchange: the NMS kernel stores the slab into every rank's buffer
        self._peer_set = int(peer_set)
        self._peer_args = None
        if peer_exchange is not None:
            self._recv = peer_exchange.recv(self._peer_set)
        self._slab = None
        self._enc_out = None
        self._aux = None

    def _buffers(self, images):
        n = self.anchors_train[0].numel()
        k, lists = self.pp_params.nms_topk, self.pp_params.num_classes - 1
        cap = images if self.images_per_rank is None else self.images_per_rank
        if self._slab is None or self._slab.images != cap:
            self._slab = DetectionSlab(cap, lists, k, self.device, buf=self._slab_buffer)
            self._aux = (torch.empty((cap, lists, k), dtype=torch.int32, device=self.device),
                         torch.empty((cap, lists, k), dtype=torch.int32, device=self.device))
        if self._enc_out is None or self._enc_out[0].shape[0] != images:
            d = self.device
            self._enc_out = (torch.empty((images, n, 4), dtype=torch.float32, device=d),
                             torch.empty((images, n), dtype=torch.int64, device=d),
                             torch.empty((images, n), dtype=torch.float32, device=d),
                             torch.empty((images, n, 4), dtype=torch.float32, device=d),
                             None)

    def step(self, gt_boxes, gt_offsets, cls_pred, loc_pred, profile=False):
        """One pass over this rank's images.  Everything is enqueued (current stream + one forked side stream).
        profile=True runs serially and returns per-kernel CUDA-event durations as a third value (synchronises)."""
        images = gt_offsets.numel() - 1
        self._buffers(images)
        counts, scores, boxes = self._slab.views()
        if images < self._slab.images:
            # a rank that owns fewer images than the slab holds: the tail travels in the collective, it must not carry
            # stale detections
            counts[images:].zero_()
            scores[images:].zero_()
            boxes[images:].zero_()
        det_out = (boxes[:images], scores[:images], counts[:images], self._aux[0][:images], self._aux[1][:images])
        if profile or not self.overlap:
            enc = F.encode_batch(self.enc_params, *self.anchors_train, self.inside_mask, gt_boxes, gt_offsets,
                                 out=self._enc_out, workspace=self.ws_enc, profile=profile)
            det = F.postprocess_batch(self.pp_params, cls_pred, loc_pred=loc_pred, anchors=self.anchors_eval,
                                      out=det_out, workspace=self.ws_pp, profile=profile,
                                      peers=None if profile else self._peer_arguments())
            if profile:
                return enc[0], det[0], {"enc_pass1": enc[1][0], "enc_pass2": enc[1][1], "enc_pass3": enc[1][2],
                                        "pp_filter": det[1][0], "nms_greedy": det[1][1]}
            self._enqueue_gather()
            return enc, det
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            enc = F.encode_batch(self.enc_params, *self.anchors_train, self.inside_mask, gt_boxes, gt_offsets,
                                 out=self._enc_out, workspace=self.ws_enc)
        det = F.postprocess_batch(self.pp_params, cls_pred, loc_pred=loc_pred, anchors=self.anchors_eval,
                                  out=det_out, workspace=self.ws_pp, peers=self._peer_arguments())
        self._enqueue_gather()      # on the stream that just ran the NMS kernel; the encode branch is not waited for
        main.wait_stream(self._side)
        return enc, det

    def _peer_arguments(self):
        if self._peers is None:
            return None
        if self._peer_args is None:
            self._peer_args = self._peers.args(self._peer_set, self._slab.buf)
        return self._peer_args

    def _enqueue_gather(self):
        """The only exchange of the path: all-gather of the detection slabs, enqueued behind the NMS kernel (NCCL), or
        the wait for the slabs the ranks' NMS kernels stored into each other's buffers (peer exchange)."""
        if self._peers is not None:
            self._peers.wait(self._peer_set)
            return
        if self._gather is None:
            return
        if self._recv is None:
            self._recv = torch.empty(self._gather.world_size * self._slab.words, dtype=torch.float32, device=self.device)
        self._gather.gather(self._slab.buf, self._recv)

    def gather(self, world_size, group=None):
        return gather_detections(self._slab, world_size, group)

    def gathered(self):
        """Views (rank order) of the slabs the in-graph all-gather delivered (device_gather mode)."""
        w = self._slab.words
        world = self._peers.world_size if self._peers is not None else self._gather.world_size
        return [self._slab.views(self._recv[r * w:(r + 1) * w]) for r in range(world)]
