"""dist3d -- motion's full-volume 3-D DCT (`motion -b 0x0x0`, /root/reference/motion/motion.c:525-554, 641, 753)
sharded over the GPUs of one node: one process per GPU, torch.distributed (NCCL over NVLink) for the exchange.

Decomposition: rank r owns frames [r D/G, (r+1) D/G) of the [D][H][W] volume.
  forward : (1) 2-D REDFT10 over (h, w) of every local frame        -- one batched plan, no communication
            (2) all-to-all: rank r now owns every frame of the flattened spatial positions [r HW/G, (r+1) HW/G)
            (3) 1-D REDFT10 of length D along time over those positions -- strided-axis (column) kernel
  inverse : (3) (2) (1) mirrored with REDFT01.
The flattened h*w index is partitioned (not rows), so chroma planes whose height does not divide by G still shard.
With G = 1 the same object runs one rank-3 plan.  The reference has no distributed path; its single buffer
(`coeffs`, motion.c:500) is what the slabs partition.

Two exchange modes:
  * "nccl" : pack + `all_to_all_single` + local pass (the baseline; also what the gloo CPU tests exercise).
  * "peer" : the exchange is FUSED into the transform.  The receive buffers live in symmetric memory
    (torch.distributed._symmetric_memory: every rank maps every peer's buffer), and the last local pass before the
    exchange stores each run of its output axis straight into the owning GPU's buffer over NVLink
    (dsp_dct_set_output_segments): forward, the h-pass of frame d writes rows [g H/G, (g+1) H/G) into rank g's
    [D][HW/G] array; inverse, the temporal pass writes frames [g D/G, (g+1) D/G) into rank g's slab.  No pack kernel,
    no separate collective: one device-side barrier before and after.  Needs H % G == 0, float, CUDA tensors.
"""
import ctypes
import os

import numpy as np

import torch
import torch.distributed as dist

from . import capi
from .plan import Plan


def _ptr(t):
    return t.data_ptr()


def motion_params(dims, damp=1.0, boost=1.0, bandpass=None, threshold=(0.0, 0.0), quant=0.0, preserve_dc=0,
                  float_pixels=False):
    """dsp_motion_params of a full-volume block (`motion -b 0x0x0`: block == scaled == the plane volume)."""
    mp = capi.MotionParams()
    bp = bandpass if bandpass is not None else ((0, 0, 0), tuple(dims))
    for i in range(3):
        mp.block[i] = mp.scaled[i] = int(dims[i])
        mp.bp_begin[i], mp.bp_end[i] = int(bp[0][i]), int(bp[1][i])
    mp.float_pixels = int(bool(float_pixels))
    mp.damp, mp.boost = float(damp), float(boost)
    mp.threshold_min, mp.threshold_max = float(threshold[0]), float(threshold[1])
    mp.quant = float(quant)
    mp.preserve_dc = int(preserve_dc)
    return mp


class Dist3D:
    def __init__(self, D, H, W, prec="f", group=None, lib=None, exchange="auto", device=None, motion=None):
        self.D, self.H, self.W = int(D), int(H), int(W)
        self.prec = prec
        self.lib = lib if lib is not None else capi.load()
        self.tdt = torch.float32 if prec == "f" else torch.float64
        self.group = group
        self.G = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.G > 1 else 0
        G = self.G
        if self.D % G or (self.H * self.W) % G:
            raise ValueError("frames (%d) and h*w (%d) must divide by the number of ranks (%d)" % (D, H * W, G))
        self.Dl, self.Pl = self.D // G, (self.H * self.W) // G
        kw = dict(lib=self.lib)
        if G == 1:
            self.fwd3 = Plan(prec, [D, H, W], [capi.REDFT10] * 3, **kw)
            self.inv3 = Plan(prec, [D, H, W], [capi.REDFT01] * 3, **kw)
        else:
            hw = self.H * self.W
            self.fwd2 = Plan(prec, [H, W], [capi.REDFT10] * 2, 1, None, 1, 0, None, 1, 0, self.Dl, hw, hw, **kw)
            self.inv2 = Plan(prec, [H, W], [capi.REDFT01] * 2, 1, None, 1, 0, None, 1, 0, self.Dl, hw, hw, **kw)
            # time axis of a [D][Pl] array: stride Pl, Pl adjacent columns
            self.fwdt = Plan(prec, [D], [capi.REDFT10], self.Pl, None, self.Pl, 1, None, self.Pl, 1, **kw)
            self.invt = Plan(prec, [D], [capi.REDFT01], self.Pl, None, self.Pl, 1, None, self.Pl, 1, **kw)
        # motion's pel endpoints and coefficient stages (motion.c:618-624, 617-751, 757-776) fused into the plans either
        # side of the exchange: 8-bit pels in, 8-bit pels out, see process()
        self.motion = motion
        if motion is not None:
            if G == 1:
                # the coefficient stage runs as its own sweep between the plans (dsp_motion_coeff_stage): carried in the
                # temporal pass it cost 2.4 ms on top of the plain pass for the 256 x 1080 x 1920 volume, the sweep costs ~1 ms
                self.fwd3.fuse_pel_load(bool(motion.float_pixels))
                self.inv3.fuse_pel_store(motion)
            else:
                # (the coefficient stages run as a sweep over this rank's [D][Pl] coefficients with flat (y, x) coordinates,
                # dsp_motion_coeff_stage_flat in process(); set DSP_DIST_FUSE_COEFF=1 to carry them in the temporal pass's store)
                self.fwd2.fuse_pel_load(bool(motion.float_pixels))
                self.fuse_coeff = bool(os.environ.get("DSP_DIST_FUSE_COEFF"))
                if self.fuse_coeff:
                    self.fwdt.fuse_motion_coeff(motion, None, W, self.rank * self.Pl)
                self.inv2.fuse_pel_store(motion)
        self.a2a_bytes = 0
        self.mode = "nccl"
        if G > 1 and exchange in ("auto", "peer"):
            try:
                self._setup_peer(device)
                self.mode = "peer"
            except Exception as e:                      # no symmetric memory / layout not eligible: keep the NCCL path
                # a partial setup must not survive: a plan that kept its output segments would store into the
                # (now unused) peer buffers on the NCCL path
                for p in (self.fwd2, self.invt):
                    try:
                        p.clear_output_segments()
                    except Exception:
                        pass
                self.cols_buf = self.slab_buf = self.h_cols = self.h_slab = None
                if exchange == "peer":
                    raise
                self.peer_error = repr(e)

    def _setup_peer(self, device):
        """Symmetric receive buffers + segmented output of the two passes that precede an exchange."""
        import torch.distributed._symmetric_memory as symm_mem
        if self.prec != "f" or self.H % self.G:
            raise ValueError("peer exchange needs float data and H divisible by the number of ranks")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        G, D, Dl, Pl, H, W = self.G, self.D, self.Dl, self.Pl, self.H, self.W
        grp = self.group if self.group is not None else dist.group.WORLD
        self.cols_buf = symm_mem.empty((D * Pl,), dtype=self.tdt, device=device)       # [D][Pl]
        self.slab_buf = symm_mem.empty((Dl * H * W,), dtype=self.tdt, device=device)   # [Dl][H][W]
        self.h_cols = symm_mem.rendezvous(self.cols_buf, grp)
        self.h_slab = symm_mem.rendezvous(self.slab_buf, grp)
        es = self.cols_buf.element_size()
        Hg = H // G
        # forward: frame dl of this rank, rows [g Hg, (g+1) Hg) -> rank g's cols[rank Dl + dl][(h - g Hg) W + w]
        self.fwd2.set_output_segments(Hg, [self.h_cols.buffer_ptrs[g] + self.rank * Dl * Pl * es for g in range(G)],
                                      outer_stride=Pl)
        # inverse: temporal output frame d in [g Dl, (g+1) Dl) -> rank g's slab[d - g Dl][rank Pl + col]
        self.invt.set_output_segments(Dl, [self.h_slab.buffer_ptrs[g] + self.rank * Pl * es for g in range(G)],
                                      row_stride=H * W)

    # -- exchange ------------------------------------------------------------------------------------------------
    def _all_to_all(self, out, inp):
        """out[g] <- rank g's inp[self.rank]; inp/out are [G][...] contiguous."""
        try:
            dist.all_to_all_single(out, inp, group=self.group)
        except (RuntimeError, NotImplementedError):
            # gloo (CPU test suite) has no all_to_all: pairwise exchange
            reqs = []
            out[self.rank].copy_(inp[self.rank])
            for g in range(self.G):
                if g != self.rank:
                    reqs.append(dist.isend(inp[g].contiguous(), g, group=self.group))
                    reqs.append(dist.irecv(out[g], g, group=self.group))
            for r in reqs:
                r.wait()
        self.a2a_bytes += inp.numel() * inp.element_size() * (self.G - 1) // self.G

    def _stream(self, t):
        return torch.cuda.current_stream().cuda_stream if t.is_cuda else None

    # -- transforms ------------------------------------------------------------------------------------------------
    def process(self, pels, out=None, work=None):
        """motion -b 0x0x0 on this rank's frames (needs motion=...): pels [D/G][H][W] uint8 (float32 with float pels)
        -> forward 3-D DCT -> coefficient stages -> inverse -> pels of the same type.  `work`: float [D/G][H][W] scratch."""
        assert self.motion is not None
        t = torch
        pel_dt = t.float32 if self.motion.float_pixels else t.uint8
        assert pels.dtype == pel_dt and pels.is_contiguous() and tuple(pels.shape) == (self.Dl, self.H, self.W)
        if out is None:
            out = t.empty_like(pels)
        if work is None:
            work = t.empty((self.Dl, self.H, self.W), dtype=self.tdt, device=pels.device)
        st = self._stream(pels)
        if self.G == 1:
            self.fwd3.execute_dev(_ptr(pels), _ptr(work), st)
            if self.lib.dsp_motion_coeff_stage(self.prec.encode(), ctypes.byref(self.motion), _ptr(work), None, st) != 0:
                raise capi.DspDctError(capi.last_error(self.lib))
            self.inv3.execute_dev(_ptr(work), _ptr(out), st)
            return out
        coeffs = self._forward_from(pels, work, st)
        if not self.fuse_coeff:
            if self.lib.dsp_motion_coeff_stage_flat(self.prec.encode(), ctypes.byref(self.motion), _ptr(coeffs), self.D, self.Pl, self.W,
                                                    self.rank * self.Pl, None, st) != 0:
                raise capi.DspDctError(capi.last_error(self.lib))
        slab = self._inverse_to_slab(coeffs, st)
        self.inv2.execute_dev(_ptr(slab), _ptr(out), st)
        return out

    def forward(self, slab):
        """slab: [D/G][H][W] local frames (modified in place as scratch).  Returns the coefficients this rank owns:
        G == 1: [D][H][W];  G > 1: [D][HW/G] -- all temporal frequencies of its slice of flattened (h, w)."""
        assert slab.dtype == self.tdt and slab.is_contiguous()
        st = self._stream(slab)
        if self.G == 1:
            self.fwd3.execute_dev(_ptr(slab), _ptr(slab), st)
            return slab
        return self._forward_from(slab, slab, st)

    def _forward_from(self, src, slab, st):
        """frames `src` (same tensor as `slab`, or 8-bit pels with a fused pel load) -> coefficients [D][HW/G]"""
        G, Dl, Pl = self.G, self.Dl, self.Pl
        if self.mode == "peer":
            # (the returned array is the plan-owned symmetric buffer: valid until the next forward())
            self.h_cols.barrier(channel=0)                      # every peer is done with its previous cols
            self.fwd2.execute_dev(_ptr(src), _ptr(slab), st)    # last pass stores into the peers' cols
            self.h_cols.barrier(channel=0)                      # all runs have landed
            cols = self.cols_buf.view(self.D, Pl)
            self.fwdt.execute_dev(_ptr(cols), _ptr(cols), st)
            self.a2a_bytes += slab.numel() * slab.element_size() * (G - 1) // G
            return cols
        self.fwd2.execute_dev(_ptr(src), _ptr(slab), st)
        send = slab.view(Dl, G, Pl).permute(1, 0, 2).contiguous()          # [G][Dl][Pl]
        recv = torch.empty_like(send)
        self._all_to_all(recv, send)
        cols = recv.view(self.D, Pl)                                        # frames in global order
        self.fwdt.execute_dev(_ptr(cols), _ptr(cols), st)
        return cols

    def inverse(self, coeffs):
        """Inverse of forward() (unnormalised: returns 8 D H W times the original)."""
        assert coeffs.dtype == self.tdt and coeffs.is_contiguous()
        st = self._stream(coeffs)
        if self.G == 1:
            self.inv3.execute_dev(_ptr(coeffs), _ptr(coeffs), st)
            return coeffs
        slab = self._inverse_to_slab(coeffs, st)
        self.inv2.execute_dev(_ptr(slab), _ptr(slab), st)
        return slab

    def _inverse_to_slab(self, coeffs, st):
        """temporal inverse + exchange: coefficients [D][HW/G] -> this rank's frames [D/G][H][W], spatial axes still
        in the frequency domain"""
        G, Dl, Pl = self.G, self.Dl, self.Pl
        if self.mode == "peer":
            self.h_slab.barrier(channel=1)
            self.invt.execute_dev(_ptr(coeffs), _ptr(coeffs), st)   # stores frames into the owning ranks' slabs
            self.h_slab.barrier(channel=1)
            self.a2a_bytes += coeffs.numel() * coeffs.element_size() * (G - 1) // G
            return self.slab_buf.view(Dl, self.H, self.W)
        self.invt.execute_dev(_ptr(coeffs), _ptr(coeffs), st)
        send = coeffs.view(G, Dl, Pl)                                       # chunk g = frames of rank g
        recv = torch.empty_like(send)
        self._all_to_all(recv, send.contiguous())
        return recv.permute(1, 0, 2).contiguous().view(Dl, self.H, self.W)

    def destroy(self):
        for n in ("fwd3", "inv3", "fwd2", "inv2", "fwdt", "invt"):
            p = getattr(self, n, None)
            if p is not None:
                p.destroy()
