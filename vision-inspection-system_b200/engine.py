"""The device engine: owns tables on one GPU and launches the kernels of libvis_b200.so.

PyTorch is plumbing only (device memory, streams); all compute is in the C-ABI library.  One ``Engine`` per
process/GPU.  No CPU fallback: constructing an Engine without CUDA raises.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import threading
from dataclasses import dataclass

import numpy as np
import torch

from . import _native as N
from . import geometry as G
from . import tables as T


def _stream_ptr(stream=None) -> C.c_void_p:
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


@dataclass
class _Geometry:
    """Everything that depends only on (src_h, src_w, dst_h, dst_w) — cached per engine."""
    src_h: int
    src_w: int
    dst_h: int
    dst_w: int
    kt: int                      # tap class of the general fused kernel (6 / 8); 0: scheduled kernel or generic passes only
    hrec: torch.Tensor | None
    vrec: torch.Tensor | None
    htable: T.CoeffTable
    vtable: T.CoeffTable
    n_rows: int
    plans: dict                  # vsplit -> StripPlan
    scheds: dict                 # (pitch, n_segs) -> VisSched bytes, or None when the scheduled kernel declines
    srec: list                   # [hrec, vrec] device records of the scheduled kernel (packed on first use)


@dataclass
class _FusedLaunch:
    kt: int
    n_frames: int
    n_strips: int
    span_bytes: int
    strip_w: int
    frames: torch.Tensor         # device VisFrame[]
    strips: torch.Tensor         # device VisStrip[]


@dataclass
class _SchedLaunch:
    """One geometry, statically scheduled kernel (vis_preprocess_fused_sched)."""
    sched: np.ndarray            # the VisSched parameter block (host bytes, opaque)
    n_frames: int
    n_items: int
    frames: torch.Tensor         # device VisFrameRef[]
    hrec: torch.Tensor
    vrec: torch.Tensor
    dup: torch.Tensor | None = None   # device int64[n_frames]: first row of each frame's second copy, -1 for none


@dataclass
class BatchPlan:
    total_rows: int
    grid_thw: torch.Tensor       # int64 [B, 3] (host)
    fused: list                  # _FusedLaunch (general kernel, one per tap class) and _SchedLaunch (one per geometry)
    generic: list                # (frame index, geometry, first output row)


class Engine:
    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("vision-inspection-system_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.L = N.lib()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.lut = torch.from_numpy(T.normalize_lut()).to(self.device)
        self._dev_tables: dict = {}
        self._geoms: dict = {}
        self._batch_plans: dict = {}
        self._staging: dict = {}
        self.sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.last_launches = 0           # kernels launched by the most recent public call
        # 9+ tap geometries: the packed-byte (IDP.4A) kernel unless VIS_B200_DP4A=0 selects the 16-slot IMAD kernel
        # (developer A/B switch; both are bit-exact)
        self.dp4a = os.environ.get("VIS_B200_DP4A", "1") != "0"
        # ... and, on top of it, both passes as banded u8 x limb matrix products on the integer tensor path (IMMA.16832,
        # vis_fused_mma.cu) unless VIS_B200_MMA=0 keeps the IDP.4A kernel; the three families are bit-exact
        self.mma = os.environ.get("VIS_B200_MMA", "1") != "0"
        # preprocess_dual: CUDA streams its independent launches are spread over (1 = one after the other on the caller's)
        self.dual_streams = max(1, int(os.environ.get("VIS_B200_DUAL_STREAMS", "4")))
        self._lock = threading.RLock()   # public entry points are serialised: caches and staging buffers are shared
        self._tls = threading.local()    # per-thread state (nvJPEG handles are not thread-safe)

    def _sched_flags(self) -> int:
        return (N.SCHED_FLAG_MMA if self.mma and self.dp4a else 0) | (N.SCHED_FLAG_DP4A if self.dp4a else 0)

    def use_dp4a(self, on: bool, mma: bool | None = None) -> None:
        """Select the kernel family of 9+ tap geometries — byte limbs (default: on the integer tensor path, ``mma=False``:
        IDP.4A) or the 16-slot IMAD kernel — and drop every cached schedule and plan (A/B runs and the tests that keep
        all families green)."""
        with self._lock:
            self.dp4a = bool(on)
            if mma is not None:
                self.mma = bool(mma)
            self._geoms.clear()
            self._dev_tables.clear()
            self._batch_plans.clear()
            self.__dict__.pop("_dual_plans", None)

    # ------------------------------------------------------------------ tables
    def _device_table(self, in_size: int, out_size: int, filt: int):
        key = (in_size, out_size, filt)
        hit = self._dev_tables.get(key)
        if hit is None:
            t = T.coeff_table(in_size, out_size, filt)
            hit = (t, torch.from_numpy(t.k.copy()).to(self.device), torch.from_numpy(t.bounds.copy()).to(self.device))
            self._dev_tables[key] = hit
        return hit

    def _geometry(self, src_h: int, src_w: int, dst_h: int, dst_w: int) -> _Geometry:
        key = (src_h, src_w, dst_h, dst_w)
        g = self._geoms.get(key)
        if g is None:
            ht = T.coeff_table(src_w, dst_w, N.FILTER_BICUBIC)
            vt = T.coeff_table(src_h, dst_h, N.FILTER_BICUBIC)
            kt = T.kt_class(max(ht.max_taps, vt.max_taps))     # class of the GENERAL fused kernel (6 / 8), 0 beyond
            hrec = vrec = None
            if kt:
                hrec = torch.from_numpy(T.pack_records(ht, kt)).to(self.device)
                vrec = torch.from_numpy(T.pack_records(vt, kt)).to(self.device)
            g = _Geometry(src_h, src_w, dst_h, dst_w, kt, hrec, vrec, ht, vt,
                          (dst_h // G.PATCH_SIZE) * (dst_w // G.PATCH_SIZE), {}, {}, [])
            self._geoms[key] = g
        return g

    def _plan(self, g: _Geometry, vsplit: int) -> T.StripPlan:
        p = g.plans.get(vsplit)
        if p is None:
            p = T.plan_strips(g.dst_h, g.dst_w, g.htable, g.kt, vsplit)
            g.plans[vsplit] = p
        return p

    def _pick_segs(self, items_per_seg: int, src_h: int, max_segs: int) -> int:
        """Row segments per frame for a scheduled launch.  The persistent kernels hand items (strip x segment) to the
        CTAs round-robin, so the launch takes ceil(items / SMs) rounds: 448 items on 148 SMs run 4 rounds at 76 %
        occupancy of the machine.  Pick the count that maximises rounds-efficiency times the cost of a segment start
        (its first rows are read again by the segment above and the role pipeline refills: ~1.3 % of a 2160-row frame)."""
        best, best_score = 1, 0.0
        for segs in range(1, max(1, min(16, max_segs)) + 1):
            rounds = items_per_seg * segs / self.sm_count
            score = rounds / -(-items_per_seg * segs // self.sm_count) * src_h / (src_h + (segs - 1) * 28.0)
            if score > best_score * 1.01:                 # prefer fewer segments unless the gain is real
                best, best_score = segs, score
        return best

    def _sched_records(self, head, tables) -> list:
        """Device records of a scheduled launch for the (horizontal, vertical) coefficient tables, in the format of the
        kernel the schedule names: packed byte limbs (dp_words > 0) or int32 coefficients."""
        kt, per_index, words = int(head["kt"]), int(head["per_index"]), int(head["dp_words"])
        out = []
        for t in tables:
            if words and int(head["mma_ks"]):
                stride = N.check(self.L.vis_sched_record_stride_mma(words), "vis_sched_record_stride_mma")
                rec = np.zeros((t.out_size + 1, stride), np.int32)
                N.check(self.L.vis_sched_pack_records_mma(t.out_size, N.i32ptr(t.k), N.i32ptr(t.bounds), t.ksize, words,
                                                          N.i32ptr(rec), rec.size), "vis_sched_pack_records_mma")
            elif words:
                stride = N.check(self.L.vis_sched_record_stride_dp(words), "vis_sched_record_stride_dp")
                rec = np.zeros((t.out_size + 1, stride), np.int32)
                N.check(self.L.vis_sched_pack_records_dp(t.out_size, N.i32ptr(t.k), N.i32ptr(t.bounds), t.ksize, words,
                                                         N.i32ptr(rec), rec.size), "vis_sched_pack_records_dp")
            else:
                stride = self.L.vis_record_stride(kt)
                rec = np.zeros((t.out_size + 1, stride), np.int32)
                N.check(self.L.vis_sched_pack_records(t.out_size, N.i32ptr(t.k), N.i32ptr(t.bounds), t.ksize, kt, per_index,
                                                      N.i32ptr(rec), rec.size), "vis_sched_pack_records")
            out.append(torch.from_numpy(rec).to(self.device))
        return out

    def _sched(self, g: _Geometry, pitch: int, n_segs: int):
        """VisSched parameter block for (geometry, row pitch, row segments), or None if the geometry needs the
        general kernel (upscaling, > 8 taps, schedule too large)."""
        key = (pitch, n_segs)
        if key not in g.scheds:
            buf = np.zeros(self.L.vis_sched_sizeof(), np.uint8)
            mode = N.SCHED_OUT_PIXEL_VALUES | self._sched_flags()
            rc = self.L.vis_sched_build(g.src_h, g.src_w, g.dst_h, g.dst_w, pitch, N.i32ptr(g.htable.bounds),
                                        N.i32ptr(g.vtable.bounds), n_segs, mode, buf.ctypes.data_as(C.c_void_p))
            if rc == N.VIS_E_UNSUPPORTED:
                buf = None
            else:
                N.check(rc, "vis_sched_build")
                if not g.srec:
                    head = np.frombuffer(buf[:N.SCHED_HEAD_DTYPE.itemsize], N.SCHED_HEAD_DTYPE)[0]
                    g.srec.extend(self._sched_records(head, (g.htable, g.vtable)))
            g.scheds[key] = buf
        return g.scheds[key]

    # ------------------------------------------------------------------ fused resample (uint8 -> uint8), RGB, <= 16 taps
    def _resize_sched(self, src_h: int, src_w: int, out_h: int, out_w: int, filt: int, pitch: int, n_segs: int):
        """(VisSched bytes, hrec, vrec) of the fused uint8 resize for this geometry, or None when it needs the generic
        passes (> 16 taps, upscaling, width not a multiple of 4, vertical-first order)."""
        key = ("rs", src_h, src_w, out_h, out_w, filt, pitch, n_segs)
        hit = self._dev_tables.get(key, False)
        if hit is not False:
            return hit
        hit = None
        if G.pil_pass_order(src_h, src_w, out_h, out_w) == "hv":
            ht, vt = T.coeff_table(src_w, out_w, filt), T.coeff_table(src_h, out_h, filt)
            buf = np.zeros(self.L.vis_sched_sizeof(), np.uint8)
            mode = N.SCHED_OUT_U8 | self._sched_flags()
            rc = self.L.vis_sched_build(src_h, src_w, out_h, out_w, pitch, N.i32ptr(ht.bounds), N.i32ptr(vt.bounds), n_segs,
                                        mode, buf.ctypes.data_as(C.c_void_p))
            if rc != N.VIS_E_UNSUPPORTED:
                N.check(rc, "vis_sched_build")
                rkey = ("rsrec", src_h, src_w, out_h, out_w, filt)
                recs = self._dev_tables.get(rkey)
                if recs is None:
                    head = np.frombuffer(buf[:N.SCHED_HEAD_DTYPE.itemsize], N.SCHED_HEAD_DTYPE)[0]
                    recs = self._dev_tables[rkey] = self._sched_records(head, (ht, vt))
                hit = (buf, recs[0], recs[1])
        self._dev_tables[key] = hit
        return hit

    def _resize_plan(self, frames: list, out_h: int, out_w: int, filt: int):
        """Everything ONE fused ``vis_resize_fused_sched`` launch over same-shape RGB frames needs — (schedule, records,
        device descriptors, the [n, out_h, out_w, 3] result tensor) — or None when the geometry needs the generic passes.
        The plan stays valid while the frames keep their addresses (``preprocess_dual`` caches it)."""
        f0 = frames[0]
        fusable = all(f.dim() == 3 and f.shape == f0.shape and f.shape[2] == 3 and f.stride(2) == 1 and f.stride(1) == 3
                      and f.stride(0) == f0.stride(0) and f.stride(0) % 16 == 0 and f.data_ptr() % 16 == 0 for f in frames)
        if not fusable:
            return None
        for f in frames:
            self._check_u8(f)
        h, w = int(f0.shape[0]), int(f0.shape[1])
        head = self._resize_sched(h, w, out_h, out_w, filt, int(f0.stride(0)), 1)
        if head is None:
            return None
        per = int(np.frombuffer(head[0][:N.SCHED_HEAD_DTYPE.itemsize], N.SCHED_HEAD_DTYPE)[0]["n_strips"])
        segs = self._pick_segs(len(frames) * per, h, out_h // 14)
        sched, hrec, vrec = self._resize_sched(h, w, out_h, out_w, filt, int(f0.stride(0)), segs) or head
        out = torch.empty((len(frames), out_h, out_w, 3), dtype=torch.uint8, device=self.device)
        ref = np.zeros(len(frames), N.RESIZE_REF_DTYPE)
        ref["src"] = [f.data_ptr() for f in frames]
        ref["dst"] = out.data_ptr() + np.arange(len(frames), dtype=np.uint64) * np.uint64(out.stride(0))
        d_ref = torch.from_numpy(ref.view(np.uint8).copy()).to(self.device)
        return (sched, hrec, vrec, d_ref, len(frames), out_w * 3, out)

    def _run_resize(self, rp) -> None:
        sched, hrec, vrec, d_ref, n, dst_pitch, _ = rp
        N.check(self.L.vis_resize_fused_sched(sched.ctypes.data_as(C.c_void_p), d_ref.data_ptr(), n, dst_pitch,
                                              hrec.data_ptr(), vrec.data_ptr(), _stream_ptr()), "vis_resize_fused_sched")

    def resize_batch_u8(self, frames, out_h: int, out_w: int, filt: int = N.FILTER_LANCZOS) -> list:
        """``Image.resize((out_w, out_h), filt)`` of a list of same-shape RGB uint8 HWC CUDA frames: ONE fused launch
        (both passes) when the geometry allows, else the generic passes frame by frame.  Returns new tensors."""
        frames = list(frames)
        if not frames:
            return []
        rp = self._resize_plan(frames, out_h, out_w, filt)
        if rp is None:
            outs = []
            launches = 0
            for f in frames:
                outs.append(self.resize_u8(f, out_h, out_w, filt, fused=False))
                launches += self.last_launches
            self.last_launches = launches
            return outs
        self._run_resize(rp)
        self.last_launches = 1
        self._keepalive_r = rp[3]
        return list(rp[6].unbind(0))

    # ------------------------------------------------------------------ generic resample (uint8 -> uint8)
    def resize_u8(self, img: torch.Tensor, out_h: int, out_w: int, filt: int = N.FILTER_LANCZOS, stream=None,
                  fused: bool = True) -> torch.Tensor:
        """``PIL.Image.resize((out_w, out_h), filt, reducing_gap=None)`` for a CUDA uint8 HWC (or HW) tensor.

        RGB frames whose geometry the fused scheduled kernel takes (<= 16 taps, no upscaling) go through it (one launch,
        both passes); everything else through the two generic passes (``fused=False`` forces those)."""
        self._check_u8(img)
        if fused and stream is None and img.dim() == 3 and img.shape[2] == 3:
            return self.resize_batch_u8([img], out_h, out_w, filt)[0]
        squeeze = img.dim() == 2
        if squeeze:
            img = img.unsqueeze(-1)
        if img.stride(2) != 1 or img.stride(1) != img.shape[2]:
            img = img.contiguous()
        h, w, ch = img.shape
        order = G.pil_pass_order(h, w, out_h, out_w)
        sp = _stream_ptr(stream)
        cur, cur_h, cur_w = img, h, w
        self.last_launches = 0
        if order == "":
            cur = img.clone()
        for axis in order:
            if axis == "h":
                t, k, b = self._device_table(cur_w, out_w, filt)
                dst = torch.empty((cur_h, out_w, ch), dtype=torch.uint8, device=self.device)
                N.check(self.L.vis_resample_h_u8(cur.data_ptr(), cur.stride(0), cur_h, cur_w, ch,
                                                 dst.data_ptr(), dst.stride(0), out_w,
                                                 k.data_ptr(), b.data_ptr(), t.ksize, sp), "vis_resample_h_u8")
                cur, cur_w = dst, out_w
            else:
                t, k, b = self._device_table(cur_h, out_h, filt)
                dst = torch.empty((out_h, cur_w, ch), dtype=torch.uint8, device=self.device)
                N.check(self.L.vis_resample_v_u8(cur.data_ptr(), cur.stride(0), cur_h, cur_w * ch,
                                                 dst.data_ptr(), dst.stride(0), out_h,
                                                 k.data_ptr(), b.data_ptr(), t.ksize, sp), "vis_resample_v_u8")
                cur, cur_h = dst, out_h
            self.last_launches += 1
        return cur.squeeze(-1) if squeeze else cur

    def resize_hp(self, img: torch.Tensor, out_h: int, out_w: int, filt: int, kind: int) -> torch.Tensor:
        """``Image.resize((out_w, out_h), filt)`` for the single-channel modes Pillow resamples in double precision:
        ``img`` is the raw image memory as a CUDA uint8 tensor ``[H, W * bytes_per_pixel]`` (``kind``: N.HP_U16LE /
        HP_U16BE for "I;16" / "I;16B", HP_I32 for "I", HP_F32 for "F").  Same pass order as the 8-bit path (horizontal
        first; vertical first for frames more than 100x taller than wide), bit-exact with Pillow."""
        self._check_u8(img)
        bpp = 2 if kind in (N.HP_U16LE, N.HP_U16BE) else 4
        if img.dim() != 2 or img.stride(1) != 1 or img.shape[1] % bpp:
            raise ValueError("expected the raw image memory as a [H, W * bytes_per_pixel] uint8 tensor")
        h, w = int(img.shape[0]), int(img.shape[1]) // bpp
        sp = _stream_ptr()
        cur, cur_h, cur_w = img, h, w
        self.last_launches = 0
        order = G.pil_pass_order(h, w, out_h, out_w)
        if order == "":
            return img.clone()
        keep = []
        for axis in order:
            vertical = axis == "v"
            k, b = T.coeff_table_f64(cur_h if vertical else cur_w, out_h if vertical else out_w, filt)
            dk, db = torch.from_numpy(k).to(self.device), torch.from_numpy(b).to(self.device)
            nh, nw = (out_h, cur_w) if vertical else (cur_h, out_w)
            dst = torch.empty((nh, nw * bpp), dtype=torch.uint8, device=self.device)
            N.check(self.L.vis_resample_hp(cur.data_ptr(), cur.stride(0), cur_h, cur_w, kind, 1 if vertical else 0,
                                           dst.data_ptr(), dst.stride(0), out_h if vertical else out_w, dk.data_ptr(),
                                           db.data_ptr(), k.shape[1], sp), "vis_resample_hp")
            keep += [dk, db, cur]
            cur, cur_h, cur_w = dst, nh, nw
            self.last_launches += 1
        self._keepalive_hp = keep
        return cur

    def reduce_u8(self, img: torch.Tensor, factor, box=None) -> torch.Tensor:
        """``PIL.Image.reduce(factor, box)`` for a CUDA uint8 HWC tensor: integer box average (libImaging/Reduce.c)."""
        self._check_u8(img)
        if img.dim() != 3 or img.stride(2) != 1 or img.stride(1) != img.shape[2]:
            raise ValueError("expected a [H, W, C] uint8 tensor with contiguous pixels")
        h, w, ch = (int(v) for v in img.shape)
        fx, fy = (int(v) for v in (factor if isinstance(factor, (tuple, list)) else (factor, factor)))
        x0, y0, x1, y1 = (0, 0, w, h) if box is None else (int(v) for v in box)
        out = torch.empty((-(-(y1 - y0) // fy), -(-(x1 - x0) // fx), ch), dtype=torch.uint8, device=self.device)
        N.check(self.L.vis_reduce_u8(img.data_ptr(), img.stride(0), h, w, ch, fx, fy, x0, y0, x1, y1,
                                     out.data_ptr(), out.stride(0), _stream_ptr()), "vis_reduce_u8")
        self.last_launches = 1
        return out

    def resize_box_u8(self, img: torch.Tensor, out_h: int, out_w: int, filt: int, box) -> torch.Tensor:
        """``Image.resize((out_w, out_h), filt, box=box, reducing_gap=None)`` for a CUDA uint8 HWC tensor: the core
        ImagingResample with a fractional source box — horizontal pass over the rows the vertical pass reads, then
        the vertical pass (generic kernels, boxed coefficient tables)."""
        self._check_u8(img)
        if img.dim() != 3 or img.stride(2) != 1 or img.stride(1) != img.shape[2]:
            raise ValueError("expected a [H, W, C] uint8 tensor with contiguous pixels")
        h, w, ch = (int(v) for v in img.shape)
        box = tuple(float(np.float32(v)) for v in box)                  # ImagingResample takes float box[4]
        if h > w * 100 and out_h < h:                                   # tall image: vertical pass first (PIL:Image.py:2431-2435)
            tmp = self.resize_box_u8(img, out_h, w, filt, (0.0, box[1], float(w), box[3]))
            return self.resize_box_u8(tmp, out_h, out_w, filt, (box[0], 0.0, box[2], float(out_h)))
        need_h = out_w != w or box[0] != 0.0 or box[2] != float(out_w)
        need_v = out_h != h or box[1] != 0.0 or box[3] != float(out_h)
        sp = _stream_ptr()
        self.last_launches = 0
        vt = T.coeff_table_box(h, box[1], box[3], out_h, filt)
        y_first = int(vt.bounds[0, 0])
        y_last = int(vt.bounds[-1, 0] + vt.bounds[-1, 1])
        cur, row_off = img, 0
        if need_h:
            ht = T.coeff_table_box(w, box[0], box[2], out_w, filt)
            k, b = torch.from_numpy(ht.k).to(self.device), torch.from_numpy(ht.bounds).to(self.device)
            rows = y_last - y_first
            dst = torch.empty((rows, out_w, ch), dtype=torch.uint8, device=self.device)
            N.check(self.L.vis_resample_h_u8(img.data_ptr() + y_first * img.stride(0), img.stride(0), rows, w, ch,
                                             dst.data_ptr(), dst.stride(0), out_w, k.data_ptr(), b.data_ptr(), ht.ksize, sp),
                    "vis_resample_h_u8")
            cur, row_off = dst, y_first
            self.last_launches += 1
        if not need_v:
            return cur.clone() if cur is img else cur
        bounds = vt.bounds.copy()
        bounds[:, 0] -= row_off
        k, b = torch.from_numpy(vt.k).to(self.device), torch.from_numpy(bounds).to(self.device)
        cur_h, cur_w = int(cur.shape[0]), int(cur.shape[1])
        dst = torch.empty((out_h, cur_w, ch), dtype=torch.uint8, device=self.device)
        N.check(self.L.vis_resample_v_u8(cur.data_ptr(), cur.stride(0), cur_h, cur_w * ch, dst.data_ptr(), dst.stride(0),
                                         out_h, k.data_ptr(), b.data_ptr(), vt.ksize, sp), "vis_resample_v_u8")
        self.last_launches += 1
        self._keepalive_b = (k, b)
        return dst

    def alpha_premultiply_(self, img: torch.Tensor, forward: bool = True) -> torch.Tensor:
        """In place: premultiply (``forward``) or un-premultiply the colour channels of an RGBA / LA CUDA uint8 HWC
        tensor by its last channel, with Pillow's integer formulas (Convert.c rgbA2rgba / rgba2rgbA)."""
        self._check_u8(img)
        if img.dim() != 3 or img.shape[2] not in (2, 4) or img.stride(2) != 1 or img.stride(1) != img.shape[2]:
            raise ValueError("expected a [H, W, 2|4] uint8 tensor with contiguous pixels")
        N.check(self.L.vis_alpha_premultiply_u8(img.data_ptr(), img.stride(0), int(img.shape[0]), int(img.shape[1]),
                                                int(img.shape[2]), 1 if forward else 0, _stream_ptr()),
                "vis_alpha_premultiply_u8")
        return img

    def resize_nearest_u8(self, img: torch.Tensor, out_h: int, out_w: int, box=None) -> torch.Tensor:
        """``Image.resize((out_w, out_h), NEAREST, box)`` for a CUDA uint8 HW / HWC tensor — Pillow's path for palette
        and bilevel images whatever filter is named (index tables on the host, one gather launch)."""
        self._check_u8(img)
        squeeze = img.dim() == 2
        if squeeze:
            img = img.unsqueeze(-1)
        if img.stride(2) != 1 or img.stride(1) != img.shape[2]:
            img = img.contiguous()
        h, w, ch = (int(v) for v in img.shape)
        box = (0.0, 0.0, float(w), float(h)) if box is None else tuple(float(np.float32(v)) for v in box)
        xt, yt = np.empty(out_w, np.int32), np.empty(out_h, np.int32)
        N.check(self.L.vis_nearest_table(w, box[0], box[2], out_w, N.i32ptr(xt)), "vis_nearest_table")
        N.check(self.L.vis_nearest_table(h, box[1], box[3], out_h, N.i32ptr(yt)), "vis_nearest_table")
        d_xt, d_yt = torch.from_numpy(xt).to(self.device), torch.from_numpy(yt).to(self.device)
        out = torch.empty((out_h, out_w, ch), dtype=torch.uint8, device=self.device)
        N.check(self.L.vis_gather_u8(img.data_ptr(), img.stride(0), h, w, ch, out.data_ptr(), out.stride(0), out_h, out_w,
                                     d_xt.data_ptr(), d_yt.data_ptr(), _stream_ptr()), "vis_gather_u8")
        self.last_launches = 1
        self._keepalive_n = (d_xt, d_yt)
        return out.squeeze(-1) if squeeze else out

    def resize_reducing_u8(self, img: torch.Tensor, out_h: int, out_w: int, filt: int = N.FILTER_LANCZOS, box=None,
                           reducing_gap: float = 2.0) -> torch.Tensor:
        """``Image.resize((out_w, out_h), filt, box, reducing_gap)`` — what ``Image.thumbnail`` calls: an integer
        ``reduce`` pre-pass when the frame is >= 2 * reducing_gap times larger than the target, then the resample over
        the (fractional) box of the reduced frame.  Frames below that ratio take the fused kernels unchanged."""
        h, w = int(img.shape[0]), int(img.shape[1])
        plan = G.reducing_plan(w, h, out_w, out_h, filt, box, reducing_gap)
        if plan is None:
            return self.resize_u8(img, out_h, out_w, filt) if box is None else self.resize_box_u8(img, out_h, out_w, filt, box)
        factor, reduce_box, new_box = plan
        reduced = self.reduce_u8(img if img.stride(1) == img.shape[2] else img.contiguous(), factor, reduce_box)
        out = self.resize_box_u8(reduced, out_h, out_w, filt, new_box)
        self.last_launches += 1
        return out

    def agent_inputs(self, frames, role: str = "inspector", max_size: int | None = None) -> list:
        """The frames an agent's Qwen2-VL processor sees: ``thumbnail((S, S), LANCZOS)`` for every RGB frame larger than
        S (2048 Inspector / 1024 Auditor — src/agents/vlm_inspector.py:63-64, vlm_auditor.py:90-91), the others
        untouched.  Frames of one source geometry share ONE fused launch; frames >= 4x the limit take Pillow's
        reduce pre-pass.  Returns a list aligned with ``frames``."""
        limit = max_size or {"inspector": G.INSPECTOR_MAX_SIZE, "auditor": G.AUDITOR_MAX_SIZE}[role]
        batch = list(frames.unbind(0)) if isinstance(frames, torch.Tensor) and frames.dim() == 4 else list(frames)
        groups: dict = {}
        for i, f in enumerate(batch):
            h, w = int(f.shape[0]), int(f.shape[1])
            if max(h, w) > limit:
                groups.setdefault((h, w, f.stride(0)), []).append(i)
        launches = 0
        for (h, w, _), idx in groups.items():
            tw, th = G.thumbnail_size(w, h, limit)
            if G.reducing_plan(w, h, tw, th, N.FILTER_LANCZOS) is not None:
                outs = []
                for i in idx:
                    outs.append(self.resize_reducing_u8(batch[i], th, tw, N.FILTER_LANCZOS))
                    launches += self.last_launches
            else:
                outs = self.resize_batch_u8([batch[i] for i in idx], th, tw, N.FILTER_LANCZOS)
                launches += self.last_launches
            for i, o in zip(idx, outs):
                batch[i] = o
        self.last_launches = launches
        return batch

    # ------------------------------------------------------------------ frames -> pixel_values
    def plan_batch(self, frames, min_pixels: int = G.DEFAULT_MIN_PIXELS, max_pixels: int = G.DEFAULT_MAX_PIXELS,
                   force_generic: bool = False, vsplit: int | None = None, path: str = "auto",
                   rows=None, dup_rows=None, total_rows: int | None = None) -> "BatchPlan":
        """Host-side planning for one batch: geometry, output rows, descriptor arrays (uploaded once).

        Plans depend only on pointers and shapes, so they are cached and reused when the same device buffers are
        submitted again (a streaming loop that refills one staging buffer pays for planning once).
        ``rows`` (first output row per frame; default: frames back to back), ``dup_rows`` (first row of a second copy
        of a frame's rows, -1 for none) and ``total_rows`` place the frames in a larger tensor: ``preprocess_dual``.
        """
        uniform = isinstance(frames, torch.Tensor) and frames.dim() == 4
        placed = (None if rows is None else tuple(int(v) for v in rows),
                  None if dup_rows is None else tuple(int(v) for v in dup_rows), total_rows)
        if uniform:
            self._check_u8(frames)
            key = ("u", frames.data_ptr(), tuple(frames.shape), tuple(frames.stride()), min_pixels, max_pixels,
                   force_generic, vsplit, path, placed)
        else:
            frames = list(frames)
            for f in frames:
                self._check_u8(f)
            key = ("l", tuple((f.data_ptr(), tuple(f.shape), tuple(f.stride())) for f in frames), min_pixels,
                   max_pixels, force_generic, vsplit, path, placed)
        plan = self._batch_plans.get(key)
        if plan is not None:
            return plan
        if uniform:
            b, h, w, c = (int(v) for v in frames.shape)
            if b == 0:
                raise ValueError("no frames")
            shapes = [(h, w)] * b
            ptrs = frames.data_ptr() + np.arange(b, dtype=np.int64) * frames.stride(0)
            pitches = np.full(b, frames.stride(1), np.int64)
            ok_layout = c == 3 and frames.stride(3) == 1 and frames.stride(2) == 3
        else:
            if not frames:
                raise ValueError("no frames")
            ok_layout = all(f.dim() == 3 and f.shape[2] == 3 and f.stride(2) == 1 and f.stride(1) == 3 for f in frames)
            shapes = [(int(f.shape[0]), int(f.shape[1])) for f in frames] if ok_layout else []
            ptrs = np.array([f.data_ptr() for f in frames], np.int64)
            pitches = np.array([f.stride(0) for f in frames], np.int64)
        if not ok_layout:
            raise ValueError("frames must be [H, W, 3] uint8 with contiguous pixels (row pitch may be padded)")

        geoms, grids = [], []
        cache: dict = {}
        for hw in shapes:
            g = cache.get(hw)
            if g is None:
                dh, dw = G.smart_resize(hw[0], hw[1], G.FACTOR, min_pixels, max_pixels)
                g = cache[hw] = self._geometry(hw[0], hw[1], dh, dw)
            geoms.append(g)
            grids.append(G.grid_thw(g.dst_h, g.dst_w))
        n_rows = np.array([g.n_rows for g in geoms], np.int64)
        if rows is None:
            row0 = np.concatenate([[0], np.cumsum(n_rows)[:-1]]).astype(np.int64)
            total = int(n_rows.sum())
        else:
            row0 = np.asarray(rows, np.int64)
            if row0.shape != n_rows.shape or total_rows is None:
                raise ValueError("rows: one first row per frame, with total_rows")
            total = int(total_rows)
        dup = np.full(len(geoms), -1, np.int64) if dup_rows is None else np.asarray(dup_rows, np.int64)
        if dup.shape != n_rows.shape:
            raise ValueError("dup_rows: one entry per frame")

        plan = BatchPlan(total, torch.tensor(grids, dtype=torch.int64), [], [])
        groups: dict = {}
        for i, g in enumerate(geoms):
            # the fused kernels stage rows with bulk copies: 16-byte aligned base and pitch, horizontal pass first
            aligned = ptrs[i] % 16 == 0 and pitches[i] % 16 == 0 and pitches[i] >= g.src_w * 3
            if not force_generic and aligned and G.pil_pass_order(g.src_h, g.src_w, g.dst_h, g.dst_w) != "vh":
                groups.setdefault(id(g), (g, []))[1].append(i)
            else:
                plan.generic.append((i, g, int(row0[i])))
                if dup[i] >= 0:                       # no shared store on this path: the copy is computed again
                    plan.generic.append((i, g, int(dup[i])))
        want = 3 * self.sm_count                     # work items wanted: a few per SM
        # statically scheduled kernels (8-slot: <= 8 taps, 16-slot: 9..32 taps): one launch per (geometry, pitch)
        by_class: dict = {}
        for g, idx in groups.values():
            idx = np.asarray(idx)
            rest = idx
            if path != "general":
                rest_parts = []
                for pitch in np.unique(pitches[idx]):
                    sel = idx[pitches[idx] == pitch]
                    head = self._sched(g, int(pitch), 1)
                    if head is None:
                        rest_parts.append(sel)
                        continue
                    per = int(np.frombuffer(head[:N.SCHED_HEAD_DTYPE.itemsize], N.SCHED_HEAD_DTYPE)[0]["n_strips"])
                    segs = vsplit if vsplit is not None else self._pick_segs(len(sel) * per, g.src_h, g.dst_h // 28)
                    sched = self._sched(g, int(pitch), max(1, min(segs, g.dst_h // 14)))
                    if sched is None:                # every segment adds chunk-rounded mask bytes: the multi-segment
                        sched = head                 # schedule may not fit where the one-segment schedule did
                    hd = np.frombuffer(sched[:N.SCHED_HEAD_DTYPE.itemsize], N.SCHED_HEAD_DTYPE)[0]
                    src_ptr, first_row, d_dup = ptrs[sel], row0[sel], None
                    if (dup[sel] >= 0).any():
                        if int(hd["ring"]) == 8:     # second destination written from the same registers
                            d_dup = torch.from_numpy(dup[sel].copy()).to(self.device)
                        else:                        # 16-slot kernel: the copy is a second work item
                            extra = sel[dup[sel] >= 0]
                            src_ptr, first_row = np.concatenate([src_ptr, ptrs[extra]]), np.concatenate([first_row, dup[extra]])
                    ref = np.zeros(len(src_ptr), N.FRAME_REF_DTYPE)
                    ref["src"], ref["row0"] = src_ptr.astype(np.uint64), first_row
                    plan.fused.append(_SchedLaunch(
                        sched, len(ref), len(ref) * int(hd["n_strips"]) * int(hd["n_segs"]),
                        torch.from_numpy(ref.view(np.uint8).copy()).to(self.device), g.srec[0], g.srec[1], d_dup))
                rest = np.concatenate(rest_parts) if rest_parts else np.zeros(0, np.int64)
            if len(rest) and g.kt:
                by_class.setdefault(g.kt, []).append((g, rest))
            else:                                    # 9+ taps the schedule declined: generic passes
                for i in rest:
                    plan.generic.append((int(i), g, int(row0[i])))
                    if dup[i] >= 0:
                        plan.generic.append((int(i), g, int(dup[i])))
        # general kernel: one launch per tap class; all geometries of a class share it
        for kt, members in by_class.items():
            n_class = sum(len(idx) for _, idx in members)
            fr_parts, st_parts, span, strip_w, base = [], [], 0, 0, 0
            for g, idx in members:
                if vsplit is None:
                    per_frame = len(self._plan(g, 1).strips)
                    vs = max(1, min(g.dst_h // 56, -(-want // (n_class * per_frame))))
                else:
                    vs = vsplit
                sp = self._plan(g, vs)
                first_row = row0[idx]
                if (dup[idx] >= 0).any():             # general kernel: the copy is a second frame entry
                    extra = idx[dup[idx] >= 0]
                    idx, first_row = np.concatenate([idx, extra]), np.concatenate([first_row, dup[extra]])
                fr = np.zeros(len(idx), N.FRAME_DTYPE)
                fr["src"], fr["src_pitch"] = ptrs[idx].astype(np.uint64), pitches[idx]
                fr["src_h"], fr["src_w"], fr["dst_h"], fr["dst_w"] = g.src_h, g.src_w, g.dst_h, g.dst_w
                fr["hrec"], fr["vrec"], fr["row0"] = g.hrec.data_ptr(), g.vrec.data_ptr(), first_row
                st = np.tile(sp.strips, len(idx))
                st["frame"] = base + np.repeat(np.arange(len(idx), dtype=np.int32), len(sp.strips))
                fr_parts.append(fr)
                st_parts.append(st)
                span, strip_w, base = max(span, sp.span_bytes), max(strip_w, sp.strip_w), base + len(idx)
            fr_all, st_all = np.concatenate(fr_parts), np.concatenate(st_parts)
            plan.fused.append(_FusedLaunch(
                kt, len(fr_all), len(st_all), span, strip_w,
                torch.from_numpy(fr_all.view(np.uint8).copy()).to(self.device),
                torch.from_numpy(st_all.view(np.uint8).copy()).to(self.device)))
        if len(self._batch_plans) >= 64:
            self._batch_plans.pop(next(iter(self._batch_plans)))
        self._batch_plans[key] = plan
        return plan

    def preprocess(self, frames, min_pixels: int = G.DEFAULT_MIN_PIXELS, max_pixels: int = G.DEFAULT_MAX_PIXELS,
                   out: torch.Tensor | None = None, force_generic: bool = False, vsplit: int | None = None,
                   path: str = "auto", rows=None, dup_rows=None, total_rows: int | None = None):
        """RGB uint8 HWC CUDA frames -> (pixel_values f32 [sum N_i, 1176] on device, image_grid_thw int64 [B, 3]).

        ``frames``: a ``[B, H, W, 3]`` tensor or a list of ``[H, W, 3]`` tensors (mixed sizes allowed).
        Same result as ``Qwen2VLImageProcessorPil(size={shortest_edge: min_pixels, longest_edge: max_pixels})``.
        Work is enqueued on the current CUDA stream; nothing synchronises.  ``path``: "auto" (statically scheduled
        kernel where it applies, general fused kernel otherwise) or "general" (never the scheduled kernel; tests).
        """
        if path not in ("auto", "general"):
            raise ValueError("path must be 'auto' or 'general'")
        self._align_launches = 0
        if not force_generic:
            if isinstance(frames, torch.Tensor) and frames.dim() == 4:
                if frames.is_cuda and (frames.stride(1) % 16 or frames.stride(0) % 16 or frames.data_ptr() % 16):
                    frames = self._align_frames(list(frames.unbind(0)))          # e.g. [B, 100, 502, 3]: 1506-byte rows
            else:
                frames = self._align_frames(list(frames))
        plan = self.plan_batch(frames, min_pixels, max_pixels, force_generic, vsplit, path, rows, dup_rows, total_rows)
        out = self._run_plan(plan, frames, out, self._align_launches)
        return out, plan.grid_thw

    def _run_plan(self, plan: "BatchPlan", frames, out, launches: int = 0):
        """Enqueue the launches of a batch plan on the current stream (``frames``: what the plan was made for)."""
        if out is None:
            out = torch.empty((plan.total_rows, G.ROW_FLOATS), dtype=torch.float32, device=self.device)
        elif (tuple(out.shape) != (plan.total_rows, G.ROW_FLOATS) or out.dtype != torch.float32
              or not out.is_contiguous() or out.device != self.device):
            raise ValueError(f"out must be a contiguous float32 [{plan.total_rows}, {G.ROW_FLOATS}] tensor on {self.device}")
        sp = _stream_ptr()
        flist = frames if plan.generic else None      # indexable either way ([B,H,W,3] tensor or list)
        for i, g, r0 in plan.generic:
            resized = self.resize_u8(flist[i], g.dst_h, g.dst_w, N.FILTER_BICUBIC)
            launches += self.last_launches + 1
            N.check(self.L.vis_normalize_patchify(resized.data_ptr(), resized.stride(0), g.dst_h, g.dst_w,
                                                  self.lut.data_ptr(), out.data_ptr(), r0, sp), "vis_normalize_patchify")
        for fl in plan.fused:
            if isinstance(fl, _SchedLaunch):
                N.check(self.L.vis_preprocess_fused_sched_dup(fl.sched.ctypes.data_as(C.c_void_p), fl.frames.data_ptr(),
                                                              fl.n_frames, fl.hrec.data_ptr(), fl.vrec.data_ptr(),
                                                              self.lut.data_ptr(), out.data_ptr(),
                                                              fl.dup.data_ptr() if fl.dup is not None else None, sp),
                        "vis_preprocess_fused_sched")
                launches += 1
                continue
            N.check(self.L.vis_preprocess_fused(fl.frames.data_ptr(), fl.n_frames, fl.strips.data_ptr(), fl.n_strips,
                                                fl.kt, fl.span_bytes, fl.strip_w,
                                                self.lut.data_ptr(), out.data_ptr(), sp), "vis_preprocess_fused")
            launches += 1
        self.last_launches = launches
        return out

    def preprocess_dual(self, frames, min_pixels: int = G.DEFAULT_MIN_PIXELS, max_pixels: int = G.DEFAULT_MAX_PIXELS,
                        out: torch.Tensor | None = None):
        """Both agents' Qwen2-VL inputs of a batch of RGB uint8 HWC CUDA frames in one pass (BASELINE config 5):
        Inspector = ``thumbnail((2048, 2048), LANCZOS)`` then the processor (src/agents/vlm_inspector.py:59-69),
        Auditor = ``thumbnail((1024, 1024), LANCZOS)`` then the processor (src/agents/vlm_auditor.py:87-96).

        Returns ``{"inspector": (pixel_values, image_grid_thw), "auditor": (...)}``; the two ``pixel_values`` are views
        of ONE allocation (Inspector rows first).  Thumbnails: one fused launch per (source geometry, role that needs
        one).  Processor: ONE plan over both roles' inputs, so frames of one geometry share a launch whatever role or
        source they come from (every Auditor thumbnail of a 16:9 frame is 1024x576).  A frame neither role thumbnails
        (longer side <= 1024) is resampled ONCE and its rows are stored to both tensors from the same registers.
        The whole pass (thumbnail staging tensors, descriptors, batch plan) is cached per set of frame addresses: a
        streaming loop that refills the same buffers pays for planning once.
        """
        batch = list(frames.unbind(0)) if isinstance(frames, torch.Tensor) and frames.dim() == 4 else list(frames)
        n = len(batch)
        if n == 0:
            raise ValueError("no frames")
        for f in batch:
            self._check_u8(f)
        key = (tuple((f.data_ptr(), tuple(f.shape), tuple(f.stride())) for f in batch), min_pixels, max_pixels)
        cache = self.__dict__.setdefault("_dual_plans", {})
        dp = cache.get(key)
        if dp is None:
            dp = self._plan_dual(batch, min_pixels, max_pixels)
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            cache[key] = dp
        launches = 0
        work = dp["work"]
        if dp["slow"]:                                   # thumbnails the fused kernel declines: recomputed, not cached
            work = list(work)
            for role_offset, role, idx in dp["slow"]:
                outs = self.agent_inputs([batch[i] for i in idx], role)
                launches += self.last_launches
                for i, o in zip(idx, outs):
                    work[dp["slot"][(role, i)]] = o
            plan = self.plan_batch(work, min_pixels, max_pixels, rows=dp["rows"], dup_rows=dp["dup"], total_rows=dp["total"])
        else:
            plan = dp["plan"]
        if self.dual_streams > 1 and not dp["slow"] and dp.get("early") is not None:
            pv = self._run_dual_streams(dp, plan, out)
        else:
            for rp in dp["thumbs"]:
                self._run_resize(rp)
                launches += 1
            if dp["align"] is not None:
                self._run_align(dp["align"])
                launches += 1
            pv = self._run_plan(plan, work, out, launches)
        ti = dp["total_i"]
        return {"inspector": (pv[:ti], dp["grid_i"]), "auditor": (pv[ti:], dp["grid_a"])}

    def _run_dual_streams(self, dp: dict, plan: "BatchPlan", out):
        """The launches of a cached dual pass on ``dual_streams`` CUDA streams.  They are persistent grids of at most one
        CTA per SM whose last round leaves SMs idle (4.2 rounds = 84 % of the machine for 624 items) and several cover
        fewer items than there are SMs; launches that do not depend on each other fill those gaps when they sit on
        different streams (tools/dual_breakdown.py: 3.25 -> 3.00 ms for the 192-frame slice of bench.py).  The thumbnails, the
        re-pitch launch and the processor launches that read caller frames only go first; a processor launch that reads
        thumbnails (or re-pitched frames) waits for the events of exactly those launches.  Forked from and joined back into
        the caller's stream."""
        if out is None:
            out = torch.empty((plan.total_rows, G.ROW_FLOATS), dtype=torch.float32, device=self.device)
        elif (tuple(out.shape) != (plan.total_rows, G.ROW_FLOATS) or out.dtype != torch.float32
              or not out.is_contiguous() or out.device != self.device):
            raise ValueError(f"out must be a contiguous float32 [{plan.total_rows}, {G.ROW_FLOATS}] tensor on {self.device}")
        cur = torch.cuda.current_stream()
        side = self.__dict__.get("_side_streams")
        if side is None or len(side) != self.dual_streams - 1:
            side = self._side_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.dual_streams - 1)]
        streams = [cur] + side
        fork = cur.record_event()
        for st in side:
            st.wait_event(fork)
        turn = [0]

        def on_next(fn):
            st = streams[turn[0] % len(streams)]
            turn[0] += 1
            with torch.cuda.stream(st):
                fn()
            return st

        def launch(fl):
            N.check(self.L.vis_preprocess_fused_sched_dup(fl.sched.ctypes.data_as(C.c_void_p), fl.frames.data_ptr(), fl.n_frames,
                                                          fl.hrec.data_ptr(), fl.vrec.data_ptr(), self.lut.data_ptr(),
                                                          out.data_ptr(), fl.dup.data_ptr() if fl.dup is not None else None,
                                                          _stream_ptr()), "vis_preprocess_fused_sched")
        n = 0
        t_ev = []                                        # (stream, event) per thumbnail launch
        for rp in dp["thumbs"]:
            st = on_next(lambda rp=rp: self._run_resize(rp))
            t_ev.append((st, st.record_event()))
            n += 1
        a_ev = None
        if dp["align"] is not None:
            st = on_next(lambda: self._run_align(dp["align"]))
            a_ev = (st, st.record_event())
            n += 1
        order = sorted(range(len(plan.fused)), key=lambda j: (len(dp["deps"][j][0]) > 0 or dp["deps"][j][1], j))
        for j in order:                                  # launches that read caller frames only go first
            thumbs_read, repitched = dp["deps"][j]
            st = streams[turn[0] % len(streams)]
            for k in thumbs_read:
                if t_ev[k][0] is not st:
                    st.wait_event(t_ev[k][1])
            if repitched and a_ev is not None and a_ev[0] is not st:
                st.wait_event(a_ev[1])
            on_next(lambda fl=plan.fused[j]: launch(fl))
            n += 1
        for st in side:
            cur.wait_event(st.record_event())
        self.last_launches = n
        return out

    def _plan_dual(self, batch: list, min_pixels: int, max_pixels: int) -> dict:
        n = len(batch)
        inputs = {"inspector": list(batch), "auditor": list(batch)}
        thumbs, slow = [], []
        for role, limit in (("inspector", G.INSPECTOR_MAX_SIZE), ("auditor", G.AUDITOR_MAX_SIZE)):
            groups: dict = {}
            for i, f in enumerate(batch):
                h, w = int(f.shape[0]), int(f.shape[1])
                if max(h, w) > limit:
                    groups.setdefault((h, w, f.stride(0)), []).append(i)
            for (h, w, _), idx in groups.items():
                tw, th = G.thumbnail_size(w, h, limit)
                rp = None
                if G.reducing_plan(w, h, tw, th, N.FILTER_LANCZOS) is None:
                    rp = self._resize_plan([batch[i] for i in idx], th, tw, N.FILTER_LANCZOS)
                if rp is None:                           # reduce pre-pass (>= 4x) or generic passes: per call
                    slow.append((0, role, idx))
                    for i in idx:
                        inputs[role][i] = None
                else:
                    thumbs.append(rp)
                    for i, o in zip(idx, rp[6].unbind(0)):
                        inputs[role][i] = o
        insp, aud = inputs["inspector"], inputs["auditor"]
        shared = [insp[i] is batch[i] and aud[i] is batch[i] for i in range(n)]

        def shape_of(role, i):
            f = inputs[role][i]
            if f is not None:
                return int(f.shape[0]), int(f.shape[1])
            limit = G.INSPECTOR_MAX_SIZE if role == "inspector" else G.AUDITOR_MAX_SIZE
            tw, th = G.thumbnail_size(int(batch[i].shape[1]), int(batch[i].shape[0]), limit)
            return th, tw

        def n_rows(role, i):
            h, w = shape_of(role, i)
            dh, dw = G.smart_resize(h, w, G.FACTOR, min_pixels, max_pixels)
            return (dh // G.PATCH_SIZE) * (dw // G.PATCH_SIZE), G.grid_thw(dh, dw)
        info_i = [n_rows("inspector", i) for i in range(n)]
        info_a = [n_rows("auditor", i) for i in range(n)]
        rows_i = np.array([r for r, _ in info_i], np.int64)
        rows_a = np.array([r for r, _ in info_a], np.int64)
        total_i, total_a = int(rows_i.sum()), int(rows_a.sum())
        at_i = np.concatenate([[0], np.cumsum(rows_i)[:-1]]).astype(np.int64)
        at_a = total_i + np.concatenate([[0], np.cumsum(rows_a)[:-1]]).astype(np.int64)
        own_a = [i for i in range(n) if not shared[i]]
        work = list(insp) + [aud[i] for i in own_a]
        slot = {("inspector", i): i for i in range(n)}
        slot.update({("auditor", i): n + k for k, i in enumerate(own_a)})
        rows = np.concatenate([at_i, at_a[own_a]]).astype(np.int64)
        dup = np.full(len(work), -1, np.int64)
        dup[:n] = np.where(np.asarray(shared), at_a, -1)
        dp = {"thumbs": thumbs, "slow": slow, "slot": slot, "rows": rows, "dup": dup, "total": total_i + total_a,
              "total_i": total_i, "grid_i": torch.tensor([g for _, g in info_i], dtype=torch.int64),
              "grid_a": torch.tensor([g for _, g in info_a], dtype=torch.int64), "align": None, "plan": None}
        if not slow:
            work, dp["align"] = self._align_frames(work, plan_only=True)
            dp["plan"] = self.plan_batch(work, min_pixels, max_pixels, rows=rows, dup_rows=dup, total_rows=dp["total"])
        dp["work"] = work
        dp["early"] = None
        plan = dp["plan"]
        if plan is not None and not plan.generic and all(isinstance(fl, _SchedLaunch) for fl in plan.fused):
            # what each processor launch reads besides caller frames: which thumbnail launches, and re-pitched frames
            base = {int(f.data_ptr()) for f in batch}
            spans = [(int(rp[6].data_ptr()), int(rp[6].data_ptr()) + rp[6].numel()) for rp in thumbs]
            deps = []
            for fl in plan.fused:
                srcs = [int(p) for p in np.frombuffer(fl.frames.cpu().numpy().tobytes(), N.FRAME_REF_DTYPE)["src"]]
                reads = sorted({k for p in srcs for k, (lo, hi) in enumerate(spans) if lo <= p < hi})
                other = any(p not in base and not any(lo <= p < hi for lo, hi in spans) for p in srcs)
                deps.append((reads, other))
            dp["deps"] = deps
            dp["early"] = [not r and not o for r, o in deps]
        return dp

    def _align_frames(self, frames: list, plan_only: bool = False):
        """Frames whose base or row pitch is not a multiple of 16 bytes (e.g. 502-pixel-wide rows) cannot be staged with
        bulk copies; all of them are repacked by ONE ``vis_repitch_u8`` launch into staging tensors with a padded pitch
        (one per shape) so that they take the fused kernels like everything else.  ``plan_only``: return
        (frames, re-runnable launch or None) with PRIVATE staging tensors instead of launching (``preprocess_dual``)."""
        groups: dict = {}
        for i, f in enumerate(frames):
            if (isinstance(f, torch.Tensor) and f.is_cuda and f.dtype == torch.uint8 and f.dim() == 3 and f.shape[2] == 3
                    and f.stride(2) == 1 and f.stride(1) == 3 and (f.stride(0) % 16 or f.data_ptr() % 16)):
                groups.setdefault((int(f.shape[0]), int(f.shape[1])), []).append(i)
        if not groups:
            return (frames, None) if plan_only else frames
        frames = list(frames)
        n = sum(len(idx) for idx in groups.values())
        desc = np.zeros(n, N.REPITCH_DTYPE)
        k, biggest, sources, bufs = 0, 0, [], []
        for (h, w), idx in groups.items():
            pitch = (w * 3 + 15) // 16 * 16
            key = ("align", h, w, len(idx))
            buf = None if plan_only else self._staging.get(key)
            if buf is None:
                buf = torch.zeros((len(idx), h, pitch), dtype=torch.uint8, device=self.device)
                if not plan_only:
                    self._staging[key] = buf
            bufs.append(buf)
            biggest = max(biggest, h * pitch)
            for j, i in enumerate(idx):
                f = frames[i]
                desc[k] = (f.data_ptr(), buf[j].data_ptr(), f.stride(0), pitch, h, w * 3)
                sources.append(f)
                frames[i] = buf[j].as_strided((h, w, 3), (pitch, 3, 1))
                k += 1
        ap = (torch.from_numpy(desc.view(np.uint8).copy()).to(self.device), n, biggest, sources, bufs)
        if plan_only:
            return frames, ap
        self._run_align(ap)
        self._keepalive_a = ap
        self._align_launches = 1
        return frames

    def _run_align(self, ap) -> None:
        N.check(self.L.vis_repitch_u8(ap[0].data_ptr(), ap[1], ap[2], _stream_ptr()), "vis_repitch_u8")

    def preprocess_host(self, host_frames: torch.Tensor, min_pixels: int = G.DEFAULT_MIN_PIXELS,
                        max_pixels: int = G.DEFAULT_MAX_PIXELS, out: torch.Tensor | None = None, chunk: int = 32,
                        host_out: torch.Tensor | None = None):
        """Same as ``preprocess`` for a HOST batch ``[B, H, W, 3]`` uint8 (pinned memory recommended).

        Frames are copied to the device in chunks on a side stream into two staging buffers while the previous
        chunk is being processed, so the PCIe transfer and the kernel overlap.  ``pixel_values`` stays on the device
        (its consumer is the vision tower); ``image_grid_thw`` is returned on the host.  ``host_out`` (pinned float32
        ``[rows, 1176]``): each chunk's rows are ALSO copied back on a third stream as soon as its kernel has finished,
        while the next chunk is still going up — PCIe is full duplex, so the read-back overlaps the upload; the
        current stream waits for the last copy, so synchronising it makes ``host_out`` complete.
        """
        if host_frames.is_cuda or host_frames.dtype != torch.uint8 or host_frames.dim() != 4 or host_frames.shape[3] != 3:
            raise TypeError("preprocess_host expects a CPU uint8 [B, H, W, 3] tensor")
        if not host_frames.is_contiguous():
            host_frames = host_frames.contiguous()
        b, h, w, _ = (int(v) for v in host_frames.shape)
        dh, dw = G.smart_resize(h, w, G.FACTOR, min_pixels, max_pixels)
        rows = (dh // G.PATCH_SIZE) * (dw // G.PATCH_SIZE)
        if out is None:
            out = torch.empty((b * rows, G.ROW_FLOATS), dtype=torch.float32, device=self.device)
        chunk = max(1, min(chunk, b))
        key = (chunk, h, w)
        st = self._staging.get(key)
        if st is None:
            st = self._staging[key] = {
                "buf": [torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=self.device) for _ in range(2)],
                "ready": [torch.cuda.Event() for _ in range(2)],
                "free": [torch.cuda.Event() for _ in range(2)],
                "stream": torch.cuda.Stream(device=self.device),
                "down": torch.cuda.Stream(device=self.device), "done": torch.cuda.Event()}
        cur = torch.cuda.current_stream()
        copy_stream, down = st["stream"], st["down"]
        if host_out is not None and (host_out.is_cuda or host_out.dtype != torch.float32 or
                                     tuple(host_out.shape) != (b * rows, G.ROW_FLOATS) or not host_out.is_contiguous()):
            raise ValueError(f"host_out must be a contiguous CPU float32 [{b * rows}, {G.ROW_FLOATS}] tensor")
        for ev in st["free"]:
            ev.record(cur)
        launches = 0
        for i, c0 in enumerate(range(0, b, chunk)):
            n = min(chunk, b - c0)
            slot = i & 1
            buf = st["buf"][slot]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(st["free"][slot])
                buf[:n].copy_(host_frames[c0:c0 + n], non_blocking=True)
                st["ready"][slot].record(copy_stream)
            cur.wait_event(st["ready"][slot])
            self.preprocess(buf[:n], min_pixels, max_pixels, out=out[c0 * rows:(c0 + n) * rows])
            launches += self.last_launches
            st["free"][slot].record(cur)
            if host_out is not None:
                st["done"].record(cur)
                with torch.cuda.stream(down):
                    down.wait_event(st["done"])
                    host_out[c0 * rows:(c0 + n) * rows].copy_(out[c0 * rows:(c0 + n) * rows], non_blocking=True)
        if host_out is not None:
            cur.wait_stream(down)
        self.last_launches = launches
        grid = torch.tensor([G.grid_thw(dh, dw)] * b, dtype=torch.int64)
        return out, grid

    def preprocess_jpeg(self, streams, min_pixels: int = G.DEFAULT_MIN_PIXELS, max_pixels: int = G.DEFAULT_MAX_PIXELS,
                        out: torch.Tensor | None = None, chunk: int = 256):
        """JPEG byte streams (host) -> (pixel_values, image_grid_thw): the streams cross PCIe COMPRESSED (~0.7 MB instead
        of 6.2 MB for a 1080p frame), are decoded on the GPU in batches (nvJPEG, ``jpeg.py``: a few levels away from
        libjpeg-turbo, so this entry point is tolerance-specified like every ``codec="nvjpeg"`` path) and go through the
        same kernels as ``preprocess``.  Replaces Image.open + the processor (src/agents/vlm_inspector.py:59)."""
        streams = list(streams)
        if not streams:
            raise ValueError("no streams")
        codec = self.jpeg_codec()
        sizes = [codec.info(s)[:2] for s in streams]
        n_rows = []
        for w, h in sizes:
            dh, dw = G.smart_resize(h, w, G.FACTOR, min_pixels, max_pixels)
            n_rows.append((dh // G.PATCH_SIZE) * (dw // G.PATCH_SIZE))
        at = np.concatenate([[0], np.cumsum(n_rows)]).astype(np.int64)
        if out is None:
            out = torch.empty((int(at[-1]), G.ROW_FLOATS), dtype=torch.float32, device=self.device)
        launches, grids = 0, []
        for c0 in range(0, len(streams), chunk):
            frames = codec.decode_batch(streams[c0:c0 + chunk])
            _, grid = self.preprocess(frames, min_pixels, max_pixels, out=out[int(at[c0]):int(at[min(c0 + chunk, len(streams))])])
            launches += self.last_launches
            grids.append(grid)
        self.last_launches = launches
        return out, torch.cat(grids)

    # ------------------------------------------------------------------ defect overlay
    def _marker_sprite(self, radius: int, b: int, g: int, r: int, label: bytes):
        """The marker of (radius, colour, label) rasterised once ON THE DEVICE: its leaves (alpha 255) drawn by the overlay
        kernel on a zeroed BGRA canvas.  Returns (canvas tensor, w, h, ox, oy); cached per engine."""
        try:
            return self._render_template(("marker", radius, b, g, r, label), 4,
                                         lambda *a: self.L.vis_overlay_sprite_expand(radius, b, g, r, label, *a),
                                         "vis_overlay_sprite_expand")
        except N.VisError as e:
            if e.code != N.VIS_E_UNSUPPORTED:         # a label wider than any private canvas: expanded in place instead
                raise
            self.__dict__.setdefault("_sprites", {})[("marker", radius, b, g, r, label)] = None
            return None

    def _dash_stamp(self, dx: int, dy: int):
        """The blend chains of the dash (0,0)-(dx,dy) recorded once ON THE DEVICE (record mode of the overlay kernel: 8
        bytes per pixel).  Returns (canvas, w, h, ox, oy), or None when a chain overflowed its seven slots."""
        return self._render_template(("dash", dx, dy), 8, lambda *a: self.L.vis_overlay_stamp_expand(dx, dy, *a),
                                     "vis_overlay_stamp_expand")

    def _render_template(self, key, channels: int, expand, where: str):
        cache = self.__dict__.setdefault("_sprites", {})
        if key in cache:
            return cache[key]
        from . import overlay as O
        w, h, ox, oy, needed = (C.c_int(0) for _ in range(5))
        cap = 2048
        while True:
            leaves = np.empty(cap, N.LEAF_DTYPE)
            rc = expand(leaves.ctypes.data_as(C.c_void_p), cap, C.byref(needed), C.byref(w), C.byref(h), C.byref(ox), C.byref(oy))
            if rc == N.VIS_E_CAPACITY:
                cap = needed.value
                continue
            N.check(rc, where)
            leaves = leaves[:rc].copy()
            break
        canvas = torch.zeros((h.value, w.value, channels), dtype=torch.uint8, device=self.device)
        tiles, refs = O.touched_tiles(leaves, 1, w.value, h.value)
        tl = np.zeros(len(tiles), N.OVERLAY_TILE_DTYPE)
        tl["txy"], tl["ref_begin"], tl["ref_end"] = tiles[:, 0], tiles[:, 1], tiles[:, 2]
        desc = np.zeros(1, N.OVERLAY_FRAME_DTYPE)
        desc["src"] = desc["dst"] = canvas.data_ptr()
        desc["src_pitch"] = desc["dst_pitch"] = canvas.stride(0)
        desc["h"], desc["w"], desc["group_begin"], desc["group_end"] = h.value, w.value, 0, 1
        up = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(self.device)  # noqa: E731
        d = [up(desc), up(tl), up(np.ascontiguousarray(refs)), up(leaves)]
        N.check(self.L.vis_overlay_draw_cn(d[0].data_ptr(), 1, channels, 0, d[1].data_ptr(), len(tl), d[2].data_ptr(),
                                           d[3].data_ptr(), _stream_ptr()), "vis_overlay_draw_cn")
        hit = (canvas, w.value, h.value, ox.value, oy.value)
        if channels == 8:
            # a blend chain that overflowed its seven slots (flag 0x40 in byte 0 of the record) means this dash stays
            # leaves: checked on the host from a copy of the (at most ~1 KB) record canvas — once per dash geometry per engine
            if (canvas.cpu().numpy()[:, :, 0] & 0x40).any():
                hit = None
        # (no synchronisation: the upload tensors go back to torch's stream-ordered allocator, the canvas stays)
        if len(cache) >= 4096:
            cache.clear()
        cache[key] = hit
        return hit

    def plan_overlay(self, shapes, boxes_per_frame, confidence_threshold: str = "low", criticality: str = "medium"):
        """Host half of the overlay for a batch: box validation (reference rules) + expansion into leaves.

        ``shapes``: list of (H, W) per frame.  Returns (leaves uint8 device tensor, per-frame (begin, end) header
        ranges, number of boxes drawn, device tile list, tile count, device ref list).
        """
        from . import overlay as O
        # host rules per frame (Python, like the reference), then ONE threaded C call for expansion + tile binning
        rows_all, box_begin, hw = [], [0], np.zeros((len(shapes), 2), np.int32)
        keys, any_dashed = set(), False
        for i, ((h, w), boxes) in enumerate(zip(shapes, boxes_per_frame)):
            rows = O.box_rows(boxes, w, h, confidence_threshold, criticality)
            rows_all.extend(rows)
            box_begin.append(box_begin[-1] + len(rows))
            hw[i] = (h, w)
            if rows:
                # markers are opaque: each distinct (radius, colour, label) of the batch is rasterised once on the device
                # and referenced by one sprite leaf per box (utils/image_utils.py:292-293 for the radius rule)
                radius = max(25, min(int(max(w, h) * 0.04), 60))
                for r in rows:
                    keys.add((radius, r[4], r[5], r[6], r[8]))
                    any_dashed = any_dashed or bool(r[7])
        boxes_arr = N.host_records(rows_all, N.BOX_DTYPE, "label") if rows_all else np.zeros(1, N.BOX_DTYPE)
        # dashes (confidence == "low") are 10 px long, the last one of an edge 1..9: one recorded stamp per length and
        # orientation, shared by every colour
        dashes = [(L, 0) for L in range(1, 11)] + [(0, L) for L in range(1, 11)] if any_dashed else []
        entries, keep_sprites = [], []
        for key in sorted(keys):
            hit = self._marker_sprite(*key)
            if hit is None:
                continue
            canvas, sw, sh, sox, soy = hit
            keep_sprites.append(canvas)
            entries.append((key[0], key[1], key[2], key[3], 0, key[4], canvas.data_ptr(), sw, sh, sox, soy))
        for dx, dy in dashes:
            hit = self._dash_stamp(dx, dy)
            if hit is not None:
                keep_sprites.append(hit[0])
                entries.append((-1, 0, 0, 0, 0, f"{dx},{dy}".encode(), hit[0].data_ptr(), hit[1], hit[2], hit[3], hit[4]))
        sprites = N.host_records(entries, N.SPRITE_DTYPE, "label") if entries else np.zeros(1, N.SPRITE_DTYPE)
        box_begin = np.asarray(box_begin, np.int32)
        n = len(shapes)
        leaf_begin = np.zeros(n + 1, np.int32)
        needed = np.zeros(3, np.int64)
        cap = [max(1, int(box_begin[-1]) * 1600), max(1, int(box_begin[-1]) * 160), max(1, int(box_begin[-1]) * 640)]
        # the plan is written straight into PINNED host buffers (torch's caching host allocator reuses them from call to
        # call), so the upload below is one asynchronous DMA per array instead of a staged copy of pageable memory:
        # the leaves alone are ~280 KB per annotated 1080p frame
        while True:
            pinned = [torch.empty(max(1, c) * dt.itemsize, dtype=torch.uint8, pin_memory=True)
                      for c, dt in zip(cap, (N.LEAF_DTYPE, N.OVERLAY_TILE_DTYPE, N.OVERLAY_REF_DTYPE))]
            leaves, tiles, refs = (t.numpy().view(dt) for t, dt in
                                   zip(pinned, (N.LEAF_DTYPE, N.OVERLAY_TILE_DTYPE, N.OVERLAY_REF_DTYPE)))
            rc = self.L.vis_overlay_plan_batch_sprites(
                n, hw.ctypes.data_as(C.c_void_p), boxes_arr.ctypes.data_as(C.c_void_p), box_begin.ctypes.data_as(C.c_void_p),
                leaves.ctypes.data_as(C.c_void_p), cap[0], leaf_begin.ctypes.data_as(C.c_void_p),
                tiles.ctypes.data_as(C.c_void_p), cap[1], refs.ctypes.data_as(C.c_void_p), cap[2],
                needed.ctypes.data_as(C.c_void_p), 0, sprites.ctypes.data_as(C.c_void_p), len(entries))
            if rc == N.VIS_E_CAPACITY:
                cap = [max(1, int(v)) for v in needed]
                continue
            N.check(rc, "vis_overlay_plan_batch")
            break
        n_leaves, n_tiles, n_refs = (int(v) for v in needed)
        ranges = [(int(leaf_begin[i]), int(leaf_begin[i]) + int(box_begin[i + 1] - box_begin[i])) for i in range(n)]
        up = lambda t, k, dt: t[:max(k, 1) * dt.itemsize].to(self.device, non_blocking=True)  # noqa: E731
        return (up(pinned[0], n_leaves, N.LEAF_DTYPE), ranges, int(box_begin[-1]), up(pinned[1], n_tiles, N.OVERLAY_TILE_DTYPE),
                n_tiles, up(pinned[2], n_refs, N.OVERLAY_REF_DTYPE), keep_sprites)     # the plan keeps its sprites alive

    def annotate(self, frames, boxes_per_frame, confidence_threshold: str = "low", criticality: str = "medium",
                 inplace: bool = False, plan=None):
        """Draw the reference's defect overlay on BGR uint8 HWC CUDA frames (``[B,H,W,3]`` tensor or list).

        Returns new tensors (or the inputs when ``inplace``).  Pixels equal what the reference's
        ``draw_bounding_boxes`` holds just before ``cv2.imwrite``.
        """
        uniform = isinstance(frames, torch.Tensor) and frames.dim() == 4
        if uniform:
            # a batch tensor: one check, descriptors by arithmetic (1024 unbind() views cost more than the draw)
            self._check_u8(frames)
            if frames.shape[3] != 3 or frames.stride(3) != 1 or frames.stride(2) != 3:
                raise ValueError("frames must be [B, H, W, 3] uint8 with contiguous pixels")
            n = int(frames.shape[0])
            if n != len(boxes_per_frame):
                raise ValueError("one box list per frame expected")
            shapes = [(int(frames.shape[1]), int(frames.shape[2]))] * n
            result = frames if inplace else torch.empty_like(frames)
            idx = np.arange(n, dtype=np.uint64)
            src_ptr = np.uint64(frames.data_ptr()) + idx * np.uint64(frames.stride(0))
            dst_ptr = np.uint64(result.data_ptr()) + idx * np.uint64(result.stride(0))
            src_pitch, dst_pitch = frames.stride(1), result.stride(1)
        else:
            flist = list(frames)
            n = len(flist)
            if n != len(boxes_per_frame):
                raise ValueError("one box list per frame expected")
            for f in flist:
                self._check_u8(f)
                if f.dim() != 3 or f.shape[2] != 3 or f.stride(2) != 1 or f.stride(1) != 3:
                    raise ValueError("frames must be [H, W, 3] uint8 with contiguous pixels")
            shapes = [(int(f.shape[0]), int(f.shape[1])) for f in flist]
            outs = flist if inplace else [torch.empty_like(f) for f in flist]
            result = frames if inplace else outs
            src_ptr, dst_ptr = [f.data_ptr() for f in flist], [o.data_ptr() for o in outs]
            src_pitch, dst_pitch = [f.stride(0) for f in flist], [o.stride(0) for o in outs]
        if plan is None:
            plan = self.plan_overlay(shapes, boxes_per_frame, confidence_threshold, criticality)
        d_leaves, ranges, _, d_tiles, n_tiles, d_refs = plan[:6]
        desc = np.zeros(n, N.OVERLAY_FRAME_DTYPE)
        desc["src"], desc["dst"] = src_ptr, dst_ptr
        desc["src_pitch"], desc["dst_pitch"] = src_pitch, dst_pitch
        desc["h"] = [s_[0] for s_ in shapes]
        desc["w"] = [s_[1] for s_ in shapes]
        rng_arr = np.asarray(ranges, np.int32).reshape(-1, 2)
        desc["group_begin"], desc["group_end"] = rng_arr[:, 0], rng_arr[:, 1]
        d_desc = torch.from_numpy(desc.view(np.uint8).copy()).to(self.device)
        sp = _stream_ptr()
        if n > 65535:
            raise ValueError("at most 65535 frames per annotate call")
        N.check(self.L.vis_overlay_draw(d_desc.data_ptr(), n, 0 if inplace else 1, d_tiles.data_ptr(), n_tiles,
                                        d_refs.data_ptr(), d_leaves.data_ptr(), sp), "vis_overlay_draw")
        self.last_launches = (0 if inplace else 1) + (1 if n_tiles else 0)
        self._keepalive = (d_desc, d_leaves, d_tiles, d_refs)
        return result

    # ------------------------------------------------------------------ defect heat-map overlay
    def heatmap(self, frame: torch.Tensor, defects: list) -> torch.Tensor:
        """``create_heatmap_overlay`` for one BGR uint8 HWC CUDA frame and the reference's defect dicts (percent boxes,
        ``safety_impact``, ``confidence``, ``location``): returns the blended BGR frame (a copy when ``defects`` is
        empty, like the reference).  Tolerance-specified (float32 blur), see vis_heatmap.cu."""
        return self.heatmap_batch([frame], [defects])[0]

    def plan_heatmap(self, shapes, defects_per_frame) -> dict:
        """Host half of the heat map for a batch: the reference's per-defect scalar code (``heatmap.defect_params``), the
        Gaussian kernels, and the item / offset tables of ``vis_heatmap_batch`` (uploaded here).  ``shapes``: (H, W) per
        frame.  Reusable for any frames of those shapes."""
        from . import heatmap as H
        items, kernels, kcache, frames = [], [], {}, []
        koff = plane = tmp = tab = 0

        def kernel_offset(k: np.ndarray) -> int:
            nonlocal koff
            key = k.tobytes()
            at = kcache.get(key)
            if at is None:
                at = kcache[key] = koff
                kernels.append(k)
                koff += len(k)
            return at
        copies = []
        for i, ((h, w), defects) in enumerate(zip(shapes, defects_per_frame)):
            recs, kern, had = H.defect_params(defects, w, h)
            if not had:                               # no defects: the reference writes the image back unchanged
                copies.append(i)
                continue
            fk, fkern = H.final_blur(w, h)
            frames.append((i, h, w, plane, fk, kernel_offset(fkern)))
            if len(recs):
                rw, rh, ks = recs["x2"] - recs["x1"], recs["y2"] - recs["y1"], recs["ksize"]
                if ((rw <= 0) | (rh <= 0) | (recs["x1"] < 0) | (recs["y1"] < 0) | (recs["x2"] > w) | (recs["y2"] > h) |
                        (ks < 1) | (ks > 51) | (ks % 2 == 0)).any():
                    raise ValueError("heat-map defect with an invalid region or blur kernel")
                it = np.zeros(len(recs), H.ITEM_DTYPE)
                it["d"] = recs
                for j in np.nonzero(ks > 1)[0]:
                    o = int(recs["koff"][j])
                    it["d"]["koff"][j] = kernel_offset(kern[o:o + int(ks[j])])
                it["frame"] = len(frames) - 1
                tsz = 3 * (rw.astype(np.int64) + rh)
                it["tab_off"] = tab + np.concatenate([[0], np.cumsum(tsz)[:-1]])
                tab += int(tsz.sum())
                blurred = (recs["kind"] == 0) & (ks > 1)
                rsz = np.where(blurred, rw.astype(np.int64) * rh, 0)
                it["tmp_off"] = tmp + np.concatenate([[0], np.cumsum(rsz)[:-1]])
                tmp += int(rsz.sum())
                items.append(it)
            plane += h * w
        if len(frames) > 65535 or sum(len(it) for it in items) > 65535:
            raise ValueError("at most 65535 frames / defects per heatmap_batch call")
        it_arr = np.concatenate(items) if items else np.zeros(0, H.ITEM_DTYPE)
        up = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(self.device)  # noqa: E731
        plan = {"copies": copies, "frames": frames, "n_items": len(it_arr), "plane": plane, "tmp": tmp, "tab": tab,
                "shapes": [tuple(s_) for s_ in shapes]}
        if frames:
            plan["d_items"] = up(it_arr) if len(it_arr) else None
            plan["d_kern"] = torch.from_numpy(np.concatenate(kernels).astype(np.float32)).to(self.device)
            plan["max_rw"] = int((it_arr["d"]["x2"] - it_arr["d"]["x1"]).max()) if len(it_arr) else 0
            plan["max_rh"] = int((it_arr["d"]["y2"] - it_arr["d"]["y1"]).max()) if len(it_arr) else 0
        return plan

    def heatmap_batch(self, frames, defects_per_frame=None, plan: dict | None = None) -> list:
        """``create_heatmap_overlay`` for a batch of BGR uint8 HWC CUDA frames (a ``[B,H,W,3]`` tensor or a list, mixed
        sizes allowed): SIX launches for the whole batch.  Returns a list of new frames, aligned with the input.
        ``plan``: a ``plan_heatmap`` result for these shapes (the host half, reusable)."""
        from . import heatmap as H
        batch = list(frames.unbind(0)) if isinstance(frames, torch.Tensor) and frames.dim() == 4 else list(frames)
        for f in batch:
            self._check_u8(f)
            if f.dim() != 3 or f.shape[2] != 3 or f.stride(2) != 1 or f.stride(1) != 3:
                raise ValueError("frames must be [H, W, 3] uint8 with contiguous pixels")
        shapes = [(int(f.shape[0]), int(f.shape[1])) for f in batch]
        if plan is None:
            if defects_per_frame is None or len(batch) != len(defects_per_frame):
                raise ValueError("one defect list per frame expected")
            plan = self.plan_heatmap(shapes, defects_per_frame)
        elif plan["shapes"] != shapes:
            raise ValueError("the plan was made for other frame shapes")
        outs = [None] * len(batch)
        for i in plan["copies"]:
            outs[i] = batch[i].clone()
        if not plan["frames"]:
            self.last_launches = 0
            return outs
        fr = np.zeros(len(plan["frames"]), N.HEAT_FRAME_DTYPE)
        for j, (i, h, w, plane_off, fk, fkoff) in enumerate(plan["frames"]):
            out = torch.empty((h, w, 3), dtype=torch.uint8, device=self.device)
            outs[i] = out
            fr[j] = (batch[i].data_ptr(), out.data_ptr(), batch[i].stride(0), out.stride(0), h, w, plane_off, fk, fkoff)
        d_fr = torch.from_numpy(fr.view(np.uint8).reshape(-1).copy()).to(self.device)
        if not hasattr(self, "_jet"):
            self._jet = torch.from_numpy(H.JET_BGR.copy()).to(self.device)
        planes = torch.empty((3, plan["plane"]), dtype=torch.float32, device=self.device)
        d_tmp = torch.empty(max(plan["tmp"], 1), dtype=torch.float32, device=self.device)
        d_tab = torch.empty(max(plan["tab"], 1), dtype=torch.float64, device=self.device)
        d_max = torch.empty(len(fr), dtype=torch.int32, device=self.device)
        n_items = plan["n_items"]
        N.check(self.L.vis_heatmap_batch(d_fr.data_ptr(), len(fr), plan["d_items"].data_ptr() if n_items else None, n_items,
                                         int(fr["w"].max()), int(fr["h"].max()), plan["max_rw"], plan["max_rh"], plan["plane"],
                                         plan["d_kern"].data_ptr(), self._jet.data_ptr(), planes[0].data_ptr(),
                                         planes[1].data_ptr(), planes[2].data_ptr(), d_tmp.data_ptr(), d_tab.data_ptr(),
                                         d_max.data_ptr(), _stream_ptr()), "vis_heatmap_batch")
        self.last_launches = 3 + (3 if n_items else 0)
        self._keepalive_h = (d_fr, plan, planes, d_tmp, d_tab, d_max)
        return outs

    # ------------------------------------------------------------------ comparison panel / status stamp
    def _draw_plan(self, h: int, w: int, cmds: np.ndarray) -> dict:
        """The expanded plan of a draw list (``compare.*_commands``) on an [h, w] canvas — leaves, touched tiles, refs, as
        host arrays and device copies — cached per (canvas size, draw list content)."""
        from . import compare as CP
        from . import overlay as O
        up = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(self.device)  # noqa: E731
        cache = self.__dict__.setdefault("_draw_plans", {})
        key = (h, w, CP.commands_key(cmds))
        plan = cache.get(key)
        if plan is None:
            leaves = CP.expand_commands(cmds, w, h)
            tiles, refs = O.touched_tiles(leaves, len(cmds), w, h)
            tl = np.zeros(len(tiles), N.OVERLAY_TILE_DTYPE)
            if len(tiles):
                tl["txy"], tl["ref_begin"], tl["ref_end"] = tiles[:, 0], tiles[:, 1], tiles[:, 2]
            refs = np.ascontiguousarray(refs)
            if len(cache) >= 64:
                cache.clear()
            plan = cache[key] = {"n_tiles": len(tl), "n_cmds": len(cmds), "tiles": tl, "refs": refs, "leaves": leaves,
                                 "d_tiles": up(tl) if len(tl) else None, "d_refs": up(refs) if len(tl) else None,
                                 "d_leaves": up(leaves) if len(tl) else None}
        return plan

    def _draw_list(self, canvas: torch.Tensor, cmds: np.ndarray) -> int:
        """Apply a draw list in place on a [H, W, 3|4] uint8 CUDA canvas; returns launches."""
        h, w, cn = int(canvas.shape[0]), int(canvas.shape[1]), int(canvas.shape[2])
        plan = self._draw_plan(h, w, cmds)
        if plan["n_tiles"] == 0:
            return 0
        desc = np.zeros(1, N.OVERLAY_FRAME_DTYPE)
        desc["src"] = desc["dst"] = canvas.data_ptr()
        desc["src_pitch"] = desc["dst_pitch"] = canvas.stride(0)
        desc["h"], desc["w"] = h, w
        desc["group_begin"], desc["group_end"] = 0, plan["n_cmds"]
        d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1).copy()).to(self.device)
        N.check(self.L.vis_overlay_draw_cn(d_desc.data_ptr(), 1, cn, 0, plan["d_tiles"].data_ptr(), plan["n_tiles"],
                                           plan["d_refs"].data_ptr(), plan["d_leaves"].data_ptr(), _stream_ptr()), "vis_overlay_draw_cn")
        self._keepalive_d = (d_desc, plan)
        return 1

    def _linear_tables(self, src_size: int, dst_size: int, is_x: bool) -> list:
        """Device copies of one axis of the bilinear tables (``vis_linear_table``), cached per geometry."""
        from . import compare as CP
        cache = self.__dict__.setdefault("_linear_cache", {})
        key = (src_size, dst_size, is_x)
        if key not in cache:
            if len(cache) >= 256:
                cache.clear()
            cache[key] = [torch.from_numpy(a).to(self.device) for a in CP.linear_tables(src_size, dst_size, is_x)]
        return cache[key]

    def _pair_panels(self, original: torch.Tensor, annotated: torch.Tensor, panels: np.ndarray, keep: list):
        """Fill the two ``VisPanel`` records of a comparison canvas; returns (left_w, right_w)."""
        from . import compare as CP
        org_x = 0
        for i, f in enumerate((original, annotated)):
            self._check_u8(f)
            if f.dim() != 3 or f.shape[2] != 3 or f.stride(2) != 1 or f.stride(1) != 3:
                raise ValueError("frames must be [H, W, 3] uint8 with contiguous pixels")
            h, w = int(f.shape[0]), int(f.shape[1])
            dw = CP.panel_width(h, w)
            if dw < 1:
                raise ValueError(f"frame {w}x{h} is too narrow for an {CP.TARGET_HEIGHT}-row panel")
            mode = N.check(self.L.vis_resize_linear_mode(h, w, CP.TARGET_HEIGHT, dw), "vis_resize_linear_mode")
            p = panels[i]
            p["src"], p["src_pitch"], p["src_h"], p["src_w"] = f.data_ptr(), f.stride(0), h, w
            p["dst_h"], p["dst_w"], p["org_x"], p["org_y"], p["mode"] = CP.TARGET_HEIGHT, dw, org_x, CP.HEADER_HEIGHT, mode
            if mode == N.RESIZE_BILINEAR:
                dev = self._linear_tables(w, dw, True) + self._linear_tables(h, CP.TARGET_HEIGHT, False)
                keep += dev
                p["xofs"], p["alpha"], p["yofs"], p["beta"] = (t.data_ptr() for t in dev)
            org_x += dw + CP.DIVIDER_WIDTH
        return int(panels[0]["dst_w"]), int(panels[1]["dst_w"])

    def side_by_side(self, original: torch.Tensor, annotated: torch.Tensor, labels=None) -> torch.Tensor:
        """``create_side_by_side_comparison`` for two BGR uint8 HWC CUDA frames: both resized to a height of 800 as
        ``cv2.resize`` does, a 40-row header and a 10-column divider of gray 45, two centred white labels.  Returns the
        [840, W1 + 10 + W2, 3] canvas the reference hands to ``cv2.imwrite``; bit-exact.  Two launches."""
        from . import compare as CP
        labels = CP.DEFAULT_LABELS if labels is None else labels
        keep, panels = [], np.zeros(2, N.PANEL_DTYPE)
        left_w, right_w = self._pair_panels(original, annotated, panels, keep)
        total_w = left_w + CP.DIVIDER_WIDTH + right_w
        canvas = torch.empty((CP.HEADER_HEIGHT + CP.TARGET_HEIGHT, total_w, 3), dtype=torch.uint8, device=self.device)
        N.check(self.L.vis_compose_panels(canvas.data_ptr(), canvas.stride(0), int(canvas.shape[0]), total_w, CP.BAR_GRAY,
                                          panels.ctypes.data_as(C.c_void_p), 2, _stream_ptr()), "vis_compose_panels")
        self._keepalive_c = keep
        self.last_launches = 1 + self._draw_list(canvas, CP.header_commands(left_w, right_w, labels))
        return canvas

    def _pair_geometry(self, h1: int, w1: int, h2: int, w2: int):
        """Everything of a comparison canvas that depends on the two frame sizes only: a ``VisPanelCanvas`` record with
        the panel geometry, modes and device tables filled in (sources and canvas left open), and (left_w, right_w)."""
        from . import compare as CP
        cache = self.__dict__.setdefault("_pair_geo", {})
        key = (h1, w1, h2, w2)
        if key not in cache:
            rec = np.zeros(1, N.PANEL_CANVAS_DTYPE)
            keep, org_x = [], 0
            for i, (h, w) in enumerate(((h1, w1), (h2, w2))):
                dw = CP.panel_width(h, w)
                if dw < 1:
                    raise ValueError(f"frame {w}x{h} is too narrow for an {CP.TARGET_HEIGHT}-row panel")
                mode = N.check(self.L.vis_resize_linear_mode(h, w, CP.TARGET_HEIGHT, dw), "vis_resize_linear_mode")
                p = rec[0]["panels"][i]
                p["src_h"], p["src_w"] = h, w
                p["dst_h"], p["dst_w"], p["org_x"], p["org_y"], p["mode"] = CP.TARGET_HEIGHT, dw, org_x, CP.HEADER_HEIGHT, mode
                if mode == N.RESIZE_BILINEAR:
                    dev = self._linear_tables(w, dw, True) + self._linear_tables(h, CP.TARGET_HEIGHT, False)
                    keep += dev
                    p["xofs"], p["alpha"], p["yofs"], p["beta"] = (t.data_ptr() for t in dev)
                org_x += dw + CP.DIVIDER_WIDTH
            lw, rw = int(rec[0]["panels"][0]["dst_w"]), int(rec[0]["panels"][1]["dst_w"])
            rec["h"], rec["w"] = CP.HEADER_HEIGHT + CP.TARGET_HEIGHT, lw + CP.DIVIDER_WIDTH + rw
            rec["fill"], rec["n_panels"] = CP.BAR_GRAY, 2
            if len(cache) >= 256:
                cache.clear()
            cache[key] = (rec, lw, rw, keep)
        return cache[key]

    def _frame_table(self, frames):
        """(ptr uint64[n], pitch int64[n], h int32[n], w int32[n]) of BGR uint8 HWC CUDA frames: a [B, H, W, 3] tensor by
        arithmetic, a list frame by frame."""
        if isinstance(frames, torch.Tensor) and frames.dim() == 4:
            self._check_u8(frames)
            if frames.shape[3] != 3 or frames.stride(3) != 1 or frames.stride(2) != 3:
                raise ValueError("frames must be [B, H, W, 3] uint8 with contiguous pixels")
            n = int(frames.shape[0])
            ptr = np.uint64(frames.data_ptr()) + np.arange(n, dtype=np.uint64) * np.uint64(frames.stride(0))
            return (ptr, np.full(n, frames.stride(1), np.int64), np.full(n, int(frames.shape[1]), np.int32),
                    np.full(n, int(frames.shape[2]), np.int32))
        flist = list(frames)
        for f in flist:
            self._check_u8(f)
            if f.dim() != 3 or f.shape[2] != 3 or f.stride(2) != 1 or f.stride(1) != 3:
                raise ValueError("frames must be [H, W, 3] uint8 with contiguous pixels")
        return (np.array([f.data_ptr() for f in flist], np.uint64), np.array([f.stride(0) for f in flist], np.int64),
                np.array([f.shape[0] for f in flist], np.int32), np.array([f.shape[1] for f in flist], np.int32))

    def side_by_side_batch(self, originals, annotateds, labels=None) -> list:
        """``side_by_side`` for many (original, annotated) pairs — a report run composes one panel per inspected image:
        ONE compose launch and ONE label launch for the whole batch (``vis_compose_panels_batch`` + the overlay draw
        kernel over every canvas's header tiles).  ``originals`` / ``annotateds``: lists of [H, W, 3] frames or
        [B, H, W, 3] tensors.  Returns one [840, W, 3] canvas per pair (views of one allocation per distinct width);
        every canvas equals what ``side_by_side`` returns for its pair."""
        from . import compare as CP
        labels = CP.DEFAULT_LABELS if labels is None else labels
        pa, pitch_a, ha, wa = self._frame_table(originals)
        pb, pitch_b, hb, wb = self._frame_table(annotateds)
        n = len(pa)
        if n == 0 or n != len(pb):
            raise ValueError("one annotated frame per original expected (at least one pair)")
        if n > 65535:
            raise ValueError("at most 65535 pairs per call")
        H = CP.HEADER_HEIGHT + CP.TARGET_HEIGHT
        shapes = np.stack([ha, wa, hb, wb], axis=1)
        uniq, inverse = np.unique(shapes, axis=0, return_inverse=True)
        inverse = inverse.reshape(-1)
        recs, keep, geo = np.zeros(n, N.PANEL_CANVAS_DTYPE), [], []
        for g, (h1, w1, h2, w2) in enumerate(uniq):
            rec, lw, rw, tabs = self._pair_geometry(int(h1), int(w1), int(h2), int(w2))
            recs[inverse == g] = rec[0]
            geo.append((lw, rw))
            keep.append(tabs)
        recs["panels"]["src"][:, 0], recs["panels"]["src"][:, 1] = pa, pb
        recs["panels"]["src_pitch"][:, 0], recs["panels"]["src_pitch"][:, 1] = pitch_a, pitch_b
        totals = recs["w"].astype(np.int64)
        canvases, ptrs, pitches = [None] * n, np.zeros(n, np.uint64), np.zeros(n, np.int64)
        for tw in np.unique(totals):                      # one allocation per distinct canvas width
            idx = np.flatnonzero(totals == tw)
            block = torch.empty((len(idx), H, int(tw), 3), dtype=torch.uint8, device=self.device)
            ptrs[idx] = np.uint64(block.data_ptr()) + np.arange(len(idx), dtype=np.uint64) * np.uint64(block.stride(0))
            pitches[idx] = block.stride(1)
            for k, view in zip(idx, block.unbind(0)):
                canvases[k] = view
        recs["canvas"], recs["pitch"] = ptrs, pitches
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(self.device)  # noqa: E731
        d_recs = up(recs)
        sp = _stream_ptr()
        N.check(self.L.vis_compose_panels_batch(d_recs.data_ptr(), n, H, int(totals.max()), sp), "vis_compose_panels_batch")
        # the header labels: canvases of one geometry share one expanded draw list; its tiles are instantiated per canvas
        plans = [self._draw_plan(H, lw + CP.DIVIDER_WIDTH + rw, CP.header_commands(lw, rw, labels)) for lw, rw in geo]
        leaf_base = np.concatenate(([0], np.cumsum([len(p["leaves"]) for p in plans])))
        ref_base = np.concatenate(([0], np.cumsum([len(p["refs"]) for p in plans])))
        desc = np.zeros(n, N.OVERLAY_FRAME_DTYPE)
        desc["src"] = desc["dst"] = ptrs
        desc["src_pitch"] = desc["dst_pitch"] = pitches
        desc["h"], desc["w"] = H, totals
        desc["group_begin"] = leaf_base[inverse]
        desc["group_end"] = leaf_base[inverse] + np.array([p["n_cmds"] for p in plans])[inverse]
        tiles = []
        for g, p in enumerate(plans):
            idx = np.flatnonzero(inverse == g)
            if p["n_tiles"] == 0 or len(idx) == 0:
                continue
            t = np.tile(p["tiles"], len(idx))
            t["frame"] = np.repeat(idx, p["n_tiles"])
            t["ref_begin"] += ref_base[g]
            t["ref_end"] += ref_base[g]
            tiles.append(t)
        self.last_launches = 1
        d_draw = None
        if tiles:
            tiles = np.concatenate(tiles)
            d_draw = (up(desc), up(tiles), up(np.concatenate([p["refs"] for p in plans])),
                      up(np.concatenate([p["leaves"] for p in plans])))
            N.check(self.L.vis_overlay_draw_cn(d_draw[0].data_ptr(), n, 3, 0, d_draw[1].data_ptr(), len(tiles),
                                               d_draw[2].data_ptr(), d_draw[3].data_ptr(), sp), "vis_overlay_draw_cn")
            self.last_launches = 2
        self._keepalive_c = (keep, d_recs, d_draw)
        return canvases

    def status_stamp(self, verdict: str, size=(300, 100)) -> torch.Tensor:
        """``create_status_stamp``: the [height, width, 4] BGRA stamp (transparent background, 4-px border, verdict
        text) as a CUDA tensor; bit-exact against the reference's array."""
        from . import compare as CP
        width, height = int(size[0]), int(size[1])
        if width <= 0 or height <= 0:
            raise ValueError("stamp size must be positive")
        canvas = torch.zeros((height, width, 4), dtype=torch.uint8, device=self.device)
        self.last_launches = self._draw_list(canvas, CP.stamp_commands(verdict, width, height))
        return canvas

    # ------------------------------------------------------------------ JPEG codec stage (nvJPEG, opt-in)
    def jpeg_codec(self, backend: str = "gpu_hybrid"):
        """The engine's nvJPEG codec for ``backend`` (created on first use; chroma upsampling interpolated, the closest
        match to libjpeg-turbo).  Tolerance-specified against the host decoders: see ``jpeg.py``."""
        from .jpeg import JpegCodec
        cache = self._tls.__dict__.setdefault("jpeg", {})      # one codec per THREAD and backend (Streamlit sessions
        if backend not in cache:                               # run on separate threads; a VisJpeg handle must not be shared)
            cache[backend] = JpegCodec(self.device, backend, True)
        return cache[backend]

    # ------------------------------------------------------------------ image quality statistics
    def quality_stats(self, frames):
        """BGR uint8 HWC CUDA frames (``[B,H,W,3]`` tensor or list) -> (int64 CUDA tensor [B, 3] = sum(gray),
        sum(laplacian), sum(laplacian^2) per frame, exact; list of (H, W)).  Two launches for the batch."""
        uniform = isinstance(frames, torch.Tensor) and frames.dim() == 4
        if uniform:                                   # a batch tensor: one check, descriptors by arithmetic
            self._check_u8(frames)
            if frames.shape[0] == 0:
                raise ValueError("no frames")
            if frames.shape[3] != 3 or frames.stride(3) != 1 or frames.stride(2) != 3:
                raise ValueError("frames must be [B, H, W, 3] uint8 with contiguous pixels")
            n = int(frames.shape[0])
            shapes = [(int(frames.shape[1]), int(frames.shape[2]))] * n
            key = (frames.data_ptr(), tuple(frames.shape), frames.stride(0), frames.stride(1))
            cached = self._quality_desc if getattr(self, "_quality_desc", None) and self._quality_desc[0] == key else None
            desc = None
            if cached is None:                        # the same batch tensor again (a streaming loop): descriptors stay on the device
                desc = np.zeros(n, N.QUALITY_FRAME_DTYPE)
                desc["src"] = np.uint64(frames.data_ptr()) + np.arange(n, dtype=np.uint64) * np.uint64(frames.stride(0))
                desc["pitch"], desc["h"], desc["w"] = frames.stride(1), shapes[0][0], shapes[0][1]
            flist = range(n)
        else:
            flist = list(frames)
            if not flist:
                raise ValueError("no frames")
            for f in flist:
                self._check_u8(f)
                if f.dim() != 3 or f.shape[2] != 3 or f.stride(2) != 1 or f.stride(1) != 3:
                    raise ValueError("frames must be [H, W, 3] uint8 with contiguous pixels")
            shapes = [(int(f.shape[0]), int(f.shape[1])) for f in flist]
            desc = np.zeros(len(flist), N.QUALITY_FRAME_DTYPE)
            desc["src"] = [f.data_ptr() for f in flist]
            desc["pitch"] = [f.stride(0) for f in flist]
            desc["h"] = [s_[0] for s_ in shapes]
            desc["w"] = [s_[1] for s_ in shapes]
        if uniform and cached is not None:
            d_desc = cached[1]
        else:
            d_desc = torch.from_numpy(desc.view(np.uint8).copy()).to(self.device)
            if uniform:
                self._quality_desc = (key, d_desc)
        sums = torch.empty((len(flist), 3), dtype=torch.int64, device=self.device)
        out = []
        for b0 in range(0, len(flist), 65535):
            n = min(65535, len(flist) - b0)
            mh, mw = (shapes[0] if uniform else (max(s[0] for s in shapes[b0:b0 + n]), max(s[1] for s in shapes[b0:b0 + n])))
            N.check(self.L.vis_quality_stats(d_desc.data_ptr() + b0 * N.QUALITY_FRAME_DTYPE.itemsize, n, mh, mw,
                                             sums.data_ptr() + b0 * 24, _stream_ptr()), "vis_quality_stats")
            out.append(n)
        self.last_launches = 2 * len(out)            # the ring kernel and the kernel of the remaining strips
        self._keepalive_q = d_desc
        return sums, shapes

    # ------------------------------------------------------------------ helpers
    def _check_u8(self, t: torch.Tensor) -> None:
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.uint8:
            raise TypeError("expected a CUDA uint8 tensor (this engine has no CPU path)")
        if t.device != self.device:
            raise ValueError(f"tensor on {t.device}, engine on {self.device}")


def _locked(fn):
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with self._lock:
            return fn(self, *args, **kwargs)
    return wrapper


# The engine is shared process-wide (get_engine) and the reference's callers may sit on several threads (Streamlit runs
# one script thread per session): every public entry point holds the engine's lock while it plans and enqueues.
for _name in ("resize_batch_u8", "resize_u8", "resize_hp", "reduce_u8", "resize_box_u8", "alpha_premultiply_", "resize_nearest_u8",
              "resize_reducing_u8", "agent_inputs", "plan_batch", "preprocess", "preprocess_dual", "preprocess_host", "preprocess_jpeg", "plan_overlay", "annotate",
              "heatmap", "plan_heatmap", "heatmap_batch", "side_by_side", "side_by_side_batch", "status_stamp", "quality_stats"):
    setattr(Engine, _name, _locked(getattr(Engine, _name)))
del _name

_engines: dict = {}
_engines_lock = threading.Lock()


def get_engine(device=None) -> Engine:
    """Process-wide engine for ``device`` (default: current CUDA device)."""
    if not torch.cuda.is_available():
        raise RuntimeError("vision-inspection-system_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    with _engines_lock:
        e = _engines.get(dev)
        if e is None:
            e = _engines[dev] = Engine(dev)
    return e
