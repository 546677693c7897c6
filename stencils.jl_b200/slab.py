"""Slab-partitioned iterated sweeps over the GPUs of one box (one process per GPU).

The global array is split into contiguous slabs along its LAST (slowest) axis; rank r owns slab r. Every rank
keeps G >= R ghost planes on each side of its slab inside the same parent buffer; a sweep treats the whole parent
as one unpadded array and writes only the planes that are still exact (`region_lo/region_hi` of sb200_desc), so
the kernels read the ghost planes as ordinary neighbours, while the other axes keep the array's own boundary
condition, resolved inside the kernels.

Ghost planes are exchanged once every k = G // R steps ("wide halo"): after an exchange the ghost planes are
exact copies of the neighbours' cells, and step s of the cycle recomputes the planes that are still exact
([R*s, ext - R*s) of the parent), so after k steps exactly the owned planes are valid again. The results are
bit-identical to the single-domain sweep for any number of ranks and any G. (The global ends of the split axis
under Remove / Reflect are re-imposed by the end ranks after every step: a recomputed mirror image would fold
its neighbours in the opposite order and differ in the last bit.)

During the last step of a cycle the planes the neighbours need are computed first; their exchange (NCCL
send/recv over NVLink, on a side stream) overlaps the interior update.

The same class runs on CPU tensors with the gloo backend and an injected `compute` callable — that is how the
decomposition / exchange logic is tested without GPUs (tests/test_slab_gloo.py).
"""
from __future__ import annotations

import os

import numpy as np

from . import _abi as A
from ._desc import build_desc


class PeerMailbox:
    """Ghost-plane exchange over peer memory (NVLink P2P stores) instead of NCCL send/recv.

    Every rank owns one device allocation (sb200_malloc, exported with a CUDA IPC handle and opened by its two
    neighbours): two flag words and, per side (0 = planes arriving from the rank below, 1 = from the rank above),
    two landing slots of G planes used alternately. An exchange is, per neighbour, ONE kernel on the sender
    (sb200_push_planes: 128-bit stores of its boundary planes straight into the neighbour's landing slot, then a
    system-scope release of the neighbour's flag) and a stream-ordered acquire wait + local copy on the receiver
    (sb200_wait_flag, sb200_memcpy_d2d). Two slots suffice without credits: a rank can only start exchange c + 2
    after it has received its neighbour's planes of exchange c + 1, which the neighbour sent after it had emptied
    slot c."""
    FLAGS = 256  # bytes reserved for the flag words

    def __init__(self, plane_bytes, G, rank, world, wrap):
        import ctypes as C
        import torch.distributed as dist
        self.C = C
        self.lib = A.lib()
        self.nbytes = int(plane_bytes) * G            # one landing slot
        self.rank, self.world = rank, world
        total = self.FLAGS + 4 * self.nbytes
        self.base, self.peer, self.seq = 0, {}, 0
        self.up = (rank + 1) % world if (wrap or rank < world - 1) else None
        self.down = (rank - 1) % world if (wrap or rank > 0) else None
        # every rank takes part in both collectives below whatever happens locally, so a rank whose driver refuses
        # IPC cannot leave the others waiting; `ok` is agreed on by all ranks (SlabIterator falls back to NCCL)
        mine = None
        try:
            base = C.c_void_p()
            A.check(self.lib.sb200_malloc(C.byref(base), total))
            self.base = base.value
            A.check(self.lib.sb200_memset(base, 0, total, None))
            A.check(self.lib.sb200_stream_sync(None))
            handle = (C.c_ubyte * 64)()
            A.check(self.lib.sb200_ipc_export(base, handle))
            mine = bytes(handle)
        except Exception:
            mine = None
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        good = all(h is not None for h in handles)
        if good:
            try:
                for r in {self.up, self.down} - {None}:
                    if r == rank:
                        self.peer[r] = self.base
                        continue
                    ptr = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                    A.check(self.lib.sb200_ipc_import(buf, C.byref(ptr)))
                    self.peer[r] = ptr.value
            except Exception:
                good = False
        votes = [None] * world
        dist.all_gather_object(votes, good)
        self.ok = all(votes)
        if not self.ok:
            self.close()

    def slot(self, base, side, parity):
        return base + self.FLAGS + (2 * side + parity) * self.nbytes

    def flag(self, base, side):
        return base + 4 * side

    def push(self, top_ptr, bottom_ptr, stream):
        """Send my top owned planes up and my bottom owned planes down (exchange number self.seq + 1)."""
        self.seq += 1
        par = self.seq & 1
        if self.up is not None:   # arrives at `up` from below: side 0
            pb = self.peer[self.up]
            A.check(self.lib.sb200_push_planes(top_ptr, self.slot(pb, 0, par), self.nbytes, self.flag(pb, 0), self.seq, stream))
        if self.down is not None:  # arrives at `down` from above: side 1
            pb = self.peer[self.down]
            A.check(self.lib.sb200_push_planes(bottom_ptr, self.slot(pb, 1, par), self.nbytes, self.flag(pb, 1), self.seq, stream))

    def signal(self, seq, stream):
        """Publish exchange `seq` whose planes the sweep kernels already stored into the neighbours' slots."""
        self.seq = seq
        if self.up is not None:
            A.check(self.lib.sb200_signal_flag(self.flag(self.peer[self.up], 0), seq, stream))
        if self.down is not None:
            A.check(self.lib.sb200_signal_flag(self.flag(self.peer[self.down], 1), seq, stream))

    def pull(self, ghost_bottom_ptr, ghost_top_ptr, stream):
        """Wait for exchange self.seq from both neighbours and move the planes into my ghost zones."""
        par = self.seq & 1
        if self.down is not None:  # planes from below -> my bottom ghost
            A.check(self.lib.sb200_wait_flag(self.flag(self.base, 0), self.seq, stream))
            A.check(self.lib.sb200_memcpy_d2d(ghost_bottom_ptr, self.slot(self.base, 0, par), self.nbytes, stream))
        if self.up is not None:    # planes from above -> my top ghost
            A.check(self.lib.sb200_wait_flag(self.flag(self.base, 1), self.seq, stream))
            A.check(self.lib.sb200_memcpy_d2d(ghost_top_ptr, self.slot(self.base, 1, par), self.nbytes, stream))

    def close(self):
        for r, ptr in self.peer.items():
            if r != self.rank:
                self.lib.sb200_ipc_close(ptr)
        self.peer = {}
        if self.base:
            self.lib.sb200_free(self.base)
            self.base = 0


class SlabIterator:
    def __init__(self, local_owned, *, offsets, radius, reducer, boundary, eltype, ghost=None, rank=0, world=1,
                 compute=None, reducer_kwargs=None, padval=0, exchange="auto", overlap=None):
        """local_owned: torch tensor holding this rank's slab in column-major layout, i.e. a C-contiguous torch
        tensor of shape reversed(logical shape) (split axis first). boundary: per-axis sb200 enums of the GLOBAL
        array. compute(desc_handle, src_tensor, dst_tensor): sweep backend; None = libstencils_b200 on the
        current CUDA stream."""
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.R = int(radius)
        self.G = int(ghost) if ghost is not None else self.R
        if self.G < self.R or self.G % max(self.R, 1):
            raise A.ArgumentError("ghost thickness must be a positive multiple of the radius")
        self.k = self.G // self.R
        t = local_owned
        self.nd = t.dim()
        self.n_local = t.shape[0]
        if self.n_local < self.G:
            raise A.ArgumentError("slab thinner than the ghost zone")
        self.logical_rest = tuple(reversed(t.shape[1:]))          # sizes of axes 0..nd-2
        self.bcs = tuple(boundary)
        self.bc_split = self.bcs[-1]
        self.padval = padval
        ext = self.n_local + 2 * self.G
        self.ext = ext
        self.bufs = [torch.empty((ext,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for _ in range(2)]
        self.bufs[0][self.G:self.G + self.n_local].copy_(t)
        self.cur = 0
        self.eltype = eltype
        self.compute = compute or self._compute_cuda
        # The sweep descriptor covers the whole parent (ghost planes included) as an unpadded array and restricts the
        # OUTPUT to the planes that are still exact (`region`): those never read outside the parent, so the boundary
        # condition named for the split axis is never exercised. (Measured on B200: the same Life kernel runs 10 %
        # slower back to back through the ring form `src_off = R, boundary = USE` of the same sweep.)
        size = self.logical_rest + (ext,)
        self._mk = lambda region, flags: build_desc(
            size=size, eltype=eltype, out_eltype=eltype, offsets=offsets, radius=self.R,
            boundary=self.bcs[:-1] + (A.WRAP,), reducer=reducer, padval=padval, region=region, flags=flags,
            **(reducer_kwargs or {}))
        # Life on UInt8: after the first sweep every cell this rank reads is a 0/1 output of the kernel (own cells or
        # exchanged ghosts); stale ghost planes outside the still-exact region only feed outputs that are discarded.
        self._later_flags = A.FLAG_CELLS_01 if (reducer == A.LIFE and eltype == A.U8) else 0
        self._reducer = reducer
        # two generations per launch: Life on a device grid whose split axis is a ring (Remove / Reflect ends must be
        # re-imposed after every single generation) — the library has the last word (life2_accepts, csrc/life.cu)
        self._gen = 1
        self.double_ok = (reducer == A.LIFE and t.is_cuda and compute is None and self.bc_split == A.WRAP and self.k >= 2 and
                          self.n_local >= 4 * self.G + 64 and os.environ.get("SB200_DOUBLE_STEP", "1") != "0")
        self.quad_ok = self.double_ok and self.k >= 4 and os.environ.get("SB200_QUAD_STEP", "1") != "0"
        # eight generations per launch (one-halo-lane layout of life_bit_kernel, the default build); SB200_OCT_STEP=0 turns it off
        self.oct_ok = self.quad_ok and self.k >= 8 and os.environ.get("SB200_OCT_STEP", "1") != "0"
        # two diffusion steps per launch (csrc/stream3d2.cu): same schedule; SB200_DIFFUSION_DOUBLE_STEP=0 turns it off
        if (reducer == A.DIFFUSION and t.is_cuda and compute is None and self.bc_split == A.WRAP and self.k >= 2 and
                self.n_local >= 4 * self.G and os.environ.get("SB200_DIFFUSION_DOUBLE_STEP", A.DIFFUSION_DOUBLE_STEP_DEFAULT) != "0"):
            self.double_ok = True
        self._nsweeps = 0
        self._descs = {}
        self.is_cuda = t.is_cuda
        if self.is_cuda:
            self.comm_stream = torch.cuda.Stream(device=t.device)
        # overlap=True: on the last sweep of a cycle the boundary planes are computed first and their exchange runs on a
        # side stream under the interior sweep (three launches); False: one sweep, then the exchange (cheaper when the
        # exchange is tiny next to a sweep, e.g. 2-D rows)
        # None: overlap when a ghost zone is at least 1 MiB (measured on B200: Life rows, 512 KiB per exchange, run 1.3 %
        # faster serialised; 16 MiB diffusion planes hide completely under the interior sweep)
        self.overlap = (t[0].numel() * t.element_size() * self.G >= (1 << 20)) if overlap is None else bool(overlap)
        self.steps_since_exchange = self.k  # ghosts are not valid yet
        self.launches = 0
        # ghost exchange: peer-memory stores over NVLink when the ranks can open each other's memory, else NCCL
        self.mailbox = None
        self.fused = os.environ.get("SB200_FUSED_PUSH", "1") != "0"
        self.exchange = "local" if world == 1 else ("nccl" if self.is_cuda else "gloo")
        if world > 1 and self.is_cuda and exchange in ("auto", "p2p"):
            plane_bytes = t[0].numel() * t.element_size()
            mb = PeerMailbox(plane_bytes, self.G, rank, world, self.bc_split == A.WRAP)
            if mb.ok:
                self.mailbox, self.exchange = mb, "p2p"
            elif exchange == "p2p":
                raise A.SB200Error(A.ECUDA, "CUDA IPC peer access is not available between the ranks")

    # ---- helpers ----
    def _desc(self, lo_plane, hi_plane, mirror=None):
        """descriptor whose output region is parent planes [lo_plane, hi_plane) of the split axis; mirror =
        (peer pointer, first plane, end plane): those planes are also stored into the peer's landing slot by the sweep"""
        flags = (self._later_flags if self._nsweeps > 0 else 0) | {1: 0, 2: A.FLAG_DOUBLE_STEP, 4: A.FLAG_QUAD_STEP, 8: A.FLAG_OCT_STEP}[self._gen]
        key = (lo_plane, hi_plane, flags, mirror)
        if key not in self._descs:
            lo = (0,) * (self.nd - 1) + (lo_plane,)
            hi = self.logical_rest + (hi_plane,)
            lo, hi = lo + (0,) * (3 - self.nd), hi + (0,) * (3 - self.nd)
            h = self._mk((lo, hi), flags)
            if mirror is not None:
                h = h.copy(mirror_parent=mirror[0], mirror_lo=mirror[1], mirror_hi=mirror[2])
            self._descs[key] = h
        return self._descs[key]

    def _compute_cuda(self, h, src, dst):
        stream = self.torch.cuda.current_stream().cuda_stream
        A.check(A.lib().sb200_gather(h.ptr(), src.data_ptr(), dst.data_ptr(), stream))

    def _sweep(self, lo_plane, hi_plane, mirror=None):
        if hi_plane > lo_plane:
            self.compute(self._desc(lo_plane, hi_plane, mirror), self.bufs[self.cur], self.bufs[1 - self.cur])
            self.launches += 1

    def close(self):
        """Release the peer-memory mailbox (IPC mappings + the device allocation)."""
        if self.mailbox is not None:
            self.mailbox.close()
            self.mailbox = None

    @property
    def state(self):
        """The owned planes of the current state (torch tensor view, split axis first)."""
        return self.bufs[self.cur][self.G:self.G + self.n_local]

    # ---- ghost exchange ----
    def _fill_end_ghosts(self, buf):
        """Global ends of the split axis under Remove / Reflect are local operations."""
        G, n = self.G, self.n_local
        if self.bc_split == A.WRAP:
            return
        first, last = self.rank == 0, self.rank == self.world - 1
        if self.bc_split == A.REMOVE:
            if first:
                buf[:G] = self.padval
            if last:
                buf[G + n:] = self.padval
        elif self.bc_split == A.REFLECT:  # i<0 -> -i ; i>=s -> 2(s-1)-i, mirror without repeating the edge
            if first:
                buf[:G] = buf[G + 1:2 * G + 1].flip(0)
            if last:
                buf[G + n:] = buf[n - 1:G + n - 1].flip(0)
        else:
            raise A.ArgumentError("the split axis needs Wrap, Remove or Reflect")

    def _exchange_ops(self, buf):
        import torch.distributed as dist
        G, n, r, w = self.G, self.n_local, self.rank, self.world
        wrap = self.bc_split == A.WRAP
        up, down = (r + 1) % w, (r - 1) % w       # up = owner of the planes above mine
        ops = []
        if wrap or r < w - 1:
            ops.append(dist.P2POp(dist.isend, buf[n:n + G], up))           # my top owned planes -> up's bottom ghost
        if wrap or r > 0:
            ops.append(dist.P2POp(dist.irecv, buf[:G], down))              # my bottom ghost <- down's top planes
        if wrap or r > 0:
            ops.append(dist.P2POp(dist.isend, buf[G:2 * G], down))         # my bottom owned planes -> down's top ghost
        if wrap or r < w - 1:
            ops.append(dist.P2POp(dist.irecv, buf[G + n:], up))            # my top ghost <- up's bottom planes
        return ops

    def _exchange(self, buf):
        """Refresh the ghost planes of `buf` (blocking w.r.t. the stream it is called on)."""
        import torch.distributed as dist
        if self.world == 1:
            G, n = self.G, self.n_local
            if self.bc_split == A.WRAP:
                buf[:G].copy_(buf[n:n + G])
                buf[G + n:].copy_(buf[G:2 * G])
        elif self.mailbox is not None:
            G, n = self.G, self.n_local
            stream = self.torch.cuda.current_stream().cuda_stream
            self.mailbox.push(buf[n:n + G].data_ptr(), buf[G:2 * G].data_ptr(), stream)
            self.mailbox.pull(buf[:G].data_ptr(), buf[G + n:].data_ptr(), stream)
        else:
            ops = self._exchange_ops(buf)
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        self._fill_end_ghosts(buf)

    # ---- stepping ----
    def step(self, nsteps=1):
        torch = self.torch
        left = int(nsteps)
        while left > 0:
            if self.steps_since_exchange >= self.k:
                self._exchange(self.bufs[self.cur])
                self.steps_since_exchange = 0
            # Life: four / two generations per launch (SB200_FLAG_QUAD_STEP / _DOUBLE_STEP) while the cycle has room
            room = self.k - self.steps_since_exchange
            m = 4 if (self.quad_ok and left >= 4 and room >= 4) else (2 if (self.double_ok and left >= 2 and room >= 2) else 1)
            if self.oct_ok and left >= 8 and room >= 8:
                m = 8
            if m > 1:
                try:
                    self._step_one(torch, m)
                except A.ArgumentError:   # the library declined this layout / rule: fewer generations per launch from now on
                    if m == 8:
                        self.oct_ok = False
                    elif m == 4:
                        self.quad_ok = False
                    else:
                        self.double_ok = False
                    continue
            else:
                self._step_one(torch, 1)
            self._nsweeps += 1
            left -= m

    def _step_one(self, torch, m=1):
        self._gen = m
        if True:
            s = self.steps_since_exchange + m                 # generations since the exchange once this sweep is done
            lo, hi = self.R * s, self.ext - self.R * s        # parent planes that are still exact after this sweep
            last_of_cycle = s == self.k
            # (thin boundary sweeps of G + 1 rows are below what the multi-generation Life kernels accept: no overlap for those
            # launches, instead of a rejected launch that used to switch the multi-generation modes off for good — ADVICE r1)
            thin_ok = m == 1 or self._reducer != A.LIFE or self.G + 1 >= 16
            if last_of_cycle and self.is_cuda and self.world > 1 and self.overlap and thin_ok:
                # boundary planes first, their exchange overlaps the interior update
                # (one extra plane per side: a Reflect end mirrors planes G+1 .. 2G of the new state)
                G, n = self.G + 1, self.n_local
                nxt = self.bufs[1 - self.cur]
                mb = self.mailbox
                fused = mb is not None and self.fused and n >= 2 * self.G + 2
                if fused:
                    # the boundary sweeps store the planes the neighbours need straight into their landing slots
                    # (peer memory, NVLink) while they compute them; two 1-thread kernels publish the flags
                    stream = torch.cuda.current_stream().cuda_stream
                    seq = mb.seq + 1
                    par = seq & 1
                    m_dn = (mb.slot(mb.peer[mb.down], 1, par), self.G, 2 * self.G) if mb.down is not None else None
                    m_up = (mb.slot(mb.peer[mb.up], 0, par), n, n + self.G) if mb.up is not None else None
                    self._sweep(lo, min(lo + G, hi), m_dn)
                    self._sweep(max(hi - G, lo + G), hi, m_up)
                    mb.signal(seq, stream)
                else:
                    self._sweep(lo, min(lo + G, hi))
                    self._sweep(max(hi - G, lo + G), hi)
                ev = torch.cuda.Event()
                ev.record()
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ev)
                    if fused:
                        mb.pull(nxt[:self.G].data_ptr(), nxt[self.G + n:].data_ptr(), self.comm_stream.cuda_stream)
                        self._fill_end_ghosts(nxt)
                    else:
                        self._exchange(nxt)
                    done = torch.cuda.Event()
                    done.record()
                self._sweep(lo + G, hi - G)
                torch.cuda.current_stream().wait_event(done)
                self.cur = 1 - self.cur
                self.steps_since_exchange = 0
                return
            self._sweep(lo, hi)
            # Remove / Reflect at the global ends hold at EVERY step (the recomputed ghost planes of an end rank are
            # not the boundary values), so the end ranks refresh them after each sweep; Wrap needs nothing.
            self._fill_end_ghosts(self.bufs[1 - self.cur])
            self.cur = 1 - self.cur
            self.steps_since_exchange = s


class SlabPlan:
    """Thin caller of the C-ABI slab plan (include/stencils_b200.h: sb200_plan_*; csrc/slab_plan.cu + slab_sched.h): the
    library owns the slab parents, mailboxes, streams, events and the cycle schedule; this class only builds the descriptor
    of the UNDIVIDED array and moves handles. Two forms, like the ABI:

      SlabPlan(shape, ..., devices=[0, 1, ...])      one process drives every GPU (what a Julia session does)
      SlabPlan(shape, ..., rank=r, world=w)           one process per GPU (torchrun): IPC handles travel through
                                                      torch.distributed.all_gather_object

    `shape` is the logical (column-major) shape of the undivided array; the last axis is split."""

    def __init__(self, shape, *, offsets, radius, reducer, boundary, eltype, ghost=0, devices=None, rank=None, world=None,
                 reducer_kwargs=None, padval=0, plan_flags=0):
        import ctypes as C
        self.C = C
        self.lib = A.lib()
        self.shape = tuple(int(v) for v in shape)
        self.eltype = eltype
        self.desc = build_desc(size=self.shape, eltype=eltype, out_eltype=eltype, offsets=offsets, radius=radius,
                               boundary=tuple(boundary), reducer=reducer, padval=padval, **(reducer_kwargs or {}))
        self.handle = C.c_void_p()
        self.rank_form = devices is None
        if devices is not None:
            arr = (C.c_int32 * len(devices))(*devices)
            A.check(self.lib.sb200_plan_create(self.desc.ptr(), len(devices), arr, int(ghost), int(plan_flags), C.byref(self.handle)))
        else:
            import torch.distributed as dist
            A.check(self.lib.sb200_plan_create_rank(self.desc.ptr(), int(rank), int(world), int(ghost), int(plan_flags),
                                                    C.byref(self.handle)))
            if world > 1:
                mine = (C.c_ubyte * 64)()
                err = None
                try:
                    A.check(self.lib.sb200_plan_ipc_handle(self.handle, mine))
                except Exception as e:   # every rank still takes part in the collectives below
                    err = repr(e)
                got = [None] * world
                dist.all_gather_object(got, (bytes(mine), err))
                bad = [f"rank {r}: {e}" for r, (_, e) in enumerate(got) if e]
                if not bad:
                    blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(h for h, _ in got))
                    try:
                        A.check(self.lib.sb200_plan_connect(self.handle, blob))
                    except Exception as e:
                        err = repr(e)
                votes = [None] * world
                dist.all_gather_object(votes, err)
                bad += [f"rank {r}: {e}" for r, e in enumerate(votes) if e]
                if bad:
                    self.close()
                    raise A.SB200Error(A.ECUDA, "CUDA IPC peer access is not available between the ranks: " + "; ".join(bad))

    def nslabs(self):
        n = self.C.c_int32()
        A.check(self.lib.sb200_plan_nslabs(self.handle, self.C.byref(n)))
        return n.value

    def slab(self, i=0):
        """(lo, hi, device, device pointer of the first owned plane of the current state)"""
        C = self.C
        lo, hi, dev, ptr = C.c_int64(), C.c_int64(), C.c_int32(), C.c_void_p()
        A.check(self.lib.sb200_plan_slab(self.handle, i, C.byref(lo), C.byref(hi), C.byref(dev), C.byref(ptr)))
        return lo.value, hi.value, dev.value, ptr.value

    def mark_dirty(self):
        A.check(self.lib.sb200_plan_mark_dirty(self.handle))

    def load(self, host_array):
        """host_array: the planes this plan owns (whole array / the rank's slab), column-major NumPy array."""
        a = np.asfortranarray(host_array)
        A.check(self.lib.sb200_plan_load_host(self.handle, a.ctypes.data))

    def store(self, nplanes=None):
        dt = A.DTYPE_OF_ELTYPE[self.eltype]
        if nplanes is None:
            n = self.nslabs()
            nplanes = self.slab(n - 1)[1] - self.slab(0)[0]
        out = np.empty(self.shape[:-1] + (nplanes,), dtype=dt, order="F")
        A.check(self.lib.sb200_plan_store_host(self.handle, out.ctypes.data))
        return out

    def iterate(self, nsteps):
        A.check(self.lib.sb200_plan_iterate(self.handle, int(nsteps)))

    def sync(self):
        A.check(self.lib.sb200_plan_sync(self.handle))

    def iterate_timed(self, nsteps):
        ms = self.C.c_float()
        A.check(self.lib.sb200_plan_iterate_timed(self.handle, int(nsteps), self.C.byref(ms)))
        return ms.value

    def stats(self):
        out = (self.C.c_int64 * 8)()
        A.check(self.lib.sb200_plan_stats(self.handle, out))
        keys = ("generations", "launches", "exchanges", "ghost_planes", "generations_per_exchange", "overlap", "max_generations_per_launch", "sync")
        d = dict(zip(keys, [int(v) for v in out]))
        d["sync"] = {1: "events", 2: "flags"}.get(d["sync"], d["sync"])
        return d

    def close(self):
        if self.handle:
            self.lib.sb200_plan_destroy(self.handle)
            self.handle = self.C.c_void_p()


def slab_schedule(radius, ghost, n_min, *, split_wrap=True, overlap=False, max_gens=1, min_planes_multi=0, since=None, first_sweep=True,
                  nsteps=1):
    """The op list a slab plan runs for `nsteps` generations (sb200_slab_schedule: pure host logic, needs no GPU).
    Returns (list of op dicts, generations since the last exchange afterwards)."""
    import ctypes as C
    lib = A.lib()
    if since is None:
        since = ghost // radius
    cap = 16 * (nsteps + 2)
    ops = (A.SlabOp * cap)()
    count, since_out = C.c_int32(), C.c_int32()
    A.check(lib.sb200_slab_schedule(radius, ghost, n_min, int(split_wrap), int(overlap), max_gens, min_planes_multi, since, int(first_sweep),
                                    nsteps, ops, cap, C.byref(count), C.byref(since_out)))
    assert count.value <= cap
    out = [dict(kind=o.kind, gens=o.gens, mirror=o.mirror, buf=o.buf, async_=o.async_, first=o.first, lo=o.lo, hi=o.hi) for o in ops[:count.value]]
    return out, since_out.value


def split_axis_last(shape_global, world, rank):
    """Planes [lo, hi) of the last axis owned by `rank` (equal slabs, remainder to the first ranks)."""
    n = shape_global[-1]
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ------------------------------------------------------------------------------------------------ one-shot sweeps
def slab_gather(local_owned, *, offsets, radius, reducer, boundary, eltype, rank, world, compute=None,
                reducer_kwargs=None, padval=0, exchange="auto"):
    """One-shot (non-iterated) gather over slabs — `mapstencil(f, A)` with A split along its last axis (SURVEY §8e:
    BASELINE configs 1, 3, 4a at N GPUs): ONE pre-exchange of R ghost planes, then one sweep of the owned planes.
    Returns this rank's slab of the result (torch tensor, split axis first). Reducers that change the element type
    (mean / sum of Bool, mean of integers) are not taken here: the state buffers are typed like the source."""
    changes_eltype = (reducer == A.SUM and eltype == A.BOOL) or (reducer == A.MEAN and eltype not in (A.F32, A.F64))
    if changes_eltype:   # sum(Bool) -> Int64, mean(integer) -> Float64 (src/gatherstencil.jl:41-59)
        raise A.ArgumentError("slab_gather needs a reducer whose result has the element type of the source")
    it = SlabIterator(local_owned, offsets=offsets, radius=radius, reducer=reducer, boundary=boundary, eltype=eltype,
                      ghost=max(int(radius), 1), rank=rank, world=world, compute=compute, reducer_kwargs=reducer_kwargs,
                      padval=padval, exchange=exchange, overlap=False)
    try:
        it.step(1)
        return it.state.clone()
    finally:
        it.close()


def split_columns_for_scatter(ncols, world, rank, radius):
    """Columns [lo, hi) of the last axis owned by `rank` for slab_scatter: slab boundaries are multiples of the pass
    stride 2R+1 of _scatterstencil_cpu! (src/scatterstencil.jl:53-58), so a column's pass number is the same in the
    slab's local numbering as in the global one."""
    S = 2 * int(radius) + 1
    units = -(-int(ncols) // S)
    base, rem = divmod(units, world)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(lo_u * S, ncols), min(hi_u * S, ncols)


def slab_scatter(local_src, local_dst, *, ncols_global, offsets, radius, weights, boundary, eltype, rank, world,
                 scatter_op=A.OP_ADD, scatter_rule=A.SCATTER_CENTER_WEIGHTS, flags=0, compute=None):
    """scatterstencil!(f, op, dest, source) (src/scatterstencil.jl:36-112) with source and dest split along the last
    axis into the column ranges of split_columns_for_scatter (SURVEY §8e, config 4b at N GPUs).

    A destination cell folds, in the reference's (pass, row, k) order, the values scattered to it by the source
    columns within R of its own column. Every rank therefore receives 2R+1 ghost SOURCE columns from each neighbour
    (one point-to-point exchange; a ring when the split axis wraps), runs the ordinary single-domain scatter over
    [ghost | owned | ghost] and keeps the owned destination columns: the ghost destination columns collect partial
    folds and are dropped. Because slab boundaries and the ghost thickness are multiples of 2R+1, local and global
    pass numbers agree and the result is bit-identical to the single-domain scatter. At the global ends of a Remove /
    Reflect axis the parent simply ends (no ghost), so the end rank applies the real boundary rule; interior parent
    ends use the same rule, which only ever touches ghost destination columns. local_src / local_dst: torch tensors of
    shape (owned columns, rows), C-contiguous (= column-major (rows, columns)); local_dst is updated in place."""
    import torch
    import torch.distributed as dist
    R = int(radius)
    S = 2 * R + 1
    G = S
    bc0, bc_split = boundary
    wrap = bc_split == A.WRAP
    n = local_src.shape[0]
    if local_dst.shape != local_src.shape:
        raise A.ArgumentError("Source array sizes must match: dest slab and source slab differ")
    if wrap and ncols_global % S:
        raise A.ArgumentError(f"slab_scatter with Wrap on the split axis needs a column count that is a multiple of 2R+1 = {S} "
                              "(the reference's pass order at the wrap seam is only defined then)")
    has_lo, has_hi = wrap or rank > 0, wrap or rank < world - 1
    if world > 1 and n < G:
        raise A.ArgumentError(f"slab of {n} columns is thinner than the ghost zone ({G})")
    glo, ghi = (G if has_lo else 0), (G if has_hi else 0)
    ext = glo + n + ghi
    src = torch.empty((ext,) + tuple(local_src.shape[1:]), dtype=local_src.dtype, device=local_src.device)
    dst = torch.zeros_like(src)
    src[glo:glo + n].copy_(local_src)
    dst[glo:glo + n].copy_(local_dst)
    if world > 1:
        up, down = (rank + 1) % world, (rank - 1) % world
        ops = []
        if has_hi:
            ops.append(dist.P2POp(dist.isend, src[glo + n - G:glo + n], up))
        if has_lo:
            ops.append(dist.P2POp(dist.irecv, src[:G], down))
        if has_lo:
            ops.append(dist.P2POp(dist.isend, src[glo:glo + G], down))
        if has_hi:
            ops.append(dist.P2POp(dist.irecv, src[glo + n:], up))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    elif wrap:
        src[:G].copy_(local_src[n - G:])
        src[glo + n:].copy_(local_src[:G])
    rows = int(local_src.shape[1])
    local_bc = A.REFLECT if bc_split == A.REFLECT else A.REMOVE
    h = build_desc(size=(rows, ext), eltype=eltype, out_eltype=eltype, offsets=offsets, radius=R, boundary=(bc0, local_bc),
                   weights=weights, scatter_op=scatter_op, scatter_rule=scatter_rule, flags=flags)
    if compute is None:
        stream = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().sb200_scatter(h.ptr(), src.data_ptr(), dst.data_ptr(), stream))
    else:
        compute(h, src, dst)
    local_dst.copy_(dst[glo:glo + n])
    return local_dst
