"""Row-band sharding of one frame over the GPUs of a box, one process per GPU (torch.distributed for the plumbing).

Pixels are independent (camera.rs:70-76 has no loop-carried state), so the frame is cut into bands of `band_rows` rows
dealt cyclically to the ranks (band b -> rank b mod G: meshes sit mid-frame, contiguous slabs would not balance).  The
scene (< 2 MB) is replicated and every rank renders its bands with ONE launch (rtc_render_device + rtc_rows).  Two ways
to get the bands to rank 0:

  "peer"   (default when it can be set up) — rank 0's frame buffers are mapped into every rank over NVLink (CUDA IPC) and
           the render kernel stores each pixel STRAIGHT INTO THEM at its frame position (RTC_ROWS_FRAME): the transfer
           rides along with the computation tile by tile, there is no collective on the data path and no reassembly;
           completion is a counter the kernels' last CTAs bump over NVLink and rank 0's stream waits on (no NCCL call
           per frame); frames alternate between two buffers and a second counter keeps a rank from overwriting a frame
           rank 0 has not consumed yet (ShardedRenderer, "peer exchange").
  "gather" — every rank renders into a compact device buffer, `dist.gather` moves the buffers to rank 0 over NCCL and
           one strided copy there interleaves the bands back into frame order.
"""
import math

from ._capi import Rows


class BandPlan:
    """Which frame rows each rank renders, and how the gathered buffers interleave back into the frame."""

    def __init__(self, vsize, world_size, band_rows=8):
        if vsize < 1 or world_size < 1:
            raise ValueError("vsize and world_size must be positive")
        # bands must tile the frame exactly so every rank's compact buffer is a whole number of bands
        br = max(1, int(band_rows))
        while vsize % br:
            br -= 1
        self.vsize, self.world_size, self.band_rows = int(vsize), int(world_size), br
        self.nbands = vsize // br
        self.bands_per_rank = math.ceil(self.nbands / world_size)  # buffers are padded to this many bands

    def rows(self, rank, frame_layout=False):
        return Rows(self.band_rows, rank, self.world_size, Rows.FRAME if frame_layout else Rows.COMPACT)

    def bands_of(self, rank):
        return list(range(rank, self.nbands, self.world_size))

    def local_rows(self, rank):
        return len(self.bands_of(rank)) * self.band_rows

    @property
    def padded_rows(self):
        return self.bands_per_rank * self.band_rows

    def assemble(self, gathered):
        """gathered: [G, padded_rows, W, C] (torch, any device) -> frame [vsize, W, C] in row order (one copy)."""
        g, rows, w, c = gathered.shape
        assert g == self.world_size and rows == self.padded_rows
        v = gathered.view(g, self.bands_per_rank, self.band_rows, w, c).transpose(0, 1)
        return v.reshape(self.bands_per_rank * g * self.band_rows, w, c)[: self.vsize]


def peer_schedule(rank, world_size, f):
    """The stream operations of frame f on `rank` under the peer exchange, in stream order (ShardedRenderer.render runs
    them; tests/test_multi_cpu.py simulates them): rank 0 owns two frame buffers and the `arrive` counter, every other rank
    a `go` counter.
        ("set_go", v)        rank 0 stores v into every rank's go counter: frames < v are consumed
        ("wait_go", v)       the stream waits for this rank's go counter to reach v
        ("render", b)        render this rank's bands into frame buffer b
        ("render_notify", b) the same, and the launch's last CTA adds 1 to arrive[b]
        ("wait_arrive", (b, v)) rank 0's stream waits for arrive[b] to reach v: the frame in buffer b is complete
    arrive is per BUFFER: a rank may run a frame ahead of rank 0, and its early contribution to the next frame must not be
    counted towards this one."""
    ops = []
    if rank == 0:
        if f >= 1:
            ops.append(("set_go", f))
        ops.append(("render", f % 2))
        ops.append(("wait_arrive", (f % 2, (f // 2 + 1) * (world_size - 1))))
    else:
        if f >= 2:  # frame f - 2 lived in this buffer: rank 0 must have started frame f - 1
            ops.append(("wait_go", f - 1))
        ops.append(("render_notify", f % 2))
    return ops


class ShardedRenderer:
    """Camera::render of one frame across `world_size` ranks; the RGBA8 frame lands on rank 0's device."""

    def __init__(self, world, camera, rank=0, world_size=1, device=0, band_rows=8, mode="auto"):
        import torch
        self.torch = torch
        self.world, self.camera, self.rank, self.world_size, self.device = world, camera, rank, world_size, device
        self.plan = BandPlan(camera.vsize, world_size, band_rows)
        dev = torch.device("cuda", device)
        w = camera.hsize
        world.scene(device)  # flatten + upload now, not inside the first frame
        self.mode = "single" if world_size == 1 else mode
        self.frame = self.local = self.gathered = None
        self._shared = []  # (pointer, owner) of every library-shared allocation this rank holds
        self._f = 0        # frames rendered so far (the same number on every rank: render() is collective)
        if self.mode in ("auto", "peer"):
            import torch.distributed as dist
            err = None
            try:
                self._setup_peer(dev, w)
            except Exception as e:  # IPC not permitted in this container, no P2P, ...
                err = f"{type(e).__name__}: {e}"
            # every rank must agree on the path (and every rank has run the same collectives inside _setup_peer)
            ok = torch.tensor([0 if err else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                self.mode = "peer"
            else:
                self._close_shared()
                if mode == "peer":
                    raise RuntimeError("peer exchange could not be set up on every rank: " + (err or "another rank failed"))
                self.peer_error = err
                self.mode = "gather"
        if self.mode == "peer":
            self.rows = self.plan.rows(rank, frame_layout=True)
        else:
            self.rows = self.plan.rows(rank)
            self.local = torch.zeros((self.plan.padded_rows, w, 4), dtype=torch.uint8, device=dev)
            if world_size > 1 and rank == 0:
                self.gathered = torch.empty((world_size, self.plan.padded_rows, w, 4), dtype=torch.uint8, device=dev)

    # ---- peer exchange ---------------------------------------------------------------------------------------------------
    # Rank 0 owns TWO frame buffers (frame f goes to buffer f % 2) and an `arrive` counter per buffer; every rank maps them over NVLink
    # (CUDA IPC) and its render kernel stores each pixel straight into the frame, then adds 1 to `arrive` from its last CTA
    # once the stores are visible system-wide (rtc_render_device_notify).  Rank 0's stream waits for arrive to reach
    # (f // 2 + 1) * (G - 1) with a stream wait-value operation: no collective and no kernel on the completion path.  In the other
    # direction every rank owns a `go` counter that rank 0 maps: at the start of its frame f rank 0 stores f into all of them
    # (everything it enqueued to consume frame f - 1 precedes that store in its stream), and a rank waits for go >= f - 1
    # before it overwrites the buffer frame f - 2 lived in — a rank can run one frame ahead of rank 0, never two.
    def _share_create(self, nbytes):
        import ctypes as C
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        self.camera.api.check(self.camera.api.frame_share_create(self.device, nbytes, C.byref(ptr), handle))
        self._shared.append((ptr.value, 1))
        return ptr.value, handle.raw

    def _share_open(self, handle):
        import ctypes as C
        ptr = C.c_void_p()
        self.camera.api.check(self.camera.api.frame_share_open(self.device, handle, C.byref(ptr)))
        self._shared.append((ptr.value, 0))
        return ptr.value

    def _setup_peer(self, dev, w):
        """Collectives run in the same order on every rank whatever fails locally (a failure travels as None and is raised
        after the last collective)."""
        import torch.distributed as dist
        torch = self.torch
        nbytes = self.camera.vsize * w * 4
        failure = None
        # 1. rank 0: two frames + the arrive counter -> everyone
        box = [None]
        if self.rank == 0:
            try:
                made = [self._share_create(nbytes), self._share_create(nbytes), self._share_create(256)]
                box[0] = [h for _, h in made]
                self._frames = [made[0][0], made[1][0]]
                self._arrive = made[2][0]
            except Exception as e:
                failure = e
        dist.broadcast_object_list(box, src=0)
        if self.rank != 0:
            if box[0] is None:
                failure = RuntimeError("rank 0 could not create the shared frame")
            else:
                try:
                    ptrs = [self._share_open(h) for h in box[0]]
                    self._frames, self._arrive = ptrs[:2], ptrs[2]
                except Exception as e:
                    failure = e
        # 2. every other rank: its go counter -> rank 0
        mine = None
        if self.rank != 0 and failure is None:
            try:
                self._go, mine = self._share_create(256)
            except Exception as e:
                failure = e
        handles = [None] * self.world_size
        dist.all_gather_object(handles, mine)
        if self.rank == 0 and failure is None:
            try:
                if any(h is None for h in handles[1:]):
                    raise RuntimeError("a rank could not create its go counter")
                self._go_peers = [self._share_open(h) for h in handles[1:]]
            except Exception as e:
                failure = e
        if failure is not None:
            raise failure
        for p, owner in self._shared:  # counters start at zero, frames black
            if owner:
                torch.cuda.synchronize(self.device)
        if self.rank == 0:
            class _Iface:
                def __init__(self, ptr, shape):
                    self.__cuda_array_interface__ = {"shape": shape, "typestr": "|u1", "data": (ptr, False), "version": 2}
            self._frame_views = [torch.as_tensor(_Iface(p, (self.camera.vsize, w, 4)), device=dev) for p in self._frames]
            for v in self._frame_views:
                v.zero_()
            torch.as_tensor(_Iface(self._arrive, (256,)), device=dev).zero_()
            self.frame = self._frame_views[0]
        else:
            class _Iface:
                def __init__(self, ptr, shape):
                    self.__cuda_array_interface__ = {"shape": shape, "typestr": "|u1", "data": (ptr, False), "version": 2}
            torch.as_tensor(_Iface(self._go, (256,)), device=dev).zero_()
        torch.cuda.synchronize(self.device)

    def _close_shared(self):
        if self._shared:
            self.torch.cuda.synchronize(self.device)
            self.frame = None
            self._frame_views = None
            for ptr, owner in reversed(self._shared):
                self.camera.api.frame_share_close(self.device, ptr, owner)
            self._shared = []

    def close(self):
        """Collective: every rank unmaps before any owner frees."""
        if self._shared and self.world_size > 1:
            import torch.distributed as dist
            self.torch.cuda.synchronize(self.device)
            dist.barrier()
            mapped = [(p, o) for p, o in self._shared if not o]
            owned = [(p, o) for p, o in self._shared if o]
            self.frame = None
            self._frame_views = None
            for ptr, owner in mapped:
                self.camera.api.frame_share_close(self.device, ptr, owner)
            dist.barrier()
            for ptr, owner in owned:
                self.camera.api.frame_share_close(self.device, ptr, owner)
            self._shared = []

    def out_ptr(self):
        return self._frames[self._f % 2] if self.mode == "peer" else self.local.data_ptr()

    def render(self, stats=None, scene=None):
        """One frame.  Returns the [vsize, W, 4] uint8 device tensor on rank 0 (None elsewhere).  Asynchronous on
        torch's current stream unless `stats` is given.  `scene`: an explicit rtc_scene handle (default: the world's)."""
        import ctypes as C
        torch = self.torch
        stream = torch.cuda.current_stream(self.device).cuda_stream
        target = scene if scene is not None else self.world
        if self.mode != "peer":
            self.camera.render_device(target, d_rgba8=self.out_ptr(), rows=self.rows, stream=stream, stats=stats,
                                      device=self.device)
            return self._finish_gather()
        api, f = self.camera.api, self._f
        sp = C.c_void_p(stream) if stream else None
        for op, v in peer_schedule(self.rank, self.world_size, f):
            if op == "set_go":
                peers = (C.c_void_p * len(self._go_peers))(*self._go_peers)
                api.check(api.stream_set_counters(self.device, sp, peers, len(self._go_peers), v))
            elif op == "wait_go":
                api.check(api.stream_wait_counter(self.device, sp, C.c_void_p(self._go), v))
            elif op == "wait_arrive":
                api.check(api.stream_wait_counter(self.device, sp, C.c_void_p(self._arrive + 128 * v[0]), v[1]))
            elif op == "render" or stats is not None:
                self.camera.render_device(target, d_rgba8=self._frames[v], rows=self.rows, stream=stream, stats=stats,
                                          device=self.device)
                if op == "render_notify":  # a counting frame is a synchronous launch without notify: count this rank in
                    self._bump_arrive(sp, v)
            else:
                d = self.camera.desc()
                sc = target.scene(self.device) if hasattr(target, "scene") else target
                api.check(api.render_device_notify(sc, C.byref(d), C.byref(self.rows), C.c_void_p(self._frames[v]), None, sp,
                                                   C.c_void_p(self._arrive + 128 * v)))
        self._f = f + 1
        if self.rank == 0:
            self.frame = self._frame_views[f % 2]
            return self.frame
        return None

    def _bump_arrive(self, sp, buf):
        """arrive[buf] += 1 from the host side of a rank whose launch carried no notify (the counting frame)."""
        import ctypes as C
        d = self.camera.desc()
        empty = Rows(self.rows.band_rows, 1 << 30, self.rows.band_stride, self.rows.layout)  # selects no band: nothing renders
        sc = self.world.scene(self.device)
        self.camera.api.check(self.camera.api.render_device_notify(sc, C.byref(d), C.byref(empty), None, None, sp,
                                                                   C.c_void_p(self._arrive + 128 * buf)))

    def _finish_gather(self):
        if self.mode == "single":
            return self.local[: self.camera.vsize]
        import torch.distributed as dist
        dist.gather(self.local, list(self.gathered.unbind(0)) if self.rank == 0 else None, dst=0)
        self._f += 1
        return self.plan.assemble(self.gathered) if self.rank == 0 else None


def canvas_schedule(rank, world_size, f):
    """The host-side operations of frame f (1-based) on `rank` for a canvas in host memory shared by the ranks
    (SharedCanvasRenderer.render runs them; tests/test_multi_cpu.py simulates them under adversarial interleavings).  The
    segment holds ONE canvas, a `consumed` counter rank 0 owns and a `done` counter per rank:
        ("set_consumed", v)   rank 0: the caller asked for the next frame, so frames <= v are no longer being read
        ("wait_consumed", v)  the rank waits until frames <= v are consumed before it overwrites their pixels
        ("render",)           rtc_render with RTC_ROWS_FRAME into the canvas: returns when this rank's bands have landed
        ("set_done", v)       this rank's bands of frame v are in the canvas
        ("wait_done", v)      rank 0 waits for every other rank's done counter to reach v
    A rank can finish frame f before rank 0 has even started it, but never touches the canvas of a frame rank 0's caller may
    still be reading."""
    ops = []
    if rank == 0:
        ops.append(("set_consumed", f - 1))
    else:
        ops.append(("wait_consumed", f - 1))
    ops.append(("render",))
    ops.append(("set_done", f))
    if rank == 0 and world_size > 1:
        ops.append(("wait_done", f))
    return ops


class SharedCanvasRenderer:
    """Camera::render of one frame across `world_size` ranks (one process per GPU) with the Canvas in HOST memory shared
    by the processes: every rank renders its bands and its own copy engine writes them into the canvas over its own PCIe
    link (rtc_render + RTC_ROWS_FRAME into an rtc_host_share segment) — the frame never funnels through one GPU and one
    link, and there is no collective per frame: completion is a counter per rank in the segment itself.

    want_f64: the canvas holds the f64 colours (24 B/px — the reference's Canvas, canvas.rs:8); want_rgba8: the quantised
    RGBA8 frame (4 B/px).  Rank 0's render() returns numpy views of them; they stay valid until its next render().
    `name`: the segment's name when the launcher hands it out itself; by default rank 0 invents one and broadcasts it with
    torch.distributed (set-up only).  page_lock=False maps without cudaHostRegister (CPU tests of the protocol)."""

    HEADER = 4096  # counters: consumed at byte 0, done[r] at byte 64 * (r + 1)

    def __init__(self, world, camera, rank=0, world_size=1, device=0, band_rows=8, want_f64=True, want_rgba8=False,
                 name=None, timeout_s=120.0, page_lock=True):
        import ctypes as C
        if not (want_f64 or want_rgba8):
            raise ValueError("a canvas needs its f64 colours, its RGBA8 pixels, or both")
        if world_size > (self.HEADER // 64) - 2:
            raise ValueError("too many ranks for the counter block")
        self.world, self.camera, self.rank, self.world_size, self.device = world, camera, rank, world_size, device
        self.api = camera.api
        self.plan = BandPlan(camera.vsize, world_size, band_rows)
        self.rows = self.plan.rows(rank, frame_layout=True)
        self.timeout_s = float(timeout_s)
        px = camera.hsize * camera.vsize
        page = 4096
        self._off64 = self.HEADER
        self._off8 = self._off64 + ((px * 24 + page - 1) // page * page if want_f64 else 0)
        self.nbytes = self._off8 + ((px * 4 + page - 1) // page * page if want_rgba8 else 0)
        self.want_f64, self.want_rgba8 = bool(want_f64), bool(want_rgba8)
        self._base = None
        self._f = 0
        dev_arg = device if page_lock else -1
        failure = None
        if name is None and world_size > 1:
            import torch.distributed as dist
            box = [None]
            if rank == 0:
                name = self._fresh_name()
                try:
                    self._map(dev_arg, name, create=True)
                    box[0] = name
                except Exception as e:  # /dev/shm too small, registration refused, ...
                    failure = e
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                if box[0] is None:
                    failure = RuntimeError("rank 0 could not create the shared canvas")
                else:
                    name = box[0]
                    try:
                        self._map(dev_arg, name, create=False)
                    except Exception as e:
                        failure = e
            # every rank learns whether every rank has the canvas (same collectives on every rank whatever failed locally)
            oks = [None] * world_size
            dist.all_gather_object(oks, failure is None)
            self.name = name
            self._owner = rank == 0 and box[0] is not None
            if not all(oks):
                self._unmap(unlink=self._owner)
                raise RuntimeError("the shared canvas could not be set up on every rank: "
                                   + (f"{type(failure).__name__}: {failure}" if failure else "another rank failed"))
        else:
            self.name = name or self._fresh_name()
            self._owner = rank == 0
            self._map(dev_arg, self.name, create=self._owner)
        self._C = C

    @staticmethod
    def _fresh_name():
        import os
        import secrets
        return f"/rtc_canvas_{os.getpid()}_{secrets.token_hex(6)}"

    def _map(self, dev_arg, name, create):
        import ctypes as C
        p = C.c_void_p()
        fn = self.api.host_share_create if create else self.api.host_share_open
        self.api.check(fn(dev_arg, name.encode(), self.nbytes, C.byref(p)))
        self._base = p.value

    def _unmap(self, unlink):
        if self._base is not None:
            self.api.host_share_close(self._base, self.nbytes, self.name.encode() if unlink and self.name else None)
            self._base = None

    def _counter(self, slot):
        return self._C.c_void_p(self._base + 64 * slot)

    def views(self):
        """numpy views (f64 [vsize, W, 3] or None, rgba8 [vsize, W, 4] or None) of the shared canvas."""
        import ctypes as C
        import numpy as np
        h, w = self.camera.vsize, self.camera.hsize
        f64 = rgba = None
        if self.want_f64:
            f64 = np.ctypeslib.as_array((C.c_double * (h * w * 3)).from_address(self._base + self._off64)).reshape(h, w, 3)
        if self.want_rgba8:
            rgba = np.ctypeslib.as_array((C.c_uint8 * (h * w * 4)).from_address(self._base + self._off8)).reshape(h, w, 4)
        return f64, rgba

    def _render_rows(self, scene, stats):
        C = self._C
        target = scene if scene is not None else self.world
        sc = target.scene(self.device) if hasattr(target, "scene") else target
        d = self.camera.desc()
        self.api.check(self.api.render(sc, C.byref(d), C.byref(self.rows),
                                       C.c_void_p(self._base + self._off8) if self.want_rgba8 else None,
                                       C.c_void_p(self._base + self._off64) if self.want_f64 else None,
                                       C.byref(stats) if stats is not None else None))

    def render(self, scene=None, stats=None, _render=None):
        """One frame; collective in the sense that every rank must call it the same number of times.  Returns (f64, rgba8)
        views on rank 0 once every rank's bands have landed, None elsewhere.  `stats`: this rank's counters (one launch)."""
        api, f = self.api, self._f + 1
        for op in canvas_schedule(self.rank, self.world_size, f):
            if op[0] == "set_consumed":
                api.host_counter_store(self._counter(0), op[1])
            elif op[0] == "wait_consumed":
                api.check(api.host_counter_wait(self._counter(0), op[1], self.timeout_s))
            elif op[0] == "render":
                (_render or self._render_rows)(scene, stats)
            elif op[0] == "set_done":
                api.host_counter_store(self._counter(1 + self.rank), op[1])
            elif op[0] == "wait_done":
                for r in range(1, self.world_size):
                    api.check(api.host_counter_wait(self._counter(1 + r), op[1], self.timeout_s))
        self._f = f
        return self.views() if self.rank == 0 else None

    AUX_SLOT = 63  # a spare counter of the header for the caller's own host-side hand-shakes (signal / wait_signal)

    def signal(self, value):
        """Stores `value` into the segment's spare counter (release)."""
        self.api.host_counter_store(self._counter(self.AUX_SLOT), int(value))

    def wait_signal(self, value, timeout_s=None):
        """Waits on the CPU — no GPU work, unlike an NCCL barrier, whose kernel spins on the device — until the spare
        counter reaches `value`."""
        self.api.check(self.api.host_counter_wait(self._counter(self.AUX_SLOT), int(value),
                                                  self.timeout_s if timeout_s is None else float(timeout_s)))

    def close(self):
        """Collective when set up through torch.distributed: every rank unmaps before rank 0 unlinks the name."""
        if self._base is None:
            return
        if self.world_size > 1:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    dist.barrier()
            except ImportError:
                pass
        self._unmap(unlink=self._owner)

    def __del__(self):  # a renderer dropped without close(): unmap, and the creator removes the name (no collective here)
        try:
            if getattr(self, "_base", None) is not None:
                self._unmap(unlink=getattr(self, "_owner", False))
        except Exception:
            pass
