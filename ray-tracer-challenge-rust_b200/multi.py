"""Row-band sharding of one frame over the GPUs of a box, one process per GPU (torch.distributed for the plumbing).

Pixels are independent (camera.rs:70-76 has no loop-carried state), so the frame is cut into bands of `band_rows` rows
dealt cyclically to the ranks (band b -> rank b mod G: meshes sit mid-frame, contiguous slabs would not balance).  The
scene (< 2 MB) is replicated and every rank renders its bands with ONE launch (rtc_render_device + rtc_rows).  Two ways
to get the bands to rank 0:

  "peer"   (default when it can be set up) — rank 0's frame buffer is mapped into every rank over NVLink (CUDA IPC) and
           the render kernel stores each pixel STRAIGHT INTO IT at its frame position (RTC_ROWS_FRAME): the transfer
           rides along with the computation tile by tile, there is no collective on the data path and no reassembly;
           one stream-ordered barrier tells rank 0 the frame is complete.
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
        if self.mode in ("auto", "peer"):
            try:
                self._setup_peer(dev, w)
                self.mode = "peer"
            except Exception as e:  # IPC not permitted in this container, no P2P, ...
                if mode == "peer":
                    raise
                self.peer_error = f"{type(e).__name__}: {e}"
                self.mode = "gather"
            # every rank must agree on the path
            import torch.distributed as dist
            ok = torch.tensor([1 if self.mode == "peer" else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self.mode = "gather"
        if self.mode == "peer":
            self.rows = self.plan.rows(rank, frame_layout=True)
        else:
            self.rows = self.plan.rows(rank)
            self.local = torch.zeros((self.plan.padded_rows, w, 4), dtype=torch.uint8, device=dev)
            if world_size > 1 and rank == 0:
                self.gathered = torch.empty((world_size, self.plan.padded_rows, w, 4), dtype=torch.uint8, device=dev)

    def _setup_peer(self, dev, w):
        """Rank 0 allocates the frame through the library (plain cudaMalloc + CUDA IPC handle); every other rank maps it
        with ITS device current, which also enables peer access, so its render kernel can store through the pointer."""
        import ctypes as C
        import torch.distributed as dist
        torch = self.torch
        api = self.camera.api
        nbytes = self.camera.vsize * w * 4
        box = [None]
        ptr = C.c_void_p()
        if self.rank == 0:
            handle = C.create_string_buffer(64)
            api.check(api.frame_share_create(self.device, nbytes, C.byref(ptr), handle))
            box[0] = handle.raw
        dist.broadcast_object_list(box, src=0)
        if self.rank != 0:
            api.check(api.frame_share_open(self.device, box[0], C.byref(ptr)))
        self._frame_ptr = ptr.value
        self._frame_owner = self.rank == 0
        if self.rank == 0:  # a torch view of the library-owned buffer (for the host copy / PPM encoder)
            class _Iface:
                __cuda_array_interface__ = {"shape": (self.camera.vsize, w, 4), "typestr": "|u1",
                                            "data": (ptr.value, False), "version": 2}
            self.frame = torch.as_tensor(_Iface(), device=dev)
            self.frame.zero_()
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)

    def close(self):
        if getattr(self, "_frame_ptr", None):
            self.torch.cuda.synchronize(self.device)
            self.frame = None
            self.camera.api.frame_share_close(self.device, self._frame_ptr, int(self._frame_owner))
            self._frame_ptr = None

    def out_ptr(self):
        return self._frame_ptr if self.mode == "peer" else self.local.data_ptr()

    def finish(self):
        """After this rank's launch: make the frame complete on rank 0.  Returns it there (None elsewhere)."""
        if self.mode == "single":
            return self.local[: self.camera.vsize]
        import torch.distributed as dist
        if self.mode == "peer":
            dist.all_reduce(self._flag)  # stream-ordered barrier: every rank's stores precede rank 0's completion
            return self.frame if self.rank == 0 else None
        dist.gather(self.local, list(self.gathered.unbind(0)) if self.rank == 0 else None, dst=0)
        return self.plan.assemble(self.gathered) if self.rank == 0 else None

    def render(self, stats=None, scene=None):
        """One frame.  Returns the [vsize, W, 4] uint8 device tensor on rank 0 (None elsewhere).  Asynchronous on
        torch's current stream unless `stats` is given.  `scene`: an explicit rtc_scene handle (default: the world's)."""
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.camera.render_device(scene if scene is not None else self.world, d_rgba8=self.out_ptr(), rows=self.rows,
                                  stream=stream, stats=stats, device=self.device)
        return self.finish()
