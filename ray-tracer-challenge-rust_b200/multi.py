"""Row-band sharding of one frame over the GPUs of a box, one process per GPU (torch.distributed for the plumbing).

Pixels are independent (camera.rs:70-76 has no loop-carried state), so the frame is cut into bands of `band_rows` rows
dealt cyclically to the ranks (band b -> rank b mod G: meshes sit mid-frame, contiguous slabs would not balance).  The
scene (< 2 MB) is replicated; every rank renders its bands into a compact device buffer with ONE launch
(rtc_render_device + rtc_rows) and the only exchange is the gather of those buffers to rank 0 over NCCL / NVLink,
followed by one strided copy there that interleaves the bands back into frame order.
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

    def rows(self, rank):
        return Rows(self.band_rows, rank, self.world_size)

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

    def __init__(self, world, camera, rank=0, world_size=1, device=0, band_rows=8):
        import torch
        self.torch = torch
        self.world, self.camera, self.rank, self.world_size, self.device = world, camera, rank, world_size, device
        self.plan = BandPlan(camera.vsize, world_size, band_rows)
        self.rows = self.plan.rows(rank)
        dev = torch.device("cuda", device)
        w = camera.hsize
        self.local = torch.zeros((self.plan.padded_rows, w, 4), dtype=torch.uint8, device=dev)
        self.gathered = None
        if world_size > 1 and rank == 0:
            self.gathered = torch.empty((world_size, self.plan.padded_rows, w, 4), dtype=torch.uint8, device=dev)
        world.scene(device)  # flatten + upload now, not inside the first frame

    def render(self, stats=None):
        """One frame.  Returns the [vsize, W, 4] uint8 device tensor on rank 0 (None elsewhere).  Asynchronous on
        torch's current stream unless `stats` is given."""
        torch = self.torch
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.camera.render_device(self.world, d_rgba8=self.local.data_ptr(), rows=self.rows, stream=stream,
                                  stats=stats, device=self.device)
        if self.world_size == 1:
            return self.local[: self.camera.vsize]
        import torch.distributed as dist
        dist.gather(self.local, list(self.gathered.unbind(0)) if self.rank == 0 else None, dst=0)
        if self.rank == 0:
            return self.plan.assemble(self.gathered)
        return None
