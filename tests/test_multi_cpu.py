"""Row-band sharding (multi.py): the plan, and the world_size-2 gather + reassembly over gloo on CPU."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

multi = importlib.import_module("ray-tracer-challenge-rust_b200.multi")


@pytest.mark.parametrize("vsize,g,br", [(1080, 8, 8), (1080, 2, 8), (4320, 8, 8), (117, 4, 8), (200, 3, 16), (7, 8, 4),
                                        (1, 1, 8), (2160, 4, 8)])
def test_band_plan_partitions_the_rows(vsize, g, br):
    plan = multi.BandPlan(vsize, g, br)
    assert vsize % plan.band_rows == 0 and plan.band_rows <= br
    seen = []
    for r in range(g):
        rows = plan.rows(r)
        assert (rows.band_rows, rows.band_first, rows.band_stride) == (plan.band_rows, r, g)
        for b in plan.bands_of(r):
            seen += list(range(b * plan.band_rows, (b + 1) * plan.band_rows))
        assert plan.local_rows(r) <= plan.padded_rows
    assert sorted(seen) == list(range(vsize))
    # assemble() inverts the dealing: fill each rank's compact buffer with its frame-row numbers
    w = 3
    gathered = torch.full((g, plan.padded_rows, w, 4), -1, dtype=torch.int32)
    for r in range(g):
        k = 0
        for b in plan.bands_of(r):
            for i in range(plan.band_rows):
                gathered[r, k] = b * plan.band_rows + i
                k += 1
    frame = plan.assemble(gathered)
    assert frame.shape == (vsize, w, 4)
    assert torch.equal(frame[:, 0, 0], torch.arange(vsize, dtype=torch.int32))


def _worker(rank, world_size, port, vsize, width, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    plan = multi.BandPlan(vsize, world_size, 8)
    # stand-in for the render kernel: every rank fills its compact band buffer with (frame row, column, rank, 255)
    local = torch.zeros((plan.padded_rows, width, 4), dtype=torch.uint8)
    k = 0
    for b in plan.bands_of(rank):
        for i in range(plan.band_rows):
            row = b * plan.band_rows + i
            local[k, :, 0] = row % 251
            local[k, :, 1] = torch.arange(width) % 256
            local[k, :, 2] = rank
            local[k, :, 3] = 255
            k += 1
    gathered = torch.empty((world_size, plan.padded_rows, width, 4), dtype=torch.uint8) if rank == 0 else None
    dist.gather(local, list(gathered.unbind(0)) if rank == 0 else None, dst=0)
    if rank == 0:
        np.save(out_path, plan.assemble(gathered).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("vsize", [64, 120])
def test_gloo_world_size_2_gather_reassembles_the_frame(tmp_path, vsize):
    port = 29500 + (os.getpid() + vsize) % 2000
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, port, vsize, 16, out), nprocs=2, join=True)
    frame = np.load(out)
    assert frame.shape == (vsize, 16, 4)
    rows = np.arange(vsize)
    assert np.array_equal(frame[:, 0, 0], rows % 251)
    assert np.array_equal(frame[:, :, 1], np.tile(np.arange(16), (vsize, 1)))
    plan = multi.BandPlan(vsize, 2, 8)
    assert np.array_equal(frame[:, 0, 2], (rows // plan.band_rows) % 2)  # band b came from rank b mod 2
    assert (frame[:, :, 3] == 255).all()


@pytest.mark.parametrize("g", [2, 4, 8])
@pytest.mark.parametrize("seed", range(6))
def test_peer_exchange_schedule_never_tears_a_frame(g, seed):
    """multi.peer_schedule under adversarial interleavings: every rank's stream runs its operations in order, the ranks at
    arbitrary relative speed (a seeded scheduler picks which stream advances; a wait blocks only its own stream).  Rank 0
    consumes every frame right after its wait_arrive (in its own stream, as bench.py's device->host copy does).  No schedule
    may deadlock, hand rank 0 an incomplete frame, or let a rank write a buffer whose frame rank 0 has not consumed."""
    rng = np.random.default_rng(100 * g + seed)
    frames = 7
    streams = []
    for r in range(g):
        ops = []
        for f in range(frames):
            for op, v in multi.peer_schedule(r, g, f):
                ops.append((op, v, f))
            if r == 0:
                ops.append(("consume", f % 2, f))
        streams.append(ops)
    pc = [0] * g
    arrive, go = [0, 0], [0] * g
    written = [dict(), dict()]   # buffer -> {rank: frame it last wrote there}
    consumed = -1
    steps = 0
    while any(pc[r] < len(streams[r]) for r in range(g)):
        runnable = []
        for r in range(g):
            if pc[r] >= len(streams[r]):
                continue
            op, v, f = streams[r][pc[r]]
            if op == "wait_go" and go[r] < v:
                continue
            if op == "wait_arrive" and arrive[v[0]] < v[1]:
                continue
            runnable.append(r)
        assert runnable, "deadlock"
        r = int(rng.choice(runnable))
        op, v, f = streams[r][pc[r]]
        if op == "set_go":
            for k in range(1, g):
                go[k] = v
        elif op in ("render", "render_notify"):
            # the buffer's previous tenant (frame f - 2) must have been consumed before ANY rank overwrites its bands
            assert f < 2 or consumed >= f - 2, f"rank {r} overwrites frame {f - 2} before rank 0 consumed it"
            written[v][r] = f
            if op == "render_notify":
                arrive[v] += 1
        elif op == "consume":
            assert all(written[v].get(k) == f for k in range(g)), f"frame {f} incomplete or torn when consumed: {written[v]}"
            consumed = f
        pc[r] += 1
        steps += 1
    assert consumed == frames - 1 and steps == sum(len(s) for s in streams)


def test_hot_rectangle_tile_order_is_a_bijection():
    """render_inst.cu tile_of (queue position -> tile; the hot rectangle's tiles first) restated: every tile exactly once for
    any rectangle inside the grid, including none."""
    def tile_of(q, tiles_x, hot):
        x0, y0, x1, y1 = hot
        hw, hh = x1 - x0, y1 - y0
        n = hw * hh
        if q < n:
            return x0 + q % hw, y0 + q // hw
        q -= n
        above = y0 * tiles_x
        if n == 0 or q < above:
            return q % tiles_x, q // tiles_x
        q -= above
        side = tiles_x - hw
        if q < side * hh:
            c = q % side
            return (c if c < x0 else c + hw), y0 + q // side
        q -= side * hh
        return q % tiles_x, y1 + q // tiles_x
    rng = np.random.default_rng(3)
    for _ in range(400):
        tx, ty = int(rng.integers(1, 30)), int(rng.integers(1, 30))
        x0 = int(rng.integers(0, tx + 1)); x1 = int(rng.integers(x0, tx + 1))
        y0 = int(rng.integers(0, ty + 1)); y1 = int(rng.integers(y0, ty + 1))
        if x1 == x0 or y1 == y0:
            x0 = x1 = y0 = y1 = 0
        seen = {tile_of(q, tx, (x0, y0, x1, y1)) for q in range(tx * ty)}
        assert seen == {(a, b) for a in range(tx) for b in range(ty)}


@pytest.mark.parametrize("g", [1, 2, 4, 8])
@pytest.mark.parametrize("seed", range(6))
def test_shared_canvas_schedule_never_tears_a_frame(g, seed):
    """multi.canvas_schedule under adversarial interleavings: every rank runs its host-side operations in order, the ranks
    at arbitrary relative speed.  Rank 0's caller reads frame f from the return of render(f) until it calls render(f + 1)
    (a `read` step here, right before that call's first operation).  No schedule may deadlock, hand rank 0 an incomplete
    frame, or let any rank write into the one canvas while frame f may still be read."""
    rng = np.random.default_rng(1000 * g + seed)
    frames = 6
    procs = []
    for r in range(g):
        ops = []
        for f in range(1, frames + 1):
            ops += [(op, f) for op in multi.canvas_schedule(r, g, f)]
            if r == 0:
                ops.append((("read",), f))
        procs.append(ops)
    pc = [0] * g
    consumed, done = 0, [0] * g
    wrote = [0] * g        # frame whose bands rank r last put into the canvas
    reading = 0            # frame rank 0's caller may be reading (0: none)
    while any(pc[r] < len(procs[r]) for r in range(g)):
        runnable = []
        for r in range(g):
            if pc[r] >= len(procs[r]):
                continue
            op, f = procs[r][pc[r]]
            if op[0] == "wait_consumed" and consumed < op[1]:
                continue
            if op[0] == "wait_done" and any(done[k] < op[1] for k in range(1, g)):
                continue
            runnable.append(r)
        assert runnable, "deadlock"
        r = int(rng.choice(runnable))
        op, f = procs[r][pc[r]]
        if op[0] == "set_consumed":
            consumed, reading = op[1], 0
        elif op[0] == "render":
            assert reading == 0, f"rank {r} writes frame {f} while frame {reading} is being read"
            assert wrote[r] == f - 1
            wrote[r] = f
        elif op[0] == "set_done":
            done[r] = op[1]
        elif op[0] == "read":
            assert all(w == f for w in wrote), f"frame {f} incomplete when handed out: {wrote}"
            reading = f
        pc[r] += 1
    assert all(w == frames for w in wrote)


def _canvas_worker(rank, world_size, port, vsize, width, out_path):
    import time
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
    world, cam = rtc.build_scene("table", width, vsize)
    r = multi.SharedCanvasRenderer(world, cam, rank, world_size, 0, band_rows=8, want_f64=True, want_rgba8=True,
                                   page_lock=False, timeout_s=30.0)
    other = multi.SharedCanvasRenderer(world, cam, rank, world_size, 0, name=r.name, page_lock=False) if rank else None
    f64, rgba = r.views()
    frame_no = [0]

    def fake_render(scene, stats):  # stand-in for rtc_render + RTC_ROWS_FRAME: this rank's bands, stamped with the frame
        for b in r.plan.bands_of(rank):
            rows = slice(b * r.plan.band_rows, (b + 1) * r.plan.band_rows)
            f64[rows] = frame_no[0] + rank / 16.0
            rgba[rows] = (frame_no[0], rank, b % 256, 255)

    ok = True
    for f in range(1, 6):
        frame_no[0] = f
        out = r.render(_render=fake_render)
        if rank == 0:
            for _ in range(3):  # the caller reads for a while: nobody may write frame f + 1 under it
                owner = (np.arange(vsize) // r.plan.band_rows) % world_size
                ok &= bool((out[1][:, :, 0] == f).all() and (out[1][:, 0, 1] == owner).all())
                ok &= bool(np.array_equal(out[0][:, 0, 0], f + owner / 16.0))
                time.sleep(0.02)
    if rank == 0:
        np.save(out_path, np.array([ok]))
    if other is not None:
        other._unmap(unlink=False)
    r.close()
    dist.destroy_process_group()


def test_gloo_world_size_2_shared_canvas(tmp_path):
    """Two processes, one canvas in POSIX shared memory (not page-locked: no GPU here), the library's host counters as the
    only per-frame synchronisation; the name travels by broadcast_object_list at set-up."""
    port = 31500 + os.getpid() % 2000
    out = str(tmp_path / "ok.npy")
    mp.spawn(_canvas_worker, args=(2, port, 48, 16, out), nprocs=2, join=True)
    assert bool(np.load(out)[0])


def test_host_counter_wait_times_out_and_share_rejects_bad_names():
    rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
    import ctypes as C
    api = rtc.api()
    p = C.c_void_p()
    assert api.host_share_create(-1, b"no-slash", 4096, C.byref(p)) == rtc.RTC_ERR_INVALID
    name = f"/rtc_test_{os.getpid()}".encode()
    api.check(api.host_share_create(-1, name, 4096, C.byref(p)))
    try:
        q = C.c_void_p()
        assert api.host_share_create(-1, name, 4096, C.byref(q)) == rtc.RTC_ERR_INVALID      # exists already
        assert api.host_share_open(-1, name, 1 << 20, C.byref(q)) == rtc.RTC_ERR_INVALID     # smaller than asked for
        api.check(api.host_share_open(-1, name, 4096, C.byref(q)))
        api.host_counter_store(C.c_void_p(p.value + 64), 7)
        assert api.host_counter_load(C.c_void_p(q.value + 64)) == 7                            # the same pages
        assert api.host_counter_wait(C.c_void_p(q.value + 64), 7, 0.1) == rtc.RTC_OK
        assert api.host_counter_wait(C.c_void_p(q.value + 64), 8, 0.05) == rtc.RTC_ERR_TIMEOUT
        api.check(api.host_share_close(q, 4096, None))
    finally:
        api.check(api.host_share_close(p, 4096, name))
    assert not os.path.exists("/dev/shm/" + name.decode().lstrip("/"))
