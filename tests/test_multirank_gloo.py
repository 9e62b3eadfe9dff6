"""CPU, world_size 2 (gloo): the multi-GPU path's host logic — interleaved scanline-tile ownership, jitter-stream
determinism per rank (every rank replays the stream to ITS rows) and the framebuffer gather/assembly — with the CPU
oracle standing in for the per-rank render (tests may use the oracle; the product never does)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import conftest as T

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
    import conftest as T
    rt = T.rtds_b200
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, spp, tile = 96, 71, 2, 8
    sph, mat = T.synthetic_scene(400, 5)
    oracle = T.Oracle()
    rc, nodes, order, _ = oracle.build_bvh(sph)
    rows = rt.owned_rows(H, tile, rank, world)
    local = np.zeros((len(rows), W, 3), np.uint8)
    k = 0
    for t in range(rank, (H + tile - 1) // tile, world):          # one oracle call per owned tile
        y0, y1 = t * tile, min(H, (t + 1) * tile)
        rgb, _, _, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, y0, y1, want_hit=False)
        local[k:k + (y1 - y0)] = rgb
        k += y1 - y0
    assert k == len(rows) == rt.rows_for_rank(H, tile, rank, world)
    frame = rt.gather_frame(torch.from_numpy(local), H, W, tile, rank, world)
    if rank == 0:
        full, _, _, _ = oracle.render_rows(sph, mat, nodes, order, W, H, spp, want_hit=False)
        assert np.array_equal(frame.numpy(), full), "assembled frame differs from the single-rank frame"
        print("GATHER_OK", int(frame.numpy().astype(np.int64).sum()))
    else:
        assert frame is None
    dist.barrier()
    dist.destroy_process_group()
''')


def test_two_rank_gather_matches_single_rank(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), T.ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GATHER_OK" in outs[0]


def test_ownership_is_a_partition_with_balanced_tiles():
    rt = T.rtds_b200
    for world in (2, 4, 8):
        rows = [rt.owned_rows(2160, 8, r, world) for r in range(world)]
        assert sorted(np.concatenate(rows).tolist()) == list(range(2160))
        sizes = [len(r) for r in rows]
        assert max(sizes) - min(sizes) <= 8


SLICE_WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
    import conftest as T
    rt = T.rtds_b200
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for n in (1, 2, 5, 1001, 4096):
        rng = np.random.default_rng(n)
        table = torch.from_numpy(rng.normal(size=(n, 4)).astype(np.float32))       # every rank holds the same host table
        per, lo, hi = rt.scene_slice(n, rank, world)
        part = torch.zeros((per, 4), dtype=torch.float32)
        part[: hi - lo] = table[lo:hi]                                               # "upload" of this rank's 1/N
        full = torch.zeros((per * world, 4), dtype=torch.float32)
        rt.exchange_scene(part, full)
        assert torch.equal(full[:n], table), (n, rank)
        assert not full[n:].any()
    if rank == 0:
        print("SLICE_OK")
''')


def test_scene_exchange_reassembles_the_table_on_every_rank(tmp_path):
    """bench.py at N >= 2, end to end: every rank uploads only its 1/N of the scene tables and the ranks all-gather the parts
    (rtds_b200.scene_slice / exchange_scene; NCCL over NVLink on the GPUs). Here with gloo, world 2: every rank ends up with the
    whole table, bit for bit, for table sizes that do and do not divide by the world size."""
    script = tmp_path / "slice_worker.py"
    script.write_text(SLICE_WORKER)
    port = 29400 + os.getpid() % 500
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), T.ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SLICE_OK" in outs[0]


def test_scene_slices_partition_the_table():
    rt = T.rtds_b200
    for n in (1, 7, 8, 9, 1078411):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                per, lo, hi = rt.scene_slice(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                cover.extend(range(lo, hi)) if n < 100 else cover.append((lo, hi))
            if n < 100:
                assert cover == list(range(n))
            else:
                assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
