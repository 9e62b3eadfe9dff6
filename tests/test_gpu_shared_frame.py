"""GPU: the multi-GPU frame assembled by direct stores into rank 0's buffer (rtds_shared_frame_* / rtds_render_shared)
must equal the single-rank frame byte for byte. In-process variant: several contexts (ranks) on one device attach to the
owner's frame; IPC variant: two processes under torch.distributed (needs 2 GPUs, skipped otherwise)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("graph", [0, 1])
@pytest.mark.parametrize("world,W,H,spp,shadows", [(2, 320, 200, 4, 0), (3, 322, 203, 4, 0), (4, 640, 360, 1, 0), (2, 320, 200, 2, 1)])
def test_shared_frame_in_process_equals_single_rank(gpu_ctx, world, W, H, spp, shadows, graph):
    with T.option(gpu_ctx, "frame_graph", graph):
        _shared_frame_in_process(gpu_ctx, world, W, H, spp, shadows, graph)


def _shared_frame_in_process(gpu_ctx, world, W, H, spp, shadows, graph):
    sph, mat = T.synthetic_scene(3000, 31)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    full, _, _, _ = gpu_ctx.render(rt.LBVH, W, H, spp, shadows=shadows)
    ranks = [gpu_ctx] + [rt.Rtds(0) for _ in range(world - 1)]
    try:
        for c in ranks[1:]:
            c.set_option("frame_graph", graph)
            c.set_spheres(sph, mat)
            c.build(rt.LBVH, mode=rt.MODE_TRUE)
        handle = gpu_ctx.shared_frame_create(W, H, world)
        assert len(handle) == 64
        for r, c in enumerate(ranks[1:], 1):
            c.shared_frame_attach(gpu_ctx, r)
        for seq in (1, 2):                                    # two frames: flags are per frame
            rays = 0
            for r in range(world - 1, -1, -1):                # the owner last: it waits for every rank's flag
                p = ranks[r].render_params(W, H, spp, rank=r, world=world, shadows=shadows)
                st = ranks[r].render_shared(rt.LBVH, p, seq)
                rays += st["primary_rays"]
            assert rays == W * H * spp
            assert np.array_equal(gpu_ctx.shared_frame_read(W, H), full)
        # rtds_frame_shared: upload + build + render_shared in one call per rank gives the same frame
        for r in range(world - 1, -1, -1):
            p = ranks[r].render_params(W, H, spp, rank=r, world=world, shadows=shadows)
            bst, rst = ranks[r].frame_shared(sph, mat, rt.LBVH, p, 7, mode=rt.MODE_TRUE)
            assert bst["n_prims"] == sph.shape[0] and rst["rows"] == rt.rows_for_rank(H, 8, r, world)
        assert np.array_equal(gpu_ctx.shared_frame_read(W, H), full)
        # mismatching parameters are refused
        with pytest.raises(rt.RtdsError):
            ranks[1].render_shared(rt.LBVH, ranks[1].render_params(W, H, spp, rank=0, world=world), 3)
        with pytest.raises(rt.RtdsError):
            gpu_ctx.render_shared(rt.LBVH, gpu_ctx.render_params(W, H, spp, rank=0, world=world), 0)
    finally:
        for c in ranks[1:]:
            c.shared_frame_close()
            c.close()
        gpu_ctx.shared_frame_close()


_IPC_SCRIPT = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.environ["RTDS_ROOT"], "tests"))
import conftest as T
rt = T.rtds_b200
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
W, H, spp = 640, 360, 4
sph, mat = T.synthetic_scene(3000, 31)
ctx = rt.Rtds(rank)
ctx.set_spheres(sph, mat)
ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
box = [ctx.shared_frame_create(W, H, world) if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
if rank != 0:
    ctx.shared_frame_open(box[0], W, H, world, rank)
for seq in (1, 2, 3):
    dist.barrier()
    ctx.render_shared(rt.LBVH, ctx.render_params(W, H, spp, rank=rank, world=world), seq)
    if rank == 0:
        got = ctx.shared_frame_read(W, H)
        full, _, _, _ = ctx.render(rt.LBVH, W, H, spp)
        assert np.array_equal(got, full), "IPC shared frame differs from the single-rank frame"
dist.barrier()
ctx.shared_frame_close()
if rank == 0:
    print("IPC_OK")
dist.destroy_process_group()
'''


def test_shared_frame_ipc_two_processes(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "ipc.py"
    script.write_text(_IPC_SCRIPT)
    env = dict(os.environ, RTDS_ROOT=T.ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "IPC_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
