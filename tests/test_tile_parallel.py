"""Tile-parallel partition + exchange logic on CPU: world_size 2, gloo, oracle tile functions injected."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hunyuanvideo_efficiency_b200.vae import tile_parallel as TP
from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D
from oracle import vae_oracle as O
from oracle import weights as W


def test_tile_grid_matches_reference_counts():
    # BASELINE config 4 / 2 / 3 tile counts
    g = TP.tile_grid(129, 720, 1280, temporal=True, spatial=True, min_t=64, min_s=256, overlap=0.25)
    assert len(g) == 84 and sorted({s.t1 - s.t0 for s in g}) == [33, 65]
    g = TP.tile_grid(33, 90, 160, temporal=True, spatial=True, min_t=16, min_s=32, overlap=0.25)
    assert len(g) == 84 and sum(1 for s in g if (s.t1 - s.t0, s.h1 - s.h0, s.w1 - s.w0) == (17, 32, 32)) == 36
    g = TP.tile_grid(65, 544, 960, temporal=True, spatial=True, min_t=64, min_s=256, overlap=0.25)
    assert len(g) == 30


def test_lpt_is_balanced_for_the_720p_grid():
    g = TP.tile_grid(33, 90, 160, temporal=True, spatial=True, min_t=16, min_s=32, overlap=0.25)
    costs = [s.cost for s in g]
    for world in (2, 4, 8):
        owner = TP.lpt_assign(costs, world)
        load = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(world)]
        assert max(load) / (sum(costs) / world) < 1.04     # partition alone permits > 7.7x at 8 ranks


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = dict(W.SMALL_CONFIG)
    sd = W.make_state_dict(cfg)
    vae = AutoencoderKLCausal3D.from_config(cfg)   # host-side object only: tiling attributes + config
    vae.enable_tiling()
    tl = O.Tiling.from_cfg(cfg, True, True)
    x = W.make_video((1, 3, 29, 56, 40))
    runner = TP.TileParallelVAE(
        vae, rank, world,
        enc_tile=lambda t: O._enc_tile(sd, cfg, t, None), dec_tile=lambda t: O._dec_tile(sd, cfg, t, None),
        assemble_spatial=O.spatial_assemble,
        assemble_temporal=lambda row, e, l: O.temporal_assemble([t[:, :, off:] for t, off in row], e, l))
    with torch.no_grad():
        mom = runner.encode_moments(x)
        dec = runner.decode(mom[:, :16].contiguous())
        ref_mom = O.encode_moments(sd, cfg, x, tl)
        ok = torch.allclose(mom, ref_mom, atol=1e-5)
        if rank == 0:
            ref_dec = O.decode(sd, cfg, ref_mom[:, :16].contiguous(), tl)
            ok = ok and torch.allclose(dec, ref_dec, atol=1e-4)
        else:
            ok = ok and dec is None
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_partition_equals_single_process_oracle():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
