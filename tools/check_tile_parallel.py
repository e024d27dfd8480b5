"""torchrun target: the tile-parallel round trip over WORLD_SIZE GPUs must equal the single-GPU one bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D, tile_parallel as TP
from hunyuanvideo_efficiency_b200 import synthetic as W

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(W.SMALL_CONFIG, block_out_channels=[64, 128, 256, 256])
vae = AutoencoderKLCausal3D.from_config(cfg)
vae.load_state_dict(W.make_state_dict(cfg))
vae = vae.to(torch.bfloat16).to(dev).eval().requires_grad_(False)
vae.enable_tiling()
x = W.make_video((1, 3, 29, 72, 88)).to(dev, torch.bfloat16)
with torch.no_grad():
    runner = TP.TileParallelVAE(vae, rank, world)
    ref = vae.decode(vae.encode(x).latent_dist.mode()).sample if rank == 0 else None
    for it in range(3):      # the peer-memory arenas alternate between calls
        out = runner.roundtrip(x)
        if rank == 0:
            print(f"call {it}: exchange={runner.last_exchange} tile-parallel == single GPU:", torch.equal(out, ref), tuple(out.shape), flush=True)
            assert torch.equal(out, ref)
        else:
            assert out is None
    os.environ["HYVAE_TILE_PUSH"] = "0"     # the NCCL gather path must give the same bits
    g = TP.TileParallelVAE(vae, rank, world)
    out = g.roundtrip(x)
    if rank == 0:
        print(f"exchange={g.last_exchange} tile-parallel == single GPU:", torch.equal(out, ref), flush=True)
        assert torch.equal(out, ref)
    runner.close()
dist.barrier()
dist.destroy_process_group()
