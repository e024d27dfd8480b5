"""Host logic of the batched clip driver (hunyuanvideo_efficiency_b200/infer.py): file order, per-rank deal, on-disk
format, and the world_size-2 run (gloo) covering every clip exactly once.  The model call is injected, so no GPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hunyuanvideo_efficiency_b200 import infer as I


def _make_dir(tmp, n=7):
    d = os.path.join(tmp, "in")
    os.makedirs(d)
    for i in range(n):
        torch.save(torch.full((3, 5, 8, 8), float(i)), os.path.join(d, f"clip_{i:03d}.pt"))
    open(os.path.join(d, "notes.txt"), "w").write("ignored")
    return d


def test_listing_and_deal_follow_the_reference_dataset(tmp_path):
    d = _make_dir(str(tmp_path))
    files = I.list_clips(d)
    assert files == [f"clip_{i:03d}.pt" for i in range(7)]            # sorted *.pt only (dataset_loader.py:11-12)
    deals = [I.clips_of_rank(files, r, 3) for r in range(3)]
    assert sorted(sum(deals, [])) == files and all(len(x) in (2, 3) for x in deals)
    assert I.clips_of_rank(files, 0, 2, max_files=3) == ["clip_000.pt", "clip_002.pt"]   # --max-files truncates first


def test_single_process_roundtrip_format(tmp_path):
    d = _make_dir(str(tmp_path), n=3)
    out = os.path.join(str(tmp_path), "out")
    seen = []

    def fake_model(x):                                                 # (1, C, T, H, W) in, same out
        assert x.shape == (1, 3, 5, 8, 8) and x.dtype == torch.float16
        seen.append(float(x.flatten()[0]))
        return (x * 2).to(torch.float16)

    done = I.run_clips(fake_model, d, out, in_dtype=torch.float16)
    assert [n for n, _ in done] == ["clip_000.pt", "clip_001.pt", "clip_002.pt"] and seen == [0.0, 1.0, 2.0]
    y = torch.load(os.path.join(out, "clip_002.pt"))
    assert y.dtype == torch.float32 and y.shape == (1, 3, 5, 8, 8) and torch.all(y == 4.0)   # infer.py:63-65 format


def _worker(rank, world, port, d, out, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    done = I.run_clips(lambda x: x + 1, d, out, rank, world, in_dtype=torch.float32)
    names = [None] * world
    dist.all_gather_object(names, [n for n, _ in done])
    q.put((rank, names))
    dist.destroy_process_group()


def test_two_ranks_cover_every_clip_once(tmp_path):
    d = _make_dir(str(tmp_path), n=5)
    out = os.path.join(str(tmp_path), "out")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, d, out, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res[0] == res[1] and sorted(res[0][0] + res[0][1]) == I.list_clips(d) and not set(res[0][0]) & set(res[0][1])
    for i, name in enumerate(I.list_clips(d)):
        assert torch.all(torch.load(os.path.join(out, name)) == float(i) + 1.0)
