#!/usr/bin/env python3
"""VAE round-trip driver over a directory of `.pt` clips — the fork's `infer.py` as ONE multi-GPU job.

The reference (/root/reference/infer.py:28-72,96-123) loads the VAE in fp16 with the t-ops JSON, iterates a
`VideoTensorDataset` (sorted `*.pt`, each a `(C, T, H, W)` fp32 tensor in [-1, 1], dataset_loader.py:9-22) with batch
size 1, runs `model(x, return_dict=False, return_posterior=True, sample_posterior=False)[0]` and `torch.save`s the
fp32 reconstruction `(1, C, T, H, W)` under the same file name; its launchers get data parallelism by starting one
such process per GPU on disjoint config lists (run_experiments_pool.sh:76-120).

Here the clips of one directory are dealt round-robin to the ranks of a `torchrun` job (no collective: clips are
independent, BASELINE config 3), and every rank overlaps the three stages of a clip on its own threads / streams:

    reader thread : torch.load -> pinned host buffer                      (CPU, disk)
    main thread   : H2D on a copy stream -> model round trip on the compute stream -> D2H into a pinned buffer
    writer thread : torch.save of the fp32 reconstruction                 (CPU, disk)

On-disk formats, file naming, argument names and the model call are the reference's.  Usage:

    python -m hunyuanvideo_efficiency_b200.infer --tensor-dir IN --output-dir OUT [--vae-path P] [--config-json J]
    torchrun --nproc-per-node 8 -m hunyuanvideo_efficiency_b200.infer --tensor-dir IN --output-dir OUT ...
"""
from __future__ import annotations

import argparse
import os
import queue
import threading
from typing import Callable, List, Optional, Sequence, Tuple

import torch

_STOP = object()


def list_clips(tensor_dir: str) -> List[str]:
    """Sorted `*.pt` file names — the order VideoTensorDataset uses (dataset_loader.py:11-12)."""
    return sorted(f for f in os.listdir(tensor_dir) if f.endswith(".pt"))


def clips_of_rank(files: Sequence[str], rank: int, world: int, max_files: Optional[int] = None) -> List[str]:
    """Round-robin deal of the (optionally truncated, infer.py:43-44) file list; deterministic on every rank."""
    files = list(files if max_files is None else files[:max_files])
    return [f for i, f in enumerate(files) if i % world == rank]


def _reader(tensor_dir: str, names: Sequence[str], out_q: "queue.Queue", pin: bool):
    try:
        for name in names:
            x = torch.load(os.path.join(tensor_dir, name), weights_only=False)  # (C, T, H, W)
            if x.ndim == 4:
                x = x.unsqueeze(0)                                               # DataLoader(batch_size=1) adds this
            if pin:
                x = x.contiguous().pin_memory()
            out_q.put((name, x))
    except BaseException as e:  # surface I/O errors on the main thread
        out_q.put(e)
    out_q.put(_STOP)


def _writer(output_dir: str, in_q: "queue.Queue", errors: list):
    while True:
        item = in_q.get()
        if item is _STOP:
            return
        name, ready, host = item
        try:
            if ready is not None:
                ready.synchronize()
            torch.save(host.float(), os.path.join(output_dir, name))
        except BaseException as e:
            errors.append(e)


def run_clips(process: Callable[[torch.Tensor], torch.Tensor], tensor_dir: str, output_dir: str, rank: int = 0, world: int = 1,
              max_files: Optional[int] = None, device: Optional[torch.device] = None, in_dtype=torch.float16,
              prefetch: int = 2, log: Callable[[str], None] = lambda s: None) -> List[Tuple[str, Tuple[int, ...]]]:
    """Run `process` (device tensor (1,C,T,H,W) -> device tensor) over this rank's share of `tensor_dir` and save the
    results under `output_dir` with the input names.  With a CUDA `device` the H2D copy of clip i+1 and the D2H copy /
    torch.save of clip i-1 overlap the compute of clip i.  Returns [(name, output shape)] in processing order."""
    os.makedirs(output_dir, exist_ok=True)
    names = clips_of_rank(list_clips(tensor_dir), rank, world, max_files)
    cuda = device is not None and torch.device(device).type == "cuda"
    rq: "queue.Queue" = queue.Queue(maxsize=max(prefetch, 1))
    wq: "queue.Queue" = queue.Queue(maxsize=max(prefetch, 1))
    errors: list = []
    rt = threading.Thread(target=_reader, args=(tensor_dir, names, rq, cuda), daemon=True)
    wt = threading.Thread(target=_writer, args=(output_dir, wq, errors), daemon=True)
    rt.start(); wt.start()
    done = []
    copy_stream = torch.cuda.Stream(device) if cuda else None
    try:
        while True:
            item = rq.get()
            if item is _STOP:
                break
            if isinstance(item, BaseException):
                raise item
            name, host_x = item
            if cuda:
                with torch.cuda.stream(copy_stream):
                    x = host_x.to(device, non_blocking=True).to(in_dtype)
                torch.cuda.current_stream(device).wait_stream(copy_stream)
                x.record_stream(torch.cuda.current_stream(device))
            else:
                x = host_x.to(in_dtype) if device is None else host_x.to(device, in_dtype)
            log(f"Processing {name[:-3]}, video shape: {tuple(x.shape)}")
            with torch.no_grad():
                y = process(x)
            if cuda:
                host_y = torch.empty(y.shape, dtype=y.dtype).pin_memory()
                host_y.copy_(y, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(device))
                wq.put((name, ev, host_y))
            else:
                wq.put((name, None, y.cpu()))
            done.append((name, tuple(y.shape)))
    finally:
        wq.put(_STOP)
        wt.join()
    if errors:
        raise errors[0]
    return done


def roundtrip(vae) -> Callable[[torch.Tensor], torch.Tensor]:
    """The reference's model call (infer.py:55-60): encode -> posterior mode -> decode."""
    return lambda x: vae(x, return_dict=False, return_posterior=True, sample_posterior=False)[0]


def encode_moments(vae) -> Callable[[torch.Tensor], torch.Tensor]:
    """Dataset encode (BASELINE config 3): the posterior parameters (mean, logvar) of each clip."""
    return lambda x: vae.encode(x).latent_dist.parameters


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="VAE inference over a directory of .pt video tensors (B200-native VAE).")
    p.add_argument("--tensor-dir", type=str, required=True, help="Directory containing input .pt video tensors.")
    p.add_argument("--output-dir", type=str, required=True, help="Directory to save the reconstructed videos.")
    p.add_argument("--vae-path", type=str, default="ckpts/hunyuan-video-t2v-720p/vae",
                   help="Path to VAE checkpoint directory (contains config.json and pytorch_model.pt).")
    p.add_argument("--config-json", type=str, default=None, help="Path to the T-ops config JSON file.")
    p.add_argument("--max-files", type=int, default=None, help="Max number of input files to process.")
    p.add_argument("--vae-precision", type=str, default="fp16", choices=["fp16", "bf16", "fp32"])
    p.add_argument("--tiling", action="store_true", help="vae.enable_tiling() (needed above 256x256x64).")
    p.add_argument("--encode-only", action="store_true", help="Save the posterior moments instead of the reconstruction.")
    return p.parse_args(argv)


def main(argv=None):
    from .vae import PRECISION_TO_TYPE, load_vae
    args = parse_args(argv)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if not torch.cuda.is_available():
        raise SystemExit("the B200-native VAE has no CPU path: run on a CUDA device")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    vae, _, _, _ = load_vae(vae_type="884-16c-hy", vae_precision=args.vae_precision, vae_path=args.vae_path, device=device,
                            t_ops_config_path=args.config_json, test=args.config_json is not None)
    if args.tiling:
        vae.enable_tiling()
    fn = encode_moments(vae) if args.encode_only else roundtrip(vae)
    done = run_clips(fn, args.tensor_dir, args.output_dir, rank, world, args.max_files, device,
                     PRECISION_TO_TYPE[args.vae_precision], log=lambda s: print(f"[rank {rank}] {s}", flush=True))
    torch.cuda.synchronize()
    print(f"[rank {rank}] {len(done)} clip(s) written to {args.output_dir}", flush=True)


if __name__ == "__main__":
    main()
