#!/usr/bin/env python3
"""Temporal stride / pool sweep (BASELINE config 5) as ONE multi-GPU job.

The fork's launchers (run_experiments_pool.sh:24-120, run_experiments_stride.sh) (1) enumerate t-ops configs with
dynamic_enumeration.py / dynamic_enumeration_stride.py into exp_N.json files, (2) start one `infer.py` process per GPU
on a batch of four configs at a time (encode -> decode of every clip, reconstruction saved as .pt and .mp4) and
(3) score every output directory with evaluation/compute_metrics.py (PSNR / SSIM / LPIPS of mp4 frames).

Here the experiment configs are dealt round-robin over the ranks of a `torchrun` job (configs are independent: no
collective), each rank loads the VAE once, re-arms the t-ops hooks per config, runs the clips and scores the
reconstruction against the input on the GPU (metrics.py), writing `<metrics-dir>/<exp name>/metrics.txt` in the
reference's result format (compute_metrics.py:73-86).  The enumerators are restated so the sweep needs no
intermediate files; `--config-dir` takes exp_*.json files produced by the fork's own scripts instead.

    torchrun --nproc-per-node 8 -m hunyuanvideo_efficiency_b200.sweep --tensor-dir IN --metrics-dir OUT \
        --base-config t_ops_config.json --mode pool [--vae-path P] [--max-files 100] [--save-dir RECON]
"""
from __future__ import annotations

import argparse
import copy
import glob
import json
import os
from typing import Dict, List, Optional, Tuple

import torch

Slot = Tuple[int, int, str]


def default_t_ops_config(layers_per_block: int = 2, n_blocks: int = 4, pool_kernel: int = 3, pool_stride: int = 2,
                         strides=((1, 2, 2), (2, 2, 2), (2, 2, 2), (1, 1, 1))) -> dict:
    """An all-off base config with the schema of the fork's t_ops_config.json (one record per down / up block, resnet
    flag lists of length layers_per_block / layers_per_block + 1, the stock downsample strides, mid-block records)."""
    off = lambda n: [False] * n
    down = [{"block_type": "DownEncoderBlockCausal3D", "block_index": i, "pool_t_kernel": pool_kernel, "pool_t_stride": pool_stride,
             "enable_t_pool_before_block": off(layers_per_block), "enable_t_pool_after_block": off(layers_per_block),
             "downsample_stride": list(strides[i])} for i in range(n_blocks)]
    up = [{"block_type": "UpDecoderBlockCausal3D", "block_index": i, "enable_t_interp_before_block": off(layers_per_block + 1),
           "enable_t_interp_after_block": off(layers_per_block + 1), "interp_t_scale_factor": 2, "interp_mode": "nearest"}
          for i in range(n_blocks)]
    mid = lambda: {"mid_block_type": "UNetMidBlockCausal3D", "pool_t_kernel": pool_kernel, "pool_t_stride": pool_stride,
                   "enable_t_pool_before_block": off(2), "enable_t_pool_after_block": off(2)}
    return {"encoder": {"down_blocks": down, "mid_block": mid()}, "decoder": {"up_blocks": up, "mid_block": mid()}}


def _slots(blocks: List[dict], key_before: str, key_after: str) -> List[Slot]:
    """(block, resnet, 'before'|'after') insertion points, in the reference's order (dynamic_enumeration.py:21-60)."""
    out: List[Slot] = []
    for i, block in enumerate(blocks):
        n = min(len(block.get(key_before, [])), len(block.get(key_after, [])))
        for j in range(n):
            out += [(i, j, "before"), (i, j, "after")]
    return out


def encoder_slots(cfg: dict) -> List[Slot]:
    return _slots(cfg.get("encoder", {}).get("down_blocks", []), "enable_t_pool_before_block", "enable_t_pool_after_block")


def decoder_slots(cfg: dict) -> List[Slot]:
    return _slots(cfg.get("decoder", {}).get("up_blocks", []), "enable_t_interp_before_block", "enable_t_interp_after_block")


def _clear(blocks: List[dict], keys: Tuple[str, str]):
    for block in blocks:
        for k in keys:
            if k in block:
                block[k] = [False] * len(block[k])


def enumerate_pool_configs(base: dict, max_combos: int = 384) -> List[Tuple[str, dict]]:
    """dynamic_enumeration.py:79-118: one temporal pool in the encoder x one temporal interpolation in the decoder,
    everything else off; exp_1 .. exp_N in (encoder slot, decoder slot) order, capped at 384."""
    out = []
    for e in encoder_slots(base):
        for d in decoder_slots(base):
            if len(out) >= max_combos:
                return out
            cfg = copy.deepcopy(base)
            _clear(cfg.get("encoder", {}).get("down_blocks", []), ("enable_t_pool_before_block", "enable_t_pool_after_block"))
            _clear(cfg.get("decoder", {}).get("up_blocks", []), ("enable_t_interp_before_block", "enable_t_interp_after_block"))
            cfg["encoder"]["down_blocks"][e[0]]["enable_t_pool_" + e[2] + "_block"][e[1]] = True
            cfg["decoder"]["up_blocks"][d[0]]["enable_t_interp_" + d[2] + "_block"][d[1]] = True
            out.append((f"exp_{len(out) + 1}", cfg))
    return out


def enumerate_stride_configs(base: dict) -> List[Tuple[str, dict]]:
    """dynamic_enumeration_stride.py:63-131: the temporal stride of encoder down block 0, 1 or 2 doubled (block 0:
    1 -> 2; blocks 1, 2: 2 -> 4) x one temporal interpolation in the decoder; encoder pools off.  Like the reference,
    the decoder flags of the base config are kept as they are and one more is switched on."""
    out = []
    for eb in (0, 1, 2):
        for d in decoder_slots(base):
            cfg = copy.deepcopy(base)
            st = cfg["encoder"]["down_blocks"][eb]["downsample_stride"]
            cfg["encoder"]["down_blocks"][eb]["downsample_stride"] = [2 if eb == 0 else st[0] * 2, st[1], st[2]]
            _clear(cfg.get("encoder", {}).get("down_blocks", []), ("enable_t_pool_before_block", "enable_t_pool_after_block"))
            cfg["decoder"]["up_blocks"][d[0]]["enable_t_interp_" + d[2] + "_block"][d[1]] = True
            out.append((f"exp_{len(out) + 1}", cfg))
    return out


def enumerate_stride_pair_configs(base: dict) -> List[Tuple[str, dict]]:
    """dynamic_enumeration_stride_2.py:82-101: TWO of the encoder down blocks 0, 1, 2 get their temporal stride doubled
    (block 0: 1 -> 2; blocks 1, 2: x2) x TWO temporal interpolations in the decoder; every pool / interp flag of the base
    config is cleared first (set_all_false, :30-44).  3 block pairs x C(24, 2) slot pairs = 828 configs for the stock JSON,
    numbered exp_1.. in (block pair, slot pair) order."""
    out = []
    dslots = decoder_slots(base)
    for a, eb1 in enumerate((0, 1, 2)):
        for eb2 in (0, 1, 2)[a + 1:]:
            for j, d1 in enumerate(dslots):
                for d2 in dslots[j + 1:]:
                    cfg = copy.deepcopy(base)
                    for eb in (eb1, eb2):
                        st = cfg["encoder"]["down_blocks"][eb]["downsample_stride"]
                        cfg["encoder"]["down_blocks"][eb]["downsample_stride"] = [2 if eb == 0 else st[0] * 2, st[1], st[2]]
                    _clear(cfg.get("encoder", {}).get("down_blocks", []), ("enable_t_pool_before_block", "enable_t_pool_after_block"))
                    _clear(cfg.get("decoder", {}).get("up_blocks", []), ("enable_t_interp_before_block", "enable_t_interp_after_block"))
                    for d in (d1, d2):
                        cfg["decoder"]["up_blocks"][d[0]]["enable_t_interp_" + d[2] + "_block"][d[1]] = True
                    out.append((f"exp_{len(out) + 1}", cfg))
    return out


def configs_of_rank(configs: List[Tuple[str, dict]], rank: int, world: int) -> List[Tuple[str, dict]]:
    return [c for i, c in enumerate(configs) if i % world == rank]


def write_metrics(results: Dict[str, float], root1: str, root2: str, results_dir: str, name: str = "metrics.txt") -> str:
    """Result file in the layout of compute_metrics.py:73-86 (Root1 / Root2 / one `metric: value` line each)."""
    os.makedirs(results_dir, exist_ok=True)
    path = os.path.join(results_dir, name)
    with open(path, "w") as f:
        f.write("\n")
        f.write(f"Root1: {root1}\n")
        f.write(f"Root2: {root2}\n")
        for k, v in results.items():
            f.write(f"{k}: {v}\n")
        f.write("\n")
    return path


def run_config(vae, t_ops: Optional[dict], tensor_dir: str, names: List[str], device, in_dtype, save_dir: Optional[str] = None) -> Dict[str, float]:
    """Arm the hooks of one experiment, round-trip the clips, score them.  The original geometry (strides, hooks) is
    restored afterwards so the same module serves the next config."""
    from . import metrics as M
    from .vae import _apply_t_ops_config_to_vae
    saved = snapshot_t_ops(vae)
    if t_ops is not None:
        _apply_t_ops_config_to_vae(vae, t_ops)
    pairs_psnr: List[float] = []
    pairs_ssim: List[float] = []
    try:
        for name in names:
            x = torch.load(os.path.join(tensor_dir, name), weights_only=False)
            if x.ndim == 4:
                x = x.unsqueeze(0)
            xd = x.to(device)
            with torch.no_grad():
                rec = vae(xd.to(in_dtype), return_dict=False, return_posterior=True, sample_posterior=False)[0]
            rec32 = rec.float()  # the reference scores the fp32 tensor it saved (infer.py:63)
            p, s = M.frame_metrics(M.video_to_frames_u8(xd.float(), True), M.video_to_frames_u8(rec32, True))
            pairs_psnr += p
            pairs_ssim += s
            if save_dir is not None:
                os.makedirs(save_dir, exist_ok=True)
                torch.save(rec32.cpu(), os.path.join(save_dir, name))
    finally:
        restore_t_ops(vae, saved)
    out: Dict[str, float] = {}
    if pairs_psnr:
        out["PSNR"] = sum(pairs_psnr) / len(pairs_psnr)
        out["SSIM"] = sum(pairs_ssim) / len(pairs_ssim)
    return out


def snapshot_t_ops(vae):
    """Everything apply_t_ops_config mutates: per-resnet hook records and the downsamplers' conv strides."""
    snap = {"strides": [], "hooks": []}
    for blk in vae.encoder.down_blocks:
        for ds in (blk.downsamplers or []):
            snap["strides"].append((ds.conv.conv, tuple(ds.conv.conv.stride)))
    for blk in list(vae.encoder.down_blocks) + list(vae.decoder.up_blocks) + [vae.encoder.mid_block, vae.decoder.mid_block]:
        for attr in ("resnet_pool_configs", "resnet_pad_configs", "resnet_interp_configs"):
            if hasattr(blk, attr):
                snap["hooks"].append((blk, attr, copy.deepcopy(getattr(blk, attr))))
    return snap


def restore_t_ops(vae, snap):
    for conv, stride in snap["strides"]:
        conv.stride = stride
    for blk, attr, val in snap["hooks"]:
        setattr(blk, attr, val)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Temporal stride / pool sweep with PSNR / SSIM on the GPU.")
    p.add_argument("--tensor-dir", required=True, help="Directory of input .pt clips ((C, T, H, W) fp32 in [-1, 1]).")
    p.add_argument("--metrics-dir", required=True, help="One sub-directory with metrics.txt per experiment is written here.")
    p.add_argument("--base-config", default="t_ops_config.json", help="Base t-ops JSON the enumerators start from.")
    p.add_argument("--mode", default="pool", choices=["pool", "stride", "stride2", "dir"],
                   help="Enumerator (dynamic_enumeration.py / _stride.py / _stride_2.py), or 'dir' to read --config-dir/exp_*.json.")
    p.add_argument("--config-dir", default=None)
    p.add_argument("--vae-path", default="ckpts/hunyuan-video-t2v-720p/vae")
    p.add_argument("--vae-precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    p.add_argument("--max-files", type=int, default=100)
    p.add_argument("--max-configs", type=int, default=None)
    p.add_argument("--save-dir", default=None, help="Also save the reconstructions (.pt, infer.py's format) under <save-dir>/<exp>.")
    return p.parse_args(argv)


def main(argv=None):
    from .infer import list_clips
    from .vae import PRECISION_TO_TYPE, load_vae
    args = parse_args(argv)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if not torch.cuda.is_available():
        raise SystemExit("the B200-native VAE has no CPU path: run on a CUDA device")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if args.mode == "dir":
        files = sorted(glob.glob(os.path.join(args.config_dir, "exp_*.json")))
        configs = [(os.path.splitext(os.path.basename(f))[0], json.load(open(f))) for f in files]
    else:
        base = json.load(open(args.base_config))
        configs = {"pool": enumerate_pool_configs, "stride": enumerate_stride_configs, "stride2": enumerate_stride_pair_configs}[args.mode](base)
    if args.max_configs is not None:
        configs = configs[:args.max_configs]
    vae, _, _, _ = load_vae(vae_type="884-16c-hy", vae_precision=args.vae_precision, vae_path=args.vae_path, device=device)
    names = list_clips(args.tensor_dir)[:args.max_files]
    for name, cfg in configs_of_rank(configs, rank, world):
        res = run_config(vae, cfg, args.tensor_dir, names, device, PRECISION_TO_TYPE[args.vae_precision],
                         os.path.join(args.save_dir, name) if args.save_dir else None)
        path = write_metrics(res, args.tensor_dir, f"reconstruction of {name}", os.path.join(args.metrics_dir, name))
        print(f"[rank {rank}] {name}: {res} -> {path}", flush=True)


if __name__ == "__main__":
    main()
