"""One clip across several GPUs: the tiles of AutoencoderKLCausal3D's tiled encode / decode are independent
sub-model calls (own replicate padding, own GroupNorm statistics, own attention), so they are the unit of
partitioning (BASELINE config 4).  The reference runs the 720p decode replicated on every rank of
a torchrun job (pipeline_hunyuan_video.py:1074-1082); here each rank computes a cost-balanced subset of the
reference's tile grid and the results are exchanged with ONE collective per direction:

  encode: all_gather of the tile moments (1.1 MB each) -> every rank blends the full latent (needed by all)
  decode: every rank PUSHES each decoded tile into rank 0's tile arena over NVLink as soon as the tile is finished
          (peer-to-peer copy engine writes into rank 0's memory, mapped through CUDA IPC; no SMs, overlapped with the
          remaining tiles' compute), then ONE tiny all_reduce orders the pushes before rank 0's raster-order blend.
          Opt-in (HYVAE_TILE_PUSH=1).  The default is ONE NCCL gather of the tiles to rank 0: at 8 GPUs it measured faster
          (460 vs 480 ms per step, profiles/r02_bench_8gpu_{gather,push_v1}.json) than the first push version, whose
          cross-device torch copies enqueued event work on rank 0's GPU from seven other processes.

The blend chain is order dependent, so assembly always happens on complete tile grids, in the reference's order.
The tile functions / assemblers are injectable so the partition + exchange logic is testable on CPU (gloo).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import os

import torch
import torch.distributed as dist

from .. import _native as N
from .model import run_tiles


@dataclass(frozen=True)
class TileSpec:
    tt: int                      # temporal tile index
    i: int                       # spatial row
    j: int                       # spatial column
    t0: int
    t1: int
    h0: int
    h1: int
    w0: int
    w1: int

    @property
    def cost(self) -> int:
        return (self.t1 - self.t0) * (self.h1 - self.h0) * (self.w1 - self.w0)


def tile_grid(T: int, H: int, W: int, *, temporal: bool, spatial: bool, min_t: int, min_s: int, overlap: float) -> List[TileSpec]:
    """The reference's tile loops (autoencoder_kl_causal_3d.py:387-396,441-450,483-492,519-528) as data."""
    if temporal and T > min_t:
        tr = [(i, min(i + min_t + 1, T)) for i in range(0, T, int(min_t * (1 - overlap)))]
    else:
        tr = [(0, T)]
    specs = []
    for tt, (t0, t1) in enumerate(tr):
        if spatial and (H > min_s or W > min_s):
            s = int(min_s * (1 - overlap))
            for i, h0 in enumerate(range(0, H, s)):
                for j, w0 in enumerate(range(0, W, s)):
                    specs.append(TileSpec(tt, i, j, t0, t1, h0, min(h0 + min_s, H), w0, min(w0 + min_s, W)))
        else:
            specs.append(TileSpec(tt, 0, 0, t0, t1, 0, H, 0, W))
    return specs


def lpt_assign(costs: Sequence[int], world: int) -> List[int]:
    """Longest-processing-time-first: owner rank of every tile; deterministic, identical on all ranks."""
    load = [0] * world
    owner = [0] * len(costs)
    for k in sorted(range(len(costs)), key=lambda k: (-costs[k], k)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[k] = r
        load[r] += costs[k]
    return owner


class TileParallelVAE:
    def __init__(self, vae, rank: int, world: int, group=None,
                 enc_tile: Optional[Callable] = None, dec_tile: Optional[Callable] = None,
                 assemble_spatial: Optional[Callable] = None, assemble_temporal: Optional[Callable] = None):
        self.vae, self.rank, self.world, self.group = vae, rank, world, group
        self.enc_tile = enc_tile or vae._encode_tile
        self.dec_tile = dec_tile or vae._decode_tile
        self.assemble_spatial = assemble_spatial or vae._assemble_spatial
        self.assemble_temporal = assemble_temporal or vae._assemble_temporal

    # ---- exchange ---------------------------------------------------------------------------
    def _exchange(self, mine: dict, shapes: List[Tuple[int, ...]], owner: List[int], dtype, device, to_all: bool):
        """mine: {tile index: tensor}.  Returns the list of all tile tensors (on every rank if to_all, else on rank 0)."""
        numel = [int(torch.Size(s).numel()) for s in shapes]
        per_rank = [[k for k in range(len(shapes)) if owner[k] == r] for r in range(self.world)]
        cap = max(sum(numel[k] for k in ks) for ks in per_rank)
        send = torch.empty(max(cap, 1), dtype=dtype, device=device)
        off = 0
        for k in per_rank[self.rank]:
            send[off:off + numel[k]].copy_(mine[k].reshape(-1))
            off += numel[k]
        if to_all:
            recv = torch.empty(self.world * send.numel(), dtype=dtype, device=device)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            bufs = recv.view(self.world, -1)
        else:
            lst = [torch.empty_like(send) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(send, lst, dst=0, group=self.group)
            if self.rank != 0:
                return None
            bufs = lst
        tiles: List[Optional[torch.Tensor]] = [None] * len(shapes)
        for r, ks in enumerate(per_rank):
            off = 0
            for k in ks:
                tiles[k] = mine[k] if r == self.rank else bufs[r][off:off + numel[k]].view(shapes[k])
                off += numel[k]
        return tiles

    # ---- peer-memory push (decode direction) --------------------------------------------------
    def _push_arena(self, shapes: List[Tuple[int, ...]], owner: List[int], dtype, device):
        """Rank 0's tile arena for this tile set, mapped into every rank: [(views of arena 0), (views of arena 1)] with one
        view per tile that rank 0 does not own (None for its own), or None when peer memory is not available.

        Two arenas alternate between calls: a rank may start pushing the tiles of call i+1 while rank 0 still blends call i,
        and it cannot reach call i+2 before rank 0 has passed the ordering all_reduce of call i+1, which rank 0 enqueues
        after the blends of call i.  The arenas are allocated once per (tile set, dtype) and shared through CUDA IPC
        (torch.multiprocessing's reduce_tensor: cudaIpcGetMemHandle / cudaIpcOpenMemHandle), so a push is a plain
        device-to-device copy into peer memory: copy engines over NVLink, no SM and no NCCL kernel."""
        if os.environ.get("HYVAE_TILE_PUSH", "0") != "1" or device.type != "cuda" or self.world < 2:
            return None
        key = (tuple(shapes), tuple(owner), dtype)
        cache = self.__dict__.setdefault("_arenas", {})
        if key in cache:
            return cache[key]
        numel = [int(torch.Size(s).numel()) for s in shapes]
        offs, total = [], 0
        for k, n in enumerate(numel):
            offs.append(total)
            if owner[k] != 0:
                total += (n + 127) // 128 * 128      # keep every tile 256-byte aligned (16-byte vectors in the blend kernel)
        views, ok = None, 1

        def note(e):
            if os.environ.get("HYVAE_TILE_PUSH_DEBUG"):
                print(f"[tile_parallel] rank {self.rank}: peer-memory push unavailable: {type(e).__name__}: {e}", flush=True)

        payload = [None]
        if self.rank == 0:   # a failure here still reaches the broadcast below (payload None), so no rank waits alone
            try:
                from torch.multiprocessing.reductions import reduce_tensor
                arenas = [torch.empty(max(total, 1), dtype=dtype, device=device) for _ in range(2)]
                payload = [[reduce_tensor(a) for a in arenas]]
                self._arena_owner = getattr(self, "_arena_owner", []) + [arenas]   # rank 0 keeps the allocations alive
            except Exception as e:  # noqa: BLE001
                note(e)
                payload = [None]
        dist.broadcast_object_list(payload, src=0, group=self.group)
        try:
            if payload[0] is None:
                raise RuntimeError("rank 0 could not export its arena")
            if self.rank != 0:
                arenas = [fn(*args) for fn, args in payload[0]]
            views = [[None if owner[k] == 0 else a[offs[k]:offs[k] + numel[k]].view(shapes[k]) for k in range(len(shapes))] for a in arenas]
            if self.rank != 0:   # one small write proves the mapping (and lets torch enable peer access) before the hot loop relies on it
                probe = torch.zeros(1, dtype=dtype, device=device)
                arenas[0][:1].copy_(probe)
                torch.cuda.synchronize(device)
        except Exception as e:  # noqa: BLE001 - any failure (no peer access, IPC refused in this container) selects the gather
            note(e)
            ok, views = 0, None
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)   # every rank takes the same path
        if int(flag.item()) == 0:
            views = None
        cache[key] = views
        return views

    # ---- one direction ----------------------------------------------------------------------
    def _run(self, x: torch.Tensor, encode: bool, to_all: bool, post: bool = False):
        v = self.vae
        B, _, T, H, W = x.shape
        ov = v.tile_overlap_factor
        if encode:
            min_t, min_s, fn = v.tile_sample_min_tsize, v.tile_sample_min_size, self.enc_tile
            ext_s, lim_s = int(v.tile_latent_min_size * ov), v.tile_latent_min_size - int(v.tile_latent_min_size * ov)
            ext_t, lim_t = int(v.tile_latent_min_tsize * ov), v.tile_latent_min_tsize - int(v.tile_latent_min_tsize * ov)
            cout, r_t, r_s = 2 * v.config.latent_channels, v.config.time_compression_ratio, v.config.spatial_compression_ratio
            oshape = lambda s: (B, cout, (s.t1 - s.t0 - 1) // r_t + 1, -(-(s.h1 - s.h0) // r_s), -(-(s.w1 - s.w0) // r_s))
        else:
            min_t, min_s, fn = v.tile_latent_min_tsize, v.tile_latent_min_size, self.dec_tile
            ext_s, lim_s = int(v.tile_sample_min_size * ov), v.tile_sample_min_size - int(v.tile_sample_min_size * ov)
            ext_t, lim_t = int(v.tile_sample_min_tsize * ov), v.tile_sample_min_tsize - int(v.tile_sample_min_tsize * ov)
            cout, r_t, r_s = v.config.out_channels, v.config.time_compression_ratio, v.config.spatial_compression_ratio
            oshape = lambda s: (B, cout, (s.t1 - s.t0 - 1) * r_t + 1, (s.h1 - s.h0) * r_s, (s.w1 - s.w0) * r_s)
        specs = tile_grid(T, H, W, temporal=v.use_temporal_tiling, spatial=v.use_spatial_tiling, min_t=min_t, min_s=min_s, overlap=ov)
        owner = lpt_assign([s.cost for s in specs], self.world)
        mine_k = [k for k in range(len(specs)) if owner[k] == self.rank]
        shapes = [oshape(s) for s in specs]
        cut = lambda s: fn(x[:, :, s.t0:s.t1, s.h0:s.h1, s.w0:s.w1]).contiguous()
        out_dtype = getattr(v, "dtype", x.dtype)
        arena = None if to_all else self._push_arena(shapes, owner, out_dtype, x.device)
        if arena is not None:
            self._push_calls = getattr(self, "_push_calls", 0) + 1
            slots = arena[self._push_calls & 1]
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=x.device)
            copy_stream = self._copy_stream
            copy_stream.wait_stream(torch.cuda.current_stream(x.device))
        if not to_all:
            self.last_exchange = "push" if arena is not None else "gather"

        def push(k, t):
            """Peer write of a finished tile into rank 0's arena, on the copy stream, behind the tile's own stream."""
            if arena is not None and self.rank != 0:
                copy_stream.wait_stream(torch.cuda.current_stream(x.device))
                with torch.cuda.stream(copy_stream):
                    N.peer_copy(slots[k], t)     # one cudaMemcpyAsync on THIS device's stream; nothing runs on rank 0's GPU
                t.record_stream(copy_stream)
            return t

        outs = run_tiles([(lambda k=k: push(k, cut(specs[k]))) for k in mine_k], getattr(v, "tile_streams", 1) if x.is_cuda else 1)
        if hasattr(v, "_guard_tiles"):   # fp16-operand range guard of a bf16 model (model.py): re-run overflowing tiles in bf16
            outs = v._guard_tiles(outs, lambda i: push(mine_k[i], cut(specs[mine_k[i]])))
        mine = {}
        for k, t in zip(mine_k, outs):
            assert tuple(t.shape) == shapes[k], (tuple(t.shape), shapes[k])
            mine[k] = t
        if arena is not None:
            # ONE ordering collective: stream-ordered behind this rank's pushes; when it completes on rank 0 every peer
            # write has landed in the arena
            torch.cuda.current_stream(x.device).wait_stream(copy_stream)
            done = torch.zeros(1, dtype=torch.int32, device=x.device)
            dist.all_reduce(done, group=self.group)
            if self.rank != 0:
                return None
            tiles = [mine[k] if owner[k] == 0 else slots[k] for k in range(len(specs))]
        else:
            dtype = next(iter(mine.values())).dtype if mine else out_dtype
            tiles = self._exchange(mine, shapes, owner, dtype, x.device, to_all)
        if tiles is None:
            return None
        # assemble: spatial grids per temporal tile, then the temporal chain
        n_tt = max(s.tt for s in specs) + 1
        row = []
        for tt in range(n_tt):
            ks = [k for k, s in enumerate(specs) if s.tt == tt]
            ni = max(specs[k].i for k in ks) + 1
            grid = [[None] * (max(specs[k].j for k in ks) + 1) for _ in range(ni)]
            for k in ks:
                grid[specs[k].i][specs[k].j] = tiles[k]
            single = len(ks) == 1 and not (v.use_spatial_tiling and (H > min_s or W > min_s))
            last = n_tt == 1 and not (v.use_temporal_tiling and T > min_t)   # no temporal assembly follows
            if single:
                row.append((grid[0][0], 1 if tt > 0 else 0))
            elif post and last:
                row.append((self.assemble_spatial(grid, ext_s, lim_s, True), 0))
            else:
                row.append((self.assemble_spatial(grid, ext_s, lim_s), 1 if tt > 0 else 0))
        if n_tt == 1 and not (v.use_temporal_tiling and T > min_t):
            if post and single:
                return N.image_postprocess(row[0][0])
            return row[0][0]
        return self.assemble_temporal(row, ext_t, lim_t, True) if post else self.assemble_temporal(row, ext_t, lim_t)

    def close(self):
        """Drop the peer mappings of rank 0's arenas (call on every rank before the process group is destroyed, so that the
        consumers release the CUDA IPC handles before the producer frees the memory)."""
        if getattr(self, "_arenas", None):
            self._arenas.clear()
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self._arena_owner = []

    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """Blended moments of the whole clip, on every rank."""
        return self._run(x, True, True)

    def decode(self, z: torch.Tensor) -> Optional[torch.Tensor]:
        """Decoded clip on rank 0 (None elsewhere)."""
        return self._run(z, False, False)

    def decode_to_image(self, z: torch.Tensor, latent_scale: float = 1.0, latent_shift: float = 0.0) -> Optional[torch.Tensor]:
        """AutoencoderKLCausal3D.decode_to_image with the tiles sharded over the ranks: fp32 image on rank 0."""
        v = self.vae
        v._latent_affine = None if (latent_scale == 1.0 and latent_shift == 0.0) else (float(latent_scale), float(latent_shift))
        try:
            return self._run(z, False, False, post=True)
        finally:
            v._latent_affine = None

    def roundtrip(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        """forward(sample_posterior=False): encode -> mode() -> decode (autoencoder_kl_causal_3d.py:543-578)."""
        moments = self.encode_moments(x)
        mean = moments[:, : moments.shape[1] // 2]
        return self.decode(mean)
