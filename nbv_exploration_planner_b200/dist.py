"""Launch plumbing for one-process-per-GPU jobs (``torchrun``): ``torch.distributed`` around the library.

The multi-GPU data path lives INSIDE ``libnbvgpu.so`` (include/nbvgpu.h, "multi-GPU contexts"): candidates sharded over the
ranks, the map replicated by ``nbvgpu_commit`` with an NCCL broadcast of {descriptor table | payload}, the per-shard best
records exchanged through GPU stores into shared host memory.  What is left for this module:

  * ``create_context``   -- hand the library's NCCL unique id from rank 0 to the other ranks (one 128-byte broadcast at start-up)
    and return each rank's ``NbvGpu`` (``nbvgpu_create_rank``); with one rank, a one-device ``nbvgpu_create_multi`` context;
  * ``shard_bounds``     -- the library's sharding rule, for callers that keep their shard in device memory
    (``NbvGpu.evaluate_shard_device``);
  * ``gather_sharded``   -- the widened rows (history-graph vertices for the frontier potential, positions for the ESDF
    interpolation) shard the same way in a torch program: every rank evaluates a contiguous block of the batch against its
    replica and the per-item results are all-gathered back into batch order;
  * ``broadcast_blocks`` / ``gather_best*`` -- the torch-tensor forms of the two exchanges (round 1's data path; kept for
    callers whose map updates or results already live in torch tensors, and exercised on CPU with gloo).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def create_context(local_device: int, stream=None, group=None):
    """This rank's ``NbvGpu`` of a multi-GPU context spanning the process group (or a one-device context without one)."""
    from .planner import NbvGpu
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        g = NbvGpu.multi([local_device])
    else:
        rank = dist.get_rank(group)
        dev = torch.device("cuda", local_device)
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(NbvGpu.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0, group=group)
        g = NbvGpu.rank(local_device, rank, world, uid.cpu().numpy().tobytes())
    if stream is not None:
        g.set_stream(stream)
    return g


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous blocks of ceil(n / world) candidates per rank (keeps tree-order locality)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def pull_slice(payload_bytes: int, world: int, rank: int) -> Tuple[int, int, int]:
    """The rule of a pulled commit chunk (csrc/nbvgpu.cu, group_commit): rank ``rank`` copies bytes [lo, hi) of the chunk's
    payload over its own PCIe link into its slice of the staging buffer; slices are ``slice`` bytes apart (a multiple of 256, so
    that the in-place all-gather of ``world`` x ``slice`` bytes lands every piece where the descriptors expect it).
    Returns (lo, hi, slice)."""
    per = (payload_bytes + world - 1) // world
    sl = (per + 255) // 256 * 256
    lo = min(payload_bytes, rank * sl)
    return lo, min(payload_bytes, lo + sl), sl


_PINNED = {}


def _pinned_staging(n: int, words: int) -> torch.Tensor:
    """One reusable pinned host buffer per row width (allocating pinned memory per call costs more than
    the copy itself for large maps)."""
    buf = _PINNED.get(words)
    if buf is None or buf.shape[0] < n:
        buf = torch.empty((max(n, 64), words), dtype=torch.float32).pin_memory()
        _PINNED[words] = buf
    return buf[:n]


def broadcast_blocks(keys: Optional[np.ndarray], payload: Optional[np.ndarray], payload_words_per_block: int,
                     device: torch.device, src: int = 0, group=None):
    """Broadcast changed blocks from `src`.  Returns (n_blocks, d_keys int32[n][3], d_payload f32[n][words])."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank == src:
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        n = len(keys)
        if len(np.unique(keys, axis=0)) != n:
            raise ValueError("broadcast_blocks: a block index appears twice in one batch")
        payload = np.ascontiguousarray(payload)
        if payload.dtype != np.float32:      # PACK_VOXBLOX words: reinterpret the bits, never convert numerically
            if payload.dtype.itemsize != 4:
                raise ValueError("payload must be float32 or 32-bit words")
            payload = payload.view(np.float32)
    else:
        n = 0
    if world > 1:
        hdr = torch.tensor([n], dtype=torch.int64, device=device)
        dist.broadcast(hdr, src, group=group)
        n = int(hdr.item())
    if n == 0:
        return 0, None, None
    # one buffer: [keys as float32 bit patterns | payload] so a single collective moves everything
    words = 3 + payload_words_per_block
    if rank == src:
        if device.type == "cuda":
            stage = _pinned_staging(n, words)
            buf_h = stage.numpy()
            buf_h[:, :3] = keys.view(np.float32)
            buf_h[:, 3:] = np.asarray(payload, np.float32).reshape(n, payload_words_per_block)
            buf = stage.to(device, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()  # the staging buffer is reused by the next call
        else:
            buf_h = np.empty((n, words), np.float32)
            buf_h[:, :3] = keys.view(np.float32)
            buf_h[:, 3:] = np.asarray(payload, np.float32).reshape(n, payload_words_per_block)
            buf = torch.from_numpy(buf_h)
    else:
        buf = torch.empty((n, words), dtype=torch.float32, device=device)
    if world > 1:
        dist.broadcast(buf, src, group=group)
    d_keys = buf[:, :3].contiguous().view(torch.int32)
    d_payload = buf[:, 3:].contiguous()
    return n, d_keys, d_payload


def gather_sharded(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-item results of a batch that was split with ``shard_bounds``.

    `local`: this rank's results, shape [hi - lo, ...] for (lo, hi) = shard_bounds(n_total, world, rank), on the device
    the process group works on.  Returns the results of the whole batch in batch order, shape [n_total, ...], on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    per = (n_total + world - 1) // world
    lo, hi = shard_bounds(n_total, world, rank)
    assert local.shape[0] == hi - lo, (local.shape, lo, hi)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)   # equal-sized pieces
    pad[: hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        rlo, rhi = shard_bounds(n_total, world, r)
        pieces.append(out[r * per: r * per + (rhi - rlo)])
    return torch.cat(pieces, 0)


def _pick(records) -> dict:
    """records: iterable of (global index or -1, score, gain, yaw).  Highest score wins, ties -> lowest
    global candidate index (= the reference's first strict maximum in candidate order)."""
    best = {"index": -1, "score": None, "gain": 0.0, "yaw": 0.0}
    for idx, score, gain, yaw in records:
        if idx < 0:
            continue
        if best["index"] < 0 or score > best["score"] or (score == best["score"] and idx < best["index"]):
            best = {"index": int(idx), "score": float(score), "gain": float(gain), "yaw": float(yaw)}
    return best


def gather_best(record: torch.Tensor, index_offset: int, group=None) -> dict:
    """record: float64[4] = (local index or -1, score, gain, yaw) on this rank's device.
    Returns the global winner; index -1 if no rank found a candidate above zero_gain."""
    rec = record.detach().cpu().clone()
    if rec[0] >= 0:
        rec[0] += index_offset
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        rec = rec.to(record.device)
        out = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(out, rec, group=group)
        allr = torch.stack(out).cpu().numpy()
    else:
        allr = rec.numpy()[None]
    return _pick(allr)


def gather_best_raw(d_best: torch.Tensor, n_total: int, group=None) -> dict:
    """d_best: the 32-byte nbvgpu_best record {int64 local index, double score, gain, yaw} as a uint8[32]
    device tensor, exactly as nbvgpu_gain_eval_best_device wrote it.  One all-gather of the raw records
    (NCCL over NVLink), one device->host read, decode + shard offsets + tie-break on the host."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        out = torch.empty(world * 32, dtype=torch.uint8, device=d_best.device)
        dist.all_gather_into_tensor(out, d_best, group=group)
    else:
        out = d_best
    raw = out.cpu().numpy()
    idx = raw.view(np.int64).reshape(world, 4)[:, 0]
    val = raw.view(np.float64).reshape(world, 4)[:, 1:]
    recs = []
    for r in range(world):
        lo, _ = shard_bounds(n_total, world, r)
        recs.append((int(idx[r]) + lo if idx[r] >= 0 else -1, val[r, 0], val[r, 1], val[r, 2]))
    return _pick(recs)


def best_record_as_f64(d_best_bytes: torch.Tensor) -> torch.Tensor:
    """View the 32-byte nbvgpu_best {int64 index, double score, gain, yaw} as float64[4] with the
    index converted (exact for |index| < 2^53)."""
    raw = d_best_bytes.view(torch.int64)
    f = d_best_bytes.view(torch.float64).clone()
    f[0] = raw[0].to(torch.float64)
    return f
