"""Multi-GPU plumbing.  Self-play games are independent, so every rank runs its own batch with no
collective on the search path (the reference's N Gamer actors, Training/AlphaZero.py:525-578); the one
exchange is the merge of finished trajectories into the replay buffer (Training/Gamer.py:95 ships
each game to the ReplayBuffer actor): one variable-length all-gather of the compact move records.

Wire format = the engine's record arena (32-bit words: header, compact root state, (action, visits)
per root child — see csrc/mcts.cuh write_record), i.e. a few dozen bytes per position instead of the
reference's pickled float32 state tensors and length-A policy lists; states and policy targets are
rebuilt on arrival (nuzero_b200.selfplay.game_record / gamer.finished_game_of).
"""
import torch
import torch.distributed as dist


def all_gather_records(words, group=None):
    """words: 1-D int32 tensor (this rank's record words; CUDA for nccl, CPU for gloo).
    Returns the list of every rank's words (same order on all ranks)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [words]
    world = dist.get_world_size(group)
    n = torch.tensor([words.numel()], dtype=torch.int64, device=words.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c) for c in counts]
    width = max(max(counts), 1)
    send = torch.zeros(width, dtype=torch.int32, device=words.device)
    send[: words.numel()] = words
    recv = torch.empty(world * width, dtype=torch.int32, device=words.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return [recv[r * width: r * width + counts[r]] for r in range(world)]


def all_gather_indexed(words, offsets, group=None):
    """Record words plus the index of record offsets (engine.rec_index) of every rank: what
    DeviceReplayBuffer.ingest_parts needs.  Returns [(words_r, offsets_r)] in rank order.  One all-gather of the two
    lengths (the only host read) and one of the payload: each rank sends its words followed by its offsets in one row."""
    if not (dist.is_available() and dist.is_initialized()):
        return [(words, offsets.to(torch.int64))]
    world = dist.get_world_size(group)
    n = torch.tensor([words.numel(), offsets.numel()], dtype=torch.int64, device=words.device)
    counts = torch.zeros(world * 2, dtype=torch.int64, device=words.device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = counts.cpu().numpy().reshape(world, 2)
    ww, wo = max(int(counts[:, 0].max()), 1), max(int(counts[:, 1].max()), 1)
    send = torch.zeros(ww + wo, dtype=torch.int32, device=words.device)
    send[: words.numel()] = words
    send[ww: ww + offsets.numel()] = offsets.to(torch.int32)
    recv = torch.empty(world * (ww + wo), dtype=torch.int32, device=words.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, ww + wo)
    return [(recv[r, : int(counts[r, 0])], recv[r, ww: ww + int(counts[r, 1])].to(torch.int64)) for r in range(world)]


def global_game_index(rank, world, local_uid):
    """Rank r owns the games g = r (mod world)."""
    return local_uid * world + rank
