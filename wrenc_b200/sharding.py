"""Multi-GPU host logic: contiguous picture ranges per rank and the ordered gather of per-picture byte buffers
(SURVEY.md §8e).  Every picture is an independent IDR (reference src/main.rs:296,358), so ranks never exchange data on
the search path; the only collective is this gather of variable-length buffers to the writer rank."""
import torch
import torch.distributed as dist


def shard_range(n_pictures: int, world: int, rank: int):
    """Contiguous [begin, end) picture range of `rank`; earlier ranks take the remainder."""
    base, rem = divmod(n_pictures, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_in_order(buffers, dst=0, device=None):
    """buffers: list of bytes objects for this rank's pictures, in picture order.  Returns on `dst` the list of all
    pictures' buffers in global picture order (ranks hold contiguous ranges), None elsewhere.  Works with gloo (CPU
    tensors) and nccl (pass device)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device or torch.device("cpu")
    sizes = torch.tensor([len(b) for b in buffers], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(buffers)], dtype=torch.int64, device=dev))
    counts = [int(c.item()) for c in counts]
    maxc = max(counts) if counts else 0
    pad_sizes = torch.zeros(maxc, dtype=torch.int64, device=dev)
    pad_sizes[:len(buffers)] = sizes
    all_sizes = [torch.zeros(maxc, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, pad_sizes)
    totals = [int(s[:c].sum().item()) for s, c in zip(all_sizes, counts)]
    maxb = max(totals) if totals else 0
    payload = torch.zeros(max(maxb, 1), dtype=torch.uint8, device=dev)
    if buffers:
        flat = torch.frombuffer(bytearray(b"".join(buffers)), dtype=torch.uint8) if totals[rank] else torch.zeros(0, dtype=torch.uint8)
        payload[:flat.numel()] = flat.to(dev)
    gathered = [torch.zeros_like(payload) for _ in range(world)] if rank == dst else None
    dist.gather(payload, gathered, dst=dst)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        data = gathered[r].cpu().numpy().tobytes()
        off = 0
        for s in all_sizes[r][:counts[r]].tolist():
            out.append(data[off:off + s])
            off += s
    return out


_seq = 0


def gather_in_order_host(buffers, dst=0, barrier=None, rank=None, world=None, directory="/dev/shm"):
    """The same ordered gather for ranks of ONE node without touching a device or a collective's staging buffers
    ("D2H per GPU + host concatenation in pic_idx order"): every rank drops its pictures' sizes and
    bytes into a host shared-memory file, one barrier, and the writer rank concatenates the files in rank order.  `barrier`
    defaults to torch.distributed.barrier; rank / world default to the process group's."""
    global _seq
    import os
    import struct
    if rank is None or world is None:
        rank, world = dist.get_rank(), dist.get_world_size()
    barrier = barrier or dist.barrier
    job = os.environ.get("MASTER_PORT", "0") + "_" + os.environ.get("TORCHELASTIC_RUN_ID", "x")
    _seq += 1
    name = lambda r: os.path.join(directory, f"wrenc_b200_gather_{job}_{_seq}_{r}")  # noqa: E731
    with open(name(rank) + ".tmp", "wb") as f:
        f.write(struct.pack("<q", len(buffers)))
        f.write(struct.pack(f"<{len(buffers)}q", *[len(b) for b in buffers]))
        for b in buffers:
            f.write(b)
    os.replace(name(rank) + ".tmp", name(rank))
    barrier()
    if rank != dst:
        return None
    out = []
    for r in range(world):
        with open(name(r), "rb") as f:
            data = f.read()
        os.unlink(name(r))
        n = struct.unpack_from("<q", data, 0)[0]
        sizes = struct.unpack_from(f"<{n}q", data, 8)
        off = 8 + 8 * n
        for s in sizes:
            out.append(data[off:off + s])
            off += s
    return out
