"""N > 1 host logic on CPU: two ranks over gloo (127.0.0.1).  Covers what bench.py / a multi-GPU
host does around the engine: the window partition between ranks (no window lost or duplicated, same
sequence as the reference's GetJob, 1_9_7File.pb:2077-2092), one broadcast of the table image from
rank 0, the max-over-ranks timing reduction and 'first hit stops everybody'."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bsgs_cuda_b200 as b
    from tests import pyec
    from oracle import oracle as orc
    T, B, p, w, htsz = 8, 2, 4, 500, 7
    pk = 1 << 40
    secret = pk + 123456789
    target = pyec.mul(pyec.G, secret)
    # --- strided partition: rank r owns windows r, r+R, ...
    s = b.Search(T, B, p, w, htsz, pk, target, rank=rank, world=world)
    mine = [s.window_key(j) % (1 << 62) for j in range(5)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # --- table image: rank 0 builds it, one broadcast, everybody holds the same bytes
    nbytes = orc.lib().orc_htgpu_bytes(w, htsz)
    img = torch.zeros(nbytes, dtype=torch.uint8)
    if rank == 0:
        cpu, gpu = orc.build_ht(orc.baby_keys(w), htsz)
        img.copy_(torch.from_numpy(gpu))
    dist.broadcast(img, src=0)
    csum = int(img.to(torch.int64).sum())
    # --- timing reduction as in bench.py: max over ranks
    tt = torch.tensor([10.0 + rank, 20.0 - rank], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    # --- a hit on one rank stops all ranks (host-side flag, like `quit` F197:4287-4290)
    flag = torch.tensor([1 if rank == 1 else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if rank == 0:
        single = b.Search(T, B, p, w, htsz, pk, target)
        seq = [single.window_key(j) % (1 << 62) for j in range(5 * world)]
        out.put({"gathered": gathered, "seq": seq, "csum": csum, "tt": tt.tolist(), "flag": int(flag)})
    else:
        out.put({"csum": csum})
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_partition_broadcast_and_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    r0 = next(r for r in res if "seq" in r)
    inter = [r0["gathered"][j % world][j // world] for j in range(5 * world)]
    assert inter == r0["seq"]                                   # the union of the ranks' windows is GetJob's sequence
    assert len(set(r0["seq"])) == len(r0["seq"])
    assert len({r["csum"] for r in res}) == 1 and r0["csum"] > 0
    assert r0["tt"] == [11.0, 20.0] and r0["flag"] == 1


def test_contiguous_block_partition_of_bench():
    """bench.py gives rank r the windows [r*ceil(W/R), (r+1)*ceil(W/R)): together they cover all."""
    import bsgs_cuda_b200 as b
    from tests import pyec
    s = b.Search(512, 148, 256, 1 << 30, 28, 1 << 64, pyec.mul(pyec.G, (1 << 64) + 5))
    total = s.windows_needed((1 << 65) - 1)
    assert total == 223                                         # 2^64 / (4*N*w) = 221.4 -> 222 windows with key_j <= width, plus the one launched before the end test (1_9_7File.pb:2512-2518)
    for world in (1, 2, 4, 8):
        per = (total + world - 1) // world
        cover = set()
        for r in range(world):
            cover |= {(r * per + j) for j in range(per) if r * per + j < total}
        assert cover == set(range(total))
