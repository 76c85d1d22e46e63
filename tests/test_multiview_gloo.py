"""N>1 host logic on CPU: world_size-2 gloo processes exercise rank->view assignment, the unique-id
broadcast and the frame-ordering contract of the gather (the NCCL data path itself needs GPUs)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from bevy_gaussian_splatting_b200.multiview import MultiViewSession, broadcast_unique_id, view_for_rank

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ident = broadcast_unique_id(lambda: bytes(range(128)), rank, 0)
    sess = MultiViewSession(rank, world, 0, plugin=None)
    v = sess.view(64, 32)
    frame = np.full((32, 64, 4), rank + 1, np.uint8)
    got = sess.gather_host(frame)
    ok = ident == bytes(range(128))
    ok &= np.allclose(v.view_from_world, view_for_rank(rank, world, 64, 32).view_from_world)
    if rank == 0:
        ok &= got.shape == (world, 32, 64, 4) and all(int(got[r, 0, 0, 0]) == r + 1 for r in range(world))
    else:
        ok &= got is None
    q.put((rank, bool(ok), v.world_position.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_view_batch_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    # two distinct cameras on the r=5 circle around (0, 1.5, 0)
    (_, _, e0), (_, _, e1) = res
    assert np.allclose(e0, [0, 1.5, 5], atol=1e-5) and np.allclose(e1, [0, 1.5, -5], atol=1e-5)
