"""N > 1 host logic on CPU: world_size-2 gloo run of the stream sharding / timing reduction used by
bench.py --gpus N (one process per GPU, independent streams, no data-path collective)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from dpvo_b200 import multigpu
    r, w, _ = multigpu.init("gloo")
    mine = multigpu.streams_of_rank(5, r, w)
    secs = 1.0 + 0.5 * r                       # rank 1 is the slow one
    rate, tmax = multigpu.whole_job_rate(len(mine) * 10, secs)
    mx = multigpu.max_over_ranks([float(r), 3.0 - r])
    multigpu.barrier()
    out.put((r, mine, rate, tmax, mx))
    multigpu.finalize()


def test_two_rank_gloo_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res[0][1] + res[1][1]) == [0, 1, 2, 3, 4]            # every stream exactly once
    for r, mine, rate, tmax, mx in res:
        assert abs(tmax - 1.5) < 1e-9 and abs(rate - 50 / 1.5) < 1e-9  # all units / slowest rank
        assert mx == [1.0, 3.0]


def test_single_process_is_identity():
    from dpvo_b200 import multigpu
    assert multigpu.max_over_ranks([1.0, 2.0]) == [1.0, 2.0]
    assert multigpu.streams_of_rank(3, 0, 1) == [0, 1, 2]
    assert multigpu.whole_job_rate(10, 2.0) == (5.0, 2.0)


# ------------------------------------------------------------------ training: data-parallel gradient averaging
def _grad_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from dpvo_b200 import multigpu
    multigpu.init("gloo")
    torch.manual_seed(0)                                      # same parameters on every rank
    net = torch.nn.Sequential(torch.nn.Linear(40, 300), torch.nn.ReLU(), torch.nn.Linear(300, 300), torch.nn.Linear(300, 3))
    unused = torch.nn.Parameter(torch.ones(7))                # a parameter that gets no gradient: its bucket still reduces
    params = list(net.parameters()) + [unused]
    red = multigpu.GradReducer(params, bucket_mb=0.2)         # several buckets
    g = torch.Generator().manual_seed(100 + rank)             # different data per rank
    x = torch.randn(16, 40, generator=g)
    for p in params:
        p.grad = torch.zeros_like(p)
    red.begin()
    net(x).pow(2).mean().backward()
    local = [p.grad.clone() for p in params]
    red.finish()
    out.put((rank, [l.numpy() for l in local], [p.grad.numpy() for p in params], len(red.buckets), red.bytes_per_step))
    multigpu.barrier()
    multigpu.finalize()


def test_two_rank_gloo_gradient_allreduce_is_the_mean_of_local_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=180) for _ in ps), key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, a0, nb, nbytes), (_, l1, a1, _, _) = res
    assert nb > 1 and nbytes == 4 * sum(x.size for x in l0)
    for x0, x1, y0, y1 in zip(l0, l1, a0, a1):
        mean = 0.5 * (torch.from_numpy(x0) + torch.from_numpy(x1))
        assert torch.allclose(torch.from_numpy(y0), mean, atol=1e-7) and torch.allclose(torch.from_numpy(y1), mean, atol=1e-7)
