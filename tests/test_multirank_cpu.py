"""world_size-2 gloo test (CPU) of the data-parallel host logic: batch sharding + the K-float loss all-reduce give
the same global means as the unsharded batch, and per-sample losses of a shard equal those of the full batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import loss_port as P
    from unsupervised_depth_opticalflow_egomotion_b200 import parallel
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
    torch.set_num_threads(1)
    B = 5                                        # odd on purpose: shards of 3 and 2
    t = make_triplet(B, 32, 64, 3, 1, seed=9, flow_px=2.0)
    a, b = parallel.shard_range(B, rank, world)
    sl = lambda x: x[a:b].contiguous()
    local = P.flow_mode_loss(sl(t.img_l), sl(t.img), sl(t.img_r), [sl(f) for f in t.flows_fwd], [sl(f) for f in t.flows_bwd], 3)
    means = parallel.global_loss_means(local, B)
    total = parallel.weighted_total(means, P.FLOW_WEIGHTS)
    # the preallocated, graph-capturable form of the same collective: (K, B_local) matrix -> (K,) global means
    keys = list(P.FLOW_WEIGHTS)
    lar = parallel.LossAllReduce(len(keys), B, "cpu")
    vec = lar(torch.stack([local[k].detach() for k in keys]))
    # plain Python payloads: tensors in a Queue travel as file descriptors, which break if the sender exits first
    q.put((rank, (a, b), {k: v.tolist() for k, v in local.items()}, {k: float(v) for k, v in means.items()}, float(total),
           {k: float(vec[i]) for i, k in enumerate(keys)}))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_losses_match_unsharded():
    sys.path.insert(0, ROOT)
    from oracle import loss_port as P
    from unsupervised_depth_opticalflow_egomotion_b200 import parallel
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
    assert [parallel.shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [parallel.shard_range(32, r, 8) for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.set_num_threads(1)
    t = make_triplet(5, 32, 64, 3, 1, seed=9, flow_px=2.0)
    full = P.flow_mode_loss(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, 3)
    for rank, (a, b), local, means, total, vec in res:
        for k in full:
            assert torch.allclose(torch.tensor(local[k]), full[k][a:b], rtol=1e-6, atol=0), k      # per-sample values do not depend on the shard
            assert abs(means[k] - float(full[k].mean())) <= 1e-6 * abs(float(full[k].mean())), k
            assert abs(vec[k] - means[k]) <= 1e-6 * abs(means[k]), k                                # LossAllReduce == global_loss_means
        assert abs(total - float(P.weighted_total(full, P.FLOW_WEIGHTS))) <= 1e-6 * abs(total)
    assert res[0][3] == res[1][3]                                                    # every rank holds the same global means
