"""N>1 path on CPU: world_size-2 gloo — batch sharding + the single weight broadcast."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastvocoder_b200.sharding import shard_range


def test_shard_range_partitions_batch():
    for B in (1, 7, 8, 32, 64, 65):
        for G in (1, 2, 4, 8):
            parts = [shard_range(B, r, G) for r in range(G)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == B


def _worker(rank, world, port, q):
    import json
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from fastvocoder_b200 import build_generator
    from fastvocoder_b200.sharding import broadcast_weights, init_distributed, max_over_ranks
    from fastvocoder_b200.synthetic import synth_state_dict
    here = os.path.dirname(os.path.abspath(__file__))
    spec = json.load(open(os.path.join(here, "golden", "specs.json")))["basis-melgan-light"]
    r, _, w = init_distributed("gloo")
    assert (r, w) == (rank, world)
    model = build_generator("basis-melgan", spec["config"])
    if rank == 0:   # only rank 0 "loads the checkpoint"
        weights = synth_state_dict([(n, tuple(s)) for n, s in spec["spec_folded"]], seed=3)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    model.remove_weight_norm()
    before = model.packed_weights.clone()
    broadcast_weights(model, src=0)
    gathered = [torch.empty_like(model.packed_weights) for _ in range(world)]
    dist.all_gather(gathered, model.packed_weights)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    changed = not torch.equal(before, model.packed_weights)
    tmax = max_over_ranks(float(rank + 1), device="cpu")
    q.put((rank, same, changed, tmax, model._bound_key is None))
    dist.destroy_process_group()


def test_weight_broadcast_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]            # identical weights on both ranks
    assert res[0][2] is False and res[1][2] is True       # rank 1 received rank 0's weights
    assert [r[3] for r in res] == [2.0, 2.0]              # max over ranks
    assert all(r[4] for r in res)                         # device images are re-derived after the broadcast
