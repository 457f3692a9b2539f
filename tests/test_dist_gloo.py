"""N > 1 host logic on CPU: world_size-2 gloo process group (sharding map, max-over-ranks timing, aggregation)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from oracle import games as ogames


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, G, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from muzero_hypermodel_b200 import dist as mdist
    r, lr, w = mdist.init(backend="gloo")
    assert (r, w) == (rank, world)
    slot0 = mdist.first_slot(r, G)
    # each rank "plays" its shard: the cartpole initial states are a pure function of the GLOBAL slot
    env = ogames.CartPole(G, seed=3, slot0=slot0)
    states = env.state64()
    mdist.barrier()
    ms = 10.0 + 5.0 * rank                        # rank 1 is slower
    value, ms_max = mdist.aggregate_throughput(G * 50, ms)
    total_games = mdist.sum_over_ranks(G)
    # the two learner-side collectives: weight refresh from rank 0, gradient all-reduce (mean)
    torch.manual_seed(rank)
    sd = {"a.weight": torch.randn(3, 4), "b.bias": torch.randn(5)}
    got = mdist.broadcast_weights(sd, src=0, device="cpu")
    grads = [torch.full((2, 3), float(rank + 1)), torch.full((4,), float(10 * (rank + 1)))]
    mdist.allreduce_gradients(grads)
    out[rank] = (slot0, states, value, ms_max, total_games, {k: v.clone() for k, v in got.items()}, [g.clone() for g in grads])
    torch.distributed.destroy_process_group()


def test_world_size_2_sharding_and_aggregation():
    G, world = 6, 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, G, out), nprocs=world, join=True)
    assert out[0][0] == 0 and out[1][0] == G
    # N-invariance: the union of the shards equals one rank playing all 2G slots
    single = ogames.CartPole(2 * G, seed=3, slot0=0).state64()
    np.testing.assert_array_equal(np.concatenate([out[0][1], out[1][1]]), single)
    for r in range(world):
        slot0, _, value, ms_max, total_games, got, grads = out[r]
        torch.manual_seed(0)
        want = {"a.weight": torch.randn(3, 4), "b.bias": torch.randn(5)}
        assert all(torch.equal(got[k], want[k]) for k in want)          # every rank holds rank 0's weights
        assert torch.equal(grads[0], torch.full((2, 3), 1.5)) and torch.equal(grads[1], torch.full((4,), 15.0))
        assert ms_max == 15.0 and total_games == 2 * G                 # max over ranks, sum over ranks
        assert abs(value - (2 * G * 50) / 15e-3) < 1e-6                 # whole-job units / slowest rank's time
    from muzero_hypermodel_b200 import dist as mdist
    assert mdist.owner_of(7, G) == (1, 1) and mdist.owner_of(5, G) == (0, 5)
