"""world_size-2 gloo test (CPU) of the multi-GPU host logic: rollout sharding, the 16-byte
(cost, index) all-gather and the lowest-index tie-break every rank applies afterwards."""
import os
import socket

import numpy as np
import pytest

from bipedal_locomotion_framework_b200 import sharding


def test_shard_rollouts_partitions_exactly():
    for n in (0, 1, 7, 4096, 335544):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_rollouts(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_pair_packing_roundtrip():
    for cost, idx in ((0.0, 0), (1.5, 4), (179272.25, 2 ** 40), (float("inf"), -1)):
        assert sharding.unpack_pair(*sharding.pack_pair(cost, idx)) == (cost, idx)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rollouts, rollout_len, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ccm_oracle
        from bipedal_locomotion_framework_b200 import synthetic as syn
        first, count = sharding.shard_rollouts(n_rollouts, world, rank)
        # each rank evaluates only its own rollouts (here with the CPU oracle standing in for the
        # GPU kernel: this test is about the exchange, not the arithmetic)
        st = syn.make_states(count * rollout_len, seed=45, start=first * rollout_len)
        w = ccm_oracle.eval_batch_states(st, mask=1)["wrench"]
        cost = ccm_oracle.rollout_cost(w, rollout_len, [0, 0, 30.0, 0, 0, 0], [1.0, 10.0])
        j = int(np.argmin(cost))
        best = torch.tensor(sharding.pack_pair(float(cost[j]), first + j), dtype=torch.int64)
        gathered = sharding.all_gather_pairs(best, world, dist)
        pairs = [sharding.unpack_pair(int(a), int(b)) for a, b in gathered.tolist()]
        winner = min(pairs, key=lambda p: (p[0], p[1]))
        q.put((rank, winner, (first, count)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_argmin_exchange_matches_single_process():
    import torch.multiprocessing as mp
    from oracle import ccm_oracle
    from bipedal_locomotion_framework_b200 import synthetic as syn
    ccm_oracle.build()
    world, n_rollouts, rollout_len = 2, 37, 20
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rollouts, rollout_len, q))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process answer over the whole batch
    st = syn.make_states(n_rollouts * rollout_len, seed=45)
    w = ccm_oracle.eval_batch_states(st, mask=1)["wrench"]
    cost = ccm_oracle.rollout_cost(w, rollout_len, [0, 0, 30.0, 0, 0, 0], [1.0, 10.0])
    expect = (float(cost.min()), int(np.argmin(cost)))
    winners = {r: w_ for r, w_, _ in results}
    assert winners[0] == winners[1] == expect          # every rank picks the same global arg-min
    spans = sorted(s for _, _, s in results)
    assert spans[0][0] == 0 and spans[0][0] + spans[0][1] == spans[1][0]
    assert spans[1][0] + spans[1][1] == n_rollouts
