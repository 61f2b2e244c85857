"""N>1 path on CPU: bridges shard contiguously, per-channel summaries are
gathered to rank 0 (gloo, world_size 2 and 3 with ragged shards)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_py as O
from igate4xsoftphonedsp_b200 import sharding, synth
from igate4xsoftphonedsp_b200._native import SUMMARY_DT


def test_partition_covers_every_bridge_once():
    for world in (1, 2, 3, 4, 8):
        for B in (1, 7, 8, 1024, 16384):
            r = [sharding.bridge_range(k, world, B) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
            assert sum(sharding.shard_sizes(world, B)) == B
    assert sharding.channel_range(1, 2, 10, 4) == (20, 40)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, G, F, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b0, b1 = sharding.bridge_range(rank, world, B)
    nb = b1 - b0
    # every rank generates and processes ITS bridges only (no data-path collective)
    pcm = synth.pcm_noise_tone(F, nb * G, ch0=b0 * G)
    law = synth.laws(nb * G, ch0=b0 * G)
    codes = np.stack([O.g711_encode(pcm[:, c], int(law[c])) for c in range(nb * G)], axis=1) if nb else \
        np.zeros((F, 0, 160), np.uint8)
    gain = synth.gains(F, nb, G)
    if nb:
        _, _, meter, _ = O.process_batch(codes, law, gain, synth.out_laws(nb, b0), G)
        local = O.event_summary(meter, gain).astype(SUMMARY_DT)
    else:
        local = np.zeros(0, SUMMARY_DT)
    got = sharding.gather_records(local, B, G)
    if rank == 0:
        q.put(got.tobytes())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 6), (3, 7)])
def test_summary_gather_matches_single_process(world, B):
    G, F = 4, 30
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, G, F, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = np.frombuffer(q.get(timeout=120), dtype=SUMMARY_DT)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pcm = synth.pcm_noise_tone(F, B * G)
    law = synth.laws(B * G)
    codes = np.stack([O.g711_encode(pcm[:, c], int(law[c])) for c in range(B * G)], axis=1)
    gain = synth.gains(F, B, G)
    _, _, meter, _ = O.process_batch(codes, law, gain, synth.out_laws(B), G)
    want = O.event_summary(meter, gain)
    assert got.shape == want.shape and got.tobytes() == want.astype(SUMMARY_DT).tobytes()
