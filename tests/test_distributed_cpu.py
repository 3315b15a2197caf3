"""CPU, world_size 2 over gloo: the utterance-sharding + all-gather logic of distributed.py.
Sharding must not change a single bit of the result (no arithmetic crosses utterances); the compute
callbacks here are the CPU oracle, injected, because this suite has no GPU."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from lsm_speech_classifier_b200 import distributed as D


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 2, 5, 8, 2400, 2401, 105000):
        for ws in (1, 2, 3, 4, 8):
            seen = []
            for r in range(ws):
                lo, hi, per = D.shard_bounds(n, r, ws)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lsm_speech_classifier_b200 import synth, filterbank
        from lsm_speech_classifier_b200.reservoir import SimulationParams, build_reservoir
        from oracle import coracle

        class FakeFrontend:   # geometry only; compute is injected
            rows, steps = 128, 400

        class FakeLsm:
            num_output_neurons = 400

        rng = np.random.default_rng(3)
        spikes = (rng.random((n, 128, 400)) < 0.03).astype(np.uint8)
        r = build_reservoir(SimulationParams(mean_weight=0.0105, input_spike_times=spikes[0] if n else np.zeros((128, 400), np.uint8)))
        keys = ['spike_counts', 'mean_spike_times']

        def compute(block):
            return coracle.reservoir_run(r, block, 0b101, True, False, nthreads=2)[0]

        got = D.sharded_features(FakeLsm(), spikes, keys, compute=compute)
        want = compute(spikes) if n else np.zeros((0, 800))
        ok_feat = got.shape == want.shape and np.array_equal(got, want)

        pcm = rng.standard_normal((n, 16000)).astype(np.float32)

        def enc(block):   # any deterministic per-utterance byte function will do for the plumbing
            out = np.zeros((len(block), 128, 400), np.uint8)
            out[:, :, 0] = (np.abs(block[:, :128]) * 10).astype(np.uint8)
            return out

        got_s = D.sharded_spikes(FakeFrontend(), pcm, compute=enc)
        ok_spk = np.array_equal(got_s, enc(pcm))
        q.put((rank, bool(ok_feat), bool(ok_spk), D.is_main()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 2, 1])
def test_world_size_two_gloo_matches_single_process(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] and r[2] for r in res)
    assert [r[3] for r in res] == [True, False]
