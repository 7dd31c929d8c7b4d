"""Host-side multi-GPU logic on CPU: window-range partitioning, corpus balancing, and the
world_size-2 gather over gloo with a fake backend that follows the C library's plan."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import fake_chunk_fn
from oracle import stitch_oracle as SO


def _sharding():
    import importlib
    return importlib.import_module("qwen3-tts-axera-russian_b200.sharding")


def test_window_ranges_cover_exactly():
    S = _sharding()
    for nw in (1, 2, 5, 157, 209):
        for world in (1, 2, 4, 8):
            r = S.window_ranges(nw, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == nw
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert S.num_windows(7500) == 157 and S.num_windows(64) == 1 and S.num_windows(65) == 2
    assert S.num_windows(96) == 2 and S.num_windows(97) == 3
    assert all(S.num_windows(n) == len(SO.window_starts(n)) for n in range(1, 400))


def test_corpus_sharding_is_balanced():
    S = _sharding()
    rng = np.random.default_rng(2)
    lengths = np.clip(np.round(np.exp(rng.normal(np.log(100), 0.8, 1000))), 8, 3750).astype(int)
    bins = S.shard_corpus(lengths, 8)
    assert sorted(i for b in bins for i in b) == list(range(1000))
    loads = [sum(S.num_windows(lengths[i]) for i in b) for b in bins]
    assert max(loads) - min(loads) <= max(S.num_windows(n) for n in lengths)
    assert max(loads) / (sum(loads) / 8) < 1.02


class FakeRangeVocoder:
    """Follows voc_synthesize_range_dev's ownership rule in numpy, with a fake model."""

    def __init__(self, backend, Lc=122325):
        self.backend, self.Lc, self.fn = backend, Lc, fake_chunk_fn(Lc)

    def num_windows(self, n):
        return len(self.backend.plan(64, self.Lc, n)[0])

    def synthesize_range_pcm16(self, codes, w0, w1):
        n = len(codes)
        meta, total, _ = self.backend.plan(64, self.Lc, n)
        full = SO.to_pcm16(SO.synthesize(codes, self.fn, 64))
        if w0 >= w1:
            return 0, torch.zeros(0, dtype=torch.int16)
        begin = int(meta[w0][0])
        end = total if w1 == len(meta) else int(meta[w1][0])
        return begin, torch.from_numpy(full[begin:end].copy())


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    import importlib, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = importlib.import_module("qwen3-tts-axera-russian_b200.sharding")
        backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
        codes = (np.arange(n * 16, dtype=np.int64).reshape(n, 16) * 7919 + n) % 2048
        out = S.synthesize_sharded(FakeRangeVocoder(backend), codes, rank, world)
        if rank == 0:
            q.put(out)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [500, 97])
def test_two_rank_gloo_gather_equals_single(backend, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    codes = (np.arange(n * 16, dtype=np.int64).reshape(n, 16) * 7919 + n) % 2048
    ref = SO.to_pcm16(SO.synthesize(codes, fake_chunk_fn(122325), 64))
    assert np.array_equal(out, ref)
