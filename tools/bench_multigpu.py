#!/usr/bin/env python3
"""BASELINE.json configs[3] and configs[4] on N GPUs of one box (one process per GPU, torchrun):

  * a 10-minute utterance (7 500 frames, 157 windows) split into contiguous window ranges, each
    rank stitching the span its windows own, int16 PCM gathered to rank 0 over NCCL and checked
    bit-equal against the same request synthesised on one GPU;
  * a 1 000-utterance mixed-length corpus (lengths round(exp(N(ln 100, 0.8^2))) clipped to
    [8, 3750] frames, seed 2) sharded by longest-processing-time balance, no collective at all.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/bench_multigpu.py [--corpus 1000]

Prints one JSON line per config on rank 0.  Times are CUDA-event / synchronised wall times, max over
ranks; audio seconds are nominal (frames * 0.08 s)."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=7500)
    ap.add_argument("--corpus", type=int, default=1000)
    ap.add_argument("--wave", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--coalesce", type=int, default=64, help="requests per batched call")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
    backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
    S = importlib.import_module("qwen3-tts-axera-russian_b200.sharding")
    cfg = pkg.VocoderConfig()
    voc = backend.Vocoder(cfg, pkg.init_weights(cfg, 0), device=local, wave=args.wave)

    def sync_max(t):
        if world == 1:
            return t
        x = torch.tensor([t], dtype=torch.float64, device=dev)
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return float(x.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- config 4: one long utterance, window ranges
    n = args.frames
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (n, 16), dtype=np.int64)
    out = S.synthesize_sharded(voc, codes, rank, world)          # warm-up (allocations, tensor maps)
    ts = []
    for _ in range(args.reps):
        barrier()
        t0 = time.perf_counter()
        out = S.synthesize_sharded(voc, codes, rank, world)
        barrier()
        ts.append(sync_max(time.perf_counter() - t0))
    if rank == 0:
        ref = voc.synthesize_pcm16(codes)                         # the same request on this GPU alone
        t0 = time.perf_counter(); ref = voc.synthesize_pcm16(codes); t1 = time.perf_counter() - t0
        ok = bool(out.shape == ref.shape and np.array_equal(out, ref))
        t = min(ts)
        print(json.dumps({"config": "10-minute utterance, window ranges across GPUs (BASELINE configs[3])",
                          "n_gpus": world, "frames": n, "windows": voc.num_windows(n), "samples": int(len(out)),
                          "bit_equal_to_single_gpu": ok, "seconds": t, "xrt": n * 0.08 / t,
                          "single_gpu_seconds": t1, "single_gpu_xrt": n * 0.08 / t1,
                          "includes": "H2D codes, kernels, stitch, NCCL PCM gather to rank 0, D2H"}), flush=True)
        assert ok, "sharded PCM differs from the single-GPU result"

    # ---- config 5: corpus of whole utterances, sharded; no collective
    if args.corpus > 0:
        rng = np.random.default_rng(2)
        lengths = np.clip(np.round(np.exp(rng.normal(np.log(100), 0.8, args.corpus))), 8, 3750).astype(int)
        bins = S.shard_corpus(lengths, world, cfg.chunk_frames)
        mine = bins[rank]
        utts = [np.random.default_rng(1000 + i).integers(0, cfg.codebook_size, (int(lengths[i]), 16), dtype=np.int64)
                for i in mine]
        for u in utts[:3]:
            voc.synthesize_pcm16(u)
        barrier()
        t0 = time.perf_counter()
        total = 0
        for u in utts:
            total += len(voc.synthesize_pcm16(u))
        barrier()
        t_single = sync_max(time.perf_counter() - t0)
        # the same corpus through the batched entry point (requests coalesced 64 at a time)
        voc.synthesize_batch_pcm16(utts[:8])
        barrier()
        t0 = time.perf_counter()
        total_b = 0
        for i in range(0, len(utts), args.coalesce):
            total_b += sum(len(x) for x in voc.synthesize_batch_pcm16(utts[i:i + args.coalesce]))
        barrier()
        t_batch = sync_max(time.perf_counter() - t0)
        assert total_b == total
        frames = int(lengths.sum())
        if rank == 0:
            print(json.dumps({"config": f"{args.corpus}-utterance mixed-length corpus sharded across GPUs (BASELINE configs[4])",
                              "n_gpus": world, "frames": frames, "audio_seconds": frames * 0.08,
                              "one_request_at_a_time": {"seconds": t_single, "xrt": frames * 0.08 / t_single,
                                                        "call": "voc_synthesize_pcm16"},
                              "coalesced": {"seconds": t_batch, "xrt": frames * 0.08 / t_batch, "requests_per_call": args.coalesce,
                                            "call": "voc_synthesize_batch_pcm16"},
                              "utterances_on_rank0": len(mine),
                              "note": "host codes in, host PCM out; no collective"}), flush=True)
    voc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
