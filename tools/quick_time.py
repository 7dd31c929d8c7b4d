"""Scratch timing of the device path (not the bench): python tools/quick_time.py [B] [wave]"""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wave = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mode = sys.argv[3] if len(sys.argv) > 3 else "auto"
cfg = pkg.VocoderConfig()
voc = backend.Vocoder(cfg, None, wave=wave)
voc.set_option("gemm", mode)
codes = torch.from_numpy(np.random.default_rng(1).integers(0, 2048, (B, 64, 16), dtype=np.int64)).cuda()
out = torch.empty(B, voc.chunk_samples, dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(2):
        voc.infer_chunks_dev(codes, B, out, st.cuda_stream)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    n = 3
    for _ in range(n):
        voc.infer_chunks_dev(codes, B, out, st.cuda_stream)
    e1.record(st)
    st.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"B={B} wave={wave} gemm={mode}: {ms:.2f} ms/step  {ms/B:.2f} ms/chunk  xRT={B*5.12/(ms/1e3):.0f}  "
      f"TFLOP/s={B*cfg.flops_per_chunk()/(ms/1e3)/1e12:.1f}  out rms={out.float().pow(2).mean().sqrt().item():.3f}")
