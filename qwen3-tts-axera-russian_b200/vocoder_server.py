#!/usr/bin/env python3
"""Drop-in vocoder server: the reference's unix-socket protocol in front of the B200 backend.

Mirrors the public surface of ``/root/reference/dual_npu/vocoder_server.py`` -- class
``VocoderServer(model_path, socket_path)`` with ``max_tokens``, ``_inference_chunk``,
``synthesize``, ``serve`` and the same CLI (``--model``, ``--socket``) -- so the launcher
(``dual_npu/launch_qwen3_tts.sh:185-187``) and the client (``dual_npu/tts_client.py:78-108``)
work unchanged.  The backend is selected by the model file's suffix, exactly like the
reference selects ONNX vs RKNN (``vocoder_server.py:36``): ``*.b200voc`` -> this backend.

Wire protocol (``vocoder_server.py:8-12``), unchanged:
  client -> server : int32 LE n_tokens, then n_tokens*16 int64 LE codes (row-major [n, 16])
  server -> client : int32 LE n_samples, then n_samples int16 LE PCM
  n_tokens <= 0 or > 10000, a short body, or any error: the connection is closed without a
  reply (``:149-151,162-164,180-183``).

Usage:
  python3 vocoder_server.py --model vocoder.b200voc --socket /tmp/qwen3_voc.sock
"""
from __future__ import annotations

import argparse
import os
import signal
import socket
import struct
import time

import numpy as np

SAMPLE_RATE = 24000
SAMPLES_PER_TOKEN = 1920
MAX_REQUEST_TOKENS = 10000            # vocoder_server.py:149
CODEBOOKS = 16


def _recv_exact(conn: socket.socket, n: int) -> bytes:
    parts, got = [], 0
    while got < n:
        piece = conn.recv(min(65536, n - got))
        if not piece:
            break
        parts.append(piece)
        got += len(piece)
    return b"".join(parts)


class VocoderServer:
    def __init__(self, model_path, socket_path="/tmp/qwen3_voc.sock", device=0, wave=32,
                 install_signal_handlers=True):
        from .backend import Vocoder
        from .weights import MODEL_SUFFIX
        self.socket_path = socket_path
        self.is_onnx = False
        if not str(model_path).endswith(MODEL_SUFFIX):
            raise RuntimeError(
                f"{model_path}: this server only drives the B200 backend ({MODEL_SUFFIX} files); "
                "use the reference vocoder_server.py for .onnx/.rknn models (no CPU fallback here)")
        self.voc = Vocoder.from_file(model_path, device=device, wave=wave)
        self.max_tokens = self.voc.max_tokens
        print(f"Vocoder: B200 CUDA (device {device}), max_tokens={self.max_tokens}")
        self._running = True
        if install_signal_handlers:
            signal.signal(signal.SIGINT, self._signal_handler)
            signal.signal(signal.SIGTERM, self._signal_handler)

    def _signal_handler(self, signum, frame):
        self._running = False

    # level 1 -- the chunk interface (vocoder_server.py:67-71)
    def _inference_chunk(self, padded):
        return self.voc.infer_chunks(padded)[0]

    # level 2 -- all windows of the request in batched launches, stitched on the GPU;
    # equals the reference's loop over _inference_chunk bit for bit (tests/test_gpu_parity.py)
    def synthesize(self, codes_array):
        return self.voc.synthesize(codes_array)

    def synthesize_pcm16(self, codes_array):
        return self.voc.synthesize_pcm16(codes_array)

    def handle_connection(self, conn: socket.socket) -> None:
        """One request/reply exchange; any failure closes without a reply."""
        try:
            header = _recv_exact(conn, 4)
            if len(header) < 4:
                return
            (n_tokens,) = struct.unpack("<i", header)
            if n_tokens <= 0 or n_tokens > MAX_REQUEST_TOKENS:
                return
            want = n_tokens * CODEBOOKS * 8
            body = _recv_exact(conn, want)
            if len(body) < want:
                return
            codes = np.frombuffer(body, dtype="<i8").reshape(n_tokens, CODEBOOKS)
            t0 = time.time()
            pcm = self.synthesize_pcm16(codes)
            dt = time.time() - t0
            print(f"  Vocoder: {n_tokens} tokens -> {len(pcm)} samples ({dt:.2f}s)")
            conn.sendall(struct.pack("<i", len(pcm)))
            conn.sendall(pcm.astype("<i2", copy=False).tobytes())
        except Exception as e:  # same policy as the reference (:180-181)
            print(f"  Vocoder Error: {e}")
        finally:
            conn.close()

    def serve(self):
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)
        sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        sock.bind(self.socket_path)
        sock.listen(1)
        sock.settimeout(1.0)
        os.chmod(self.socket_path, 0o666)
        print(f"\nVocoder Server listening on {self.socket_path}")
        try:
            while self._running:
                try:
                    conn, _ = sock.accept()
                except socket.timeout:
                    continue
                self.handle_connection(conn)
        finally:
            sock.close()
            if os.path.exists(self.socket_path):
                os.unlink(self.socket_path)
            self.voc.close()
            print("Vocoder Server stopped.")


def main():
    parser = argparse.ArgumentParser(description="Qwen3-TTS Vocoder Server (B200 backend)")
    parser.add_argument("--model", required=True, help="Vocoder model (.b200voc)")
    parser.add_argument("--socket", default="/tmp/qwen3_voc.sock")
    parser.add_argument("--device", type=int, default=0)
    parser.add_argument("--wave", type=int, default=32, help="windows resident in HBM at once")
    args = parser.parse_args()
    server = VocoderServer(model_path=args.model, socket_path=args.socket, device=args.device,
                           wave=args.wave)
    server.serve()


if __name__ == "__main__":
    main()
