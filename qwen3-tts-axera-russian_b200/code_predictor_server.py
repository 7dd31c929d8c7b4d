#!/usr/bin/env python3
"""Drop-in code-predictor server: the reference's unix-socket protocol in front of the B200 backend.

Mirrors ``/root/reference/dual_npu/code_predictor_server.py``: class ``CodePredictorServer(model_dir, embeddings_dir,
socket_path, temperature, top_k, n_threads, batch_prefill)`` with ``_ort_step``, ``_sample``, ``predict`` and ``serve``,
the same files (``code_predictor_weights.npz``, ``codec_embedding.npy``) and the same CLI, so the launcher and the talker
client (``dual_npu/talker_client.py``) work unchanged.  The ONNX file is not read: the decode step runs from the arrays
of the ``.npz`` (scripts/export_code_predictor_weights.py:50-70 exports the layer weights beside the heads).

Wire protocol (``code_predictor_server.py:8-12``), unchanged:
  client -> server : hidden_size float32 LE (4096 bytes), then int32 LE code_0
  server -> client : 15 int32 LE codes (60 bytes)
  a short read, or any error: the connection is closed without a reply (:160-186)

Two samplers:
  ``sampler="host"``   -- level 1: every decode step on the GPU, logits back to the host, the reference's own
                          ``_sample`` with NumPy's global generator (same random stream as the reference);
  ``sampler="device"`` -- level 2 (default): the whole frame is one CUDA-graph launch, top-k sampling on the device
                          (same distribution, different stream; ``top_k`` <= 64).
"""
from __future__ import annotations

import argparse
import os
import signal
import socket
import struct
import time

import numpy as np

HIDDEN_SIZE = 1024


class CodePredictorServer:
    def __init__(self, model_dir, embeddings_dir, socket_path="/tmp/qwen3_cp.sock", temperature=0.1, top_k=50,
                 n_threads=1, batch_prefill=False, sampler="device", device=0, seed=0, install_signal_handlers=True,
                 weights=None, codec_embedding=None):
        from .code_predictor import CodePredictor, CPConfig
        self.socket_path = socket_path
        self.temperature = temperature
        self.top_k = top_k
        self.batch_prefill = batch_prefill
        self.sampler = sampler
        if sampler not in ("host", "device"):
            raise ValueError("sampler must be 'host' or 'device'")
        del n_threads                     # an ONNX Runtime knob; accepted for CLI compatibility, nothing to thread here
        if weights is None:
            print("Loading code predictor weights...")
            weights = dict(np.load(os.path.join(model_dir, "code_predictor_weights.npz")))
        if codec_embedding is None:
            codec_embedding = np.load(os.path.join(embeddings_dir, "codec_embedding.npy"))
        self.codec_embedding = np.asarray(codec_embedding, dtype=np.float32)
        cfg = CPConfig.from_weights(weights)
        self.cp = CodePredictor(cfg, weights, device=device)
        self.num_groups = cfg.groups
        self.num_layers = cfg.layers
        self.head_dim = cfg.head_dim
        self.num_kv_heads = cfg.kv_heads
        self.hidden_size = cfg.hidden
        self.codec_embeddings = [np.asarray(weights[f"codec_emb_{i}"], dtype=np.float32) for i in range(cfg.groups)]
        self._frame = int(seed)
        print(f"  B200 code predictor (device {device}): {self.num_layers} layers, head_dim={self.head_dim}, sampler={sampler}")
        self._running = True
        if install_signal_handlers:
            signal.signal(signal.SIGINT, self._signal_handler)
            signal.signal(signal.SIGTERM, self._signal_handler)

    def _signal_handler(self, signum, frame):
        self._running = False

    # level 1 -- code_predictor_server.py:77-85; the caches stay on the device, so `kv_caches` is only a marker:
    # None (or empty caches) starts a frame
    def _ort_step(self, hidden, position, kv_caches=None):
        if position == 0:
            self.cp.reset()
        out = self.cp.step(hidden, position)
        return out[None], {"device_cache_len": self.cp.cache_len}

    # code_predictor_server.py:87-92, verbatim arithmetic (host sampler)
    def _sample(self, logits):
        top_indices = np.argpartition(logits, -self.top_k)[-self.top_k:]
        top_logits = logits[top_indices]
        probs = np.exp((top_logits - top_logits.max()) / max(self.temperature, 1e-6))
        probs /= probs.sum()
        return int(top_indices[np.random.choice(len(top_indices), p=probs)])

    def predict(self, hidden_state, code_0):
        """Groups 1-15 of one frame from the talker's hidden state and code_0 (code_predictor_server.py:94-140)."""
        H = self.hidden_size
        code_0_embed = self.codec_embedding[code_0]
        h0 = np.asarray(hidden_state, dtype=np.float32).flatten()[:H]
        h1 = code_0_embed.flatten()[:H]
        if self.sampler == "device":
            self._frame += 1
            return [int(c) for c in self.cp.predict(h0, h1, self.temperature, self.top_k, seed=self._frame)]
        self.cp.reset()
        if self.batch_prefill:
            self.cp.step(np.stack([h0, h1]), 0)
        else:
            self.cp.step(h0[None], 0)
            self.cp.step(h1[None], 1)
        token = self._sample(self.cp.logits(0))
        predicted_tokens = [token]
        for step in range(1, self.num_groups):
            self.cp.step(self.codec_embeddings[step - 1][token][None], step + 1)
            token = self._sample(self.cp.logits(step))
            predicted_tokens.append(token)
        return predicted_tokens

    def _recv_exact(self, conn, n):
        data = b""
        while len(data) < n:
            chunk = conn.recv(n - len(data))
            if not chunk:
                break
            data += chunk
        return data

    def serve(self):
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)
        sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        sock.bind(self.socket_path)
        sock.listen(1)
        sock.settimeout(1.0)
        os.chmod(self.socket_path, 0o666)
        print(f"\nCode Predictor Server listening on {self.socket_path}")
        nbytes = self.hidden_size * 4
        while self._running:
            try:
                conn, _ = sock.accept()
            except socket.timeout:
                continue
            try:
                hidden_data = self._recv_exact(conn, nbytes)
                if len(hidden_data) < nbytes:
                    continue
                code_data = self._recv_exact(conn, 4)
                if len(code_data) < 4:
                    continue
                code_0 = struct.unpack("<i", code_data)[0]
                if not 0 <= code_0 < len(self.codec_embedding):
                    raise ValueError(f"code_0 {code_0} outside the codec embedding table")
                codes = self.predict(np.frombuffer(hidden_data, dtype=np.float32), code_0)
                conn.sendall(np.array(codes[:self.num_groups], dtype=np.int32).tobytes())
            except Exception as e:
                print(f"  CP Error: {e}")
            finally:
                conn.close()
        sock.close()
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)
        print("Code Predictor Server stopped.")


def main():
    parser = argparse.ArgumentParser(description="Qwen3-TTS Code Predictor Server (B200 backend)")
    parser.add_argument("--model_dir", required=True)
    parser.add_argument("--embeddings_dir", required=True)
    parser.add_argument("--socket", default="/tmp/qwen3_cp.sock")
    parser.add_argument("--temperature", type=float, default=0.1)
    parser.add_argument("--top_k", type=int, default=50)
    parser.add_argument("--threads", type=int, default=3)
    parser.add_argument("--batch_prefill", action="store_true")
    parser.add_argument("--sampler", choices=("device", "host"), default="device")
    parser.add_argument("--device", type=int, default=0)
    args = parser.parse_args()
    CodePredictorServer(model_dir=args.model_dir, embeddings_dir=args.embeddings_dir, socket_path=args.socket,
                        temperature=args.temperature, top_k=args.top_k, n_threads=args.threads,
                        batch_prefill=args.batch_prefill, sampler=args.sampler, device=args.device).serve()


if __name__ == "__main__":
    import importlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    importlib.import_module("qwen3-tts-axera-russian_b200.code_predictor_server").main()
