/*
 * voc_b200.h -- C ABI of the B200-native Qwen3-TTS 12 Hz codec vocoder backend.
 *
 * This is the drop-in boundary for ONE path of MasterVVK/qwen3-tts-axera-russian: the model
 * call and chunk loop of dual_npu/vocoder_server.py.  Conventions follow the reference's own
 * FFI precedent (dual_npu/llama_wrapper.c:1-6 and dual_npu/llama_cpp_bindings.py:41-81):
 * scalar / pointer arguments only, no structs by value, opaque handles, caller-allocated
 * outputs, `int` status (0 = ok, negative = error), create/destroy pairs, loadable with
 * ctypes.CDLL.  No torch types appear in any signature.
 *
 * All "host" entry points take ordinary (pageable or pinned) host pointers and perform the
 * H2D / D2H copies themselves; the "_dev" twins take device pointers on the handle's device
 * and a cudaStream_t passed as void* (NULL = the handle's own stream) and do not synchronise.
 *
 * There is no CPU fallback: every entry point that computes fails with VOC_E_CUDA when no
 * sm_100 device is usable.
 */
#ifndef VOC_B200_H
#define VOC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define VOC_OK          0
#define VOC_E_INVALID  (-1)  /* bad argument, or a code outside [0, codebook_size): ORT's Gather
                                would throw there (SURVEY 8b "Errors")                          */
#define VOC_E_CUDA     (-2)  /* CUDA runtime / driver failure, or no sm_100 device              */
#define VOC_E_STATE    (-3)  /* wrong call order: tensor missing, not finalized ...             */
#define VOC_E_NOMEM    (-4)

/* ABI version of this header (bumped on any signature change). */
int voc_abi_version(void);

/* ---- lifetime ------------------------------------------------------------------------
 * Replaces: ort.InferenceSession(model_path, ...)       dual_npu/vocoder_server.py:39-44
 *   cfg_json  : architecture JSON (VocoderConfig.to_json()); NULL or "" = default architecture
 *   device    : CUDA device ordinal
 *   wave      : how many 64-frame windows are resident in HBM at once (activations for
 *               `wave` windows are pre-allocated; larger batches are processed in waves)
 * Returns NULL on failure (message on stderr, like the reference's wrappers).            */
void* voc_create(const char* cfg_json, int device, int wave);
void  voc_destroy(void* h);

/* voc_create + voc_set_tensor for every tensor + voc_finalize from a .b200voc model file
 * (weights.py:save_model; safetensors byte layout with the architecture JSON in the metadata).
 * Replaces: ort.InferenceSession(model_path) for a native caller      vocoder_server.py:39-44 */
void* voc_create_from_file(const char* path, int device, int wave);

/* Upload one FP32 tensor in torch layout (names and shapes: weights.py:weight_shapes).
 * Replaces the weights baked into vocoder_traced_64.onnx
 * (scripts/export_vocoder_traced.py:74-99).                                               */
int voc_set_tensor(void* h, const char* name, const float* data, long long n_elem);
/* Build the kernel-ready weight layouts.  Must be called once after all voc_set_tensor.   */
int voc_finalize(void* h);

/* ---- shape queries -------------------------------------------------------------------
 * voc_max_tokens    replaces  sess.get_inputs()[0].shape[1]   vocoder_server.py:45-46
 * voc_chunk_samples = L, the float samples the graph emits per window
 * voc_out_samples   = len(VocoderServer.synthesize(codes[n])) incl. the short-last-window
 *                     duplication quirk                        vocoder_server.py:73-121     */
int       voc_max_tokens(void* h);
long long voc_chunk_samples(void* h);
long long voc_out_samples(void* h, int n_tokens);
int       voc_num_windows(void* h, int n_tokens);

/* ---- level 1: the chunk interface ----------------------------------------------------
 * Replaces: sess.run(None, {'audio_codes': padded})[0]        vocoder_server.py:67-71
 *   codes : int64 [B][max_tokens][16], C-contiguous   (graph input `audio_codes`,
 *           scripts/export_vocoder_traced.py:95)
 *   out   : float32 [B][voc_chunk_samples()]          (graph output `audio_values`)       */
int voc_infer_chunks(void* h, const long long* codes, int B, float* out);
int voc_infer_chunks_dev(void* h, const long long* d_codes, int B, float* d_out, void* stream);

/* ---- level 2: whole request, all windows in batched launches -------------------------
 * Replaces: VocoderServer.synthesize(codes_array)             vocoder_server.py:73-121
 *           + np.clip(audio*32767,...).astype(int16)          vocoder_server.py:175
 *   codes : int64 [n_tokens][16]
 *   out   : caller-allocated, capacity `cap` elements; *n_out receives the element count
 * Results equal level 1 + the reference's Python stitching bit for bit.                   */
int voc_synthesize_f32(void* h, const long long* codes, int n_tokens, float* out,
                       long long cap, long long* n_out);
int voc_synthesize_pcm16(void* h, const long long* codes, int n_tokens, short* out,
                         long long cap, long long* n_out);
/* Device twin.  Either of d_out_f32 / d_out_i16 may be NULL.                              */
int voc_synthesize_dev(void* h, const long long* d_codes, int n_tokens, float* d_out_f32,
                       short* d_out_i16, long long cap, long long* n_out, void* stream);

/* ---- level 2, batched: many requests in one call -------------------------
 * What a server does with the streaming client's concurrent 64-token requests
 * (dual_npu/tts_client.py:188-197) or a corpus of utterances: the windows of all requests share
 * batched launches and one stitch.  codes = the requests' [n_tokens[u]][16] arrays concatenated;
 * request u's PCM is out[out_offsets[u] .. out_offsets[u+1]) (out_offsets has n_requests + 1
 * entries), each bit-identical to voc_synthesize_pcm16 on that request alone.                */
int voc_synthesize_batch_pcm16(void* h, const long long* codes, const int* n_tokens, int n_requests,
                               short* out, long long cap, long long* out_offsets);

/* ---- carried-state decode (OPT-IN: the output differs from the reference's) --------
 * The reference decodes a long request as 64-frame windows with a stride of 48 and crossfades the 16
 * recomputed frames (dual_npu/vocoder_server.py:83-119): 25 % of the frames are computed twice and the
 * window edges differ from an un-chunked decode.  These entry points decode ONE sequence incrementally
 * instead: every causal layer keeps the rows of left context it needs ((k-1)*dilation input rows per
 * convolution, one per transposed convolution, sliding_window-1 rows of K/V for the attention) from one
 * call to the next, so that the concatenated output of successive calls equals the decoder run once on
 * the whole sequence -- each call emits exactly n_tokens * 1920 samples, nothing is recomputed, there is
 * no crossfade and no short-last-window duplication.  Needs transconv_trim = "right" (the
 * causal, length-preserving trim); VOC_E_STATE otherwise.  voc_stream_reset starts a new sequence.
 * Host pointers; any n_tokens >= 1 per call, at most 10240 frames per sequence.                   */
int       voc_stream_reset(void* h);
long long voc_stream_position(void* h);      /* frames decoded since the last reset */
int       voc_stream_decode_f32(void* h, const long long* codes, int n_tokens, float* out, long long cap,
                                long long* n_out);
int       voc_stream_decode_pcm16(void* h, const long long* codes, int n_tokens, short* out, long long cap,
                                  long long* n_out);

/* The _dev twins do not synchronise, so an out-of-range code cannot be reported by their
 * return value.  voc_check_dev synchronises `stream` and returns VOC_E_INVALID if any launch
 * since the last check met a code outside [0, codebook_size) (and clears the flag).        */
int voc_check_dev(void* h, void* stream);

/* ---- multi-GPU: a contiguous range of windows of one long request --------
 * Computes windows [w0, w1) of the n_tokens request (plus window w0-1 when the overlap with
 * it must be blended) and writes the output samples those windows own:
 *   out[0 .. *n_out) == synthesize(codes)[*out_offset .. *out_offset + *n_out)
 * Ranges of consecutive ranks tile the full output exactly; no inter-GPU traffic needed.  */
int voc_synthesize_range_dev(void* h, const long long* d_codes, int n_tokens, int w0, int w1,
                             float* d_out_f32, short* d_out_i16, long long cap,
                             long long* out_offset, long long* n_out, void* stream);

/* ---- host-side planning, usable without a GPU or a handle ---------------------------
 * voc_plan restates the window loop of synthesize() (vocoder_server.py:83-119) for a model
 * that emits `chunk_samples` per `max_tokens`-frame window.  meta (may be NULL) receives
 * 6 ints per window: {dst, a_len, blended, next_blended, prev_a_len, start_frame}.
 * Returns the number of windows (or a negative error); *total = output samples;
 * *pairwise = 1 when every crossfade reads un-blended samples of the previous window.
 * voc_fade_tables fills np.linspace(1,0,ov,dtype=f32) and 1-fade_out (:108-109).          */
int voc_plan(int max_tokens, long long chunk_samples, int n_tokens, int meta_cap, int* meta,
             long long* total, int* pairwise);
int voc_fade_tables(int ov, float* fade_out, float* fade_in);

/* ---- diagnostics ---------------------------------------------------------------------*/
const char* voc_last_error(void* h);       /* NULL handle: error of the last failed voc_create */
long long   voc_kernel_launches(void* h);  /* kernels launched by this handle so far            */
/* Request-path launches of a dense layer that did NOT run on the tcgen05 kernel (possible only with
 * gemm = "auto" on architectures whose channel counts the tensor-core tiles do not take; with
 * gemm = "tc" such a layer is an error, VOC_E_INVALID).  0 on the production architecture: asserted by
 * tests/test_gpu_parity.py and reported by bench.py.  No reference counterpart (ORT has one CPU EP). */
long long   voc_simt_launches(void* h);
/* Options: "gemm" = "auto" | "simt" | "tc": kernel family of the dense layers.  "tc" runs them as tcgen05
 *            tensor-core tiles on split-fp16 operands with FP32 accumulation and FAILS (VOC_E_INVALID) on a
 *            layer shape that kernel does not take; "auto" does the same but lets such a layer (only tiny
 *            test architectures have one) run on the CUDA-core kernel, counted by voc_simt_launches;
 *            "simt" is the all-float32 CUDA-core path (the on-device cross-check).
 *          "tc_flags" = experiment switches (bit 0 no tap reuse, bit 1 / 2 force 64- / 32-wide K
 *            chunks, bit 3 run-time epilogue only, bit 4 no double-length head segments, bit 7 no
 *            cta_group::2 pairs, bit 8 / bit 9 always the widest / the narrowest column tile of a layer's
 *            family -- default: chosen per launch from the number of tiles, bit-identical results either
 *            way --, bits 16.. = MMAs accumulated in the tensor core per round-to-nearest flush, default 24)
 *          "operand_stats" = "0" | "1": see voc_operand_report
 *          "fuse_ru" = "1" | "0": residual units of the blocks with C <= 192 as ONE kernel (conv7 -> Snake ->
 *            conv1 -> + residual, the intermediate operand in shared memory) or as two tap-GEMM launches; the
 *            results are bit-identical
 *          "graphs" = "1" | "0": replay recurring waves of <= "graph_max_wave" (default 4) windows as
 *            CUDA graphs (batch-1 streaming latency is launch-bound)
 *          "front_wave" = windows per launch of the stages before the decoder blocks (codebook sum,
 *            transformer, up-sampling); default min(256, 8 * wave); set before voc_finalize
 *          "profile" = "0" | "1", "debug" = "0" | "1"                                      */
int         voc_set_option(void* h, const char* key, const char* value);
/* The handle's own stream (cudaStream_t as void*), so a caller can bracket the host entry
 * points with CUDA events.                                                                */
void*       voc_stream(void* h);
/* Per-launch CUDA-event profile.  After voc_set_option(h,"profile","1") every kernel launch is
 * bracketed by an event pair; voc_profile_report synchronises, aggregates by layer tag and
 * writes a JSON array [{"tag","calls","ms","flops","bytes"}...] (algorithmic FLOPs / bytes of
 * the launches) into buf, clearing the records.  buf = NULL returns the size needed.        */
long long   voc_profile_report(void* h, char* buf, long long cap);
/* Operand-range diagnostic for checkpoints other than the random-init one.  GEMM operands are stored as two unscaled
 * fp16 planes (hi, lo): absolute error <= 2^-25, but a layer whose activations sit far below 2^-3 keeps fewer
 * significant bits and one beyond 65504 saturates.  After voc_set_option(h,"operand_stats","1") every launch that
 * writes such an operand (dense layers, fused residual units) is followed by a counting pass; voc_operand_report
 * writes a JSON array [{"tag","elements","saturated","hi_subnormal","lo_subnormal","rms","max_abs"}...] per layer tag
 * into buf and clears the counters.  buf = NULL returns the size needed.  Diagnostic mode: no CUDA graphs.          */
long long   voc_operand_report(void* h, char* buf, long long cap);
/* Intermediate activations for parity tests: copies stage `name` ("rvq","pre_conv","xf",
 * "up0","up1","conv_in_s","dec0".."dec3") of the last wave into `out` (channels-last
 * [windows][time][channels]); returns the element count or a negative error.               */
long long   voc_debug_stage(void* h, const char* name, float* out, long long cap);

/* How the tcgen05 kernel would tile a dense layer (N output channels, K input channels, ntaps taps, M time steps per
 * window) for a batch of B windows on `sms` SMs -- host arithmetic only, callable without a GPU.  out5 = {column tile,
 * K chunk, 1 if cta_group::2 pairs, 1 if the 3-pass form runs on a 96-column tile, 1 if the layer's MMA form is the
 * 3-pass one (else concatenated)}.  The form and the pairing never depend on B (they fix the rounding); the column
 * tile may (narrower tiles of one form are bit-identical and fill the machine at small batches).               */
int voc_tc_plan(int N, int K, int ntaps, int M, int B, int sms, int tc_flags, int* out5);

/* The shared-memory plan of the fused residual-unit kernel (C = 96 or 192 channels, kernel size, dilation): out5 = {rows
 * of the halo tile one TMA box loads, halo stages, weight stages, 1 if the intermediate operand T overlays the halo ring
 * (the order of work is then conv7, T, 1x1 per tile; else conv7 of the next tile runs under the hand-over), dynamic
 * shared-memory bytes}.  Host arithmetic only; VOC_E_INVALID for a shape the fused kernel does not take.            */
int voc_ru_plan(int C, int ksz, int dil, int* out5);

/* Kernel-level test / micro-benchmark hook: one "tap GEMM" (the contraction every dense layer of
 * the graph maps onto: causal dilated Conv1d, phase-decomposed ConvTranspose1d, Linear) on caller
 * data, through the FP32 CUDA-core kernel (mode 0), the CUDA-core kernel on split-fp16 operands
 * (mode 1) or the tcgen05 kernel (mode 2).  Layouts: A [B][a_rows][K], W [ntaps*K][N],
 * R / Y / S [B][M][N]; optional pointers may be NULL.  iters > 0 times that many launches with
 * CUDA events (*ms = mean milliseconds).  Returns 0, a negative error, or 1 if mode 2 does not
 * take the shape.  No reference counterpart: ONNX Runtime's operators are not individually
 * callable from dual_npu/vocoder_server.py.                                                   */
int voc_test_tapgemm(int device, int mode, int tc_flags, int B, int a_rows, int K, int N, int M,
                     int a_row0, int ntaps, const int* tap_off, const float* A, const float* W,
                     const float* bias, const float* scale, int act_kind, const float* R,
                     const float* sn_a, const float* sn_invb, float* Y, float* S, int iters,
                     float* ms);

/* Kernel-level test / micro-benchmark hook for one residual unit:
 *   T = Snake2(conv_k(A) + b7) with dilation `dil`;  x' = R + conv1x1(T) + b1;  Y = x';  S = Snake_next(x')
 * fused = 1 runs the single fused kernel (ru_fused.cu), fused = 0 the two tap-GEMM launches it replaces; both on the
 * tcgen05 path and bit-identical.  A, R, Y, S are [B][L][C] float32 (A is split to fp16 hi/lo here), W7 [ksz*C][C],
 * W1 [C][C].  Returns 0, a negative error, or 1 if the fused kernel does not take the shape.  No reference counterpart
 * (ONNX Runtime's operators are not individually callable from dual_npu/vocoder_server.py).                         */
int voc_test_ru(int device, int fused, int tc_flags, int B, int L, int C, int ksz, int dil, const float* A,
                const float* W7, const float* b7, const float* sn2_a, const float* sn2_invb, const float* W1,
                const float* b1, const float* R, const float* snn_a, const float* snn_invb, float* Y, float* S,
                int iters, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* VOC_B200_H */
