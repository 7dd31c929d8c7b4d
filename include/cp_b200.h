/*
 * cp_b200.h -- C ABI of the B200-native code predictor (widening step N4).
 *
 * The second path of MasterVVK/qwen3-tts-axera-russian the vocoder's caller spends its time in: the 5-layer
 * transformer that predicts codec groups 1-15 of every frame, which the reference runs as
 * code_predictor_decode_step.onnx on ONNX Runtime (dual_npu/code_predictor_server.py:55-140).  Same conventions as
 * voc_b200.h: scalars and plain pointers, opaque handle, caller-allocated outputs, int status, no torch types,
 * loadable with ctypes.CDLL; host pointers in, host pointers out.  There is no CPU fallback.
 *
 * The KV cache lives on the device inside the handle (the reference passes past_k / past_v through NumPy on every
 * call, code_predictor_server.py:80-85); cp_reset empties it, which is what predict() does at the start of each frame.
 */
#ifndef CP_B200_H
#define CP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CP_OK          0
#define CP_E_INVALID  (-1)
#define CP_E_CUDA     (-2)
#define CP_E_STATE    (-3)
#define CP_E_NOMEM    (-4)

/* Replaces: ort.InferenceSession(model_path, providers=['CPUExecutionProvider'])   code_predictor_server.py:55-58
 *   cfg_json : {"hidden","layers","heads","kv_heads","head_dim","inter","vocab","groups","rms_eps","rope_theta",
 *               "max_positions"}; NULL or "" = the production shape (1024 / 5 / 16 / 8 / 128 / 3072 / 2048 / 15)
 * Returns NULL on failure (cp_last_error(NULL) holds the reason).                                              */
void* cp_create(const char* cfg_json, int device);
void  cp_destroy(void* h);

/* One float32 array of code_predictor_weights.npz, by the name scripts/export_code_predictor_weights.py:50-70 gives
 * it (layer_{i}_{input_ln,q_proj,k_proj,v_proj,o_proj,q_norm,k_norm,post_ln,gate_proj,up_proj,down_proj},
 * final_norm, codec_emb_{g}, lm_head_{g}); torch [out, in] layout.
 * Replaces: np.load(weights_path) + the weights baked into the ONNX file   code_predictor_server.py:44-53        */
int cp_set_tensor(void* h, const char* name, const float* data, long long n_elem);
int cp_finalize(void* h);

/* ---- level 1: the decode step ------------------------------------------------------------------------------
 * Replaces: CodePredictorServer._ort_step(hidden, position, past_k, past_v)       code_predictor_server.py:77-85
 *   hidden_in  : float32 [S][hidden], S = 1 (decode) or 2 (the reference's batch prefill, :112-120)
 *   position   : position of the first token; must equal cp_cache_len() (the cache only grows)
 *   hidden_out : float32 [S][hidden], the final-normed hidden states (graph output `hidden_out`)
 * The grown past_k / past_v stay on the device.                                                                */
int cp_reset(void* h);
int cp_cache_len(void* h);
int cp_step(void* h, const float* hidden_in, int S, int position, float* hidden_out);

/* Replaces: hidden_out[0, -1, :] @ self.lm_heads[group].T                      code_predictor_server.py:124, :135
 *   logits_out : float32 [vocab], from the last token of the last cp_step / cp_predict                         */
int cp_logits(void* h, int group, float* logits_out);

/* ---- level 2: a whole frame ---------------------------------------------------------------------------------
 * Replaces: CodePredictorServer.predict(hidden_state, code_0_embed, temperature, top_k)   code_predictor_server.py:94-140
 *   hidden_state, code0_embed : float32 [hidden]
 *   codes_out : int32 [groups]
 * One launch from the host: 2 prefill positions, then per group lm_head -> top-k sample -> embedding -> decode step,
 * the sampler on the device (top_k <= 64).  top_k = 1 is greedy decoding and is deterministic; for top_k > 1 the
 * draw comes from a counter-based generator keyed by `seed` and the group index: same distribution as the reference's
 * np.random.choice over the renormalised top-k, a different random stream.                                     */
int cp_predict(void* h, const float* hidden_state, const float* code0_embed, float temperature, int top_k,
               unsigned long long seed, int* codes_out);

/* No counterpart in the reference (it serves one stream): the frames of B <= cp_max_batch() independent streams in one
 * launch.  Every kernel carries B input vectors, so the weights -- all of this path's memory traffic -- are streamed once
 * for all B; each stream has its own KV cache and its own seed.  Row b of every array belongs to stream b; results equal
 * B calls of cp_predict.
 *   hidden_states, code0_embeds : float32 [B][hidden];  seeds : [B];  codes_out : int32 [B][groups]            */
int cp_max_batch(void* h);
int cp_predict_batch(void* h, int B, const float* hidden_states, const float* code0_embeds, float temperature, int top_k,
                     const unsigned long long* seeds, int* codes_out);

/* cp_predict runs as a CUDA graph of ~430 launches chained by programmatic dependent launch (default), or -- opt-in,
 * when the shape fits the device -- as ONE persistent cooperative kernel (432 phases separated by grid barriers,
 * next-phase weights loaded under each barrier).  Same results, same speed on a B200; the graph leaves
 * SMs to other streams.  cp_set_option(h, "predict", "graph" | "persistent") selects; cp_predict_path reports.  */
int cp_set_option(void* h, const char* key, const char* value);
const char* cp_predict_path(void* h);

int cp_hidden_size(void* h);
int cp_num_groups(void* h);
int cp_vocab_size(void* h);
/* kernels launched (or captured into the frame graph) by this handle so far */
long long cp_launches(void* h);
const char* cp_last_error(void* h);
void* cp_stream(void* h);

#ifdef __cplusplus
}
#endif
#endif
