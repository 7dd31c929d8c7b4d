// Engine + C ABI of the B200 vocoder backend (see include/voc_b200.h for the contract and the
// reference call sites each entry point replaces).
//
// Data layout in HBM: every activation is channels-last FP32 [window][time][channel]; weights
// are re-laid out once at voc_finalize() into the [tap*K + k][n] matrices the tap-GEMM kernels
// consume; the RVQ out-projections are folded into the codebooks.  Three large ping-pong
// buffers (X = residual stream, S = Snake-activated operand, T = intermediate operand) hold
// the decoder blocks of one *wave* of windows; the tiny front end (RVQ .. conv-in) has its own
// pool.  Nothing is allocated on the request path.
#include "voc_common.cuh"
#include "../../include/voc_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <tuple>
#include <vector>

namespace {

// --------------------------------------------------------------------------------------
// minimal flat-JSON reader (numbers, booleans, strings, arrays of numbers)
// --------------------------------------------------------------------------------------
struct JVal { std::vector<double> nums; std::string str; bool is_str = false; };
static bool parse_flat_json(const char* s, std::map<std::string, JVal>& out, std::string& err) {
    const char* p = s;
    auto ws = [&]() { while (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') ++p; };
    auto str = [&](std::string& o) -> bool {
        if (*p != '"') return false;
        ++p; o.clear();
        while (*p && *p != '"') { if (*p == '\\' && p[1]) ++p; o.push_back(*p++); }
        if (*p != '"') return false;
        ++p; return true;
    };
    auto num = [&](double& v) -> bool {
        char* e = nullptr; v = strtod(p, &e);
        if (e == p) return false;
        p = e; return true;
    };
    ws(); if (*p != '{') { err = "config JSON must be an object"; return false; }
    ++p; ws();
    if (*p == '}') return true;
    for (;;) {
        ws(); std::string key;
        if (!str(key)) { err = "bad key in config JSON"; return false; }
        ws(); if (*p != ':') { err = "missing ':' in config JSON"; return false; }
        ++p; ws();
        JVal v;
        if (*p == '"') { v.is_str = true; if (!str(v.str)) { err = "bad string"; return false; } }
        else if (*p == '[') {
            ++p; ws();
            while (*p && *p != ']') { double d; if (!num(d)) { err = "bad array"; return false; }
                v.nums.push_back(d); ws(); if (*p == ',') { ++p; ws(); } }
            if (*p != ']') { err = "bad array"; return false; }
            ++p;
        } else if (!strncmp(p, "true", 4)) { v.nums.push_back(1); p += 4; }
        else if (!strncmp(p, "false", 5)) { v.nums.push_back(0); p += 5; }
        else if (!strncmp(p, "null", 4)) { p += 4; }
        else { double d; if (!num(d)) { err = "bad value for " + key; return false; } v.nums.push_back(d); }
        out[key] = v;
        ws();
        if (*p == ',') { ++p; continue; }
        if (*p == '}') return true;
        err = "bad config JSON near key " + key; return false;
    }
}

// --------------------------------------------------------------------------------------
// architecture (mirror of config.py:VocoderConfig)
// --------------------------------------------------------------------------------------
struct Cfg {
    int codebook_size = 2048, codebook_dim = 256, num_quantizers = 16, num_semantic = 1, rvq_dim = 512;
    int latent_dim = 1024, pre_conv_kernel = 3;
    int pre_transformer = 1, xf_hidden = 512, xf_inter = 1024, xf_layers = 8, xf_heads = 16, xf_head_dim = 64;
    double rope_theta = 10000.0, rms_eps = 1e-5;
    int sliding_window = 72;
    std::vector<int> upsampling_ratios{2, 2};
    int convnext = 1, convnext_mult = 4;
    double ln_eps = 1e-6;
    int decoder_dim = 1536;
    std::vector<int> upsample_rates{8, 5, 4, 3};
    std::vector<int> dilations{1, 3, 9};
    int conv_kernel = 7;
    double snake_eps = 1e-9;
    int trim_both = 1;
    int chunk_frames = 64, sample_rate = 24000;

    int attn_dim() const { return xf_heads * xf_head_dim; }
    int samples_per_frame() const { int p = 1; for (int r : upsampling_ratios) p *= r; for (int r : upsample_rates) p *= r; return p; }
    int tlen(int L, int s) const { return trim_both ? (L - 1) * s : L * s; }
    int chunk_samples_for(int frames) const {
        int t = frames; for (int r : upsampling_ratios) t *= r;
        for (int s : upsample_rates) t = tlen(t, s);
        return t;
    }
    int chunk_samples() const { return chunk_samples_for(chunk_frames); }
    // How many frames of a window must be computed for its first `len` frames' samples to come out exactly as from
    // the full chunk_frames-frame window (the reference pads a short window with code 0 and slices the output,
    // vocoder_server.py:77-81,92-99).  Every layer is causal except the transposed convolutions under trim "both",
    // which look ahead by one input step each: less than two steps of the decoder's input rate in total, i.e.
    // ceil(2 / prod(upsampling_ratios)) frames.  Rounded up to a multiple of 8 frames so that windows bucket.
    int frames_needed(int len) const {
        if (len >= chunk_frames) return chunk_frames;
        int up = 1; for (int r : upsampling_ratios) up *= r;
        const int la = trim_both ? (2 + up - 1) / up : 0;
        int t = std::min(chunk_frames, (len + la + 7) / 8 * 8);
        const long long want = std::min<long long>(chunk_samples(), (long long)len * 1920);
        while (t < chunk_frames && chunk_samples_for(t) < want) t = std::min(chunk_frames, t + 8);
        return t;
    }
};

static bool cfg_from_json(const char* js, Cfg& c, std::string& err) {
    if (!js || !*js) return true;
    std::map<std::string, JVal> m;
    if (!parse_flat_json(js, m, err)) return false;
    auto geti = [&](const char* k, int& v) { auto it = m.find(k); if (it != m.end() && !it->second.nums.empty()) v = (int)it->second.nums[0]; };
    auto getd = [&](const char* k, double& v) { auto it = m.find(k); if (it != m.end() && !it->second.nums.empty()) v = it->second.nums[0]; };
    auto getv = [&](const char* k, std::vector<int>& v) { auto it = m.find(k); if (it != m.end() && !it->second.is_str) { v.clear(); for (double d : it->second.nums) v.push_back((int)d); } };
    geti("codebook_size", c.codebook_size); geti("codebook_dim", c.codebook_dim);
    geti("num_quantizers", c.num_quantizers); geti("num_semantic", c.num_semantic); geti("rvq_dim", c.rvq_dim);
    geti("latent_dim", c.latent_dim); geti("pre_conv_kernel", c.pre_conv_kernel);
    geti("pre_transformer", c.pre_transformer); geti("xf_hidden", c.xf_hidden); geti("xf_inter", c.xf_inter);
    geti("xf_layers", c.xf_layers); geti("xf_heads", c.xf_heads); geti("xf_head_dim", c.xf_head_dim);
    getd("rope_theta", c.rope_theta); getd("rms_eps", c.rms_eps); geti("sliding_window", c.sliding_window);
    getv("upsampling_ratios", c.upsampling_ratios); geti("convnext", c.convnext); geti("convnext_mult", c.convnext_mult);
    getd("ln_eps", c.ln_eps); geti("decoder_dim", c.decoder_dim); getv("upsample_rates", c.upsample_rates);
    getv("dilations", c.dilations); geti("conv_kernel", c.conv_kernel); getd("snake_eps", c.snake_eps);
    geti("chunk_frames", c.chunk_frames); geti("sample_rate", c.sample_rate);
    auto it = m.find("transconv_trim");
    if (it != m.end() && it->second.is_str) {
        if (it->second.str == "both") c.trim_both = 1;
        else if (it->second.str == "right") c.trim_both = 0;
        else { err = "transconv_trim must be 'both' or 'right'"; return false; }
    }
    if (c.num_quantizers != 16) { err = "the chunk interface carries exactly 16 codebooks per frame"; return false; }
    if (c.conv_kernel > VOC_MAX_TAPS || c.pre_conv_kernel > VOC_MAX_TAPS) { err = "kernel size > 8 taps"; return false; }
    if (c.xf_head_dim != 64 && c.xf_head_dim != 32 && c.xf_head_dim != 16) { err = "xf_head_dim must be 16/32/64"; return false; }
    const int dd = c.decoder_dim >> c.upsample_rates.size();
    if (dd < 4 || (dd % 4) || (c.decoder_dim % (1 << c.upsample_rates.size()))) { err = "decoder_dim too small"; return false; }
    if (c.codebook_dim % 4 || c.rvq_dim % 4 || c.latent_dim % 4 || c.xf_hidden % 4 || c.xf_inter % 4) { err = "channel counts must be multiples of 4"; return false; }
    if (c.chunk_frames < 1 || c.chunk_frames > 4096) { err = "bad chunk_frames"; return false; }
    return true;
}

// --------------------------------------------------------------------------------------
struct DevBuf {
    float* p = nullptr; size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
};

struct GemmW {                 // one dense layer, kernel-ready
    float* W = nullptr;        // CUDA-core form  [ntaps*K][N]
    __half* Wtc = nullptr;     // tensor-core form [2 planes][ntaps][N][K] of W * 2^wexp (hi, lo)
    long long wtc_plane = 0;   // elements per plane
    float wscale = 1.f;        // 2^-wexp
    float* bias = nullptr;     // [N] or null
    int K = 0, N = 0, ntaps = 1;
    int tap_off[VOC_MAX_TAPS] = {0};
};
struct SnakeP { float* a = nullptr; float* invb = nullptr; };

static thread_local std::string g_create_error;
static void fade_tables(int ov, float* fo, float* fi);

struct Engine {
    Cfg cfg;
    int device = 0, wave = 1;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;
    long long simt_launches = 0;                // request-path launches that left the tcgen05 family (gemm = "auto")
    std::set<std::string> simt_warned;
    // CUDA graphs of small waves (the batch-1 streaming case is launch-bound: ~120 tiny kernels per window)
    struct WaveGraph { cudaGraphExec_t exec = nullptr; long long launches = 0; int seen = 0; long long last_use = 0; };
    std::map<std::tuple<const void*, const void*, int, int, int, int, int, int>, WaveGraph> graphs;
    bool use_graphs = true;
    int graph_cache_max = 64;  long long graph_clock = 0;
    void drop_graphs() { for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec); graphs.clear(); }
    int front_wave = 0;                        // windows per launch of the stages before the decoder blocks (>= wave)
    int graph_max_wave = 4;                    // waves of at most this many windows are replayed as graphs
    bool finalized = false;
    int gemm_mode = 0;          // 0 auto (= tc), 1 simt (FP32 CUDA cores), 2 tc (tcgen05, split fp16)
    int tc_flags = 0;           // VOC_TC_* experiment switches
    bool fuse_ru = true;        // residual units with C <= 192 as one kernel (ru_fused.cu)
    bool short_windows = true;  // a request's short last window computes only the frames it needs
    bool fold_head = true;      // the output head rides on the last residual unit (ru_fused.cu, HEAD)
    int num_sms = 148;
    bool debug = false;
    bool tc() const { return gemm_mode != 1; }
    std::map<const float*, size_t> cap;              // capacity (elements) of every activation buffer

    std::map<std::string, std::vector<float>> raw;   // host staging until finalize
    std::vector<void*> owned;                         // device allocations (weights)

    // kernel-ready weights
    float* rvq_tables = nullptr;                      // [16][codebook][rvq_dim]
    GemmW pre_conv, xf_in, xf_out, conv_in;
    struct XfLayer { float *ln1, *ln2, *ls_attn, *ls_mlp; GemmW qkv, o, gu, down; };
    std::vector<XfLayer> xf;
    float* xf_norm = nullptr; float* rope_cos = nullptr; float* rope_sin = nullptr;
    struct Up { GemmW convt; float *dw_w, *dw_b, *ln_w, *ln_b, *gamma; GemmW pw1, pw2; };
    std::vector<Up> ups;
    struct RU { GemmW c1, c2; SnakeP s1, s2; };
    struct Block { SnakeP s_in; SnakeP s_ru0_tiled; GemmW convt; std::vector<RU> ru; int cin, cout, stride; };
    std::vector<Block> blocks;
    SnakeP head_snake; float* head_w = nullptr; float head_b = 0.f;

    // activations
    DevBuf front, big[4], chunks, stitch_f32;
    float *f_rvq, *f_pre, *f_h, *f_hn, *f_qkv, *f_att, *f_gu, *f_act, *f_x, *f_ln, *f_mid, *f_x2;
    size_t big_elems = 0;
    int* d_err = nullptr;
    int* d_strm_pos = nullptr;                  // carried-state decode: frames decoded so far, read by the attention kernel
    int* d_meta = nullptr; size_t meta_cap = 0;
    float* d_fade_out = nullptr; float* d_fade_in = nullptr;
    long long* d_codes = nullptr; size_t codes_cap = 0;
    short* d_pcm = nullptr; size_t pcm_cap = 0;
    std::map<std::string, std::pair<std::shared_ptr<DevBuf>, size_t>> dbg;

    // carried-state decode: frames decoded so far and, per convolution / attention layer, the rows of
    // its input that the next segment's left context needs (created on first use, in the order stream_segment walks)
    struct Stream { long long pos = 0; std::vector<std::pair<float*, size_t>> halos; } strm;
    int rope_positions = 0;

    // operand-range statistics (option "operand_stats" = "1"), read by voc_operand_report: one slot of six counters per
    // layer tag, filled by a counting pass after every launch that writes a split-fp16 operand
    bool opstats = false;
    static constexpr int OPSTAT_SLOTS = 96, OPSTAT_WORDS = 6;
    unsigned long long* d_opstats = nullptr;
    std::vector<std::string> opstat_tags;
    std::string opstat_json;
    // per-launch CUDA-event profile (option "profile" = "1"), read by voc_profile_report
    bool profile = false;
    struct ProfRec { const char* tag; double flops, bytes; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    std::string prof_json;
    cudaEvent_t take_event() {
        if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr; cudaEventCreate(&e); return e;
    }

    ~Engine() {
        for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        for (void* p : owned) cudaFree(p);
        for (auto& hl : strm.halos) cudaFree(hl.first);
        if (d_err) cudaFree(d_err);
        if (d_strm_pos) cudaFree(d_strm_pos);
        if (d_opstats) cudaFree(d_opstats);
        if (d_meta) cudaFree(d_meta);
        if (d_fade_out) cudaFree(d_fade_out);
        if (d_fade_in) cudaFree(d_fade_in);
        if (d_codes) cudaFree(d_codes);
        if (d_pcm) cudaFree(d_pcm);
        for (auto& r : prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        for (auto e : ev_pool) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

// RAII bracket: records an event pair around one kernel launch when profiling is on
struct ProfScope {
    Engine* E; cudaStream_t st; cudaEvent_t e1 = nullptr;
    ProfScope(Engine* E_, cudaStream_t st_, const char* tag, double flops, double bytes) : E(E_), st(st_) {
        E->launches++;
        if (!E->profile) return;
        cudaEvent_t e0 = E->take_event(); e1 = E->take_event();
        cudaEventRecord(e0, st);
        E->prof.push_back({tag, flops, bytes, e0, e1});
    }
    ~ProfScope() { if (e1) cudaEventRecord(e1, st); }
};

#define CK(expr)                                                                         \
    do { cudaError_t _e = (expr);                                                        \
         if (_e != cudaSuccess) {                                                        \
             E->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                \
             fprintf(stderr, "voc_b200: %s\n", E->err.c_str());                          \
             return VOC_E_CUDA; } } while (0)

static int fail(Engine* E, int code, const std::string& msg) {
    E->err = msg;
    fprintf(stderr, "voc_b200: %s\n", msg.c_str());
    return code;
}

// operand-range statistics of the tensor a launch has just written (diagnostic mode only; not counted as a launch)
static cudaError_t opstat(Engine* E, cudaStream_t st, const char* tag, const __half* hi, const __half* lo, int B,
                          long long rows, int cols, int ld, long long bstride) {
    if (!E->opstats || !hi || !lo) return cudaSuccess;
    if (!E->d_opstats) {
        const size_t bytes = (size_t)Engine::OPSTAT_SLOTS * Engine::OPSTAT_WORDS * sizeof(unsigned long long);
        cudaError_t e = cudaMalloc(&E->d_opstats, bytes);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(E->d_opstats, 0, bytes, st);
        if (e != cudaSuccess) return e;
    }
    size_t slot = 0;
    while (slot < E->opstat_tags.size() && E->opstat_tags[slot] != tag) ++slot;
    if (slot == E->opstat_tags.size()) {
        if (slot >= (size_t)Engine::OPSTAT_SLOTS) return cudaSuccess;       // more layer tags than slots: not recorded
        E->opstat_tags.push_back(tag);
    }
    return voc_launch_operand_stats(hi, lo, B, rows, cols, ld, bstride, E->d_opstats + slot * Engine::OPSTAT_WORDS, st);
}

static float* upload(Engine* E, const std::vector<float>& v) {
    float* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(float)) != cudaSuccess) return nullptr;
    // stream-ordered with the finalize-time kernels: a plain cudaMemcpy from pageable memory may
    // return before the DMA lands and is not ordered against a non-blocking stream
    if (!v.empty() && cudaMemcpyAsync(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, E->stream) != cudaSuccess) {
        cudaFree(d); return nullptr;
    }
    E->owned.push_back(d);
    return d;
}

// Tensor-core form of a weight matrix given its CUDA-core form t[(tap*K + k)*N + n]:
// two fp16 planes [plane][tap][n][k] of t * 2^wexp, wexp chosen so that max|t| * 2^wexp is in
// [2^11, 2^12): hi is far from overflow and lo = fp16(w - hi) stays a normal number for every
// weight within 2^-14 of the largest.
static bool make_wtc(Engine* E, const std::vector<float>& t, GemmW& g) {
    const size_t n = t.size();
    float mx = 0.f;
    for (float v : t) mx = std::max(mx, std::fabs(v));
    int wexp = 0;
    if (mx > 0.f && std::isfinite(mx)) { int e; std::frexp(mx, &e); wexp = 12 - e; }   // mx = f * 2^e, f in [0.5,1)
    const float up = std::ldexp(1.0f, wexp);
    g.wscale = std::ldexp(1.0f, -wexp);
    g.wtc_plane = (long long)n;
    std::vector<__half> h(2 * n);
    const int K = g.K, N = g.N;
    for (int tap = 0; tap < g.ntaps; ++tap)
        for (int k = 0; k < K; ++k) {
            const float* src = &t[((size_t)tap * K + k) * N];
            for (int nn = 0; nn < N; ++nn) {
                const float w = src[nn] * up;
                const __half hi = __float2half_rn(w);
                const __half lo = __float2half_rn(w - __half2float(hi));
                const size_t o = ((size_t)tap * N + nn) * K + k;
                h[o] = hi; h[n + o] = lo;
            }
        }
    __half* d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(__half)) != cudaSuccess) return false;
    E->owned.push_back(d);
    if (cudaMemcpyAsync(d, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice, E->stream) != cudaSuccess) return false;
    if (cudaStreamSynchronize(E->stream) != cudaSuccess) return false;     // h goes out of scope
    g.Wtc = d;
    return true;
}

static const std::vector<float>* get_raw(Engine* E, const std::string& name, size_t n) {
    auto it = E->raw.find(name);
    if (it == E->raw.end()) { E->err = "missing tensor " + name; return nullptr; }
    if (it->second.size() != n) { E->err = "tensor " + name + " has wrong size"; return nullptr; }
    return &it->second;
}

// Conv1d weight [Co][Ci][K] (+ bias [Co]) -> tap GEMM, causal taps with dilation d
static bool make_conv(Engine* E, const std::string& p, int Co, int Ci, int K, int dil, bool has_bias, GemmW& g) {
    auto* w = get_raw(E, p + ".w", (size_t)Co * Ci * K); if (!w) return false;
    std::vector<float> t((size_t)K * Ci * Co);
    for (int co = 0; co < Co; ++co) for (int ci = 0; ci < Ci; ++ci) for (int j = 0; j < K; ++j)
        t[((size_t)j * Ci + ci) * Co + co] = (*w)[((size_t)co * Ci + ci) * K + j];
    g.W = upload(E, t); g.K = Ci; g.N = Co; g.ntaps = K;
    if (!make_wtc(E, t, g)) return false;
    for (int j = 0; j < K; ++j) g.tap_off[j] = -(K - 1 - j) * dil;
    if (has_bias) { auto* b = get_raw(E, p + ".b", Co); if (!b) return false; g.bias = upload(E, *b); if (!g.bias) return false; }
    return g.W != nullptr;
}
// Linear [out][in] (several stacked along out) -> ntaps = 1
static bool make_linear(Engine* E, const std::vector<std::string>& names, int out_each, int in, const char* bias_name, GemmW& g) {
    const int N = out_each * (int)names.size();
    std::vector<float> t((size_t)in * N);
    for (size_t s = 0; s < names.size(); ++s) {
        auto* w = get_raw(E, names[s], (size_t)out_each * in); if (!w) return false;
        for (int o = 0; o < out_each; ++o) for (int i = 0; i < in; ++i)
            t[(size_t)i * N + s * out_each + o] = (*w)[(size_t)o * in + i];
    }
    g.W = upload(E, t); g.K = in; g.N = N; g.ntaps = 1; g.tap_off[0] = 0;
    if (!make_wtc(E, t, g)) return false;
    if (bias_name) { auto* b = get_raw(E, bias_name, N); if (!b) return false; g.bias = upload(E, *b); if (!g.bias) return false; }
    return g.W != nullptr;
}
// ConvTranspose1d [Ci][Co][k], stride s, k = s (1 tap) or k = 2s (2 taps): N = s*Co
static bool make_convt(Engine* E, const std::string& p, int Ci, int Co, int k, int s, GemmW& g) {
    auto* w = get_raw(E, p + ".w", (size_t)Ci * Co * k); if (!w) return false;
    auto* b = get_raw(E, p + ".b", Co); if (!b) return false;
    const int taps = k / s, N = s * Co;
    if (taps * s != k || taps < 1 || taps > 2) { E->err = "transposed conv needs k = s or k = 2s"; return false; }
    std::vector<float> t((size_t)taps * Ci * N);
    for (int tap = 0; tap < taps; ++tap) for (int ci = 0; ci < Ci; ++ci) for (int ph = 0; ph < s; ++ph) for (int co = 0; co < Co; ++co)
        t[((size_t)tap * Ci + ci) * N + ph * Co + co] = (*w)[((size_t)ci * Co + co) * k + ph + tap * s];
    std::vector<float> bt(N);
    for (int ph = 0; ph < s; ++ph) for (int co = 0; co < Co; ++co) bt[ph * Co + co] = (*b)[co];
    g.W = upload(E, t); g.bias = upload(E, bt); g.K = Ci; g.N = N; g.ntaps = taps;
    if (!make_wtc(E, t, g)) return false;
    g.tap_off[0] = 0; g.tap_off[1] = -1;
    return g.W && g.bias;
}
static bool make_snake(Engine* E, const std::string& p, int C, int tile, SnakeP& sp) {
    auto* al = get_raw(E, p + ".alpha", C); if (!al) return false;
    auto* be = get_raw(E, p + ".beta", C); if (!be) return false;
    std::vector<float> a((size_t)C * tile), ib((size_t)C * tile);
    for (int r = 0; r < tile; ++r) for (int c = 0; c < C; ++c) {
        a[(size_t)r * C + c] = 2.0f * expf((*al)[c]);      // device convention: the doubled argument (voc_common.cuh)
        ib[(size_t)r * C + c] = 1.0f / (expf((*be)[c]) + (float)E->cfg.snake_eps);
    }
    sp.a = upload(E, a); sp.invb = upload(E, ib);
    return sp.a && sp.invb;
}
static float* upload_named(Engine* E, const std::string& name, size_t n) {
    auto* v = get_raw(E, name, n); if (!v) return nullptr;
    return upload(E, *v);
}

static cudaError_t run_gemm(Engine* E, TapGemmParams& p, cudaStream_t st, const char* tag = "gemm.other", bool force_simt = false) {
    const double flops = 2.0 * p.B * (double)p.M * p.N * p.K * p.ntaps;
    // algorithmic bytes: A once, W once, outputs / residual once
    double bytes = 4.0 * ((double)p.B * p.M * p.K + (double)p.ntaps * p.K * p.N);
    if (p.Y) bytes += 4.0 * p.B * (double)p.M * p.N;
    if (p.S) bytes += 4.0 * p.B * (double)p.M * p.N;
    if (p.R) bytes += 4.0 * p.B * (double)p.M * p.N;
    ProfScope ps(E, st, tag, flops, bytes);
    if (force_simt || !E->tc()) return voc_launch_tapgemm_simt(p, st);
    // Tensor-core family.  north_star: "no multi-backend dispatch" -- with gemm = "tc" a shape the tcgen05 kernel
    // does not take is an error, never a silent change of kernel family.  gemm = "auto" (small test
    // architectures whose channel counts are not multiples of 32) may use the FP32 CUDA-core kernel for such a
    // layer; every such launch is counted (voc_simt_launches) and named once on stderr.
    cudaError_t e = cudaErrorNotSupported;
    if (voc_tc_eligible(p)) {
        e = voc_launch_tapgemm_tc(p, st, E->num_sms, E->tc_flags);
        if (e == cudaSuccess) return opstat(E, st, tag, p.S_hi, p.S_lo, p.B, p.M, p.N, p.lds, p.s_bstride);
        if (e != cudaErrorNotSupported) return e;
        (void)cudaGetLastError();
    }
    if (E->gemm_mode == 2) {
        E->err = std::string("layer ") + tag + " (K " + std::to_string(p.K) + ", N " + std::to_string(p.N) + ", taps " +
                 std::to_string(p.ntaps) + ") is not eligible for the tcgen05 kernel and gemm = \"tc\" forbids another kernel family";
        return cudaErrorNotSupported;
    }
    E->simt_launches++;
    if (E->simt_warned.insert(tag).second && getenv("VOC_VERBOSE"))
        fprintf(stderr, "voc_b200: layer %s (K %d, N %d) runs on the FP32 CUDA-core kernel (not tcgen05-eligible)\n", tag, p.K, p.N);
    return voc_launch_tapgemm_simt(p, st);
}

// run_gemm with the engine's error conventions: a refused shape under gemm = "tc" is the caller's mistake
// (VOC_E_INVALID, message names the layer), anything else is a CUDA failure
static int gemm_ck(Engine* E, TapGemmParams& p, cudaStream_t st, const char* tag) {
    const cudaError_t e = run_gemm(E, p, st, tag);
    if (e == cudaSuccess) return VOC_OK;
    if (e == cudaErrorNotSupported && E->gemm_mode == 2) { const std::string m = E->err; return fail(E, VOC_E_INVALID, m); }
    E->err = std::string("tap-GEMM ") + tag + ": " + cudaGetErrorString(e);
    fprintf(stderr, "voc_b200: %s\n", E->err.c_str());
    return VOC_E_CUDA;
}

// An activation buffer in the format of the current mode: float32 on the CUDA-core path, two
// fp16 planes (hi at the base, lo `capacity` halves later -- the same bytes) on the tensor path.
static VocAct act(Engine* E, float* base) {
    VocAct a{nullptr, nullptr, nullptr};
    if (!E->tc()) { a.f = base; return a; }
    a.hi = reinterpret_cast<__half*>(base);
    a.lo = a.hi + E->cap.at(base);
    return a;
}
static VocAct act_f32(float* base) { return VocAct{base, nullptr, nullptr}; }

static TapGemmParams gp(const GemmW& g, VocAct A, long long a_bs, int a_rows, int a_row0, int M, int B) {
    TapGemmParams p;
    memset(&p, 0, sizeof(p));
    p.A = A.f; p.A_hi = A.hi; p.A_lo = A.lo;
    p.a_bstride = a_bs; p.a_rows = a_rows; p.lda = g.K; p.K = g.K; p.a_row0 = a_row0;
    p.ntaps = g.ntaps; for (int i = 0; i < VOC_MAX_TAPS; ++i) p.tap_off[i] = g.tap_off[i];
    p.W = g.W; p.Wtc = g.Wtc; p.wtc_plane = g.wtc_plane; p.wscale = g.wscale;
    p.N = g.N; p.M = M; p.B = B; p.bias = g.bias;
    return p;
}
static void setY(TapGemmParams& p, float* Y) { p.Y = Y; p.ldy = p.N; p.y_bstride = (long long)p.M * p.N; }
// S = snake(v) (sp given) or v itself (sp null), in the operand format of the mode
static void setS(TapGemmParams& p, VocAct S, const SnakeP* sp) {
    p.S = S.f; p.S_hi = S.hi; p.S_lo = S.lo; p.lds = p.N; p.s_bstride = (long long)p.M * p.N;
    p.sn_a = sp ? sp->a : nullptr; p.sn_invb = sp ? sp->invb : nullptr;
}
static void setR(TapGemmParams& p, const float* R) { p.R = R; p.ldr = p.N; p.r_bstride = (long long)p.M * p.N; }

static int dbg_capture(Engine* E, const char* name, VocAct src, size_t n, cudaStream_t st) {
    if (!E->debug) return VOC_OK;
    auto b = std::make_shared<DevBuf>();
    CK(cudaMalloc(&b->p, n * sizeof(float))); b->n = n;
    if (src.f) CK(cudaMemcpyAsync(b->p, src.f, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else CK(voc_launch_unsplit(src.hi, src.lo, b->p, (long long)n, st));
    E->dbg[name] = {b, n};
    return VOC_OK;
}

// --------------------------------------------------------------------------------------
// finalize: build every kernel-ready tensor
// --------------------------------------------------------------------------------------
static int engine_finalize(Engine* E) {
    const Cfg& c = E->cfg;
    CK(cudaSetDevice(E->device));
    E->err.clear();
#define REQ(x) do { if (!(x)) return fail(E, E->err.rfind("missing", 0) == 0 || E->err.rfind("tensor", 0) == 0 ? VOC_E_STATE : VOC_E_CUDA, E->err.empty() ? std::string("finalize failed: " #x) : E->err); } while (0)

    // ---- RVQ: fold the out-projections into the codebooks with the tap-GEMM itself
    {
        const size_t tbl = (size_t)c.codebook_size * c.rvq_dim;
        float* tables = nullptr;
        CK(cudaMalloc(&tables, tbl * c.num_quantizers * sizeof(float)));
        E->owned.push_back(tables);
        GemmW ps, pa;
        REQ(make_linear(E, {"rvq.proj_sem.w"}, c.rvq_dim, c.codebook_dim, nullptr, ps));
        REQ(make_linear(E, {"rvq.proj_ac.w"}, c.rvq_dim, c.codebook_dim, nullptr, pa));
        for (int q = 0; q < c.num_quantizers; ++q) {
            float* cb = upload_named(E, "rvq.codebook." + std::to_string(q), (size_t)c.codebook_size * c.codebook_dim);
            REQ(cb);
            TapGemmParams p = gp(q < c.num_semantic ? ps : pa, act_f32(cb), 0, c.codebook_size, 0, c.codebook_size, 1);
            setY(p, tables + q * tbl);
            // weight preparation, once per handle: an exact FP32 fold on the CUDA cores by design (its operand
            // is the float32 codebook itself), not a request-path launch
            CK(run_gemm(E, p, E->stream, "finalize.fold", true));
        }
        CK(cudaStreamSynchronize(E->stream));
        E->rvq_tables = tables;
    }
    REQ(make_conv(E, "pre_conv", c.latent_dim, c.rvq_dim, c.pre_conv_kernel, 1, true, E->pre_conv));
    if (c.pre_transformer) {
        REQ(make_linear(E, {"xf.in_proj.w"}, c.xf_hidden, c.latent_dim, "xf.in_proj.b", E->xf_in));
        REQ(make_linear(E, {"xf.out_proj.w"}, c.latent_dim, c.xf_hidden, "xf.out_proj.b", E->xf_out));
        E->xf.resize(c.xf_layers);
        for (int l = 0; l < c.xf_layers; ++l) {
            const std::string p = "xf." + std::to_string(l) + ".";
            auto& L = E->xf[l];
            REQ(L.ln1 = upload_named(E, p + "ln1.w", c.xf_hidden));
            REQ(L.ln2 = upload_named(E, p + "ln2.w", c.xf_hidden));
            REQ(L.ls_attn = upload_named(E, p + "ls_attn", c.xf_hidden));
            REQ(L.ls_mlp = upload_named(E, p + "ls_mlp", c.xf_hidden));
            REQ(make_linear(E, {p + "q.w", p + "k.w", p + "v.w"}, c.attn_dim(), c.xf_hidden, nullptr, L.qkv));
            REQ(make_linear(E, {p + "o.w"}, c.xf_hidden, c.attn_dim(), nullptr, L.o));
            REQ(make_linear(E, {p + "gate.w", p + "up.w"}, c.xf_inter, c.xf_hidden, nullptr, L.gu));
            REQ(make_linear(E, {p + "down.w"}, c.xf_hidden, c.xf_inter, nullptr, L.down));
        }
        REQ(E->xf_norm = upload_named(E, "xf.norm.w", c.xf_hidden));
        // rotary table evaluated in float64 then cast to float32 (angles t * theta^(-2d/hd))
        // positions 0 .. 10 239: a window's frames, or the absolute frame index of a carried-state decode (the
        // protocol's longest request is 10 000 frames, vocoder_server.py:149)
        const int T = std::max(c.chunk_frames, 10240), H2 = c.xf_head_dim / 2;
        E->rope_positions = T;
        std::vector<float> cs((size_t)T * H2), sn((size_t)T * H2);
        for (int t = 0; t < T; ++t) for (int d = 0; d < H2; ++d) {
            const double inv = 1.0 / std::pow(c.rope_theta, (2.0 * d) / c.xf_head_dim);
            cs[(size_t)t * H2 + d] = (float)std::cos(t * inv);
            sn[(size_t)t * H2 + d] = (float)std::sin(t * inv);
        }
        REQ(E->rope_cos = upload(E, cs)); REQ(E->rope_sin = upload(E, sn));
    }
    E->ups.resize(c.upsampling_ratios.size());
    for (size_t u = 0; u < c.upsampling_ratios.size(); ++u) {
        const std::string p = "up." + std::to_string(u) + ".";
        auto& U = E->ups[u];
        const int C = c.latent_dim, r = c.upsampling_ratios[u];
        REQ(make_convt(E, p + "convt", C, C, r, r, U.convt));
        if (c.convnext) {
            auto* dw = get_raw(E, p + "dw.w", (size_t)C * c.conv_kernel); REQ(dw);
            std::vector<float> t((size_t)c.conv_kernel * C);
            for (int ch = 0; ch < C; ++ch) for (int j = 0; j < c.conv_kernel; ++j) t[(size_t)j * C + ch] = (*dw)[(size_t)ch * c.conv_kernel + j];
            REQ(U.dw_w = upload(E, t));
            REQ(U.dw_b = upload_named(E, p + "dw.b", C));
            REQ(U.ln_w = upload_named(E, p + "ln.w", C));
            REQ(U.ln_b = upload_named(E, p + "ln.b", C));
            REQ(U.gamma = upload_named(E, p + "gamma", C));
            const std::string b1 = p + "pw1.b", b2 = p + "pw2.b";
            REQ(make_linear(E, {p + "pw1.w"}, c.convnext_mult * C, C, b1.c_str(), U.pw1));
            REQ(make_linear(E, {p + "pw2.w"}, C, c.convnext_mult * C, b2.c_str(), U.pw2));
        }
    }
    REQ(make_conv(E, "dec.conv_in", c.decoder_dim, c.latent_dim, c.conv_kernel, 1, true, E->conv_in));
    E->blocks.resize(c.upsample_rates.size());
    for (size_t b = 0; b < c.upsample_rates.size(); ++b) {
        const std::string p = "dec." + std::to_string(b) + ".";
        auto& Bk = E->blocks[b];
        Bk.cin = c.decoder_dim >> b; Bk.cout = c.decoder_dim >> (b + 1); Bk.stride = c.upsample_rates[b];
        REQ(make_snake(E, p + "snake", Bk.cin, 1, Bk.s_in));
        REQ(make_convt(E, p + "convt", Bk.cin, Bk.cout, 2 * Bk.stride, Bk.stride, Bk.convt));
        REQ(make_snake(E, p + "ru.0.snake1", Bk.cout, Bk.stride, Bk.s_ru0_tiled));
        Bk.ru.resize(c.dilations.size());
        for (size_t j = 0; j < c.dilations.size(); ++j) {
            const std::string r = p + "ru." + std::to_string(j) + ".";
            REQ(make_snake(E, r + "snake1", Bk.cout, 1, Bk.ru[j].s1));
            REQ(make_snake(E, r + "snake2", Bk.cout, 1, Bk.ru[j].s2));
            REQ(make_conv(E, r + "conv1", Bk.cout, Bk.cout, c.conv_kernel, c.dilations[j], true, Bk.ru[j].c1));
            REQ(make_conv(E, r + "conv2", Bk.cout, Bk.cout, 1, 1, true, Bk.ru[j].c2));
        }
    }
    const int ch = c.decoder_dim >> c.upsample_rates.size();
    REQ(make_snake(E, "head.snake", ch, 1, E->head_snake));
    {
        auto* w = get_raw(E, "head.conv.w", (size_t)ch * c.conv_kernel); REQ(w);
        std::vector<float> t((size_t)c.conv_kernel * ch);
        for (int cc = 0; cc < ch; ++cc) for (int j = 0; j < c.conv_kernel; ++j) t[(size_t)j * ch + cc] = (*w)[(size_t)cc * c.conv_kernel + j];
        REQ(E->head_w = upload(E, t));
        auto* b = get_raw(E, "head.conv.b", 1); REQ(b);
        E->head_b = (*b)[0];
    }
#undef REQ
    E->raw.clear();

    // ---- activation pools
    // Two wave sizes: the decoder blocks (gigabytes of activations per launch) run `wave` windows at a time, the
    // stages before them (codebook sum, transformer, up-sampling: 10 MB per window, ~100 small launches) run
    // `front_wave` windows at a time so that their GEMMs fill the machine.
    if (E->front_wave < E->wave) E->front_wave = std::max(E->wave, std::min(256, 8 * E->wave));
    const int W = E->front_wave, T = c.chunk_frames;
    int Tup = T; for (int r : c.upsampling_ratios) Tup *= r;
    size_t off = 0;
    auto carve = [&](size_t n) { size_t o = off; off += (n + 63) / 64 * 64; return o; };
    const size_t o_rvq = carve((size_t)W * T * c.rvq_dim), o_pre = carve((size_t)W * T * c.latent_dim),
                 o_h = carve((size_t)W * T * c.xf_hidden), o_hn = carve((size_t)W * T * c.xf_hidden),
                 o_qkv = carve((size_t)W * T * 3 * c.attn_dim()), o_att = carve((size_t)W * T * c.attn_dim()),
                 o_gu = carve((size_t)W * T * 2 * c.xf_inter), o_act = carve((size_t)W * T * c.xf_inter),
                 o_x = carve((size_t)W * Tup * c.latent_dim), o_x2 = carve((size_t)W * Tup * c.latent_dim),
                 o_ln = carve((size_t)W * Tup * c.latent_dim),
                 o_mid = carve((size_t)W * Tup * c.latent_dim * c.convnext_mult);
    CK(cudaMalloc(&E->front.p, off * sizeof(float))); E->front.n = off;
    float* f = E->front.p;
    E->f_rvq = f + o_rvq; E->f_pre = f + o_pre; E->f_h = f + o_h; E->f_hn = f + o_hn; E->f_qkv = f + o_qkv;
    E->f_att = f + o_att; E->f_gu = f + o_gu; E->f_act = f + o_act; E->f_x = f + o_x; E->f_x2 = f + o_x2;
    E->f_ln = f + o_ln; E->f_mid = f + o_mid;
    E->cap[E->f_rvq] = (size_t)W * T * c.rvq_dim;      E->cap[E->f_pre] = (size_t)W * T * c.latent_dim;
    E->cap[E->f_hn] = (size_t)W * T * c.xf_hidden;     E->cap[E->f_att] = (size_t)W * T * c.attn_dim();
    E->cap[E->f_act] = (size_t)W * T * c.xf_inter;     E->cap[E->f_x] = (size_t)W * Tup * c.latent_dim;
    E->cap[E->f_x2] = (size_t)W * Tup * c.latent_dim;  E->cap[E->f_ln] = (size_t)W * Tup * c.latent_dim;
    E->cap[E->f_mid] = (size_t)W * Tup * c.latent_dim * c.convnext_mult;
    size_t mx = (size_t)Tup * c.decoder_dim;
    int L = Tup;
    for (size_t b = 0; b < c.upsample_rates.size(); ++b) {
        L = c.tlen(L, c.upsample_rates[b]);
        if (L < 1) return fail(E, VOC_E_INVALID, "chunk_frames too small for transconv_trim='both'");
        mx = std::max(mx, (size_t)L * (c.decoder_dim >> (b + 1)));
    }
    E->big_elems = mx;
    for (int i = 0; i < 4; ++i) {
        CK(cudaMalloc(&E->big[i].p, mx * E->wave * sizeof(float))); E->big[i].n = mx * E->wave;
        E->cap[E->big[i].p] = mx * E->wave;
    }
    CK(cudaMalloc(&E->d_err, sizeof(int)));

    // ---- crossfade tables (fade_tables below restates numpy's linspace)
    {
        const int ov = 16 * c.samples_per_frame();
        std::vector<float> fo(ov), fi(ov);
        fade_tables(ov, fo.data(), fi.data());
        CK(cudaMalloc(&E->d_fade_out, ov * sizeof(float))); CK(cudaMalloc(&E->d_fade_in, ov * sizeof(float)));
        CK(cudaMemcpyAsync(E->d_fade_out, fo.data(), ov * sizeof(float), cudaMemcpyHostToDevice, E->stream));
        CK(cudaMemcpyAsync(E->d_fade_in, fi.data(), ov * sizeof(float), cudaMemcpyHostToDevice, E->stream));
        CK(cudaStreamSynchronize(E->stream));
    }
    CK(cudaMemsetAsync(E->d_err, 0, sizeof(int), E->stream));
    CK(cudaDeviceSynchronize());       // weights are visible to any stream the caller brings
    E->finalized = true;
    return VOC_OK;
}

// --------------------------------------------------------------------------------------
// one front wave: windows [w_begin, w_begin + nw) of a request (nw <= front_wave) through the codebook sum,
// pre-conv, transformer and up-sampling stages -> *x_out [nw][*L_out][latent] in operand format
// --------------------------------------------------------------------------------------
static int run_front(Engine* E, const long long* d_codes, int n_frames, int win_step, int w_begin, int nw,
                     cudaStream_t st, const int* d_wmeta, float** x_out, int* L_out, int T) {
    const Cfg& c = E->cfg;
#define KLAUNCH(tag, flops, bytes, call) do { ProfScope _ps(E, st, tag, flops, bytes); CK(call); } while (0)
#define GEMM(tag, p) do { if (int _r = gemm_ck(E, p, st, tag)) return _r; } while (0)
    // Operand tensors (everything a GEMM reads) are in the mode's operand format -- act(); the
    // residual streams (f_h, the up-sampled x while ConvNeXt needs it, bX) and the tensors the
    // non-GEMM kernels read (f_qkv, f_gu) stay float32.
    // K1: RVQ gather-sum
    KLAUNCH("rvq_gather", 0.0, 4.0 * nw * T * c.rvq_dim * (c.num_quantizers + 1.0),
            voc_launch_rvq_gather(d_wmeta ? d_codes : d_codes + (long long)w_begin * win_step * c.num_quantizers,
                                  n_frames - (d_wmeta ? 0 : w_begin * win_step), T, win_step, nw, c.num_quantizers,
                                  c.codebook_size, E->rvq_tables, c.rvq_dim, act(E, E->f_rvq), E->d_err, st,
                                  d_wmeta ? d_wmeta + 2 * (size_t)w_begin : nullptr));
    if (int r = dbg_capture(E, "rvq", act(E, E->f_rvq), (size_t)nw * T * c.rvq_dim, st)) return r;
    // K2: pre-conv
    {
        TapGemmParams p = gp(E->pre_conv, act(E, E->f_rvq), (long long)T * c.rvq_dim, T, 0, T, nw);
        setS(p, act(E, E->f_pre), nullptr);
        GEMM("pre_conv", p);
    }
    if (int r = dbg_capture(E, "pre_conv", act(E, E->f_pre), (size_t)nw * T * c.latent_dim, st)) return r;
    float* x = E->f_pre;                 // [nw][T][latent], operand format
    if (c.pre_transformer) {
        const int rows = nw * T, H = c.xf_hidden, A = c.attn_dim();
        const double nb = 8.0 * rows * H;
        { TapGemmParams p = gp(E->xf_in, act(E, x), 0, rows, 0, rows, 1); setY(p, E->f_h); GEMM("xf.gemm", p); }
        for (int l = 0; l < c.xf_layers; ++l) {
            auto& Ly = E->xf[l];
            KLAUNCH("xf.norm", 0.0, nb, voc_launch_rmsnorm(E->f_h, Ly.ln1, act(E, E->f_hn), rows, H, (float)c.rms_eps, st));
            { TapGemmParams p = gp(Ly.qkv, act(E, E->f_hn), 0, rows, 0, rows, 1); setY(p, E->f_qkv); GEMM("xf.gemm", p); }
            KLAUNCH("xf.attn", 4.0 * nw * c.xf_heads * (double)T * T * c.xf_head_dim / 2, 16.0 * rows * A,
                    voc_launch_attention(E->f_qkv, act(E, E->f_att), nw, T, c.xf_heads, c.xf_head_dim, E->rope_cos, E->rope_sin, c.sliding_window, st));
            { TapGemmParams p = gp(Ly.o, act(E, E->f_att), 0, rows, 0, rows, 1); p.scale = Ly.ls_attn; setR(p, E->f_h); setY(p, E->f_h); GEMM("xf.gemm", p); }
            KLAUNCH("xf.norm", 0.0, nb, voc_launch_rmsnorm(E->f_h, Ly.ln2, act(E, E->f_hn), rows, H, (float)c.rms_eps, st));
            { TapGemmParams p = gp(Ly.gu, act(E, E->f_hn), 0, rows, 0, rows, 1); setY(p, E->f_gu); GEMM("xf.gemm", p); }
            KLAUNCH("xf.swiglu", 0.0, 12.0 * rows * c.xf_inter, voc_launch_swiglu(E->f_gu, act(E, E->f_act), rows, c.xf_inter, st));
            { TapGemmParams p = gp(Ly.down, act(E, E->f_act), 0, rows, 0, rows, 1); p.scale = Ly.ls_mlp; setR(p, E->f_h); setY(p, E->f_h); GEMM("xf.gemm", p); }
        }
        KLAUNCH("xf.norm", 0.0, nb, voc_launch_rmsnorm(E->f_h, E->xf_norm, act(E, E->f_hn), rows, H, (float)c.rms_eps, st));
        { TapGemmParams p = gp(E->xf_out, act(E, E->f_hn), 0, rows, 0, rows, 1); setS(p, act(E, E->f_x), nullptr); GEMM("xf.gemm", p); }
        x = E->f_x;
        if (int r = dbg_capture(E, "xf", act(E, x), (size_t)nw * T * c.latent_dim, st)) return r;
    }
    // K3: upsample stages (k = s transposed conv = GEMM + interleave, then ConvNeXt)
    int L = T;
    for (size_t u = 0; u < E->ups.size(); ++u) {
        auto& U = E->ups[u];
        const int C = c.latent_dim, r = c.upsampling_ratios[u];
        float* d1 = (x == E->f_x) ? E->f_x2 : E->f_x;        // transposed-conv output
        float* d2 = (d1 == E->f_x) ? E->f_x2 : E->f_x;       // ConvNeXt output (x is dead by then)
        {
            TapGemmParams p = gp(U.convt, act(E, x), (long long)L * C, L, 0, L, nw);
            if (c.convnext) setY(p, d1); else setS(p, act(E, d1), nullptr);
            GEMM("up.convt", p);
        }
        L *= r;
        x = d1;                           // [nw][L][C]
        if (c.convnext) {
            KLAUNCH("up.dwconv_ln", 2.0 * nw * L * C * c.conv_kernel, 8.0 * nw * L * C,
                    voc_launch_dwconv_ln(d1, U.dw_w, U.dw_b, U.ln_w, U.ln_b, act(E, E->f_ln), nw, L, C, c.conv_kernel, (float)c.ln_eps, st));
            const int rows = nw * L;
            { TapGemmParams p = gp(U.pw1, act(E, E->f_ln), 0, rows, 0, rows, 1); p.act = VOC_ACT_GELU; setS(p, act(E, E->f_mid), nullptr); GEMM("up.pw1", p); }
            { TapGemmParams p = gp(U.pw2, act(E, E->f_mid), 0, rows, 0, rows, 1); p.scale = U.gamma; setR(p, d1); setS(p, act(E, d2), nullptr); GEMM("up.pw2", p); }
            x = d2;
        }
        char nm[32]; snprintf(nm, sizeof nm, "up%d", (int)u);
        if (int r2 = dbg_capture(E, nm, act(E, x), (size_t)nw * L * C, st)) return r2;
    }
    *x_out = x; *L_out = L;
    return VOC_OK;
}

// an operand tensor `elems` elements into a pool buffer (both planes move together)
static VocAct act_at(Engine* E, float* base, size_t elems) {
    VocAct a = act(E, base);
    if (a.f) a.f += elems;
    if (a.hi) { a.hi += elems; a.lo += elems; }
    return a;
}

// --------------------------------------------------------------------------------------
// one wave of the decoder: windows [x_win0, x_win0 + nw) of a front wave's output (nw <= wave) through conv-in,
// the decoder blocks and the head -> chunk_out[nw][Lc]
// --------------------------------------------------------------------------------------
static int run_back(Engine* E, float* x, int x_win0, int L, int nw, float* chunk_out, long long o_bstride, cudaStream_t st) {
    const Cfg& c = E->cfg;
    // K4: decoder conv-in, emits only Snake_0(conv_in(x)) -- the operand of block 0
    float* bX = E->big[0].p; float* bS = E->big[1].p; float* bT = E->big[2].p; float* bX2 = E->big[3].p;
    {
        TapGemmParams p = gp(E->conv_in, act_at(E, x, (size_t)x_win0 * L * c.latent_dim), (long long)L * c.latent_dim, L, 0, L, nw);
        setS(p, act(E, bS), &E->blocks[0].s_in);
        if (E->debug) setY(p, bX);
        GEMM("conv_in", p);
        if (E->debug) if (int r = dbg_capture(E, "conv_in", act_f32(bX), (size_t)nw * L * c.decoder_dim, st)) return r;
    }
    // K5/K6: decoder blocks
    bool head_folded = false;
    static const char* const T_CONVT[] = {"dec0.convt", "dec1.convt", "dec2.convt", "dec3.convt", "decN.convt"};
    static const char* const T_C7[] = {"dec0.ru.conv7", "dec1.ru.conv7", "dec2.ru.conv7", "dec3.ru.conv7", "decN.ru.conv7"};
    static const char* const T_C1[] = {"dec0.ru.conv1", "dec1.ru.conv1", "dec2.ru.conv1", "dec3.ru.conv1", "decN.ru.conv1"};
    for (size_t b = 0; b < E->blocks.size(); ++b) {
        auto& Bk = E->blocks[b];
        const size_t ti = std::min<size_t>(b, 4);
        const int Lout = c.tlen(L, Bk.stride);
        const int Mrows = Lout / Bk.stride;                 // GEMM rows (one per input step)
        const int row0 = c.trim_both ? 1 : 0;
        {   // Snake'd input (bS, [nw][L][cin]) -> X' (bX) and Snake1_ru0(X') (bT)
            TapGemmParams p = gp(Bk.convt, act(E, bS), (long long)L * Bk.cin, L, row0, Mrows, nw);
            setY(p, bX); setS(p, act(E, bT), &Bk.s_ru0_tiled);
            GEMM(T_CONVT[ti], p);
        }
        L = Lout;
        std::swap(bS, bT);                                   // bS = Snake1(X'), bT free
        const int C = Bk.cout;
        for (size_t j = 0; j < Bk.ru.size(); ++j) {
            auto& R = Bk.ru[j];
            const bool last_ru_f = (j + 1 == Bk.ru.size());
            const SnakeP& nxt_f = !last_ru_f ? Bk.ru[j + 1].s1
                                : (b + 1 < E->blocks.size() ? E->blocks[b + 1].s_in : E->head_snake);
            if (E->tc() && E->fuse_ru) {
                // the whole unit in one kernel where a tile spans all channels (C <= 192): conv7 -> Snake2 -> conv1 ->
                // + residual, the Snake2'd operand staying in shared memory (ru_fused.cu)
                RuFusedParams f;
                memset(&f, 0, sizeof f);
                const VocAct Ain = act(E, bS), Sout = act(E, bT);
                f.A_hi = Ain.hi; f.A_lo = Ain.lo; f.L = L; f.C = C; f.B = nw; f.dil = c.dilations[j]; f.ksz = c.conv_kernel;
                f.W7tc = R.c1.Wtc; f.w7_plane = R.c1.wtc_plane; f.w7scale = R.c1.wscale; f.bias7 = R.c1.bias;
                f.sn2_a = R.s2.a; f.sn2_invb = R.s2.invb;
                f.W1tc = R.c2.Wtc; f.w1_plane = R.c2.wtc_plane; f.w1scale = R.c2.wscale; f.bias1 = R.c2.bias;
                f.R = bX; f.Y = (!last_ru_f || E->debug) ? bX2 : nullptr;
                f.S_hi = Sout.hi; f.S_lo = Sout.lo; f.snn_a = nxt_f.a; f.snn_invb = nxt_f.invb;
                // the very last unit can take the output head with it: per-tap partial sums (into bT, which S would have
                // used) instead of the 4 B per element of S that the head kernel would read straight back
                const bool very_last = last_ru_f && b + 1 == E->blocks.size();
                if (very_last && E->fold_head && !E->debug && c.conv_kernel == 7) {
                    RuFusedParams g = f;
                    g.S_hi = nullptr; g.S_lo = nullptr; g.head_w = E->head_w; g.head_part = bT; g.head_taps = c.conv_kernel;
                    if (voc_ru_fused_eligible(g)) { f = g; head_folded = true; }
                }
                if (voc_ru_fused_eligible(f)) {
                    static const char* const T_FU[] = {"dec0.ru.fused", "dec1.ru.fused", "dec2.ru.fused", "dec3.ru.fused", "decN.ru.fused"};
                    const double el = (double)nw * L * C;
                    ProfScope ps(E, st, T_FU[ti], 2.0 * el * C * (c.conv_kernel + 1), 4.0 * el * (f.Y ? 4.0 : f.head_w ? 2.0 : 3.0));
                    CK(voc_launch_ru_fused(f, st, E->num_sms, E->tc_flags));
                    CK(opstat(E, st, T_FU[ti], f.S_hi, f.S_lo, f.B, f.L, f.C, f.C, (long long)f.L * f.C));
                    std::swap(bS, bT);
                    if (f.Y) std::swap(bX, bX2);
                    continue;
                }
            }
            {   // conv k7 dilated on Snake1(x) -> Snake2(.) only
                TapGemmParams p = gp(R.c1, act(E, bS), (long long)L * C, L, 0, L, nw);
                setS(p, act(E, bT), &R.s2);
                GEMM(T_C7[ti], p);
            }
            {   // conv k1 + residual -> x and the next consumer's Snake.  The residual stream ping-pongs between
                // two buffers: updating it in place (a lane's float32 stores landing in the 128-byte lines its
                // next residual loads are about to read) ran the fast epilogue 3.4x slower at C = 192.
                TapGemmParams p = gp(R.c2, act(E, bT), (long long)L * C, L, 0, L, nw);
                setR(p, bX);
                const bool last_ru = (j + 1 == Bk.ru.size());
                const SnakeP& nxt = !last_ru ? Bk.ru[j + 1].s1
                                  : (b + 1 < E->blocks.size() ? E->blocks[b + 1].s_in : E->head_snake);
                setS(p, act(E, bS), &nxt);
                const bool writes_y = !last_ru || E->debug;   // the residual stream ends with the block
                if (writes_y) setY(p, bX2);
                GEMM(T_C1[ti], p);
                if (writes_y) std::swap(bX, bX2);
            }
        }
        char nm[32]; snprintf(nm, sizeof nm, "dec%d", (int)b);
        if (E->debug) if (int r = dbg_capture(E, nm, act_f32(bX), (size_t)nw * L * C, st)) return r;
    }
    // K7: head
    const int ch = c.decoder_dim >> c.upsample_rates.size();
    if (head_folded) {
        // (after the swap the partial sums are in bS)
        KLAUNCH("head", 2.0 * nw * L * 2 * c.conv_kernel, 4.0 * nw * L * (2.0 * c.conv_kernel + 1.0),
                voc_launch_head_finish(bS, 2, c.conv_kernel, L, E->head_b, chunk_out, o_bstride, nw, st));
    } else {
        KLAUNCH("head", 2.0 * nw * L * ch * c.conv_kernel, 4.0 * nw * L * (ch + 1.0),
                voc_launch_head(act(E, bS), (long long)L * ch, L, ch, c.conv_kernel, E->head_w, E->head_b, chunk_out, o_bstride, nw, st));
    }
#undef KLAUNCH
#undef GEMM
    return VOC_OK;
}

// windows [w_begin, w_begin + nw) of a request -> chunk_out[nw][Lc]; T = frames computed per window (chunk_frames, or
// fewer for windows whose tail is padding: Cfg::frames_needed), nw <= front_capacity(T)
static int front_capacity(const Engine* E, int T) { return std::max(1, (int)((long long)E->front_wave * E->cfg.chunk_frames / T)); }
static int back_capacity(const Engine* E, int T) { return std::max(1, (int)((long long)E->wave * E->cfg.chunk_frames / T)); }
static int run_wave(Engine* E, const long long* d_codes, int n_frames, int win_step, int w_begin, int nw,
                    float* chunk_out, cudaStream_t st, const int* d_wmeta = nullptr, int T = 0) {
    if (T <= 0) T = E->cfg.chunk_frames;
    float* x = nullptr; int L = 0;
    if (int r = run_front(E, d_codes, n_frames, win_step, w_begin, nw, st, d_wmeta, &x, &L, T)) return r;
    const long long Lc = E->cfg.chunk_samples();
    const int bw = E->debug ? E->wave : back_capacity(E, T);
    for (int s = 0; s < nw; s += bw) {
        if (int r = run_back(E, x, s, L, std::min(bw, nw - s), chunk_out + (long long)s * Lc, Lc, st)) return r;
    }
    return VOC_OK;
}

// --------------------------------------------------------------------------------------
// Carried-state decode (opt-in, needs transconv_trim = "right"): ONE sequence is decoded segment by
// segment; instead of the reference's 64-frame windows with 16 frames recomputed and crossfaded
// (vocoder_server.py:83-119), every causal layer finds its left context -- (k-1)*dilation rows of its input, one row
// for a transposed convolution, sliding_window-1 rows of K/V for the attention -- in a halo that precedes row 0 of its
// input buffer and was saved from the previous segment.  Output = the un-chunked decoder on the whole sequence.
// --------------------------------------------------------------------------------------
// The segment depends on the stream position only through the attention kernel, which reads it from device memory
// (E->d_strm_pos, written by the caller before the segment): the same launch sequence serves every position, so the
// live pieces of a stream (a few frames per call, ~280 small launches and copies) replay as one CUDA graph.
static int stream_segment(Engine* E, const long long* d_codes, int n, float* out, cudaStream_t st) {
    const Cfg& c = E->cfg;
    size_t hidx = 0;
    // the state buffer of the next layer in walking order (zeros = the causal zero padding of frame 0)
    auto halo_state = [&](size_t elems, float** ptr) -> int {
        if (hidx == E->strm.halos.size()) {
            float* d = nullptr;
            CK(cudaMalloc(&d, std::max<size_t>(elems, 1) * sizeof(float)));
            E->strm.halos.push_back({d, elems});
            CK(cudaMemsetAsync(d, 0, std::max<size_t>(elems, 1) * sizeof(float), st));
        }
        if (E->strm.halos[hidx].second != elems) return fail(E, VOC_E_STATE, "stream state does not match the architecture");
        *ptr = E->strm.halos[hidx++].first;
        return VOC_OK;
    };
    auto cpy2 = [&](void* d0, const void* s0, void* d1, const void* s1, size_t bytes) -> int {
        if (bytes) CK(voc_launch_copy_pair(d0, s0, d1, s1, bytes, st));
        E->launches++;
        return VOC_OK;
    };
    auto cpy = [&](void* d, const void* sr, size_t bytes) -> int { return cpy2(d, sr, nullptr, nullptr, bytes); };
    // operand tensors (two fp16 planes, or float32 on the CUDA-core path): rows [0, H) <- state; state <- rows [L, L + H)
    auto restore_op = [&](float* buf, const float* stt, int H, int C) -> int {
        const VocAct a = act(E, buf);
        const size_t n_el = (size_t)H * C;
        if (a.f) return cpy(a.f, stt, n_el * 4);
        return cpy2(a.hi, stt, a.lo, reinterpret_cast<const __half*>(stt) + n_el, n_el * 2);
    };
    auto save_op = [&](float* buf, float* stt, int H, int C, int L) -> int {
        const VocAct a = act(E, buf);
        const size_t n_el = (size_t)H * C, off = (size_t)L * C;
        if (a.f) return cpy(stt, a.f + off, n_el * 4);
        return cpy2(stt, a.hi + off, reinterpret_cast<__half*>(stt) + n_el, a.lo + off, n_el * 2);
    };
#define KLAUNCH(tag, flops, bytes, call) do { ProfScope _ps(E, st, tag, flops, bytes); CK(call); } while (0)
#define GEMM(tag, p) do { if (int _r = gemm_ck(E, p, st, tag)) return _r; } while (0)
    const int T = n;
    // ---- codebook sum -> pre-conv (k taps: k - 1 frames of history)
    const int Hpc = c.pre_conv_kernel - 1;
    KLAUNCH("rvq_gather", 0.0, 4.0 * T * c.rvq_dim * (c.num_quantizers + 1.0),
            voc_launch_rvq_gather(d_codes, n, T, T, 1, c.num_quantizers, c.codebook_size, E->rvq_tables, c.rvq_dim,
                                  act_at(E, E->f_rvq, (size_t)Hpc * c.rvq_dim), E->d_err, st, nullptr));
    {
        float* stt; if (int r = halo_state((size_t)Hpc * c.rvq_dim, &stt)) return r;
        if (int r = restore_op(E->f_rvq, stt, Hpc, c.rvq_dim)) return r;
        TapGemmParams p = gp(E->pre_conv, act(E, E->f_rvq), 0, Hpc + T, Hpc, T, 1);
        setS(p, act(E, E->f_pre), nullptr);
        GEMM("pre_conv", p);
        if (int r = save_op(E->f_rvq, stt, Hpc, c.rvq_dim, T)) return r;
    }
    float* x = E->f_pre;
    if (c.pre_transformer) {
        const int rows = T, H = c.xf_hidden, A = c.attn_dim();
        const int Hq = std::max(c.sliding_window - 1, 0);
        const double nb = 8.0 * rows * H;
        { TapGemmParams p = gp(E->xf_in, act(E, x), 0, rows, 0, rows, 1); setY(p, E->f_h); GEMM("xf.gemm", p); }
        for (int l = 0; l < c.xf_layers; ++l) {
            auto& Ly = E->xf[l];
            float* stt; if (int r = halo_state((size_t)Hq * 3 * A, &stt)) return r;
            KLAUNCH("xf.norm", 0.0, nb, voc_launch_rmsnorm(E->f_h, Ly.ln1, act(E, E->f_hn), rows, H, (float)c.rms_eps, st));
            float* qkv_new = E->f_qkv + (size_t)Hq * 3 * A;          // this segment's rows follow the history rows
            { TapGemmParams p = gp(Ly.qkv, act(E, E->f_hn), 0, rows, 0, rows, 1); setY(p, qkv_new); GEMM("xf.gemm", p); }
            if (int r = cpy(E->f_qkv, stt, (size_t)Hq * 3 * A * 4)) return r;
            KLAUNCH("xf.attn", 4.0 * c.xf_heads * (double)T * std::min(T + Hq, c.sliding_window) * c.xf_head_dim, 16.0 * rows * A,
                    voc_launch_attention_stream(E->f_qkv, act(E, E->f_att), T, c.xf_heads, c.xf_head_dim,
                                                E->rope_cos, E->rope_sin, c.sliding_window, Hq, E->d_strm_pos, st));
            if (int r = cpy(stt, E->f_qkv + (size_t)T * 3 * A, (size_t)Hq * 3 * A * 4)) return r;
            { TapGemmParams p = gp(Ly.o, act(E, E->f_att), 0, rows, 0, rows, 1); p.scale = Ly.ls_attn; setR(p, E->f_h); setY(p, E->f_h); GEMM("xf.gemm", p); }
            KLAUNCH("xf.norm", 0.0, nb, voc_launch_rmsnorm(E->f_h, Ly.ln2, act(E, E->f_hn), rows, H, (float)c.rms_eps, st));
            { TapGemmParams p = gp(Ly.gu, act(E, E->f_hn), 0, rows, 0, rows, 1); setY(p, E->f_gu); GEMM("xf.gemm", p); }
            KLAUNCH("xf.swiglu", 0.0, 12.0 * rows * c.xf_inter, voc_launch_swiglu(E->f_gu, act(E, E->f_act), rows, c.xf_inter, st));
            { TapGemmParams p = gp(Ly.down, act(E, E->f_act), 0, rows, 0, rows, 1); p.scale = Ly.ls_mlp; setR(p, E->f_h); setY(p, E->f_h); GEMM("xf.gemm", p); }
        }
        KLAUNCH("xf.norm", 0.0, nb, voc_launch_rmsnorm(E->f_h, E->xf_norm, act(E, E->f_hn), rows, H, (float)c.rms_eps, st));
        { TapGemmParams p = gp(E->xf_out, act(E, E->f_hn), 0, rows, 0, rows, 1); setS(p, act(E, E->f_x), nullptr); GEMM("xf.gemm", p); }
        x = E->f_x;
    }
    // ---- up-sampling stages: k = s transposed conv (no context), ConvNeXt (depth-wise conv: k - 1 rows)
    const int Hk = c.conv_kernel - 1;
    int L = T;
    for (size_t u = 0; u < E->ups.size(); ++u) {
        auto& U = E->ups[u];
        const int C = c.latent_dim, r = c.upsampling_ratios[u];
        float* d1 = (x == E->f_x) ? E->f_x2 : E->f_x;
        float* d2 = (d1 == E->f_x) ? E->f_x2 : E->f_x;
        const bool last_up = (u + 1 == E->ups.size());
        const size_t out_off = last_up ? (size_t)Hk * C : 0;            // conv-in wants Hk rows of history before its input
        if (!c.convnext) {
            TapGemmParams p = gp(U.convt, act(E, x), (long long)L * C, L, 0, L, 1);
            setS(p, act_at(E, d1, out_off), nullptr);
            GEMM("up.convt", p);
            L *= r; x = d1;
            continue;
        }
        float* d1n = d1 + (size_t)Hk * C;                               // float32, Hk history rows in front
        { TapGemmParams p = gp(U.convt, act(E, x), (long long)L * C, L, 0, L, 1); setY(p, d1n); GEMM("up.convt", p); }
        L *= r;
        float* stt; if (int rr = halo_state((size_t)Hk * C, &stt)) return rr;
        if (int rr = cpy(d1, stt, (size_t)Hk * C * 4)) return rr;
        KLAUNCH("up.dwconv_ln", 2.0 * L * C * c.conv_kernel, 8.0 * L * C,
                voc_launch_dwconv_ln(d1n, U.dw_w, U.dw_b, U.ln_w, U.ln_b, act(E, E->f_ln), 1, L, C, c.conv_kernel, (float)c.ln_eps, st, Hk));
        if (int rr = cpy(stt, d1 + (size_t)L * C, (size_t)Hk * C * 4)) return rr;
        { TapGemmParams p = gp(U.pw1, act(E, E->f_ln), 0, L, 0, L, 1); p.act = VOC_ACT_GELU; setS(p, act(E, E->f_mid), nullptr); GEMM("up.pw1", p); }
        { TapGemmParams p = gp(U.pw2, act(E, E->f_mid), 0, L, 0, L, 1); p.scale = U.gamma; setR(p, d1n); setS(p, act_at(E, d2, out_off), nullptr); GEMM("up.pw2", p); }
        x = d2;
    }
    // ---- decoder conv-in (k taps) -> Snake of block 0 -> the operand of block 0's transposed conv (1 row of history)
    float* bX = E->big[0].p; float* bS = E->big[1].p; float* bT = E->big[2].p; float* bX2 = E->big[3].p;
    {
        float* stt; if (int r = halo_state((size_t)Hk * c.latent_dim, &stt)) return r;
        if (int r = restore_op(x, stt, Hk, c.latent_dim)) return r;
        TapGemmParams p = gp(E->conv_in, act(E, x), 0, Hk + L, Hk, L, 1);
        setS(p, act_at(E, bS, (size_t)1 * c.decoder_dim), &E->blocks[0].s_in);
        GEMM("conv_in", p);
        if (int r = save_op(x, stt, Hk, c.latent_dim, L)) return r;
    }
    static const char* const T_CONVT[] = {"dec0.convt", "dec1.convt", "dec2.convt", "dec3.convt", "decN.convt"};
    static const char* const T_C7[] = {"dec0.ru.conv7", "dec1.ru.conv7", "dec2.ru.conv7", "dec3.ru.conv7", "decN.ru.conv7"};
    static const char* const T_C1[] = {"dec0.ru.conv1", "dec1.ru.conv1", "dec2.ru.conv1", "dec3.ru.conv1", "decN.ru.conv1"};
    static const char* const T_FU[] = {"dec0.ru.fused", "dec1.ru.fused", "dec2.ru.fused", "dec3.ru.fused", "decN.ru.fused"};
    for (size_t b = 0; b < E->blocks.size(); ++b) {
        auto& Bk = E->blocks[b];
        const size_t ti = std::min<size_t>(b, 4);
        const int C = Bk.cout;
        const int H0 = Hk * c.dilations[0];
        {   // transposed conv: out[t s + j] = x[t] W[j] + x[t - 1] W[j + s]
            float* stt; if (int r = halo_state((size_t)Bk.cin, &stt)) return r;
            if (int r = restore_op(bS, stt, 1, Bk.cin)) return r;
            TapGemmParams p = gp(Bk.convt, act(E, bS), 0, 1 + L, 1, L, 1);
            setY(p, bX); setS(p, act_at(E, bT, (size_t)H0 * C), &Bk.s_ru0_tiled);
            GEMM(T_CONVT[ti], p);
            if (int r = save_op(bS, stt, 1, Bk.cin, L)) return r;
        }
        L *= Bk.stride;
        std::swap(bS, bT);
        for (size_t j = 0; j < Bk.ru.size(); ++j) {
            auto& R = Bk.ru[j];
            const int H = Hk * c.dilations[j];
            const bool last_ru = (j + 1 == Bk.ru.size());
            const SnakeP& nxt = !last_ru ? Bk.ru[j + 1].s1 : (b + 1 < E->blocks.size() ? E->blocks[b + 1].s_in : E->head_snake);
            const int Hn = !last_ru ? Hk * c.dilations[j + 1] : (b + 1 < E->blocks.size() ? 1 : Hk);   // history the next consumer wants
            float* stt; if (int r = halo_state((size_t)H * C, &stt)) return r;
            if (int r = restore_op(bS, stt, H, C)) return r;
            bool done = false;
            if (E->tc() && E->fuse_ru) {
                RuFusedParams f;
                memset(&f, 0, sizeof f);
                const VocAct Ain = act(E, bS), Sout = act_at(E, bT, (size_t)Hn * C);
                f.A_hi = Ain.hi; f.A_lo = Ain.lo; f.L = L; f.C = C; f.B = 1; f.dil = c.dilations[j]; f.ksz = c.conv_kernel; f.a_halo = H;
                f.W7tc = R.c1.Wtc; f.w7_plane = R.c1.wtc_plane; f.w7scale = R.c1.wscale; f.bias7 = R.c1.bias;
                f.sn2_a = R.s2.a; f.sn2_invb = R.s2.invb;
                f.W1tc = R.c2.Wtc; f.w1_plane = R.c2.wtc_plane; f.w1scale = R.c2.wscale; f.bias1 = R.c2.bias;
                f.R = bX; f.Y = !last_ru ? bX2 : nullptr;
                f.S_hi = Sout.hi; f.S_lo = Sout.lo; f.snn_a = nxt.a; f.snn_invb = nxt.invb;
                if (voc_ru_fused_eligible(f)) {
                    const double el = (double)L * C;
                    ProfScope ps(E, st, T_FU[ti], 2.0 * el * C * (c.conv_kernel + 1), 4.0 * el * (f.Y ? 4.0 : 3.0));
                    CK(voc_launch_ru_fused(f, st, E->num_sms, E->tc_flags));
                    CK(opstat(E, st, T_FU[ti], f.S_hi, f.S_lo, f.B, f.L, f.C, f.C, (long long)f.L * f.C));
                    done = true;
                }
            }
            if (!done) {
                // two launches (small test architectures): conv7 -> Snake2 -> bT (the 1x1 conv needs no history); the
                // unit's output operand then overwrites bS, whose tail is saved first
                TapGemmParams p = gp(R.c1, act(E, bS), 0, H + L, H, L, 1);
                setS(p, act(E, bT), &R.s2);
                GEMM(T_C7[ti], p);
                if (int r = save_op(bS, stt, H, C, L)) return r;
                TapGemmParams p2 = gp(R.c2, act(E, bT), 0, L, 0, L, 1);
                setR(p2, bX);
                if (!last_ru) setY(p2, bX2);
                setS(p2, act_at(E, bS, (size_t)Hn * C), &nxt);
                GEMM(T_C1[ti], p2);
            } else {
                if (int r = save_op(bS, stt, H, C, L)) return r;
                std::swap(bS, bT);
            }
            if (!last_ru) std::swap(bX, bX2);
        }
    }
    // ---- head: conv k taps on the Snake'd signal + clamp
    const int ch = c.decoder_dim >> c.upsample_rates.size();
    {
        float* stt; if (int r = halo_state((size_t)Hk * ch, &stt)) return r;
        if (int r = restore_op(bS, stt, Hk, ch)) return r;
        KLAUNCH("head", 2.0 * L * ch * c.conv_kernel, 4.0 * L * (ch + 1.0),
                voc_launch_head(act_at(E, bS, (size_t)Hk * ch), (long long)L * ch, L, ch, c.conv_kernel, E->head_w, E->head_b, out, L, 1, st, Hk));
        if (int r = save_op(bS, stt, Hk, ch, L)) return r;
    }
#undef KLAUNCH
#undef GEMM
    return VOC_OK;
}

static int check_codes_flag(Engine* E, cudaStream_t st) {
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, E->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (flag) {
        CK(cudaMemsetAsync(E->d_err, 0, sizeof(int), st));
        return fail(E, VOC_E_INVALID, "audio code outside [0, codebook_size)");
    }
    return VOC_OK;
}

// A launch sequence that recurs with the same buffers (the streaming client's one-window requests, the live pieces of
// the carried-state decode) is launch-bound, so it is replayed as a CUDA graph: first sight of a key runs eagerly (lazy
// set-up: kernel attributes, tensor maps, state buffers), second sight is captured, later ones replay.  `body` must
// depend on nothing but the key.  The cache is bounded: the least recently used entry goes.
typedef std::tuple<const void*, const void*, int, int, int, int, int, int> GraphKey;
template <class F>
static int graph_or_run(Engine* E, const GraphKey& key, cudaStream_t st, F&& body) {
    auto git = E->graphs.find(key);
    if (git != E->graphs.end() && git->second.exec) {
        git->second.last_use = ++E->graph_clock;
        CK(cudaGraphLaunch(git->second.exec, st));
        E->launches += git->second.launches;
        return VOC_OK;
    }
    if (git == E->graphs.end()) {
        if (E->graphs.size() >= (size_t)E->graph_cache_max) {
            auto lru = E->graphs.begin();
            for (auto it = E->graphs.begin(); it != E->graphs.end(); ++it)
                if (it->second.last_use < lru->second.last_use) lru = it;
            if (lru->second.exec) cudaGraphExecDestroy(lru->second.exec);
            E->graphs.erase(lru);
        }
        Engine::WaveGraph& N = E->graphs[key];
        N.seen = 1; N.last_use = ++E->graph_clock;
        return body();
    }
    Engine::WaveGraph& G = git->second;
    G.last_use = ++E->graph_clock;
    if (G.seen < 0) return body();                    // known not to be capturable
    const long long l0 = E->launches;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    const int r = body();
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (r) { if (graph) cudaGraphDestroy(graph); return r; }
    if (ce != cudaSuccess || !graph) {                // not capturable here: stay eager for this key
        (void)cudaGetLastError();
        G.seen = -1000000;
        return body();
    }
    G.launches = E->launches - l0;
    E->launches = l0;
    CK(cudaGraphInstantiate(&G.exec, graph, 0));
    cudaGraphDestroy(graph);
    CK(cudaGraphLaunch(G.exec, st));
    E->launches += G.launches;
    return VOC_OK;
}

static int run_windows(Engine* E, const long long* d_codes, int n_frames, int win_step, int w0, int w1,
                       float* chunk_out, cudaStream_t st, int T = 0) {
    if (T <= 0) T = E->cfg.chunk_frames;
    const long long Lc = E->cfg.chunk_samples();
    const int fw = E->debug ? E->wave : front_capacity(E, T);      // debug captures describe one decoder wave
    for (int w = w0; w < w1; w += fw) {
        const int nw = std::min(fw, w1 - w);
        float* out = chunk_out + (long long)(w - w0) * Lc;
        const bool graphable = E->use_graphs && nw <= E->graph_max_wave && !E->profile && !E->debug && !E->opstats;
        auto body = [&]() -> int { return run_wave(E, d_codes, n_frames, win_step, w, nw, out, st, nullptr, T); };
        if (!graphable) {
            if (int r = body()) return r;
            continue;
        }
        const GraphKey key = std::make_tuple((const void*)d_codes, (const void*)out, n_frames, win_step, w, nw,
                                             E->gemm_mode, E->tc_flags * 128 + T);
        if (int r = graph_or_run(E, key, st, body)) return r;
    }
    return VOC_OK;
}

// --------------------------------------------------------------------------------------
// the chunk / stitch plan of VocoderServer.synthesize (vocoder_server.py:73-121)
// --------------------------------------------------------------------------------------
struct Plan {
    int n_windows = 0, step = 0, ov = 0;
    std::vector<int> start, len, a_len, blended; std::vector<long long> dst;
    long long total = 0;
    bool pairwise = true;
};
static Plan make_plan_raw(int mt, long long Lc, int n) {
    Plan P;
    const int spt = 1920;                                 // SAMPLES_PER_TOKEN is a constant there (:30)
    P.ov = 16 * spt; P.step = mt - 16;
    if (n <= mt) {
        P.n_windows = 1; P.step = mt; P.start = {0}; P.len = {n};
        P.a_len = {(int)std::min<long long>(Lc, (long long)n * spt)}; P.blended = {0}; P.dst = {0};
        P.total = P.a_len[0];
        return P;
    }
    long long total = 0;
    for (int s = 0; s < n; s += P.step) {
        const int ln = std::min(s + mt, n) - s;
        const int a = (int)std::min<long long>(Lc, (long long)ln * spt);
        int bl = 0; long long d;
        if (s == 0) { d = 0; total = a; }
        else if (total >= P.ov && a >= P.ov) { bl = 1; d = total - P.ov; total = total + a - P.ov; }
        else { d = total; total += a; }
        P.start.push_back(s); P.len.push_back(ln); P.a_len.push_back(a); P.blended.push_back(bl); P.dst.push_back(d);
    }
    P.n_windows = (int)P.start.size(); P.total = total;
    // the single-launch stitcher needs every blend to read un-blended samples of window w-1
    for (int w = 1; w < P.n_windows; ++w)
        if (P.blended[w]) {
            const int need = P.ov + (P.blended[w - 1] ? P.ov : 0);
            if (P.a_len[w - 1] < need) P.pairwise = false;
        }
    return P;
}

static Plan make_plan(const Cfg& c, int n) { return make_plan_raw(c.chunk_frames, c.chunk_samples(), n); }

static void fade_tables(int ov, float* fo, float* fi) {
    // exactly as numpy builds them (vocoder_server.py:108-109): linspace in float64
    // (i*step + start, last element := stop), cast to f32; fade_in = 1 - fade_out in f32
    const double step = (0.0 - 1.0) / (double)(ov - 1);
    for (int i = 0; i < ov; ++i) {
        volatile double prod = (double)i * step;           // keep mul and add separately rounded
        double y = prod + 1.0;
        if (i == ov - 1) y = 0.0;
        fo[i] = (float)y;
        fi[i] = 1.0f - fo[i];
    }
}

static int ensure_i(Engine* E, size_t n) {
    if (E->meta_cap >= n) return VOC_OK;
    if (E->d_meta) cudaFree(E->d_meta);
    E->d_meta = nullptr; E->meta_cap = 0;
    CK(cudaMalloc(&E->d_meta, n * sizeof(int))); E->meta_cap = n;
    return VOC_OK;
}
static int ensure_buf(Engine* E, DevBuf& b, size_t n) {
    if (b.n >= n) return VOC_OK;
    if (b.p) { CK(cudaStreamSynchronize(E->stream)); E->drop_graphs(); cudaFree(b.p); }   // captured graphs point into it
    b.p = nullptr; b.n = 0;
    CK(cudaMalloc(&b.p, n * sizeof(float))); b.n = n;
    return VOC_OK;
}

// windows [w0,w1) of the plan -> output samples they own, written to out (offset *out_off)
static int synth_range(Engine* E, const long long* d_codes, int n, int w0, int w1, float* d_f32, short* d_i16,
                       long long cap, long long* out_off, long long* n_out, cudaStream_t st) {
    if (!E->finalized) return fail(E, VOC_E_STATE, "voc_finalize has not been called");
    if (n <= 0 || n > 10000) return fail(E, VOC_E_INVALID, "n_tokens must be in 1..10000 (vocoder_server.py:149)");
    CK(cudaSetDevice(E->device));
    const Plan P = make_plan(E->cfg, n);
    if (w0 < 0) w0 = 0;
    if (w1 > P.n_windows) w1 = P.n_windows;
    if (w0 >= w1) { if (out_off) *out_off = 0; if (n_out) *n_out = 0; return VOC_OK; }
    const long long Lc = E->cfg.chunk_samples();
    const bool whole = (w0 == 0 && w1 == P.n_windows);
    if (!P.pairwise && !whole) return fail(E, VOC_E_INVALID, "window ranges need the pairwise-overlap regime");
    // owned output span
    const long long o_begin = P.dst[w0];
    long long o_end;
    if (w1 == P.n_windows) o_end = P.total;
    else o_end = P.dst[w1];                         // next range starts at its first window's dst
    const long long cnt = o_end - o_begin;
    if (cnt > cap) return fail(E, VOC_E_INVALID, "output buffer too small");
    const int wc0 = (w0 > 0 && P.blended[w0]) ? w0 - 1 : w0;   // recompute the neighbour for the blend
    const int nwc = w1 - wc0;
    if (int r = ensure_buf(E, E->chunks, (size_t)nwc * Lc)) return r;
    {
        // every window but possibly the request's last is full; a short last window computes only the frames its
        // samples depend on (bit-identical to computing all chunk_frames and slicing: Cfg::frames_needed)
        const int last = P.n_windows - 1;
        const int Tl = (w1 == P.n_windows && E->short_windows) ? E->cfg.frames_needed(P.len[last]) : E->cfg.chunk_frames;
        const int w_full = Tl < E->cfg.chunk_frames ? w1 - 1 : w1;
        if (w_full > wc0) if (int r = run_windows(E, d_codes, n, P.step, wc0, w_full, E->chunks.p, st)) return r;
        if (w_full < w1 && w_full >= wc0)
            if (int r = run_windows(E, d_codes, n, P.step, w_full, w1, E->chunks.p + (long long)(w_full - wc0) * Lc, st, Tl)) return r;
    }
    if (P.pairwise) {
        std::vector<int> meta((size_t)nwc * 8);
        int max_a = 0;
        for (int w = wc0; w < w1; ++w) {
            int* m = &meta[(size_t)(w - wc0) * 8];
            const bool owned = w >= w0;
            m[0] = (int)(P.dst[w] - o_begin);
            m[1] = owned ? P.a_len[w] : 0;           // the recomputed neighbour writes nothing
            m[2] = P.blended[w];
            // the trailing ov of the last owned window belongs to the next range's first window
            m[3] = (w + 1 < P.n_windows) ? P.blended[w + 1] : 0;
            m[4] = w > 0 ? P.a_len[w - 1] : 0;
            m[5] = w - wc0; m[6] = w - wc0 - 1; m[7] = 0;    // chunk slots of this window and of the one before it
            max_a = std::max(max_a, m[1]);
        }
        if (int r = ensure_i(E, meta.size())) return r;
        CK(cudaMemcpyAsync(E->d_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        {
            ProfScope ps(E, st, "stitch", 0.0, (double)cnt * (4.0 + (d_f32 ? 4.0 : 0.0) + (d_i16 ? 2.0 : 0.0)));
            CK(voc_launch_stitch(E->chunks.p, Lc, E->d_meta, nwc, P.ov, E->d_fade_out, E->d_fade_in, d_f32, d_i16, max_a, st));
        }
        // (cudaMemcpyAsync from pageable memory returns once `meta` has been staged)
    } else {
        // general regime (window shorter than two overlaps): replay the loop on the device
        float* res = d_f32;
        if (!res) { if (int r = ensure_buf(E, E->stitch_f32, (size_t)P.total)) return r; res = E->stitch_f32.p; }
        long long len = 0;
        for (int w = 0; w < P.n_windows; ++w) {
            E->launches++;
            CK(voc_launch_append_window(res, len, E->chunks.p + (long long)w * Lc, P.a_len[w], P.ov, P.blended[w], E->d_fade_out, E->d_fade_in, st));
            len = P.blended[w] ? len + P.a_len[w] - P.ov : len + P.a_len[w];
        }
        if (d_i16) { E->launches++; CK(voc_launch_pcm16(res, d_i16, P.total, st)); }
    }
    if (out_off) *out_off = o_begin;
    if (n_out) *n_out = cnt;
    return VOC_OK;
}

// Many requests in one go (what a server does with concurrent connections): the windows of
// all requests share waves, and one stitch launch -- driven by the same per-window metadata as a single
// request, with output offsets running on across requests -- writes every request's samples.
// lens[u] frames per request; out_off[u] = first output sample of request u (out_off[n_utt] = total).
static int synth_batch(Engine* E, const long long* d_codes, const int* lens, int n_utt, float* d_f32, short* d_i16,
                       long long cap, long long* out_off, cudaStream_t st) {
    if (!E->finalized) return fail(E, VOC_E_STATE, "voc_finalize has not been called");
    CK(cudaSetDevice(E->device));
    const long long Lc = E->cfg.chunk_samples();
    // every window of every request: {first frame, frames}, frames to compute, stitch record, previous window
    struct BW { int first, len, T, dst, a_len, blended, next_blended, prev_a, prev; };
    std::vector<BW> win;
    std::vector<int> utt_first;                       // index of each request's first window (+ sentinel)
    long long total = 0, frame0 = 0;
    int max_a = 0;
    for (int u = 0; u < n_utt; ++u) {
        const int n = lens[u];
        if (n <= 0 || n > 10000) return fail(E, VOC_E_INVALID, "n_tokens must be in 1..10000 (vocoder_server.py:149)");
        const Plan P = make_plan(E->cfg, n);
        if (!P.pairwise) return fail(E, VOC_E_INVALID, "batched synthesis needs the pairwise-overlap regime");
        out_off[u] = total;
        utt_first.push_back((int)win.size());
        for (int w = 0; w < P.n_windows; ++w) {
            const long long dst = total + P.dst[w];
            if (dst + P.a_len[w] > 0x7fffffffLL) return fail(E, VOC_E_INVALID, "batch output exceeds 2^31 samples");
            BW x;
            x.first = (int)(frame0 + P.start[w]); x.len = P.len[w];
            x.T = E->short_windows ? E->cfg.frames_needed(P.len[w]) : E->cfg.chunk_frames;
            x.dst = (int)dst; x.a_len = P.a_len[w]; x.blended = P.blended[w];
            x.next_blended = w + 1 < P.n_windows ? P.blended[w + 1] : 0;
            x.prev_a = w > 0 ? P.a_len[w - 1] : 0;
            x.prev = w > 0 ? (int)win.size() - 1 : -1;
            win.push_back(x);
            max_a = std::max(max_a, P.a_len[w]);
        }
        total += P.total; frame0 += n;
    }
    utt_first.push_back((int)win.size());
    out_off[n_utt] = total;
    if (total > cap) return fail(E, VOC_E_INVALID, "output buffer too small");
    // Groups of whole requests bound the chunk buffer (0.49 MB per window).  Inside a group the windows are
    // processed in order of the frames they compute -- full 64-frame windows first, then the requests' short last
    // windows bucketed by length -- so that every wave is dense; the stitch finds a window's chunk and its
    // predecessor's through slot indices.
    const int group_cap = E->debug ? std::max(E->wave, 1) * 8 : 4096;
    std::vector<int> order, slot(win.size()), wmeta, smeta;
    for (size_t u0 = 0; u0 + 1 < utt_first.size();) {
        size_t u1 = u0 + 1;
        while (u1 + 1 < utt_first.size() && utt_first[u1 + 1] - utt_first[u0] <= group_cap) ++u1;
        const int g0 = utt_first[u0], g1 = utt_first[u1], ng = g1 - g0;
        order.resize(ng);
        for (int i = 0; i < ng; ++i) order[i] = g0 + i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return win[x].T > win[y].T; });
        for (int i = 0; i < ng; ++i) slot[order[i]] = i;
        wmeta.resize((size_t)ng * 2); smeta.resize((size_t)ng * 8);
        for (int i = 0; i < ng; ++i) {
            const BW& x = win[order[i]];
            wmeta[2 * i] = x.first; wmeta[2 * i + 1] = x.len;
            int* m = &smeta[(size_t)i * 8];
            m[0] = x.dst; m[1] = x.a_len; m[2] = x.blended; m[3] = x.next_blended; m[4] = x.prev_a;
            m[5] = i; m[6] = x.prev >= 0 ? slot[x.prev] : 0; m[7] = 0;
        }
        if (int r = ensure_i(E, wmeta.size() + smeta.size())) return r;
        int* d_wmeta = E->d_meta; int* d_smeta = E->d_meta + wmeta.size();
        CK(cudaMemcpyAsync(d_wmeta, wmeta.data(), wmeta.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_smeta, smeta.data(), smeta.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));                 // the host vectors are reused by the next group
        if (int r = ensure_buf(E, E->chunks, (size_t)ng * Lc)) return r;
        for (int p0 = 0; p0 < ng;) {
            const int T = win[order[p0]].T;
            int p1 = p0;
            while (p1 < ng && win[order[p1]].T == T) ++p1;
            const int fw = E->debug ? E->wave : front_capacity(E, T);
            for (int w = p0; w < p1; w += fw) {
                const int nw = std::min(fw, p1 - w);
                if (int r = run_wave(E, d_codes, (int)frame0, 0, w, nw, E->chunks.p + (long long)w * Lc, st, d_wmeta, T)) return r;
            }
            p0 = p1;
        }
        {
            ProfScope ps(E, st, "stitch", 0.0, 0.0);
            CK(voc_launch_stitch(E->chunks.p, Lc, d_smeta, ng, 16 * 1920, E->d_fade_out, E->d_fade_in, d_f32, d_i16, max_a, st));
        }
        u0 = u1;
    }
    return VOC_OK;
}

static int ensure_codes(Engine* E, size_t n) {
    if (E->codes_cap >= n) return VOC_OK;
    if (E->d_codes) { CK(cudaStreamSynchronize(E->stream)); E->drop_graphs(); cudaFree(E->d_codes); }
    E->d_codes = nullptr; E->codes_cap = 0;
    CK(cudaMalloc(&E->d_codes, n * sizeof(long long))); E->codes_cap = n;
    return VOC_OK;
}

}  // namespace

// No C++ exception may cross the C ABI (a std::bad_alloc / length_error from a vector or map on these paths would
// otherwise reach std::terminate inside the server or the ctypes host).
template <class F>
static int guarded(void* h, F&& f) {
    try { return f(); }
    catch (const std::bad_alloc&) { if (h) ((Engine*)h)->err = "out of host memory"; return VOC_E_NOMEM; }
    catch (const std::exception& e) { if (h) ((Engine*)h)->err = std::string("internal error: ") + e.what(); return VOC_E_INVALID; }
    catch (...) { if (h) ((Engine*)h)->err = "internal error"; return VOC_E_INVALID; }
}

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" {

int voc_abi_version(void) { return 1; }

static void* create_impl(const char* cfg_json, int device, int wave) {
    auto E = std::make_unique<Engine>();
    std::string err;
    if (!cfg_from_json(cfg_json, E->cfg, err)) { g_create_error = err; fprintf(stderr, "voc_create: %s\n", err.c_str()); return nullptr; }
    if (wave < 1) wave = 1;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        g_create_error = std::string("no usable CUDA device (there is no CPU fallback): ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range");
        fprintf(stderr, "voc_create: %s\n", g_create_error.c_str());
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        g_create_error = "device is not sm_100 (this library is built for B200 only)";
        fprintf(stderr, "voc_create: %s\n", g_create_error.c_str());
        return nullptr;
    }
    E->device = device; E->wave = wave; E->num_sms = prop.multiProcessorCount;
    // experiment hooks (the documented switch is voc_set_option)
    if (const char* g = getenv("VOC_GEMM")) E->gemm_mode = !strcmp(g, "simt") ? 1 : !strcmp(g, "tc") ? 2 : 0;
    if (const char* f = getenv("VOC_TC_FLAGS")) E->tc_flags = atoi(f);
    if (const char* f = getenv("VOC_FUSE_RU")) E->fuse_ru = atoi(f) != 0;
    if (const char* f = getenv("VOC_GRAPH_MAX_WAVE")) E->graph_max_wave = atoi(f);
    if (const char* f = getenv("VOC_FRONT_WAVE")) E->front_wave = atoi(f);
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "cudaSetDevice / stream creation failed";
        fprintf(stderr, "voc_create: %s\n", g_create_error.c_str());
        return nullptr;
    }
    return E.release();
}

void* voc_create(const char* cfg_json, int device, int wave) {
    try { return create_impl(cfg_json, device, wave); }
    catch (const std::exception& e) { g_create_error = std::string("voc_create: ") + e.what(); }
    catch (...) { g_create_error = "voc_create: internal error"; }
    fprintf(stderr, "%s\n", g_create_error.c_str());
    return nullptr;
}

void voc_destroy(void* h) {
    if (!h) return;
    Engine* E = (Engine*)h;
    cudaSetDevice(E->device);
    cudaDeviceSynchronize();
    delete E;
}

int voc_set_tensor(void* h, const char* name, const float* data, long long n_elem) {
    Engine* E = (Engine*)h;
    if (!E || !name || !data || n_elem <= 0) return VOC_E_INVALID;
    if (E->finalized) return fail(E, VOC_E_STATE, "voc_set_tensor after voc_finalize");
    return guarded(h, [&]() -> int { E->raw[name].assign(data, data + n_elem); return VOC_OK; });
}

int voc_finalize(void* h) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    if (E->finalized) return fail(E, VOC_E_STATE, "already finalized");
    return guarded(h, [&]() -> int { return engine_finalize(E); });
}

// ---- model container (.b200voc = safetensors byte layout, written by weights.py:save_model) -----------
// u64 LE header length, JSON header {"__metadata__": {"voc_config": "<config JSON>", ...},
// "<tensor>": {"dtype": "F32", "shape": [...], "data_offsets": [a, b]}, ...}, then the tensor bytes.
namespace {
struct JsonCursor {
    const char* p; const char* end; bool ok = true;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool eat(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
    std::string str() {
        std::string o; ws();
        if (p >= end || *p != '"') { ok = false; return o; }
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                ++p;
                switch (*p) { case 'n': o.push_back('\n'); break; case 't': o.push_back('\t'); break;
                              case 'r': o.push_back('\r'); break; case 'b': o.push_back('\b'); break;
                              case 'f': o.push_back('\f'); break;
                              case 'u': { unsigned v = 0; for (int i = 0; i < 4 && p + 1 < end; ++i) { ++p; v = v * 16 + (unsigned)(isdigit(*p) ? *p - '0' : (tolower(*p) - 'a' + 10)); }
                                          o.push_back((char)(v < 128 ? v : '?')); break; }
                              default: o.push_back(*p); }
                ++p;
            } else o.push_back(*p++);
        }
        if (p >= end) { ok = false; return o; }
        ++p; return o;
    }
    // skips any value; for numbers / arrays of numbers collects them
    void value(std::vector<double>* nums, std::string* sval, std::map<std::string, std::string>* obj_strings,
               std::map<std::string, std::vector<double>>* obj_nums) {
        ws();
        if (p >= end) { ok = false; return; }
        if (*p == '"') { std::string v = str(); if (sval) *sval = v; return; }
        if (*p == '{') {
            ++p;
            if (eat('}')) return;
            for (;;) {
                std::string k = str(); if (!ok || !eat(':')) { ok = false; return; }
                std::vector<double> n; std::string sv;
                value(&n, &sv, nullptr, nullptr);
                if (!ok) return;
                if (obj_strings && !sv.empty()) (*obj_strings)[k] = sv;
                if (obj_nums && !n.empty()) (*obj_nums)[k] = n;
                if (eat(',')) continue;
                if (eat('}')) return;
                ok = false; return;
            }
        }
        if (*p == '[') {
            ++p;
            if (eat(']')) return;
            for (;;) {
                value(nums, nullptr, nullptr, nullptr);
                if (!ok) return;
                if (eat(',')) continue;
                if (eat(']')) return;
                ok = false; return;
            }
        }
        if (!strncmp(p, "true", 4)) { p += 4; if (nums) nums->push_back(1); return; }
        if (!strncmp(p, "false", 5)) { p += 5; if (nums) nums->push_back(0); return; }
        if (!strncmp(p, "null", 4)) { p += 4; return; }
        char* e = nullptr; const double d = strtod(p, &e);
        if (e == p) { ok = false; return; }
        p = e; if (nums) nums->push_back(d);
    }
};
}  // namespace

static void* create_from_file_impl(const char* path, int device, int wave) {
    auto bad = [&](const std::string& m) -> void* {
        g_create_error = std::string(path ? path : "(null)") + ": " + m;
        fprintf(stderr, "voc_create_from_file: %s\n", g_create_error.c_str());
        return nullptr;
    };
    if (!path) return bad("no path");
    FILE* f = fopen(path, "rb");
    if (!f) return bad("cannot open");
    unsigned long long hl = 0;
    if (fread(&hl, 8, 1, f) != 1 || hl == 0 || hl > (64ull << 20)) { fclose(f); return bad("not a .b200voc container"); }
    std::string hdr((size_t)hl, '\0');
    if (fread(&hdr[0], 1, (size_t)hl, f) != (size_t)hl) { fclose(f); return bad("truncated header"); }
    // top-level object: "__metadata__" -> strings, every other key -> {"data_offsets": [a, b], ...}
    JsonCursor c{hdr.data(), hdr.data() + hdr.size()};
    std::string cfg_json;
    std::vector<std::pair<std::string, std::pair<long long, long long>>> tensors;
    if (!c.eat('{')) { fclose(f); return bad("header is not a JSON object"); }
    if (!c.eat('}')) for (;;) {
        const std::string key = c.str();
        if (!c.ok || !c.eat(':')) { c.ok = false; break; }
        std::map<std::string, std::string> strs; std::map<std::string, std::vector<double>> nums;
        c.value(nullptr, nullptr, &strs, &nums);
        if (!c.ok) break;
        if (key == "__metadata__") cfg_json = strs["voc_config"];
        else {
            auto it = nums.find("data_offsets");
            if (it == nums.end() || it->second.size() != 2) { c.ok = false; break; }
            if (strs["dtype"] != "F32") { fclose(f); return bad("tensor " + key + " is not F32"); }
            // offsets are JSON numbers: integral, non-negative, ordered, 4-byte multiples, matching the shape
            const double a = it->second[0], b = it->second[1];
            if (!(a >= 0.0) || !(b > a) || b > 9.0e15 || a != std::floor(a) || b != std::floor(b)) { fclose(f); return bad("tensor " + key + ": bad data_offsets"); }
            long long elems = 1;
            auto sh = nums.find("shape");
            if (sh != nums.end()) for (double d : sh->second) {
                if (!(d >= 0.0) || d != std::floor(d) || d > 4.0e9 || (d > 0 && elems > (1LL << 40) / (long long)d)) { fclose(f); return bad("tensor " + key + ": bad shape"); }
                elems *= (long long)d;
            }
            const long long la = (long long)a, lb = (long long)b;
            if ((lb - la) % 4 || (sh != nums.end() && elems * 4 != lb - la)) { fclose(f); return bad("tensor " + key + ": data_offsets do not match shape"); }
            tensors.push_back({key, {la, lb}});
        }
        if (c.eat(',')) continue;
        if (c.eat('}')) break;
        c.ok = false; break;
    }
    if (!c.ok) { fclose(f); return bad("malformed header"); }
    if (cfg_json.empty()) { fclose(f); return bad("no voc_config in metadata"); }
    const long long base = 8 + (long long)hl;
    long long fsize = 0;
    if (fseeko(f, 0, SEEK_END) != 0 || (fsize = (long long)ftello(f)) < base) { fclose(f); return bad("cannot size the file"); }
    for (auto& t : tensors)
        if (t.second.second > fsize - base) { fclose(f); return bad("tensor " + t.first + " runs past the end of the file"); }
    void* h = voc_create(cfg_json.c_str(), device, wave);
    if (!h) { fclose(f); return nullptr; }
    std::vector<float> buf;
    for (auto& t : tensors) {
        const long long n = (t.second.second - t.second.first) / 4;
        buf.resize((size_t)n);
        if (fseeko(f, (off_t)(base + t.second.first), SEEK_SET) != 0 || fread(buf.data(), 4, (size_t)n, f) != (size_t)n) {
            fclose(f); voc_destroy(h); return bad("truncated tensor " + t.first);
        }
        if (voc_set_tensor(h, t.first.c_str(), buf.data(), n) != VOC_OK) { fclose(f); voc_destroy(h); return bad("voc_set_tensor failed"); }
    }
    fclose(f);
    if (voc_finalize(h) != VOC_OK) {
        g_create_error = ((Engine*)h)->err;
        voc_destroy(h);
        return nullptr;
    }
    return h;
}

void* voc_create_from_file(const char* path, int device, int wave) {
    try { return create_from_file_impl(path, device, wave); }
    catch (const std::exception& e) { g_create_error = std::string("voc_create_from_file: ") + e.what(); }
    catch (...) { g_create_error = "voc_create_from_file: internal error"; }
    fprintf(stderr, "%s\n", g_create_error.c_str());
    return nullptr;
}

int voc_max_tokens(void* h) { return h ? ((Engine*)h)->cfg.chunk_frames : VOC_E_INVALID; }
long long voc_chunk_samples(void* h) { return h ? ((Engine*)h)->cfg.chunk_samples() : VOC_E_INVALID; }
long long voc_out_samples(void* h, int n) {
    if (!h || n <= 0) return VOC_E_INVALID;
    return make_plan(((Engine*)h)->cfg, n).total;
}
int voc_num_windows(void* h, int n) {
    if (!h || n <= 0) return VOC_E_INVALID;
    return make_plan(((Engine*)h)->cfg, n).n_windows;
}

int voc_infer_chunks_dev(void* h, const long long* d_codes, int B, float* d_out, void* stream) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    if (!E->finalized) return fail(E, VOC_E_STATE, "voc_finalize has not been called");
    if (!d_codes || !d_out || B <= 0) return fail(E, VOC_E_INVALID, "bad argument");
    CK(cudaSetDevice(E->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : E->stream;
    const int T = E->cfg.chunk_frames;
    return guarded(h, [&]() -> int { return run_windows(E, d_codes, B * T, T, 0, B, d_out, st); });
}

int voc_infer_chunks(void* h, const long long* codes, int B, float* out) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    if (!E->finalized) return fail(E, VOC_E_STATE, "voc_finalize has not been called");
    if (!codes || !out || B <= 0) return fail(E, VOC_E_INVALID, "bad argument");
    CK(cudaSetDevice(E->device));
    const int T = E->cfg.chunk_frames;
    const long long Lc = E->cfg.chunk_samples();
    const size_t nc = (size_t)B * T * 16;
    if (int r = ensure_codes(E, nc)) return r;
    if (int r = ensure_buf(E, E->chunks, (size_t)B * Lc)) return r;
    // a flag left behind by an unchecked *_dev call must not fail this unrelated request
    CK(cudaMemsetAsync(E->d_err, 0, sizeof(int), E->stream));
    CK(cudaMemcpyAsync(E->d_codes, codes, nc * sizeof(long long), cudaMemcpyHostToDevice, E->stream));
    if (int r = guarded(h, [&]() -> int { return run_windows(E, E->d_codes, B * T, T, 0, B, E->chunks.p, E->stream); })) return r;
    if (int r = check_codes_flag(E, E->stream)) return r;
    CK(cudaMemcpyAsync(out, E->chunks.p, (size_t)B * Lc * sizeof(float), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return VOC_OK;
}

int voc_synthesize_range_dev(void* h, const long long* d_codes, int n_tokens, int w0, int w1, float* d_out_f32,
                             short* d_out_i16, long long cap, long long* out_offset, long long* n_out, void* stream) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    if (!d_codes || (!d_out_f32 && !d_out_i16)) return fail(E, VOC_E_INVALID, "bad argument");
    cudaStream_t st = stream ? (cudaStream_t)stream : E->stream;
    return guarded(h, [&]() -> int { return synth_range(E, d_codes, n_tokens, w0, w1, d_out_f32, d_out_i16, cap, out_offset, n_out, st); });
}

int voc_synthesize_dev(void* h, const long long* d_codes, int n_tokens, float* d_out_f32, short* d_out_i16,
                       long long cap, long long* n_out, void* stream) {
    long long off = 0;
    return voc_synthesize_range_dev(h, d_codes, n_tokens, 0, 1 << 30, d_out_f32, d_out_i16, cap, &off, n_out, stream);
}

static int synth_host(void* h, const long long* codes, int n, float* of, short* oi, long long cap, long long* n_out) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    if (!E->finalized) return fail(E, VOC_E_STATE, "voc_finalize has not been called");
    if (!codes || (!of && !oi) || !n_out) return fail(E, VOC_E_INVALID, "bad argument");
    if (n <= 0 || n > 10000) return fail(E, VOC_E_INVALID, "n_tokens must be in 1..10000 (vocoder_server.py:149)");
    CK(cudaSetDevice(E->device));
    const long long total = make_plan(E->cfg, n).total;
    if (total > cap) return fail(E, VOC_E_INVALID, "output buffer too small");
    if (int r = ensure_codes(E, (size_t)n * 16)) return r;
    CK(cudaMemsetAsync(E->d_err, 0, sizeof(int), E->stream));
    CK(cudaMemcpyAsync(E->d_codes, codes, (size_t)n * 16 * sizeof(long long), cudaMemcpyHostToDevice, E->stream));
    float* df = nullptr; short* di = nullptr;
    if (of) { if (int r = ensure_buf(E, E->stitch_f32, (size_t)total)) return r; df = E->stitch_f32.p; }
    if (oi) {
        if (E->pcm_cap < (size_t)total) {
            if (E->d_pcm) cudaFree(E->d_pcm);
            E->d_pcm = nullptr; E->pcm_cap = 0;
            CK(cudaMalloc(&E->d_pcm, (size_t)total * sizeof(short))); E->pcm_cap = (size_t)total;
        }
        di = E->d_pcm;
    }
    long long off = 0, cnt = 0;
    if (int r = guarded(h, [&]() -> int { return synth_range(E, E->d_codes, n, 0, 1 << 30, df, di, total, &off, &cnt, E->stream); })) return r;
    if (int r = check_codes_flag(E, E->stream)) return r;
    if (of) CK(cudaMemcpyAsync(of, df, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost, E->stream));
    if (oi) CK(cudaMemcpyAsync(oi, di, (size_t)cnt * sizeof(short), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    *n_out = cnt;
    return VOC_OK;
}

int voc_synthesize_f32(void* h, const long long* codes, int n_tokens, float* out, long long cap, long long* n_out) {
    return synth_host(h, codes, n_tokens, out, nullptr, cap, n_out);
}
int voc_synthesize_pcm16(void* h, const long long* codes, int n_tokens, short* out, long long cap, long long* n_out) {
    return synth_host(h, codes, n_tokens, nullptr, out, cap, n_out);
}

int voc_synthesize_batch_pcm16(void* h, const long long* codes, const int* n_tokens, int n_requests, short* out,
                               long long cap, long long* out_offsets) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    if (!E->finalized) return fail(E, VOC_E_STATE, "voc_finalize has not been called");
    if (!codes || !n_tokens || !out || !out_offsets || n_requests <= 0) return fail(E, VOC_E_INVALID, "bad argument");
    CK(cudaSetDevice(E->device));
    long long frames = 0, total = 0;
    for (int u = 0; u < n_requests; ++u) {
        if (n_tokens[u] <= 0 || n_tokens[u] > 10000) return fail(E, VOC_E_INVALID, "n_tokens must be in 1..10000 (vocoder_server.py:149)");
        frames += n_tokens[u];
        total += make_plan(E->cfg, n_tokens[u]).total;
    }
    if (total > cap) return fail(E, VOC_E_INVALID, "output buffer too small");
    if (frames > 0x7fffffffLL / 16) return fail(E, VOC_E_INVALID, "too many frames in one batch");
    if (int r = ensure_codes(E, (size_t)frames * 16)) return r;
    CK(cudaMemsetAsync(E->d_err, 0, sizeof(int), E->stream));
    CK(cudaMemcpyAsync(E->d_codes, codes, (size_t)frames * 16 * sizeof(long long), cudaMemcpyHostToDevice, E->stream));
    if (E->pcm_cap < (size_t)total) {
        if (E->d_pcm) cudaFree(E->d_pcm);
        E->d_pcm = nullptr; E->pcm_cap = 0;
        CK(cudaMalloc(&E->d_pcm, (size_t)total * sizeof(short))); E->pcm_cap = (size_t)total;
    }
    if (int r = guarded(h, [&]() -> int { return synth_batch(E, E->d_codes, n_tokens, n_requests, nullptr, E->d_pcm, total, out_offsets, E->stream); })) return r;
    if (int r = check_codes_flag(E, E->stream)) return r;
    CK(cudaMemcpyAsync(out, E->d_pcm, (size_t)total * sizeof(short), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return VOC_OK;
}

// ---- carried-state decode (opt-in; changes the output: no windows, no crossfade) --------------------------------
int voc_stream_reset(void* h) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    if (!E->finalized) return fail(E, VOC_E_STATE, "voc_finalize has not been called");
    CK(cudaSetDevice(E->device));
    for (auto& hl : E->strm.halos) CK(cudaMemsetAsync(hl.first, 0, std::max<size_t>(hl.second, 1) * sizeof(float), E->stream));
    CK(cudaStreamSynchronize(E->stream));
    E->strm.pos = 0;
    return VOC_OK;
}

long long voc_stream_position(void* h) { return h ? ((Engine*)h)->strm.pos : VOC_E_INVALID; }

static int stream_host(void* h, const long long* codes, int n, float* of, short* oi, long long cap, long long* n_out) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    if (!E->finalized) return fail(E, VOC_E_STATE, "voc_finalize has not been called");
    if (!codes || (!of && !oi) || !n_out || n <= 0) return fail(E, VOC_E_INVALID, "bad argument");
    const Cfg& c = E->cfg;
    if (c.trim_both) return fail(E, VOC_E_STATE, "carried-state decode needs transconv_trim = \"right\" (a causal, length-preserving decoder)");
    if (E->strm.pos + n > E->rope_positions) return fail(E, VOC_E_INVALID, "stream longer than the rotary table (10240 frames): call voc_stream_reset");
    const long long spf = c.samples_per_frame();
    if ((long long)n * spf > cap) return fail(E, VOC_E_INVALID, "output buffer too small");
    CK(cudaSetDevice(E->device));
    // segment length: what the activation pools hold for one sequence (the decoder wave, the front pool's qkv rows)
    const int seg_max = std::max(1, std::min(E->wave * c.chunk_frames - 1, E->front_wave * c.chunk_frames - c.sliding_window - 8));
    if (int r = ensure_codes(E, (size_t)n * 16)) return r;
    CK(cudaMemsetAsync(E->d_err, 0, sizeof(int), E->stream));
    CK(cudaMemcpyAsync(E->d_codes, codes, (size_t)n * 16 * sizeof(long long), cudaMemcpyHostToDevice, E->stream));
    if (int r = ensure_buf(E, E->chunks, (size_t)std::min(n, seg_max) * spf)) return r;
    if (oi && E->pcm_cap < (size_t)std::min(n, seg_max) * spf) {
        if (E->d_pcm) cudaFree(E->d_pcm);
        E->d_pcm = nullptr; E->pcm_cap = 0;
        CK(cudaMalloc(&E->d_pcm, (size_t)std::min(n, seg_max) * spf * sizeof(short))); E->pcm_cap = (size_t)std::min(n, seg_max) * spf;
    }
    if (!E->d_strm_pos) CK(cudaMalloc(&E->d_strm_pos, sizeof(int)));
    for (int f0 = 0; f0 < n; f0 += seg_max) {
        const int m = std::min(seg_max, n - f0);
        const int pos_now = (int)E->strm.pos;             // (pageable source: staged by the call, safe to let go)
        CK(cudaMemcpyAsync(E->d_strm_pos, &pos_now, sizeof(int), cudaMemcpyHostToDevice, E->stream));
        const long long* seg_codes = E->d_codes + (size_t)f0 * 16;
        auto body = [&]() -> int { return stream_segment(E, seg_codes, m, E->chunks.p, E->stream); };
        // live pieces (a few frames per call) are launch-bound: replay them as graphs; long segments are not
        const bool graphable = E->use_graphs && m <= E->graph_max_wave * c.chunk_frames && !E->strm.halos.empty() &&
                               !E->profile && !E->debug && !E->opstats;
        const GraphKey key = std::make_tuple((const void*)seg_codes, (const void*)E->chunks.p, m, -1, 0, 0, E->gemm_mode,
                                             E->tc_flags * 128);
        if (int r = guarded(h, [&]() -> int { return graphable ? graph_or_run(E, key, E->stream, body) : body(); })) return r;
        E->strm.pos += m;
        const size_t cnt = (size_t)m * spf;
        if (of) CK(cudaMemcpyAsync(of + (size_t)f0 * spf, E->chunks.p, cnt * sizeof(float), cudaMemcpyDeviceToHost, E->stream));
        if (oi) {
            E->launches++;
            CK(voc_launch_pcm16(E->chunks.p, E->d_pcm, (long long)cnt, E->stream));
            CK(cudaMemcpyAsync(oi + (size_t)f0 * spf, E->d_pcm, cnt * sizeof(short), cudaMemcpyDeviceToHost, E->stream));
        }
    }
    if (int r = check_codes_flag(E, E->stream)) return r;
    CK(cudaStreamSynchronize(E->stream));
    *n_out = (long long)n * spf;
    return VOC_OK;
}
int voc_stream_decode_f32(void* h, const long long* codes, int n_tokens, float* out, long long cap, long long* n_out) {
    return stream_host(h, codes, n_tokens, out, nullptr, cap, n_out);
}
int voc_stream_decode_pcm16(void* h, const long long* codes, int n_tokens, short* out, long long cap, long long* n_out) {
    return stream_host(h, codes, n_tokens, nullptr, out, cap, n_out);
}

int voc_check_dev(void* h, void* stream) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    CK(cudaSetDevice(E->device));
    return check_codes_flag(E, stream ? (cudaStream_t)stream : E->stream);
}

int voc_plan(int max_tokens, long long chunk_samples, int n_tokens, int meta_cap, int* meta,
             long long* total, int* pairwise) {
    if (max_tokens < 17 || chunk_samples < 1 || n_tokens < 1) return VOC_E_INVALID;
    const Plan P = make_plan_raw(max_tokens, chunk_samples, n_tokens);
    if (total) *total = P.total;
    if (pairwise) *pairwise = P.pairwise ? 1 : 0;
    if (meta) {
        if (meta_cap < P.n_windows * 6) return VOC_E_INVALID;
        for (int w = 0; w < P.n_windows; ++w) {
            int* m = meta + (size_t)w * 6;
            m[0] = (int)P.dst[w]; m[1] = P.a_len[w]; m[2] = P.blended[w];
            m[3] = (w + 1 < P.n_windows) ? P.blended[w + 1] : 0;
            m[4] = w > 0 ? P.a_len[w - 1] : 0; m[5] = P.start[w];
        }
    }
    return P.n_windows;
}

int voc_fade_tables(int ov, float* fade_out, float* fade_in) {
    if (ov < 2 || !fade_out || !fade_in) return VOC_E_INVALID;
    fade_tables(ov, fade_out, fade_in);
    return VOC_OK;
}

const char* voc_last_error(void* h) {
    if (!h) return g_create_error.c_str();
    return ((Engine*)h)->err.c_str();
}
long long voc_kernel_launches(void* h) { return h ? ((Engine*)h)->launches : 0; }
long long voc_simt_launches(void* h) { return h ? ((Engine*)h)->simt_launches : 0; }

int voc_set_option(void* h, const char* key, const char* value) {
    Engine* E = (Engine*)h;
    if (!E || !key || !value) return VOC_E_INVALID;
    const std::string k = key, v = value;
    if (k == "gemm") {
        if (v == "auto") E->gemm_mode = 0; else if (v == "simt") E->gemm_mode = 1; else if (v == "tc") E->gemm_mode = 2;
        else return fail(E, VOC_E_INVALID, "gemm must be auto|simt|tc");
        return VOC_OK;
    }
    if (k == "tc_flags") { E->tc_flags = atoi(v.c_str()); return VOC_OK; }
    if (k == "fuse_ru") { E->fuse_ru = (v == "1"); return VOC_OK; }
    if (k == "short_windows") { E->short_windows = (v == "1"); return VOC_OK; }
    if (k == "fold_head") { E->fold_head = (v == "1"); return VOC_OK; }
    if (k == "graphs") { E->use_graphs = (v == "1"); return VOC_OK; }
    if (k == "graph_max_wave") { E->graph_max_wave = atoi(v.c_str()); return VOC_OK; }
    if (k == "front_wave") {
        if (E->finalized) return fail(E, VOC_E_INVALID, "front_wave must be set before the weights are finalized");
        E->front_wave = atoi(v.c_str()); return VOC_OK;
    }
    if (k == "profile") { E->profile = (v == "1"); return VOC_OK; }
    if (k == "operand_stats") { E->opstats = (v == "1"); return VOC_OK; }
    if (k == "debug") { E->debug = (v == "1"); if (!E->debug) E->dbg.clear(); return VOC_OK; }
    return fail(E, VOC_E_INVALID, "unknown option " + k);
}

void* voc_stream(void* h) { return h ? (void*)((Engine*)h)->stream : nullptr; }

long long voc_profile_report(void* h, char* buf, long long cap) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    CK(cudaSetDevice(E->device));
    if (!E->prof.empty()) {
        CK(cudaDeviceSynchronize());
        struct Agg { long long calls = 0; double ms = 0, flops = 0, bytes = 0; };
        std::map<std::string, Agg> agg;
        for (auto& r : E->prof) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, r.e0, r.e1));
            Agg& a = agg[r.tag]; a.calls++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
            E->ev_pool.push_back(r.e0); E->ev_pool.push_back(r.e1);
        }
        E->prof.clear();
        std::string js = "[";
        char tmp[512];
        bool first = true;
        for (auto& kv : agg) {
            snprintf(tmp, sizeof tmp, "%s{\"tag\":\"%s\",\"calls\":%lld,\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e}",
                     first ? "" : ",", kv.first.c_str(), kv.second.calls, kv.second.ms, kv.second.flops, kv.second.bytes);
            js += tmp; first = false;
        }
        js += "]";
        E->prof_json = js;
    } else if (E->prof_json.empty()) {
        E->prof_json = "[]";
    }
    const long long need = (long long)E->prof_json.size() + 1;
    if (!buf) return need;
    if (cap < need) return fail(E, VOC_E_INVALID, "buffer too small");
    memcpy(buf, E->prof_json.c_str(), (size_t)need);
    E->prof_json.clear();
    return need;
}

long long voc_operand_report(void* h, char* buf, long long cap) {
    Engine* E = (Engine*)h;
    if (!E) return VOC_E_INVALID;
    CK(cudaSetDevice(E->device));
    if (E->opstat_json.empty()) {
        std::string js = "[";
        if (E->d_opstats && !E->opstat_tags.empty()) {
            const size_t n = E->opstat_tags.size() * Engine::OPSTAT_WORDS;
            std::vector<unsigned long long> hst(n);
            CK(cudaStreamSynchronize(E->stream));
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hst.data(), E->d_opstats, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            CK(cudaMemset(E->d_opstats, 0, (size_t)Engine::OPSTAT_SLOTS * Engine::OPSTAT_WORDS * sizeof(unsigned long long)));
            char tmp[512];
            for (size_t i = 0; i < E->opstat_tags.size(); ++i) {
                const unsigned long long* w = hst.data() + i * Engine::OPSTAT_WORDS;
                double ss; memcpy(&ss, &w[4], sizeof ss);
                const unsigned hb = (unsigned)w[5] & 0x7FFF;       // largest |hi| as fp16 bits -> float
                const int ex = (int)(hb >> 10), ma = (int)(hb & 0x3FF);
                const double mx = ex == 0 ? ldexp((double)ma, -24) : ldexp(1.0 + ma / 1024.0, ex - 15);
                snprintf(tmp, sizeof tmp, "%s{\"tag\":\"%s\",\"elements\":%llu,\"saturated\":%llu,\"hi_subnormal\":%llu,"
                         "\"lo_subnormal\":%llu,\"rms\":%.6e,\"max_abs\":%.6e}", i ? "," : "", E->opstat_tags[i].c_str(),
                         w[0], w[1], w[2], w[3], w[0] ? sqrt(ss / (double)w[0]) : 0.0, mx);
                js += tmp;
            }
            E->opstat_tags.clear();
        }
        js += "]";
        E->opstat_json = js;
    }
    const long long need = (long long)E->opstat_json.size() + 1;
    if (!buf) return need;
    if (cap < need) return fail(E, VOC_E_INVALID, "buffer too small");
    memcpy(buf, E->opstat_json.c_str(), (size_t)need);
    E->opstat_json.clear();
    return need;
}

long long voc_debug_stage(void* h, const char* name, float* out, long long cap) {
    Engine* E = (Engine*)h;
    if (!E || !name) return VOC_E_INVALID;
    auto it = E->dbg.find(name);
    if (it == E->dbg.end()) return fail(E, VOC_E_STATE, std::string("no captured stage ") + name);
    const long long n = (long long)it->second.second;
    if (!out) return n;
    if (n > cap) return fail(E, VOC_E_INVALID, "buffer too small");
    CK(cudaSetDevice(E->device));
    CK(cudaStreamSynchronize(E->stream));
    CK(cudaMemcpy(out, it->second.first->p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return n;
}


// The tile plan of a dense layer for a given batch (host arithmetic only: callable without a GPU).
int voc_tc_plan(int N, int K, int ntaps, int M, int B, int sms, int tc_flags, int* out5) {
    if (!out5) return VOC_E_INVALID;
    const TcTilePlan t = voc_tc_plan_tile(N, K, ntaps, M, B, sms, tc_flags);
    if (!t.BN) return VOC_E_INVALID;
    out5[0] = t.BN; out5[1] = t.BK; out5[2] = t.pair ? 1 : 0; out5[3] = t.p3 ? 1 : 0; out5[4] = t.three_pass ? 1 : 0;
    return VOC_OK;
}

// The shared-memory plan of the fused residual unit for a channel count, kernel size and dilation (host arithmetic).
int voc_ru_plan(int C, int ksz, int dil, int* out5) {
    if (!out5) return VOC_E_INVALID;
    return voc_ru_fused_plan(C, ksz, dil, out5) ? VOC_OK : VOC_E_INVALID;
}

// ---- kernel-level hook: one tap-GEMM on caller data, through either kernel family -------------
// mode 0 = CUDA cores, float32 operands; 1 = CUDA cores, split-fp16 operands; 2 = tcgen05.
// A [B][a_rows][K], W [ntaps*K][N] (CUDA-core layout), R / Y / S [B][M][N]; any of bias, scale, R,
// sn_a, sn_invb, Y, S may be NULL.  iters > 0 additionally times `iters` back-to-back launches with
// CUDA events and stores the mean milliseconds in *ms.  Returns 0, a negative VOC_E_*, or 1 when
// mode 2 was asked for a shape the tensor-core kernel does not take.
int voc_test_tapgemm(int device, int mode, int tc_flags, int B, int a_rows, int K, int N, int M, int a_row0,
                     int ntaps, const int* tap_off, const float* A, const float* W, const float* bias,
                     const float* scale, int act_kind, const float* R, const float* sn_a, const float* sn_invb,
                     float* Y, float* S, int iters, float* ms) {
    if (!A || !W || !tap_off || ntaps < 1 || ntaps > VOC_MAX_TAPS || B < 1 || M < 1) return VOC_E_INVALID;
    Engine EE; Engine* E = &EE;
    E->device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    E->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking));
    E->gemm_mode = mode == 0 ? 1 : 2;
    E->tc_flags = tc_flags;
    GemmW g; g.K = K; g.N = N; g.ntaps = ntaps;
    for (int i = 0; i < ntaps; ++i) g.tap_off[i] = tap_off[i];
    std::vector<float> wt(W, W + (size_t)ntaps * K * N);
    g.W = upload(E, wt);
    if (!g.W || !make_wtc(E, wt, g)) return fail(E, VOC_E_CUDA, "weight upload failed");
    auto up = [&](const float* h, size_t n) -> float* { if (!h) return nullptr; return upload(E, std::vector<float>(h, h + n)); };
    g.bias = up(bias, N);
    float* d_scale = up(scale, N);
    float* d_R = up(R, (size_t)B * M * N);
    auto up_doubled = [&](const float* h, size_t n) -> float* {      // device convention: 2 * e^alpha
        if (!h) return nullptr;
        std::vector<float> v(h, h + n);
        for (auto& x : v) x *= 2.0f;
        return upload(E, v);
    };
    SnakeP sp; sp.a = up_doubled(sn_a, N); sp.invb = up(sn_invb, N);
    const size_t na = ((size_t)B * a_rows * K + 63) / 64 * 64, no = ((size_t)B * M * N + 63) / 64 * 64;
    float *dA = nullptr, *dY = nullptr, *dS = nullptr;
    CK(cudaMalloc(&dA, na * 4)); E->owned.push_back(dA); E->cap[dA] = na;
    // Both output buffers are filled with a byte pattern and sit between guard bands of it; after the launches every
    // byte the layer does not own -- the bands, the padding behind the float32 output and behind each operand plane --
    // must still hold the pattern (compute-sanitizer is not available on the pool this library is developed on).
    const size_t G = 16384;                                  // guard floats on each side
    auto guarded_alloc = [&](float** q, size_t n) -> int {
        float* raw = nullptr;
        CK(cudaMalloc(&raw, (n + 2 * G) * 4)); E->owned.push_back(raw);
        CK(cudaMemsetAsync(raw, 0xA5, (n + 2 * G) * 4, E->stream));
        *q = raw + G;
        return VOC_OK;
    };
    // [lo, hi) in bytes relative to q must be untouched
    auto untouched = [&](const float* q, long long lo, long long hi, const char* what) -> int {
        if (hi <= lo) return VOC_OK;
        std::vector<unsigned char> g((size_t)(hi - lo));
        CK(cudaMemcpy(g.data(), reinterpret_cast<const unsigned char*>(q) + lo, g.size(), cudaMemcpyDeviceToHost));
        for (unsigned char c : g) if (c != 0xA5) return fail(E, VOC_E_CUDA, std::string("guard bytes of ") + what + " overwritten: out-of-bounds store");
        return VOC_OK;
    };
    if (int r = guarded_alloc(&dY, no)) return r;
    if (int r = guarded_alloc(&dS, no)) return r;
    E->cap[dS] = no;
    if (mode == 0) {
        CK(cudaMemcpyAsync(dA, A, (size_t)B * a_rows * K * 4, cudaMemcpyHostToDevice, E->stream));
    } else {
        std::vector<__half> h(2 * na);
        for (size_t i = 0; i < (size_t)B * a_rows * K; ++i) {
            const float v = std::min(65504.f, std::max(-65504.f, A[i]));
            h[i] = __float2half_rn(v);
            h[na + i] = __float2half_rn(v - __half2float(h[i]));
        }
        CK(cudaMemcpyAsync(dA, h.data(), h.size() * 2, cudaMemcpyHostToDevice, E->stream));
        CK(cudaStreamSynchronize(E->stream));
    }
    TapGemmParams p = gp(g, act(E, dA), (long long)a_rows * K, a_rows, a_row0, M, B);
    p.act = act_kind; p.scale = d_scale;
    if (d_R) setR(p, d_R);
    // VOC_TEST_INPLACE (timing only): the float32 output overwrites the residual, as the residual units do
    if (Y) setY(p, (d_R && getenv("VOC_TEST_INPLACE")) ? d_R : dY);
    if (S) setS(p, act(E, dS), sp.a ? &sp : nullptr);
    if (mode == 2 && !voc_tc_eligible(p)) return 1;
    CK(run_gemm(E, p, E->stream, "test", mode != 2));
    CK(cudaStreamSynchronize(E->stream));
    if (iters > 0 && ms) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, E->stream));
        for (int i = 0; i < iters; ++i) CK(run_gemm(E, p, E->stream, "test", mode != 2));
        CK(cudaEventRecord(e1, E->stream));
        CK(cudaStreamSynchronize(E->stream));
        float t = 0.f; CK(cudaEventElapsedTime(&t, e0, e1));
        *ms = t / iters;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    {
        const long long nb = (long long)B * M * N, gb = (long long)G * 4, ob = (long long)no * 4;
        const bool y_direct = Y && !(d_R && getenv("VOC_TEST_INPLACE"));
        if (int r = untouched(dY, -gb, y_direct ? 0 : ob, "Y (front)")) return r;
        if (int r = untouched(dY, y_direct ? nb * 4 : ob, ob + gb, "Y (back)")) return r;
        if (int r = untouched(dS, -gb, S ? 0 : ob, "S (front)")) return r;
        if (S && mode != 0) {                                   // two fp16 planes: [0, 2 nb) and [2 no, 2 no + 2 nb)
            if (int r = untouched(dS, nb * 2, (long long)no * 2, "S (behind the hi plane)")) return r;
            if (int r = untouched(dS, (long long)no * 2 + nb * 2, ob + gb, "S (behind the lo plane)")) return r;
        } else {
            if (int r = untouched(dS, S ? nb * 4 : ob, ob + gb, "S (back)")) return r;
        }
    }
    if (Y) CK(cudaMemcpy(Y, dY, (size_t)B * M * N * 4, cudaMemcpyDeviceToHost));
    if (S) {
        if (mode == 0) CK(cudaMemcpy(S, dS, (size_t)B * M * N * 4, cudaMemcpyDeviceToHost));
        else {
            float* tmp = nullptr;
            CK(cudaMalloc(&tmp, no * 4)); E->owned.push_back(tmp);
            const VocAct sa = act(E, dS);
            CK(voc_launch_unsplit(sa.hi, sa.lo, tmp, (long long)B * M * N, E->stream));
            CK(cudaStreamSynchronize(E->stream));
            CK(cudaMemcpy(S, tmp, (size_t)B * M * N * 4, cudaMemcpyDeviceToHost));
        }
    }
    CK(cudaDeviceSynchronize());
    return VOC_OK;
}


// ---- kernel-level hook: one residual unit on caller data, fused (1) or as the two tap-GEMM launches (0) ----------
// A [B][L][C] = Snake1(x) (float32, split here), W7 [ksz*C][C] and W1 [C][C] in the CUDA-core layout, R = x [B][L][C].
// Outputs Y = x' and S = Snake_next(x') [B][L][C] float32.  Returns 0, a negative VOC_E_*, or 1 when the fused kernel
// does not take the shape.
int voc_test_ru(int device, int fused, int tc_flags, int B, int L, int C, int ksz, int dil, const float* A,
                const float* W7, const float* b7, const float* sn2_a, const float* sn2_invb, const float* W1,
                const float* b1, const float* R, const float* snn_a, const float* snn_invb, float* Y, float* S,
                int iters, float* ms) {
    if (!A || !W7 || !W1 || !b7 || !b1 || !R || !sn2_a || !sn2_invb || !snn_a || !snn_invb || B < 1 || L < 1 ||
        ksz < 1 || ksz > VOC_MAX_TAPS) return VOC_E_INVALID;
    Engine EE; Engine* E = &EE;
    E->device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    E->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking));
    E->gemm_mode = 2; E->tc_flags = tc_flags;
    GemmW g7, g1;
    g7.K = C; g7.N = C; g7.ntaps = ksz;
    for (int i = 0; i < ksz; ++i) g7.tap_off[i] = -(ksz - 1 - i) * dil;
    g1.K = C; g1.N = C; g1.ntaps = 1; g1.tap_off[0] = 0;
    std::vector<float> w7(W7, W7 + (size_t)ksz * C * C), w1(W1, W1 + (size_t)C * C);
    g7.W = upload(E, w7); g1.W = upload(E, w1);
    if (!g7.W || !g1.W || !make_wtc(E, w7, g7) || !make_wtc(E, w1, g1)) return fail(E, VOC_E_CUDA, "weight upload failed");
    auto up = [&](const float* h, size_t n) -> float* { return upload(E, std::vector<float>(h, h + n)); };
    g7.bias = up(b7, C); g1.bias = up(b1, C);
    auto up_doubled = [&](const float* h, size_t n) -> float* {      // device convention: 2 * e^alpha
        std::vector<float> v(h, h + n);
        for (auto& x : v) x *= 2.0f;
        return upload(E, v);
    };
    SnakeP s2, sn; s2.a = up_doubled(sn2_a, C); s2.invb = up(sn2_invb, C); sn.a = up_doubled(snn_a, C); sn.invb = up(snn_invb, C);
    const size_t ne = (size_t)B * L * C, na = (ne + 63) / 64 * 64;
    float* dR = up(R, ne);
    // Every output buffer sits between guard bands of a known byte pattern that are checked after the launches: an
    // out-of-bounds store of the kernels under test fails the call (compute-sanitizer is not available on the GPU
    // pool this library is developed on; profiles/r2_compute_sanitizer_refused.txt).
    const size_t G = 16384;                                  // guard floats on each side
    auto guarded_alloc = [&](float** p, size_t n) -> int {
        float* raw = nullptr;
        CK(cudaMalloc(&raw, (n + 2 * G) * 4)); E->owned.push_back(raw);
        CK(cudaMemsetAsync(raw, 0xA5, (n + 2 * G) * 4, E->stream));
        *p = raw + G;
        return VOC_OK;
    };
    auto guards_intact = [&](const float* p, size_t n, const char* what) -> int {
        std::vector<unsigned char> g(2 * G * 4);
        CK(cudaMemcpy(g.data(), p - G, G * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g.data() + G * 4, p + n, G * 4, cudaMemcpyDeviceToHost));
        for (unsigned char c : g) if (c != 0xA5) return fail(E, VOC_E_CUDA, std::string("guard band of ") + what + " overwritten: out-of-bounds store");
        return VOC_OK;
    };
    float *dA = nullptr, *dT = nullptr, *dY = nullptr, *dS = nullptr, *tmp = nullptr;
    CK(cudaMalloc(&dA, na * 4)); E->owned.push_back(dA); E->cap[dA] = na;
    if (int r = guarded_alloc(&dT, na)) return r;            // operand tensors: two fp16 planes in the bytes of na floats
    E->cap[dT] = na;
    if (int r = guarded_alloc(&dS, na)) return r;
    E->cap[dS] = na;
    if (int r = guarded_alloc(&dY, na)) return r;
    CK(cudaMalloc(&tmp, na * 4)); E->owned.push_back(tmp);
    CK(cudaMemsetAsync(dY, 0, na * 4, E->stream)); CK(cudaMemsetAsync(dS, 0, na * 4, E->stream));
    CK(cudaMemsetAsync(dT, 0, na * 4, E->stream));
    {
        std::vector<__half> h(2 * na);
        for (size_t i = 0; i < ne; ++i) {
            const float v = std::min(65504.f, std::max(-65504.f, A[i]));
            h[i] = __float2half_rn(v);
            h[na + i] = __float2half_rn(v - __half2float(h[i]));
        }
        CK(cudaMemcpyAsync(dA, h.data(), h.size() * 2, cudaMemcpyHostToDevice, E->stream));
        CK(cudaStreamSynchronize(E->stream));
    }
    RuFusedParams f;
    memset(&f, 0, sizeof f);
    const VocAct Ain = act(E, dA), Sout = act(E, dS);
    f.A_hi = Ain.hi; f.A_lo = Ain.lo; f.L = L; f.C = C; f.B = B; f.dil = dil; f.ksz = ksz;
    f.W7tc = g7.Wtc; f.w7_plane = g7.wtc_plane; f.w7scale = g7.wscale; f.bias7 = g7.bias; f.sn2_a = s2.a; f.sn2_invb = s2.invb;
    f.W1tc = g1.Wtc; f.w1_plane = g1.wtc_plane; f.w1scale = g1.wscale; f.bias1 = g1.bias;
    f.R = dR; f.Y = Y ? dY : nullptr; f.S_hi = Sout.hi; f.S_lo = Sout.lo; f.snn_a = sn.a; f.snn_invb = sn.invb;
    if (fused && !voc_ru_fused_eligible(f)) return 1;
    auto once = [&]() -> int {
        if (fused) { CK(voc_launch_ru_fused(f, E->stream, E->num_sms, E->tc_flags)); return VOC_OK; }
        TapGemmParams p = gp(g7, act(E, dA), (long long)L * C, L, 0, L, B);
        setS(p, act(E, dT), &s2);
        CK(run_gemm(E, p, E->stream, "test.conv7"));
        TapGemmParams p2 = gp(g1, act(E, dT), (long long)L * C, L, 0, L, B);
        setR(p2, dR); if (Y) setY(p2, dY); setS(p2, act(E, dS), &sn);
        CK(run_gemm(E, p2, E->stream, "test.conv1"));
        return VOC_OK;
    };
    if (int r = once()) return r;
    CK(cudaStreamSynchronize(E->stream));
    if (iters > 0 && ms) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, E->stream));
        for (int i = 0; i < iters; ++i) if (int r = once()) return r;
        CK(cudaEventRecord(e1, E->stream));
        CK(cudaStreamSynchronize(E->stream));
        float t = 0.f; CK(cudaEventElapsedTime(&t, e0, e1));
        *ms = t / iters;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    if (int r = guards_intact(dY, na, "Y")) return r;
    if (int r = guards_intact(dS, na, "S")) return r;
    if (int r = guards_intact(dT, na, "T")) return r;
    if (Y) CK(cudaMemcpy(Y, dY, ne * 4, cudaMemcpyDeviceToHost));
    if (S) {
        CK(voc_launch_unsplit(Sout.hi, Sout.lo, tmp, (long long)ne, E->stream));
        CK(cudaStreamSynchronize(E->stream));
        CK(cudaMemcpy(S, tmp, ne * 4, cudaMemcpyDeviceToHost));
    }
    CK(cudaDeviceSynchronize());
    return VOC_OK;
}

}  // extern "C"
