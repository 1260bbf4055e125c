// Model plan, parameter table, weight packing and the C ABI (include/eabnet_b200.h).
//
// The architecture walk below restates EaBNet.__init__/forward (EaBNet.py:9-125 and the sub-modules at
// :157-624) as a list of layer descriptors that (a) declare the reference state_dict entries in the reference's
// registration order, (b) pack them into the kernels' layouts, and (c) launch the kernels.  Activations are
// channels-last [B,T,F,C] fp32 and are stored RAW next to their (sum, sumsq) statistics; normalisation and
// PReLU are applied by the consumer while it stages its operand (see common.cuh Xform).
// Split over translation units: model_build.cu (layer descriptors / parameter table), model_pack.cu (weight packing),
// model_run.cu (forward orchestration, planning, streaming); this file holds the per-thread runtime state (errors, launch
// counters, profiler, per-device launch state) and the C ABI.
#include "model_internal.h"

namespace eab {

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static thread_local int g_launches = 0;
void set_error(const std::string& m) { g_err = m; }
int fail(const std::string& m) { g_err = m; return 1; }
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
}
namespace {
struct DevState { int sms = 0; std::map<const void*, int> smem; };
std::mutex g_dev_mu;
std::map<int, DevState> g_dev;
}  // namespace
int device_sm_count(int* sms) {
    int dev = 0;
    EAB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_dev_mu);
    DevState& d = g_dev[dev];
    if (!d.sms) EAB_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    *sms = d.sms;
    return 0;
}
int ensure_dynamic_smem(const void* kernel, int bytes) {
    int dev = 0;
    EAB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_dev_mu);
    int& cur = g_dev[dev].smem[kernel];
    if (bytes > cur) {
        EAB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        cur = bytes;
    }
    return 0;
}
void count_launch(int n) { g_launches += n; }

// ------------------------------------------------------------------------------------------------ profiler
struct ProfRec { const char* cat; double flops, bytes, moved; cudaEvent_t a, b; };
struct Profiler {
    bool on = false;
    bool detail = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        cudaEvent_t e;
        if (!pool.empty()) { e = pool.back(); pool.pop_back(); return e; }
        cudaEventCreate(&e);
        return e;
    }
    void clear() {
        for (auto& r : recs) { pool.push_back(r.a); pool.push_back(r.b); }
        recs.clear();
    }
};
static thread_local Profiler g_prof;

ProfScope::ProfScope(const char* category, double flops, double bytes, cudaStream_t s, double moved) : rec(nullptr), st(s) {
    if (!g_prof.on) return;
    ProfRec r{category, flops, bytes, moved < 0 ? bytes : moved, g_prof.get(), g_prof.get()};
    cudaEventRecord(r.a, st);
    g_prof.recs.push_back(r);
    rec = reinterpret_cast<void*>(g_prof.recs.size());      // index + 1
}
ProfScope::~ProfScope() {
    if (!rec) return;
    cudaEventRecord(g_prof.recs[reinterpret_cast<size_t>(rec) - 1].b, st);
}
bool g_prof_on() { return g_prof.on; }
int launch_count() { return g_launches; }
void reset_launch_count() { g_launches = 0; }

}  // namespace eab

// =================================================================================================== C ABI
extern "C" {

size_t eab_workspace_bytes(const eab_model* m, int B, int T);

int eab_create(const eab_config* cfg, eab_model** out) {
    if (!cfg || !out) return fail("eab_create: null argument");
    std::unique_ptr<eab_model> m(new eab_model());
    m->cfg = *cfg;
    if (m->cfg.n_freq <= 0) m->cfg.n_freq = 161;
    if (build(m.get())) return 1;
    // the 1x1-conv heads feed decoder round-off straight into the beam weights (no recurrent smoothing, and "miso"
    // sums 161 bins): keep their decoder fp32-grade as well
    if (!(m->cfg.topo_type == 0 && m->cfg.bf_type == 0)) m->opt_dec_passes = 3;
    *out = m.release();
    return 0;
}

int eab_gag_create(const eab_gag_config* cfg, eab_model** out) {
    if (!cfg || !out) return fail("eab_gag_create: null argument");
    std::unique_ptr<eab_model> m(new eab_model());
    m->kind = 1;
    m->gcfg = *cfg;
    if (build_gag(m.get())) return 1;
    // the post-filter refines an estimate that is already within tolerance: keep its whole encoder fp32-grade
    m->opt_enc_passes = 3;
    m->opt_inner_passes = 3;
    *out = m.release();
    return 0;
}

size_t eab_gag_workspace_bytes(const eab_model* m, int B, int T) {
    if (!m || m->kind != 1) return 0;
    return eab_workspace_bytes(m, B, T);
}

int eab_gag_forward(eab_model* m, const float* inpt, const int64_t inpt_strides[4], const float* pre, float* out, int B, int T,
                    void* ws, size_t ws_bytes, void* stream) {
    if (!m || !inpt || !inpt_strides || !pre || !out || !ws) return fail("eab_gag_forward: null argument");
    const long long st4[4] = {(long long)inpt_strides[0], (long long)inpt_strides[1], (long long)inpt_strides[2], (long long)inpt_strides[3]};
    reset_launch_count();
    const int rc = gag_forward(m, inpt, st4, pre, out, B, T, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    m->last_launches = launch_count();
    return rc;
}

void eab_destroy(eab_model* m) {
    if (!m) return;
    if (m->blob) cudaFree(m->blob);
    if (m->scratch) cudaFree(m->scratch);
    for (auto& sg : m->slot_graph) if (sg.exec) cudaGraphExecDestroy(sg.exec);
    if (m->s_in) {
        cudaStreamDestroy(m->s_in);
        cudaStreamDestroy(m->s_out);
        cudaStreamDestroy(m->s_comp);
        cudaStreamDestroy(m->s_comp2);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(m->ev_in[i]); cudaEventDestroy(m->ev_comp[i]); cudaEventDestroy(m->ev_out[i]); }
    }
    delete m;
}

int eab_param_count(const eab_model* m) { return m ? (int)m->params.size() : 0; }

int eab_param_info(const eab_model* m, int i, const char** name, int* ndim, int64_t shape[4], int* kind, int* fan_in) {
    if (!m || i < 0 || i >= (int)m->params.size()) return fail("eab_param_info: index out of range");
    const Param& p = m->params[i];
    if (name) *name = p.name.c_str();
    if (ndim) *ndim = p.ndim;
    if (shape) for (int k = 0; k < 4; ++k) shape[k] = p.shape[k];
    if (kind) *kind = p.kind;
    if (fan_in) *fan_in = p.fan_in;
    return 0;
}

int eab_set_param(eab_model* m, const char* name, const float* host, int64_t numel) {
    if (!m || !name) return fail("eab_set_param: null argument");
    auto it = m->index.find(name);
    if (it == m->index.end()) return fail(std::string("unexpected key in state_dict: ") + name);
    Param& p = m->params[it->second];
    if (p.kind == EAB_P_BN_COUNT) { p.set = true; return 0; }
    if (numel != p.numel()) return fail(std::string("size mismatch for ") + name);
    if (!host) return fail("eab_set_param: null data");
    p.host.assign(host, host + numel);
    p.set = true;
    m->dirty = true;
    return 0;
}

int eab_commit_params(eab_model* m, void* stream) {
    if (!m) return fail("null model");
    return commit(m, static_cast<cudaStream_t>(stream));
}

size_t eab_workspace_bytes(const eab_model* m, int B, int T) {
    size_t sb = 0, tb = 0;
    if (!m || B < 1 || T < 1) return 0;
    if (plan(const_cast<eab_model*>(m), B, T, &sb, &tb)) return 0;
    return tb;
}

int eab_forward(eab_model* m, const float* inpt, float* out, int B, int T, void* ws, size_t ws_bytes, void* stream) {
    if (!m || !inpt || !out || !ws) return fail("eab_forward: null argument");
    reset_launch_count();
    const int rc = forward(m, inpt, out, B, T, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    m->last_launches = launch_count();
    return rc;
}

int eab_stft(const float* wave, float* spec, int B, int M, int L, void* stream) {
    if (!wave || !spec) return fail("eab_stft: null argument");
    return launch_stft(wave, spec, B, M, L, static_cast<cudaStream_t>(stream));
}

int eab_istft(const float* spec, float* wave, int B, int T, void* stream) {
    if (!spec || !wave) return fail("eab_istft: null argument");
    return launch_istft(spec, wave, B, T, static_cast<cudaStream_t>(stream));
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

size_t eab_enhance_workspace_bytes(const eab_model* m, int B, int L) {
    if (!m || B < 1 || L < 161) return 0;
    const int T = 1 + L / 160;
    const size_t fw = eab_workspace_bytes(m, B, T);
    if (!fw) return 0;
    const size_t spec = align256((size_t)B * T * m->cfg.n_freq * m->cfg.M * 2 * sizeof(float));
    const size_t outp = align256((size_t)B * 2 * T * m->cfg.n_freq * sizeof(float));
    return spec + outp + fw;
}

// wave -> wave on device buffers.  Input: fp32 [B][M][L], or (pcm != null) int16 PCM in file channel order with microphone
// m = file channel order[m]; output: fp32 samples, or (enhanced16 != null) the reference's int16 writer.
static int enhance_dev(eab_model* m, const float* wave, const int16_t* pcm, const int* order, float* enhanced, int16_t* enhanced16,
                       int B, int L, void* ws, size_t ws_bytes, void* stream) {
    if (!m || (!wave && !pcm) || (!enhanced && !enhanced16) || !ws) return fail("eab_enhance: null argument");
    if (m->kind != 0) return fail("eab_enhance needs an EaBNet handle");
    if (m->cfg.topo_type == 1) return fail("eab_enhance: the 'miso' topology returns [B,2,T], which has no iSTFT");
    if (m->cfg.n_freq != 161) return fail("eab_enhance: the 320-point STFT gives 161 bins");
    if (L < 161) return fail("eab_enhance: need at least 161 samples");
    const size_t need = eab_enhance_workspace_bytes(m, B, L);
    if (!need) return 1;
    if (ws_bytes < need) return fail("workspace too small: need " + std::to_string(need) + " bytes");
    const int T = 1 + L / 160;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(ws);
    float* spec = reinterpret_cast<float*>(p);
    p += align256((size_t)B * T * 161 * m->cfg.M * 2 * sizeof(float));
    float* outp = reinterpret_cast<float*>(p);
    p += align256((size_t)B * 2 * T * 161 * sizeof(float));
    reset_launch_count();
    // the forward workspace is idle until the STFT has finished: it doubles as the STFT's plane scratch
    const size_t rest = ws_bytes - (size_t)(p - static_cast<char*>(ws));
    int rc = pcm ? launch_stft_pcm16(pcm, order, spec, B, m->cfg.M, L, st, p, rest) : launch_stft(wave, spec, B, m->cfg.M, L, st, p, rest);
    if (!rc) rc = forward(m, spec, outp, B, T, p, rest, st);
    if (!rc) rc = launch_istft(outp, enhanced, B, T, st, enhanced16);
    m->last_launches = launch_count();
    return rc;
}

int eab_enhance(eab_model* m, const float* wave, float* enhanced, int B, int L, void* ws, size_t ws_bytes, void* stream) {
    if (!wave || !enhanced) return fail("eab_enhance: null argument");
    return enhance_dev(m, wave, nullptr, nullptr, enhanced, nullptr, B, L, ws, ws_bytes, stream);
}

// enhance.py:49-62 on device buffers: STFT + compression, EaBNet, GaGNet on (reference microphone, EaBNet estimate), iSTFT
// of the last glance-gaze module's estimate.
size_t eab_enhance_postnet_workspace_bytes(const eab_model* eab, const eab_model* gag, int B, int L) {
    if (!eab || !gag || eab->kind != 0 || gag->kind != 1 || B < 1 || L < 161) return 0;
    const int T = 1 + L / 160;
    const size_t f1 = eab_workspace_bytes(eab, B, T), f2 = eab_workspace_bytes(gag, B, T);
    if (!f1 || !f2) return 0;
    const size_t spec = align256((size_t)B * T * eab->cfg.n_freq * eab->cfg.M * 2 * sizeof(float));
    const size_t est = align256((size_t)B * 2 * T * eab->cfg.n_freq * sizeof(float));
    return spec + est * (1 + gag->cfg.q) + std::max(f1, f2);
}

static int enhance_postnet_dev(eab_model* eab, eab_model* gag, int ref_mic, const float* wave, const int16_t* pcm, const int* order,
                               float* enhanced, int16_t* enhanced16, int B, int L, void* ws, size_t ws_bytes, void* stream) {
    if (!eab || !gag || (!wave && !pcm) || (!enhanced && !enhanced16) || !ws) return fail("eab_enhance_postnet: null argument");
    if (eab->kind != 0 || gag->kind != 1) return fail("eab_enhance_postnet: needs an EaBNet handle and a GaGNet handle");
    if (eab->cfg.topo_type == 1) return fail("eab_enhance_postnet: the 'miso' topology returns [B,2,T]");
    if (eab->cfg.n_freq != 161 || gag->cfg.n_freq != 161) return fail("eab_enhance_postnet: the 320-point STFT gives 161 bins");
    if (ref_mic < 0 || ref_mic >= eab->cfg.M) return fail("eab_enhance_postnet: ref_mic out of range");
    const size_t need = eab_enhance_postnet_workspace_bytes(eab, gag, B, L);
    if (!need) return fail("eab_enhance_postnet: bad shape");
    if (ws_bytes < need) return fail("workspace too small: need " + std::to_string(need) + " bytes");
    const int T = 1 + L / 160, F = 161, M = eab->cfg.M;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(ws);
    float* spec = reinterpret_cast<float*>(p);
    p += align256((size_t)B * T * F * M * 2 * sizeof(float));
    const size_t est = align256((size_t)B * 2 * T * F * sizeof(float));
    float* est0 = reinterpret_cast<float*>(p);
    p += est;
    float* stages = reinterpret_cast<float*>(p);
    p += est * gag->cfg.q;
    // (the q estimates are written back to back, B*2*T*F floats each; only the last one is read here)
    const size_t rest = ws_bytes - (size_t)(p - static_cast<char*>(ws));
    reset_launch_count();
    // (the forward workspace is idle until the STFT has finished: it doubles as the STFT's plane scratch)
    int rc = pcm ? launch_stft_pcm16(pcm, order, spec, B, M, L, st, p, rest) : launch_stft(wave, spec, B, M, L, st, p, rest);
    if (!rc) rc = forward(eab, spec, est0, B, T, p, rest, st);
    const long long strides[4] = {(long long)T * F * M * 2, 1, (long long)F * M * 2, (long long)M * 2};
    if (!rc) rc = gag_forward(gag, spec + (size_t)ref_mic * 2, strides, est0, stages, B, T, p, rest, st);
    if (!rc) rc = launch_istft(stages + (size_t)(gag->cfg.q - 1) * B * 2 * T * F, enhanced, B, T, st, enhanced16);
    const int n = launch_count();
    eab->last_launches = n;
    gag->last_launches = n;
    return rc;
}

int eab_enhance_postnet(eab_model* eab, eab_model* gag, int ref_mic, const float* wave, float* enhanced, int B, int L, void* ws,
                        size_t ws_bytes, void* stream) {
    if (!wave || !enhanced) return fail("eab_enhance_postnet: null argument");
    return enhance_postnet_dev(eab, gag, ref_mic, wave, nullptr, nullptr, enhanced, nullptr, B, L, ws, ws_bytes, stream);
}

// Host front door.  Batch i+1 is uploaded (copy stream) and batch i-1 downloaded (second copy stream) while batch i
// computes on the caller's stream: two device input slots, two output slots, one workspace.
// pcm: the batches are int16 PCM in file channel order (microphone m = file channel order[m]) and the results int16 PCM
static int host_batches(eab_model* m, const void* const* waves_host, void* const* enhanced_host, bool pcm, const int* order,
                        int n_batches, int B, int L, void* stream) {
    if (!m || !waves_host || !enhanced_host) return fail("eab_enhance_host_batches: null argument");
    if (m->kind != 0) return fail("eab_enhance_host_batches needs an EaBNet handle");
    int ord[64];
    unsigned long long ord_hash = pcm ? 1469598103934665603ull : 0ull;
    if (pcm) {
        if (m->cfg.M > 64) return fail("pcm16: at most 64 microphones");
        for (int i = 0; i < m->cfg.M; ++i) {
            ord[i] = order ? order[i] : i;
            if (ord[i] < 0 || ord[i] >= m->cfg.M) return fail("pcm16: mic_order entries must be in [0, M)");
            ord_hash = (ord_hash ^ (unsigned long long)(ord[i] + 1)) * 1099511628211ull;
        }
    }
    const size_t esz = pcm ? sizeof(int16_t) : sizeof(float);
    if (n_batches < 1) return 0;
    for (int i = 0; i < n_batches; ++i)
        if (!waves_host[i] || !enhanced_host[i]) return fail("eab_enhance_host_batches: null batch pointer");
    const size_t need = eab_enhance_workspace_bytes(m, B, L);
    if (!need) return fail("eab_enhance_host_batches: bad shape");
    const int nslot = n_batches > 1 ? 2 : 1;
    const size_t in_bytes = (size_t)B * m->cfg.M * L * esz;
    const size_t out_bytes = (size_t)B * 160 * (L / 160) * esz;
    const size_t in_b = align256(in_bytes), out_b = align256(out_bytes);
    // two compute streams (even / odd batches, a workspace each): consecutive batches overlap wherever one leaves SMs idle
    // (the LSTM runs on 108 of the 148 SMs for 4 ms of a 22 ms step)
    const bool dual = m->opt_dual_stream && nslot == 2;
    const size_t total = nslot * (in_b + out_b) + (dual ? 2 : 1) * align256(need);
    if (m->scratch_bytes < total) {
        if (m->scratch) cudaFree(m->scratch);
        m->scratch = nullptr;
        m->scratch_bytes = 0;
        EAB_CUDA(cudaMalloc(&m->scratch, total));
        m->scratch_bytes = total;
    }
    if (!m->s_in) {
        EAB_CUDA(cudaStreamCreateWithFlags(&m->s_in, cudaStreamNonBlocking));
        EAB_CUDA(cudaStreamCreateWithFlags(&m->s_out, cudaStreamNonBlocking));
        EAB_CUDA(cudaStreamCreateWithFlags(&m->s_comp, cudaStreamNonBlocking));
        EAB_CUDA(cudaStreamCreateWithFlags(&m->s_comp2, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            EAB_CUDA(cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming));
            EAB_CUDA(cudaEventCreateWithFlags(&m->ev_comp[i], cudaEventDisableTiming));
            EAB_CUDA(cudaEventCreateWithFlags(&m->ev_out[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t caller = static_cast<cudaStream_t>(stream);
    // the legacy default stream cannot be captured into a graph: compute on the library's own stream then, ordered after the
    // caller's queued work (the call synchronises before it returns, so the caller's stream semantics are unchanged)
    cudaStream_t st = caller ? caller : m->s_comp;
    if (!caller) {
        EAB_CUDA(cudaEventRecord(m->ev_out[0], caller));
        EAB_CUDA(cudaStreamWaitEvent(st, m->ev_out[0], 0));
        stream = st;
    }
    char* p = static_cast<char*>(m->scratch);
    char* din[2] = {p, p + (nslot - 1) * in_b};
    char* dout[2] = {p + nslot * in_b, p + nslot * in_b + (nslot - 1) * out_b};
    char* ws_slot[2] = {p + nslot * (in_b + out_b), p + nslot * (in_b + out_b) + (dual ? align256(need) : 0)};
    cudaStream_t cs[2] = {st, dual ? m->s_comp2 : st};
    // the copy streams start after everything already queued on the caller's stream (scratch may still be in use)
    EAB_CUDA(cudaEventRecord(m->ev_comp[0], st));
    EAB_CUDA(cudaStreamWaitEvent(m->s_in, m->ev_comp[0], 0));
    if (dual) EAB_CUDA(cudaStreamWaitEvent(m->s_comp2, m->ev_comp[0], 0));
    int launches = 0;
    for (int i = 0; i < n_batches; ++i) {
        const int s = i & 1;
        cudaStream_t cst = cs[s];
        char* ws = ws_slot[s];
        if (i >= 2) EAB_CUDA(cudaStreamWaitEvent(m->s_in, m->ev_comp[s], 0));          // batch i-2 has consumed this input slot
        EAB_CUDA(cudaMemcpyAsync(din[s], waves_host[i], in_bytes, cudaMemcpyHostToDevice, m->s_in));
        EAB_CUDA(cudaEventRecord(m->ev_in[s], m->s_in));
        EAB_CUDA(cudaStreamWaitEvent(cst, m->ev_in[s], 0));
        if (i >= 2) EAB_CUDA(cudaStreamWaitEvent(cst, m->ev_out[s], 0));               // batch i-2 has left this output slot
        // the slot's step: replayed from a CUDA graph once the slot has run it directly (same buffers, same shape, same weights)
        {
            eab_model::SlotGraph& sg = m->slot_graph[s];
            const bool match = sg.exec && sg.in == din[s] && sg.out == dout[s] && sg.ws == ws && sg.B == B && sg.L == L &&
                               sg.version == m->param_version && sg.mode == ord_hash;
            auto step = [&]() {
                return pcm ? enhance_dev(m, nullptr, reinterpret_cast<const int16_t*>(din[s]), ord, nullptr, reinterpret_cast<int16_t*>(dout[s]),
                                         B, L, ws, need, cst)
                           : enhance_dev(m, reinterpret_cast<const float*>(din[s]), nullptr, nullptr, reinterpret_cast<float*>(dout[s]), nullptr,
                                         B, L, ws, need, cst);
            };
            if (match) {
                EAB_CUDA(cudaGraphLaunch(sg.exec, cst));
                m->last_launches = sg.launches;
            } else if (m->opt_host_graph && i >= 2 && !g_prof_on()) {
                if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
                cudaGraph_t graph = nullptr;
                bool ok = cudaStreamBeginCapture(cst, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                int rc = 1;
                if (ok) {
                    rc = step();
                    ok = cudaStreamEndCapture(cst, &graph) == cudaSuccess && rc == 0 && graph != nullptr;
                }
                if (ok) ok = cudaGraphInstantiate(&sg.exec, graph, 0) == cudaSuccess;
                if (graph) cudaGraphDestroy(graph);
                if (ok) {
                    sg.in = din[s]; sg.out = dout[s]; sg.ws = ws; sg.B = B; sg.L = L; sg.version = m->param_version; sg.mode = ord_hash;
                    sg.launches = m->last_launches;
                    EAB_CUDA(cudaGraphLaunch(sg.exec, cst));
                } else {
                    sg.exec = nullptr;
                    cudaGetLastError();                        // capture refused (e.g. a legacy-stream caller): run the step directly
                    EAB_TRY(step());
                }
            } else {
                EAB_TRY(step());
            }
        }
        launches += m->last_launches;
        EAB_CUDA(cudaEventRecord(m->ev_comp[s], cst));
        EAB_CUDA(cudaStreamWaitEvent(m->s_out, m->ev_comp[s], 0));
        EAB_CUDA(cudaMemcpyAsync(enhanced_host[i], dout[s], out_bytes, cudaMemcpyDeviceToHost, m->s_out));
        EAB_CUDA(cudaEventRecord(m->ev_out[s], m->s_out));
    }
    m->last_launches = launches;
    EAB_CUDA(cudaStreamSynchronize(m->s_out));
    EAB_CUDA(cudaStreamSynchronize(st));
    if (dual) EAB_CUDA(cudaStreamSynchronize(m->s_comp2));
    return 0;
}

int eab_enhance_host_batches(eab_model* m, const float* const* waves_host, float* const* enhanced_host, int n_batches,
                             int B, int L, void* stream) {
    return host_batches(m, reinterpret_cast<const void* const*>(waves_host), reinterpret_cast<void* const*>(enhanced_host), false, nullptr,
                        n_batches, B, L, stream);
}

// The same pipeline on the 16-bit PCM wire format (half the H2D / D2H bytes): the int16 -> float conversion and the microphone
// permutation happen in the STFT's operand staging, the int16 writer in the iSTFT's store.
int eab_enhance_host_batches_pcm16(eab_model* m, const int16_t* const* pcm_host, const int* mic_order, int16_t* const* enhanced_pcm_host,
                                   int n_batches, int B, int L, void* stream) {
    return host_batches(m, reinterpret_cast<const void* const*>(pcm_host), reinterpret_cast<void* const*>(enhanced_pcm_host), true, mic_order,
                        n_batches, B, L, stream);
}

// enhance.py:35-43 + the int16 writer of the dataset tools, end to end on HOST 16-bit PCM buffers: H2D of the PCM (half the
// bytes of the fp32 front door), int16 -> float / 32768 with the microphone permutation, EaBNet (+ GaGNet), float -> int16, D2H.
int eab_enhance_host_pcm16(eab_model* m, eab_model* gag, int ref_mic, const int16_t* pcm_host, const int* mic_order,
                           int16_t* enhanced_pcm_host, int B, int L, void* stream) {
    if (!m || !pcm_host || !enhanced_pcm_host) return fail("eab_enhance_host_pcm16: null argument");
    if (m->kind != 0 || (gag && gag->kind != 1)) return fail("eab_enhance_host_pcm16: needs an EaBNet handle (and optionally a GaGNet handle)");
    if (!gag) return eab_enhance_host_batches_pcm16(m, &pcm_host, mic_order, &enhanced_pcm_host, 1, B, L, stream);
    const int M = m->cfg.M;
    if (M > 64) return fail("eab_enhance_host_pcm16: at most 64 microphones");
    int order[64];
    for (int i = 0; i < M; ++i) {
        order[i] = mic_order ? mic_order[i] : i;
        if (order[i] < 0 || order[i] >= M) return fail("eab_enhance_host_pcm16: mic_order entries must be in [0, M)");
    }
    const size_t need = eab_enhance_postnet_workspace_bytes(m, gag, B, L);
    if (!need) return fail("eab_enhance_host_pcm16: bad shape");
    const size_t n_in = (size_t)B * M * L, n_out = (size_t)B * 160 * (L / 160);
    const size_t o_pcm = 0, o_epcm = align256(n_in * 2), o_ws = o_epcm + align256(n_out * 2), total = o_ws + need;
    if (m->scratch_bytes < total) {
        if (m->scratch) cudaFree(m->scratch);
        m->scratch = nullptr;
        m->scratch_bytes = 0;
        EAB_CUDA(cudaMalloc(&m->scratch, total));
        m->scratch_bytes = total;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(m->scratch);
    int16_t* d_pcm = reinterpret_cast<int16_t*>(p + o_pcm);
    int16_t* d_epcm = reinterpret_cast<int16_t*>(p + o_epcm);
    EAB_CUDA(cudaMemcpyAsync(d_pcm, pcm_host, n_in * 2, cudaMemcpyHostToDevice, st));
    EAB_TRY(enhance_postnet_dev(m, gag, ref_mic, nullptr, d_pcm, order, nullptr, d_epcm, B, L, p + o_ws, need, stream));
    EAB_CUDA(cudaMemcpyAsync(enhanced_pcm_host, d_epcm, n_out * 2, cudaMemcpyDeviceToHost, st));
    EAB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int eab_enhance_host(eab_model* m, const float* wave_host, float* enhanced_host, int B, int L, void* stream) {
    if (!m || !wave_host || !enhanced_host) return fail("eab_enhance_host: null argument");
    return eab_enhance_host_batches(m, &wave_host, &enhanced_host, 1, B, L, stream);
}

size_t eab_stream_state_bytes(const eab_model* m, int n_streams) {
    StreamLayout L;
    if (!m || stream_layout(const_cast<eab_model*>(m), n_streams, &L)) return 0;
    return L.total;
}

static int stream_check(eab_model* m, void* state, size_t state_bytes, int S, StreamLayout* L) {
    if (!m || !state) return fail("eab_stream: null argument");
    EAB_TRY(stream_layout(m, S, L));
    if (state_bytes < L->total) return fail("stream state too small: need " + std::to_string(L->total) + " bytes");
    if ((reinterpret_cast<uintptr_t>(state) & 255) != 0) return fail("stream state must be 256-byte aligned");
    return 0;
}

int eab_stream_reset(eab_model* m, void* state, size_t state_bytes, int n_streams, void* stream) {
    StreamLayout L;
    EAB_TRY(stream_check(m, state, state_bytes, n_streams, &L));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EAB_CUDA(cudaMemsetAsync(state, 0, L.total, st));
    if (!L.descs.empty()) {
        EAB_CUDA(cudaMemcpyAsync(static_cast<char*>(state) + L.off_desc, L.descs.data(), L.descs.size() * sizeof(TcmStreamDesc),
                                 cudaMemcpyHostToDevice, st));
        EAB_CUDA(cudaStreamSynchronize(st));       // the descriptor vector dies with this scope
    }
    return 0;
}

int eab_stream_step_spec(eab_model* m, void* state, size_t state_bytes, const float* frame, float* out_frame, int n_streams,
                         void* stream) {
    StreamLayout L;
    if (m && m->kind != 0) return fail("eab_stream_step_spec needs an EaBNet handle (a GaGNet steps through eab_gag_stream_step_spec)");
    EAB_TRY(stream_check(m, state, state_bytes, n_streams, &L));
    if (!frame || !out_frame) return fail("eab_stream_step_spec: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(state);
    const eab_config& c = m->cfg;
    reset_launch_count();
    {   // new spectrum frame -> slot (*step % 2) of the spectrum ring
        CombineArgs a;
        memset(&a, 0, sizeof(a));
        a.nsrc = 1;
        a.src[0].x = frame; a.src[0].C = 2 * c.M; a.src[0].xf = xform_identity(); a.src[0].RT = 1;
        a.B = n_streams; a.P = c.n_freq; a.C = 2 * c.M;
        a.out = reinterpret_cast<float*>(p + L.off_spec); a.out_RT = 2;
        a.step = reinterpret_cast<const int*>(p);
        EAB_TRY(launch_combine(a, st));
    }
    EAB_TRY(stream_forward(m, p, L, n_streams, st));
    const size_t out_bytes = (size_t)n_streams * (c.topo_type == 1 ? 2 : 2 * c.n_freq) * sizeof(float);
    EAB_CUDA(cudaMemcpyAsync(out_frame, p + L.off_out, out_bytes, cudaMemcpyDeviceToDevice, st));
    EAB_TRY(launch_step_advance(reinterpret_cast<int*>(p), st));
    m->last_launches = launch_count();
    return 0;
}

static int stream_step_any(eab_model* m, void* state, size_t state_bytes, eab_model* gag, void* gstate, size_t gstate_bytes, int ref_mic,
                           const float* hop, const int16_t* hop16, float* enhanced_hop, int16_t* enhanced16, int n_streams, void* stream) {
    StreamLayout L, GL;
    if (m && m->kind != 0) return fail("eab_stream_step needs an EaBNet handle");
    EAB_TRY(stream_check(m, state, state_bytes, n_streams, &L));
    if ((!hop && !hop16) || (!enhanced_hop && !enhanced16)) return fail("eab_stream_step: null argument");
    const eab_config& c = m->cfg;
    if (c.topo_type == 1) return fail("eab_stream_step: the 'miso' topology returns [B,2,T], which has no iSTFT");
    if (c.n_freq != 161) return fail("eab_stream_step: the 320-point STFT gives 161 bins");
    if (gag) {
        if (gag->kind != 1) return fail("eab_stream_step_postnet: the post-filter handle is not a GaGNet");
        if (gag->cfg.n_freq != c.n_freq) return fail("eab_stream_step_postnet: EaBNet and GaGNet disagree on the number of bins");
        if (ref_mic < 0 || ref_mic >= c.M) return fail("eab_stream_step_postnet: reference microphone out of range");
        EAB_TRY(stream_check(gag, gstate, gstate_bytes, n_streams, &GL));
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(state);
    int* step = reinterpret_cast<int*>(p);
    const int* start = reinterpret_cast<const int*>(p + L.off_start);
    reset_launch_count();
    EAB_TRY(launch_stft_frame(hop, hop16, reinterpret_cast<float*>(p + L.off_prev), reinterpret_cast<float*>(p + L.off_spec), 2, step,
                              start, n_streams, c.M, st));
    EAB_TRY(stream_forward(m, p, L, n_streams, st));
    const float* est = reinterpret_cast<const float*>(p + L.off_out);
    if (gag) {
        // enhance.py:49-62 frame by frame: GaGNet on (reference microphone of the compressed spectrum - read in place from the
        // spectrum ring's current slot -, EaBNet's estimate); the last module's estimate goes to the iSTFT
        char* g = static_cast<char*>(gstate);
        const long long FM2 = (long long)c.n_freq * c.M * 2;
        const long long strides[4] = {2 * FM2, 1, 0, (long long)c.M * 2};
        EAB_TRY(gag_stream_forward(gag, g, GL, reinterpret_cast<const float*>(p + L.off_spec) + ref_mic * 2, strides, 2, FM2, est, n_streams, st));
        est = reinterpret_cast<const float*>(g + GL.off_out) + (gag->gags.size() - 1) * (size_t)n_streams * 2 * c.n_freq;
        EAB_TRY(launch_step_advance(reinterpret_cast<int*>(g), st));
    }
    EAB_TRY(launch_istft_frame(est, reinterpret_cast<float*>(p + L.off_tail), enhanced_hop, enhanced16, step, start, n_streams, st));
    EAB_TRY(launch_step_advance(step, st));
    m->last_launches = launch_count();
    return 0;
}

int eab_stream_step(eab_model* m, void* state, size_t state_bytes, const float* hop, float* enhanced_hop, int n_streams,
                    void* stream) {
    if (!hop || !enhanced_hop) return fail("eab_stream_step: null argument");
    return stream_step_any(m, state, state_bytes, nullptr, nullptr, 0, 0, hop, nullptr, enhanced_hop, nullptr, n_streams, stream);
}

// the same step on the 16-bit PCM wire format: hop [S][M][160] int16 (sample / 32768), enhanced hop [S][160] int16
int eab_stream_step_pcm16(eab_model* m, void* state, size_t state_bytes, const int16_t* hop, int16_t* enhanced_hop, int n_streams,
                          void* stream) {
    if (!hop || !enhanced_hop) return fail("eab_stream_step_pcm16: null argument");
    return stream_step_any(m, state, state_bytes, nullptr, nullptr, 0, 0, nullptr, hop, nullptr, enhanced_hop, n_streams, stream);
}

// EaBNet + GaGNet post-filter, one hop per stream (enhance.py:49-62 as a causal stream).  Both states carry their own frame
// counter: reset them together (eab_stream_reset / eab_stream_reset_one on each).
int eab_stream_step_postnet(eab_model* eabnet, void* state, size_t state_bytes, eab_model* gagnet, void* gag_state, size_t gag_state_bytes,
                            int ref_mic, const float* hop, float* enhanced_hop, int n_streams, void* stream) {
    if (!hop || !enhanced_hop || !gagnet) return fail("eab_stream_step_postnet: null argument");
    return stream_step_any(eabnet, state, state_bytes, gagnet, gag_state, gag_state_bytes, ref_mic, hop, nullptr, enhanced_hop, nullptr,
                           n_streams, stream);
}
int eab_stream_step_postnet_pcm16(eab_model* eabnet, void* state, size_t state_bytes, eab_model* gagnet, void* gag_state,
                                  size_t gag_state_bytes, int ref_mic, const int16_t* hop, int16_t* enhanced_hop, int n_streams, void* stream) {
    if (!hop || !enhanced_hop || !gagnet) return fail("eab_stream_step_postnet_pcm16: null argument");
    return stream_step_any(eabnet, state, state_bytes, gagnet, gag_state, gag_state_bytes, ref_mic, nullptr, hop, nullptr, enhanced_hop,
                           n_streams, stream);
}

// GaGNet.forward one frame at a time: inpt_frame / pre_frame [S][2][F] -> the q modules' estimates [q][S][2][F]
int eab_gag_stream_step_spec(eab_model* gagnet, void* gag_state, size_t gag_state_bytes, const float* inpt_frame, const float* pre_frame,
                             float* out_frames, int n_streams, void* stream) {
    if (!gagnet || gagnet->kind != 1) return fail("eab_gag_stream_step_spec: the handle is not a GaGNet");
    if (!inpt_frame || !pre_frame || !out_frames) return fail("eab_gag_stream_step_spec: null argument");
    StreamLayout GL;
    EAB_TRY(stream_check(gagnet, gag_state, gag_state_bytes, n_streams, &GL));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* g = static_cast<char*>(gag_state);
    const int F = gagnet->cfg.n_freq;
    const long long strides[4] = {2LL * F, (long long)F, 0, 1};
    reset_launch_count();
    EAB_TRY(gag_stream_forward(gagnet, g, GL, inpt_frame, strides, 1, 0, pre_frame, n_streams, st));
    EAB_CUDA(cudaMemcpyAsync(out_frames, g + GL.off_out, gagnet->gags.size() * (size_t)n_streams * 2 * F * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
    EAB_TRY(launch_step_advance(reinterpret_cast<int*>(g), st));
    gagnet->last_launches = launch_count();
    return 0;
}

// One stream leaves and a new one joins in its slot: from the next step on stream `idx` starts over (its frame 0), the other
// streams carry on untouched.  Its history rings need no clearing: every kernel treats frames before the stream's start index
// as the literal zeros of the causal padding; the carried LSTM state is zeroed.  Stream-ordered, no synchronisation.
int eab_stream_reset_one(eab_model* m, void* state, size_t state_bytes, int n_streams, int idx, void* stream) {
    StreamLayout L;
    EAB_TRY(stream_check(m, state, state_bytes, n_streams, &L));
    if (idx < 0 || idx >= n_streams) return fail("eab_stream_reset_one: stream index out of range");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(state);
    EAB_TRY(launch_stream_restart(reinterpret_cast<int*>(p + L.off_start), reinterpret_cast<const int*>(p), idx, st));
    for (const auto& ps : L.per_stream)
        EAB_CUDA(cudaMemsetAsync(p + L.off_act + ps.first + (size_t)idx * ps.second, 0, ps.second, st));
    return 0;
}

int eab_last_launch_count(const eab_model* m) { return m ? m->last_launches : 0; }

// InstanceNorm statistics of the last eab_forward (option "norm_log" = 1 before it): the i-th normalisation layer the
// forward applied, in execution order.  sums_host [C][2] = (sum, sum of squares) of the layer's input over `count`
// positions (all batch items pooled).  Synchronises the stream; the workspace of that forward must still be intact.
int eab_norm_stats_count(const eab_model* m) { return m ? (int)m->norm_log.size() : 0; }
int eab_norm_stats(eab_model* m, int i, const char** weight_name, int* C, int64_t* count, double* sums_host, void* stream) {
    if (!m) return fail("eab_norm_stats: null handle");
    if (i < 0 || i >= (int)m->norm_log.size()) return fail("eab_norm_stats: index out of range (set option norm_log = 1 before the forward)");
    const auto& r = m->norm_log[i];
    if (weight_name) *weight_name = m->params[r.gamma].name.c_str();
    if (C) *C = r.C;
    if (count) *count = (int64_t)r.count * r.B;
    if (sums_host) {
        std::vector<double> h((size_t)r.B * r.C * 2);
        EAB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
        EAB_CUDA(cudaMemcpy(h.data(), r.stats, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int c = 0; c < r.C * 2; ++c) sums_host[c] = 0.0;
        for (int b = 0; b < r.B; ++b)
            for (int c = 0; c < r.C * 2; ++c) sums_host[c] += h[(size_t)b * r.C * 2 + c];
    }
    return 0;
}

int64_t eab_debug_tap(eab_model* m, const char* name, float* dst, int64_t capacity, void* stream) {
    if (!m || !name || !dst) { fail("eab_debug_tap: null argument"); return -1; }
    auto it = m->taps.find(name);
    if (it == m->taps.end()) { fail(std::string("no such tap: ") + name); return -1; }
    const Tap& t = it->second;
    const int64_t n = (int64_t)t.B * t.T * t.act.F * t.act.C;
    if (n > capacity) { fail("eab_debug_tap: destination too small"); return -1; }
    CombineArgs a;
    memset(&a, 0, sizeof(a));
    a.nsrc = 1;
    a.src[0].x = t.act.data; a.src[0].C = t.act.C; a.src[0].xf = t.act.xf;
    if (t.act.data2) { a.nsrc = 2; a.src[1].x = t.act.data2; a.src[1].C = t.act.C; a.src[1].xf = t.act.xf2; }
    a.B = t.B; a.P = t.T * t.act.F; a.C = t.act.C; a.out = dst;
    if (launch_combine(a, static_cast<cudaStream_t>(stream))) return -1;
    return n;
}

int eab_set_option(eab_model* m, const char* name, int value) {
    if (!m || !name) return fail("eab_set_option: null argument");
    const std::string n(name);
    if (n == "umma") m->opt_umma = value != 0;
    else if (n == "staged") m->opt_staged = value != 0;
    else if (n == "raw") m->opt_raw = value != 0;
    else if (n == "raw_grid") m->opt_raw_grid = value;
    else if (n == "lazy") m->opt_lazy = value != 0;
    else if (n == "norm_log") { m->opt_norm_log = value != 0; m->norm_log.clear(); }
    else if (n == "tcm_chain") m->opt_tcm_chain = value;
    else if (n == "host_graph") m->opt_host_graph = value != 0;
    else if (n == "dual_stream") m->opt_dual_stream = value != 0;
    else if (n == "stream_tcm") m->opt_stream_tcm = value != 0;
    else if (n == "stream_umma") m->opt_stream_umma = value != 0;
    else if (n == "stream_lstm") m->opt_stream_lstm = value != 0;
    else if (n == "stream_pair") m->opt_stream_pair = value != 0;
    else if (n == "stream_fuse") m->opt_stream_fuse = value != 0;
    else if (n == "lstm_pp") m->opt_lstm_pp = value != 0;
    else if (n == "lstm_exp") m->opt_lstm_exp = value;
    else if (n == "stft_tc") g_stft_tc = value != 0;
    else if (n == "istft_tc") g_istft_tc = value != 0;
    else if (n == "fused_head") m->opt_fused_head = value != 0;
    else if (n == "head_w_tap") m->opt_head_w_tap = value != 0;
    else if (n == "enc_passes" && (value == 1 || value == 3)) { m->opt_enc_passes = value; m->opt_inner_passes = value; }
    else if (n == "dec_passes" && (value == 1 || value == 3)) m->opt_dec_passes = value;
    else if (n == "inner_passes" && (value == 1 || value == 3)) m->opt_inner_passes = value;
    else if (n == "first_passes" && (value == 1 || value == 3)) m->opt_first_passes = value;
    else if (n == "dbg_launch") {
        m->opt_dbg_launch = value;
        if (!m->dbg_buf) { if (check_cuda(cudaMalloc(&m->dbg_buf, 16 * sizeof(unsigned long long)), "dbg alloc")) return 1; }
        cudaMemset(m->dbg_buf, 0, 16 * sizeof(unsigned long long));
    }
    else return fail("eab_set_option: unknown option or bad value: " + n);
    for (auto& sg : m->slot_graph)               // graphs captured under the old options must not be replayed
        if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
    return 0;
}

int eab_debug_counters(eab_model* m, unsigned long long* out16) {
    if (!m || !out16 || !m->dbg_buf) return fail("eab_debug_counters: not enabled");
    EAB_CUDA(cudaDeviceSynchronize());
    EAB_CUDA(cudaMemcpy(out16, m->dbg_buf, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}

int eab_profile_enable(eab_model* m, int on) {
    (void)m;
    g_prof.clear();
    g_prof.on = on != 0;
    g_prof.detail = on == 2;          // 2: one summary entry per launch instead of per kernel family
    return 0;
}

int64_t eab_profile_summary(eab_model* m, char* buf, int64_t cap) {
    (void)m;
    if (!buf || cap < 2) { fail("eab_profile_summary: bad buffer"); return -1; }
    if (check_cuda(cudaDeviceSynchronize(), "profile sync")) return -1;
    struct Agg { int n = 0; double ms = 0, flops = 0, bytes = 0, moved = 0; };
    std::map<std::string, Agg> agg;
    int seq = 0;
    for (auto& r : g_prof.recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
        char key[64];
        if (g_prof.detail) snprintf(key, sizeof(key), "%04d:%s", seq++, r.cat); else snprintf(key, sizeof(key), "%s", r.cat);
        Agg& a = agg[key];
        a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes; a.moved += r.moved;
    }
    std::string js = "[";
    bool first = true;
    for (auto& kv : agg) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s{\"kernel\":\"%s\",\"launches\":%d,\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e,\"moved_bytes\":%.6e}",
                 first ? "" : ",", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes, kv.second.moved);
        js += tmp;
        first = false;
    }
    js += "]";
    g_prof.clear();
    if ((int64_t)js.size() + 1 > cap) { fail("eab_profile_summary: buffer too small"); return -1; }
    memcpy(buf, js.c_str(), js.size() + 1);
    return (int64_t)js.size();
}

const char* eab_last_error(void) { return g_err.c_str(); }

const char* eab_build_info(void) {
#define EAB_STR2(x) #x
#define EAB_STR(x) EAB_STR2(x)
    return "sm_100a;" __DATE__ ";nvcc " EAB_STR(__CUDACC_VER_MAJOR__) "." EAB_STR(__CUDACC_VER_MINOR__);
}

}  // extern "C"
