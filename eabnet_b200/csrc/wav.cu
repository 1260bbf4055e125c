// I/O edges of the reference's enhance.py: the wav container and the sample-rate conversion.
//
//   enhance.py:35     noisy, sr = torchaudio.load(path)         -> eab_wav_info / eab_wav_decode (planar float32 [C][frames],
//                                                                  integer PCM scaled like torchaudio: int16 / 32768 ...)
//   enhance.py:36-37  torchaudio.transforms.Resample(sr, 16000) -> eab_resample (sinc interpolation with a Hann window,
//                                                                  lowpass_filter_width 6, rolloff 0.99: torchaudio's defaults)
//   enhance.py:63     wavfile.write(path, 16000, esti_wav[0])   -> eab_wav_encode (scipy's writer, byte for byte: a float32
//                                                                  array becomes a WAVE_FORMAT_IEEE_FLOAT file with a fact chunk)
//
// The container code is host C++ (a few hundred bytes of header around a memcpy); the resampler is a CUDA kernel: a
// polyphase FIR, out[n new + j] = sum_k kernel[j][k] x[n orig + k - width], one thread per output sample with the phase
// table stored tap-major so the lanes of a warp read consecutive table entries and share their input window.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/eabnet_b200.h"
#include "common.cuh"

namespace eab {

namespace {

inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline void wr16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
inline void wr32(uint8_t* p, uint32_t v) { wr16(p, v & 0xffffu); wr16(p + 2, v >> 16); }

struct WavFmt {
    int tag = 0, channels = 0, rate = 0, bits = 0, block = 0;
    const uint8_t* data = nullptr;
    uint64_t data_bytes = 0;
};

// RIFF/WAVE chunk walk: `fmt ` (PCM, IEEE float, or WAVE_FORMAT_EXTENSIBLE wrapping one of the two) and `data`; every other
// chunk (LIST, fact, ...) is skipped, chunks are padded to even sizes
int parse_wav(const uint8_t* b, size_t n, WavFmt* f) {
    if (!b || n < 12 || memcmp(b, "RIFF", 4) != 0 || memcmp(b + 8, "WAVE", 4) != 0) return fail("wav: not a RIFF/WAVE file");
    size_t pos = 12;
    bool have_fmt = false;
    while (pos + 8 <= n) {
        const uint8_t* ck = b + pos;
        const uint64_t sz = rd32(ck + 4);
        const size_t body = pos + 8;
        if (memcmp(ck, "fmt ", 4) == 0) {
            if (sz < 16 || body + 16 > n) return fail("wav: truncated fmt chunk");
            f->tag = rd16(b + body); f->channels = rd16(b + body + 2); f->rate = (int)rd32(b + body + 4);
            f->block = rd16(b + body + 12); f->bits = rd16(b + body + 14);
            if (f->tag == 0xFFFE) {                          // extensible: the sub-format GUID starts with the real tag
                if (sz < 40 || body + 40 > n) return fail("wav: truncated extensible fmt chunk");
                f->tag = rd16(b + body + 24);
            }
            have_fmt = true;
        } else if (memcmp(ck, "data", 4) == 0) {
            if (!have_fmt) return fail("wav: data chunk before fmt chunk");
            f->data = b + body;
            f->data_bytes = std::min<uint64_t>(sz, n - body);        // a streamed file may carry 0xFFFFFFFF here
            break;
        }
        pos = body + (size_t)sz + (sz & 1);
    }
    if (!have_fmt || !f->data) return fail("wav: no fmt / data chunk");
    if (f->channels < 1 || f->rate < 1) return fail("wav: bad channel count / sample rate");
    const bool pcm = f->tag == 1 && (f->bits == 8 || f->bits == 16 || f->bits == 24 || f->bits == 32);
    const bool flt = f->tag == 3 && (f->bits == 32 || f->bits == 64);
    if (!pcm && !flt) return fail("wav: unsupported encoding (format tag " + std::to_string(f->tag) + ", " + std::to_string(f->bits) + " bits)");
    if (f->block != f->channels * f->bits / 8) return fail("wav: block alignment does not match channels x bits");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------- resampler
struct ResampleTable {
    int orig = 0, neu = 0, width = 0, K = 0;
    float* dev = nullptr;                                    // [K][neu] (tap-major)
};
std::mutex g_rs_mu;
std::map<std::tuple<int, int, int>, ResampleTable> g_rs;    // (device, orig / gcd, new / gcd): never freed (graphs may hold it)

// torchaudio.functional._get_sinc_resample_kernel, resampling_method "sinc_interp_hann", evaluated in float64 and rounded
// to float32 at the end exactly like torchaudio does when no dtype is given
void build_kernel(int orig, int neu, std::vector<float>* tab, int* width_out) {
    const int lowpass_filter_width = 6;
    const double rolloff = 0.99;
    const double pi = 3.14159265358979323846;
    const double base_freq = std::min(orig, neu) * rolloff;
    const int width = (int)std::ceil(lowpass_filter_width * (double)orig / base_freq);
    const int K = 2 * width + orig;
    tab->assign((size_t)K * neu, 0.f);
    const double scale = base_freq / orig;
    for (int j = 0; j < neu; ++j)
        for (int k = 0; k < K; ++k) {
            double t = (-(double)j / neu + (double)(k - width) / orig) * base_freq;
            t = std::max(-(double)lowpass_filter_width, std::min((double)lowpass_filter_width, t));
            const double c = std::cos(t * pi / lowpass_filter_width / 2);
            const double window = c * c;
            t *= pi;
            const double s = t == 0.0 ? 1.0 : std::sin(t) / t;
            (*tab)[(size_t)k * neu + j] = (float)(s * window * scale);
        }
    *width_out = width;
}

__global__ void resample_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ tab, int rows,
                                long long L_in, long long L_out, int orig, int neu, int width, int K) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (o >= L_out || row >= rows) return;
    const long long n = o / neu;
    const int j = (int)(o - n * neu);
    const float* xr = x + (size_t)row * L_in;
    const long long i0 = n * orig - width;                   // first input sample of the window (zero padding outside)
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        const long long i = i0 + k;
        const float v = (i >= 0 && i < L_in) ? __ldg(xr + i) : 0.f;
        acc = fmaf(v, __ldg(tab + (size_t)k * neu + j), acc);
    }
    y[(size_t)row * L_out + o] = acc;
}

}  // namespace

}  // namespace eab

using namespace eab;

extern "C" {

int eab_wav_info(const void* bytes, size_t n, int* channels, int* sample_rate, int64_t* frames, int* bits_per_sample, int* is_float) {
    WavFmt f;
    EAB_TRY(parse_wav(static_cast<const uint8_t*>(bytes), n, &f));
    if (channels) *channels = f.channels;
    if (sample_rate) *sample_rate = f.rate;
    if (frames) *frames = (int64_t)(f.data_bytes / (uint64_t)f.block);
    if (bits_per_sample) *bits_per_sample = f.bits;
    if (is_float) *is_float = f.tag == 3;
    return 0;
}

int eab_wav_decode(const void* bytes, size_t n, float* planar, int16_t* planar_pcm16) {
    WavFmt f;
    EAB_TRY(parse_wav(static_cast<const uint8_t*>(bytes), n, &f));
    if (!planar && !planar_pcm16) return fail("eab_wav_decode: null output");
    if (planar_pcm16 && !(f.tag == 1 && f.bits == 16)) return fail("eab_wav_decode: the int16 output needs a 16-bit PCM file");
    const int64_t frames = (int64_t)(f.data_bytes / (uint64_t)f.block);
    const int C = f.channels, bps = f.bits / 8;
    for (int64_t i = 0; i < frames; ++i)
        for (int c = 0; c < C; ++c) {
            const uint8_t* p = f.data + ((size_t)i * C + c) * bps;
            float v = 0.f;
            if (f.tag == 3) {
                if (f.bits == 32) { float t; memcpy(&t, p, 4); v = t; }
                else { double t; memcpy(&t, p, 8); v = (float)t; }
            } else if (f.bits == 16) {
                const int16_t s = (int16_t)rd16(p);
                if (planar_pcm16) planar_pcm16[(size_t)c * frames + i] = s;
                v = (float)s * (1.f / 32768.f);
            } else if (f.bits == 8) {
                v = ((float)p[0] - 128.f) * (1.f / 128.f);
            } else if (f.bits == 24) {
                const int32_t s = (int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24);     // sign in the top byte
                v = (float)s * (1.f / 2147483648.f);
            } else {
                v = (float)(int32_t)rd32(p) * (1.f / 2147483648.f);
            }
            if (planar) planar[(size_t)c * frames + i] = v;
        }
    return 0;
}

size_t eab_wav_encode_bytes(int64_t frames, int channels, int pcm16) {
    if (frames < 0 || channels < 1) return 0;
    const size_t data = (size_t)frames * channels * (pcm16 ? 2 : 4);
    return (pcm16 ? 44 : 58) + data + (data & 1);
}

int eab_wav_encode(const float* interleaved, const int16_t* interleaved_pcm16, int64_t frames, int channels, int sample_rate,
                   void* out, size_t capacity) {
    const bool pcm = interleaved_pcm16 != nullptr;
    if ((!interleaved && !pcm) || !out) return fail("eab_wav_encode: null argument");
    if (frames < 0 || channels < 1 || sample_rate < 1) return fail("eab_wav_encode: bad shape");
    const size_t total = eab_wav_encode_bytes(frames, channels, pcm);
    if (capacity < total) return fail("eab_wav_encode: output buffer too small");
    if (total - 8 > 0xFFFFFFFFull) return fail("eab_wav_encode: data exceeds the 4 GB RIFF limit");
    uint8_t* b = static_cast<uint8_t*>(out);
    const int bps = pcm ? 2 : 4;
    const size_t data = (size_t)frames * channels * bps;
    size_t p = 0;
    memcpy(b, "RIFF", 4); wr32(b + 4, (uint32_t)(total - 8)); memcpy(b + 8, "WAVEfmt ", 8); p = 16;
    wr32(b + p, pcm ? 16u : 18u); p += 4;
    wr16(b + p, pcm ? 1u : 3u); wr16(b + p + 2, (uint32_t)channels); wr32(b + p + 4, (uint32_t)sample_rate);
    wr32(b + p + 8, (uint32_t)(sample_rate * channels * bps)); wr16(b + p + 12, (uint32_t)(channels * bps)); wr16(b + p + 14, (uint32_t)(bps * 8));
    p += 16;
    if (!pcm) {                                              // non-PCM: cbSize = 0 and a fact chunk with the frame count
        wr16(b + p, 0); p += 2;
        memcpy(b + p, "fact", 4); wr32(b + p + 4, 4); wr32(b + p + 8, (uint32_t)frames); p += 12;
    }
    memcpy(b + p, "data", 4); wr32(b + p + 4, (uint32_t)data); p += 8;
    memcpy(b + p, pcm ? static_cast<const void*>(interleaved_pcm16) : static_cast<const void*>(interleaved), data); p += data;
    if (data & 1) b[p++] = 0;
    return 0;
}

int64_t eab_resample_length(int64_t length, int orig_freq, int new_freq) {
    if (length < 0 || orig_freq < 1 || new_freq < 1) return -1;
    const int g = std::gcd(orig_freq, new_freq);
    const long long o = orig_freq / g, nw = new_freq / g;
    return (int64_t)((nw * (long long)length + o - 1) / o);  // ceil(new * length / orig), torchaudio's target_length
}

int eab_resample(const float* wave_dev, float* out_dev, int rows, int64_t length, int orig_freq, int new_freq, void* stream) {
    if (rows < 0 || length < 0 || orig_freq < 1 || new_freq < 1) return fail("eab_resample: bad shape / rates");
    if (rows > 65535) return fail("eab_resample: more than 65535 rows");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t L_out = eab_resample_length(length, orig_freq, new_freq);
    if (rows == 0 || L_out == 0) return 0;                   // (empty tensors have null pointers)
    if (!wave_dev || !out_dev) return fail("eab_resample: null argument");
    if (orig_freq == new_freq) {                             // Resample.forward returns the input untouched
        EAB_CUDA(cudaMemcpyAsync(out_dev, wave_dev, (size_t)rows * length * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const int g = std::gcd(orig_freq, new_freq);
    const int orig = orig_freq / g, neu = new_freq / g;
    int dev = 0;
    EAB_CUDA(cudaGetDevice(&dev));
    ResampleTable t;
    {
        std::lock_guard<std::mutex> lk(g_rs_mu);
        auto key = std::make_tuple(dev, orig, neu);
        auto it = g_rs.find(key);
        if (it == g_rs.end()) {
            std::vector<float> tab;
            ResampleTable nt;
            nt.orig = orig; nt.neu = neu;
            build_kernel(orig, neu, &tab, &nt.width);
            nt.K = 2 * nt.width + orig;
            EAB_CUDA(cudaMalloc(&nt.dev, tab.size() * sizeof(float)));
            EAB_CUDA(cudaMemcpy(nt.dev, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
            it = g_rs.emplace(key, nt).first;
        }
        t = it->second;
    }
    const int threads = 256;
    const long long blocks = (L_out + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return fail("eab_resample: signal too long");
    resample_kernel<<<dim3((unsigned)blocks, (unsigned)rows), threads, 0, st>>>(wave_dev, out_dev, t.dev, rows, (long long)length, (long long)L_out,
                                                                                 t.orig, t.neu, t.width, t.K);
    EAB_LAUNCH_CHECK("resample_kernel");
    return 0;
}

}  // extern "C"
