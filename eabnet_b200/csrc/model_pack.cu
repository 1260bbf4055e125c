// Weight packing: the reference state_dict -> the kernels' layouts (one device blob per model), eab_commit_params.
#include "model_internal.h"

namespace eab {
namespace detail {

// ================================================================================================ pack
struct Packer {
    eab_model* m;
    std::vector<float> blob;
    size_t alloc(size_t n) {
        size_t off = (blob.size() + 63) / 64 * 64;
        blob.resize(off + n, 0.f);
        return off;
    }
    const std::vector<float>& P(int i) const { return m->params[i].host; }

    void normact(NormAct& na) {
        const int C = na.C;
        if (na.has_norm) {
            na.off_scale = alloc(C);
            na.off_shift = alloc(C);
            for (int c = 0; c < C; ++c) {
                if (m->cfg.norm_type == 1) {          // BatchNorm eval: fold running statistics
                    const double s = (double)P(na.gamma)[c] / sqrt((double)P(na.var)[c] + 1e-5);
                    blob[na.off_scale + c] = (float)s;
                    blob[na.off_shift + c] = (float)((double)P(na.beta)[c] - (double)P(na.mean)[c] * s);
                } else {
                    blob[na.off_scale + c] = P(na.gamma)[c];
                    blob[na.off_shift + c] = P(na.beta)[c];
                }
            }
        }
        na.off_alpha = alloc(C);
        na.alpha01 = true;
        for (int c = 0; c < C; ++c) {
            const float al = P(na.alpha)[c];
            blob[na.off_alpha + c] = al;
            if (!(al >= 0.f && al <= 1.f)) na.alpha01 = false;
        }
    }

    void conv(ConvLayer& L) {
        const int cout_t = L.gated ? 2 * L.cout : L.cout;
        L.N = L.gated ? 2 * ceil64(L.cout) : pad_n(L.cout);
        L.gate_off = L.gated ? ceil64(L.cout) : 0;
        auto col_of = [&](int n_orig) { return (L.gated && n_orig >= L.cout) ? L.gate_off + (n_orig - L.cout) : n_orig; };
        auto cin_of = [&](int cin_ref) {          // reference input channel -> memory channel
            if (!L.perm_ri) return cin_ref;
            const int ri = cin_ref / L.M, mic = cin_ref - ri * L.M;
            return mic * 2 + ri;
        };
        const std::vector<float>& W = P(L.w);
        L.nvar = L.deconv ? 2 : 1;
        for (int v = 0; v < L.nvar; ++v) {
            int nt = 0;
            std::vector<int> kj, kk;
            for (int j = 0; j < L.kt; ++j)
                for (int k = 0; k < L.kf; ++k) {
                    if (L.deconv) {
                        if ((k & 1) != v) continue;
                        L.dt[v][nt] = j;                  // transposed conv + chomp: tap j reads frame t - j
                        L.df[v][nt] = -(k / 2);           // fo = 2e + v, fi = e - (k - v)/2
                    } else {
                        L.dt[v][nt] = L.kt - 1 - j;       // top padding kt-1: tap j reads frame t - (kt-1-j)
                        L.df[v][nt] = k;                  // fi = 2 fo + k
                    }
                    kj.push_back(j); kk.push_back(k);
                    ++nt;
                }
            L.ntaps[v] = nt;
            L.off_w[v] = alloc((size_t)(nt > 0 ? nt : 1) * L.cin * L.N);
            for (int tp = 0; tp < nt; ++tp)
                for (int ci = 0; ci < L.cin; ++ci)
                    for (int n = 0; n < cout_t; ++n) {
                        const size_t src = L.deconv
                            ? (((size_t)ci * cout_t + n) * L.kt + kj[tp]) * L.kf + kk[tp]
                            : (((size_t)n * L.cin + ci) * L.kt + kj[tp]) * L.kf + kk[tp];
                        blob[L.off_w[v] + ((size_t)tp * L.cin + cin_of(ci)) * L.N + col_of(n)] = W[src];
                    }
        }
        L.off_b = alloc(L.N);
        for (int n = 0; n < cout_t; ++n) blob[L.off_b + col_of(n)] = P(L.b)[n];
        pack_umma(L, W);
        normact(L.na);
    }

    // tcgen05 weight images: [variant][tap][slab][N rows][32 k], 128B-swizzled, hi = tf32(w), lo = tf32(w - hi)
    void pack_umma(ConvLayer& L, const std::vector<float>& W) {
        const int co = L.cout;
        L.umma_ok = false;
        if (co != 16 && co != 32 && co != 64 && co != 128) return;
        if (L.gated && co > 128) return;
        L.wide = L.perm_ri;
        if (!L.wide && (L.cin % 64 != 0)) return;
        if (L.wide && L.deconv) return;
        const int cout_t = L.gated ? 2 * co : co;
        L.u_N = cout_t;
        L.u_gate_off = L.gated ? co : 0;
        L.u_kwidth = L.wide ? L.kf * L.cin : 0;
        L.u_nslab = L.wide ? (L.u_kwidth + 63) / 64 : L.cin / 64;
        auto cin_mem = [&](int cin_ref) {
            if (!L.perm_ri) return cin_ref;
            const int ri = cin_ref / L.M, mic = cin_ref - ri * L.M;
            return mic * 2 + ri;
        };
        for (int v = 0; v < L.nvar; ++v) {
            // taps: wide mode has one tap per temporal tap (window over kf positions); otherwise as the generic path
            std::vector<int> tj, tk;
            int nt = 0;
            if (L.wide) {
                for (int j = 0; j < L.kt; ++j) { L.u_dt[v][nt] = L.kt - 1 - j; L.u_df[v][nt] = 0; tj.push_back(j); tk.push_back(0); ++nt; }
            } else {
                nt = L.ntaps[v];
                int q = 0;
                for (int j = 0; j < L.kt; ++j)
                    for (int k = 0; k < L.kf; ++k) {
                        if (L.deconv && (k & 1) != v) continue;
                        L.u_dt[v][q] = L.dt[v][q]; L.u_df[v][q] = L.df[v][q];
                        tj.push_back(j); tk.push_back(k);
                        ++q;
                    }
            }
            L.u_ntaps[v] = nt;
            if (nt == 0) return;
            const size_t img = (size_t)nt * L.u_nslab * cout_t * 32;      // floats: N rows x 128 B per (tap, slab)
            L.off_whi[v] = alloc(img);
            L.off_wlo[v] = alloc(img);
            // dense [tap][kk][n] first (kk = K index inside the tap in MEMORY order), then swizzle per slab
            const int kper = L.u_nslab * 64;
            std::vector<float> dense((size_t)nt * kper * cout_t, 0.f);
            for (int tp = 0; tp < nt; ++tp)
                for (int ci = 0; ci < L.cin; ++ci)
                    for (int n = 0; n < cout_t; ++n) {
                        if (L.wide) {
                            for (int k = 0; k < L.kf; ++k) {
                                const size_t src = (((size_t)n * L.cin + ci) * L.kt + tj[tp]) * L.kf + k;
                                dense[((size_t)tp * kper + (size_t)k * L.cin + cin_mem(ci)) * cout_t + n] = W[src];
                            }
                        } else {
                            const size_t src = L.deconv ? (((size_t)ci * cout_t + n) * L.kt + tj[tp]) * L.kf + tk[tp]
                                                        : (((size_t)n * L.cin + ci) * L.kt + tj[tp]) * L.kf + tk[tp];
                            dense[((size_t)tp * kper + ci) * cout_t + n] = W[src];
                        }
                    }
            __half* img_hi = reinterpret_cast<__half*>(blob.data() + L.off_whi[v]);
            __half* img_lo = reinterpret_cast<__half*>(blob.data() + L.off_wlo[v]);
            for (int tp = 0; tp < nt; ++tp)
                for (int sl = 0; sl < L.u_nslab; ++sl) {
                    const size_t base = ((size_t)tp * L.u_nslab + sl) * cout_t * 64;       // in halves
                    for (int n = 0; n < cout_t; ++n)
                        for (int k = 0; k < 64; ++k) {
                            const float w = dense[((size_t)tp * kper + sl * 64 + k) * cout_t + n];
                            const __half hi = __float2half_rn(w);
                            img_hi[base + sw128_index_h(n, k)] = hi;
                            img_lo[base + sw128_index_h(n, k)] = __float2half_rn(w - __half2float(hi));
                        }
                }
        }
        L.off_ub = alloc(cout_t);
        for (int n = 0; n < cout_t; ++n) blob[L.off_ub + n] = P(L.b)[n];
        L.umma_ok = true;
        if (L.wide && !L.deconv && 2 * L.cin <= 64 && L.kt * ((L.kf + 1) / 2) <= kMaxTaps) {
            const int ns = (L.kf + 1) / 2;
            // kt * 2 cin <= 64 (and 8-float chunks that stay inside a frame): the kt frames of a pair are stacked along K,
            // the row carries no zero padding and only the ns frequency shifts remain as taps (half the MMAs for kt = 2)
            const bool stack = L.kt > 1 && L.kt * 2 * L.cin <= 64 && (2 * L.cin) % 8 == 0;
            L.p_stack = stack ? L.kt : 1;
            L.p_ntaps = stack ? ns : L.kt * ns;
            const size_t img = (size_t)L.p_ntaps * cout_t * 32;            // floats: one 64-wide slab per tap
            L.off_phi = alloc(img);
            L.off_plo = alloc(img);
            __half* phi = reinterpret_cast<__half*>(blob.data() + L.off_phi);
            __half* plo = reinterpret_cast<__half*>(blob.data() + L.off_plo);
            int q = 0;
            for (int j0 = 0; j0 < (stack ? 1 : L.kt); ++j0)
                for (int sft = 0; sft < ns; ++sft, ++q) {
                    L.p_dt[q] = stack ? 0 : L.kt - 1 - j0;
                    L.p_ds[q] = sft;
                    const size_t base = (size_t)q * cout_t * 64;
                    for (int n = 0; n < cout_t; ++n)
                        for (int k = 0; k < 64; ++k) {
                            const int j = stack ? k / (2 * L.cin) : j0;           // kernel row (frame t - (kt - 1 - j))
                            const int kk = stack ? k - j * 2 * L.cin : k;
                            const int pos = kk / L.cin, cm = kk - pos * L.cin;    // position inside the pair, memory channel
                            float w = 0.f;
                            if (j < L.kt && pos < 2 && 2 * sft + pos < L.kf) {
                                int ci = cm;                                      // memory channel -> reference channel
                                if (L.perm_ri) { const int mic = cm / 2, ri = cm - 2 * mic; ci = ri * L.M + mic; }
                                w = W[(((size_t)n * L.cin + ci) * L.kt + j) * L.kf + 2 * sft + pos];
                            }
                            const __half hi = __float2half_rn(w);
                            phi[base + sw128_index_h(n, k)] = hi;
                            plo[base + sw128_index_h(n, k)] = __float2half_rn(w - __half2float(hi));
                        }
                }
            L.pair_ok = true;
        }
    }

    void tcm(TcmLayer& t) {
        const eab_config& c = m->cfg;
        const int Fb = m->Fb, cd = c.cd1, df = c.d_feat, kd = c.kd1;
        const int Nin = pad_n(cd);
        // residual-stream channel of reference channel cr: bottleneck order f*64 + cc for cr = cc*Fb + f (EaBNet), or cr itself
        auto mem_ch = [&](int cr) { if (!t.perm) return cr; const int cc = cr / Fb, f = cr - cc * Fb; return f * 64 + cc; };
        // 1x1 squeeze
        t.off_in = alloc((size_t)df * Nin);
        for (int n = 0; n < cd; ++n)
            for (int cr = 0; cr < df; ++cr)
                blob[t.off_in + (size_t)mem_ch(cr) * Nin + n] = P(t.w_in)[(size_t)n * df + cr];
        if (t.single) {
            // one dilated branch, no gate (GaGNet.py:310-315)
            t.off_dil = alloc((size_t)kd * cd * Nin);
            for (int k = 0; k < kd; ++k)
                for (int ci = 0; ci < cd; ++ci)
                    for (int n = 0; n < cd; ++n)
                        blob[t.off_dil + ((size_t)k * cd + ci) * Nin + n] = P(t.w_left)[((size_t)n * cd + ci) * kd + k];
        } else {
            // dilated pair as one gated conv over K = [left-branch channels | right-branch channels]
            const int Nd = 2 * ceil64(cd), goff = ceil64(cd);
            t.off_dil = alloc((size_t)kd * 2 * cd * Nd);
            for (int k = 0; k < kd; ++k)
                for (int ci = 0; ci < cd; ++ci)
                    for (int n = 0; n < cd; ++n) {
                        blob[t.off_dil + ((size_t)k * 2 * cd + ci) * Nd + n] = P(t.w_left)[((size_t)n * cd + ci) * kd + k];
                        blob[t.off_dil + ((size_t)k * 2 * cd + cd + ci) * Nd + goff + n] = P(t.w_right)[((size_t)n * cd + ci) * kd + k];
                    }
        }
        // 1x1 expand
        t.off_out = alloc((size_t)cd * df);
        for (int nr = 0; nr < df; ++nr)
            for (int ci = 0; ci < cd; ++ci) blob[t.off_out + (size_t)ci * df + mem_ch(nr)] = P(t.w_out)[(size_t)nr * cd + ci];
        t.u_in = umma_images(t.off_in, 1, df, Nin, cd, false, 0, nullptr);
        if (t.single) t.u_dil = umma_images(t.off_dil, kd, cd, Nin, cd, false, 0, nullptr);
        else {
            t.u_dil = umma_images(t.off_dil, kd, 2 * cd, 2 * ceil64(cd), 2 * cd, true, ceil64(cd), nullptr);
            if (cd == 64) {
                // the branches separately (the merged form above is block-diagonal: half of its MMAs multiply zeros)
                for (int br = 0; br < 2; ++br) {
                    const size_t off = alloc((size_t)kd * cd * Nin);
                    const std::vector<float>& W = P(br ? t.w_right : t.w_left);
                    for (int k = 0; k < kd; ++k)
                        for (int ci = 0; ci < cd; ++ci)
                            for (int n = 0; n < cd; ++n) blob[off + ((size_t)k * cd + ci) * Nin + n] = W[((size_t)n * cd + ci) * kd + k];
                    (br ? t.u_dr : t.u_dl) = umma_images(off, kd, cd, Nin, cd, false, 0, nullptr);
                }
            }
        }
        t.u_out = umma_images(t.off_out, 1, cd, df, df, false, 0, nullptr);
        normact(t.na_left);
        if (!t.single) normact(t.na_right);
        normact(t.na_out);
    }

    // GaGNet glance / gaze input convs (GaGNet.py:161-165, 190): K = [encoder feature, bottleneck order f*64+c | pre_x row
    // ri*F+f, zero-padded to KP], columns of split sp = value channels 64 sp .. | gate channels 64 sp ..
    void gag_in(GagIn& in) {
        const eab_config& c = m->cfg;
        const int Fb = m->Fb, df = c.d_feat, Fq = c.n_freq, KP = ceil64(2 * Fq), ci = 2 * Fq + df;
        // widest gated split the tensor-core kernel takes: 128 value + 128 gate columns (each split re-reads the whole input)
        const int SW = df % 128 == 0 ? 128 : 64;
        in.K = df + KP;
        in.SW = SW;
        in.nsplit = df / SW;
        for (int sp = 0; sp < in.nsplit; ++sp) {
            in.off_dense[sp] = alloc((size_t)in.K * 2 * SW);
            std::vector<float> bias(2 * SW);
            for (int n = 0; n < 2 * SW; ++n) {
                const bool gate = n >= SW;
                const int co = sp * SW + (n % SW);
                const std::vector<float>& W = P(gate ? in.w_gate : in.w_main);
                bias[n] = P(gate ? in.b_gate : in.b_main)[co];
                for (int cr = 0; cr < df; ++cr) {
                    const int cc = cr / Fb, f = cr - cc * Fb;
                    blob[in.off_dense[sp] + (size_t)(f * 64 + cc) * 2 * SW + n] = W[(size_t)co * ci + cr];
                }
                for (int k = 0; k < 2 * Fq; ++k) blob[in.off_dense[sp] + (size_t)(df + k) * 2 * SW + n] = W[(size_t)co * ci + df + k];
            }
            in.u[sp] = umma_images(in.off_dense[sp], 1, in.K, 2 * SW, 2 * SW, true, SW, bias.data());
        }
    }

    void gag_lin(GagLin& l) {
        const eab_config& c = m->cfg;
        l.off_w = linear(l.w, l.b, c.n_freq, c.d_feat, &l.N, &l.off_b);
        std::vector<float> b0(blob.begin() + l.off_b, blob.begin() + l.off_b + l.N);
        l.u = umma_images(l.off_w, 1, c.d_feat, l.N, c.n_freq, false, 0, b0.data());
    }

    // Build fp16 hi/lo images from a dense [ntaps][K][ldn] fp32 matrix that already sits in the blob at `off`
    // (the layout of the CUDA-core path).  Columns [0, ncols) are used; gated => value|gate halves of `cout` each
    // located at columns [0,cout) and [gate_col, gate_col+cout) of the dense matrix.
    UmmaW umma_images(size_t off, int ntaps, int K, int ldn, int ncols, bool gated, int gate_col, const float* bias) {
        UmmaW u;
        if (K % 64 != 0) return u;
        const int cout = gated ? ncols / 2 : ncols;
        int padded = cout <= 16 ? 16 : cout <= 32 ? 32 : cout <= 64 ? 64 : (cout + 127) / 128 * 128;
        if (gated && padded != cout) return u;
        if (gated && cout > 128) return u;
        u.ntaps = ntaps;
        u.nslab = K / 64;
        u.gate_off = gated ? cout : 0;
        u.nsplit = (!gated && padded > 128) ? padded / 128 : 1;
        if (u.nsplit > 4) return u;
        u.cout = padded / u.nsplit;
        u.ncol = gated ? 2 * cout : u.cout;
        u.ld = padded;
        u.has_bias = bias != nullptr;
        for (int sp = 0; sp < u.nsplit; ++sp) {
            const size_t img = (size_t)ntaps * u.nslab * u.ncol * 32;
            u.off_hi[sp] = alloc(img);
            u.off_lo[sp] = alloc(img);
            u.off_bias[sp] = alloc(u.ncol);
            __half* hi = reinterpret_cast<__half*>(blob.data() + u.off_hi[sp]);
            __half* lo = reinterpret_cast<__half*>(blob.data() + u.off_lo[sp]);
            for (int n = 0; n < u.ncol; ++n) {
                // column of the dense matrix feeding image row n
                int col;
                if (gated) col = n < cout ? n : gate_col + (n - cout);
                else col = sp * u.cout + n;
                const bool real = gated ? true : col < ncols;
                blob[u.off_bias[sp] + n] = (bias && real) ? bias[col] : 0.f;
                for (int tp = 0; tp < ntaps; ++tp)
                    for (int sl = 0; sl < u.nslab; ++sl) {
                        const size_t base = ((size_t)tp * u.nslab + sl) * u.ncol * 64;
                        for (int k = 0; k < 64; ++k) {
                            const float w = real ? blob[off + ((size_t)tp * K + sl * 64 + k) * ldn + col] : 0.f;
                            const __half h = __float2half_rn(w);
                            hi[base + sw128_index_h(n, k)] = h;
                            lo[base + sw128_index_h(n, k)] = __float2half_rn(w - __half2float(h));
                        }
                    }
            }
        }
        u.ok = true;
        return u;
    }

    size_t linear(int w, int b, int nout, int nin, int* N, size_t* off_b) {
        *N = pad_n(nout);
        const size_t off = alloc((size_t)nin * *N);
        for (int n = 0; n < nout; ++n)
            for (int k = 0; k < nin; ++k) blob[off + (size_t)k * *N + n] = P(w)[(size_t)n * nin + k];
        *off_b = alloc(*N);
        for (int n = 0; n < nout; ++n) blob[*off_b + n] = P(b)[n];
        return off;
    }

    void head() {
        const eab_config& c = m->cfg;
        if (m->rnn[0][0] >= 0) {
            const int H = 64;
            for (int r = 0; r < 2; ++r) {
                const int E = r ? H : c.embed_dim;
                m->off_rnn[r][0] = alloc((size_t)E * H * 4);
                m->off_rnn[r][1] = alloc((size_t)H * H * 4);
                m->off_rnn[r][2] = alloc((size_t)H * 4);
                for (int g = 0; g < 4; ++g)
                    for (int j = 0; j < H; ++j) {
                        for (int k = 0; k < E; ++k)
                            blob[m->off_rnn[r][0] + ((size_t)k * H + j) * 4 + g] = P(m->rnn[r][0])[(size_t)(g * H + j) * E + k];
                        for (int k = 0; k < H; ++k)
                            blob[m->off_rnn[r][1] + ((size_t)k * H + j) * 4 + g] = P(m->rnn[r][1])[(size_t)(g * H + j) * H + k];
                        blob[m->off_rnn[r][2] + (size_t)j * 4 + g] = P(m->rnn[r][2])[g * H + j] + P(m->rnn[r][3])[g * H + j];
                    }
            }
            // tcgen05 LSTM: image row n = half*128 + quarter*32 + gate*8 + j  <->  torch row gate*64 + (quarter*16 + half*8 + j);
            // K slab 0 = W_ih (input channels), slab 1 = W_hh
            m->rnn_umma_ok = c.embed_dim == 64;
            if (m->rnn_umma_ok) {
                for (int r = 0; r < 2; ++r) {
                    m->off_rnn_img[r] = alloc((size_t)4 * 256 * 32);
                    m->off_rnn_ubias[r] = alloc(256);
                    __half* img = reinterpret_cast<__half*>(blob.data() + m->off_rnn_img[r]);
                    for (int n = 0; n < 256; ++n) {
                        const int hf_ = n >> 7, qtr_ = (n >> 5) & 3, g = (n >> 3) & 3, jj = n & 7;
                        const int row = g * H + qtr_ * 16 + hf_ * 8 + jj;
                        blob[m->off_rnn_ubias[r] + n] = P(m->rnn[r][2])[row] + P(m->rnn[r][3])[row];
                        for (int slab = 0; slab < 2; ++slab)
                            for (int k = 0; k < 64; ++k) {
                                const float w = P(m->rnn[r][slab])[(size_t)row * 64 + k];
                                const __half hi = __float2half_rn(w);
                                const __half lo = __float2half_rn(w - __half2float(hi));
                                img[((size_t)(0 * 2 + slab) * 256) * 64 + sw128_index_h(n, k)] = hi;
                                img[((size_t)(1 * 2 + slab) * 256) * 64 + sw128_index_h(n, k)] = lo;
                            }
                    }
                }
            }
            // streaming step on the tensor cores: the gate GEMM [x_t, h_{t-1}] (K = 64 + 64) x dense [128][256] (column = torch
            // gate row g*64 + j), bias b_ih + b_hh
            if (m->rnn_umma_ok) {
                for (int r = 0; r < 2; ++r) {
                    m->off_rnn_step[r] = alloc((size_t)128 * 256);
                    std::vector<float> bsum(256);
                    for (int n = 0; n < 256; ++n) {
                        bsum[n] = P(m->rnn[r][2])[n] + P(m->rnn[r][3])[n];
                        for (int k = 0; k < 64; ++k) {
                            blob[m->off_rnn_step[r] + (size_t)k * 256 + n] = P(m->rnn[r][0])[(size_t)n * 64 + k];
                            blob[m->off_rnn_step[r] + (size_t)(64 + k) * 256 + n] = P(m->rnn[r][1])[(size_t)n * 64 + k];
                        }
                    }
                    m->off_rnn_step_b[r] = alloc(256);
                    for (int n = 0; n < 256; ++n) blob[m->off_rnn_step_b[r] + n] = bsum[n];
                    m->u_rnn_step[r] = umma_images(m->off_rnn_step[r], 1, 128, 256, 256, false, 0, bsum.data());
                }
            }
            m->off_dnn_w[0] = linear(m->dnn_w[0], m->dnn_b[0], H, H, &m->dnn_N[0], &m->off_dnn_b[0]);
            m->off_dnn_w[1] = linear(m->dnn_w[1], m->dnn_b[1], 2 * c.M, H, &m->dnn_N[1], &m->off_dnn_b[1]);
            {
                std::vector<float> b0(blob.begin() + m->off_dnn_b[0], blob.begin() + m->off_dnn_b[0] + m->dnn_N[0]);
                std::vector<float> b1(blob.begin() + m->off_dnn_b[1], blob.begin() + m->off_dnn_b[1] + m->dnn_N[1]);
                m->u_dnn[0] = umma_images(m->off_dnn_w[0], 1, H, m->dnn_N[0], H, false, 0, b0.data());
                m->u_dnn[1] = umma_images(m->off_dnn_w[1], 1, H, m->dnn_N[1], 2 * c.M, false, 0, b1.data());
            }
            m->off_ln_g = alloc(c.embed_dim);
            m->off_ln_b = alloc(c.embed_dim);
            for (int i = 0; i < c.embed_dim; ++i) {
                blob[m->off_ln_g + i] = P(m->ln_g)[i];
                blob[m->off_ln_b + i] = P(m->ln_b)[i];
            }
        } else {
            const int n = c.topo_type == 0 ? 2 * c.M : 2;
            m->off_cnn_w = linear(m->cnn_w, m->cnn_b, n, c.embed_dim, &m->cnn_N, &m->off_cnn_b);
            {
                std::vector<float> b0(blob.begin() + m->off_cnn_b, blob.begin() + m->off_cnn_b + m->cnn_N);
                m->u_cnn = umma_images(m->off_cnn_w, 1, c.embed_dim, m->cnn_N, n, false, 0, b0.data());
            }
        }
    }
};

int commit(eab_model* m, cudaStream_t st) {
    for (const Param& p : m->params)
        if (!p.set && p.kind != EAB_P_BN_COUNT) return fail("parameter not set: " + p.name);
    Packer pk{m};
    for (auto& U : m->en_mod) { pk.conv(U.in_conv); for (auto& L : U.enco) pk.conv(L); for (auto& L : U.deco) pk.conv(L); }
    for (auto& U : m->de_mod) { pk.conv(U.in_conv); for (auto& L : U.enco) pk.conv(L); for (auto& L : U.deco) pk.conv(L); }
    if (m->cfg.is_u2) { pk.conv(m->en_last); if (m->kind == 0) pk.conv(m->de_last); }
    for (auto& L : m->en_plain) pk.conv(L);
    for (auto& L : m->de_plain) pk.conv(L);
    for (auto& t : m->tcms) pk.tcm(t);
    if (m->kind == 0) pk.head();
    for (auto& G : m->gags) {
        pk.gag_in(G.in_g);
        pk.gag_in(G.in_z);
        for (auto& t : G.tcn_g) pk.tcm(t);
        for (auto& t : G.tcm_r) pk.tcm(t);
        for (auto& t : G.tcm_i) pk.tcm(t);
        pk.gag_lin(G.lin_g);
        pk.gag_lin(G.lin_r);
        pk.gag_lin(G.lin_i);
    }
    if (m->blob && m->blob_floats < pk.blob.size()) { cudaFree(m->blob); m->blob = nullptr; }
    if (!m->blob) {
        EAB_CUDA(cudaMalloc(&m->blob, pk.blob.size() * sizeof(float)));
        m->blob_floats = pk.blob.size();
    }
    EAB_CUDA(cudaMemcpyAsync(m->blob, pk.blob.data(), pk.blob.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    EAB_CUDA(cudaStreamSynchronize(st));        // the staging vector dies with this scope
    m->dirty = false;
    ++m->param_version;
    return 0;
}


}  // namespace detail
}  // namespace eab
