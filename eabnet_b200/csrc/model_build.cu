// Architecture walk: EaBNet.__init__ / GaGNet.__init__ restated as layer descriptors that declare the reference state_dict
// entries in the reference's registration order (EaBNet.py:9-125, 157-624; GaGNet.py:5-326).  See model.cu for the overview.
#include "model_internal.h"

namespace eab {
namespace detail {

// ================================================================================================ declare
struct Builder {
    eab_model* m;
    int add(const std::string& name, std::initializer_list<int64_t> shape, int kind, int fan_in) {
        Param p;
        p.name = name;
        p.ndim = (int)shape.size();
        int i = 0;
        for (auto s : shape) p.shape[i++] = s;
        p.kind = kind;
        p.fan_in = fan_in;
        m->index[name] = (int)m->params.size();
        m->params.push_back(p);
        return (int)m->params.size() - 1;
    }
    void norm(const std::string& pfx, int C, NormAct& na) {
        na.has_norm = true;
        na.C = C;
        na.gamma = add(pfx + ".norm.weight", {C}, EAB_P_NORM_G, C);
        na.beta = add(pfx + ".norm.bias", {C}, EAB_P_NORM_B, C);
        if (m->cfg.norm_type == 1) {
            na.mean = add(pfx + ".norm.running_mean", {C}, EAB_P_BN_MEAN, C);
            na.var = add(pfx + ".norm.running_var", {C}, EAB_P_BN_VAR, C);
            add(pfx + ".norm.num_batches_tracked", {}, EAB_P_BN_COUNT, 1);
        }
    }
    // Sequential(gated (de)conv, [norm], PReLU)   (EaBNet.py:185-189, 214-231, 267-271, 351-358)
    ConvLayer gated(const std::string& pfx, int cin, int cout, int kt, int kf, bool deconv, bool with_norm) {
        ConvLayer L;
        L.cin = cin; L.cout = cout; L.kt = kt; L.kf = kf; L.deconv = deconv; L.gated = true;
        const std::string sub = kt > 1 ? (deconv ? ".0.conv.0" : ".0.conv.1") : ".0.conv";
        const int fan = (deconv ? 2 * cout : cin) * kt * kf;      // torch: weight.size(1) * receptive field
        if (deconv) L.w = add(pfx + sub + ".weight", {cin, 2 * cout, kt, kf}, EAB_P_CONV_W, fan);
        else        L.w = add(pfx + sub + ".weight", {2 * cout, cin, kt, kf}, EAB_P_CONV_W, fan);
        L.b = add(pfx + sub + ".bias", {2 * cout}, EAB_P_CONV_B, fan);
        L.na.C = cout;
        if (with_norm) {
            norm(pfx + ".1", cout, L.na);
            L.na.alpha = add(pfx + ".2.weight", {cout}, EAB_P_PRELU, cout);
        } else {
            L.na.alpha = add(pfx + ".1.weight", {cout}, EAB_P_PRELU, cout);
        }
        return L;
    }
    // Conv2dunit / Deconv2dunit (EaBNet.py:391-431)
    ConvLayer unit(const std::string& pfx, int cin, int cout, int kt, int kf, bool deconv) {
        ConvLayer L;
        L.cin = cin; L.cout = cout; L.kt = kt; L.kf = kf; L.deconv = deconv; L.gated = false;
        const int fan = (deconv ? cout : cin) * kt * kf;
        if (deconv) L.w = add(pfx + ".0.weight", {cin, cout, kt, kf}, EAB_P_CONV_W, fan);
        else        L.w = add(pfx + ".0.weight", {cout, cin, kt, kf}, EAB_P_CONV_W, fan);
        L.b = add(pfx + ".0.bias", {cout}, EAB_P_CONV_B, fan);
        norm(pfx + ".1", cout, L.na);
        L.na.alpha = add(pfx + ".2.weight", {cout}, EAB_P_PRELU, cout);
        return L;
    }
    UnetModule module(const std::string& pfx, int cin, int kt, int kf, int scale, bool deconv) {
        const eab_config& c = m->cfg;
        UnetModule U;
        U.in_conv = gated(pfx + ".in_conv", cin, c.c, kt, kf, deconv, true);
        for (int i = 0; i < scale; ++i)
            U.enco.push_back(unit(pfx + ".enco." + std::to_string(i) + ".conv", c.c, c.c, c.k2_t, c.k2_f, false));
        for (int i = 0; i < scale; ++i) {
            const int cin_d = (i == 0 || c.intra_connect == 1) ? c.c : 2 * c.c;
            U.deco.push_back(unit(pfx + ".deco." + std::to_string(i) + ".deconv", cin_d, c.c, c.k2_t, c.k2_f, true));
        }
        return U;
    }
};


int build(eab_model* m) {
    const eab_config& c = m->cfg;
    if (c.c < 1 || c.c > 128) return fail("c must be in 1..128");
    if (c.embed_dim < 1 || c.embed_dim > 128) return fail("embed_dim must be in 1..128");
    if (c.M < 1 || c.M > 64) return fail("M must be in 1..64");
    if (c.kd1 < 1 || c.kd1 > kMaxTaps) return fail("kd1 must be in 1..16");
    if (c.k1_t < 1 || c.k1_t > 2 || c.k2_t < 1 || c.k2_t > 2) return fail("temporal kernel sizes above 2 are not supported");
    if (c.k1_t * c.k1_f > kMaxTaps || c.k2_t * c.k2_f > kMaxTaps || c.k1_f < 1 || c.k2_f < 1) return fail("kernel too large");
    if (c.cd1 < 1 || c.cd1 > 128) return fail("cd1 must be in 1..128");
    if (c.p < 1 || c.q < 1 || c.q > 3) return fail("p >= 1 and 1 <= q <= 3 required");
    if (c.norm_type != 0 && c.norm_type != 1)
        return fail("norm_type 'cLN' cannot be constructed in the reference either (EaBNet.py:689,691)");
    Builder bd{m};
    // encoder F chain
    m->Fchain.clear();
    m->Fchain.push_back(c.n_freq);
    {
        int F = conv_out_f(c.n_freq, 5);
        m->Fchain.push_back(F);
        for (int i = 0; i < 4; ++i) { F = F > 0 ? conv_out_f(F, c.k1_f) : -1; m->Fchain.push_back(F); }
        if (F < 1) return fail("n_freq too small for the five stride-2 encoder stages");
        m->Fb = F;
    }
    if (c.d_feat != 64 * m->Fb)
        return fail("d_feat must equal 64 * bottleneck_F (the reference fails at run time otherwise, EaBNet.py:100,549)");
    if (c.is_u2) {
        m->en_mod.push_back(bd.module("en.meta_unet_list.0", 2 * c.M, 2, 5, 4, false));
        for (int i = 1; i < 4; ++i)
            m->en_mod.push_back(bd.module("en.meta_unet_list." + std::to_string(i), c.c, c.k1_t, c.k1_f, 4 - i, false));
        m->en_mod[0].in_conv.perm_ri = true;
        m->en_mod[0].in_conv.M = c.M;
        m->en_mod[0].in_conv.zone = 3;
        m->en_last = bd.gated("en.last_conv", c.c, 64, c.k1_t, c.k1_f, false, true);
        m->de_mod.push_back(bd.module("de.meta_unet_list.0", 128, c.k1_t, c.k1_f, 1, true));
        for (int i = 1; i < 4; ++i)
            m->de_mod.push_back(bd.module("de.meta_unet_list." + std::to_string(i), 2 * c.c, c.k1_t, c.k1_f, i + 1, true));
        m->de_last = bd.gated("de.last_conv", 2 * c.c, c.embed_dim, 2, 5, true, true);
        for (auto& U : m->de_mod) { U.in_conv.zone = 1; for (auto& L : U.enco) L.zone = 1; for (auto& L : U.deco) L.zone = 1; }
        m->de_last.zone = 1;
        // zone 2: the inner U-Nets of the encoder modules (their result is the residual branch of x0 + y)
        for (auto& U : m->en_mod) { for (auto& L : U.enco) L.zone = 2; for (auto& L : U.deco) L.zone = 2; }
    } else {
        m->en_plain.push_back(bd.gated("en.unet_list.0", 2 * c.M, c.c, 2, 5, false, true));
        m->en_plain[0].perm_ri = true;
        m->en_plain[0].M = c.M;
        m->en_plain.push_back(bd.gated("en.unet_list.1", c.c, c.c, c.k1_t, c.k1_f, false, false));
        m->en_plain.push_back(bd.gated("en.unet_list.2", c.c, c.c, c.k1_t, c.k1_f, false, false));
        m->en_plain.push_back(bd.gated("en.unet_list.3", c.c, c.c, c.k1_t, c.k1_f, false, true));
        m->en_plain.push_back(bd.gated("en.unet_list.4", c.c, 64, c.k1_t, c.k1_f, false, true));
        m->de_plain.push_back(bd.gated("de.unet_list.0", 128, c.c, c.k1_t, c.k1_f, true, true));
        for (int i = 1; i < 4; ++i)
            m->de_plain.push_back(bd.gated("de.unet_list." + std::to_string(i), 2 * c.c, c.c, c.k1_t, c.k1_f, true, true));
        m->de_plain.push_back(bd.gated("de.unet_list.4", 2 * c.c, c.embed_dim, 2, 5, true, true));
        for (auto& L : m->de_plain) L.zone = 1;
    }
    // head (EaBNet.py:75-81, 581-598)
    if (c.topo_type == 0 && c.bf_type == 0) {
        const int H = 64;
        for (int r = 0; r < 2; ++r) {
            const std::string p = std::string("bf_map.rnn") + (r ? "2" : "1");
            const int cin = r ? H : c.embed_dim;
            m->rnn[r][0] = bd.add(p + ".weight_ih_l0", {4 * H, cin}, EAB_P_LSTM, H);
            m->rnn[r][1] = bd.add(p + ".weight_hh_l0", {4 * H, H}, EAB_P_LSTM, H);
            m->rnn[r][2] = bd.add(p + ".bias_ih_l0", {4 * H}, EAB_P_LSTM, H);
            m->rnn[r][3] = bd.add(p + ".bias_hh_l0", {4 * H}, EAB_P_LSTM, H);
        }
        m->dnn_w[0] = bd.add("bf_map.w_dnn.0.weight", {H, H}, EAB_P_LIN_W, H);
        m->dnn_b[0] = bd.add("bf_map.w_dnn.0.bias", {H}, EAB_P_LIN_B, H);
        m->dnn_w[1] = bd.add("bf_map.w_dnn.2.weight", {2 * c.M, H}, EAB_P_LIN_W, H);
        m->dnn_b[1] = bd.add("bf_map.w_dnn.2.bias", {2 * c.M}, EAB_P_LIN_B, H);
        m->ln_g = bd.add("bf_map.norm.weight", {c.embed_dim}, EAB_P_NORM_G, c.embed_dim);
        m->ln_b = bd.add("bf_map.norm.bias", {c.embed_dim}, EAB_P_NORM_B, c.embed_dim);
    } else {
        const int n = c.topo_type == 0 ? 2 * c.M : 2;
        m->cnn_w = bd.add("bf_map.weight", {n, c.embed_dim, 1, 1}, EAB_P_CONV_W, c.embed_dim);
        m->cnn_b = bd.add("bf_map.bias", {n}, EAB_P_CONV_B, c.embed_dim);
    }
    // squeezed TCMs (EaBNet.py:83-86, 506-571)
    for (int g = 0; g < c.q; ++g)
        for (int i = 0; i < c.p; ++i) {
            if (i > 24) return fail("p too large (dilation 2^i overflows)");
            TcmLayer t;
            t.dilation = 1 << i;
            const std::string p = "stcns." + std::to_string(g) + ".tcm_list." + std::to_string(i);
            t.w_in = bd.add(p + ".in_conv.weight", {c.cd1, c.d_feat, 1}, EAB_P_CONV_W, c.d_feat);
            t.na_left.C = t.na_right.C = t.na_out.C = c.cd1;
            t.na_left.alpha = bd.add(p + ".left_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
            bd.norm(p + ".left_conv.1", c.cd1, t.na_left);
            t.w_left = bd.add(p + ".left_conv.3.weight", {c.cd1, c.cd1, c.kd1}, EAB_P_CONV_W, c.cd1 * c.kd1);
            t.na_right.alpha = bd.add(p + ".right_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
            bd.norm(p + ".right_conv.1", c.cd1, t.na_right);
            t.w_right = bd.add(p + ".right_conv.3.weight", {c.cd1, c.cd1, c.kd1}, EAB_P_CONV_W, c.cd1 * c.kd1);
            t.na_out.alpha = bd.add(p + ".out_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
            bd.norm(p + ".out_conv.1", c.cd1, t.na_out);
            t.w_out = bd.add(p + ".out_conv.2.weight", {c.d_feat, c.cd1, 1}, EAB_P_CONV_W, c.cd1);
            const int span = (c.kd1 - 1) * t.dilation;
            if (!c.is_causal && (span & 1)) return fail("non-causal TCM needs an even (kd1-1)*dilation");
            const int pad_left = c.is_causal ? span : span / 2;
            for (int k = 0; k < c.kd1; ++k) t.dt[k] = pad_left - k * t.dilation;
            m->tcms.push_back(t);
        }
    return 0;
}

// GaGNet.__init__ (GaGNet.py:69-73): encoder on cat(inpt, pre_x), then q glance-gaze modules.  m->cfg carries the
// settings the shared builder / packer / runner code reads (M = cin makes the first layer's 2M input channels the
// reference's cin*2 and its (ri, m) -> (m, ri) weight permutation the one gag_pack_kernel's channel order needs).
int build_gag(eab_model* m) {
    const eab_gag_config& g = m->gcfg;
    eab_config& c = m->cfg;
    if (g.cin != 2) return fail("GaGNet: cin must be 2 (the reference's glance / gaze 1x1 convs take 2*(fft_num/2+1) + d_feat channels, GaGNet.py:160,224)");
    if (g.n_dilas < 1 || g.n_dilas > 8) return fail("GaGNet: 1..8 dilation rates");
    if (g.fft_num < 2 || (g.fft_num & 1)) return fail("GaGNet: fft_num must be even");
    if (g.acti_type < 0 || g.acti_type > 2) return fail("GaGNet: a activation function must be assigned! (GaGNet.py:171-172)");
    memset(&c, 0, sizeof(c));
    c.k1_t = g.k1_t; c.k1_f = g.k1_f; c.k2_t = g.k2_t; c.k2_f = g.k2_f; c.c = g.c; c.M = g.cin; c.embed_dim = 64;
    c.kd1 = g.kd1; c.cd1 = g.cd1; c.d_feat = g.d_feat; c.p = g.p; c.q = g.q; c.is_causal = g.is_causal; c.is_u2 = g.is_u2;
    c.intra_connect = g.intra_connect; c.norm_type = g.norm_type; c.n_freq = g.fft_num / 2 + 1;
    if (c.c < 1 || c.c > 128) return fail("c must be in 1..128");
    if (c.kd1 < 1 || c.kd1 > kMaxTaps) return fail("kd1 must be in 1..16");
    if (c.k1_t < 1 || c.k1_t > 2 || c.k2_t < 1 || c.k2_t > 2) return fail("temporal kernel sizes above 2 are not supported");
    if (c.k1_t * c.k1_f > kMaxTaps || c.k2_t * c.k2_f > kMaxTaps || c.k1_f < 1 || c.k2_f < 1) return fail("kernel too large");
    if (c.cd1 < 1 || c.cd1 > 128) return fail("cd1 must be in 1..128");
    if (c.p < 1 || c.q < 1 || c.q > 16) return fail("p >= 1 and 1 <= q <= 16 required");
    if (c.norm_type != 0 && c.norm_type != 1) return fail("norm_type must be 'IN' or 'BN'");
    Builder bd{m};
    m->Fchain.clear();
    m->Fchain.push_back(c.n_freq);
    {
        int F = conv_out_f(c.n_freq, 5);
        m->Fchain.push_back(F);
        for (int i = 0; i < 4; ++i) { F = F > 0 ? conv_out_f(F, c.k1_f) : -1; m->Fchain.push_back(F); }
        if (F < 1) return fail("fft_num too small for the five stride-2 encoder stages");
        m->Fb = F;
    }
    if (c.d_feat != 64 * m->Fb) return fail("d_feat must equal 64 * bottleneck_F (the reference fails at run time otherwise)");
    if (c.d_feat / 64 > 8) return fail("GaGNet: d_feat above 512 is not supported");
    if (c.is_u2) {
        m->en_mod.push_back(bd.module("en.meta_unet_list.0", 2 * c.M, 2, 5, 4, false));
        for (int i = 1; i < 4; ++i)
            m->en_mod.push_back(bd.module("en.meta_unet_list." + std::to_string(i), c.c, c.k1_t, c.k1_f, 4 - i, false));
        m->en_mod[0].in_conv.perm_ri = true;
        m->en_mod[0].in_conv.M = c.M;
        m->en_last = bd.gated("en.last_conv", c.c, 64, c.k1_t, c.k1_f, false, true);
        for (auto& U : m->en_mod) { for (auto& L : U.enco) L.zone = 2; for (auto& L : U.deco) L.zone = 2; }
    } else {
        // UNet_Encoder of GaGNet.py:368-413: every layer carries its norm (EaBNet's variant drops two of them)
        m->en_plain.push_back(bd.gated("en.unet_list.0", 2 * c.M, c.c, 2, 5, false, true));
        m->en_plain[0].perm_ri = true;
        m->en_plain[0].M = c.M;
        for (int i = 1; i < 4; ++i)
            m->en_plain.push_back(bd.gated("en.unet_list." + std::to_string(i), c.c, c.c, c.k1_t, c.k1_f, false, true));
        m->en_plain.push_back(bd.gated("en.unet_list.4", c.c, 64, c.k1_t, c.k1_f, false, true));
    }
    const int Fq = c.n_freq, ci = 2 * Fq + c.d_feat;
    auto tcm_groups = [&](const std::string& pfx, std::vector<TcmLayer>& dst) -> int {
        for (int gi = 0; gi < c.p; ++gi)
            for (int i = 0; i < g.n_dilas; ++i) {
                TcmLayer t;
                t.single = true;
                t.perm = false;
                t.dilation = g.dilas[i];
                if (t.dilation < 1) return fail("GaGNet: dilation rates must be positive");
                const std::string p = pfx + "." + std::to_string(gi) + ".tcns." + std::to_string(i);
                t.w_in = bd.add(p + ".in_conv.weight", {c.cd1, c.d_feat, 1}, EAB_P_CONV_W, c.d_feat);
                t.na_left.C = t.na_out.C = c.cd1;
                t.na_left.alpha = bd.add(p + ".d_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
                bd.norm(p + ".d_conv.1", c.cd1, t.na_left);
                t.w_left = bd.add(p + ".d_conv.3.weight", {c.cd1, c.cd1, c.kd1}, EAB_P_CONV_W, c.cd1 * c.kd1);
                t.na_out.alpha = bd.add(p + ".out_conv.0.weight", {c.cd1}, EAB_P_PRELU, c.cd1);
                bd.norm(p + ".out_conv.1", c.cd1, t.na_out);
                t.w_out = bd.add(p + ".out_conv.2.weight", {c.d_feat, c.cd1, 1}, EAB_P_CONV_W, c.cd1);
                const int span = (c.kd1 - 1) * t.dilation;
                if (!c.is_causal && (span & 1)) return fail("non-causal TCM needs an even (kd1-1)*dilation");
                const int pad_left = c.is_causal ? span : span / 2;
                for (int k = 0; k < c.kd1; ++k) t.dt[k] = pad_left - k * t.dilation;
                dst.push_back(t);
            }
        return 0;
    };
    auto in_convs = [&](const std::string& p, GagIn& in) {
        in.w_main = bd.add(p + ".in_conv_main.weight", {c.d_feat, ci, 1}, EAB_P_CONV_W, ci);
        in.b_main = bd.add(p + ".in_conv_main.bias", {c.d_feat}, EAB_P_CONV_B, ci);
        in.w_gate = bd.add(p + ".in_conv_gate.0.weight", {c.d_feat, ci, 1}, EAB_P_CONV_W, ci);
        in.b_gate = bd.add(p + ".in_conv_gate.0.bias", {c.d_feat}, EAB_P_CONV_B, ci);
    };
    auto lin = [&](const std::string& p, GagLin& l) {
        l.w = bd.add(p + ".weight", {Fq, c.d_feat, 1}, EAB_P_CONV_W, c.d_feat);
        l.b = bd.add(p + ".bias", {Fq}, EAB_P_CONV_B, c.d_feat);
    };
    m->gags.resize(c.q);
    for (int i = 0; i < c.q; ++i) {
        GagModule& G = m->gags[i];
        std::string p = "gags." + std::to_string(i) + ".glance_block";
        in_convs(p, G.in_g);
        EAB_TRY(tcm_groups(p + ".tcn_g", G.tcn_g));
        lin(p + ".linear_g.0", G.lin_g);
        p = "gags." + std::to_string(i) + ".gaze_block";
        in_convs(p, G.in_z);
        if (g.is_squeezed) {
            EAB_TRY(tcm_groups(p + ".tcm_ri", G.tcm_r));
        } else {
            EAB_TRY(tcm_groups(p + ".tcm_r", G.tcm_r));
            EAB_TRY(tcm_groups(p + ".tcm_i", G.tcm_i));
        }
        lin(p + ".linear_r", G.lin_r);
        lin(p + ".linear_i", G.lin_i);
    }
    return 0;
}


}  // namespace detail
}  // namespace eab
