// Forward orchestration: layer descriptors -> kernel launches (offline, planning and streaming passes share one walk).
#include "model_internal.h"

namespace eab {
namespace detail {

// ================================================================================================ run
struct Ctx {
    eab_model* m;
    bool dry;                 // size planning only: no launches, no dereference
    char* base;
    size_t stats_off = 0, stats_cap = 0;      // [0, stats_cap): zeroed once per forward
    size_t act_off = 0;
    int B, T;
    cudaStream_t st;

    // streaming (eab_stream_step): T == 1, every activation is a persistent ring of `last_RT` frames in the caller's
    // state blob (same allocation order every step => same addresses), nothing is reused, tensor-core kernels are off
    bool streaming = false;
    const int* step = nullptr;
    const int* start = nullptr;   // [streams] first absolute frame of each stream (eab_stream_reset_one)
    int gag_in_RT = 1; long long gag_in_slot = 0;      // streaming GaGNet: ring geometry of the `inpt` frame source
    int next_RT = 0;          // ring size of the next allocation (0 = the default of 2: current + previous frame)
    int last_RT = 0;          // ring size of the last allocation (0 offline)
    bool tensor_ok() const { return m->opt_umma && !streaming; }
    bool stream_umma() const { return m->opt_umma && streaming && m->opt_stream_umma; }      // conv_umma with ring addressing
    std::vector<TcmStreamDesc>* tcm_desc = nullptr;     // planning pass of eab_stream_reset: receives the descriptors
    std::vector<std::pair<size_t, size_t>>* per_stream = nullptr;      // planning pass: carried per-stream state to zero on a restart
    const TcmStreamDesc* tcm_desc_dev = nullptr;        // step: the table inside the state blob

    size_t act_peak = 0;
    float* alloc_act(size_t floats) {
        last_RT = streaming ? (next_RT ? next_RT : 2) : 0;
        next_RT = 0;
        if (streaming) floats *= last_RT;
        const size_t bytes = (floats * sizeof(float) + 255) / 256 * 256;
        float* p = reinterpret_cast<float*>(base + act_off);
        act_off += bytes;
        if (act_off > act_peak) act_peak = act_off;
        return p;
    }
    // scoped reuse: everything allocated after mark() is dead at release() (single stream => later kernels that
    // overwrite the region are ordered after the kernels that read it)
    size_t mark() const { return act_off; }
    void release(size_t m) { if (!streaming) act_off = m; }
    double* alloc_stats(int C) {
        const size_t bytes = ((size_t)B * C * 2 * sizeof(double) + 255) / 256 * 256;
        double* p = reinterpret_cast<double*>(base + stats_off);
        stats_off += bytes;
        return p;
    }
    const float* W(size_t off) const { return m->blob + off; }
};

// the Xform a consumer uses for the output of a conv followed by NormAct (2-D convention: norm -> PReLU)
Xform xf_after(Ctx& cx, const NormAct& na, double* stats, int count, int prelu_pos) {
    Xform x = xform_identity();
    if (na.has_norm) {
        if (cx.m->cfg.norm_type == 0) {
            x.affine = 1; x.stats = stats; x.inv_count = 1.f / (float)count;
            if (cx.m->opt_norm_log && !cx.dry && stats) cx.m->norm_log.push_back({na.gamma, stats, na.C, count, cx.B});
        }
        else x.affine = 2;
        x.scale = cx.W(na.off_scale);
        x.shift = cx.W(na.off_shift);
    }
    x.alpha = cx.W(na.off_alpha);
    x.alpha01 = na.alpha01 ? 1 : 0;
    x.prelu = prelu_pos;
    return x;
}

// Re-express a per-tap gather launch in the padded-pitch row space of conv_raw / the staged pair; false if the shape does not qualify.
bool to_plane_args(const UmmaConvArgs& u, PlaneConvArgs* p, int force_P = 0) {
    if (u.wide) return false;
    memset(p, 0, sizeof(*p));
    p->nsrc = u.nsrc;
    for (int i = 0; i < u.nsrc; ++i) p->src[i] = u.src[i];
    p->B = u.B; p->T = u.T; p->Fin = u.Fin; p->E = u.E;
    int min_df = 0, max_df = 0;
    for (int i = 0; i < u.ntaps; ++i) { min_df = std::min(min_df, u.df[i]); max_df = std::max(max_df, u.df[i]); }
    if (u.in_stride == 2) {
        if (min_df < 0) return false;
        p->nplanes = 2;
        p->plane_cols[0] = (u.Fin + 1) / 2; p->plane_cols[1] = u.Fin / 2;
        p->col_stride = 2; p->col_off[0] = 0; p->col_off[1] = 1;
        p->P = std::max(std::max(u.E + max_df / 2, p->plane_cols[0]), force_P);
    } else if (u.in_stride == 1) {
        if (max_df > 0) return false;
        p->nplanes = 1;
        p->plane_cols[0] = u.Fin; p->plane_cols[1] = 0;
        p->col_stride = 1; p->col_off[0] = 0; p->col_off[1] = 0;
        p->P = std::max(std::max(u.E, u.Fin - min_df), force_P);      // the pad columns [Fin, P) absorb the negative column offsets
    } else {
        return false;
    }
    p->ntaps = u.ntaps;
    int back = 0, fwd = 0;
    for (int i = 0; i < u.ntaps; ++i) {
        p->tap_plane[i] = u.in_stride == 2 ? (u.df[i] & 1) : 0;
        p->tap_shift[i] = -u.dt[i] * p->P + (u.in_stride == 2 ? (u.df[i] >> 1) : u.df[i]);
        back = std::max(back, -p->tap_shift[i]);
        fwd = std::max(fwd, p->tap_shift[i]);
    }
    p->back = back; p->fwd = fwd;
    p->out_stride = u.out_stride; p->out_off = u.out_off; p->Fout = u.Fout;
    p->nslab = u.nslab; p->ncoef = u.ncoef; p->npass = u.npass;
    p->Whi = u.Whi; p->Wlo = u.Wlo; p->bias = u.bias;
    p->Cout = u.Cout; p->N = u.N; p->gate_off = u.gate_off; p->relu = u.relu; p->algo_frac = u.algo_frac;
    p->out = u.out; p->out_ld = u.out_ld; p->out_coff = u.out_coff; p->resid = u.resid;
    p->nstats = u.nstats;
    p->stats_ld = u.stats_ld; p->stats_coff = u.stats_coff;
    for (int i = 0; i < 2; ++i) { p->stats[i] = u.stats[i]; p->stat_alpha[i] = u.stat_alpha[i]; }
    p->tiles_per_b = (int)(((long long)u.T * p->P + 127) / 128);
    p->nbuf = 1;
    return plane_conv_supported(*p);
}

// streaming: the variants of one layer on the tcgen05 gather kernel - all streams share one row space (B = 1, "frames" = streams)
int run_umma_stream(Ctx& cx, UmmaConvArgs* us, int n, int out_RT, int resid_RT) {
    if (cx.dry) return 0;
    for (int i = 0; i < n; ++i) {
        UmmaConvArgs& u = us[i];
        for (int k = 0; k < u.nsrc; ++k)
            if (u.src[k].x2) return fail("internal: lazy residual sum in a streaming step");
        u.B = 1; u.T = cx.B;
        u.tiles_per_b = (int)(((long long)cx.B * u.E + 127) / 128);
        u.step = cx.step; u.start = cx.start; u.out_RT = out_RT; u.resid_RT = resid_RT;
        u.nstats = 0;
        if (cx.m->opt_dbg_launch <= -300 && cx.m->umma_launch_idx++ == -300 - cx.m->opt_dbg_launch && cx.m->dbg_buf) u.dbg = cx.m->dbg_buf;
        if (n == 2 && cx.m->opt_stream_pair && !u.wide && us[0].out_stride == 2 && us[1].out_stride == 2 && us[0].bias == us[1].bias &&
            us[0].out == us[1].out && us[0].N == us[1].N) {
            if (i == 1) EAB_TRY(launch_conv_umma_pair(us[0], us[1], cx.st));     // both output parities of a transposed conv in one grid
            continue;
        }
        EAB_TRY(launch_conv_umma(u, cx.st));
    }
    return 0;
}

int launch_tensor_conv(eab_model* m, const UmmaConvArgs& u, cudaStream_t st) {
    (void)m;
    return launch_conv_umma(u, st);      // per-tap gather ring: what neither conv_raw nor the staged pair takes (K = 578, `staged` = 0)
}

// Launch the 1-4 tensor-core variants of one layer (output parities of a transposed conv, column splits of a wide
// 1x1) that read the same inputs.  Preferred path: stage the normalised fp16 planes ONCE (stage_kernel) and run the
// TMA-fed GEMM per variant; otherwise the fused-producer kernels.  Also runs in planning mode (allocations only).
// Can the 1-4 variants of a layer run as one stage launch + conv_tma launches?  Fills their plane arguments.
bool plan_planes(const UmmaConvArgs* us, int n, PlaneConvArgs* p, PlaneConvArgs* ps_out) {
    if (us[0].wide || n > 4) return false;
    bool ok = true;
    int P = 0;
    for (int i = 0; i < n; ++i) { ok = ok && to_plane_args(us[i], &p[i]); if (ok) P = std::max(P, p[i].P); }
    if (ok)
        for (int i = 0; i < n; ++i)
            if (p[i].P != P) ok = ok && to_plane_args(us[i], &p[i], P);
    if (ok)
        for (int i = 1; i < n; ++i)
            ok = ok && p[i].nplanes == p[0].nplanes && p[i].plane_cols[0] == p[0].plane_cols[0] &&
                 p[i].plane_cols[1] == p[0].plane_cols[1] && p[i].nslab == p[0].nslab && p[i].npass == p[0].npass &&
                 p[i].tiles_per_b == p[0].tiles_per_b;
    if (!ok) return false;
    PlaneConvArgs ps = p[0];
    for (int i = 1; i < n; ++i) { ps.back = std::max(ps.back, p[i].back); ps.fwd = std::max(ps.fwd, p[i].fwd); }
    *ps_out = ps;
    return true;
}

int run_tensor_convs(Ctx& cx, UmmaConvArgs* us, int n) {
    eab_model* m = cx.m;
    PlaneConvArgs p[4];
    PlaneConvArgs ps;
    const bool planes_ok = plan_planes(us, n, p, &ps);
    if (planes_ok && m->opt_raw && raw_conv_supported(p, n)) {
        {
            // one launch per layer: raw tiles -> norm + PReLU -> fp16 operand in shared memory -> GEMM (both parities)
            if (cx.dry) return 0;
            unsigned long long* dbg = nullptr;
            if (m->umma_launch_idx++ == m->opt_dbg_launch && m->dbg_buf) dbg = m->dbg_buf;
            return launch_conv_raw(p, n, cx.st, dbg, m->opt_raw_grid);
        }
    }
    bool staged_ok = planes_ok && m->opt_staged && staged_conv_supported(ps);
    for (int i = 0; staged_ok && i < n; ++i) staged_ok = staged_conv_supported(p[i]);
    if (staged_ok) {
        int front = 0;
        const int rows = staged_rows(ps, &front);
        const int nimg = ps.nplanes * ps.nslab * (ps.npass == 3 ? 2 : 1);
        ps.np_rows = rows; ps.np_front = front;
        const size_t scratch = cx.mark();           // the staged planes die with this layer
        for (int k = 0; k < nimg; ++k) ps.np[k] = cx.alloc_act((size_t)cx.B * rows * 32);     // 128 B per row
        cx.release(scratch);
        if (cx.dry) return 0;
        EAB_TRY(launch_stage(ps, cx.st));
        for (int i = 0; i < n; ++i) {
            if (m->umma_launch_idx++ == m->opt_dbg_launch && m->dbg_buf) p[i].dbg = m->dbg_buf;
            p[i].np_rows = rows; p[i].np_front = front;
            for (int k = 0; k < nimg; ++k) p[i].np[k] = ps.np[k];
            p[i].algo_in_share = 1.f / (float)n;
            EAB_TRY(launch_conv_staged(p[i], cx.st));
        }
        return 0;
    }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < us[i].nsrc; ++k)
            if (us[i].src[k].x2) return fail("internal: lazy residual sum reached a kernel that cannot read it");
    if (cx.dry) return 0;
    for (int i = 0; i < n; ++i) EAB_TRY(launch_tensor_conv(m, us[i], cx.st));
    return 0;
}

inline int zone_passes(const eab_model* m, int zone) {
    return zone == 3 ? m->opt_first_passes : zone == 0 ? m->opt_enc_passes : zone == 1 ? m->opt_dec_passes : m->opt_inner_passes;
}

// one 2-D layer: conv/deconv (+gate) -> raw output + statistics; returns the Act a consumer should read
int materialize(Ctx& cx, Act* a);
// fuse_res / fuse_out (streaming on the tensor cores only, see fuse_ok in run_module): the launch writes xform(own result) +
// xform(*fuse_res) straight into *fuse_out's buffer - the module's residual sum without a combine launch
int run_conv2d(Ctx& cx, const ConvLayer& L, const Act* srcs_in, int nsrc, Act* out, float* prealloc = nullptr,
               const Act* fuse_res = nullptr, const Act* fuse_out = nullptr) {
    Act srcs[2];
    for (int i = 0; i < nsrc; ++i) srcs[i] = srcs_in[i];
    if (!(cx.tensor_ok() && L.umma_ok && cx.m->opt_staged && !L.wide))
        for (int i = 0; i < nsrc; ++i) EAB_TRY(materialize(cx, &srcs[i]));       // only conv_raw / the staged pair read lazy sums
    const int Fin = srcs[0].F;
    int cin = 0;
    for (int i = 0; i < nsrc; ++i) {
        if (srcs[i].F != Fin) return fail("skip connection width mismatch (the reference's torch.cat would raise too)");
        cin += srcs[i].C;
    }
    if (cin != L.cin) return fail("internal: channel mismatch in conv layer");
    const int Fout = L.deconv ? deconv_out_f(Fin, L.kf) : conv_out_f(Fin, L.kf);
    if (Fout < 1) return fail("frequency axis too short for this layer");
    out->F = Fout;
    out->C = L.cout;
    out->data2 = nullptr;
    const size_t out_elems = (size_t)cx.B * cx.T * Fout * L.cout;
    const bool in_stats = L.na.has_norm && cx.m->cfg.norm_type == 0;
    double* stats = in_stats ? cx.alloc_stats(L.cout) : nullptr;
    out->xf = xf_after(cx, L.na, stats, cx.T * Fout, 2);
    auto allocate_out = [&]() {
        out->data = prealloc ? prealloc : cx.alloc_act(out_elems);
        out->RT = prealloc ? 0 : cx.last_RT;
    };
    if (cx.tensor_ok() && L.umma_ok && L.wide && cx.m->opt_staged && nsrc == 1 && !L.deconv &&
        srcs[0].xf.affine == 0 && srcs[0].xf.prelu == 0 && !srcs[0].data2 && (srcs[0].C * 2) % 2 == 0) {
        // first layer (2M input channels) on the staged path: a plane row is a frequency PAIR (or the whole kf x C tap window)
        allocate_out();
        PlaneConvArgs p;
        memset(&p, 0, sizeof(p));
        p.nsrc = 1;
        set_src(p.src[0], srcs[0]);
        const bool pair = L.pair_ok;
        p.B = cx.B; p.T = cx.T; p.Fin = Fin; p.E = Fout;
        p.nplanes = 1; p.col_stride = 2; p.col_off[0] = 0;
        int back = 0, fwd = 0;
        if (pair) {
            // rows = frequency PAIRS (2 x cin values each): 1.8x the input bytes per pass-plane instead of 3.6x for window rows
            p.P = (Fin + 1) / 2;
            p.plane_cols[0] = p.P;
            p.ntaps = L.p_ntaps;
            for (int i = 0; i < p.ntaps; ++i) {
                p.tap_plane[i] = 0; p.tap_shift[i] = -L.p_dt[i] * p.P + L.p_ds[i];
                back = std::max(back, -p.tap_shift[i]); fwd = std::max(fwd, p.tap_shift[i]);
            }
            p.nslab = 1;
            p.Whi = cx.W(L.off_phi); p.Wlo = cx.W(L.off_plo);
            p.algo_frac = (float)(L.p_stack * L.kf * cin) / (float)(((L.kf + 1) / 2) * 64);
            p.wide_k = 2 * cin;
            p.wide_kt = L.p_stack;
            if (L.p_stack * 2 * cin <= 48 && L.p_stack * 2 * cin > 32) p.ksteps = 3;      // 36 of 64 columns (M = 9): the fourth K step is padding
        } else {
            p.P = Fout;
            p.plane_cols[0] = Fout;
            p.ntaps = L.u_ntaps[0];
            for (int i = 0; i < p.ntaps; ++i) { p.tap_plane[i] = 0; p.tap_shift[i] = -L.u_dt[0][i] * p.P; back = std::max(back, -p.tap_shift[i]); }
            p.nslab = L.u_nslab;
            p.Whi = cx.W(L.off_whi[0]); p.Wlo = cx.W(L.off_wlo[0]);
            p.algo_frac = (float)L.u_kwidth / (float)(L.u_nslab * 64);
            p.wide_k = L.u_kwidth;
        }
        p.back = back; p.fwd = fwd;
        p.out_stride = 1; p.out_off = 0; p.Fout = Fout;
        p.ncoef = cin; p.npass = zone_passes(cx.m, L.zone);
        p.bias = cx.W(L.off_ub);
        p.Cout = L.cout; p.N = L.u_N; p.gate_off = L.u_gate_off;
        p.out = out->data; p.out_ld = L.cout; p.out_coff = 0;
        if (stats) { p.nstats = 1; p.stats[0] = stats; }
        p.tiles_per_b = (int)(((long long)cx.T * p.P + 127) / 128);
        const bool geom_ok = (srcs[0].C * p.col_stride) % 2 == 0 && (Fin * srcs[0].C) % 2 == 0 &&      // 8-byte aligned windows
                             (Fout - 1) * 2 + L.kf <= Fin && p.P >= 1 &&
                             ((long long)cx.T * p.P + p.back + 4 * 128 + 2ll * p.P) * p.P < (1ll << 31);
        if (geom_ok && staged_conv_fits(p)) {
            int front = 0;
            const int rows = staged_rows(p, &front);
            const int nimg = p.nslab * (p.npass == 3 ? 2 : 1);
            p.np_rows = rows; p.np_front = front;
            const size_t scratch = cx.mark();
            for (int k = 0; k < nimg; ++k) p.np[k] = cx.alloc_act((size_t)cx.B * rows * 32);
            cx.release(scratch);
            if (cx.dry) return 0;
            EAB_TRY(launch_stage(p, cx.st));
            if (cx.m->umma_launch_idx++ == cx.m->opt_dbg_launch && cx.m->dbg_buf) p.dbg = cx.m->dbg_buf;
            return launch_conv_staged(p, cx.st);
        }
    }
    if ((cx.tensor_ok() || cx.stream_umma()) && L.umma_ok) {
        UmmaConvArgs us[4];
        bool all_ok = true;
        int nus = 0;
        for (int v = 0; v < L.nvar; ++v) {
            const int npass = zone_passes(cx.m, L.zone);
            {
                UmmaConvArgs& u = us[nus++];
                memset(&u, 0, sizeof(u));
                u.nsrc = nsrc;
                for (int i = 0; i < nsrc; ++i) set_src(u.src[i], srcs[i]);
                u.B = cx.B; u.T = cx.T; u.Fin = Fin; u.Fout = Fout;
                if (L.deconv) { u.in_stride = 1; u.out_stride = 2; u.out_off = v; u.E = (Fout - v + 1) / 2; }
                else          { u.in_stride = 2; u.out_stride = 1; u.out_off = 0; u.E = Fout; }
                u.ntaps = L.u_ntaps[v];
                for (int i = 0; i < u.ntaps; ++i) { u.dt[i] = L.u_dt[v][i]; u.df[i] = L.u_df[v][i]; }
                u.wide = L.wide; u.kwidth = L.u_kwidth; u.nslab = L.u_nslab; u.ncoef = cin;
                u.npass = npass;
                u.algo_frac = 1.f;
                u.out = out->data; u.out_ld = L.cout;
                u.Whi = cx.W(L.off_whi[v]); u.Wlo = cx.W(L.off_wlo[v]); u.bias = cx.W(L.off_ub);
                u.Cout = L.cout; u.N = L.u_N; u.gate_off = L.u_gate_off; u.out_coff = 0;
                if (stats) { u.nstats = 1; u.stats[0] = stats; }
                u.tiles_per_b = (cx.T * u.E + 127) / 128;
                all_ok = all_ok && umma_conv_supported(u);
            }
        }
        if (all_ok && fuse_res) {
            if (!cx.streaming) return fail("internal: fused residual sum outside a streaming step");
            out->data = fuse_out->data; out->RT = fuse_out->RT;
            for (int i = 0; i < nus; ++i) {
                us[i].out = out->data;
                us[i].post = 1; us[i].post_xf = out->xf; us[i].resid = fuse_res->data; us[i].resid_xf = fuse_res->xf;
            }
            out->xf = xform_identity();
            return run_umma_stream(cx, us, nus, out->RT, fuse_res->RT);
        }
        if (fuse_res) return fail("internal: fused residual sum on a layer the tensor-core path does not take");
        if (all_ok) {
            allocate_out();
            for (int i = 0; i < nus; ++i) us[i].out = out->data;
            if (cx.streaming) return run_umma_stream(cx, us, nus, out->RT, 0);
            return run_tensor_convs(cx, us, nus);
        }
    }
    allocate_out();
    for (int i = 0; i < nsrc; ++i) EAB_TRY(materialize(cx, &srcs[i]));
    if (cx.dry) return 0;
    for (int v = 0; v < L.nvar; ++v) {
        ConvArgs a;
        memset(&a, 0, sizeof(a));
        a.nsrc = nsrc;
        for (int i = 0; i < nsrc; ++i) { a.src[i].x = srcs[i].data; a.src[i].C = srcs[i].C; a.src[i].xf = srcs[i].xf; a.src[i].RT = srcs[i].RT; }
        a.step = cx.step; a.start = cx.start; a.out_RT = out->RT;
        a.B = cx.B; a.T = cx.T; a.Fin = Fin; a.Fout = Fout;
        if (L.deconv) { a.in_stride = 1; a.out_stride = 2; a.out_off = v; a.E = (Fout - v + 1) / 2; }
        else          { a.in_stride = 2; a.out_stride = 1; a.out_off = 0; a.E = Fout; }
        a.ntaps = L.ntaps[v];
        for (int i = 0; i < a.ntaps; ++i) { a.dt[i] = L.dt[v][i]; a.df[i] = L.df[v][i]; }
        if (a.ntaps == 0) return fail("transposed conv with kf == 1 is not supported");
        a.W = cx.W(L.off_w[v]);
        a.bias = cx.W(L.off_b);
        a.Cout = L.cout; a.N = L.N; a.gate_off = L.gate_off;
        a.algo_frac = 1.f;
        a.out = out->data;
        if (stats) { a.nstats = 1; a.stats[0] = stats; }
        EAB_TRY(launch_conv(a, cx.st));
    }
    return 0;
}

int run_combine_into(Ctx& cx, const Act* srcs, int nsrc, Act* out);
int run_combine(Ctx& cx, const Act* srcs, int nsrc, Act* out) {
    out->F = srcs[0].F;
    out->C = srcs[0].C;
    out->xf = xform_identity();
    out->data = cx.alloc_act((size_t)cx.B * cx.T * out->F * out->C);
    out->RT = cx.last_RT;
    return run_combine_into(cx, srcs, nsrc, out);
}

// same, into an Act whose buffer (F, C, data) the caller has already allocated
int run_combine_into(Ctx& cx, const Act* srcs, int nsrc, Act* out) {
    if (cx.dry) return 0;
    CombineArgs a;
    memset(&a, 0, sizeof(a));
    int n = 0;
    for (int i = 0; i < nsrc; ++i) {
        if (srcs[i].F != out->F || srcs[i].C != out->C) return fail("internal: combine shape mismatch");
        if (n + (srcs[i].data2 ? 2 : 1) > 3) return fail("internal: too many addends in combine");
        a.src[n].x = srcs[i].data; a.src[n].C = srcs[i].C; a.src[n].xf = srcs[i].xf; a.src[n].RT = srcs[i].RT; ++n;
        if (srcs[i].data2) { a.src[n].x = srcs[i].data2; a.src[n].C = srcs[i].C; a.src[n].xf = srcs[i].xf2; a.src[n].RT = srcs[i].RT; ++n; }
    }
    a.nsrc = n;
    a.B = cx.B; a.P = cx.T * out->F; a.C = out->C; a.out = out->data;
    a.step = cx.step; a.out_RT = out->RT;
    return launch_combine(a, cx.st);
}

// materialise a lazy residual sum (only needed in front of kernels that cannot read one)
int materialize(Ctx& cx, Act* a) {
    if (!a->data2) return 0;
    Act src = *a;
    Act dst;
    EAB_TRY(run_combine(cx, &src, 1, &dst));
    *a = dst;
    return 0;
}

// En_unet_module.forward (EaBNet.py:372-388)
int run_combine_into(Ctx& cx, const Act* srcs, int nsrc, Act* out);
int run_module(Ctx& cx, const UnetModule& U, const Act* srcs, int nsrc, Act* out) {
    // the module output is allocated first; everything else (in_conv output, inner U-Net maps) is scoped scratch
    const int Fin = srcs[0].F;
    const int Fw = U.in_conv.deconv ? deconv_out_f(Fin, U.in_conv.kf) : conv_out_f(Fin, U.in_conv.kf);
    if (Fw < 1) return fail("frequency axis too short for this layer");
    out->F = Fw;
    out->C = U.in_conv.cout;
    out->xf = xform_identity();
    // lazy mode: the module result x0 + y is never written; its two addends (in_conv output, last inner deconv output)
    // outlive the module instead and every consumer's stage kernel sums them while staging
    const bool lazy = cx.m->opt_lazy && cx.tensor_ok() && cx.m->opt_staged && U.in_conv.umma_ok && !U.deco.empty() &&
                      U.deco.back().umma_ok;
    const size_t nel = (size_t)cx.B * cx.T * Fw * out->C;
    float* buf_x0 = nullptr;
    float* buf_y = nullptr;
    if (lazy) { buf_x0 = cx.alloc_act(nel); buf_y = cx.alloc_act(nel); out->data = nullptr; }
    else { out->data = cx.alloc_act(nel); out->RT = cx.last_RT; }
    const size_t scope = cx.mark();
    Act x0;
    EAB_TRY(run_conv2d(cx, U.in_conv, srcs, nsrc, &x0, buf_x0));
    Act y = x0;
    std::vector<Act> keep;
    for (size_t i = 0; i < U.enco.size(); ++i) {
        Act z;
        EAB_TRY(run_conv2d(cx, U.enco[i], &y, 1, &z, nullptr));
        keep.push_back(z);
        y = z;
    }
    // streaming on the tensor cores (static normalisation): the last inner deconv's epilogue adds the transformed x0 and writes the
    // module output itself - no combine launch (9 of a step's launches)
    const bool fuse_ok = !lazy && cx.streaming && cx.stream_umma() && cx.m->opt_stream_fuse && !U.deco.empty() && U.deco.back().umma_ok &&
                         !U.deco.back().wide && (U.deco.size() == 1 || cx.m->cfg.intra_connect == 0) && cx.m->cfg.norm_type != 0 &&
                         U.deco.back().cout == out->C;
    for (size_t i = 0; i < U.deco.size(); ++i) {
        Act z;
        float* pre = (lazy && i + 1 == U.deco.size()) ? buf_y : nullptr;
        const bool fuse = fuse_ok && i + 1 == U.deco.size();
        if (i == 0) {
            EAB_TRY(run_conv2d(cx, U.deco[i], &y, 1, &z, pre, fuse ? &x0 : nullptr, fuse ? out : nullptr));
        } else {
            Act pair[2] = {y, keep[keep.size() - 1 - i]};
            if (cx.m->cfg.intra_connect == 0) {
                EAB_TRY(run_conv2d(cx, U.deco[i], pair, 2, &z, pre, fuse ? &x0 : nullptr, fuse ? out : nullptr));
            } else {
                Act sum;
                EAB_TRY(run_combine(cx, pair, 2, &sum));
                EAB_TRY(run_conv2d(cx, U.deco[i], &sum, 1, &z, pre));
            }
        }
        y = z;
    }
    if (lazy) {
        out->data = x0.data; out->xf = x0.xf;
        out->data2 = y.data; out->xf2 = y.xf;
    } else if (!fuse_ok) {
        Act pair[2] = {x0, y};
        EAB_TRY(run_combine_into(cx, pair, 2, out));
    }
    cx.release(scope);
    return 0;
}

// 1x1 "conv" over positions with optional bias / relu / residual / statistics (used by TCMs and the head)
int run_pointwise(Ctx& cx, const Act* srcs, int nsrc, const float* W, const float* bias, int Cout, int N, int gate_off,
                  int ntaps, const int* dt, int relu, const float* resid, int nstats, double** stats,
                  const float** stat_alpha, Act* out, const UmmaW* uw = nullptr, bool preallocated = false, int resid_RT = 0) {
    const bool use_umma = uw && uw->ok && (cx.tensor_ok() || cx.stream_umma()) && (resid == nullptr || uw->ld == Cout);
    if (!preallocated) {
        out->F = srcs[0].F;
        out->C = use_umma ? uw->ld : Cout;      // the tcgen05 path may pad the channel count (e.g. 18 -> 32, zeros)
        out->xf = xform_identity();
        out->data = cx.alloc_act((size_t)cx.B * cx.T * out->F * out->C);
        out->RT = cx.last_RT;
    }
    if (use_umma) {
        UmmaConvArgs us[4];
        for (int sp = 0; sp < uw->nsplit; ++sp) {
            UmmaConvArgs& u = us[sp];
            memset(&u, 0, sizeof(u));
            u.nsrc = nsrc;
            int cin = 0;
            for (int i = 0; i < nsrc; ++i) { set_src(u.src[i], srcs[i]); cin += srcs[i].C; }
            u.B = cx.B; u.T = cx.T; u.Fin = srcs[0].F; u.Fout = srcs[0].F; u.E = srcs[0].F;
            u.in_stride = 1; u.out_stride = 1; u.out_off = 0;
            u.ntaps = uw->ntaps;
            for (int i = 0; i < u.ntaps; ++i) { u.dt[i] = dt ? dt[i] : 0; u.df[i] = 0; }
            u.wide = 0; u.kwidth = 0; u.nslab = uw->nslab; u.ncoef = cin;
            u.npass = 3;                        // these layers are < 5 % of the FLOPs: keep them fp32-grade
            u.Whi = cx.W(uw->off_hi[sp]); u.Wlo = cx.W(uw->off_lo[sp]);
            u.bias = uw->has_bias ? cx.W(uw->off_bias[sp]) : nullptr;
            u.Cout = uw->cout; u.N = uw->ncol; u.gate_off = uw->gate_off; u.relu = relu;
            u.algo_frac = (nsrc == 2 && gate_off > 0) ? 0.5f : 1.f;
            u.out = out->data; u.out_ld = uw->ld; u.out_coff = sp * uw->cout;
            u.resid = resid;
            u.nstats = nstats;
            for (int i = 0; i < nstats; ++i) { u.stats[i] = stats[i]; u.stat_alpha[i] = stat_alpha[i]; }
            u.tiles_per_b = (cx.T * u.E + 127) / 128;
            if (!umma_conv_supported(u)) return fail("internal: pointwise layer rejected by the tcgen05 path");
        }
        if (cx.streaming) return run_umma_stream(cx, us, uw->nsplit, out->RT, resid_RT);
        return run_tensor_convs(cx, us, uw->nsplit);
    }
    if (cx.dry) return 0;
    for (int i = 0; i < nsrc; ++i)
        if (srcs[i].data2) return fail("internal: lazy activation reached the CUDA-core pointwise kernel");
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.nsrc = nsrc;
    for (int i = 0; i < nsrc; ++i) { a.src[i].x = srcs[i].data; a.src[i].C = srcs[i].C; a.src[i].xf = srcs[i].xf; a.src[i].RT = srcs[i].RT; }
    a.step = cx.step; a.start = cx.start; a.out_RT = out->RT; a.resid_RT = resid_RT;
    a.B = cx.B; a.T = cx.T; a.Fin = srcs[0].F; a.Fout = srcs[0].F; a.E = srcs[0].F;
    a.in_stride = 1; a.out_stride = 1; a.out_off = 0;
    a.ntaps = ntaps;
    for (int i = 0; i < ntaps; ++i) { a.dt[i] = dt ? dt[i] : 0; a.df[i] = 0; }
    a.W = W; a.bias = bias; a.Cout = Cout; a.N = N; a.gate_off = gate_off; a.relu = relu;
    a.algo_frac = (nsrc == 2 && gate_off > 0) ? 0.5f : 1.f;      // merged TCM branches: block-diagonal weights
    a.out = out->data; a.resid = resid;
    a.nstats = nstats;
    for (int i = 0; i < nstats; ++i) { a.stats[i] = stats[i]; a.stat_alpha[i] = stat_alpha[i]; }
    return launch_conv(a, cx.st);
}

// SqueezedTCM.forward (EaBNet.py:572-578) on the channels-last residual stream x [B,T,1,d_feat]
int run_tcm(Ctx& cx, const TcmLayer& t, const Act& x, Act* out) {
    const eab_config& c = cx.m->cfg;
    const bool in_stats = c.norm_type == 0;
    // the residual-stream output first, the squeezed intermediates are scoped scratch
    out->F = x.F; out->C = c.d_feat; out->xf = xform_identity();
    out->data = cx.alloc_act((size_t)cx.B * cx.T * x.F * c.d_feat);
    out->RT = cx.last_RT;
    const size_t scope = cx.mark();
    if (cx.streaming) {                        // the dilated taps reach (kd1-1)*d frames back into the squeezed tensor
        int back = 0;
        for (int i = 0; i < c.kd1; ++i) back = std::max(back, t.dt[i]);
        cx.next_RT = back + 1;
    }
    Act z;
    double* st_o[1] = {nullptr};
    if (t.single) {
        // GaGNet's SqueezedTCM: squeeze 1x1 (statistics of PReLU(y)), one plain dilated conv
        double* st_d[1] = {in_stats ? cx.alloc_stats(c.cd1) : nullptr};
        const float* al_d[1] = {cx.W(t.na_left.off_alpha)};
        Act y;
        EAB_TRY(run_pointwise(cx, &x, 1, cx.W(t.off_in), nullptr, c.cd1, pad_n(c.cd1), 0, 1, nullptr, 0, nullptr,
                              in_stats ? 1 : 0, st_d, al_d, &y, &t.u_in));
        y.xf = xf_after(cx, t.na_left, st_d[0], cx.T, 1);
        st_o[0] = in_stats ? cx.alloc_stats(c.cd1) : nullptr;
        const float* al_o[1] = {cx.W(t.na_out.off_alpha)};
        EAB_TRY(run_pointwise(cx, &y, 1, cx.W(t.off_dil), nullptr, c.cd1, pad_n(c.cd1), 0, c.kd1, t.dt, 0, nullptr,
                              in_stats ? 1 : 0, st_o, al_o, &z, &t.u_dil));
    } else {
    // squeeze 1x1; statistics of PReLU_left(y) and PReLU_right(y) for the two branch norms
    double* st_lr[2] = {in_stats ? cx.alloc_stats(c.cd1) : nullptr, in_stats ? cx.alloc_stats(c.cd1) : nullptr};
    const float* al_lr[2] = {cx.W(t.na_left.off_alpha), cx.W(t.na_right.off_alpha)};
    Act y;
    EAB_TRY(run_pointwise(cx, &x, 1, cx.W(t.off_in), nullptr, c.cd1, pad_n(c.cd1), 0, 1, nullptr, 0, nullptr,
                          in_stats ? 2 : 0, st_lr, al_lr, &y, &t.u_in));
    // both dilated branches as one gated conv: value = left branch, gate = right branch (sigmoid)
    Act br[2] = {y, y};
    br[0].xf = xf_after(cx, t.na_left, st_lr[0], cx.T, 1);
    br[1].xf = xf_after(cx, t.na_right, st_lr[1], cx.T, 1);
    st_o[0] = in_stats ? cx.alloc_stats(c.cd1) : nullptr;
    const float* al_o[1] = {cx.W(t.na_out.off_alpha)};
    EAB_TRY(run_pointwise(cx, br, 2, cx.W(t.off_dil), nullptr, c.cd1, 2 * ceil64(c.cd1), ceil64(c.cd1), c.kd1, t.dt, 0,
                          nullptr, in_stats ? 1 : 0, st_o, al_o, &z, &t.u_dil));
    }
    // expand 1x1 + residual
    z.xf = xf_after(cx, t.na_out, st_o[0], cx.T, 1);
    EAB_TRY(run_pointwise(cx, &z, 1, cx.W(t.off_out), nullptr, c.d_feat, c.d_feat, 0, 1, nullptr, 0, x.data, 0, nullptr,
                          nullptr, out, &t.u_out, /*preallocated=*/true, x.RT));
    cx.release(scope);
    return 0;
}

void tap(Ctx& cx, const char* name, const Act& a) {
    if (cx.dry || cx.streaming) return;
    Tap t;
    t.act = a; t.B = cx.B; t.T = cx.T;
    cx.m->taps[name] = t;
}

struct ChainRef { const TcmLayer* l; int n; };
bool tcm_chain_ok(Ctx& cx, const ChainRef* chains, int nch);
int run_tcm_chains(Ctx& cx, const ChainRef* chains, int nch, const Act* ins, Act* outs);

int run_forward(Ctx& cx, const float* inpt, float* out_dev) {
    eab_model* m = cx.m;
    const eab_config& c = m->cfg;
    Act x;
    x.data = const_cast<float*>(inpt);
    x.F = c.n_freq;
    x.C = 2 * c.M;
    x.RT = cx.streaming ? 2 : 0;          // streaming: inpt is the [S][2][F][M][2] spectrum ring of the stream state
    const int inpt_RT = x.RT;
    std::vector<Act> skips;
    // ---------------- encoder (EaBNet.py:190-197 / :234-239)
    if (c.is_u2) {
        for (size_t i = 0; i < m->en_mod.size(); ++i) {
            Act y;
            EAB_TRY(run_module(cx, m->en_mod[i], &x, 1, &y));
            skips.push_back(y);
            tap(cx, ("en." + std::to_string(i)).c_str(), y);
            x = y;
        }
        Act y;
        EAB_TRY(run_conv2d(cx, m->en_last, &x, 1, &y));
        skips.push_back(y);
        tap(cx, "en.4", y);
        x = y;
    } else {
        for (size_t i = 0; i < m->en_plain.size(); ++i) {
            Act y;
            EAB_TRY(run_conv2d(cx, m->en_plain[i], &x, 1, &y));
            skips.push_back(y);
            tap(cx, ("en." + std::to_string(i)).c_str(), y);
            x = y;
        }
    }
    if (x.F != m->Fb || x.C != 64) return fail("internal: bottleneck shape");
    // ---------------- squeezed TCM stack (EaBNet.py:99-106); channel order f*64+c is kept both ways
    Act r;                                            // residual stream [B,T,1,d_feat], finalised
    EAB_TRY(run_combine(cx, &x, 1, &r));
    r.F = 1;
    r.C = c.d_feat;
    r.xf = xform_identity();
    Act acc;
    if (cx.streaming && tcm_stream_supported(c.cd1, c.d_feat, c.kd1) && c.norm_type == 1 && m->opt_stream_tcm) {
        // one launch for the whole stack (tcm_stream.cu): per-TCM history rings of the squeezed tensor + the group sum
        std::vector<TcmStreamDesc> descs;
        for (const TcmLayer& t : m->tcms) {
            TcmStreamDesc d;
            memset(&d, 0, sizeof(d));
            int back = 0;
            for (int i = 0; i < c.kd1; ++i) { d.dt[i] = t.dt[i]; back = std::max(back, t.dt[i]); }
            cx.next_RT = back + 1;
            float* ring = cx.alloc_act((size_t)cx.B * c.cd1);
            d.RT = cx.last_RT;
            d.ring_off = (long long)((reinterpret_cast<char*>(ring) - cx.base) / (ptrdiff_t)sizeof(float));
            d.W_in = (long long)t.off_in; d.W_dil = (long long)t.off_dil; d.W_out = (long long)t.off_out;
            d.sL = (long long)t.na_left.off_scale; d.hL = (long long)t.na_left.off_shift; d.aL = (long long)t.na_left.off_alpha;
            d.sR = (long long)t.na_right.off_scale; d.hR = (long long)t.na_right.off_shift; d.aR = (long long)t.na_right.off_alpha;
            d.sO = (long long)t.na_out.off_scale; d.hO = (long long)t.na_out.off_shift; d.aO = (long long)t.na_out.off_alpha;
            descs.push_back(d);
        }
        acc.F = 1; acc.C = c.d_feat; acc.xf = xform_identity();
        acc.data = cx.alloc_act((size_t)cx.B * c.d_feat);
        acc.RT = cx.last_RT;
        if (cx.tcm_desc) *cx.tcm_desc = descs;
        if (!cx.dry) {
            TcmStreamArgs a;
            memset(&a, 0, sizeof(a));
            a.desc = cx.tcm_desc_dev; a.blob = m->blob;
            a.ntcm = (int)m->tcms.size(); a.p = c.p; a.kd = c.kd1; a.S = cx.B;
            a.step = cx.step; a.start = cx.start; a.act_base = reinterpret_cast<float*>(cx.base);
            a.x = r.data; a.x_RT = r.RT; a.out = acc.data; a.out_RT = acc.RT;
            EAB_TRY(launch_tcm_stream(a, cx.st));
        }
    } else {
    std::vector<Act> group_out;
    size_t ti = 0;
    for (int g = 0; g < c.q; ++g) {
        const ChainRef chain = {m->tcms.data() + (size_t)g * c.p, c.p};
        if (!cx.streaming && tcm_chain_ok(cx, &chain, 1)) {
            // a whole group (p gated TCMs) as one persistent cooperative launch (tcm_chain.cu)
            Act nx;
            EAB_TRY(run_tcm_chains(cx, &chain, 1, &r, &nx));
            r = nx;
            ti += c.p;
        } else {
            for (int i = 0; i < c.p; ++i) {
                Act nx;
                EAB_TRY(run_tcm(cx, m->tcms[ti++], r, &nx));
                r = nx;
            }
        }
        group_out.push_back(r);
    }
    if (c.q == 1) acc = group_out[0];
    else EAB_TRY(run_combine(cx, group_out.data(), c.q, &acc));
    }
    acc.F = m->Fb;
    acc.C = 64;
    tap(cx, "tcm", acc);
    // ---------------- decoder (EaBNet.py:273-279 / :324-328)
    x = acc;
    Act emb;
    if (c.is_u2) {
        for (size_t i = 0; i < m->de_mod.size(); ++i) {
            Act pair[2] = {x, skips[skips.size() - 1 - i]};
            Act y;
            EAB_TRY(run_module(cx, m->de_mod[i], pair, 2, &y));
            tap(cx, ("de." + std::to_string(i)).c_str(), y);
            x = y;
        }
        Act pair[2] = {x, skips[0]};
        EAB_TRY(run_conv2d(cx, m->de_last, pair, 2, &emb));
    } else {
        for (size_t i = 0; i < m->de_plain.size(); ++i) {
            Act pair[2] = {x, skips[skips.size() - 1 - i]};
            Act y;
            EAB_TRY(run_conv2d(cx, m->de_plain[i], pair, 2, &y));
            if (i + 1 < m->de_plain.size()) tap(cx, ("de." + std::to_string(i)).c_str(), y);
            x = y;
        }
        emb = x;
    }
    if (emb.F != c.n_freq) return fail("decoder output width differs from the input width (the reference would fail in filter-and-sum)");
    tap(cx, "embed", emb);
    // ---------------- beam-weight head + filter-and-sum (EaBNet.py:108-125, 600-614)
    Act w;
    if (c.topo_type == 0 && c.bf_type == 0) {
        Act h[2];
        for (int l = 0; l < 2; ++l) {
            h[l].F = c.n_freq; h[l].C = 64; h[l].xf = xform_identity();
            h[l].data = cx.alloc_act((size_t)cx.B * cx.T * c.n_freq * 64);
            h[l].RT = cx.last_RT;
            float* hc_state[2] = {nullptr, nullptr};
            if (cx.streaming)
                for (int k = 0; k < 2; ++k) {
                    cx.next_RT = 1; hc_state[k] = cx.alloc_act((size_t)cx.B * c.n_freq * 64);
                    if (cx.per_stream) cx.per_stream->push_back({(size_t)(reinterpret_cast<char*>(hc_state[k]) - cx.base), (size_t)c.n_freq * 64 * sizeof(float)});
                }
            if (!cx.dry) {
                LstmArgs a;
                memset(&a, 0, sizeof(a));
                const Act& src = l ? h[0] : emb;
                a.src.x = src.data; a.src.C = src.C; a.src.xf = src.xf; a.src.RT = src.RT;
                a.step = cx.step; a.h_state = hc_state[0]; a.c_state = hc_state[1]; a.out_RT = h[l].RT;
                a.layer_norm = l == 0;
                a.ln_g = cx.W(m->off_ln_g); a.ln_b = cx.W(m->off_ln_b);
                a.Wx = cx.W(m->off_rnn[l][0]); a.Wh = cx.W(m->off_rnn[l][1]); a.bias = cx.W(m->off_rnn[l][2]);
                a.B = cx.B; a.T = cx.T; a.F = c.n_freq; a.E = src.C;
                a.out = h[l].data;
                if (cx.tensor_ok() && m->rnn_umma_ok) {
                    LstmArgs u = a;
                    if (m->opt_dbg_launch == -100 - l && m->dbg_buf) u.dbg = m->dbg_buf;
                    u.exp_flags = m->opt_lstm_exp;
                    u.Wimg = cx.W(m->off_rnn_img[l]);
                    u.bias = cx.W(m->off_rnn_ubias[l]);
                    if (m->opt_lstm_pp && lstm_pp_supported(u)) {
                        EAB_TRY(launch_lstm_pp(u, cx.st));
                        tap(cx, l ? "h2" : "h1", h[l]);
                        continue;
                    }
                    if (lstm_umma_supported(u)) {
                        EAB_TRY(launch_lstm_umma(u, cx.st));
                        tap(cx, l ? "h2" : "h1", h[l]);
                        continue;
                    }
                }
                if (!(cx.stream_umma() && m->rnn_umma_ok && m->u_rnn_step[l].ok && m->opt_stream_lstm)) EAB_TRY(launch_lstm(a, cx.st));
            }
            if (cx.stream_umma() && m->rnn_umma_ok && (cx.dry || m->u_rnn_step[l].ok) && m->opt_stream_lstm) {
                // streaming step on the tensor cores: [LayerNorm kernel ->] gate GEMM over all (stream, f) rows -> cell kernel
                const Act& src = l ? h[0] : emb;
                Act xin = src;
                if (l == 0) {
                    xin.F = c.n_freq; xin.C = 64; xin.xf = xform_identity();
                    cx.next_RT = 1; xin.data = cx.alloc_act((size_t)cx.B * c.n_freq * 64); xin.RT = 1;
                    if (!cx.dry) {
                        LstmFrameArgs f;
                        memset(&f, 0, sizeof(f));
                        f.x = src.data; f.x_RT = src.RT; f.xf = src.xf; f.layer_norm = 1;
                        f.ln_g = cx.W(m->off_ln_g); f.ln_b = cx.W(m->off_ln_b);
                        f.out = xin.data; f.step = cx.step; f.rows = cx.B * c.n_freq; f.F = c.n_freq;
                        EAB_TRY(launch_lstm_ln_frame(f, cx.st));
                    }
                }
                Act hprev;
                hprev.F = c.n_freq; hprev.C = 64; hprev.xf = xform_identity(); hprev.data = hc_state[0]; hprev.RT = 1;
                Act pair[2] = {xin, hprev};
                Act gates;
                cx.next_RT = 1;
                EAB_TRY(run_pointwise(cx, pair, 2, cx.W(m->off_rnn_step[l]), cx.W(m->off_rnn_step_b[l]), 256, 256, 0, 1, nullptr, 0, nullptr, 0,
                                      nullptr, nullptr, &gates, &m->u_rnn_step[l]));
                if (!cx.dry) {
                    LstmFrameArgs f;
                    memset(&f, 0, sizeof(f));
                    f.gates = gates.data; f.gates_ld = gates.C; f.c_state = hc_state[1]; f.h_state = hc_state[0];
                    f.out = h[l].data; f.out_RT = h[l].RT; f.step = cx.step; f.rows = cx.B * c.n_freq; f.F = c.n_freq;
                    EAB_TRY(launch_lstm_cell_frame(f, cx.st));
                }
            }
            tap(cx, l ? "h2" : "h1", h[l]);
        }
        // fused w_dnn + filter-and-sum (head_fused.cu): one read of h2, nothing else touches HBM
        if (cx.tensor_ok() && m->opt_fused_head && m->u_dnn[0].ok && m->u_dnn[1].ok && m->u_dnn[0].nsplit == 1 &&
            m->u_dnn[1].nsplit == 1 && m->u_dnn[0].ncol == 64 && m->u_dnn[1].ncol == 32 && h[1].xf.affine == 0 &&
            h[1].xf.prelu == 0 && c.M <= 16) {
            float* wtap = nullptr;
            if (m->opt_head_w_tap) wtap = cx.alloc_act((size_t)cx.B * cx.T * c.n_freq * 32);
            if (!cx.dry) {
                HeadArgs a;
                memset(&a, 0, sizeof(a));
                a.h = h[1].data; a.inpt = inpt; a.out = out_dev; a.w_out = wtap; a.w_ld = 32;
                a.rows = (long long)cx.B * cx.T * c.n_freq; a.T = cx.T; a.F = c.n_freq; a.M = c.M;
                a.W1hi = cx.W(m->u_dnn[0].off_hi[0]); a.W1lo = cx.W(m->u_dnn[0].off_lo[0]);
                a.W2hi = cx.W(m->u_dnn[1].off_hi[0]); a.W2lo = cx.W(m->u_dnn[1].off_lo[0]);
                a.b1 = cx.W(m->u_dnn[0].off_bias[0]); a.b2 = cx.W(m->u_dnn[1].off_bias[0]);
                EAB_TRY(launch_head_fused(a, cx.st));
                if (wtap) { Act wa; wa.data = wtap; wa.F = c.n_freq; wa.C = 32; tap(cx, "w", wa); }
            }
            return 0;
        }
        Act u;
        EAB_TRY(run_pointwise(cx, &h[1], 1, cx.W(m->off_dnn_w[0]), cx.W(m->off_dnn_b[0]), 64, m->dnn_N[0], 0, 1, nullptr, 1,
                              nullptr, 0, nullptr, nullptr, &u, &m->u_dnn[0]));
        EAB_TRY(run_pointwise(cx, &u, 1, cx.W(m->off_dnn_w[1]), cx.W(m->off_dnn_b[1]), 2 * c.M, m->dnn_N[1], 0, 1, nullptr, 0,
                              nullptr, 0, nullptr, nullptr, &w, &m->u_dnn[1]));
    } else {
        const int n = c.topo_type == 0 ? 2 * c.M : 2;
        EAB_TRY(run_pointwise(cx, &emb, 1, cx.W(m->off_cnn_w), cx.W(m->off_cnn_b), n, m->cnn_N, 0, 1, nullptr, 0, nullptr, 0,
                              nullptr, nullptr, &w, &m->u_cnn));
    }
    tap(cx, "w", w);
    if (!cx.dry) {
        BeamArgs a;
        memset(&a, 0, sizeof(a));
        a.step = cx.step; a.w_RT = w.RT; a.inpt_RT = inpt_RT;
        a.w = w.data; a.w_ld = w.C; a.inpt = inpt; a.B = cx.B; a.T = cx.T; a.F = c.n_freq; a.M = c.M; a.miso = c.topo_type == 1;
        a.out = out_dev;
        EAB_TRY(launch_beam(a, cx.st));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- GaGNet post-filter
// in_conv_main(cat(feat, pre)) * sigmoid(in_conv_gate(cat(feat, pre)))  (GaGNet.py:189-191, 249-251): d_feat/64 gated
// tcgen05 launches over K = d_feat + KP (two sources, the concat is never materialised)
int run_gag_in(Ctx& cx, const GagIn& in, const Act& feat, const Act& pre, Act* out) {
    const eab_config& c = cx.m->cfg;
    out->F = 1; out->C = c.d_feat; out->xf = xform_identity();
    out->data = cx.alloc_act((size_t)cx.B * cx.T * c.d_feat);
    out->RT = cx.last_RT;
    if (!cx.tensor_ok() && !cx.stream_umma()) {
        // CUDA-core path (streaming with stream_umma = 0, option umma = 0): the dense [K][value | gate] matrix of every column split; the split's
        // SW output channels are position `sp` of a [.., nsplit, SW] view of the d_feat-wide row
        if (cx.dry) return 0;
        for (int sp = 0; sp < in.nsplit; ++sp) {
            if (!in.u[sp].ok) return fail("internal: GaGNet input conv bias images missing");
            ConvArgs a;
            memset(&a, 0, sizeof(a));
            a.nsrc = 2;
            const Act* srcs[2] = {&feat, &pre};
            for (int i = 0; i < 2; ++i) { a.src[i].x = srcs[i]->data; a.src[i].C = srcs[i]->C; a.src[i].xf = srcs[i]->xf; a.src[i].RT = srcs[i]->RT; }
            a.step = cx.step; a.start = cx.start; a.out_RT = out->RT;
            a.B = cx.B; a.T = cx.T; a.Fin = 1; a.E = 1; a.Fout = in.nsplit;
            a.in_stride = 1; a.out_stride = 1; a.out_off = sp;
            a.ntaps = 1; a.dt[0] = 0; a.df[0] = 0;
            a.W = cx.W(in.off_dense[sp]); a.bias = cx.W(in.u[sp].off_bias[0]);
            a.Cout = in.SW; a.N = 2 * in.SW; a.gate_off = in.SW;
            a.algo_frac = (float)(c.d_feat + 2 * c.n_freq) / (float)in.K;
            a.out = out->data;
            EAB_TRY(launch_conv(a, cx.st));
        }
        return 0;
    }
    for (int sp = 0; sp < in.nsplit; ++sp) {
        if (!cx.dry && !in.u[sp].ok) return fail("internal: GaGNet input conv images missing");
        UmmaConvArgs u;
        memset(&u, 0, sizeof(u));
        u.nsrc = 2;
        set_src(u.src[0], feat);
        set_src(u.src[1], pre);
        u.B = cx.B; u.T = cx.T; u.Fin = 1; u.Fout = 1; u.E = 1;
        u.in_stride = 1; u.out_stride = 1; u.out_off = 0;
        u.ntaps = 1; u.dt[0] = 0; u.df[0] = 0;
        u.nslab = in.K / 64; u.ncoef = in.K; u.npass = 3;
        u.Whi = cx.W(in.u[sp].off_hi[0]); u.Wlo = cx.W(in.u[sp].off_lo[0]); u.bias = cx.W(in.u[sp].off_bias[0]);
        u.Cout = in.SW; u.N = 2 * in.SW; u.gate_off = in.SW;
        u.algo_frac = (float)(c.d_feat + 2 * c.n_freq) / (float)in.K;
        u.out = out->data; u.out_ld = c.d_feat; u.out_coff = sp * in.SW;
        u.tiles_per_b = (cx.T + 127) / 128;
        if (feat.C + pre.C != in.K || !umma_conv_supported(u)) return fail("internal: GaGNet input conv rejected by the tcgen05 path");
        if (cx.streaming) EAB_TRY(run_umma_stream(cx, &u, 1, out->RT, 0));
        else EAB_TRY(run_tensor_convs(cx, &u, 1));
    }
    return 0;
}

int run_gag_lin(Ctx& cx, const GagLin& l, const Act& x, Act* out) {
    const eab_config& c = cx.m->cfg;
    return run_pointwise(cx, &x, 1, cx.W(l.off_w), cx.W(l.off_b), c.n_freq, l.N, 0, 1, nullptr, 0, nullptr, 0, nullptr, nullptr,
                         out, &l.u);
}

// 1-3 equally long chains of single-branch TCMs as one cooperative launch (tcm_chain.cu)
bool tcm_chain_ok(Ctx& cx, const ChainRef* chains, int nch) {
    const eab_config& c = cx.m->cfg;
    if (!cx.m->opt_tcm_chain || !cx.tensor_ok() || c.cd1 != 64 || c.d_feat != 256 || c.kd1 > 8) return false;
    const bool gated = !chains[0].l[0].single;
    if (gated ? (c.kd1 != 3 && c.kd1 != 5) : c.kd1 > 4) return false;
    const size_t nl = (size_t)chains[0].n;
    if (nl < 1 || nch < 1 || nch > 3 || nch * nl > (size_t)kMaxChainLayers) return false;
    for (int i = 0; i < nch; ++i) {
        if ((size_t)chains[i].n != nl) return false;
        for (int li = 0; li < chains[i].n; ++li) {
            const TcmLayer& t = chains[i].l[li];
            // (planning runs before the weights are packed: the image flags are only known once committed)
            if (t.single == gated) return false;
            if (!cx.m->dirty && !(t.u_in.ok && t.u_out.ok && t.u_out.nsplit == 2 && t.u_in.nslab == 4)) return false;
            if (!cx.m->dirty && !(gated ? (t.u_dl.ok && t.u_dr.ok) : t.u_dil.ok)) return false;
            for (int k = 0; k < c.kd1; ++k) if (t.dt[k] > 30000 || t.dt[k] < -30000) return false;
        }
    }
    return true;
}

int run_tcm_chains(Ctx& cx, const ChainRef* chains, int nch, const Act* ins, Act* outs) {
    const eab_config& c = cx.m->cfg;
    const int nl = chains[0].n;
    const size_t rows = (size_t)cx.B * cx.T;
    TcmChainArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < nch; ++i) {
        if (ins[i].xf.affine != 0 || ins[i].xf.prelu != 0 || ins[i].data2 || ins[i].C != 256)
            return fail("internal: TCM chain input must be a plain fp32 [B,T,256] tensor");
        outs[i].F = 1; outs[i].C = c.d_feat; outs[i].xf = xform_identity();
        outs[i].data = cx.alloc_act(rows * 256);
        outs[i].RT = cx.last_RT;
        a.x_in[i] = ins[i].data;
        a.x_buf[i] = outs[i].data;
    }
    const size_t scope = cx.mark();
    for (int i = 0; i < nch; ++i) { a.y[i] = cx.alloc_act(rows * 64); a.z[i] = cx.alloc_act(rows * 64); }
    cx.release(scope);                                   // y / z die with the launch
    const bool in_stats = c.norm_type == 0;
    double* sbase = reinterpret_cast<double*>(cx.base);
    a.barrier = reinterpret_cast<unsigned*>(cx.alloc_stats(1));
    for (int i = 0; i < nch; ++i)
        for (int l = 0; l < nl; ++l) {
            const TcmLayer& t = chains[i].l[l];
            TcmChainLayer& L = a.L[i * nl + l];
            L.win_hi = (unsigned)t.u_in.off_hi[0]; L.win_lo = (unsigned)t.u_in.off_lo[0];
            const bool gated = !t.single;
            L.wd_hi = (unsigned)(gated ? t.u_dl : t.u_dil).off_hi[0]; L.wd_lo = (unsigned)(gated ? t.u_dl : t.u_dil).off_lo[0];
            if (gated) {
                L.wr_hi = (unsigned)t.u_dr.off_hi[0]; L.wr_lo = (unsigned)t.u_dr.off_lo[0];
                L.sc_r = (unsigned)t.na_right.off_scale; L.sh_r = (unsigned)t.na_right.off_shift; L.al_r = (unsigned)t.na_right.off_alpha;
                if (in_stats) L.st_r = (unsigned)(cx.alloc_stats(64) - sbase);
            }
            for (int sp = 0; sp < 2; ++sp) { L.wo_hi[sp] = (unsigned)t.u_out.off_hi[sp]; L.wo_lo[sp] = (unsigned)t.u_out.off_lo[sp]; }
            L.sc_d = (unsigned)t.na_left.off_scale; L.sh_d = (unsigned)t.na_left.off_shift; L.al_d = (unsigned)t.na_left.off_alpha;
            L.sc_o = (unsigned)t.na_out.off_scale; L.sh_o = (unsigned)t.na_out.off_shift; L.al_o = (unsigned)t.na_out.off_alpha;
            if (in_stats) {
                L.st_d = (unsigned)(cx.alloc_stats(64) - sbase);
                L.st_o = (unsigned)(cx.alloc_stats(64) - sbase);
            }
            for (int k = 0; k < c.kd1; ++k) L.dt[k] = (short)t.dt[k];
        }
    if (cx.dry) return 0;
    a.blob = cx.m->blob; a.stats = sbase;
    a.nchains = nch; a.nlayers = nl; a.kd = c.kd1; a.B = cx.B; a.T = cx.T;
    a.gated = chains[0].l[0].single ? 0 : 1;
    a.instance_norm = in_stats ? 1 : 0; a.inv_count = 1.f / (float)cx.T;
    if (cx.m->opt_dbg_launch == -200 && cx.m->dbg_buf) a.dbg = cx.m->dbg_buf;
    a.no_cluster = cx.m->opt_tcm_chain == 3 ? 1 : 0;
    if (cx.m->opt_tcm_chain == 2 && nch > 1) {
        // one chain per launch: a single chain's residual stream + scratch (59 MB at 64 x 6 s) stays in the 126 MB L2
        for (int i = 0; i < nch; ++i) {
            TcmChainArgs s = a;
            s.nchains = 1;
            s.x_in[0] = a.x_in[i]; s.x_buf[0] = a.x_buf[i]; s.y[0] = a.y[i]; s.z[0] = a.z[i];
            for (int l = 0; l < nl; ++l) s.L[l] = a.L[i * nl + l];
            s.barrier = a.barrier + 4 * i;
            EAB_TRY(launch_tcm_chain(s, cx.st));
        }
        return 0;
    }
    return launch_tcm_chain(a, cx.st);
}

// GaGNet.forward (GaGNet.py:75-89).  inpt through strides sb, sc, st, sf (floats); pre [B,2,T,F]; out [q][B,2,T,F].
int run_gag_forward(Ctx& cx, const float* inpt, const long long* strides, const float* pre_in, float* out_dev) {
    eab_model* m = cx.m;
    const eab_config& c = m->cfg;
    const int F = c.n_freq, KP = ceil64(2 * F);
    Act x;
    x.F = F; x.C = 4; x.xf = xform_identity();
    x.data = cx.alloc_act((size_t)cx.B * cx.T * F * 4);          // streaming: a ring of 2 frames (the kt = 2 first conv)
    x.RT = cx.last_RT;
    Act pre;
    pre.F = 1; pre.C = KP; pre.xf = xform_identity();
    if (cx.streaming) cx.next_RT = 1;                            // read by 1x1 convs and the elementwise kernels only
    pre.data = cx.alloc_act((size_t)cx.B * cx.T * KP);
    pre.RT = cx.last_RT;
    if (!cx.dry) {
        GagPackArgs a;
        memset(&a, 0, sizeof(a));
        a.inpt = inpt; a.sb = strides[0]; a.sc = strides[1]; a.st = strides[2]; a.sf = strides[3];
        a.pre = pre_in; a.x4 = x.data; a.pre_row = pre.data;
        a.B = cx.B; a.T = cx.T; a.F = F; a.KP = KP; a.KP2 = KP / 2;
        if (cx.streaming) { a.step = cx.step; a.in_RT = cx.gag_in_RT; a.in_slot = cx.gag_in_slot; a.x_RT = x.RT; }
        EAB_TRY(launch_gag_pack(a, cx.st));
    }
    // ---------------- encoder (GaGNet.py:361-365 / :408-412): only the bottleneck is used
    if (c.is_u2) {
        for (size_t i = 0; i < m->en_mod.size(); ++i) {
            Act y;
            EAB_TRY(run_module(cx, m->en_mod[i], &x, 1, &y));
            tap(cx, ("en." + std::to_string(i)).c_str(), y);
            x = y;
        }
        Act y;
        EAB_TRY(run_conv2d(cx, m->en_last, &x, 1, &y));
        x = y;
    } else {
        for (size_t i = 0; i < m->en_plain.size(); ++i) {
            Act y;
            EAB_TRY(run_conv2d(cx, m->en_plain[i], &x, 1, &y));
            x = y;
        }
    }
    tap(cx, "en.4", x);
    if (x.F != m->Fb || x.C != 64) return fail("internal: bottleneck shape");
    Act feat;                                             // [B,T,1,d_feat] finalised, channel f*64+c
    EAB_TRY(run_combine(cx, &x, 1, &feat));
    feat.F = 1; feat.C = c.d_feat; feat.xf = xform_identity();
    // ---------------- glance-gaze modules (GaGNet.py:84-88, 120-134)
    const size_t stage_elems = (size_t)cx.B * 2 * cx.T * F;
    for (size_t gi = 0; gi < m->gags.size(); ++gi) {
        const GagModule& G = m->gags[gi];
        const bool last = gi + 1 == m->gags.size();
        Act next;
        next.F = 1; next.C = KP; next.xf = xform_identity();
        if (cx.streaming) cx.next_RT = 1;
        next.data = last ? nullptr : cx.alloc_act((size_t)cx.B * cx.T * KP);      // outlives the module's scratch
        next.RT = cx.last_RT;
        const size_t scope = cx.mark();
        Act xg, xz;
        EAB_TRY(run_gag_in(cx, G.in_g, feat, pre, &xg));
        EAB_TRY(run_gag_in(cx, G.in_z, feat, pre, &xz));
        tap(cx, ("g.in_g." + std::to_string(gi)).c_str(), xg);
        Act xr = xz, xi = xz;
        {
            const ChainRef chains[3] = {{G.tcn_g.data(), (int)G.tcn_g.size()}, {G.tcm_r.data(), (int)G.tcm_r.size()},
                                        {G.tcm_i.data(), (int)G.tcm_i.size()}};
            const int nch = m->gcfg.is_squeezed ? 2 : 3;
            if (tcm_chain_ok(cx, chains, nch)) {
                Act ins[3] = {xg, xz, xz}, outs[3];
                EAB_TRY(run_tcm_chains(cx, chains, nch, ins, outs));
                xg = outs[0]; xr = outs[1]; xi = nch == 3 ? outs[2] : outs[1];
            } else {
                for (const TcmLayer& t : G.tcn_g) { Act nx; EAB_TRY(run_tcm(cx, t, xg, &nx)); xg = nx; }
                for (const TcmLayer& t : G.tcm_r) { Act nx; EAB_TRY(run_tcm(cx, t, xr, &nx)); xr = nx; }
                if (m->gcfg.is_squeezed) xi = xr;
                else for (const TcmLayer& t : G.tcm_i) { Act nx; EAB_TRY(run_tcm(cx, t, xi, &nx)); xi = nx; }
            }
        }
        tap(cx, ("g.tcn_g." + std::to_string(gi)).c_str(), xg);
        Act gain;
        if (cx.streaming) cx.next_RT = 1;                        // (gag_crm_kernel reads plain rows)
        EAB_TRY(run_gag_lin(cx, G.lin_g, xg, &gain));
        Act rr, ri;
        if (cx.streaming) cx.next_RT = 1;
        EAB_TRY(run_gag_lin(cx, G.lin_r, xr, &rr));
        if (cx.streaming) cx.next_RT = 1;
        EAB_TRY(run_gag_lin(cx, G.lin_i, xi, &ri));
        tap(cx, ("g.gain." + std::to_string(gi)).c_str(), gain);
        tap(cx, ("g.res_r." + std::to_string(gi)).c_str(), rr);
        if (rr.C != ri.C) return fail("internal: GaGNet residual widths differ");
        if (!cx.dry) {
            GagCrmArgs a;
            memset(&a, 0, sizeof(a));
            a.pre_row = pre.data; a.gain = gain.data; a.res_r = rr.data; a.res_i = ri.data;
            a.ld_g = gain.C; a.ld_r = rr.C; a.acti = m->gcfg.acti_type;
            a.next_row = next.data; a.out = out_dev + gi * stage_elems;
            a.B = cx.B; a.T = cx.T; a.F = F; a.KP = KP; a.KP2 = KP / 2;
            EAB_TRY(launch_gag_crm(a, cx.st));
        }
        cx.release(scope);
        pre = next;
    }
    return 0;
}

int plan(eab_model* m, int B, int T, size_t* stats_bytes, size_t* total_bytes) {
    Ctx cx;
    cx.m = m; cx.dry = true; cx.base = nullptr; cx.B = B; cx.T = T; cx.st = nullptr;
    if (m->kind == 1) {
        const long long zero[4] = {0, 0, 0, 0};
        EAB_TRY(run_gag_forward(cx, nullptr, zero, nullptr, nullptr));
    } else
    EAB_TRY(run_forward(cx, nullptr, nullptr));
    *stats_bytes = cx.stats_off;
    *total_bytes = cx.stats_off + cx.act_peak;
    return 0;
}

// ---------------------------------------------------------------------------------------------- streaming
// State blob layout (device, caller-owned): [0,256) absolute frame counter | carried hop [S][M][160] | iSTFT tail
// [S][160] | spectrum ring [S][2][F][M][2] | output frame [S][2][F] | activation rings + LSTM state (run_forward order)

int stream_layout(eab_model* m, int S, StreamLayout* L) {
    const eab_config& c = m->cfg;
    if (S < 1) return fail("stream: need at least one stream");
    if (c.norm_type != 1) return fail("streaming needs norm_type='BN': InstanceNorm statistics span the whole utterance (EaBNet.py:684-686)");
    if (!c.is_causal) return fail("streaming needs is_causal=True");
    size_t o = 256;
    L->off_start = o; o += up256((size_t)S * sizeof(int));
    if (m->kind == 1) {
        // GaGNet post-filter: [0,256) frame counter | start | estimates of the q modules [q][S][2][F] | activation rings
        L->off_desc = L->off_prev = L->off_tail = L->off_spec = o;
        L->off_out = o; o += up256(m->gags.size() * (size_t)S * 2 * c.n_freq * sizeof(float));
        L->off_act = o;
        Ctx cx;
        cx.m = m; cx.dry = true; cx.base = nullptr; cx.B = S; cx.T = 1; cx.st = nullptr; cx.streaming = true;
        const long long zero[4] = {0, 0, 0, 0};
        EAB_TRY(run_gag_forward(cx, nullptr, zero, nullptr, nullptr));
        if (cx.stats_off != 0) return fail("internal: streaming plan allocated statistics");
        L->total = o + cx.act_peak;
        return 0;
    }
    L->off_desc = o; o += up256(m->tcms.size() * sizeof(TcmStreamDesc));
    L->off_prev = o; o += up256((size_t)S * c.M * 160 * sizeof(float));
    L->off_tail = o; o += up256((size_t)S * 160 * sizeof(float));
    L->off_spec = o; o += up256((size_t)S * 2 * c.n_freq * c.M * 2 * sizeof(float));
    L->off_out = o;  o += up256((size_t)S * 2 * c.n_freq * sizeof(float));
    L->off_act = o;
    Ctx cx;
    cx.m = m; cx.dry = true; cx.base = nullptr; cx.B = S; cx.T = 1; cx.st = nullptr; cx.streaming = true;
    cx.tcm_desc = &L->descs;
    cx.per_stream = &L->per_stream;
    EAB_TRY(run_forward(cx, nullptr, nullptr));
    if (cx.stats_off != 0) return fail("internal: streaming plan allocated statistics");
    L->total = o + cx.act_peak;
    return 0;
}

int stream_forward(eab_model* m, char* state, const StreamLayout& L, int S, cudaStream_t st) {
    m->umma_launch_idx = 0;
    if (m->dirty) return fail("parameters not committed: call eab_commit_params first");
    Ctx cx;
    cx.m = m; cx.dry = false; cx.base = state + L.off_act; cx.B = S; cx.T = 1; cx.st = st;
    cx.streaming = true; cx.step = reinterpret_cast<const int*>(state);
    cx.start = reinterpret_cast<const int*>(state + L.off_start);
    cx.tcm_desc_dev = reinterpret_cast<const TcmStreamDesc*>(state + L.off_desc);
    return run_forward(cx, reinterpret_cast<const float*>(state + L.off_spec), reinterpret_cast<float*>(state + L.off_out));
}

// one frame of every stream through the post-filter: inpt element (s, ri, f) at inpt[s*strides[0] + ri*strides[1] + f*strides[3]
// + slot*in_slot] (slot = *step % in_RT: the spectrum ring of an EaBNet stream state, or in_RT = 1 for a plain frame);
// pre [S][2][F]; the q estimates go to the state's [q][S][2][F] block
int gag_stream_forward(eab_model* m, char* state, const StreamLayout& L, const float* inpt, const long long* strides, int in_RT,
                       long long in_slot, const float* pre, int S, cudaStream_t st) {
    Ctx cx;
    cx.m = m; cx.dry = false; cx.base = state + L.off_act; cx.B = S; cx.T = 1; cx.st = st;
    cx.streaming = true; cx.step = reinterpret_cast<const int*>(state);
    cx.start = reinterpret_cast<const int*>(state + L.off_start);
    cx.gag_in_RT = in_RT; cx.gag_in_slot = in_slot;
    return run_gag_forward(cx, inpt, strides, pre, reinterpret_cast<float*>(state + L.off_out));
}

int gag_forward(eab_model* m, const float* inpt, const long long* strides, const float* pre, float* out, int B, int T, void* ws,
                size_t ws_bytes, cudaStream_t st) {
    if (m->kind != 1) return fail("eab_gag_forward: the handle is not a GaGNet (create it with eab_gag_create)");
    if (B < 1 || T < 1) return fail("forward: B and T must be positive");
    if (m->cfg.norm_type == 0 && T < 2) return fail("InstanceNorm1d needs more than one frame (the reference raises too)");
    if (m->dirty) return fail("parameters not committed: call eab_commit_params first");
    size_t sb = 0, tb = 0;
    EAB_TRY(plan(m, B, T, &sb, &tb));
    if (ws_bytes < tb) return fail("workspace too small: need " + std::to_string(tb) + " bytes");
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail("workspace must be 256-byte aligned");
    m->taps.clear();
    m->norm_log.clear();
    m->umma_launch_idx = 0;
    if (sb) EAB_CUDA(cudaMemsetAsync(ws, 0, sb, st));
    Ctx cx;
    cx.m = m; cx.dry = false; cx.base = static_cast<char*>(ws); cx.B = B; cx.T = T; cx.st = st;
    cx.stats_off = 0; cx.stats_cap = sb; cx.act_off = sb; cx.act_peak = sb;
    return run_gag_forward(cx, inpt, strides, pre, out);
}

int forward(eab_model* m, const float* inpt, float* out, int B, int T, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (m->kind != 0) return fail("this entry point needs an EaBNet handle (eab_create); a GaGNet runs through eab_gag_forward");
    if (B < 1 || T < 1) return fail("forward: B and T must be positive");
    if (m->cfg.norm_type == 0 && T < 2)
        return fail("InstanceNorm1d needs more than one frame (the reference raises too)");
    if (m->dirty) return fail("parameters not committed: call eab_commit_params first");
    size_t sb = 0, tb = 0;
    EAB_TRY(plan(m, B, T, &sb, &tb));
    if (ws_bytes < tb) return fail("workspace too small: need " + std::to_string(tb) + " bytes");
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail("workspace must be 256-byte aligned");
    m->taps.clear();
    m->umma_launch_idx = 0;
    if (sb) EAB_CUDA(cudaMemsetAsync(ws, 0, sb, st));
    Ctx cx;
    cx.m = m; cx.dry = false; cx.base = static_cast<char*>(ws); cx.B = B; cx.T = T; cx.st = st;
    cx.stats_off = 0; cx.stats_cap = sb; cx.act_off = sb; cx.act_peak = sb;
    return run_forward(cx, inpt, out);
}


}  // namespace detail
}  // namespace eab
