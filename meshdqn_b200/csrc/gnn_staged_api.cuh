// gnn_staged_api.cuh -- host side of the staged tensor-core Q-path (included by gnn_fused.cu after its kernels).
namespace {

struct StgGeom {
    int B, max_n, max_e, R1, R2, EC, GS1, GS2;
};

struct StgWs {
    float *x1, *r0, *r1, *x2, *x3;
    unsigned short *e1, *e2;
    int *e1n, *e2n;
    float *h1k, *s1k, *z1k, *h2k, *s2k, *z2k, *dX2, *dR, *lterm;
    unsigned char *amax1, *amax2, *perm2;
    int64_t floats;
};

int stg_env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}

void stg_geom(const mdq_net_t &net, int B, int max_n, int max_e, StgGeom &G)
{
    G.B = B; G.max_n = max_n; G.max_e = max_e;
    G.R1 = topk_count(net.ratio, max_n);
    G.R2 = topk_count(net.ratio, G.R1);
    G.EC = (max_e > 0 ? max_e : 1);
    G.EC = (G.EC + 7) & ~7;
    // graphs per tail CTA: 4 for replay minibatches (64 CTAs beside the other kernels of the step), 8 once the batch alone
    // fills the machine (8192 candidates: 7.3 vs 6.6 M graphs/s)
    int gs1 = stg_env_int("MDQ_STG_GS1", 2), gs2 = stg_env_int("MDQ_STG_GS2", B >= 1024 ? 8 : 4);
    while (gs1 > 1 && gs1 * G.R1 > 128) --gs1;
    while (gs2 > 1 && gs2 * G.R2 > 32) --gs2;
    G.GS1 = gs1 < 1 ? 1 : gs1;
    G.GS2 = gs2 < 1 ? 1 : gs2;
}

// carve the staged sections out of `base` (floats); wd != nullptr: backward mode (x2 / x3 live in the weight-gradient rows)
void stg_carve(const StgGeom &G, float *base, bool bwd, StgWs &w)
{
    int64_t o = 0;
    auto takef = [&](int64_t n) { float *p = base ? base + o : nullptr; o += (n + 3) & ~(int64_t)3; return p; };
    const int64_t B = G.B;
    w.x1 = takef(B * G.R1 * 128);
    w.r0 = takef(B * 256);
    w.r1 = takef(B * 256);
    w.e1 = reinterpret_cast<unsigned short *>(takef((B * G.EC + 1) / 2));
    w.e2 = reinterpret_cast<unsigned short *>(takef((B * G.EC + 1) / 2));
    w.e1n = reinterpret_cast<int *>(takef(B));
    w.e2n = reinterpret_cast<int *>(takef(B));
    if (!bwd) {
        w.x2 = takef(B * G.R2 * 128);
        w.x3 = takef(B * 128);
        w.h1k = w.s1k = w.z1k = w.h2k = w.s2k = w.z2k = w.dX2 = w.dR = w.lterm = nullptr;
        w.amax1 = w.amax2 = w.perm2 = nullptr;
    } else {
        w.x2 = w.x3 = nullptr;
        w.h1k = takef(B * G.R1 * 128);
        w.s1k = takef(B * G.R1);
        w.z1k = takef(B * G.R1);
        w.h2k = takef(B * G.R2 * 128);
        w.s2k = takef(B * G.R2);
        w.z2k = takef(B * G.R2);
        w.dX2 = takef(B * G.R2 * 128);
        w.dR = takef(B * 256);
        w.lterm = takef(B);
        w.amax1 = reinterpret_cast<unsigned char *>(takef(B * 32));
        w.amax2 = reinterpret_cast<unsigned char *>(takef(B * 32));
        w.perm2 = reinterpret_cast<unsigned char *>(takef((B * G.R2 + 3) / 4));
    }
    w.floats = o;
}

template <typename K>
int stg_smem_attr(K kern, int bytes, const char *name)
{
    if (bytes > 227 * 1024) {
        mdq::set_error("%s: %d bytes of shared memory per CTA (> 227 KB)", name, bytes);
        return MDQ_ESMEM;
    }
    // remembered per kernel (function address): the attribute calls cost ~10 us of host time each, 13 launches per step
    static std::mutex mu;
    static std::map<const void *, int> done;
    const void *key = reinterpret_cast<const void *>(kern);
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = done.find(key);
        if (it != done.end() && it->second >= bytes) return MDQ_OK;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { mdq::set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e)); return MDQ_ECUDA; }
    {
        std::lock_guard<std::mutex> lk(mu);
        done[key] = bytes;
    }
    return MDQ_OK;
}

struct StgCall {
    const mdq_net_t *net;
    const float *params, *wsplit, *x;
    const void *esrc, *edst;      // int64 (PyG edge_index rows) or int32 when edge_i32
    int edge_i32;
    const int32_t *nptr, *eptr;
    float *out, *emb;
    int32_t *amax;
};

// stages 0 and 1 (SAVE: with the backward saves); fills s2 with everything stage 2 needs from them
template <bool SAVE>
int stg_launch_01(const StgCall &c, const StgGeom &G, const stg::Plan &P, const StgWs &w, float *c1k, float *c2k, float *x2,
                  cudaStream_t st)
{
    const mdq_net_t &net = *c.net;
    {
        stg::S0Args a;
        memset(&a, 0, sizeof(a));
        a.trace = g_trace;
        a.params = c.params; a.wsplit = c.wsplit;
        a.w_off = (unsigned)P.m[stg::M_C1].off;
        a.pf_floats = (unsigned)P.total; a.n_params = (unsigned)net.n_params;
        a.F = P.F; a.Fp = P.Fp; a.nch = P.k1pad / 4;
        a.x_stride = net.x_stride; a.col0 = net.in_col0; a.b_off = net.blk[0].b_off; a.pool_off = net.blk[0].pool_off;
        a.ratio = net.ratio;
        a.x = c.x; a.esrc = (const long long *)c.esrc; a.edst = (const long long *)c.edst; a.edge_i32 = c.edge_i32;
        a.nptr = c.nptr; a.eptr = c.eptr;
        a.B = G.B; a.R1 = G.R1; a.EC1 = G.EC;
        a.x1 = w.x1; a.e1 = w.e1; a.e1n = w.e1n; a.r0 = w.r0;
        a.h1k = w.h1k; a.c1k = c1k; a.s1k = w.s1k; a.z1k = w.z1k; a.amax1 = w.amax1;
        const int bytes = stg::s0_layout(a, G.max_n, G.max_e, G.R1);
        if (bytes < 0) { mdq::set_error("staged stage 0: layout"); return MDQ_ESMEM; }
        int rc = stg_smem_attr(stg::k_stage0<SAVE>, bytes, "k_stage0");
        if (rc != MDQ_OK) return rc;
        stg::k_stage0<SAVE><<<G.B, stg::NTH, bytes, st>>>(a);
        if ((rc = mdq::check_launch("k_stage0")) != MDQ_OK) return rc;
    }
    {
        stg::S1Args a;
        memset(&a, 0, sizeof(a));
        a.trace = g_trace;
        a.params = c.params; a.wsplit = c.wsplit;
        a.nblk = stg::add_blocks(P.m[stg::M_C2F], 0, a.blk, 0);
        a.b_off = net.blk[1].b_off; a.pool_off = net.blk[1].pool_off; a.ratio = net.ratio;
        a.nptr = c.nptr; a.B = G.B; a.GS = G.GS1; a.R1 = G.R1; a.EC1 = G.EC; a.R2 = G.R2; a.EC2 = G.EC;
        a.NRMAX = G.GS1 * G.R1;
        a.x1 = w.x1; a.e1 = w.e1; a.e1n = w.e1n;
        a.x2 = x2; a.e2 = w.e2; a.e2n = w.e2n; a.r1 = w.r1;
        a.h2k = w.h2k; a.c2k = c2k; a.s2k = w.s2k; a.z2k = w.z2k; a.perm2 = w.perm2; a.amax2 = w.amax2;
        int bytes = -1;
        for (a.sta = stg_env_int("MDQ_STG_STA", 5); a.sta >= 2; --a.sta) {   // deepest weight ring that fits
            bytes = stg::s1_layout(a);
            if (bytes > 0 && bytes <= 227 * 1024) break;
        }
        if (bytes < 0 || a.sta < 2) { mdq::set_error("staged stage 1: layout"); return MDQ_ESMEM; }
        int rc = stg_smem_attr(stg::k_stage1<SAVE>, bytes, "k_stage1");
        if (rc != MDQ_OK) return rc;
        stg::k_stage1<SAVE><<<(G.B + G.GS1 - 1) / G.GS1, stg::NTH, bytes, st>>>(a);
        if ((rc = mdq::check_launch("k_stage1")) != MDQ_OK) return rc;
    }
    return MDQ_OK;
}

void stg_tail_common(stg::TArgs &a, const StgCall &c, const StgGeom &G, const stg::Plan &P, const StgWs &w, const float *x2,
                     float *x3, bool bwd)
{
    const mdq_net_t &net = *c.net;
    memset(&a, 0, sizeof(a));
    a.trace = g_trace;
    a.params = c.params; a.wsplit = c.wsplit;
    int n = 0;
    n = stg::add_rblocks(net.blk[2].w_off, 128, 128, a.blk, n);
    n = stg::add_rblocks(net.blk[3].w_off, 128, 128, a.blk, n);
    n = stg::add_rblocks(net.lin_off[0], 256, 128, a.blk, n);
    n = stg::add_rblocks(net.lin_off[1], 128, 64, a.blk, n);
    n = stg::add_rblocks(net.lin_off[2], 64, net.out_dim, a.blk, n);
    if (bwd) {
        // transposed copies [out][in]: rows = out features (the reduction axis of the backward products)
        n = stg::add_rblocks(P.t[stg::T_L3].off, net.out_dim, 64, a.blk, n, true);
        n = stg::add_rblocks(P.t[stg::T_L2].off, 64, 128, a.blk, n, true);
        n = stg::add_rblocks(P.t[stg::T_L1].off, 128, 256, a.blk, n, true);
        n = stg::add_rblocks(P.t[stg::T_C5].off, 128, 128, a.blk, n, true);
        n = stg::add_rblocks(P.t[stg::T_C4].off, 128, 128, a.blk, n, true);
    }
    a.nblk = n;
    a.b4_off = net.blk[2].b_off; a.p4_off = net.blk[2].pool_off; a.b5_off = net.blk[3].b_off; a.p5_off = net.blk[3].pool_off;
    for (int i = 0; i < 3; ++i) a.lb_off[i] = net.lin_boff[i];
    a.A = net.out_dim; a.softmax = net.softmax;
    a.ratio = net.ratio; a.nptr = c.nptr; a.B = G.B; a.GS = G.GS2; a.R2 = G.R2; a.EC2 = G.EC;
    a.x2 = x2; a.e2 = w.e2; a.e2n = w.e2n; a.r0 = w.r0; a.r1 = w.r1; a.x3 = x3;
    a.out = c.out; a.emb = c.emb; a.amax_out = c.amax;
}

template <bool BWD>
int stg_launch_tail(stg::TArgs &a, const StgGeom &G, cudaStream_t st)
{
    const int bytes = stg::tail_layout(a);
    int rc = stg_smem_attr(stg::k_tail<BWD>, bytes, "k_tail");
    if (rc != MDQ_OK) return rc;
    stg::k_tail<BWD><<<(G.B + G.GS2 - 1) / G.GS2, stg::NTH_TAIL, bytes, st>>>(a);
    return mdq::check_launch("k_tail");
}

// phase 0: everything; 1: stages 0 / 1 only (independent of Q_other: may run beside the other net's forward);
// 2: the rest (tail backward, backward 1, weight gradients); 3: stages 0 / 1 + tail backward (with `sync`: launched before
// Q_other exists, the tail waits for mdq_stream_post where the loss needs it); 4: backward 1 + weight gradients.  zero_grad: memset the flat gradient first (the unused
// blocks' entries); the replay path keeps a persistent, pre-zeroed gradient buffer instead.
struct StgLoss { int batch; const int32_t *next_slot; float *loss; };
int stg_backward_launch(const StgCall &c, int B, int max_n, int max_e, stg::TArgs &a2, float *grad, float *workspace,
                        cudaStream_t st, int phase, bool zero_grad, const StgLoss &ls, unsigned *sync = nullptr)
{
    const bool do_stages = phase == 0 || phase == 1 || phase == 3, do_tail = phase == 0 || phase == 2 || phase == 3;
    const bool do_rest = phase == 0 || phase == 2 || phase == 4;
    const mdq_net_t &net = *c.net;
    if (!stg::supported(net, max_n, max_e)) { mdq::set_error("staged path: unsupported network / graph size"); return MDQ_EINVAL; }
    stg::Plan P;
    stg::build_plan(net, P);
    StgGeom G;
    stg_geom(net, B, max_n, max_e, G);
    WDesc wd;
    build_wdesc(net, B, max_n, wd);
    if (wd.l[0].rpg != G.R1 || wd.l[1].rpg != G.R2 || wd.l[2].rpg != G.R2 || wd.l[3].rpg != 1 || wd.l[0].K != 2 * P.F) {
        mdq::set_error("staged path: weight-gradient row geometry mismatch");
        return MDQ_EINVAL;
    }
    int n_tasks = 0;
    for (int i = 0; i < wd.nl; ++i) {
        const WLayer &l = wd.l[i];
        wd.tstart[i] = n_tasks;
        if (l.rpg == 0) continue;
        const int S = (B * l.rpg + WG_TASK - 1) / WG_TASK;
        const int nkg = (l.K + 1 + WG_KG - 1) / WG_KG;
        n_tasks += S * nkg;
    }
    wd.tstart[wd.nl] = n_tasks;
    const int64_t fused_ws = mdq_qnet_bwd_workspace_floats(&net, B, max_n);
    float *d_partial = workspace + wd.total;
    StgWs w;
    stg_carve(G, workspace + ((fused_ws + 3) & ~(int64_t)3), true, w);
    float *ws = workspace;
    if (zero_grad && do_stages) {
        cudaError_t e = cudaMemsetAsync(grad, 0, (size_t)net.n_params * sizeof(float), st);
        if (e != cudaSuccess) { mdq::set_error("memset grad: %s", cudaGetErrorString(e)); return MDQ_ECUDA; }
    }
    float *x2 = ws + wd.l[2].i_off, *x3 = ws + wd.l[3].i_off;
    int rc = MDQ_OK;
    if (do_stages) {
        rc = stg_launch_01<true>(c, G, P, w, ws + wd.l[0].i_off, ws + wd.l[1].i_off, x2, st);
        if (rc != MDQ_OK || phase == 1) return rc;
    }
    if (do_tail) {
        stg::TArgs keep = a2;   // the caller filled the loss-gradient fields
        stg_tail_common(a2, c, G, P, w, x2, x3, true);
        a2.mode = keep.mode; a2.gout = keep.gout; a2.rp_action = keep.rp_action; a2.rp_reward = keep.rp_reward;
        a2.rp_index = keep.rp_index; a2.rp_qother = keep.rp_qother; a2.rp_gamma = keep.rp_gamma;
        a2.rp_inv_batch = keep.rp_inv_batch; a2.rp_scalar = keep.rp_scalar; a2.lterm = w.lterm;
        const int nb = 4;
        for (int i = 0; i < 3; ++i) { a2.lin_in[i] = ws + wd.l[2 * nb + i].i_off; a2.lin_d[i] = ws + wd.l[2 * nb + i].d_off; }
        a2.c5_d = ws + wd.l[3].d_off; a2.c4_d = ws + wd.l[2].d_off;
        a2.pool4_d = ws + wd.l[nb + 2].d_off; a2.pool5_d = ws + wd.l[nb + 3].d_off;
        a2.bias4_d = ws + wd.l[2 * nb + 3 + 2].d_off; a2.bias5_d = ws + wd.l[2 * nb + 3 + 3].d_off;
        a2.dX2 = w.dX2; a2.dR = w.dR;
        a2.sync = sync;
        if ((rc = stg_launch_tail<true>(a2, G, st)) != MDQ_OK) return rc;
    }
    if (!do_rest) return MDQ_OK;
    {
        stg::B1Args a;
        memset(&a, 0, sizeof(a));
        a.trace = g_trace;
        a.params = c.params; a.wsplit = c.wsplit;
        a.nblk = stg::add_rblocks(P.t[stg::T_C2].off, 128, 256, a.blk, 0, true);
        a.p1_off = net.blk[0].pool_off; a.p2_off = net.blk[1].pool_off; a.ratio = net.ratio; a.nptr = c.nptr;
        a.B = B; a.GS = G.GS1; a.R1 = G.R1; a.EC1 = G.EC; a.R2 = G.R2;
        a.e1 = w.e1; a.e1n = w.e1n; a.dX2 = w.dX2; a.dR = w.dR;
        a.h2k = w.h2k; a.s2k = w.s2k; a.z2k = w.z2k; a.perm2 = w.perm2; a.amax2 = w.amax2;
        a.h1k = w.h1k; a.s1k = w.s1k; a.z1k = w.z1k; a.amax1 = w.amax1;
        a.c2_d = ws + wd.l[1].d_off; a.pool2_d = ws + wd.l[4 + 1].d_off;
        a.c1_d = ws + wd.l[0].d_off; a.pool1_d = ws + wd.l[4 + 0].d_off;
        a.mode = a2.mode; a.batch = ls.batch; a.A = net.out_dim; a.lterm = w.lterm; a.rp_qother = a2.rp_qother;
        a.rp_reward = a2.rp_reward; a.rp_action = a2.rp_action; a.next_slot = ls.next_slot; a.loss = ls.loss;
        const int bytes = stg::b1_layout(a);
        if ((rc = stg_smem_attr(stg::k_bwd1, bytes, "k_bwd1")) != MDQ_OK) return rc;
        // + 1: the loss-reduction CTA (modes 1 / 2)
        stg::k_bwd1<<<(B + G.GS1 - 1) / G.GS1 + (a.mode != 0 ? 1 : 0), stg::NTH_TAIL, bytes, st>>>(a);
        if ((rc = mdq::check_launch("k_bwd1")) != MDQ_OK) return rc;
    }
    wgrad_partial_kernel<<<n_tasks > 0 ? n_tasks : 1, 256, 0, st>>>(wd, B, workspace, d_partial);
    if ((rc = mdq::check_launch("wgrad_partial_kernel")) != MDQ_OK) return rc;
    dim3 rg(32, wd.nl);
    wgrad_reduce_kernel<<<rg, 256, 0, st>>>(wd, B, d_partial, grad);
    return mdq::check_launch("wgrad_reduce_kernel");
}

}  // namespace

extern "C" {

int mdq_qnet_staged_supported(const mdq_net_t *net, int max_n, int max_e)
{
    return (net && stg::supported(*net, max_n, max_e)) ? 1 : 0;
}

int64_t mdq_qnet_staged_wsplit_floats(const mdq_net_t *net)
{
    if (!net || !stg::supported(*net, 1, 0)) return -1;
    stg::Plan P;
    stg::build_plan(*net, P);
    return P.total;
}

int mdq_qnet_staged_wsplit(const mdq_net_t *net, const float *params, float *wsplit, void *stream)
{
    if (!net || !params || !wsplit || !stg::supported(*net, 1, 0) || (reinterpret_cast<uintptr_t>(wsplit) & 15)) {
        mdq::set_error("mdq_qnet_staged_wsplit: bad argument / unsupported network");
        return MDQ_EINVAL;
    }
    stg::Plan P;
    stg::build_plan(*net, P);
    const int items = P.item0[stg::M_COUNT];
    stg::k_wsplit<<<(items + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, params, wsplit);
    return mdq::check_launch("k_wsplit");
}

int64_t mdq_qnet_staged_workspace_floats(const mdq_net_t *net, int n_graphs, int max_n, int max_e, int backward)
{
    if (!net || n_graphs < 1 || !stg::supported(*net, max_n, max_e)) return -1;
    StgGeom G;
    stg_geom(*net, n_graphs, max_n, max_e, G);
    StgWs w;
    stg_carve(G, nullptr, backward != 0, w);
    int64_t total = w.floats + 8;
    if (backward) total += (mdq_qnet_bwd_workspace_floats(net, n_graphs, max_n) + 3) & ~(int64_t)3;
    return total;
}

int mdq_qnet_staged_forward(const mdq_net_t *net, const float *params, const float *wsplit, const float *x,
                            const void *edge_src, const void *edge_dst, int edge_i32, const int32_t *node_ptr,
                            const int32_t *edge_ptr, int n_graphs, int max_n, int max_e, float *out, float *embedding,
                            int32_t *argmax, float *workspace, void *stream)
{
    if (!net || !params || !wsplit || !x || !node_ptr || !edge_ptr || !workspace || n_graphs < 1 ||
        (reinterpret_cast<uintptr_t>(workspace) & 15)) {
        mdq::set_error("mdq_qnet_staged_forward: null / misaligned argument or empty batch");
        return MDQ_EINVAL;
    }
    if (!stg::supported(*net, max_n, max_e)) {
        mdq::set_error("mdq_qnet_staged_forward: unsupported network / graph size (use mdq_qnet_forward)");
        return MDQ_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    stg::Plan P;
    stg::build_plan(*net, P);
    StgGeom G;
    stg_geom(*net, n_graphs, max_n, max_e, G);
    StgWs w;
    stg_carve(G, workspace, false, w);
    StgCall c{net, params, wsplit, x, edge_src, edge_dst, edge_i32, node_ptr, edge_ptr, out, embedding, argmax};
    int rc = stg_launch_01<false>(c, G, P, w, nullptr, nullptr, w.x2, st);
    if (rc != MDQ_OK) return rc;
    stg::TArgs a2;
    stg_tail_common(a2, c, G, P, w, w.x2, w.x3, false);
    return stg_launch_tail<false>(a2, G, st);
}

int mdq_qnet_staged_backward(const mdq_net_t *net, const float *params, const float *wsplit, const float *x,
                             const void *edge_src, const void *edge_dst, int edge_i32, const int32_t *node_ptr,
                             const int32_t *edge_ptr, int n_graphs, int max_n, int max_e, const float *grad_out, float *grad,
                             float *workspace, void *stream)
{
    if (!net || !params || !wsplit || !x || !grad_out || !grad || !workspace || n_graphs < 1 ||
        (reinterpret_cast<uintptr_t>(workspace) & 15)) {
        mdq::set_error("mdq_qnet_staged_backward: null / misaligned argument or empty batch");
        return MDQ_EINVAL;
    }
    StgCall c{net, params, wsplit, x, edge_src, edge_dst, edge_i32, node_ptr, edge_ptr, nullptr, nullptr, nullptr};
    stg::TArgs a2;
    memset(&a2, 0, sizeof(a2));
    a2.mode = 0; a2.gout = grad_out;
    return stg_backward_launch(c, n_graphs, max_n, max_e, a2, grad, workspace, (cudaStream_t)stream, 0, true, StgLoss{0, nullptr, nullptr});
}

int mdq_qnet_staged_replay_backward(const mdq_net_t *net, const float *params, const float *wsplit, const float *x,
                                    const void *edge_src, const void *edge_dst, int edge_i32, const int32_t *node_ptr,
                                    const int32_t *edge_ptr, int n_graphs, int max_n, int max_e, int mode,
                                    const int32_t *action, const float *reward, const int32_t *index,
                                    const int32_t *next_slot, const float *q_other, int batch, float gamma, float *scalar,
                                    float *loss, float *grad, float *workspace, int phase, uint32_t *tail_sync, void *stream)
{
    if (!net || !params || !wsplit || !x || !grad || !workspace || !action || !reward || !index || !next_slot || !scalar ||
        !loss || n_graphs < 1 || batch < 1 || (mode != 1 && mode != 2) || (mode == 2 && phase != 1 && !q_other) ||
        phase < 0 || phase > 4 || (tail_sync && phase != 3) || (reinterpret_cast<uintptr_t>(workspace) & 15)) {
        mdq::set_error("mdq_qnet_staged_replay_backward: bad argument");
        return MDQ_EINVAL;
    }
    StgCall c{net, params, wsplit, x, edge_src, edge_dst, edge_i32, node_ptr, edge_ptr, nullptr, nullptr, nullptr};
    stg::TArgs a2;
    memset(&a2, 0, sizeof(a2));
    a2.mode = mode; a2.rp_action = action; a2.rp_reward = reward; a2.rp_index = index; a2.rp_qother = q_other;
    a2.rp_gamma = gamma; a2.rp_inv_batch = 1.f / (float)batch; a2.rp_scalar = scalar;
    return stg_backward_launch(c, n_graphs, max_n, max_e, a2, grad, workspace, (cudaStream_t)stream, phase, false,
                               StgLoss{batch, next_slot, loss}, tail_sync);
}

int mdq_stream_post(uint32_t *sync, void *stream)
{
    if (!sync) { mdq::set_error("mdq_stream_post: null argument"); return MDQ_EINVAL; }
    stg::k_post<<<1, 1, 0, (cudaStream_t)stream>>>(sync);
    return mdq::check_launch("k_post");
}

}  // extern "C"
