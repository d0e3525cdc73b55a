// C ABI of the fused ConvSP group (include/spnb.h): signature matching and dispatch.  The kernels are templates in
// convsp_group.cuh, instantiated per signature in convsp_group_inst_*.cu (separate translation units so that they
// compile in parallel).
#include "convsp_group.cuh"

namespace spnb {
namespace grp {

// Distinct-data numbering of a host-side layer list -> signature + kernel arguments.
static bool make_signature(const float* locs, int D, int nl, const SpnbGroupLayer* layers, Signature& sg,
                           GroupArgs& ga, float radius, bool backward)
{
    if (nl < 1 || nl > kMaxLayers) return false;
    sg.D = D;
    sg.NL = nl;
    sg.CS = sg.SS = sg.FS = sg.NS = sg.DM = 0;
    int nsrc = 0;
    for (int l = 0; l < nl; ++l) {
        const SpnbGroupLayer& L = layers[l];
        if (L.nchannels < 1 || L.nchannels > 4 || L.nkernels < 1) return false;
        unsigned s;
        if (L.data == nullptr) {
            if (L.nchannels != 1) return false;  // implicit ones: one channel
            s = kSrcOnes;
        } else if (L.data == locs && L.nchannels == D) {
            s = kSrcLocs;
        } else {
            int found = -1;
            for (int t = 0; t < nsrc; ++t)
                if (ga.src[t] == L.data) found = t;
            if (found < 0) {
                found = nsrc;
                ga.src[nsrc++] = L.data;
            }
            s = (unsigned)found;
        }
        sg.CS |= (unsigned)L.nchannels << (4 * l);
        sg.SS |= s << (4 * l);
        sg.FS |= (unsigned)L.kernel_fn << (4 * l);
        sg.NS |= (L.dis_norm ? 1u : 0u) << l;
        // a layer's data gets a gradient whenever it is a tensor (the caller provides the buffer)
        if (s != kSrcOnes) sg.DM |= 1u << l;
        if (backward && s != kSrcOnes && !L.ddata) return false;
        const SphParams p = make_sph_params(L.kernel_fn, radius);
        LayerArgs& A = ga.l[l];
        A.data = L.data; A.weight = L.weight; A.bias = L.bias; A.out = L.out;
        A.grad_out = L.grad_out; A.ddata = L.ddata;
        A.C = L.nchannels; A.O = L.nkernels;
        A.w_expr = p.w_expr; A.dw_expr = p.dw_expr; A.dis_norm = L.dis_norm ? 1 : 0;
        double wc = p.w_coef, dwc = p.dw_coef;
        if (p.w_expr == E_DSPIKY) wc /= (double)radius;
        if (p.dw_expr == E_DSPIKY) dwc /= (double)radius;
        A.wc = (float)wc; A.dwc = (float)dwc;
    }
    // same tensor used with different channel counts -> not representable
    for (int l = 0; l < nl; ++l)
        for (int m = 0; m < l; ++m) {
            const unsigned sl = (sg.SS >> (4 * l)) & 15, sm = (sg.SS >> (4 * m)) & 15;
            if (sl == sm && sl < kSrcOnes && layers[l].nchannels != layers[m].nchannels) return false;
        }
    ga.H = radius; ga.invH = 1.0f / radius; ga.H2 = radius * radius; ga.rad2 = radius * radius;
    return true;
}

static const SigEntry* find_in(const SigEntry* tab, int n, const Signature& sg, bool wild_fn)
{
    for (int i = 0; i < n; ++i) {
        const Signature& t = tab[i].sg;
        if (t.D != sg.D || t.NL != sg.NL || t.CS != sg.CS || t.SS != sg.SS || t.DM != sg.DM) continue;
        if (wild_fn || (t.FS == sg.FS && t.NS == sg.NS)) return &tab[i];
    }
    return nullptr;
}

// Exact signatures first (kernel ids compiled in), then the single-layer family (kernel id and dis_norm at run time).
static const SigEntry* find_signature(const Signature& sg)
{
    const SigEntry* e = nullptr;
    if (sg.D == 3) e = find_in(kSigsFluid3, kNumSigsFluid3, sg, false);
    if (sg.D == 2) e = find_in(kSigsFluid2, kNumSigsFluid2, sg, false);
    if (!e && sg.NL == 1 && sg.D == 3) e = find_in(kSigsSingle3, kNumSigsSingle3, sg, true);
    if (!e && sg.NL == 1 && sg.D == 2) e = find_in(kSigsSingle2, kNumSigsSingle2, sg, true);
    return e;
}

}  // namespace grp
}  // namespace spnb

using namespace spnb;
using namespace spnb::grp;

extern "C" {

size_t spnb_convsp_group_workspace_bytes(const float* locs, int batch_size, int N, int ndims, float radius,
                                         int nlayers, const SpnbGroupLayer* layers, int backward)
{
    Signature sg;
    GroupArgs ga;
    memset(&ga, 0, sizeof(ga));
    if (!layers || !make_signature(locs, ndims, nlayers, layers, sg, ga, radius, false)) return 0;
    const SigEntry* e = find_signature(sg);
    if (!e) return 0;
    return sizeof(float) * 4 * (size_t)(backward ? e->bwd_vec : e->fwd_vec) * (size_t)batch_size * N;
}

int spnb_convsp_group_forward(const float* locs, const float* neighbors, int B, int N, int D, int K,
                              float radius, int nlayers, const SpnbGroupLayer* layers, void* workspace,
                              size_t workspace_bytes, const void* tile_lists, void* stream_)
{
    Signature sg;
    RunArgs a;
    memset(&a.ga, 0, sizeof(a.ga));
    if (!locs || !neighbors || !layers || B <= 0 || N <= 0 || K <= 0) {
        set_error("spnb_convsp_group_forward: bad arguments");
        return 0;
    }
    for (int l = 0; l < nlayers; ++l)
        if (!layers[l].weight || !layers[l].out || layers[l].kernel_fn < 0 ||
            layers[l].kernel_fn >= SPNB_NUM_KERNEL_FNS) {
            set_error("spnb_convsp_group_forward: layer %d: null pointer or bad kernel id", l);
            return 0;
        }
    const SigEntry* e = make_signature(locs, D, nlayers, layers, sg, a.ga, radius, false) ? find_signature(sg) : nullptr;
    if (!e) {
        set_error("spnb_convsp_group_forward: unsupported group signature");
        return 0;
    }
    const size_t need = sizeof(float) * 4 * (size_t)e->fwd_vec * (size_t)B * N;
    if (!workspace || workspace_bytes < need) {
        set_error("spnb_convsp_group_forward: workspace too small (%zu < %zu)", workspace_bytes, need);
        return 0;
    }
    a.locs = locs; a.neighbors = neighbors; a.B = B; a.N = N; a.K = K; a.rec = (float*)workspace;
    a.dlocs = nullptr; a.sym_flag = nullptr; a.tiles = tile_lists; a.stream = (cudaStream_t)stream_;
    const int nl = e->fwd(a);
    if (nl < 0) return 0;
    count_launches(nl);
    return check_launch("spnb_convsp_group_forward") ? 1 : 0;
}

int spnb_convsp_group_backward(const float* locs, const float* neighbors, int B, int N, int D, int K,
                               float radius, int nlayers, const SpnbGroupLayer* layers, float* dlocs,
                               const int* sym_flag, void* workspace, size_t workspace_bytes,
                               const void* tile_lists, void* stream_)
{
    Signature sg;
    RunArgs a;
    memset(&a.ga, 0, sizeof(a.ga));
    if (!locs || !neighbors || !layers || !dlocs || B <= 0 || N <= 0 || K <= 0) {
        set_error("spnb_convsp_group_backward: bad arguments");
        return 0;
    }
    for (int l = 0; l < nlayers; ++l)
        if (!layers[l].weight || !layers[l].grad_out || layers[l].kernel_fn < 0 ||
            layers[l].kernel_fn >= SPNB_NUM_KERNEL_FNS) {
            set_error("spnb_convsp_group_backward: layer %d: null pointer or bad kernel id", l);
            return 0;
        }
    const SigEntry* e = make_signature(locs, D, nlayers, layers, sg, a.ga, radius, true) ? find_signature(sg) : nullptr;
    if (!e) {
        set_error("spnb_convsp_group_backward: unsupported group signature (or a data tensor without ddata buffer)");
        return 0;
    }
    const size_t need = sizeof(float) * 4 * (size_t)e->bwd_vec * (size_t)B * N;
    if (!workspace || workspace_bytes < need) {
        set_error("spnb_convsp_group_backward: workspace too small (%zu < %zu)", workspace_bytes, need);
        return 0;
    }
    a.locs = locs; a.neighbors = neighbors; a.B = B; a.N = N; a.K = K; a.rec = (float*)workspace;
    a.dlocs = dlocs; a.sym_flag = sym_flag; a.tiles = tile_lists; a.stream = (cudaStream_t)stream_;
    // (the pack pre-pass zero-fills the scatter targets when the atomics mode is going to run)
    const int nl = e->bwd(a);
    if (nl < 0) return 0;
    count_launches(nl);
    return check_launch("spnb_convsp_group_backward") ? 1 : 0;
}

}  // extern "C"
