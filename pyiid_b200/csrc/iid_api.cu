// iid_api.cu -- C ABI (include/iid_b200.h) over the sm_100a kernels.
// Host-side bookkeeping only: element sort + padding, work-item lists, the
// closed-form normaliser, launch geometry, staging buffers.  No CPU compute
// path exists behind any entry point.
#include "../../include/iid_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "iid_debye.cuh"
#include "iid_debye2.cuh"
#include "iid_debye64.cuh"
#include "iid_force_table.cuh"
#include "iid_fq_hist.cuh"
#include "iid_fused.cuh"
#include "iid_sampler.cuh"
#include "iid_small.cuh"
#include "iid_spring.cuh"
#include "iid_ubench.cuh"

using namespace iid;

static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                              \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) {                                              \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);       \
            return (int)e_;                                                   \
        }                                                                     \
    } while (0)

#define NEED(h)                                                               \
    do {                                                                      \
        if (!(h)) return fail(IID_E_BADARG, "null handle");                   \
        if (!(h)->subs.empty())                                               \
            return fail(IID_E_BADARG, std::string(__func__) +                 \
                        ": not available on a multi-device handle");          \
        CU(cudaSetDevice((h)->device));                                       \
        if ((h)->chain_left > 0) chain_drain(h); /* a leapfrog chain in flight */ \
    } while (0)

template <typename X>
static int dev_alloc(X **p, size_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) count = 1;
    CU(cudaMalloc((void **)p, count * sizeof(X)));
    return 0;
}

// A CUDA graph of one fixed launch sequence, instantiated once the sequence
// has run twice with the same key.
struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    int key_pot = -1;
    double key_conv = 0.0;
    bool key_flag = false;
    int warm = 0;
    int64_t launches = 0;  // kernels per replay
    void drop()
    {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
        warm = 0;
    }
};

struct iid_handle {
    int device = 0;
    int precision = IID_FP32;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    int rank = 0, world = 1;
    // multi-device handle (iid_create_multi): one complete sub-handle per GPU,
    // sub d computing shard (d, active); this handle only coordinates
    std::vector<iid_handle *> subs;
    int n_devices = 1;                 // devices a multi handle may use (subs are created on demand)
    std::vector<std::pair<std::string, int64_t>> multi_options;
    double multi_rs_k[IID_MAX_RESTRAINTS] = {0};
    int active = 1;                    // sub-handles used for the current structure
    std::vector<double> inv_na_host;   // [nq] 1/na for the host-side F = 2 S / na
    std::vector<double> T_host;        // multi: the transform, for devices activated later
    // structure
    int64_t n = 0, np = 0, nq = 0, qp = 0, ntypes = 0;
    double qbin = 0.0;
    double *x = nullptr, *y = nullptr, *z = nullptr;
    float *valid = nullptr;
    int *orig = nullptr, *tile_type = nullptr;
    void *ftab = nullptr, *inv_na = nullptr;
    double *inv_na_d = nullptr;
    WorkItem *items_tri = nullptr;
    int64_t n_items_tri = 0;
    // full-gradient pass: row jobs of THIS shard (iid_debye.cuh), rebuilt when
    // the structure or the shard changes
    std::vector<int> run_begin, run_end, run_type_v;
    RowJob *jobs = nullptr;
    RowSeg *segs = nullptr;
    RowFix *fixes = nullptr;
    int64_t n_jobs = 0, n_segs = 0, n_fixes = 0, n_pieces = 0, n_jobs_total = 0;
    double *Spart = nullptr;     // [n_jobs][qp]
    size_t Spart_count = 0;
    void *Gside = nullptr;       // [n_pieces][32][3][nq]
    size_t Gside_bytes = 0;
    int piece_div = 128;         // smallest piece of a split row = a slot's share / piece_div
    bool zero_copy = true;       // gradient rows straight into pinned, mapped host arrays
    int acc_j = 4096;            // FP32 row jobs park their partial sums every acc_j j atoms
    float *Gscr = nullptr;       // [n_slots][3 * 32][block threads] parked partial sums
    size_t Gscr_count = 0;
    int *slot_busy = nullptr;
    int n_slots = 0;
    // transform
    int64_t nr = 0;
    double *T = nullptr;
    // scratch (device)
    double *pos = nullptr, *S = nullptr, *F = nullptr, *Gr = nullptr,
           *cr = nullptr, *wq = nullptr, *out4 = nullptr, *force = nullptr,
           *target = nullptr;
    void *Gfull = nullptr;
    size_t Gfull_bytes = 0;
    // tabulated radial function of the force pass (iid_force_table.cuh)
    float *phi_tab = nullptr;
    double *phi_info = nullptr;
    // Q-space shortcut of the chain-rule weights (fused host path):
    // M = T^T T, vgo = T^T target, coef from the potential kernel
    double *Mq = nullptr, *vgo = nullptr, *coef = nullptr;
    bool vgo_valid = false;
    double gogo = 0.0;            // target . target (Q-space potential of the fused kernel)
    double *MF = nullptr, *wq_blk = nullptr;  // fused kernel: (T^T T) F and per-block weights
    bool use_fused = true;        // one cooperative launch per evaluation of a small structure
    bool fused_det = true;        // ... bit-reproducible: per-item partial sums, fixed order
    bool det_fq = true;           // stand-alone F(Q) pass: per-item partial sums + fixed-order reduce
    double *Sitem_fq = nullptr;   // [this shard's items][qp]
    size_t Sitem_fq_count = 0;
    double *ft_part = nullptr;    // tabulated force pass: [jsplit][np][3] partial sums
    size_t ft_part_count = 0;
    // F(Q) pass through a radial pair histogram (iid_fq_hist.cuh): FP32 mode, large structures
    bool fq_hist = true;
    int64_t fq_hist_min_n = 1200;
    double *hist_info = nullptr;            // [8] grid step, its inverse, nodes in use, gate, stencil points
    unsigned long long *hist_C = nullptr;   // [element pairs][FH_CAP] fixed point, zero between passes
    int *hist_order = nullptr;              // [n_items_tri] item indices sorted by element pair
    double *hist_Spart = nullptr;           // [element pairs][chunks][qp] partial sums of the transform
    // the fused evaluation kernel's fixed-point accumulators (zero between launches)
    unsigned long long *Sfix = nullptr, *Ffix = nullptr;
    double *gforce = nullptr;  // [qp] per-bin bound of a pair's force per unit weight
    double *phi_tab_d = nullptr;  // its float64 radial force table
    unsigned long long *fused_C = nullptr;  // its radial pair histogram (F(Q) phase)
    bool fused_hist = true;
    double *ext_ref = nullptr;    // [4] box centre of its previous evaluation + valid flag
    bool fused_table = true;
    std::vector<double> gforce_host;
    double s_fix_scale = 1.0;
    int64_t tri_maxlen = 0;       // longest j range of a triangle item
    // zero-copy I/O of the fused kernel, set by the host entry points around
    // enqueue_eval_device (pinned staging of the handle; null = copy nodes)
    const double *zc_pos_in = nullptr;
    double *zc_force_out = nullptr, *zc_out = nullptr;
    const double *zc_ctl = nullptr;   // leapfrog: step parameters in pinned memory
    double *zc_mirror = nullptr;      // leapfrog: (q, p, scalars) of the new state in pinned memory
    bool zero_copy_small = true;
    double *zc_lf_mirror = nullptr;   // leapfrog finished inside the fused launch, mirrored here
    unsigned long long *stamps = nullptr;  // IID_FUSED_STAMPS=1: phase times of the fused kernel
    double stamp_sum[12] = {0};
    int64_t stamp_n = 0;
    // spring restraints fused into iid_energy_forces_host (iid_spring.cuh)
    int n_restraints = 0;
    int rs_type[IID_MAX_RESTRAINTS] = {0};
    double rs_k[IID_MAX_RESTRAINTS] = {0}, rs_rt[IID_MAX_RESTRAINTS] = {0};
    // scratch of the stand-alone spring calls
    double *sp_buf = nullptr;
    size_t sp_count = 0;
    bool use_force_table = true;
    int64_t force_table_min_n = 600;  // below this the direct kernel is faster
    // pinned staging for the gradient's way back to pageable host memory
    unsigned char *pinG = nullptr;
    size_t pinG_bytes = 0;
    std::vector<cudaEvent_t> chunk_ev;
    // pinned host staging
    double *pin = nullptr;
    size_t pin_count = 0;
    // tunables
    int nw_max = 12;
    bool qspace_wq = true;  // fused host path: chain-rule weights from T^T T (no R x Q pass)
    int slab_override = 0;
    // CUDA graph of the fused energy+forces sequence (small-N latency)
    GraphSlot ef;  // iid_energy_forces_host
    GraphSlot lf[IID_LF_CHAIN];  // iid_leapfrog_host / _chain_host: one per ring slot of the pinned staging
    GraphSlot lf_chain[IID_LF_CHAIN];  // [k-1]: a chain of k steps inside one fused launch
    int zc_chain = 1;
    // iid_leapfrog_chain_begin / _next: steps enqueued, steps not yet handed out
    int chain_n = 0, chain_left = 0;
    bool chain_flags = false;   // the launch raises a flag per step (in-kernel chain)
    bool chain_synced = false;  // (otherwise) the stream was synchronised for this chain
    double lf_seq = 0.0;        // flag value of the last step handed out
    int64_t chain_id = 0;       // the chain begun last
    std::vector<double> lf_mass_h;  // 1/m: kinetic energy of a step finished inside the fused launch
    bool chain_in_kernel = true;
    bool use_graph = true;
    // device-resident sampler states (iid_leapfrog_host): slot = (q, p, f)
    double *lf_slab = nullptr;   // [lf_slots][3][3n]
    int64_t lf_slots = 0;
    double *lf_mass = nullptr;   // [n]
    double *lf_ctl = nullptr;    // step, src, dst, centre flag, cell centre xyz
    double *lf_mirror = nullptr; // q [3n] | p [3n] | scalars [16] of the new state: one D2H
    double *lf_pin = nullptr;    // pinned: ctl [8] | mirror [6n] | out [16]
    size_t lf_pin_count = 0;
    bool lf_system = false;
    bool cheb = true;
    bool grad_split = true;  // full gradient: F(Q) from the lower-triangle items only
    bool prod_unroll = true; // gradient kernel: two pair set-ups per producer iteration
    int grad_nw_max = 12;    // warps per gradient block (12: one block covers the 330-bin PDF grid)
    // instrumentation
    int64_t launches = 0;
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_pending = false;
    double last_pairq = 0.0;
};

static int upload_row_jobs(iid_handle *h);
static void drop_graph(iid_handle *h);
static void chain_drain(iid_handle *h);
static int multi_destroy(iid_handle *h);
static int multi_need_subs(iid_handle *h, int count);
static int multi_set_structure(iid_handle *h, int64_t n, const int32_t *type_index,
                               int64_t n_types, const double *ftable, const double *norm_table,
                               int64_t nq, double qbin);
static int multi_set_transform(iid_handle *h, int64_t nr, int64_t nq, const double *T);
static int multi_fq_host(iid_handle *h, const double *pos_host, double *F_host, double *pdf_host);
static int multi_grad_fq_host(iid_handle *h, const double *pos_host, void *G_host,
                              double *F_host);
static int multi_energy_forces_host(iid_handle *h, const double *pos_host,
                                    const double *target_host, int potential, double conv,
                                    double *out_host, double *forces_host, double *pdf_host);
#define MULTI(h) ((h) && !(h)->subs.empty())

// ---------------------------------------------------------------------------
extern "C" int iid_version(void) { return 200; }

// Host-only: the FT_PTS Lagrange weights of the radial stencil (iid_stencil.cuh)
// for a point u in [0, 1) past node k: node k - *left + i gets w[i].  The force
// table of the fused kernel interpolates with them, the F(Q) pair histogram
// spreads with them; exposed so that the stencil can be checked without a device.
extern "C" int iid_stencil_weights(double u, double *w, int *n_points, int *left, double *qmax_h)
{
    if (!w) return fail(IID_E_BADARG, "null pointer");
    double d[FT_PTS], pre[FT_PTS];
    for (int i = 0; i < FT_PTS; ++i) d[i] = u - (double)(i - FT_LEFT);
    pre[0] = 1.0;
    for (int i = 1; i < FT_PTS; ++i) pre[i] = pre[i - 1] * d[i - 1];
    double suf = 1.0;
    for (int i = FT_PTS - 1; i >= 0; --i) {
        w[i] = ft_bary(i) * pre[i] * suf;
        suf *= d[i];
    }
    if (n_points) *n_points = FT_PTS;
    if (left) *left = FT_LEFT;
    if (qmax_h) *qmax_h = FT_QH;
    return 0;
}
// Host-only: the same for the stencils of the F(Q) pair histogram (iid_fq_hist.cuh):
// tier 0 = the finest grid (FH_PTS_FINEST points), 1 = the fine grid (FH_PTS_FINE),
// 2 = the coarse grid (FT_PTS points, the stencil above).
template <int P>
static void hist_stencil(double u, double *w)
{
    constexpr int LEFT = P / 2 - 1;
    double d[P], pre[P];
    for (int i = 0; i < P; ++i) d[i] = u - (double)(i - LEFT);
    pre[0] = 1.0;
    for (int i = 1; i < P; ++i) pre[i] = pre[i - 1] * d[i - 1];
    double suf = 1.0;
    for (int i = P - 1; i >= 0; --i) {
        w[i] = lagrange_bary<P>(i) * pre[i] * suf;
        suf *= d[i];
    }
}
extern "C" int iid_hist_stencil_weights(int tier, double u, double *w, int *n_points, int *left,
                                        double *qmax_h)
{
    if (!w || tier < 0 || tier > 2) return fail(IID_E_BADARG, "null pointer or unknown tier");
    const int pts = tier == 0 ? FH_PTS_FINEST : tier == 1 ? FH_PTS_FINE : FT_PTS;
    if (tier == 0) hist_stencil<FH_PTS_FINEST>(u, w);
    else if (tier == 1) hist_stencil<FH_PTS_FINE>(u, w);
    else hist_stencil<FT_PTS>(u, w);
    if (n_points) *n_points = pts;
    if (left) *left = pts / 2 - 1;
    if (qmax_h) *qmax_h = tier == 0 ? FH_QH_FINEST : tier == 1 ? FH_QH_FINE : FT_QH;
    return 0;
}
extern "C" const char *iid_last_error(void) { return g_err.c_str(); }

extern "C" int iid_device_count(int *count)
{
    if (!count) return fail(IID_E_BADARG, "null count");
    CU(cudaGetDeviceCount(count));
    return 0;
}

extern "C" int iid_device_info(int device, int *sm_count, int *clock_khz, int *cc)
{
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        *clock_khz = khz;
    }
    if (cc) *cc = prop.major * 10 + prop.minor;
    return 0;
}

extern "C" int iid_create(int device, int precision, iid_handle **out)
{
    if (!out) return fail(IID_E_BADARG, "null out");
    if (precision != IID_FP32 && precision != IID_FP64)
        return fail(IID_E_BADARG, "precision must be IID_FP32 or IID_FP64");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(IID_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= count) return fail(IID_E_BADARG, "bad device index");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(IID_E_NODEVICE,
                    "device is not sm_100-class; kernels are built for sm_100a only");
    iid_handle *h = new iid_handle();
    h->device = device;
    h->precision = precision;
    h->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    CU(cudaMalloc((void **)&h->out4, 8 * sizeof(double)));
    CU(cudaMemset(h->out4, 0, 8 * sizeof(double)));
    if (const char *s = getenv("IID_NW")) h->nw_max = std::max(1, std::min(12, atoi(s)));
    h->nw_max = std::min(h->nw_max, 12);
    if (const char *s = getenv("IID_GRAPH")) h->use_graph = atoi(s) != 0;
    if (const char *s = getenv("IID_FORCE_TABLE")) h->use_force_table = atoi(s) != 0;
    if (const char *s = getenv("IID_FORCE_TABLE_MIN_N")) h->force_table_min_n = atoll(s);
    if (const char *s = getenv("IID_CHEB")) h->cheb = atoi(s) != 0;
    if (const char *s = getenv("IID_GRAD_SPLIT")) h->grad_split = atoi(s) != 0;
    if (const char *s = getenv("IID_PROD_UNROLL")) h->prod_unroll = atoi(s) != 0;
    if (const char *s = getenv("IID_GRAD_NW")) h->grad_nw_max = std::max(1, std::min(12, atoi(s)));
    if (const char *s = getenv("IID_QSPACE_WQ")) h->qspace_wq = atoi(s) != 0;
    if (const char *s = getenv("IID_SLAB")) h->slab_override = std::max(0, atoi(s));
    if (const char *s = getenv("IID_PIECE_DIV")) h->piece_div = std::max(1, atoi(s));
    if (const char *s = getenv("IID_ZERO_COPY")) h->zero_copy = atoi(s) != 0;
    if (const char *s = getenv("IID_ACC_J")) h->acc_j = std::max(0, atoi(s));
    if (const char *s = getenv("IID_FUSED")) h->use_fused = atoi(s) != 0;
    if (const char *s = getenv("IID_FUSED_DET")) h->fused_det = atoi(s) != 0;
    if (const char *s = getenv("IID_CHAIN_IN_KERNEL")) h->chain_in_kernel = atoi(s) != 0;
    if (const char *s = getenv("IID_FUSED_TABLE")) h->fused_table = atoi(s) != 0;
    if (const char *s = getenv("IID_FUSED_HIST")) h->fused_hist = atoi(s) != 0;
    if (const char *s = getenv("IID_FQ_HIST")) h->fq_hist = atoi(s) != 0;
    if (const char *s = getenv("IID_FQ_HIST_MIN_N")) h->fq_hist_min_n = std::max(2, atoi(s));
    if (const char *s = getenv("IID_DET_FQ")) h->det_fq = atoi(s) != 0;
    if (const char *s = getenv("IID_ZERO_COPY_SMALL")) h->zero_copy_small = atoi(s) != 0;
    *out = h;
    return 0;
}

extern "C" int iid_destroy(iid_handle *h)
{
    if (MULTI(h)) return multi_destroy(h);
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    void *ptrs[] = {h->x, h->y, h->z, h->valid, h->orig, h->tile_type, h->ftab,
                    h->inv_na, h->inv_na_d, h->items_tri, h->jobs, h->segs, h->fixes, h->Spart,
                    h->Gside, h->Gscr, h->slot_busy, h->MF, h->wq_blk, h->Sfix, h->Ffix, h->gforce, h->phi_tab_d, h->fused_C, h->ext_ref, h->hist_info, h->hist_C, h->hist_order, h->hist_Spart, h->Sitem_fq, h->ft_part, h->T,
                    h->pos, h->S, h->F, h->Gr, h->cr, h->wq, h->out4, h->force,
                    h->target, h->Gfull, h->phi_tab, h->phi_info, h->Mq, h->vgo, h->coef, h->sp_buf,
                    h->lf_slab, h->lf_mass, h->lf_ctl, h->lf_mirror};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->pinG) cudaFreeHost(h->pinG);
    if (h->stamps) cudaFree(h->stamps);
    for (cudaEvent_t e : h->chunk_ev) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    drop_graph(h);
    if (h->lf_pin) cudaFreeHost(h->lf_pin);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" int iid_get_stream(iid_handle *h, void **stream)
{
    if (MULTI(h)) return iid_get_stream(h->subs[0], stream);
    if (!h || !stream) return fail(IID_E_BADARG, "null argument");
    *stream = (void *)h->stream;
    return 0;
}

extern "C" int iid_synchronize(iid_handle *h)
{
    if (MULTI(h)) {
        for (iid_handle *sh : h->subs) {
            int rc = iid_synchronize(sh);
            if (rc) return rc;
        }
        return 0;
    }
    NEED(h);
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

static void drop_graph(iid_handle *h)
{
    h->ef.drop();
    for (GraphSlot &g : h->lf) g.drop();
    for (GraphSlot &g : h->lf_chain) g.drop();
}

extern "C" int iid_set_shard(iid_handle *h, int rank, int world)
{
    if (MULTI(h)) return fail(IID_E_BADARG, "a multi-device handle shards over its own devices");
    if (!h) return fail(IID_E_BADARG, "null handle");
    if (world < 1 || rank < 0 || rank >= world)
        return fail(IID_E_BADARG, "need 0 <= rank < world");
    const bool changed = rank != h->rank || world != h->world;
    if (changed) drop_graph(h);
    h->rank = rank;
    h->world = world;
    if (changed && h->np > 0) {
        CU(cudaSetDevice(h->device));
        CU(cudaStreamSynchronize(h->stream));
        return upload_row_jobs(h);
    }
    return 0;
}

// Work items: (i-tile, j-slab) with one element type per slab.  Triangle
// lists (F(Q), force) take the j tiles strictly below the i-tile plus one
// diagonal item; square lists (full gradient) add the j tiles above it as
// gradient-only items (ITEM_NOF): F(Q) is summed once per pair, below the
// diagonal, and the kernels run a shorter bin loop above it.
static void build_items(const std::vector<int> &run_begin,
                        const std::vector<int> &run_end, int np, int slab,
                        bool triangle, std::vector<WorkItem> &out, int gran = TILE_I)
{
    const int ntile = np / TILE_I;
    out.clear();
    // equal slabs of at most `slab` atoms, multiples of `gran` (32, or the
    // 8-atom j tile of the F(Q) / force kernels for small structures)
    auto slabs = [&](int it, int rb, int re, int info) {
        if (re <= rb) return;
        const int len = re - rb;
        const int nsl = (len + slab - 1) / slab;
        const int per = ((len + nsl - 1) / nsl + gran - 1) / gran * gran;
        for (int j0 = rb; j0 < re; j0 += per)
            out.push_back({it, j0, std::min(re, j0 + per), info});
    };
    for (int it = 0; it < ntile; ++it) {
        const int lo = it * TILE_I, hi = lo + TILE_I;
        for (size_t b = 0; b < run_begin.size(); ++b) {
            // below the i-tile: every pair once, F counts 1
            slabs(it, run_begin[b], std::min(run_end[b], lo), (int)b);
            // above it (square lists): the same pairs again from the other
            // side; they only add to the gradient rows of this i-tile
            if (!triangle) slabs(it, std::max(run_begin[b], hi), run_end[b], (int)b | ITEM_NOF);
        }
        // the i-tile against itself: both orders present, F counts 1/2
        int b = 0;
        for (size_t k = 0; k < run_begin.size(); ++k)
            if (lo >= run_begin[k] && lo < run_end[k]) b = (int)k;
        out.push_back({it, lo, hi, b | ITEM_DIAG});
    }
    // longest first: the hardware block scheduler then fills the tail with
    // short items
    std::stable_sort(out.begin(), out.end(), [](const WorkItem &a, const WorkItem &b) {
        return (a.jend - a.jbegin) > (b.jend - b.jbegin);
    });
}

// Small structures (a few waves of blocks at most): the triangle list of
// 32-atom slabs leaves the machine half idle in its last wave (Au561: 171
// items on 148 SMs).  Choose the slab cap, in units of the 8-atom j tile, that
// minimises the makespan of a longest-first schedule on `slots` SMs, counting
// a fixed per-item cost (prologue + flush, about one and a half j tiles).
static int pick_small_slab(const std::vector<int> &run_begin, const std::vector<int> &run_end,
                           int np, int slots)
{
    constexpr int GRAN = 8, FIXED = 12;
    const int ntile = np / TILE_I;
    int best = TILE_I, best1 = TILE_I;
    long best_cost = -1, best1_cost = -1;
    std::vector<int> lens;
    std::vector<long> load;
    for (int cap = GRAN; cap <= 512; cap += GRAN) {
        lens.clear();
        for (int it = 0; it < ntile; ++it) {
            const int jlimit = it * TILE_I;
            for (size_t b = 0; b < run_begin.size(); ++b) {
                const int rb = run_begin[b], re = std::min(run_end[b], jlimit);
                if (re <= rb) continue;
                const int len = re - rb, nsl = (len + cap - 1) / cap;
                const int per = ((len + nsl - 1) / nsl + GRAN - 1) / GRAN * GRAN;
                for (int j0 = 0; j0 < len; j0 += per) lens.push_back(std::min(per, len - j0));
            }
            lens.push_back(TILE_I);  // diagonal item
        }
        std::sort(lens.begin(), lens.end(), std::greater<int>());
        load.assign((size_t)slots, 0);
        std::make_heap(load.begin(), load.end(), std::greater<long>());
        for (int l : lens) {  // the next block goes to the SM that frees up first
            std::pop_heap(load.begin(), load.end(), std::greater<long>());
            load.back() += l + FIXED;
            std::push_heap(load.begin(), load.end(), std::greater<long>());
        }
        const long cost = *std::max_element(load.begin(), load.end());
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = cap;
        }
        // one item per SM lets the whole evaluation run as ONE cooperative launch
        // (iid_fused.cuh), which is worth more than a tighter makespan of the
        // pair passes alone: remember the best single-wave list as well
        if ((int)lens.size() <= slots && (best1_cost < 0 || cost < best1_cost)) {
            best1_cost = cost;
            best1 = cap;
        }
    }
    // (cost is in j atoms per SM, ~0.25 us each over the two pair passes; the
    // single launch saves ~35 us of launch gaps and one-block stages)
    if (best1_cost > 0 && best1_cost - best_cost < 140) return best1;
    return best;
}

// ---------------------------------------------------------------------------
// Row jobs of the full-gradient pass (iid_debye.cuh).  Rank `rank` of `world`
// owns the i-tiles t with t % world == rank: every gradient row is computed by
// exactly one rank, so there is no gradient collective and every rank writes
// (or downloads) only its own rows.  A row = the i-tile against ALL j: per
// element run the range below the tile (F(Q) + gradient), the tile itself
// (F(Q) weight 1/2) and the range above it (gradient only).  Most rows are one
// job each (plain stores into G); the rows kept back for load balance are cut
// into pieces of decreasing size (guided self-scheduling: piece = remaining
// tail work / slots, not below row / piece_div) that are dispatched after the
// whole rows and meet in rows_fixup_kernel.
struct RowPlan {
    std::vector<RowJob> jobs;
    std::vector<RowSeg> segs;
    std::vector<RowFix> fixes;
    int64_t n_pieces = 0, pair_slots = 0, n_jobs_total = 0;
};

static double seg_cost(const RowSeg &g)
{
    // packed instructions per bin step: 8.25 with F(Q), 6.7 gradient only; one
    // 16-j tile of pipeline restart per segment
    return (double)(g.jend - g.jbegin + 16) * ((g.info & ITEM_NOF) ? 6.7 : 8.25);
}

static void row_segments(const std::vector<int> &run_begin, const std::vector<int> &run_end,
                         const std::vector<int> &run_type, int it, std::vector<RowSeg> &out)
{
    const int lo = it * TILE_I, hi = lo + TILE_I;
    out.clear();
    for (size_t b = 0; b < run_begin.size(); ++b) {
        const int rb = run_begin[b], re = run_end[b], ty = run_type[b];
        const size_t first = out.size();
        if (re <= lo) out.push_back({rb, re, ty, 0});
        else if (rb >= hi) out.push_back({rb, re, ty | ITEM_NOF, 0});
        else {
            if (lo > rb) out.push_back({rb, lo, ty, 0});
            out.push_back({lo, hi, ty | ITEM_DIAG, 0});
            if (re > hi) out.push_back({hi, re, ty | ITEM_NOF, 0});
        }
        if (out.size() > first) out.back().info |= SEG_FLUSH;
    }
}

static void build_row_jobs(const std::vector<int> &run_begin, const std::vector<int> &run_end,
                           const std::vector<int> &run_type, int np, int rank, int world,
                           int slots, int piece_div, RowPlan &P)
{
    P = RowPlan();
    const int ntile = np / TILE_I;
    slots = std::max(1, slots);
    piece_div = std::max(1, piece_div);
    struct Row { int it; double cost; std::vector<RowSeg> segs; };
    std::vector<Row> rows;
    for (int it = rank; it < ntile; it += world) {
        Row r;
        r.it = it;
        row_segments(run_begin, run_end, run_type, it, r.segs);
        r.cost = 0.0;
        for (const RowSeg &g : r.segs) r.cost += seg_cost(g);
        rows.push_back(std::move(r));
    }
    std::stable_sort(rows.begin(), rows.end(),
                     [](const Row &a, const Row &b) { return a.cost > b.cost; });
    const size_t nrows = rows.size();
    // whole rows: all but the last 0.5 .. 1.5 rows per slot
    size_t nwhole = 0;
    if (nrows >= (size_t)slots) {
        const double waves = (double)nrows / slots;
        nwhole = (size_t)std::max(0.0, std::floor(waves - 0.5)) * slots;
    }
    auto emit = [&](int it, const std::vector<RowSeg> &sg, int dest) {
        RowJob j;
        j.itile = it;
        j.seg_begin = (int)P.segs.size();
        for (const RowSeg &g : sg) {
            P.segs.push_back(g);
            P.pair_slots += (int64_t)(g.jend - g.jbegin) * TILE_I;
        }
        j.seg_end = (int)P.segs.size();
        j.dest = dest;
        P.jobs.push_back(j);
    };
    for (size_t k = 0; k < nwhole; ++k) emit(rows[k].it, rows[k].segs, -1);
    double remaining = 0.0, total = 0.0;
    for (size_t k = 0; k < nrows; ++k) total += rows[k].cost;
    for (size_t k = nwhole; k < nrows; ++k) remaining += rows[k].cost;
    // smallest piece: the schedule ends within about one piece of the ideal, so
    // 1/piece_div of a slot's share (not below 64 j)
    const double min_piece = std::max(64 * 8.25, total / slots / piece_div);
    for (size_t k = nwhole; k < nrows; ++k) {
        const Row &r = rows[k];
        // cut the row's segment list into pieces
        std::vector<std::vector<RowSeg>> pieces;
        std::vector<RowSeg> cur;
        double cur_cost = 0.0, row_left = r.cost;  // row_left: not yet in a closed piece
        double target = std::max(min_piece, remaining / slots);
        auto close_piece = [&]() {
            pieces.push_back(cur);
            row_left -= cur_cost;
            remaining = std::max(0.0, remaining - cur_cost);
            cur.clear();
            cur_cost = 0.0;
            target = std::max(min_piece, remaining / slots);
        };
        for (size_t si = 0; si < r.segs.size(); ++si) {
            RowSeg g = r.segs[si];
            for (;;) {
                const double c = seg_cost(g);
                if (row_left <= 1.25 * target || cur_cost + c <= 1.1 * target) {
                    cur.push_back(g);  // the rest of the row is the last piece / the segment fits
                    cur_cost += c;
                    break;
                }
                if (cur_cost >= 0.9 * target) {
                    close_piece();
                    continue;
                }
                // cut inside the segment at a multiple of 32 j
                const double per_j = (g.info & ITEM_NOF) ? 6.7 : 8.25;
                int take = (int)((target - cur_cost) / per_j) / TILE_I * TILE_I;
                take = std::max(TILE_I, take);
                if (take >= g.jend - g.jbegin) {
                    cur.push_back(g);
                    cur_cost += c;
                    break;
                }
                RowSeg head = g;
                head.jend = g.jbegin + take;
                head.info &= ~SEG_FLUSH;
                cur.push_back(head);
                cur_cost += seg_cost(head);
                g.jbegin = head.jend;
                close_piece();
            }
        }
        if (!cur.empty()) close_piece();
        if (pieces.size() == 1) {
            emit(r.it, pieces[0], -1);
        } else {
            RowFix f;
            f.itile = r.it;
            f.d0 = (int)P.n_pieces;
            for (auto &pc : pieces) emit(r.it, pc, (int)P.n_pieces++);
            f.d1 = (int)P.n_pieces;
            f.pad = 0;
            P.fixes.push_back(f);
        }
    }
    // longest first: the block scheduler hands out the jobs in this order
    auto job_cost = [&](const RowJob &j) {
        double c = 0.0;
        for (int k = j.seg_begin; k < j.seg_end; ++k) c += seg_cost(P.segs[k]);
        return c;
    };
    std::vector<std::pair<double, RowJob>> order;
    for (const RowJob &j : P.jobs) order.push_back({job_cost(j), j});
    std::stable_sort(order.begin(), order.end(),
                     [](const std::pair<double, RowJob> &a, const std::pair<double, RowJob> &b) {
                         return a.first > b.first;
                     });
    for (size_t k = 0; k < order.size(); ++k) P.jobs[k] = order[k].second;
    P.n_jobs_total = (int64_t)P.jobs.size();
}

// Host-side plan of one structure: element-sorted, per-element padded atom
// order and the two work-item lists.  No device needed.
struct Layout {
    int64_t np = 0;
    std::vector<int64_t> count;
    std::vector<int> run_type, run_begin, run_end, orig, tile_type;
    std::vector<WorkItem> tri;
};

static int build_layout(int64_t n, const int32_t *type_index, int64_t n_types, int sm_count,
                        int slab_override, Layout &L)
{
    if (n < 1 || n_types < 1 || !type_index) return fail(IID_E_BADARG, "bad structure arguments");
    if (n_types > 65535) return fail(IID_E_BADARG, "more than 65535 element types");
    if (n > (int64_t)1 << 30) return fail(IID_E_BADARG, "too many atoms");
    L.count.assign(n_types, 0);
    for (int64_t i = 0; i < n; ++i) {
        if (type_index[i] < 0 || type_index[i] >= n_types)
            return fail(IID_E_BADARG, "type_index out of range");
        ++L.count[type_index[i]];
    }
    std::vector<int> &run_begin = L.run_begin, &run_end = L.run_end;
    run_begin.clear();
    run_end.clear();
    int64_t np = 0;
    std::vector<int64_t> start(n_types, 0);
    L.run_type.clear();
    for (int64_t e = 0; e < n_types; ++e) {
        start[e] = np;
        if (L.count[e] == 0) continue;
        const int64_t padded = (L.count[e] + TILE_I - 1) / TILE_I * TILE_I;
        run_begin.push_back((int)np);
        run_end.push_back((int)(np + padded));
        L.run_type.push_back((int)e);
        np += padded;
    }
    L.np = np;
    L.orig.assign(np, -1);
    L.tile_type.assign(np / TILE_I, 0);
    {
        std::vector<int64_t> fill(start);
        for (int64_t i = 0; i < n; ++i) L.orig[fill[type_index[i]]++] = (int)i;
        for (size_t b = 0; b < run_begin.size(); ++b)
            for (int t = run_begin[b] / TILE_I; t < run_end[b] / TILE_I; ++t)
                L.tile_type[t] = L.run_type[b];
    }
    // slab length: enough items to fill the machine ~12x over
    const int64_t ntile = np / TILE_I;
    const int64_t want = (int64_t)std::max(1, sm_count) * 12;
    auto pick = [&](int64_t total_j) {
        int64_t s = total_j / std::max<int64_t>(1, want);
        s = (s + TILE_I - 1) / TILE_I * TILE_I;
        s = std::max<int64_t>(TILE_I, std::min<int64_t>(4096, s));
        if (slab_override > 0) s = (slab_override + TILE_I - 1) / TILE_I * TILE_I;
        return (int)s;
    };
    if (ntile <= 64 && slab_override <= 0) {
        const int cap = pick_small_slab(run_begin, run_end, (int)np, std::max(1, sm_count));
        build_items(run_begin, run_end, (int)np, cap, true, L.tri, 8);
    } else {
        build_items(run_begin, run_end, (int)np, pick(ntile * np / 2), true, L.tri);
    }
    // the item info field stores the run's ELEMENT type
    for (auto &w : L.tri) {
        const int b = w.info & 0xffff;
        w.info = (w.info & (ITEM_DIAG | ITEM_NOF)) | L.run_type[b];
    }
    return 0;
}

// Host-only view of the sharding (no device): how many work items rank `rank`
// of `world` takes and how many (i, j) slots they cover.
extern "C" int iid_plan_shard(int64_t n, const int32_t *type_index, int64_t n_types,
                              int sm_count, int triangle, int rank, int world,
                              int64_t *n_items_total, int64_t *n_items_mine,
                              int64_t *pair_slots_mine, int64_t *padded_atoms)
{
    if (world < 1 || rank < 0 || rank >= world) return fail(IID_E_BADARG, "need 0 <= rank < world");
    Layout L;
    int rc = build_layout(n, type_index, n_types, sm_count, 0, L);
    if (rc) return rc;
    int64_t mine = 0, slots = 0, total = 0;
    if (triangle) {
        const std::vector<WorkItem> &items = L.tri;
        for (size_t k = (size_t)rank; k < items.size(); k += (size_t)world) {
            ++mine;
            slots += (int64_t)(items[k].jend - items[k].jbegin) * TILE_I;
        }
        total = (int64_t)items.size();
    } else {
        // full gradient: row jobs, i-tiles dealt cyclically to the ranks
        RowPlan P;
        for (int r = 0; r < world; ++r) {
            build_row_jobs(L.run_begin, L.run_end, L.run_type, (int)L.np, r, world, sm_count, 128, P);
            total += (int64_t)P.jobs.size();
            if (r == rank) {
                mine = (int64_t)P.jobs.size();
                slots = P.pair_slots;
            }
        }
    }
    if (n_items_total) *n_items_total = total;
    if (n_items_mine) *n_items_mine = mine;
    if (pair_slots_mine) *pair_slots_mine = slots;
    if (padded_atoms) *padded_atoms = L.np;
    return 0;
}

// Host-only: the row jobs rank `rank` of `world` would run (4 int32 per job:
// itile, seg_begin, seg_end, dest; per segment: jbegin, jend, info, 0; per
// split row: itile, first slot, one past the last slot, 0).  Arrays may be
// NULL to query the counts only.
extern "C" int iid_plan_rows(int64_t n, const int32_t *type_index, int64_t n_types, int sm_count,
                             int rank, int world, int piece_div, int64_t *n_jobs,
                             int64_t *n_segs, int64_t *n_fixes, int32_t *jobs, int64_t jobs_cap,
                             int32_t *segs, int64_t segs_cap, int32_t *fixes, int64_t fixes_cap)
{
    if (world < 1 || rank < 0 || rank >= world) return fail(IID_E_BADARG, "need 0 <= rank < world");
    Layout L;
    int rc = build_layout(n, type_index, n_types, sm_count, 0, L);
    if (rc) return rc;
    RowPlan P;
    build_row_jobs(L.run_begin, L.run_end, L.run_type, (int)L.np, rank, world, sm_count,
                   piece_div > 0 ? piece_div : 128, P);
    if (n_jobs) *n_jobs = (int64_t)P.jobs.size();
    if (n_segs) *n_segs = (int64_t)P.segs.size();
    if (n_fixes) *n_fixes = (int64_t)P.fixes.size();
    static_assert(sizeof(RowJob) == 16 && sizeof(RowSeg) == 16 && sizeof(RowFix) == 16, "layout");
    if (jobs) memcpy(jobs, P.jobs.data(), std::min<size_t>(jobs_cap, P.jobs.size()) * 16);
    if (segs) memcpy(segs, P.segs.data(), std::min<size_t>(segs_cap, P.segs.size()) * 16);
    if (fixes) memcpy(fixes, P.fixes.data(), std::min<size_t>(fixes_cap, P.fixes.size()) * 16);
    return 0;
}

// (Re)build this shard's row jobs of the full-gradient pass and upload them.
static int upload_row_jobs(iid_handle *h)
{
    if (h->np == 0) return 0;
    RowPlan P;
    build_row_jobs(h->run_begin, h->run_end, h->run_type_v, (int)h->np, h->rank, h->world,
                   h->sm_count, h->piece_div, P);
    int rc;
    if ((rc = dev_alloc(&h->jobs, P.jobs.size())) || (rc = dev_alloc(&h->segs, P.segs.size())) ||
        (rc = dev_alloc(&h->fixes, P.fixes.size())))
        return rc;
    CU(cudaMemcpy(h->jobs, P.jobs.data(), P.jobs.size() * sizeof(RowJob), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->segs, P.segs.data(), P.segs.size() * sizeof(RowSeg), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->fixes, P.fixes.data(), P.fixes.size() * sizeof(RowFix),
                  cudaMemcpyHostToDevice));
    h->n_jobs = (int64_t)P.jobs.size();
    h->n_segs = (int64_t)P.segs.size();
    h->n_fixes = (int64_t)P.fixes.size();
    h->n_pieces = P.n_pieces;
    CU(cudaStreamSynchronize(0));
    return 0;
}

extern "C" int iid_set_structure(iid_handle *h, int64_t n, const int32_t *type_index,
                                 int64_t n_types, const double *ftable, int64_t nq,
                                 double qbin)
{
    return iid_set_structure_norm(h, n, type_index, n_types, ftable, nullptr, nq, qbin);
}

// The pair sums take ftable, the normaliser na takes norm_table (null: ftable).  The
// two differ when a per-atom factor multiplies the pair term but not <f>^2: the
// Debye-Waller factor tau = t_i t_j of isotropic displacements in
// fq = norm * omega * tau (kernels/cpu_nxn.py:114-121), with norm = f_i f_j alone
// in na (flat_multi_cpu_wrap.py:52-55).
extern "C" int iid_set_structure_norm(iid_handle *h, int64_t n, const int32_t *type_index,
                                      int64_t n_types, const double *ftable,
                                      const double *norm_table, int64_t nq, double qbin)
{
    if (MULTI(h))
        return multi_set_structure(h, n, type_index, n_types, ftable, norm_table, nq, qbin);
    NEED(h);
    if (n < 1 || n_types < 1 || nq < 1 || !type_index || !ftable)
        return fail(IID_E_BADARG, "bad structure arguments");
    CU(cudaStreamSynchronize(h->stream));
    drop_graph(h);
    h->lf_system = false;  // state slots are sized for the previous structure
    Layout L;
    {
        int rc0 = build_layout(n, type_index, n_types, h->sm_count, h->slab_override, L);
        if (rc0) return rc0;
    }
    const int64_t np = L.np;
    const std::vector<int64_t> &count = L.count;
    const std::vector<int> &orig = L.orig, &tile_type = L.tile_type;
    const std::vector<WorkItem> &tri = L.tri;
    const int64_t qp = (nq + 31) / 32 * 32;
    // normaliser, closed form of N * mean_pairs(f_i f_j):
    //   na = ((sum_i f_i)^2 - sum_i f_i^2) / (N - 1)
    std::vector<double> inv_na(qp, 0.0);
    for (int64_t m = 0; m < nq; ++m) {
        double s1 = 0.0, s2 = 0.0;
        for (int64_t e = 0; e < n_types; ++e) {
            const double f = (norm_table ? norm_table : ftable)[e * nq + m];
            s1 += (double)count[e] * f;
            s2 += (double)count[e] * f * f;
        }
        const double na = n > 1 ? (s1 * s1 - s2) / (double)(n - 1) : 0.0;
        inv_na[m] = (na != 0.0 && std::isfinite(na)) ? 1.0 / na : 0.0;
    }
    if (nq != h->nq) h->nr = 0;  // a transform built for another Q grid is void
    h->n = n; h->np = np; h->nq = nq; h->qp = qp; h->ntypes = n_types; h->qbin = qbin;
    h->inv_na_host.assign(inv_na.begin(), inv_na.begin() + nq);
    int rc;
    if ((rc = dev_alloc(&h->x, np)) || (rc = dev_alloc(&h->y, np)) ||
        (rc = dev_alloc(&h->z, np)) || (rc = dev_alloc(&h->valid, np)) ||
        (rc = dev_alloc(&h->orig, np)) || (rc = dev_alloc(&h->tile_type, np / TILE_I)) ||
        (rc = dev_alloc(&h->inv_na_d, qp)) || (rc = dev_alloc(&h->items_tri, tri.size())) ||
        (rc = dev_alloc(&h->pos, 3 * n)) ||
        (rc = dev_alloc(&h->S, qp)) || (rc = dev_alloc(&h->F, qp)) ||
        (rc = dev_alloc(&h->wq, qp)) || (rc = dev_alloc(&h->force, 3 * n)))
        return rc;
    h->n_items_tri = (int64_t)tri.size();
    {
        // the histogram F(Q) pass walks the items grouped by element pair (a block
        // keeps ONE pair's histogram in shared memory), longest first within a pair
        std::vector<int> order(tri.size());
        for (size_t k = 0; k < tri.size(); ++k) order[k] = (int)k;
        auto pair_of = [&](int k) {
            const int ta = tile_type[tri[k].itile], tb = tri[k].info & 0xffff;
            return ta >= tb ? ta * (ta + 1) / 2 + tb : tb * (tb + 1) / 2 + ta;
        };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            const int pa = pair_of(a), pb = pair_of(b);
            if (pa != pb) return pa < pb;
            return (tri[a].jend - tri[a].jbegin) > (tri[b].jend - tri[b].jbegin);
        });
        int rc2 = dev_alloc(&h->hist_order, std::max<size_t>(1, order.size()));
        if (rc2) return rc2;
        if (!order.empty())
            CU(cudaMemcpy(h->hist_order, order.data(), order.size() * sizeof(int),
                          cudaMemcpyHostToDevice));
        for (double **b : {&h->hist_info, &h->hist_Spart})
            if (*b) { cudaFree(*b); *b = nullptr; }
        if (h->hist_C) { cudaFree(h->hist_C); h->hist_C = nullptr; }
    }
    h->tri_maxlen = 0;
    for (const WorkItem &w : tri) h->tri_maxlen = std::max<int64_t>(h->tri_maxlen, w.jend - w.jbegin);
    // the fused evaluation kernel's buffers are sized per structure
    for (double **b : {&h->MF, &h->wq_blk, &h->gforce, &h->phi_tab_d, &h->ext_ref})
        if (*b) { cudaFree(*b); *b = nullptr; }
    for (unsigned long long **b : {&h->Sfix, &h->Ffix, &h->fused_C})
        if (*b) { cudaFree(*b); *b = nullptr; }
    {
        // fixed-point scale of its F(Q) accumulators: |S[m]| <= pairs * f^2 * Q
        // (|sin(Qr)/r| <= Q), with room for float32 rounding; and the per-bin
        // bound of a pair's force per unit weight (force_fix_scale, iid_fused.cuh)
        h->gforce_host.assign(qp, 0.0);
        double bound = 0.0;
        for (int64_t m = 0; m < nq; ++m) {
            double f2 = 0.0;
            for (int64_t e = 0; e < n_types; ++e)
                f2 = std::max(f2, ftable[e * nq + m] * ftable[e * nq + m]);
            const double Q = (double)m * qbin;
            bound = std::max(bound, f2 * std::max(Q, 1e-3));
            h->gforce_host[m] = 0.6 * Q * Q * f2 * std::fabs(inv_na[m]);
        }
        bound *= 4.0 * (double)np * (double)np;
        int e2 = 0;
        if (bound > 0.0 && std::isfinite(bound)) std::frexp(bound, &e2);
        h->s_fix_scale = std::ldexp(1.0, 60 - e2);
    }
    h->run_begin = L.run_begin;
    h->run_end = L.run_end;
    h->run_type_v = L.run_type;
    CU(cudaMemcpy(h->orig, orig.data(), np * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->tile_type, tile_type.data(), tile_type.size() * sizeof(int),
                  cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->inv_na_d, inv_na.data(), qp * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->items_tri, tri.data(), tri.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
    if ((rc = upload_row_jobs(h))) return rc;
    CU(cudaMemset(h->S, 0, qp * sizeof(double)));
    CU(cudaMemset(h->F, 0, qp * sizeof(double)));
    CU(cudaMemset(h->wq, 0, qp * sizeof(double)));
    const size_t esz = h->precision == IID_FP32 ? sizeof(float) : sizeof(double);
    if (h->ftab) { cudaFree(h->ftab); h->ftab = nullptr; }
    if (h->inv_na) { cudaFree(h->inv_na); h->inv_na = nullptr; }
    CU(cudaMalloc(&h->ftab, (size_t)n_types * qp * esz));
    CU(cudaMalloc(&h->inv_na, (size_t)qp * esz));
    if (h->precision == IID_FP32) {
        std::vector<float> ft((size_t)n_types * qp, 0.f), in(qp, 0.f);
        for (int64_t e = 0; e < n_types; ++e)
            for (int64_t m = 0; m < nq; ++m) ft[e * qp + m] = (float)ftable[e * nq + m];
        for (int64_t m = 0; m < qp; ++m) in[m] = (float)inv_na[m];
        CU(cudaMemcpy(h->ftab, ft.data(), ft.size() * esz, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->inv_na, in.data(), in.size() * esz, cudaMemcpyHostToDevice));
    } else {
        std::vector<double> ft((size_t)n_types * qp, 0.0);
        for (int64_t e = 0; e < n_types; ++e)
            for (int64_t m = 0; m < nq; ++m) ft[e * qp + m] = ftable[e * nq + m];
        CU(cudaMemcpy(h->ftab, ft.data(), ft.size() * esz, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->inv_na, inv_na.data(), qp * esz, cudaMemcpyHostToDevice));
    }
    CU(cudaStreamSynchronize(0));  // NULL-stream copies done before the handle's stream runs
    // pinned staging: positions / forces (3n), F (qp), G(r) (nr, grown later), 4
    const size_t need = (size_t)6 * n + 2 * qp + 8 + 2 * (size_t)h->nr;
    if (need > h->pin_count) {
        if (h->pin) cudaFreeHost(h->pin);
        h->pin = nullptr;
        CU(cudaMallocHost((void **)&h->pin, need * sizeof(double)));
        h->pin_count = need;
    }
    return 0;
}

extern "C" int iid_set_transform(iid_handle *h, int64_t nr, int64_t nq, const double *T)
{
    if (MULTI(h)) return multi_set_transform(h, nr, nq, T);
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (nq != h->nq) return fail(IID_E_BADARG, "transform nq differs from structure nq");
    if (nr < 1 || !T) return fail(IID_E_BADARG, "bad transform arguments");
    CU(cudaStreamSynchronize(h->stream));
    drop_graph(h);
    int rc;
    if ((rc = dev_alloc(&h->T, (size_t)nr * h->qp)) || (rc = dev_alloc(&h->Gr, nr)) ||
        (rc = dev_alloc(&h->cr, nr)) || (rc = dev_alloc(&h->target, nr)))
        return rc;
    CU(cudaMemset(h->T, 0, (size_t)nr * h->qp * sizeof(double)));
    CU(cudaMemcpy2D(h->T, h->qp * sizeof(double), T, nq * sizeof(double),
                    nq * sizeof(double), nr, cudaMemcpyHostToDevice));
    h->nr = nr;
    if ((rc = dev_alloc(&h->Mq, (size_t)h->qp * h->qp)) || (rc = dev_alloc(&h->vgo, h->qp)) ||
        (rc = dev_alloc(&h->coef, 2)))
        return rc;
    // the copies above ran on the NULL stream; the handle's stream is
    // non-blocking and does not order itself after them
    CU(cudaStreamSynchronize(0));
    {
        const unsigned g = (unsigned)((h->qp + 15) / 16);
        ttt_kernel<<<dim3(g, g), 256, 0, h->stream>>>(h->T, (int)nr, (int)h->nq, (int)h->qp, h->Mq);
        ++h->launches;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    }
    h->vgo_valid = false;
    const size_t need = (size_t)6 * h->n + 2 * h->qp + 8 + 2 * (size_t)nr;
    if (need > h->pin_count) {
        if (h->pin) cudaFreeHost(h->pin);
        h->pin = nullptr;
        CU(cudaMallocHost((void **)&h->pin, need * sizeof(double)));
        h->pin_count = need;
    }
    return 0;
}

extern "C" int iid_get_sizes(iid_handle *h, int64_t *n, int64_t *nq, int64_t *nr,
                             int64_t *n_items_fq, int64_t *n_items_grad)
{
    if (MULTI(h)) {
        int64_t fq = 0, gr = 0, a = 0, b = 0;
        for (int d = 0; d < h->active; ++d) {
            iid_get_sizes(h->subs[d], n, nq, nr, &a, &b);
            fq = a;  // the triangle list is shared, the row jobs are per device
            gr += b;
        }
        if (n_items_fq) *n_items_fq = fq;
        if (n_items_grad) *n_items_grad = gr;
        return 0;
    }
    if (!h) return fail(IID_E_BADARG, "null handle");
    if (n) *n = h->n;
    if (nq) *nq = h->nq;
    if (nr) *nr = h->nr;
    if (n_items_fq) *n_items_fq = h->n_items_tri;
    if (n_items_grad) *n_items_grad = h->n_jobs;
    return 0;
}

// ---------------------------------------------------------------------------
static int stage_positions(iid_handle *h, const double *pos_dev, cudaStream_t st)
{
    const int np = (int)h->np;
    prep_kernel<<<(np + 255) / 256, 256, 0, st>>>(pos_dev, h->orig, np,
                                                  h->precision == IID_FP32, h->x, h->y,
                                                  h->z, h->valid);
    ++h->launches;
    CU(cudaGetLastError());
    return 0;
}

template <int C, int MODE, int MAXT, int MINB, int TJ, bool CHEB, int PU = 1>
static int launch_debye2_v(iid_handle *h, const DebyeParams &p, dim3 grid, dim3 block, int nw,
                           cudaStream_t st)
{
    const size_t smem = 2 * debye2_buf_bytes(nw, TJ) +
                        (MODE == MODE_FORCE ? debye2_phi_bytes(nw, TJ) : 0);
    // function attributes are per device: remember which devices are done
    static bool attr_done[64] = {false};
    if (!attr_done[h->device & 63]) {
        CU(cudaFuncSetAttribute(debye2_kernel<C, MODE, MAXT, MINB, TJ, CHEB, PU>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done[h->device & 63] = true;
    }
    debye2_kernel<C, MODE, MAXT, MINB, TJ, CHEB, PU><<<grid, block, smem, st>>>(p);
    return 0;
}

template <int C, int MODE, bool CHEB>
static int launch_debye2_t(iid_handle *h, const DebyeParams &p, int64_t nblocks,
                           cudaStream_t st)
{
    // warps per block = Q chunks per block.  The gradient needs 4C accumulators
    // per thread: 8 warps at 220 registers, or 9-12 warps at 168 (a few spilled
    // words) when that lets ONE block cover the grid -- the 330-bin PDF grid is
    // 11 chunks: 12.1 ms at Au 10k against 15.8 ms as two 6-warp blocks that
    // each produce the pair records.  F(Q) / force blocks take up to 12 warps.
    const int nchunk = (int)((h->nq + C - 1) / C);
    const int nwmax = MODE == MODE_GRAD ? std::min(h->nw_max, h->grad_nw_max) : h->nw_max;
    const int gy = (nchunk + nwmax - 1) / nwmax;
    const int nw = (nchunk + gy - 1) / gy;
    dim3 grid((unsigned)nblocks, (unsigned)gy, 1), block(32 * nw, 1, 1);
    if (h->timing) CU(cudaEventRecord(h->ev0, st));
    int rc;
    if constexpr (MODE == MODE_GRAD) {
        // 4C accumulators: one block of <= 8 warps (255 registers) or <= 12 (168)
        // two pair set-ups side by side in the producer (prod_unroll, default on)
        rc = nw > 8 ? launch_debye2_v<C, MODE, 384, 1, 8, CHEB>(h, p, grid, block, nw, st)
             : h->prod_unroll
                 ? launch_debye2_v<C, MODE, 256, 1, 16, CHEB, 2>(h, p, grid, block, nw, st)
                 : launch_debye2_v<C, MODE, 256, 1, 16, CHEB, 1>(h, p, grid, block, nw, st);
    } else {
        // F(Q) / force: few accumulators, short dependent chains -> two blocks
        // per SM (8-j tiles keep two blocks' pair tables in shared memory)
        // (F(Q) also with 12 warps: 80 registers, 2 x 111 KB of pair tables)
        constexpr int MINB12 = MODE == MODE_FQ ? 2 : 1;
        rc = nw <= 8 ? launch_debye2_v<C, MODE, 256, 2, 8, CHEB>(h, p, grid, block, nw, st)
                     : launch_debye2_v<C, MODE, 384, MINB12, 8, CHEB>(h, p, grid, block, nw, st);
    }
    if (rc) return rc;
    ++h->launches;
    CU(cudaGetLastError());
    if (h->timing) {
        CU(cudaEventRecord(h->ev1, st));
        h->ev_pending = true;
    }
    return 0;
}

// FP64: producer/consumer kernel (iid_debye64.cuh).
template <int C, int MODE>
static int launch_debye64_t(iid_handle *h, const DebyeParams &p, int64_t nblocks,
                            cudaStream_t st)
{
    constexpr int TJ = 8;
    // F(Q) only keeps 16 accumulators per thread (116 registers): one 16-warp
    // block covers the 250-bin grid and the pair records are produced once
    // (3.27 -> 2.7 ms at Au 10k); gradient and force blocks stay at 8 warps
    constexpr int MAXW = MODE == MODE_FQ ? 16 : 8;
    constexpr int MAXT = 32 * MAXW;
    const int nchunk = (int)((h->nq + C - 1) / C);
    const int nwmax = MODE == MODE_FQ ? MAXW : std::min(h->nw_max, 8);
    const int gy = (nchunk + nwmax - 1) / nwmax;
    const int nw = (nchunk + gy - 1) / gy;
    dim3 grid((unsigned)nblocks, (unsigned)gy, 1), block(32 * nw, 1, 1);
    const size_t smem = 2 * debye64_buf_bytes(nw, TJ) +
                        (MODE == MODE_FORCE ? debye64_phi_bytes(nw, TJ) : 0);
    static bool attr_done[64] = {false};
    if (!attr_done[h->device & 63]) {
        CU(cudaFuncSetAttribute(debye64_kernel<C, MODE, MAXT, TJ>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done[h->device & 63] = true;
    }
    if (h->timing) CU(cudaEventRecord(h->ev0, st));
    debye64_kernel<C, MODE, MAXT, TJ><<<grid, block, smem, st>>>(p);
    ++h->launches;
    CU(cudaGetLastError());
    if (h->timing) {
        CU(cudaEventRecord(h->ev1, st));
        h->ev_pending = true;
    }
    return 0;
}

constexpr int C32 = 32;  // Q bins per warp, float32 kernels
constexpr int C64 = 16;  // Q bins per warp, float64 kernels

// F(Q) pass through the radial pair histogram (iid_fq_hist.cuh): FP32 mode,
// large structures.  Leaves this shard's pair sums in S; `gate` tells the direct
// kernel launched behind it whether it still has to run.
static bool fq_hist_applicable(const iid_handle *h)
{
    return h->fq_hist && h->precision == IID_FP32 && h->n >= h->fq_hist_min_n && h->qp <= 384 &&
           h->ntypes <= 5;
}

constexpr int FH_NCHUNK = (FH_CAP + FHT_E - 1) / FHT_E;

// Buffers of the histogram pass (sized per structure; before the first launch,
// outside any stream capture).
static int fq_hist_prepare(iid_handle *h)
{
    int rc;
    const int ntp = (int)(h->ntypes * (h->ntypes + 1) / 2);
    constexpr int NCHUNK = FH_NCHUNK;
    if (!h->hist_info) {
        if ((rc = dev_alloc(&h->hist_info, 8)) ||
            (rc = dev_alloc(&h->hist_C, (size_t)ntp * FH_CAP)) ||
            (rc = dev_alloc(&h->hist_Spart, (size_t)ntp * NCHUNK * h->qp)))
            return rc;
        CU(cudaMemset(h->hist_C, 0, (size_t)ntp * FH_CAP * sizeof(unsigned long long)));
        CU(cudaMemset(h->hist_info, 0, 8 * sizeof(double)));
        CU(cudaStreamSynchronize(0));
        static bool attr_done[64] = {false};
        if (!attr_done[h->device & 63]) {
            CU(cudaFuncSetAttribute(fq_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    2 * FH_CAP * (int)sizeof(int)));
            attr_done[h->device & 63] = true;
        }
    }
    return 0;
}

static int launch_fq_hist(iid_handle *h, double *S, cudaStream_t st)
{
    const int ntp = (int)(h->ntypes * (h->ntypes + 1) / 2);
    constexpr int NCHUNK = FH_NCHUNK;
    const double hstep = FT_QH / (h->qbin * (double)h->nq);
    fq_hist_grid_kernel<<<1, 1024, 0, st>>>(h->x, h->y, h->z, h->valid, (int)h->np, hstep,
                                            h->hist_info);
    HistParams p;
    p.x = h->x; p.y = h->y; p.z = h->z; p.valid = h->valid;
    p.tile_type = h->tile_type;
    p.items = h->items_tri;
    p.order = h->hist_order;
    p.n_items = (int)h->n_items_tri;
    p.rank = h->rank; p.world = h->world;
    p.info = h->hist_info;
    p.C = h->hist_C;
    p.stride = FH_CAP;
    fq_hist_kernel<<<h->sm_count, FH_THREADS, 2 * FH_CAP * sizeof(int), st>>>(p);
    fq_hist_transform_kernel<<<dim3(NCHUNK, ntp), 384, 0, st>>>(
        h->hist_C, FH_CAP, h->hist_info, reinterpret_cast<const float *>(h->ftab), (int)h->ntypes,
        (int)h->nq, (int)h->qp, h->qbin, h->hist_Spart);
    reduce_spart_kernel<<<(unsigned)((h->nq + 31) / 32), 1024, 0, st>>>(
        h->hist_Spart, ntp * NCHUNK, (int)h->nq, (int)h->qp, S);
    h->launches += 4;
    CU(cudaGetLastError());
    return 0;
}

static int launch_debye(iid_handle *h, int mode, void *G, double *S, const double *wq,
                        double *force, cudaStream_t st)
{
    DebyeParams p;
    p.x = h->x; p.y = h->y; p.z = h->z; p.valid = h->valid; p.orig = h->orig;
    p.tile_type = h->tile_type;
    const bool rows = mode == MODE_GRAD;
    p.items = h->items_tri;
    p.item_begin = h->rank;
    p.item_stride = h->world;
    p.ftab = h->ftab; p.inv_na = h->inv_na; p.wq = wq;
    p.nq = (int)h->nq; p.qp = (int)h->qp;
    p.qbin = h->qbin;
    p.qbin_turns = h->qbin / 6.283185307179586476925286766559;
    p.G = G; p.S = S; p.force = force;
    p.grad_split = h->grad_split ? 1 : 0;
    p.jobs = h->jobs; p.segs = h->segs; p.Gside = nullptr;
    int64_t mine;
    bool hist = false;
    if (rows) {
        // row jobs of this shard; S = per-job partial sums; the pieces of the
        // split rows go through the side buffer
        mine = h->n_jobs;
        const size_t esz = h->precision == IID_FP32 ? sizeof(float) : sizeof(double);
        const size_t side = (size_t)h->n_pieces * 32 * 3 * (size_t)h->nq * esz;
        if (side > h->Gside_bytes) {
            if (h->Gside) cudaFree(h->Gside);
            h->Gside = nullptr;
            h->Gside_bytes = 0;
            CU(cudaMalloc(&h->Gside, side));
            h->Gside_bytes = side;
        }
        p.Gside = h->Gside;
        p.Gscr = nullptr; p.slot_busy = nullptr; p.n_slots = 0; p.acc_j = 0;
        if (h->precision == IID_FP32 && h->acc_j > 0 && h->np > h->acc_j) {
            // scratch slots for the parked float32 partial sums (iid_debye2.cuh)
            const int nchunk = (int)((h->nq + C32 - 1) / C32);
            const int nwmax = std::min(h->nw_max, h->grad_nw_max);
            const int gy = (nchunk + nwmax - 1) / nwmax;
            const int nw = (nchunk + gy - 1) / gy;
            const int n_slots = h->sm_count * 4;
            const size_t cnt = (size_t)n_slots * 3 * C32 * 32 * nw;
            if (cnt > h->Gscr_count || n_slots != h->n_slots) {
                int rc = dev_alloc(&h->Gscr, cnt);
                if (!rc) rc = dev_alloc(&h->slot_busy, n_slots);
                h->Gscr_count = rc ? 0 : cnt;
                h->n_slots = rc ? 0 : n_slots;
                if (rc) return rc;
                CU(cudaMemsetAsync(h->slot_busy, 0, n_slots * sizeof(int), st));
            }
            p.Gscr = h->Gscr; p.slot_busy = h->slot_busy; p.n_slots = h->n_slots;
            p.acc_j = h->acc_j;
        }
        if (S) {
            const size_t cnt = (size_t)std::max<int64_t>(1, h->n_jobs) * h->qp;
            if (cnt > h->Spart_count) {
                int rc = dev_alloc(&h->Spart, cnt);
                h->Spart_count = rc ? 0 : cnt;
                if (rc) return rc;
            }
            p.S = h->Spart;
        }
    } else {
        const int64_t nitems = h->n_items_tri;
        mine = nitems > h->rank ? (nitems - h->rank + h->world - 1) / h->world : 0;
        hist = mode == MODE_FQ && S && mine > 0 && fq_hist_applicable(h);
        if (hist) {
            // radial pair histogram (iid_fq_hist.cuh); the direct kernel below only
            // runs if the structure does not fit it (gate), adding into S
            int rc0 = fq_hist_prepare(h);
            if (rc0) return rc0;
            p.gate = h->hist_info;
        } else if (mode == MODE_FQ && h->det_fq && S && mine > 0) {
            // deterministic F(Q): every item stores its partial sums, added in
            // item order by reduce_spart_kernel (no atomics)
            const size_t cnt = (size_t)mine * h->qp;
            if (cnt > h->Sitem_fq_count) {
                int rc = dev_alloc(&h->Sitem_fq, cnt);
                h->Sitem_fq_count = rc ? 0 : cnt;
                if (rc) return rc;
            }
            p.Sitem = h->Sitem_fq;
        }
    }
    h->last_pairq = 0.5 * (double)h->n * (double)(h->n - 1) * (double)h->nq / h->world;
    int rc = 0;
    const bool timing = h->timing;
    if (hist) {
        if (timing) CU(cudaEventRecord(h->ev0, st));
        if ((rc = launch_fq_hist(h, S, st))) return rc;
        h->timing = false;  // (the events bracket the whole pass, not the gated kernel)
    }
    if (mine > 0) {
        if (h->precision == IID_FP32 && h->cheb) {
            if (mode == MODE_FQ) rc = launch_debye2_t<C32, MODE_FQ, true>(h, p, mine, st);
            else if (mode == MODE_GRAD) rc = launch_debye2_t<C32, MODE_GRAD, true>(h, p, mine, st);
            else rc = launch_debye2_t<C32, MODE_FORCE, true>(h, p, mine, st);
        } else if (h->precision == IID_FP32) {  // IID_CHEB=0: rotation recurrence everywhere
            if (mode == MODE_FQ) rc = launch_debye2_t<C32, MODE_FQ, false>(h, p, mine, st);
            else if (mode == MODE_GRAD) rc = launch_debye2_t<C32, MODE_GRAD, false>(h, p, mine, st);
            else rc = launch_debye2_t<C32, MODE_FORCE, false>(h, p, mine, st);
        } else {
            if (mode == MODE_FQ) rc = launch_debye64_t<C64, MODE_FQ>(h, p, mine, st);
            else if (mode == MODE_GRAD) rc = launch_debye64_t<C64, MODE_GRAD>(h, p, mine, st);
            else rc = launch_debye64_t<C64, MODE_FORCE>(h, p, mine, st);
        }
        if (rc) return rc;
    }
    if (hist) {
        h->timing = timing;
        if (timing) {
            CU(cudaEventRecord(h->ev1, st));
            h->ev_pending = true;
        }
    }
    if (!rows && p.Sitem != nullptr) {
        reduce_spart_kernel<<<(unsigned)((h->nq + 31) / 32), 1024, 0, st>>>(
            h->Sitem_fq, (int)mine, (int)h->nq, (int)h->qp, S);
        ++h->launches;
        CU(cudaGetLastError());
    }
    if (rows) {
        if (h->n_fixes > 0) {
            if (h->precision == IID_FP32)
                rows_fixup_kernel<float><<<dim3((unsigned)h->n_fixes, FIX_SPLIT), 256, 0, st>>>(
                    h->fixes, h->orig, (const float *)h->Gside, (int)h->nq, (float *)G);
            else
                rows_fixup_kernel<double><<<dim3((unsigned)h->n_fixes, FIX_SPLIT), 256, 0, st>>>(
                    h->fixes, h->orig, (const double *)h->Gside, (int)h->nq, (double *)G);
            ++h->launches;
            CU(cudaGetLastError());
        }
        if (S) {
            reduce_spart_kernel<<<(unsigned)((h->nq + 31) / 32), 1024, 0, st>>>(
                h->Spart, (int)h->n_jobs, (int)h->nq, (int)h->qp, S);
            ++h->launches;
            CU(cudaGetLastError());
        }
    }
    return 0;
}

// Force pass.  FP32 mode with a handful of element types: sum over the Q bins
// first (radial table, O(K Q)), then one interpolation per ordered pair
// (O(N^2)); otherwise the direct O(N^2 Q) kernel.
static int launch_force(iid_handle *h, const double *wq, double *force, cudaStream_t st)
{
    // (phi_table_kernel keeps one weight per Q bin in dynamic shared memory)
    const bool table = h->use_force_table && h->precision == IID_FP32 && h->ntypes <= 4 &&
                       h->n >= h->force_table_min_n && h->nq * sizeof(double) <= 48 * 1024;
    if (!table) return launch_debye(h, MODE_FORCE, nullptr, nullptr, wq, force, st);
    const size_t tstride = PHI_KMAX + 2 * PHI_PAD;
    if (!h->phi_tab) {
        CU(cudaMalloc((void **)&h->phi_tab, sizeof(float) * tstride * 16));
        CU(cudaMalloc((void **)&h->phi_info, sizeof(double) * 4));
    }
    const int np = (int)h->np, E = (int)h->ntypes;
    const double qmax = h->qbin * (double)h->nq;
    const double h_target = 0.05 / (qmax > 0.0 ? qmax : 1.0);
    h->last_pairq = 0.5 * (double)h->n * (double)(h->n - 1) * (double)h->nq / h->world;
    if (h->timing) CU(cudaEventRecord(h->ev0, st));
    phi_grid_kernel<<<1, 1024, 0, st>>>(h->x, h->y, h->z, h->valid, np, h_target, h->phi_info);
    phi_table_kernel<<<dim3((unsigned)(4 * h->sm_count), (unsigned)(E * E)), 128,
                       h->nq * sizeof(double), st>>>(wq, (const float *)h->ftab,
                                                     (const float *)h->inv_na, (int)h->nq,
                                                     (int)h->qp, E, h->qbin, h->phi_info,
                                                     h->phi_tab);
    const int rows = (np + FT_BLOCK - 1) / FT_BLOCK;
    const int mine = rows > h->rank ? (rows - h->rank + h->world - 1) / h->world : 0;
    if (mine > 0) {
        const int jsplit =
            std::max(1, std::min(np / TILE_I, (4 * h->sm_count + mine - 1) / mine));
        // deterministic: the j shares of a row store their partial forces, added in
        // share order afterwards (no atomics)
        double *fpart = nullptr;
        if (h->det_fq) {
            const size_t cnt = (size_t)jsplit * np * 3;
            if (cnt > h->ft_part_count) {
                int rc = dev_alloc(&h->ft_part, cnt);
                h->ft_part_count = rc ? 0 : cnt;
                if (rc) return rc;
            }
            fpart = h->ft_part;
        }
        force_table_kernel<<<dim3((unsigned)mine, (unsigned)jsplit), FT_BLOCK, 0, st>>>(
            h->x, h->y, h->z, h->valid, h->orig, h->tile_type, np, E, h->phi_info, h->phi_tab,
            jsplit, h->rank, h->world, force, wq, (const float *)h->ftab, (const float *)h->inv_na,
            (int)h->nq, (int)h->qp, h->qbin, fpart);
        if (fpart) {
            force_table_reduce_kernel<<<(3 * np + 255) / 256, 256, 0, st>>>(
                fpart, jsplit, np, h->orig, h->rank, h->world, force);
            ++h->launches;
        }
    }
    h->launches += 3;
    CU(cudaGetLastError());
    if (h->timing) {
        CU(cudaEventRecord(h->ev1, st));
        h->ev_pending = true;
    }
    return 0;
}

static cudaStream_t pick(iid_handle *h, void *stream)
{
    return stream ? (cudaStream_t)stream : h->stream;
}

extern "C" int iid_fq_partial(iid_handle *h, const double *pos_dev, double *S_dev, void *stream)
{
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (!pos_dev || !S_dev) return fail(IID_E_BADARG, "null pointer");
    cudaStream_t st = pick(h, stream);
    int rc = stage_positions(h, pos_dev, st);
    if (rc) return rc;
    CU(cudaMemsetAsync(S_dev, 0, h->nq * sizeof(double), st));
    return launch_debye(h, MODE_FQ, nullptr, S_dev, nullptr, nullptr, st);
}

extern "C" int iid_fq_finish(iid_handle *h, const double *S_dev, double *F_dev, void *stream)
{
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (!S_dev || !F_dev) return fail(IID_E_BADARG, "null pointer");
    cudaStream_t st = pick(h, stream);
    finish_fq_kernel<<<(int)(h->nq + 127) / 128, 128, 0, st>>>(S_dev, h->inv_na_d, (int)h->nq, 0, F_dev);
    ++h->launches;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int iid_grad_fq_partial(iid_handle *h, const double *pos_dev, void *G_dev,
                                   double *S_dev, void *stream)
{
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (!pos_dev || !G_dev) return fail(IID_E_BADARG, "null pointer");
    cudaStream_t st = pick(h, stream);
    int rc = stage_positions(h, pos_dev, st);
    if (rc) return rc;
    // every row of this shard's i-tiles is written exactly once (plain stores):
    // no zero-fill; rows of other shards are not touched
    return launch_debye(h, MODE_GRAD, G_dev, S_dev, nullptr, nullptr, st);
}

extern "C" int iid_force_partial(iid_handle *h, const double *pos_dev, const double *wq_dev,
                                 double *force_dev, void *stream)
{
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (!pos_dev || !wq_dev || !force_dev) return fail(IID_E_BADARG, "null pointer");
    cudaStream_t st = pick(h, stream);
    int rc = stage_positions(h, pos_dev, st);
    if (rc) return rc;
    CU(cudaMemsetAsync(force_dev, 0, (size_t)h->n * 3 * sizeof(double), st));
    return launch_force(h, wq_dev, force_dev, st);
}

extern "C" int iid_fq_to_gr(iid_handle *h, const double *F_dev, double *G_dev, void *stream)
{
    NEED(h);
    if (h->nr == 0) return fail(IID_E_NOTRANSFORM, "call iid_set_transform first");
    if (!F_dev || !G_dev) return fail(IID_E_BADARG, "null pointer");
    cudaStream_t st = pick(h, stream);
    const int64_t threads = h->nr * 32;
    gr_kernel<double><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(h->T, F_dev, h->nr,
                                                                 (int)h->nq, (int)h->qp, G_dev);
    ++h->launches;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int iid_potential(iid_handle *h, const double *G_dev, const double *target_dev,
                             int potential, double conv, double *out_dev, double *wq_dev,
                             void *stream)
{
    NEED(h);
    if (h->nr == 0) return fail(IID_E_NOTRANSFORM, "call iid_set_transform first");
    if (!G_dev || !target_dev || !out_dev) return fail(IID_E_BADARG, "null pointer");
    if (potential != IID_POT_RW && potential != IID_POT_CHI_SQ)
        return fail(IID_E_BADARG, "unknown potential");
    cudaStream_t st = pick(h, stream);
    potential_kernel<<<1, 1024, 0, st>>>(G_dev, target_dev, (int)h->nr, potential, conv,
                                         out_dev, h->cr, nullptr);
    ++h->launches;
    CU(cudaGetLastError());
    if (wq_dev) {
        CU(cudaMemsetAsync(wq_dev, 0, h->nq * sizeof(double), st));
        wq_kernel<<<(unsigned)((h->nr + WQ_ROWS - 1) / WQ_ROWS), 352, 0, st>>>(
            h->T, h->cr, (int)h->nr, (int)h->nq, (int)h->qp, conv, wq_dev);
        ++h->launches;
        CU(cudaGetLastError());
    }
    return 0;
}

extern "C" int iid_grad_pdf(iid_handle *h, const void *grad_fq_dev, int64_t rows,
                            double *grad_pdf_dev, void *stream)
{
    if (MULTI(h)) return iid_grad_pdf(h->subs[0], grad_fq_dev, rows, grad_pdf_dev, stream);
    NEED(h);
    if (h->nr == 0) return fail(IID_E_NOTRANSFORM, "call iid_set_transform first");
    if (!grad_fq_dev || !grad_pdf_dev || rows < 0) return fail(IID_E_BADARG, "bad argument");
    if (rows == 0) return 0;
    cudaStream_t st = pick(h, stream);
    dim3 grid((unsigned)((h->nr + GP_BN - 1) / GP_BN), (unsigned)((rows + GP_BM - 1) / GP_BM));
    if (h->precision == IID_FP32)
        grad_pdf_kernel<float><<<grid, 256, 0, st>>>((const float *)grad_fq_dev, h->T, rows,
                                                     (int)h->nq, (int)h->qp, (int)h->nr,
                                                     grad_pdf_dev);
    else
        grad_pdf_kernel<double><<<grid, 256, 0, st>>>((const double *)grad_fq_dev, h->T, rows,
                                                      (int)h->nq, (int)h->qp, (int)h->nr,
                                                      grad_pdf_dev);
    ++h->launches;
    CU(cudaGetLastError());
    return 0;
}

// Device -> pinned staging in chunks; each chunk is copied on to the caller's
// (pageable, usually freshly allocated) array by a few host threads while the
// next chunks are still in flight.  A direct cudaMemcpy into pageable memory
// runs at ~5 GB/s here; this path at ~15-20 GB/s.  Enqueued on h->stream; the
// data is complete when the function returns.
static int download_pipelined(iid_handle *h, const void *dev, void *host, size_t bytes)
{
    static const size_t CH = []() {
        const char *e = getenv("IID_DL_CHUNK_MB");
        return (size_t)std::max(1, e ? atoi(e) : 32) << 20;
    }();
    static const int nthreads = []() {
        const char *e = getenv("IID_DL_THREADS");
        return std::max(1, std::min(32, e ? atoi(e) : 8));
    }();
    const size_t nchunks = (bytes + CH - 1) / CH;
    if (bytes > h->pinG_bytes) {
        if (h->pinG) cudaFreeHost(h->pinG);
        h->pinG = nullptr;
        h->pinG_bytes = 0;
        if (cudaMallocHost((void **)&h->pinG, bytes) == cudaSuccess) h->pinG_bytes = bytes;
        else cudaGetLastError();
    }
    if (h->pinG_bytes >= bytes && nchunks > 1) {
        while (h->chunk_ev.size() < nchunks) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            h->chunk_ev.push_back(e);
        }
        for (size_t c = 0; c < nchunks; ++c) {
            const size_t off = c * CH, len = std::min(CH, bytes - off);
            CU(cudaMemcpyAsync(h->pinG + off, (const unsigned char *)dev + off, len,
                               cudaMemcpyDeviceToHost, h->stream));
            CU(cudaEventRecord(h->chunk_ev[c], h->stream));
        }
        for (size_t c = 0; c < nchunks; ++c) {
            const size_t off = c * CH, len = std::min(CH, bytes - off);
            CU(cudaEventSynchronize(h->chunk_ev[c]));
            std::vector<std::thread> workers((size_t)nthreads);
            const size_t part = (len + nthreads - 1) / nthreads;
            for (int t = 0; t < nthreads; ++t) {
                const size_t o = (size_t)t * part;
                const size_t l = o < len ? std::min(part, len - o) : 0;
                unsigned char *dst = (unsigned char *)host + off + o;
                const unsigned char *src = h->pinG + off + o;
                workers[t] = std::thread([=]() {
                    if (l) memcpy(dst, src, l);
                });
            }
            for (int t = 0; t < nthreads; ++t) workers[t].join();
        }
    } else {
        CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

// Copy a device array of this handle's device to pageable host memory through
// the pipelined pinned staging (ordered after the work enqueued on the
// handle's stream).
extern "C" int iid_download_host(iid_handle *h, const void *dev, void *host, int64_t bytes)
{
    if (MULTI(h)) return iid_download_host(h->subs[0], dev, host, bytes);
    NEED(h);
    if (!dev || !host || bytes < 0) return fail(IID_E_BADARG, "bad argument");
    if (bytes == 0) return 0;
    return download_pipelined(h, dev, host, (size_t)bytes);
}

// --- host-buffer entry points ------------------------------------------------
// pinned layout: [0,3n) positions in, [3n,6n) forces out, then F (qp), out4 (8),
// G(r) (nr), target (nr)
static int upload_positions(iid_handle *h, const double *pos_host)
{
    memcpy(h->pin, pos_host, (size_t)3 * h->n * sizeof(double));
    CU(cudaMemcpyAsync(h->pos, h->pin, (size_t)3 * h->n * sizeof(double),
                       cudaMemcpyHostToDevice, h->stream));
    return 0;
}

extern "C" int iid_fq_host(iid_handle *h, const double *pos_host, double *F_host)
{
    if (MULTI(h)) return multi_fq_host(h, pos_host, F_host, nullptr);
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (!pos_host || !F_host) return fail(IID_E_BADARG, "null pointer");
    int rc;
    if ((rc = upload_positions(h, pos_host))) return rc;
    if ((rc = iid_fq_partial(h, h->pos, h->S, nullptr))) return rc;
    if ((rc = iid_fq_finish(h, h->S, h->F, nullptr))) return rc;
    double *pf = h->pin + 6 * h->n;
    CU(cudaMemcpyAsync(pf, h->F, h->nq * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(F_host, pf, h->nq * sizeof(double));
    return 0;
}

// Device alias of a host pointer the GPU can write directly (pinned + mapped:
// iid_host_alloc / iid_host_register), else nullptr.
static void *mapped_alias(const void *host)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

extern "C" int iid_grad_fq_host(iid_handle *h, const double *pos_host, void *G_host,
                                double *F_host)
{
    if (MULTI(h)) return multi_grad_fq_host(h, pos_host, G_host, F_host);
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (!pos_host || !G_host) return fail(IID_E_BADARG, "null pointer");
    const size_t esz = h->precision == IID_FP32 ? sizeof(float) : sizeof(double);
    const size_t bytes = (size_t)h->n * 3 * h->nq * esz;
    // Pinned, mapped host memory: the kernel stores every gradient row once,
    // in coalesced 128-byte pieces, straight into the caller's array (0.8 GB/s
    // of posted PCIe writes spread over the whole pass) -- no device copy of
    // the gradient, no download.  Pageable memory: device array + pipelined
    // download through pinned staging.
    void *alias = h->zero_copy ? mapped_alias(G_host) : nullptr;
    if (!alias && bytes > h->Gfull_bytes) {
        if (h->Gfull) cudaFree(h->Gfull);
        h->Gfull = nullptr;
        h->Gfull_bytes = 0;
        CU(cudaMalloc(&h->Gfull, bytes));
        h->Gfull_bytes = bytes;
    }
    int rc;
    if ((rc = upload_positions(h, pos_host))) return rc;
    if ((rc = iid_grad_fq_partial(h, h->pos, alias ? alias : h->Gfull, h->S, nullptr))) return rc;
    if ((rc = iid_fq_finish(h, h->S, h->F, nullptr))) return rc;
    double *pf = h->pin + 6 * h->n;
    CU(cudaMemcpyAsync(pf, h->F, h->nq * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (!alias && (rc = download_pipelined(h, h->Gfull, G_host, bytes))) return rc;
    CU(cudaStreamSynchronize(h->stream));
    if (F_host) memcpy(F_host, pf, h->nq * sizeof(double));
    return 0;
}

// Pinned, mapped, portable host memory: an output array allocated here (or a
// host mapping registered here, e.g. POSIX shared memory that several ranks
// write their rows into) is written directly by the gradient kernel.
extern "C" int iid_host_alloc(int64_t bytes, void **ptr)
{
    if (!ptr || bytes < 1) return fail(IID_E_BADARG, "bad argument");
    *ptr = nullptr;
    CU(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    return 0;
}

extern "C" int iid_host_free(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return 0;
}

extern "C" int iid_host_register(void *ptr, int64_t bytes)
{
    if (!ptr || bytes < 1) return fail(IID_E_BADARG, "bad argument");
    CU(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return 0;
}

extern "C" int iid_host_unregister(void *ptr)
{
    if (ptr) CU(cudaHostUnregister(ptr));
    return 0;
}

// device pointer under which the current device writes to a mapped host array
extern "C" int iid_host_device_pointer(void *host, void **dev)
{
    if (!host || !dev) return fail(IID_E_BADARG, "null pointer");
    *dev = mapped_alias(host);
    if (!*dev) return fail(IID_E_BADARG, "not pinned, mapped host memory");
    return 0;
}

extern "C" int iid_pdf_host(iid_handle *h, const double *pos_host, double *pdf_host,
                            double *F_host)
{
    if (MULTI(h)) {
        if (!pdf_host) return fail(IID_E_BADARG, "null pointer");
        return multi_fq_host(h, pos_host, F_host, pdf_host);
    }
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (h->nr == 0) return fail(IID_E_NOTRANSFORM, "call iid_set_transform first");
    if (!pos_host || !pdf_host) return fail(IID_E_BADARG, "null pointer");
    int rc;
    if ((rc = upload_positions(h, pos_host))) return rc;
    if ((rc = iid_fq_partial(h, h->pos, h->S, nullptr))) return rc;
    if ((rc = iid_fq_finish(h, h->S, h->F, nullptr))) return rc;
    if ((rc = iid_fq_to_gr(h, h->F, h->Gr, nullptr))) return rc;
    double *pf = h->pin + 6 * h->n;
    double *pg = pf + h->qp + 8;
    CU(cudaMemcpyAsync(pf, h->F, h->nq * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(pg, h->Gr, h->nr * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(pdf_host, pg, h->nr * sizeof(double));
    if (F_host) memcpy(F_host, pf, h->nq * sizeof(double));
    return 0;
}

// --- spring restraints (iid_spring.cuh) ----------------------------------------
static int launch_spring(iid_handle *h, const double *pos, int64_t n, int sp_type, double k,
                         double rt, const double *com, double *energy, double *force,
                         double *atomwise, cudaStream_t st)
{
    if (n <= 0) return 0;
    const int rows = (int)((n + SP_BLOCK - 1) / SP_BLOCK);
    const int my_rows = (rows - h->rank + h->world - 1) / h->world;  // rows rank, rank+W, ...
    if (my_rows <= 0) return 0;
    const bool f32 = h->precision == IID_FP32;
    if (sp_type == SPRING_COM) {
        if (!com) return fail(IID_E_BADARG, "the com spring needs the centre of mass");
        if (f32)
            spring_com_kernel<true><<<my_rows, SP_BLOCK, 0, st>>>(
                pos, (int)n, k, rt, com[0], com[1], com[2], h->rank, h->world, energy, force,
                atomwise);
        else
            spring_com_kernel<false><<<my_rows, SP_BLOCK, 0, st>>>(
                pos, (int)n, k, rt, com[0], com[1], com[2], h->rank, h->world, energy, force,
                atomwise);
    } else {
        // split the j range until the grid covers the SMs about twice
        const int units = (int)((n + SP_JUNIT - 1) / SP_JUNIT);
        int jsplit = std::max(1, std::min(units, (2 * h->sm_count + my_rows - 1) / my_rows));
        const dim3 grid((unsigned)my_rows, (unsigned)jsplit);
        const int att = sp_type == SPRING_ATT;
        if (f32)
            spring_pair_kernel<true><<<grid, SP_BLOCK, 0, st>>>(
                pos, (int)n, att, k, rt, jsplit, h->rank, h->world, energy, force, atomwise);
        else
            spring_pair_kernel<false><<<grid, SP_BLOCK, 0, st>>>(
                pos, (int)n, att, k, rt, jsplit, h->rank, h->world, energy, force, atomwise);
    }
    ++h->launches;
    CU(cudaGetLastError());
    return 0;
}

static int check_spring(int sp_type, int64_t n)
{
    if (sp_type != SPRING_REP && sp_type != SPRING_COM && sp_type != SPRING_ATT)
        return fail(IID_E_BADARG, "unknown spring type");
    if (n < 0 || n > (int64_t)1 << 30) return fail(IID_E_BADARG, "bad atom count");
    return 0;
}

extern "C" int iid_spring_partial(iid_handle *h, const double *pos_dev, int64_t n, int sp_type,
                                  double k, double rt, const double *com, double *energy_dev,
                                  double *force_dev, double *atomwise_dev, void *stream)
{
    NEED(h);
    int rc;
    if ((rc = check_spring(sp_type, n))) return rc;
    if (n > 0 && !pos_dev) return fail(IID_E_BADARG, "null positions");
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (energy_dev) CU(cudaMemsetAsync(energy_dev, 0, sizeof(double), st));
    if (force_dev && n) CU(cudaMemsetAsync(force_dev, 0, (size_t)n * 3 * sizeof(double), st));
    if (atomwise_dev && n) CU(cudaMemsetAsync(atomwise_dev, 0, (size_t)n * sizeof(double), st));
    return launch_spring(h, pos_dev, n, sp_type, k, rt, com, energy_dev, force_dev, atomwise_dev,
                         st);
}

static int spring_scratch(iid_handle *h, size_t count)
{
    if (count <= h->sp_count) return 0;
    int rc = dev_alloc(&h->sp_buf, count);
    h->sp_count = rc ? 0 : count;
    return rc;
}

extern "C" int iid_spring_host(iid_handle *h, const double *pos_host, int64_t n, int sp_type,
                               double k, double rt, const double *com, double *energy_host,
                               double *forces_host, double *atomwise_host)
{
    if (MULTI(h)) {
        // O(N^2) with a tiny constant: one device, all rows
        iid_handle *s0 = h->subs[0];
        const int r = s0->rank, w = s0->world;
        s0->rank = 0;
        s0->world = 1;
        const int rc = iid_spring_host(s0, pos_host, n, sp_type, k, rt, com, energy_host,
                                       forces_host, atomwise_host);
        s0->rank = r;
        s0->world = w;
        return rc;
    }
    NEED(h);
    int rc;
    if ((rc = check_spring(sp_type, n))) return rc;
    if (n > 0 && !pos_host) return fail(IID_E_BADARG, "null positions");
    if (h->world != 1)
        return fail(IID_E_BADARG, "iid_spring_host needs all rows (world == 1)");
    if (n == 0) {
        if (energy_host) *energy_host = 0.0;
        return 0;
    }
    // [0,3n) positions, [3n,6n) force, [6n,7n) atomwise, [7n] energy
    if ((rc = spring_scratch(h, (size_t)7 * n + 1))) return rc;
    double *b = h->sp_buf;
    CU(cudaMemcpyAsync(b, pos_host, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice,
                       h->stream));
    if ((rc = iid_spring_partial(h, b, n, sp_type, k, rt, com, energy_host ? b + 7 * n : nullptr,
                                 forces_host ? b + 3 * n : nullptr,
                                 atomwise_host ? b + 6 * n : nullptr, nullptr)))
        return rc;
    if (energy_host)
        CU(cudaMemcpyAsync(energy_host, b + 7 * n, sizeof(double), cudaMemcpyDeviceToHost,
                           h->stream));
    if (forces_host)
        CU(cudaMemcpyAsync(forces_host, b + 3 * n, (size_t)3 * n * sizeof(double),
                           cudaMemcpyDeviceToHost, h->stream));
    if (atomwise_host)
        CU(cudaMemcpyAsync(atomwise_host, b + 6 * n, (size_t)n * sizeof(double),
                           cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int iid_spring_voxel_host(iid_handle *h, const double *pos_host, int64_t n, int sp_type,
                                     double k, double rt, const double *com, double resolution,
                                     int64_t nx, int64_t ny, int64_t nz, double *voxels_host)
{
    if (MULTI(h))
        return iid_spring_voxel_host(h->subs[0], pos_host, n, sp_type, k, rt, com, resolution, nx,
                                     ny, nz, voxels_host);
    NEED(h);
    int rc;
    if ((rc = check_spring(sp_type, n))) return rc;
    if (!voxels_host || nx < 0 || ny < 0 || nz < 0 || !(resolution > 0.0))
        return fail(IID_E_BADARG, "bad voxel grid");
    if (n > 0 && !pos_host) return fail(IID_E_BADARG, "null positions");
    if (sp_type == SPRING_COM && !com)
        return fail(IID_E_BADARG, "the com spring needs the centre of mass");
    const int64_t nv = nx * ny * nz;
    if (nv == 0) return 0;
    if (nx > 65535 || ny > 65535 || nz > 65535 || nv > (int64_t)1 << 31)
        return fail(IID_E_BADARG, "voxel grid too large");
    if ((rc = spring_scratch(h, (size_t)3 * n + (size_t)nv))) return rc;
    double *b = h->sp_buf, *vox = b + 3 * n;
    if (n)
        CU(cudaMemcpyAsync(b, pos_host, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice,
                           h->stream));
    const unsigned blocks = (unsigned)((nv + SP_BLOCK - 1) / SP_BLOCK);
    const double c0 = com ? com[0] : 0.0, c1 = com ? com[1] : 0.0, c2 = com ? com[2] : 0.0;
    if (h->precision == IID_FP32)
        spring_voxel_kernel<true><<<blocks, SP_BLOCK, 0, h->stream>>>(
            b, (int)n, sp_type, k, rt, c0, c1, c2, resolution, (int)nx, (int)ny, (int)nz, vox);
    else
        spring_voxel_kernel<false><<<blocks, SP_BLOCK, 0, h->stream>>>(
            b, (int)n, sp_type, k, rt, c0, c1, c2, resolution, (int)nx, (int)ny, (int)nz, vox);
    ++h->launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(voxels_host, vox, (size_t)nv * sizeof(double), cudaMemcpyDeviceToHost,
                       h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int iid_set_restraints(iid_handle *h, int count, const int *sp_type, const double *k,
                                  const double *rt)
{
    if (MULTI(h)) {
        for (iid_handle *sh : h->subs) {
            int rc = iid_set_restraints(sh, count, sp_type, k, rt);
            if (rc) return rc;
        }
        h->n_restraints = count;
        for (int s = 0; s < count; ++s) {  // for devices that join later
            h->rs_type[s] = sp_type[s];
            h->multi_rs_k[s] = k[s];
            h->rs_rt[s] = rt[s];
        }
        return 0;
    }
    if (!h) return fail(IID_E_BADARG, "null handle");
    if (count < 0 || count > IID_MAX_RESTRAINTS)
        return fail(IID_E_BADARG, "at most IID_MAX_RESTRAINTS restraints");
    if (count && (!sp_type || !k || !rt)) return fail(IID_E_BADARG, "null pointer");
    for (int s = 0; s < count; ++s)
        if (sp_type[s] != SPRING_REP && sp_type[s] != SPRING_ATT)
            return fail(IID_E_BADARG, "only rep / att springs can be fused");
    bool same = count == h->n_restraints;
    for (int s = 0; same && s < count; ++s)
        same = sp_type[s] == h->rs_type[s] && k[s] == h->rs_k[s] && rt[s] == h->rs_rt[s];
    if (same) return 0;
    h->n_restraints = count;
    for (int s = 0; s < count; ++s) {
        h->rs_type[s] = sp_type[s];
        h->rs_k[s] = k[s];
        h->rs_rt[s] = rt[s];
    }
    drop_graph(h);  // k / rt are kernel arguments baked into the graph
    return 0;
}

extern "C" int iid_get_restraint_energy(iid_handle *h, double *energy)
{
    if (MULTI(h)) {
        if (!energy) return fail(IID_E_BADARG, "null argument");
        *energy = h->n_restraints ? h->rs_k[0] : 0.0;  // summed over the devices by the last call
        return 0;
    }
    if (!h || !energy) return fail(IID_E_BADARG, "null argument");
    *energy = h->n_restraints && h->pin ? h->pin[6 * h->n + h->qp + 4] : 0.0;
    return 0;
}

// The device part of one Calc1D evaluation, positions already in h->pos:
// F(Q) pass -> F -> G(r) -> Rw / chi^2 (h->out4) -> chain-rule weights -> force
// pass (h->force) -> fused restraints.  Enqueued on the handle's stream; shared
// by iid_energy_forces_host and iid_leapfrog_host (both replay it from a graph).
// phase: EVAL_ALL, or the two halves around the host-side sum of the F(Q)
// partials of a multi-device handle (EVAL_FQ: staging + F(Q) pass; EVAL_REST:
// everything after it, h->S already holding the summed pair sums).
enum { EVAL_ALL = 0, EVAL_FQ = 1, EVAL_REST = 2 };

// The whole evaluation as ONE cooperative launch (iid_fused.cuh) when the
// structure is small enough for one work item per SM: FP32 mode, one shard,
// direct force pass, Q-space weights, the caller wants forces and no G(r).
// Dynamic shared memory of the fused launch: pair records (two buffers), the
// force pass's per-pair scalars, and the positions of the item's atoms.
static size_t fused_smem_bytes(const iid_handle *h, int *ps_off, int *pl)
{
    const int nchunk = (int)((h->nq + C32 - 1) / C32);
    int off = (int)(2 * debye2_buf_bytes(nchunk, 8) + debye2_phi_bytes(nchunk, 8));
    // the Q-space stages and the table force pass reuse the front of it: F, M F,
    // a reduction row, the weights of each element pair, the warps' i-side forces
    const int64_t front = ((2 + h->ntypes * h->ntypes) * h->qp + 32 + (int64_t)nchunk * 96) * 8;
    off = (int)std::max<int64_t>(off, (front + 15) / 16 * 16);
    const int len = (int)((TILE_I + h->tri_maxlen + 1) / 2 * 2);
    if (ps_off) *ps_off = off;
    if (pl) *pl = len;
    return (size_t)off + (4 * (size_t)len + 3 * (size_t)h->n) * sizeof(double);
}

static bool fused_applicable(const iid_handle *h, bool want_forces, bool want_pdf)
{
    const bool table = h->use_force_table && h->ntypes <= 4 && h->n >= h->force_table_min_n;
    // (with its own in-launch radial table the fused kernel also covers the sizes
    // where the launch sequence would switch to the tabulated force pass)
    const bool own_table = h->fused_table && h->ntypes <= 2;
    return h->use_fused && h->precision == IID_FP32 && h->cheb && h->world == 1 &&
           want_forces && !want_pdf && h->qspace_wq && (!table || own_table) && !h->timing &&
           h->n_items_tri <= h->sm_count && (h->nq + C32 - 1) / C32 <= 12 &&
           h->np <= (int64_t)h->sm_count * 384 &&
           fused_smem_bytes(h, nullptr, nullptr) <= 220 * 1024;
}

static int launch_fused(iid_handle *h, int potential, double conv, bool lf)
{
    int rc;
    if (!h->MF) {
        if ((rc = dev_alloc(&h->MF, h->qp)) ||
            (rc = dev_alloc(&h->wq_blk, (size_t)h->sm_count * h->qp)))
            return rc;
    }
    FusedParams q;
    DebyeParams &p = q.fq;
    p.x = h->x; p.y = h->y; p.z = h->z; p.valid = h->valid; p.orig = h->orig;
    p.tile_type = h->tile_type;
    p.items = h->items_tri;
    p.item_begin = 0;
    p.item_stride = 1;
    p.ftab = h->ftab; p.inv_na = h->inv_na; p.wq = nullptr;
    p.nq = (int)h->nq; p.qp = (int)h->qp;
    p.qbin = h->qbin;
    p.qbin_turns = h->qbin / 6.283185307179586476925286766559;
    p.G = nullptr; p.S = h->S; p.force = nullptr;
    p.grad_split = 1;
    p.jobs = nullptr; p.segs = nullptr; p.Gside = nullptr;
    p.Gscr = nullptr; p.slot_busy = nullptr; p.n_slots = 0; p.acc_j = 0;
    if (!h->Sfix) {
        if ((rc = dev_alloc(&h->Sfix, 2 * h->qp)) || (rc = dev_alloc(&h->Ffix, 3 * h->n)) ||
            (rc = dev_alloc(&h->gforce, h->qp)))
            return rc;
        CU(cudaMemset(h->Sfix, 0, 2 * h->qp * sizeof(unsigned long long)));
        CU(cudaMemset(h->Ffix, 0, 3 * h->n * sizeof(unsigned long long)));
        CU(cudaMemcpy(h->gforce, h->gforce_host.data(), h->qp * sizeof(double),
                      cudaMemcpyHostToDevice));
        if ((rc = dev_alloc(&h->ext_ref, 8))) return rc;
        CU(cudaMemset(h->ext_ref, 0, 8 * sizeof(double)));
        CU(cudaStreamSynchronize(0));
    }
    p.Sitem = nullptr;
    p.Sfix = h->Sfix;
    p.Ffix = nullptr;
    p.fix_scale = h->s_fix_scale;
    q.fo = p;
    q.fo.S = nullptr;
    q.fo.Sfix = nullptr;
    q.fo.force = h->force;
    q.fo.Ffix = h->Ffix;
    q.s_scale_inv = 1.0 / h->s_fix_scale;
    q.gforce = h->gforce;
    // the force pass as a float64 radial table built inside the launch
    constexpr int PHI_CAP_FUSED = 32768;  // 436 A at Q_max = 25 (64 chain steps x 262 KB per element pair)
    const bool tab = h->fused_table && h->ntypes <= 2;
    if (tab && !h->phi_tab_d) {
        if ((rc = dev_alloc(&h->phi_tab_d, (size_t)IID_LF_CHAIN * h->ntypes * h->ntypes *
                                               (PHI_CAP_FUSED + 2 * FT_PAD))))
            return rc;
    }
    q.phi_tab = tab ? h->phi_tab_d : nullptr;
    q.phi_cap = PHI_CAP_FUSED;
    q.phi_stride = PHI_CAP_FUSED + 2 * FT_PAD;
    q.ntypes = (int)h->ntypes;
    q.tab_h = FT_QH / (h->qbin * (double)h->nq);
    q.ext_ref = h->ext_ref;
    // F(Q) phase through the pair histogram: as many nodes as the pair-record
    // region of the shared memory holds in two 32-bit words each
    {
        int ps_off0 = 0;
        fused_smem_bytes(h, &ps_off0, nullptr);
        q.fq_cstride = (ps_off0 / 8) & ~31;
    }
    if (h->fused_hist && !h->fused_C) {
        const size_t cnt = (size_t)(h->ntypes * (h->ntypes + 1) / 2) * q.fq_cstride;
        if ((rc = dev_alloc(&h->fused_C, cnt))) return rc;
        CU(cudaMemset(h->fused_C, 0, cnt * sizeof(unsigned long long)));
        CU(cudaStreamSynchronize(0));
    }
    q.fq_C = h->fused_hist ? h->fused_C : nullptr;
    q.n_items = (int)h->n_items_tri;
    q.lf = lf ? 1 : 0;
    q.ctl = h->zc_ctl ? h->zc_ctl : h->lf_ctl; q.slab = h->lf_slab; q.mass = h->lf_mass;
    q.pos = h->pos;
    q.pos_in = lf ? nullptr : h->zc_pos_in;
    q.force_out = h->n_restraints ? nullptr : h->zc_force_out;
    q.out_host = h->zc_out;
    q.lf_mirror = (lf && !h->n_restraints) ? h->zc_lf_mirror : nullptr;
    q.n_chain = (lf && q.lf_mirror) ? h->zc_chain : 1;
    q.chain_stride = (int)(LF_CTL + 6 * h->n + 16);
    q.n = (int)h->n; q.np = (int)h->np; q.round_f32 = 1;
    q.inv_na_d = h->inv_na_d; q.Mq = h->Mq; q.vgo = h->vgo; q.gogo = h->gogo;
    q.MF = h->MF; q.wq_blk = h->wq_blk;
    q.potential = potential; q.conv = conv; q.out4 = h->out4;
    static const bool want_stamps = getenv("IID_FUSED_STAMPS") != nullptr;
    if (want_stamps && !h->stamps) {
        CU(cudaMalloc((void **)&h->stamps, 24 * sizeof(unsigned long long)));
        CU(cudaMemset(h->stamps, 0, 24 * sizeof(unsigned long long)));
    }
    q.stamps = h->stamps;
    const int nchunk = (int)((h->nq + C32 - 1) / C32);
    const size_t smem = fused_smem_bytes(h, &q.ps_off, &q.pl);
    static bool attr_done[64] = {false};
    if (!attr_done[h->device & 63]) {
        CU(cudaFuncSetAttribute(fused_eval_kernel<true>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done[h->device & 63] = true;
    }
    void *args[] = {(void *)&q};
    CU(cudaLaunchCooperativeKernel((const void *)fused_eval_kernel<true>, dim3(h->sm_count),
                                   dim3(32 * nchunk), args, smem, h->stream));
    ++h->launches;
    if (h->stamps) {  // developer timing (block 0's view): synchronous read-back
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(h->stream, &cs);
        if (cs == cudaStreamCaptureStatusNone) {
            unsigned long long t[24];
            CU(cudaStreamSynchronize(h->stream));
            CU(cudaMemcpy(t, h->stamps, sizeof(t), cudaMemcpyDeviceToHost));
            const int ns = q.lf_mirror ? 12 : 9;
            for (int k = 1; k < ns; ++k) h->stamp_sum[k] += (double)(t[k] - t[k - 1]);
            if (++h->stamp_n % 200 == 0) {
                fprintf(stderr, "fused phases (us, mean of %lld):", (long long)h->stamp_n);
                const char *nm[12] = {"", "stage", "sync", "fq", "sync", "MF", "sync", "pot", "force",
                                      "sync", "kick", "sync"};
                for (int k = 1; k < ns; ++k)
                    fprintf(stderr, " %s %.1f", nm[k], 1e-3 * h->stamp_sum[k] / h->stamp_n);
                if (q.phi_tab) {
                    double er[8];
                    cudaMemcpy(er, h->ext_ref, sizeof(er), cudaMemcpyDeviceToHost);
                    fprintf(stderr, " | table: radius %.2f A, %d entries; from pot: build %.1f (last "
                            "block %.1f), sync -> %.1f, forces %.1f (last block %.1f)", er[4],
                            (int)er[5], 1e-3 * (double)(t[16] - t[7]), 1e-3 * (double)(t[18] - t[7]),
                            1e-3 * (double)(t[17] - t[7]), 1e-3 * (double)(t[8] - t[7]),
                            1e-3 * (double)(t[19] - t[7]));
                }
                if (q.n_chain >= 4)
                    fprintf(stderr, " | steps 0-2 of this chain: %.1f %.1f %.1f",
                            1e-3 * (double)(t[13] - t[12]), 1e-3 * (double)(t[14] - t[13]),
                            1e-3 * (double)(t[15] - t[14]));
                fprintf(stderr, "\n");
            }
        }
    }
    return 0;
}

// lf: the positions come out of the leapfrog's half kick + drift (state slab,
// h->lf_ctl) instead of h->pos.
static int enqueue_eval_device(iid_handle *h, int potential, double conv, bool want_forces,
                               bool lf = false, int phase = EVAL_ALL, bool want_pdf = false)
{
    int rc2;
    cudaStream_t st = h->stream;
    if (phase == EVAL_ALL && fused_applicable(h, want_forces, want_pdf)) {
        if ((rc2 = launch_fused(h, potential, conv, lf))) return rc2;
        for (int s = 0; s < h->n_restraints; ++s)
            if ((rc2 = launch_spring(h, h->pos, h->n, h->rs_type[s], h->rs_k[s], h->rs_rt[s],
                                     nullptr, h->out4 + 4, h->force, nullptr, h->stream)))
                return rc2;
        return 0;
    }
    // staging also clears S and the force accumulator (no memset nodes)
    if (lf) {
        const int np = (int)h->np;
        lf_stage_kernel<<<(np + 255) / 256, 256, 0, st>>>(
            h->lf_ctl, h->lf_slab, h->lf_mass, (int)h->n, h->orig, np, h->precision == IID_FP32,
            h->pos, h->x, h->y, h->z, h->valid, h->S, (int)h->qp, h->force, 3 * (int)h->n);
        ++h->launches;
        CU(cudaGetLastError());
    } else if (phase != EVAL_REST) {
        const int np = (int)h->np;
        prep_kernel<<<(np + 255) / 256, 256, 0, st>>>(h->pos, h->orig, np,
                                                      h->precision == IID_FP32, h->x, h->y, h->z,
                                                      h->valid, h->S, (int)h->qp, h->force,
                                                      want_forces ? 3 * (int)h->n : 0);
        ++h->launches;
        CU(cudaGetLastError());
    }
    if (phase != EVAL_REST &&
        (rc2 = launch_debye(h, MODE_FQ, nullptr, h->S, nullptr, nullptr, st)))
        return rc2;
    if (phase == EVAL_FQ) return 0;
    // F = 2 S / na and G = T F in one launch
    {
        const int64_t threads = h->nr * 32;
        gr_from_s_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
            h->T, h->S, h->inv_na_d, h->nr, (int)h->nq, (int)h->qp, h->F, h->Gr);
        ++h->launches;
        CU(cudaGetLastError());
    }
    // Rw / chi^2 in r space; the chain-rule weights in Q space:
    // wq = conv T^T c = conv (coef0 T^T go - coef1 (T^T T) F)
    if (!h->qspace_wq) {
        if (h->n_restraints)
            CU(cudaMemsetAsync(h->out4 + 4, 0, sizeof(double), h->stream));
        if ((rc2 = iid_potential(h, h->Gr, h->target, potential, conv, h->out4,
                                 want_forces ? h->wq : nullptr, nullptr)))
            return rc2;
    } else {
        potential_kernel<<<1, 1024, 0, h->stream>>>(h->Gr, h->target, (int)h->nr, potential,
                                                    conv, h->out4, h->cr, h->coef);
        ++h->launches;
        CU(cudaGetLastError());
        if (want_forces) {
            // a separate 11-block launch: appended to the one-block potential
            // kernel the 0.9 MB mat-vec ran at one SM's bandwidth (39 us under ncu)
            wq_from_q_kernel<<<(unsigned)((h->nq + 31) / 32), 1024, 0, h->stream>>>(
                h->Mq, h->F, h->vgo, h->coef, (int)h->nq, (int)h->qp, conv, h->wq);
            ++h->launches;
            CU(cudaGetLastError());
        }
    }
    if (!want_forces && h->n_restraints) {
        for (int s = 0; s < h->n_restraints; ++s)
            if ((rc2 = launch_spring(h, h->pos, h->n, h->rs_type[s], h->rs_k[s], h->rs_rt[s],
                                     nullptr, h->out4 + 4, nullptr, nullptr, h->stream)))
                return rc2;
    }
    if (want_forces) {
        // positions are staged and h->force is cleared: enqueue the force pass
        if ((rc2 = launch_force(h, h->wq, h->force, h->stream))) return rc2;
        if (h->n_restraints) {
            // out4[4] was zeroed by potential_kernel (Q-space path) / the memset above
            for (int s = 0; s < h->n_restraints; ++s)
                if ((rc2 = launch_spring(h, h->pos, h->n, h->rs_type[s], h->rs_k[s],
                                         h->rs_rt[s], nullptr, h->out4 + 4, h->force, nullptr,
                                         h->stream)))
                    return rc2;
        }
    }
    return 0;
}

// Upload a new target PDF (nullptr = the resident one) and keep T^T target
// current (Q-space chain-rule weights, once per target).
static int refresh_target(iid_handle *h, const double *target_host, double *pinned)
{
    if (target_host) {
        memcpy(pinned, target_host, h->nr * sizeof(double));
        CU(cudaMemcpyAsync(h->target, pinned, h->nr * sizeof(double), cudaMemcpyHostToDevice,
                           h->stream));
        h->vgo_valid = false;
        double c = 0.0;
        for (int64_t r = 0; r < h->nr; ++r) c = fma(target_host[r], target_host[r], c);
        h->gogo = c;
    }
    if (!h->vgo_valid) {
        CU(cudaMemsetAsync(h->vgo, 0, h->qp * sizeof(double), h->stream));
        wq_kernel<<<(unsigned)((h->nr + WQ_ROWS - 1) / WQ_ROWS), 352, 0, h->stream>>>(
            h->T, h->target, (int)h->nr, (int)h->nq, (int)h->qp, 1.0, h->vgo);
        ++h->launches;
        CU(cudaGetLastError());
        h->vgo_valid = true;
    }
    return 0;
}

// Run `enqueue` on the handle's stream: eagerly the first two times, then
// captured into a graph (key: potential, conv, flag) and replayed.  At a few
// hundred atoms the evaluation is bound by launch latency, not by arithmetic.
template <class F>
static int run_graphed(iid_handle *h, GraphSlot &g, bool graphable, int potential, double conv,
                       bool flag, F &&enqueue)
{
    int rc;
    if (graphable && g.exec &&
        (g.key_pot != potential || g.key_conv != conv || g.key_flag != flag))
        g.drop();
    if (graphable && g.exec) {
        CU(cudaGraphLaunch(g.exec, h->stream));
        h->launches += g.launches;
    } else if (graphable && g.warm >= 2) {
        cudaGraph_t graph = nullptr;
        const int64_t launches0 = h->launches;
        CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue();
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        g.launches = h->launches - launches0;  // counted while capturing, not yet run
        h->launches = launches0;
        if (rc == 0 && ce == cudaSuccess && graph &&
            cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
            g.key_pot = potential;
            g.key_conv = conv;
            g.key_flag = flag;
            cudaGraphDestroy(graph);
            CU(cudaGraphLaunch(g.exec, h->stream));
            h->launches += g.launches;
        } else {
            // capture refused: fall back to plain launches for good
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            g.exec = nullptr;
            h->use_graph = false;
            if ((rc = enqueue())) return rc;
        }
    } else {
        if ((rc = enqueue())) return rc;
        ++g.warm;
    }
    return 0;
}

extern "C" int iid_energy_forces_host(iid_handle *h, const double *pos_host,
                                      const double *target_host, int potential, double conv,
                                      double *out_host, double *forces_host, double *pdf_host)
{
    if (MULTI(h))
        return multi_energy_forces_host(h, pos_host, target_host, potential, conv, out_host,
                                        forces_host, pdf_host);
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (h->nr == 0) return fail(IID_E_NOTRANSFORM, "call iid_set_transform first");
    if (!pos_host || !out_host) return fail(IID_E_BADARG, "null pointer");
    if (h->world != 1)
        return fail(IID_E_BADARG, "iid_energy_forces_host needs the whole pair list (world == 1)");
    int rc;
    double *pfor = h->pin + 3 * h->n;
    double *pf = h->pin + 6 * h->n;
    double *po = pf + h->qp;
    double *pg = po + 8;
    double *pt = pg + h->nr;
    if ((rc = refresh_target(h, target_host, pt))) return rc;
    // The whole sequence (H2D, 7 kernels, 3 memsets, D2H) is replayed from a
    // CUDA graph once it has run twice with the same shape: at a few hundred
    // atoms the evaluation is bound by launch latency, not by arithmetic.
    const bool graphable = h->use_graph && !h->timing && forces_host != nullptr;
    memcpy(h->pin, pos_host, (size_t)3 * h->n * sizeof(double));
    // small structures: the fused kernel reads the positions from, and writes
    // forces and scalars to, the pinned staging directly -- no copy nodes
    const bool zc = h->zero_copy_small && fused_applicable(h, forces_host != nullptr,
                                                           pdf_host != nullptr) &&
                    h->n_restraints == 0;
    auto enqueue = [&]() -> int {
        int rc2;
        h->zc_pos_in = zc ? h->pin : nullptr;
        h->zc_force_out = zc ? pfor : nullptr;
        h->zc_out = zc ? po : nullptr;
        if (!zc)
            CU(cudaMemcpyAsync(h->pos, h->pin, (size_t)3 * h->n * sizeof(double),
                               cudaMemcpyHostToDevice, h->stream));
        rc2 = enqueue_eval_device(h, potential, conv, forces_host != nullptr, false, EVAL_ALL,
                                  pdf_host != nullptr);
        h->zc_pos_in = nullptr;
        h->zc_force_out = h->zc_out = nullptr;
        if (rc2) return rc2;
        if (zc) return 0;
        if (forces_host) {
            CU(cudaMemcpyAsync(pfor, h->force, (size_t)3 * h->n * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
        }
        CU(cudaMemcpyAsync(po, h->out4, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (pdf_host)
            CU(cudaMemcpyAsync(pg, h->Gr, h->nr * sizeof(double), cudaMemcpyDeviceToHost,
                               h->stream));
        return 0;
    };
    // (a one-node graph still beats a direct cooperative launch: 63 against 73 us
    // per call at Au561)
    if ((rc = run_graphed(h, h->ef, graphable, potential, conv, pdf_host != nullptr, enqueue)))
        return rc;
    CU(cudaStreamSynchronize(h->stream));
    memcpy(out_host, po, 4 * sizeof(double));
    if (forces_host) memcpy(forces_host, pfor, (size_t)3 * h->n * sizeof(double));
    if (pdf_host) memcpy(pdf_host, pg, h->nr * sizeof(double));
    return 0;
}

// --- device-resident sampler states ------------------------------------------
// pyiid/sim/__init__.py:10-38 (leapfrog) keeps an Atoms object per phase-space
// point; here a point is a slot (q, p, f) of a device slab and one leapfrog is
// one graph replay: kick + drift, the fused energy + forces sequence, kick +
// kinetic energy + centring, one device-to-host copy of (q, p, scalars).
extern "C" int iid_sampler_setup(iid_handle *h, int64_t n_slots, const double *masses_host,
                                 const double *cell_centre)
{
    if (MULTI(h)) {
        if (h->active != 1)
            return fail(IID_E_BADARG, "device-resident sampler states need a one-device structure");
        return iid_sampler_setup(h->subs[0], n_slots, masses_host, cell_centre);
    }
    NEED(h);
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (n_slots < 2 || n_slots > 4096 || !masses_host || !cell_centre)
        return fail(IID_E_BADARG, "need 2 <= n_slots <= 4096, masses and the cell centre");
    for (int64_t i = 0; i < h->n; ++i)
        if (!(masses_host[i] > 0.0)) return fail(IID_E_BADARG, "masses must be positive");
    CU(cudaStreamSynchronize(h->stream));
    for (GraphSlot &g : h->lf) g.drop();
    for (GraphSlot &g : h->lf_chain) g.drop();
    static_assert(LF_CHAIN_MAX == IID_LF_CHAIN, "ring of the pinned staging");
    h->lf_mass_h.resize(h->n);
    for (int64_t i = 0; i < h->n; ++i) h->lf_mass_h[i] = 1.0 / masses_host[i];
    const size_t n3 = (size_t)3 * h->n;
    int rc;
    if (h->lf_slab) { cudaFree(h->lf_slab); h->lf_slab = nullptr; }
    if ((rc = dev_alloc(&h->lf_slab, (size_t)n_slots * 3 * n3)) ||
        (rc = dev_alloc(&h->lf_mass, h->n)) || (rc = dev_alloc(&h->lf_ctl, LF_CTL)) ||
        (rc = dev_alloc(&h->lf_mirror, 2 * n3 + 16)))
        return rc;
    h->lf_slots = n_slots;
    // a ring of IID_LF_CHAIN staging slots: step parameters in, mirror of the new state out
    const size_t need = (size_t)IID_LF_CHAIN * (LF_CTL + 2 * n3 + 16);
    if (need > h->lf_pin_count) {
        if (h->lf_pin) cudaFreeHost(h->lf_pin);
        h->lf_pin = nullptr;
        CU(cudaMallocHost((void **)&h->lf_pin, need * sizeof(double)));
        h->lf_pin_count = need;
    }
    CU(cudaMemcpy(h->lf_mass, masses_host, h->n * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemset(h->lf_slab, 0, (size_t)n_slots * 3 * n3 * sizeof(double)));
    CU(cudaStreamSynchronize(0));
    for (int r = 0; r < IID_LF_CHAIN; ++r)
        for (int w = 0; w < 3; ++w) h->lf_pin[(size_t)r * (LF_CTL + 2 * n3 + 16) + 4 + w] = cell_centre[w];
    h->lf_system = true;
    return 0;
}

static int check_slot(iid_handle *h, int slot)
{
    if (!h->lf_system) return fail(IID_E_BADARG, "call iid_sampler_setup first");
    if (slot < 0 || slot >= h->lf_slots) return fail(IID_E_BADARG, "state slot out of range");
    return 0;
}

extern "C" int iid_state_upload(iid_handle *h, int slot, const double *q_host,
                                const double *p_host, const double *f_host)
{
    if (MULTI(h)) return iid_state_upload(h->subs[0], slot, q_host, p_host, f_host);
    NEED(h);
    int rc = check_slot(h, slot);
    if (rc) return rc;
    if (!q_host || !p_host || !f_host) return fail(IID_E_BADARG, "null pointer");
    const size_t n3 = (size_t)3 * h->n, bytes = n3 * sizeof(double);
    double *base = h->lf_slab + (size_t)slot * 3 * n3;
    // ordered after the steps already enqueued; pageable source: synchronous
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(base, q_host, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(base + n3, p_host, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(base + 2 * n3, f_host, bytes, cudaMemcpyHostToDevice));
    CU(cudaStreamSynchronize(0));  // pageable copies: the DMA may still be in flight
    return 0;
}

extern "C" int iid_state_download(iid_handle *h, int slot, double *q_host, double *p_host,
                                  double *f_host)
{
    if (MULTI(h)) return iid_state_download(h->subs[0], slot, q_host, p_host, f_host);
    NEED(h);
    int rc = check_slot(h, slot);
    if (rc) return rc;
    const size_t n3 = (size_t)3 * h->n, bytes = n3 * sizeof(double);
    const double *base = h->lf_slab + (size_t)slot * 3 * n3;
    CU(cudaStreamSynchronize(h->stream));
    if (q_host) CU(cudaMemcpy(q_host, base, bytes, cudaMemcpyDeviceToHost));
    if (p_host) CU(cudaMemcpy(p_host, base + n3, bytes, cudaMemcpyDeviceToHost));
    if (f_host) CU(cudaMemcpy(f_host, base + 2 * n3, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

// Ring slot `ring` of the pinned staging: the step's parameters, then the mirror
// of the state it produces.
static inline size_t lf_ring_stride(const iid_handle *h) { return LF_CTL + 6 * (size_t)h->n + 16; }

static void leapfrog_fill(iid_handle *h, int ring, int src, int dst, double step, int centre)
{
    double *ctl = h->lf_pin + (size_t)ring * lf_ring_stride(h);
    ctl[0] = step;
    ctl[1] = (double)src;
    ctl[2] = (double)dst;
    ctl[3] = centre ? 1.0 : 0.0;
}

// The fused kernel can walk a chain of steps inside ONE cooperative launch.
static bool leapfrog_in_kernel(const iid_handle *h)
{
    return h->zero_copy_small && !h->n_restraints && h->chain_in_kernel &&
           fused_applicable(h, true, false);
}

// Enqueue n_chain leapfrog steps whose parameters / mirrors live in ring slots
// ring .. ring + n_chain - 1 of the pinned staging (no synchronisation).
// n_chain > 1 only with leapfrog_in_kernel().
static int leapfrog_enqueue(iid_handle *h, int ring, int n_chain, int potential, double conv)
{
    const size_t n3 = (size_t)3 * h->n;
    double *ctl = h->lf_pin + (size_t)ring * lf_ring_stride(h), *mir = ctl + LF_CTL;
    const int n = (int)h->n;
    // small structures: the kernels read the step parameters from, and write the
    // new state's mirror to, the pinned staging directly -- no copy nodes
    const bool zc = h->zero_copy_small && fused_applicable(h, true, false);
    auto enqueue = [&]() -> int {
        int rc2;
        if (!zc)
            CU(cudaMemcpyAsync(h->lf_ctl, ctl, LF_CTL * sizeof(double), cudaMemcpyHostToDevice,
                               h->stream));
        h->zc_ctl = zc ? ctl : nullptr;
        h->zc_lf_mirror = zc ? mir : nullptr;  // no restraints: the step finishes in this launch
        h->zc_chain = n_chain;
        rc2 = enqueue_eval_device(h, potential, conv, true, true);
        h->zc_ctl = nullptr;
        h->zc_lf_mirror = nullptr;
        h->zc_chain = 1;
        if (rc2) return rc2;
        if (zc && !h->n_restraints) return 0;
        lf_finish_kernel<<<1, 1024, 0, h->stream>>>(zc ? ctl : h->lf_ctl, h->lf_slab, h->lf_mass, n,
                                                    h->pos, h->force, zc ? mir : h->lf_mirror,
                                                    h->out4);
        ++h->launches;
        CU(cudaGetLastError());
        if (!zc)
            CU(cudaMemcpyAsync(mir, h->lf_mirror, (2 * n3 + 16) * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
        return 0;
    };
    const bool graphable = h->use_graph && !h->timing;
    // one graph per (ring slot, 1 step) and one per (ring slot 0, chain length)
    GraphSlot &g = n_chain > 1 ? h->lf_chain[n_chain - 1] : h->lf[ring];
    return run_graphed(h, g, graphable, potential, conv, false, enqueue);
}

static void leapfrog_collect(iid_handle *h, int ring, double *out_host, double *q_host,
                             double *p_host)
{
    const size_t n3 = (size_t)3 * h->n;
    const double *mir = h->lf_pin + (size_t)ring * lf_ring_stride(h) + LF_CTL;
    const double *po = mir + 2 * n3;
    // out: energy, scale, -, -, restraint energy, kinetic energy, shift x y z
    for (int k = 0; k < 9; ++k) out_host[k] = po[k];
    if (!h->n_restraints) out_host[4] = 0.0;
    if (h->zero_copy_small && !h->n_restraints && fused_applicable(h, true, false)) {
        // the fused launch finishes the step spread over its grid and leaves the
        // kinetic energy sum p.p/m / 2 to the host (fixed order: reproducible)
        const double *p = mir + n3, *im = h->lf_mass_h.data();  // (inverse masses)
        double ke[4] = {0.0, 0.0, 0.0, 0.0};
        for (int64_t a = 0; a < h->n; ++a) {
            const double px = p[3 * a], py = p[3 * a + 1], pz = p[3 * a + 2];
            ke[a & 3] = fma(fma(px, px, fma(py, py, pz * pz)), im[a], ke[a & 3]);
        }
        out_host[5] = 0.5 * ((ke[0] + ke[1]) + (ke[2] + ke[3]));
    }
    if (q_host) memcpy(q_host, mir, n3 * sizeof(double));
    if (p_host) memcpy(p_host, mir + n3, n3 * sizeof(double));
}

extern "C" int iid_leapfrog_host(iid_handle *h, int src, int dst, double step, int centre,
                                 const double *target_host, int potential, double conv,
                                 double *out_host, double *q_host, double *p_host)
{
    return iid_leapfrog_chain_host(h, src, &dst, 1, step, centre, target_host, potential, conv,
                                   out_host, q_host, p_host);
}

// n_steps consecutive leapfrog steps src -> dst[0] -> dst[1] -> ... (a NUTS
// subtree of depth j is 2^j such steps in a row, pyiid/sim/nuts_hmc.py:15-88).
// Small structures walk the whole chain inside ONE cooperative launch; the
// launch latency, the wake-up of the host and its per-call work are paid once
// per chain instead of once per step.
//
// _begin enqueues the chain and returns; _next hands out the steps in order as
// they complete: the fused launch raises a flag in the pinned staging after
// each step, so the host works on step i (kinetic energy, the caller's tree
// logic) while the device computes step i + 1.  Any other call on the handle
// first waits for the chain (NEED -> chain_drain).
extern "C" int iid_leapfrog_chain_begin(iid_handle *h, int src, const int *dst, int n_steps,
                                        double step, int centre, const double *target_host,
                                        int potential, double conv, int64_t *chain_id)
{
    if (MULTI(h)) {
        if (h->active != 1)
            return fail(IID_E_BADARG, "device-resident sampler states need a one-device structure");
        return iid_leapfrog_chain_begin(h->subs[0], src, dst, n_steps, step, centre, target_host,
                                        potential, conv, chain_id);
    }
    NEED(h);
    if (h->nr == 0) return fail(IID_E_NOTRANSFORM, "call iid_set_transform first");
    if (h->world != 1)
        return fail(IID_E_BADARG, "iid_leapfrog_host needs the whole pair list (world == 1)");
    if (!dst || n_steps < 1 || n_steps > IID_LF_CHAIN)
        return fail(IID_E_BADARG, "need 1 <= n_steps <= IID_LF_CHAIN and dst");
    int rc;
    if ((rc = check_slot(h, src))) return rc;
    for (int i = 0; i < n_steps; ++i) {
        if ((rc = check_slot(h, dst[i]))) return rc;
        if (dst[i] == (i ? dst[i - 1] : src))
            return fail(IID_E_BADARG, "source and destination slot must differ");
    }
    double *pt = h->pin + 6 * h->n + h->qp + 8 + h->nr;  // target staging of the handle
    if ((rc = refresh_target(h, target_host, pt))) return rc;
    for (int i = 0; i < n_steps; ++i) {
        leapfrog_fill(h, i, i ? dst[i - 1] : src, dst[i], step, centre);
        h->lf_pin[(size_t)i * lf_ring_stride(h) + 7] = h->lf_seq + (double)(i + 1);
    }
    h->chain_flags = n_steps > 1 && leapfrog_in_kernel(h);
    if (h->chain_flags) {
        // the whole chain in one cooperative launch
        if ((rc = leapfrog_enqueue(h, 0, n_steps, potential, conv))) return rc;
    } else {
        for (int i = 0; i < n_steps; ++i)
            if ((rc = leapfrog_enqueue(h, i, 1, potential, conv))) return rc;
    }
    h->chain_n = h->chain_left = n_steps;
    h->chain_synced = false;
    ++h->chain_id;
    if (chain_id) *chain_id = h->chain_id;
    return 0;
}

static void chain_drain(iid_handle *h)
{
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->lf_seq += (double)h->chain_n;
    h->chain_left = h->chain_n = 0;
}

extern "C" int iid_leapfrog_chain_next(iid_handle *h, int64_t chain_id, double *out_host,
                                       double *q_host, double *p_host)
{
    if (MULTI(h)) return iid_leapfrog_chain_next(h->subs[0], chain_id, out_host, q_host, p_host);
    if (!h || !out_host) return fail(IID_E_BADARG, "null pointer");
    if (h->chain_left <= 0 || chain_id != h->chain_id)
        return fail(IID_E_NOCHAIN, "this leapfrog chain is not in flight (any more)");
    CU(cudaSetDevice(h->device));
    static const bool stats = getenv("IID_CHAIN_STATS") != nullptr;  // developer timing
    const auto t_in = std::chrono::steady_clock::now();
    const int i = h->chain_n - h->chain_left;
    if (h->chain_flags && i + 1 < h->chain_n) {
        // step i of a chain inside one launch: its flag follows its mirror
        const volatile double *flag =
            h->lf_pin + (size_t)i * lf_ring_stride(h) + LF_CTL + 6 * (size_t)h->n + 15;
        const double want = h->lf_seq + (double)(i + 1);
        for (unsigned spin = 0; *flag != want; ++spin) {
            if ((spin & 0xfff) == 0xfff) {
                // the launch ended without raising the flag: an error, or nothing to wait for
                const cudaError_t e = cudaStreamQuery(h->stream);
                if (e == cudaSuccess) {
                    if (*flag == want) break;
                    chain_drain(h);
                    g_err = "leapfrog chain ended without completing its steps";
                    return (int)cudaErrorUnknown;
                }
                if (e != cudaErrorNotReady) {
                    chain_drain(h);
                    g_err = std::string("leapfrog chain: ") + cudaGetErrorString(e);
                    return (int)e;
                }
            }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    } else if (!h->chain_synced) {
        const cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) {
            chain_drain(h);
            g_err = std::string("leapfrog chain: ") + cudaGetErrorString(e);
            return (int)e;
        }
        h->chain_synced = true;
    }
    const auto t_mid = std::chrono::steady_clock::now();
    leapfrog_collect(h, i, out_host, q_host, p_host);
    if (stats) {
        static double wait_us = 0.0, collect_us = 0.0, last_wait = 0.0;
        static long calls = 0, lasts = 0;
        const auto t_out = std::chrono::steady_clock::now();
        const double w = std::chrono::duration<double, std::micro>(t_mid - t_in).count();
        if (i + 1 < h->chain_n) wait_us += w;
        else { last_wait += w; ++lasts; }
        collect_us += std::chrono::duration<double, std::micro>(t_out - t_mid).count();
        if (++calls % 1000 == 0)
            fprintf(stderr, "chain_next x%ld: wait %.1f us (inner steps), %.1f us (last step of a "
                    "chain, %ld of them), collect %.1f us\n", calls,
                    wait_us / std::max(1L, calls - lasts), last_wait / std::max(1L, lasts), lasts,
                    collect_us / calls);
    }
    if (--h->chain_left == 0) {
        h->lf_seq += (double)h->chain_n;
        h->chain_n = 0;
    }
    return 0;
}

extern "C" int iid_leapfrog_chain_host(iid_handle *h, int src, const int *dst, int n_steps,
                                       double step, int centre, const double *target_host,
                                       int potential, double conv, double *out_host,
                                       double *q_host, double *p_host)
{
    if (!out_host) return fail(IID_E_BADARG, "null pointer");
    int64_t id = 0;
    int rc = iid_leapfrog_chain_begin(h, src, dst, n_steps, step, centre, target_host, potential,
                                      conv, &id);
    if (rc) return rc;
    iid_handle *hh = MULTI(h) ? h->subs[0] : h;
    const size_t n3 = (size_t)3 * hh->n;
    for (int i = 0; i < n_steps; ++i)
        if ((rc = iid_leapfrog_chain_next(h, id, out_host + 9 * (size_t)i,
                                          q_host ? q_host + n3 * i : nullptr,
                                          p_host ? p_host + n3 * i : nullptr)))
            return rc;
    return 0;
}

// Rw / chi^2 of two host vectors and the chain-rule vector c (calc/__init__.py
// wrap_rw / wrap_chi_sq :10-54 and the c of wrap_grad_* :56-105).
extern "C" int iid_rw_host(iid_handle *h, const double *gcalc_host, const double *gobs_host,
                           int64_t len, int potential, double conv, double *out_host,
                           double *c_host)
{
    if (MULTI(h))
        return iid_rw_host(h->subs[0], gcalc_host, gobs_host, len, potential, conv, out_host, c_host);
    NEED(h);
    if (!gcalc_host || !gobs_host || !out_host || len < 1)
        return fail(IID_E_BADARG, "bad argument");
    if (potential != IID_POT_RW && potential != IID_POT_CHI_SQ)
        return fail(IID_E_BADARG, "unknown potential");
    double *buf = nullptr;
    CU(cudaMalloc((void **)&buf, (3 * len + 4) * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(buf, gcalc_host, len * sizeof(double),
                                    cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(buf + len, gobs_host, len * sizeof(double),
                            cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        potential_kernel<<<1, 1024, 0, h->stream>>>(buf, buf + len, (int)len, potential, conv,
                                                    buf + 3 * len, buf + 2 * len, nullptr);
        ++h->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(out_host, buf + 3 * len, 4 * sizeof(double),
                            cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && c_host)
        e = cudaMemcpyAsync(c_host, buf + 2 * len, len * sizeof(double),
                            cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(buf);
    if (e != cudaSuccess) {
        g_err = std::string("iid_rw_host: ") + cudaGetErrorString(e);
        return (int)e;
    }
    return 0;
}

// out[row] = sum_k A[row][k] c[k]: the contraction of wrap_grad_rw /
// wrap_grad_chi_sq (master_kernel.py:334-347, 367-375) over a host array.
extern "C" int iid_contract_host(iid_handle *h, const void *A_host, int a_is_f32,
                                 int64_t rows, int64_t len, const double *c_host,
                                 double *out_host)
{
    if (MULTI(h))
        return iid_contract_host(h->subs[0], A_host, a_is_f32, rows, len, c_host, out_host);
    NEED(h);
    if (!A_host || !c_host || !out_host || rows < 0 || len < 1)
        return fail(IID_E_BADARG, "bad argument");
    if (rows == 0) return 0;
    const size_t esz = a_is_f32 ? sizeof(float) : sizeof(double);
    void *A = nullptr;
    double *cv = nullptr;
    CU(cudaMalloc(&A, (size_t)rows * len * esz));
    cudaError_t e = cudaMalloc((void **)&cv, (len + rows) * sizeof(double));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(A, A_host, (size_t)rows * len * esz, cudaMemcpyHostToDevice,
                            h->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(cv, c_host, len * sizeof(double), cudaMemcpyHostToDevice,
                            h->stream);
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
        if (a_is_f32)
            gr_kernel<float><<<blocks, 256, 0, h->stream>>>((const float *)A, cv, rows, (int)len,
                                                            (int)len, cv + len);
        else
            gr_kernel<double><<<blocks, 256, 0, h->stream>>>((const double *)A, cv, rows,
                                                             (int)len, (int)len, cv + len);
        ++h->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(out_host, cv + len, rows * sizeof(double), cudaMemcpyDeviceToHost,
                            h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(A);
    if (cv) cudaFree(cv);
    if (e != cudaSuccess) {
        g_err = std::string("iid_contract_host: ") + cudaGetErrorString(e);
        return (int)e;
    }
    return 0;
}

// G(r) from a host F(Q) (used when noise is added to F(Q) on the host,
// elasticscatter/__init__.py:371-390)
extern "C" int iid_fq_to_gr_host(iid_handle *h, const double *F_host, double *pdf_host)
{
    if (MULTI(h)) return iid_fq_to_gr_host(h->subs[0], F_host, pdf_host);
    NEED(h);
    if (h->nr == 0) return fail(IID_E_NOTRANSFORM, "call iid_set_transform first");
    if (!F_host || !pdf_host) return fail(IID_E_BADARG, "null pointer");
    double *pf = h->pin + 6 * h->n;
    double *pg = pf + h->qp + 8;
    memcpy(pf, F_host, h->nq * sizeof(double));
    CU(cudaMemcpyAsync(h->F, pf, h->nq * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int rc = iid_fq_to_gr(h, h->F, h->Gr, nullptr);
    if (rc) return rc;
    CU(cudaMemcpyAsync(pg, h->Gr, h->nr * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(pdf_host, pg, h->nr * sizeof(double));
    return 0;
}

// Tunables (also read from IID_* environment variables at iid_create).
extern "C" int iid_set_option(iid_handle *h, const char *key, int64_t value)
{
    if (MULTI(h)) {
        if (!key) return fail(IID_E_BADARG, "null argument");
        for (iid_handle *sh : h->subs) {
            int rc = iid_set_option(sh, key, value);
            if (rc) return rc;
        }
        h->multi_options.push_back({std::string(key), value});
        return 0;
    }
    if (!h || !key) return fail(IID_E_BADARG, "null argument");
    const std::string k(key);
    if (k == "force_table") h->use_force_table = value != 0;
    else if (k == "force_table_min_n") h->force_table_min_n = value;
    else if (k == "graph") h->use_graph = value != 0;
    else if (k == "cheb") h->cheb = value != 0;
    else if (k == "grad_split") h->grad_split = value != 0;
    else if (k == "prod_unroll") h->prod_unroll = value != 0;
    else if (k == "grad_nw_max") h->grad_nw_max = (int)std::max<int64_t>(1, std::min<int64_t>(12, value));
    else if (k == "qspace_wq") h->qspace_wq = value != 0;
    else if (k == "nw_max") h->nw_max = (int)std::max<int64_t>(1, std::min<int64_t>(12, value));
    else if (k == "zero_copy") h->zero_copy = value != 0;
    else if (k == "fused") h->use_fused = value != 0;
    else if (k == "fused_det") h->fused_det = value != 0;
    else if (k == "chain_in_kernel") h->chain_in_kernel = value != 0;
    else if (k == "fused_table") h->fused_table = value != 0;
    else if (k == "fused_hist") h->fused_hist = value != 0;
    else if (k == "fq_hist") h->fq_hist = value != 0;
    else if (k == "fq_hist_min_n") h->fq_hist_min_n = std::max<int64_t>(2, value);
    else if (k == "det_fq") h->det_fq = value != 0;
    else if (k == "acc_j") h->acc_j = (int)std::max<int64_t>(0, value);
    else if (k == "piece_div") {
        h->piece_div = (int)std::max<int64_t>(1, std::min<int64_t>(1024, value));
        if (h->np > 0) {
            CU(cudaSetDevice(h->device));
            CU(cudaStreamSynchronize(h->stream));
            int rc = upload_row_jobs(h);
            if (rc) return rc;
        }
    }
    else return fail(IID_E_BADARG, "unknown option: " + k);
    drop_graph(h);
    return 0;
}

// --- instrumentation -----------------------------------------------------------
extern "C" int iid_launch_count(iid_handle *h, int64_t *count)
{
    if (MULTI(h)) {
        if (!count) return fail(IID_E_BADARG, "null argument");
        *count = 0;
        for (iid_handle *sh : h->subs) *count += sh->launches;
        return 0;
    }
    if (!h || !count) return fail(IID_E_BADARG, "null argument");
    *count = h->launches;
    return 0;
}

extern "C" int iid_set_timing(iid_handle *h, int enabled)
{
    if (MULTI(h)) {
        h->timing = enabled != 0;
        for (iid_handle *sh : h->subs) sh->timing = enabled != 0;
        return 0;
    }
    if (!h) return fail(IID_E_BADARG, "null handle");
    h->timing = enabled != 0;
    return 0;
}

extern "C" int iid_last_kernel_ms(iid_handle *h, float *ms, double *pairq)
{
    if (MULTI(h)) {
        if (!ms) return fail(IID_E_BADARG, "null argument");
        *ms = 0.f;
        double pq = 0.0;
        for (int d = 0; d < h->active; ++d) {  // the slowest device, all devices' pair*Q
            float m = 0.f;
            double q = 0.0;
            int rc = iid_last_kernel_ms(h->subs[d], &m, &q);
            if (rc) return rc;
            *ms = std::max(*ms, m);
            pq += q;
        }
        if (pairq) *pairq = pq;
        return 0;
    }
    NEED(h);
    if (!ms) return fail(IID_E_BADARG, "null argument");
    *ms = 0.f;
    if (h->ev_pending) {
        CU(cudaEventSynchronize(h->ev1));
        CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    }
    if (pairq) *pairq = h->last_pairq;
    return 0;
}

// Measured issue rates of the pipes that bound the pair sums, in lane-FMAs per
// second: out[0] scalar FFMA, out[1] packed FFMA2 (two lanes per thread
// instruction), out[2] DFMA.  Best of three timed launches each, CUDA events on
// the handle's stream; ~30 ms in all.
extern "C" int iid_measure_peaks(iid_handle *h, double *out)
{
    if (MULTI(h)) return iid_measure_peaks(h->subs[0], out);
    NEED(h);
    if (!out) return fail(IID_E_BADARG, "null pointer");
    float *buf = nullptr;
    CU(cudaMalloc((void **)&buf, 1024 * sizeof(float)));
    std::vector<float> init(64);
    for (int i = 0; i < 32; ++i) {
        init[i] = 0.999f + 1e-5f * i;   // |a| < 1: the chains stay finite
        init[32 + i] = 1e-3f * (i + 1);
    }
    CU(cudaMemcpy(buf, init.data(), 64 * sizeof(float), cudaMemcpyHostToDevice));
    const int blocks = h->sm_count * 4, threads = 256;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    for (int kind = 0; kind < 3; ++kind) {
        const int trips = kind == UB_DFMA ? 512 : 2048;
        double best = 0.0;
        for (int rep = 0; rep < 4; ++rep) {
            CU(cudaEventRecord(e0, h->stream));
            if (kind == UB_FFMA)
                pipe_peak_kernel<UB_FFMA><<<blocks, threads, 0, h->stream>>>(buf, trips, buf + 64);
            else if (kind == UB_FFMA2)
                pipe_peak_kernel<UB_FFMA2><<<blocks, threads, 0, h->stream>>>(buf, trips, buf + 64);
            else
                pipe_peak_kernel<UB_DFMA><<<blocks, threads, 0, h->stream>>>(buf, trips, buf + 64);
            CU(cudaEventRecord(e1, h->stream));
            CU(cudaEventSynchronize(e1));
            ++h->launches;
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            const double fmas = (double)blocks * threads * trips * UB_UNROLL * UB_CHAINS *
                                (kind == UB_FFMA2 ? 2.0 : 1.0);
            if (rep > 0 && ms > 0.f) best = std::max(best, fmas / (ms * 1e-3));
        }
        out[kind] = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return 0;
}

// --- multi-device handle ---------------------------------------------------------
// One process, every GPU of the box (reference: the thread-per-GPU task farm of
// gpu_wrappers/gpu_wrap.py:119-156, 287-314).  The handle owns one complete
// sub-handle per device; sub d computes shard (d, active) of each pass.  The
// only cross-device data are the F(Q) pair sums (nq doubles), the force array
// (3n doubles) and four scalars; they are summed on the host in device order
// (deterministic), so no collective library is involved.  The gradient rows
// are disjoint per device and go straight into the caller's array.
static int multi_need_subs(iid_handle *h, int count)
{
    while ((int)h->subs.size() < count) {
        iid_handle *sh = nullptr;
        int rc = iid_create((int)h->subs.size(), h->precision, &sh);
        if (rc) return rc;
        // options set on the multi handle so far apply to late joiners as well
        for (const auto &kv : h->multi_options) iid_set_option(sh, kv.first.c_str(), kv.second);
        if (h->n_restraints)
            iid_set_restraints(sh, h->n_restraints, h->rs_type, h->multi_rs_k, h->rs_rt);
        sh->timing = h->timing;
        h->subs.push_back(sh);
    }
    return 0;
}

extern "C" int iid_create_multi(int n_devices, int precision, iid_handle **out)
{
    if (!out) return fail(IID_E_BADARG, "null out");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(IID_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    if (n_devices <= 0) n_devices = count;
    if (n_devices > count) return fail(IID_E_BADARG, "more devices requested than present");
    iid_handle *h = new iid_handle();
    h->precision = precision;
    h->n_devices = n_devices;
    // sub-handles (one CUDA context each) are created when a structure first
    // needs them: a small structure never touches devices 1 .. n-1
    int rc = multi_need_subs(h, 1);
    if (rc) {
        delete h;
        return rc;
    }
    h->device = 0;
    h->sm_count = h->subs[0]->sm_count;
    *out = h;
    return 0;
}

extern "C" int iid_handle_devices(iid_handle *h, int *n_devices, int *active)
{
    if (!h) return fail(IID_E_BADARG, "null handle");
    if (n_devices) *n_devices = h->subs.empty() ? 1 : h->n_devices;
    if (active) *active = h->subs.empty() ? 1 : h->active;
    return 0;
}

static int multi_destroy(iid_handle *h)
{
    for (iid_handle *sh : h->subs) iid_destroy(sh);
    if (h->pinG) {
        cudaSetDevice(0);
        cudaFreeHost(h->pinG);
    }
    delete h;
    return 0;
}

static int multi_set_structure(iid_handle *h, int64_t n, const int32_t *type_index,
                               int64_t n_types, const double *ftable, const double *norm_table,
                               int64_t nq, double qbin)
{
    // small structures do not amortise the per-device launches and the host
    // hops: about 1500 atoms per device at least
    int active = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->n_devices, n / 1500));
    if (const char *s = getenv("IID_MULTI_ACTIVE"))
        active = std::max(1, std::min(h->n_devices, atoi(s)));
    {
        int rc = multi_need_subs(h, active);
        if (rc) return rc;
    }
    h->active = active;
    for (int d = 0; d < active; ++d) {
        iid_handle *sh = h->subs[d];
        sh->rank = d;  // before the structure: the row jobs are built once
        sh->world = active;
        int rc = iid_set_structure_norm(sh, n, type_index, n_types, ftable, norm_table, nq, qbin);
        if (rc) return rc;
    }
    if (nq != h->nq) h->nr = 0;
    // a device that joins with this structure gets the transform the others hold
    for (int d = 0; d < active && h->nr > 0; ++d)
        if (h->subs[d]->nr != h->nr) {
            int rc = iid_set_transform(h->subs[d], h->nr, nq, h->T_host.data());
            if (rc) return rc;
        }
    h->n = n;
    h->nq = nq;
    h->qp = h->subs[0]->qp;
    h->np = h->subs[0]->np;
    h->inv_na_host = h->subs[0]->inv_na_host;
    return 0;
}

static int multi_set_transform(iid_handle *h, int64_t nr, int64_t nq, const double *T)
{
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (nr < 1 || !T || nq != h->nq) return fail(IID_E_BADARG, "bad transform arguments");
    for (int d = 0; d < h->active; ++d) {
        int rc = iid_set_transform(h->subs[d], nr, nq, T);
        if (rc) return rc;
    }
    for (size_t d = (size_t)h->active; d < h->subs.size(); ++d) h->subs[d]->nr = 0;  // stale
    h->T_host.assign(T, T + (size_t)nr * nq);
    h->nr = nr;
    return 0;
}

// F(Q) pair sums of every active device -> their sum in `S` (host, device order)
static int multi_fq_pass(iid_handle *h, const double *pos_host, std::vector<double> &S)
{
    int rc;
    for (int d = 0; d < h->active; ++d) {
        iid_handle *sh = h->subs[d];
        CU(cudaSetDevice(sh->device));
        if ((rc = upload_positions(sh, pos_host))) return rc;
        if ((rc = iid_fq_partial(sh, sh->pos, sh->S, nullptr))) return rc;
        CU(cudaMemcpyAsync(sh->pin + 6 * sh->n, sh->S, sh->nq * sizeof(double),
                           cudaMemcpyDeviceToHost, sh->stream));
    }
    S.assign((size_t)h->nq, 0.0);
    for (int d = 0; d < h->active; ++d) {
        iid_handle *sh = h->subs[d];
        CU(cudaSetDevice(sh->device));
        CU(cudaStreamSynchronize(sh->stream));
        const double *ps = sh->pin + 6 * sh->n;
        for (int64_t m = 0; m < h->nq; ++m) S[m] += ps[m];
    }
    return 0;
}

static int multi_fq_host(iid_handle *h, const double *pos_host, double *F_host, double *pdf_host)
{
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (pdf_host && h->nr == 0) return fail(IID_E_NOTRANSFORM, "call iid_set_transform first");
    if (!pos_host || (!F_host && !pdf_host)) return fail(IID_E_BADARG, "null pointer");
    std::vector<double> S;
    int rc = multi_fq_pass(h, pos_host, S);
    if (rc) return rc;
    std::vector<double> F((size_t)h->nq);
    for (int64_t m = 0; m < h->nq; ++m) F[m] = 2.0 * S[m] * h->inv_na_host[m];
    if (F_host) memcpy(F_host, F.data(), h->nq * sizeof(double));
    if (pdf_host) return iid_fq_to_gr_host(h->subs[0], F.data(), pdf_host);
    return 0;
}

static int multi_grad_fq_host(iid_handle *h, const double *pos_host, void *G_host, double *F_host)
{
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (!pos_host || !G_host) return fail(IID_E_BADARG, "null pointer");
    const size_t esz = h->precision == IID_FP32 ? sizeof(float) : sizeof(double);
    const size_t bytes = (size_t)h->n * 3 * h->nq * esz;
    CU(cudaSetDevice(h->subs[0]->device));
    void *alias = mapped_alias(G_host);
    unsigned char *stage = nullptr;
    if (!alias) {
        // pageable destination: all devices write their rows into one pinned,
        // mapped staging array, copied on by a few host threads afterwards
        if (bytes > h->pinG_bytes) {
            if (h->pinG) cudaFreeHost(h->pinG);
            h->pinG = nullptr;
            h->pinG_bytes = 0;
            CU(cudaHostAlloc((void **)&h->pinG, bytes,
                             cudaHostAllocPortable | cudaHostAllocMapped));
            h->pinG_bytes = bytes;
        }
        stage = h->pinG;
    }
    int rc;
    for (int d = 0; d < h->active; ++d) {
        iid_handle *sh = h->subs[d];
        CU(cudaSetDevice(sh->device));
        void *dst = mapped_alias(stage ? (void *)stage : G_host);
        if (!dst) return fail(IID_E_BADARG, "output array is not mapped on every device");
        if ((rc = upload_positions(sh, pos_host))) return rc;
        if ((rc = iid_grad_fq_partial(sh, sh->pos, dst, sh->S, nullptr))) return rc;
        CU(cudaMemcpyAsync(sh->pin + 6 * sh->n, sh->S, sh->nq * sizeof(double),
                           cudaMemcpyDeviceToHost, sh->stream));
    }
    std::vector<double> S((size_t)h->nq, 0.0);
    for (int d = 0; d < h->active; ++d) {
        iid_handle *sh = h->subs[d];
        CU(cudaSetDevice(sh->device));
        CU(cudaStreamSynchronize(sh->stream));
        const double *ps = sh->pin + 6 * sh->n;
        for (int64_t m = 0; m < h->nq; ++m) S[m] += ps[m];
    }
    if (F_host)
        for (int64_t m = 0; m < h->nq; ++m) F_host[m] = 2.0 * S[m] * h->inv_na_host[m];
    if (stage) {
        const int nthreads = 8;
        std::vector<std::thread> workers;
        const size_t part = (bytes + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; ++t) {
            const size_t o = (size_t)t * part;
            if (o >= bytes) break;
            const size_t l = std::min(part, bytes - o);
            workers.emplace_back([=]() { memcpy((unsigned char *)G_host + o, stage + o, l); });
        }
        for (auto &w : workers) w.join();
    }
    return 0;
}

static int multi_energy_forces_host(iid_handle *h, const double *pos_host,
                                    const double *target_host, int potential, double conv,
                                    double *out_host, double *forces_host, double *pdf_host)
{
    if (h->n == 0) return fail(IID_E_NOSTRUCT, "call iid_set_structure first");
    if (h->active == 1) {  // small structure: the one-device graph replay
        int rc = iid_energy_forces_host(h->subs[0], pos_host, target_host, potential, conv,
                                        out_host, forces_host, pdf_host);
        if (rc == 0 && h->n_restraints) iid_get_restraint_energy(h->subs[0], &h->rs_k[0]);
        return rc;
    }
    if (h->nr == 0) return fail(IID_E_NOTRANSFORM, "call iid_set_transform first");
    if (!pos_host || !out_host) return fail(IID_E_BADARG, "null pointer");
    if (potential != IID_POT_RW && potential != IID_POT_CHI_SQ)
        return fail(IID_E_BADARG, "unknown potential");
    int rc;
    const bool want_forces = forces_host != nullptr;
    // phase 1 on every device: staging + its share of the F(Q) pass
    for (int d = 0; d < h->active; ++d) {
        iid_handle *sh = h->subs[d];
        CU(cudaSetDevice(sh->device));
        double *pt = sh->pin + 6 * sh->n + sh->qp + 8 + sh->nr;
        if ((rc = refresh_target(sh, target_host, pt))) return rc;
        if ((rc = upload_positions(sh, pos_host))) return rc;
        if ((rc = enqueue_eval_device(sh, potential, conv, want_forces, false, EVAL_FQ))) return rc;
        CU(cudaMemcpyAsync(sh->pin + 6 * sh->n, sh->S, sh->nq * sizeof(double),
                           cudaMemcpyDeviceToHost, sh->stream));
    }
    std::vector<double> S((size_t)h->qp, 0.0);
    for (int d = 0; d < h->active; ++d) {
        iid_handle *sh = h->subs[d];
        CU(cudaSetDevice(sh->device));
        CU(cudaStreamSynchronize(sh->stream));
        const double *ps = sh->pin + 6 * sh->n;
        for (int64_t m = 0; m < h->nq; ++m) S[m] += ps[m];
    }
    // phase 2 on every device: summed pair sums -> G(r), Rw, weights (redundant,
    // 11 MB of T) -> its share of the force pass and of the restraints
    for (int d = 0; d < h->active; ++d) {
        iid_handle *sh = h->subs[d];
        CU(cudaSetDevice(sh->device));
        double *pf = sh->pin + 6 * sh->n;
        memcpy(pf, S.data(), h->qp * sizeof(double));
        CU(cudaMemcpyAsync(sh->S, pf, sh->qp * sizeof(double), cudaMemcpyHostToDevice,
                           sh->stream));
        if ((rc = enqueue_eval_device(sh, potential, conv, want_forces, false, EVAL_REST)))
            return rc;
        if (want_forces)
            CU(cudaMemcpyAsync(sh->pin + 3 * sh->n, sh->force, (size_t)3 * sh->n * sizeof(double),
                               cudaMemcpyDeviceToHost, sh->stream));
        CU(cudaMemcpyAsync(pf + sh->qp, sh->out4, 5 * sizeof(double), cudaMemcpyDeviceToHost,
                           sh->stream));
        if (pdf_host && d == 0)
            CU(cudaMemcpyAsync(pf + sh->qp + 8, sh->Gr, sh->nr * sizeof(double),
                               cudaMemcpyDeviceToHost, sh->stream));
    }
    if (want_forces) memset(forces_host, 0, (size_t)3 * h->n * sizeof(double));
    double e_rest = 0.0;
    for (int d = 0; d < h->active; ++d) {
        iid_handle *sh = h->subs[d];
        CU(cudaSetDevice(sh->device));
        CU(cudaStreamSynchronize(sh->stream));
        const double *po = sh->pin + 6 * sh->n + sh->qp;
        if (d == 0) {
            memcpy(out_host, po, 4 * sizeof(double));
            if (pdf_host) memcpy(pdf_host, po + 8, h->nr * sizeof(double));
        }
        e_rest += po[4];
        if (want_forces) {
            const double *pfor = sh->pin + 3 * sh->n;
            for (int64_t k = 0; k < 3 * h->n; ++k) forces_host[k] += pfor[k];
        }
    }
    h->rs_k[0] = e_rest;  // read back by iid_get_restraint_energy
    return 0;
}
