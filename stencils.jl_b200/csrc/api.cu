// api.cu — the extern "C" surface of libstencils_b200.so (include/stencils_b200.h): validation, the plan
// cache (device copies of offset/weight tables + dispatch), host-buffer entry points and memory helpers.
#include <algorithm>
#include <cstdlib>
#include <array>
#include <atomic>
#include <cstdarg>
#include <cmath>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "common.cuh"

namespace sb {

static thread_local char g_err[512] = "";
static thread_local char g_kernel[96] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void set_kernel_name(const char* name) { snprintf(g_kernel, sizeof(g_kernel), "%s", name); }
void count_launch(int n) { g_launches += n; }

int num_sms() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        cached = prop.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

// ------------------------------------------------------------------------------------------------ shapes
// Shape predicates over the box (-R:R)^N iterated with axis 0 fastest (src/stencils/window.jl:4-8,
// moore.jl:5-18, vonneumman.jl:5-15, shapes.jl:2-175).
static int shape_keep(int shape, int R, int RI, int N, const int* t) {
    int manh = 0, maxabs = 0, zeros = 0, sq = 0;
    for (int a = 0; a < N; a++) {
        const int v = std::abs(t[a]);
        manh += v;
        maxabs = std::max(maxabs, v);
        zeros += t[a] == 0;
        sq += t[a] * t[a];
    }
    auto count_tail = [&](auto pred) { int m = 0; for (int a = 1; a < N; a++) m += pred(t[a]) ? 1 : 0; return m; };
    switch (shape) {
    case SB200_WINDOW: return 1;
    case SB200_MOORE: return manh != 0;
    case SB200_VONNEUMANN: return manh >= 1 && manh <= R;
    case SB200_CROSS: return zeros >= N - 1;
    case SB200_ANGLEDCROSS: return count_tail([&](int x) { return std::abs(x) == std::abs(t[0]); }) == N - 1;
    case SB200_FORWARDSLASH: return count_tail([&](int x) { return x == -t[0]; }) == N - 1;
    case SB200_BACKSLASH: return count_tail([&](int x) { return x == t[0]; }) == N - 1;
    case SB200_CIRCLE: return std::sqrt((double)sq) < R + 0.5;
    case SB200_VERTICAL: return (N > 1 && t[1] == 0) || (N == 1 && t[0] == 0);
    case SB200_HORIZONTAL: return N > 1 && t[0] == 0;
    case SB200_DIAMOND: return manh <= R;
    case SB200_ANNULUS: { const double dist = std::sqrt((double)sq); return dist < R + 0.5 && dist >= RI + 0.5; }
    case SB200_CARDINAL: return manh == R && maxabs == R;
    case SB200_ORDINAL: return manh == R * N && maxabs == R;
    default: return -1;
    }
}

static int gen_offsets(int shape, int R, int RI, int N, std::vector<int>& out) {
    out.clear();
    const int D = 2 * R + 1;
    long long total = 1;
    for (int a = 0; a < N; a++) total *= D;
    for (long long lin = 0; lin < total; lin++) {
        int t[3] = {0, 0, 0};
        long long rem = lin;
        for (int a = 0; a < N; a++) { t[a] = (int)(rem % D) - R; rem /= D; }
        const int keep = shape_keep(shape, R, RI, N, t);
        if (keep < 0) return SB200_EUNSUPPORTED;
        if (keep) { out.push_back(t[0]); out.push_back(t[1]); out.push_back(t[2]); }
    }
    return SB200_OK;
}

static int out_eltype_of(int reducer, int eltype, int* out) {
    if (eltype < SB200_BOOL || eltype > SB200_F64) { set_error("unknown eltype %d", eltype); return SB200_EUNSUPPORTED; }
    const bool isf = eltype == SB200_F32 || eltype == SB200_F64;
    switch (reducer) {
    case SB200_SUM: *out = eltype == SB200_BOOL ? SB200_I64 : eltype; return SB200_OK;
    case SB200_MEAN: *out = isf ? eltype : SB200_F64; return SB200_OK;
    case SB200_MIN: case SB200_MAX: case SB200_LIFE: *out = eltype; return SB200_OK;
    case SB200_KERNELDOT:
        if (eltype == SB200_BOOL || eltype == SB200_U8) { set_error("kernelproduct needs Int32/Int64/Float32/Float64"); return SB200_EUNSUPPORTED; }
        *out = eltype; return SB200_OK;
    case SB200_DIFFUSION:
        if (!isf) { set_error("diffusion needs Float32/Float64"); return SB200_EUNSUPPORTED; }
        *out = eltype; return SB200_OK;
    default:
        set_error("unsupported user function (reducer %d): only sum, mean, minimum, maximum, kernelproduct, "
                  "Life and Diffusion lower to CUDA kernels; there is no fallback", reducer);
        return SB200_EUNSUPPORTED;
    }
}

// ------------------------------------------------------------------------------------------------ plans
enum PlanKind { PK_GATHER = 0, PK_HALO = 1, PK_SCATTER = 2 };

static int validate(const sb200_desc* d, int kind) {
    if (!d) { set_error("descriptor is NULL"); return SB200_EINVAL; }
    if (d->struct_size != (int)sizeof(sb200_desc)) { set_error("sb200_desc.struct_size %d != %zu", d->struct_size, sizeof(sb200_desc)); return SB200_EINVAL; }
    if (d->ndim < 1 || d->ndim > 3) { set_error("ndim must be 1..3, got %d", d->ndim); return SB200_EUNSUPPORTED; }
    if (elsize(d->eltype) == 0) { set_error("unknown eltype %d", d->eltype); return SB200_EUNSUPPORTED; }
    if (kind != PK_HALO) {
        if (d->noffsets < 1 || d->noffsets > SB200_MAX_OFFSETS || !d->offsets_host) { set_error("offset table missing or larger than %d", SB200_MAX_OFFSETS); return SB200_EINVAL; }
        int maxabs = 0;
        for (int k = 0; k < d->noffsets; k++)
            for (int a = 0; a < 3; a++) {
                const int o = d->offsets_host[3 * k + a];
                if (a >= d->ndim && o != 0) { set_error("stencil has more dimensions than the array"); return SB200_EINVAL; }
                maxabs = std::max(maxabs, std::abs(o));
            }
        if (d->radius < maxabs) { set_error("radius %d smaller than the largest offset %d", d->radius, maxabs); return SB200_EINVAL; }
    }
    if (d->radius < 0) { set_error("negative radius"); return SB200_EINVAL; }
    for (int a = 0; a < d->ndim; a++) {
        if (d->size[a] < 1) { set_error("empty axis %d", a); return SB200_ESIZE; }
        if (d->src_off[a] < 0 || d->dst_off[a] < 0) { set_error("negative offset"); return SB200_EINVAL; }
        if (d->boundary[a] < SB200_REMOVE || d->boundary[a] > SB200_USE) { set_error("unknown boundary %d", d->boundary[a]); return SB200_EUNSUPPORTED; }
        bool used = kind == PK_HALO;
        if (kind != PK_HALO)
            for (int k = 0; k < d->noffsets; k++) used |= d->offsets_host[3 * k + a] != 0;
        if (d->src_off[a] > 0) {
            if (d->src_off[a] < d->radius && used) { set_error("halo ring %d thinner than the radius %d on axis %d", d->src_off[a], d->radius, a); return SB200_ESIZE; }
            if (d->src_ext[a] < d->size[a] + d->src_off[a] + (used ? d->radius : 0)) { set_error("source parent too small on axis %d", a); return SB200_ESIZE; }
            // update_boundary! reads A[bounded_index(I)] which must land inside the inner region
            if (kind == PK_HALO && (d->boundary[a] == SB200_WRAP || d->boundary[a] == SB200_REFLECT)) {
                const long long hi = d->src_ext[a] - d->src_off[a] - d->size[a];
                const long long need = std::max<long long>(d->src_off[a], hi) + (d->boundary[a] == SB200_REFLECT);
                if (need > d->size[a]) { set_error("axis %d of size %lld is smaller than its halo ring", a, (long long)d->size[a]); return SB200_ESIZE; }
            }
        } else {
            if (d->boundary[a] == SB200_USE) {
                set_error("Use boundary needs Halo padding (no getneighbor method for Use + Conditional, src/array.jl:133-138)");
                return SB200_EUNSUPPORTED;
            }
            if (d->src_ext[a] != d->size[a]) { set_error("Source array sizes must match on axis %d: %lld vs %lld", a, (long long)d->src_ext[a], (long long)d->size[a]); return SB200_ESIZE; }
            if (used && d->radius >= d->size[a]) { set_error("stencil radius is larger than array axis %lld", (long long)d->size[a]); return SB200_ESIZE; }
        }
        if (kind != PK_HALO && d->dst_ext[a] < d->size[a] + d->dst_off[a]) { set_error("Source array sizes must match: dest too small on axis %d", a); return SB200_ESIZE; }
    }
    if (kind == PK_GATHER) {
        int want = 0;
        const int rc = out_eltype_of(d->reducer, d->eltype, &want);
        if (rc) return rc;
        if (want != d->out_eltype) { set_error("out_eltype %d does not match the reducer's result type %d", d->out_eltype, want); return SB200_EINVAL; }
        if (d->reducer == SB200_KERNELDOT && !d->weights_host) { set_error("kernelproduct needs weights"); return SB200_EINVAL; }
        if (d->reducer == SB200_LIFE && d->noffsets > 31) { set_error("Life rule needs at most 31 neighbours"); return SB200_EUNSUPPORTED; }
        bool has_region = false;
        for (int a = 0; a < d->ndim; a++) has_region |= d->region_lo[a] != 0 || d->region_hi[a] != 0;
        if (has_region)
            for (int a = 0; a < d->ndim; a++)
                if (d->region_lo[a] < 0 || d->region_hi[a] > d->size[a] || d->region_lo[a] > d->region_hi[a]) { set_error("bad region on axis %d", a); return SB200_EINVAL; }
    }
    if (kind == PK_SCATTER) {
        if (d->ndim != 2) { set_error("scatterstencil! is 2-D only (src/scatterstencil.jl:49-51)"); return SB200_EUNSUPPORTED; }
        if (d->out_eltype != d->eltype) { set_error("scatterstencil! needs eltype(dest) == eltype(source)"); return SB200_EINVAL; }
        if (!d->weights_host) { set_error("scatter needs weights"); return SB200_EINVAL; }
        if (d->eltype < SB200_I32) { set_error("scatterstencil! supports Int32/Int64/Float32/Float64"); return SB200_EUNSUPPORTED; }
        if (d->scatter_op < SB200_OP_ADD || d->scatter_op > SB200_OP_MIN) { set_error("unsupported scatter op %d", d->scatter_op); return SB200_EUNSUPPORTED; }
        if (d->scatter_rule < 0 || d->scatter_rule > SB200_SCATTER_CENTER_WEIGHTS) { set_error("unsupported scatter rule %d", d->scatter_rule); return SB200_EUNSUPPORTED; }
        if (d->radius > 63 || d->size[0] >= (1LL << 24) || d->size[1] >= (1LL << 23)) { set_error("scatter: radius/size out of range"); return SB200_EUNSUPPORTED; }
    }
    return SB200_OK;
}

static std::mutex g_plan_mu;
thread_local MirrorReq g_mirror;
static std::unordered_map<std::string, Plan*> g_plans;

static unsigned long long g_plan_stamp = 0;   // guarded by g_plan_mu

static void free_plan(Plan* pl) {
    cudaFree(pl->offs_dev); cudaFree(pl->weights_dev); cudaFree(pl->scatter_order_dev);
    delete[] pl->d.offsets_host;
    delete[] (const char*)pl->d.weights_host;
    delete pl;
}

static std::string plan_key(const sb200_desc* d, int kind, int dev) {
    std::string k;
    sb200_desc c = *d;
    c.offsets_host = nullptr;
    c.weights_host = nullptr;
    c.mirror_parent = nullptr; c.mirror_lo = c.mirror_hi = 0;  // per-call, not part of the plan
    k.append((const char*)&kind, sizeof(kind));
    k.append((const char*)&dev, sizeof(dev));
    k.append((const char*)&c, sizeof(c));
    if (kind != PK_HALO) {
        k.append((const char*)d->offsets_host, sizeof(int32_t) * 3 * d->noffsets);
        if (d->weights_host) k.append((const char*)d->weights_host, elsize(d->eltype) * d->noffsets);
    }
    return k;
}

// One-entry memo per thread in front of the cache: a repeated call with the same descriptor (the launch-bound case: one
// small sweep called in a loop) skips validation, key building, the mutex and the hash lookup. An eviction anywhere
// invalidates every memo (g_evict_epoch).
static std::atomic<unsigned long long> g_evict_epoch{0};
struct PlanMemo {
    sb200_desc d;
    int kind = -1, dev = -1;
    unsigned long long epoch = 0;
    Plan* pl = nullptr;
};
static thread_local PlanMemo g_memo;

static bool memo_hit(const sb200_desc* d, int kind, int dev) {
    const PlanMemo& m = g_memo;
    if (!m.pl || m.kind != kind || m.dev != dev || m.epoch != g_evict_epoch.load(std::memory_order_relaxed)) return false;
    sb200_desc c = *d;
    c.offsets_host = m.d.offsets_host; c.weights_host = m.d.weights_host;
    c.mirror_parent = m.d.mirror_parent; c.mirror_lo = m.d.mirror_lo; c.mirror_hi = m.d.mirror_hi;
    if (memcmp(&c, &m.d, sizeof(c)) != 0) return false;
    if ((d->weights_host != nullptr) != (m.pl->d.weights_host != nullptr)) return false;
    if (kind != PK_HALO) {
        if (!d->offsets_host || memcmp(d->offsets_host, m.pl->d.offsets_host, sizeof(int32_t) * 3 * d->noffsets) != 0) return false;
        if (d->weights_host && memcmp(d->weights_host, m.pl->d.weights_host, elsize(d->eltype) * d->noffsets) != 0) return false;
    }
    return true;
}

static int get_plan(const sb200_desc* d, int kind, Plan** out) {
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (d && d->struct_size == (int)sizeof(sb200_desc) && memo_hit(d, kind, dev)) { *out = g_memo.pl; return SB200_OK; }
    int rc = validate(d, kind);
    if (rc) return rc;
    std::string key = plan_key(d, kind, dev);
    std::lock_guard<std::mutex> lock(g_plan_mu);
    auto it = g_plans.find(key);
    auto remember = [&](Plan* pl) {
        g_memo.d = *d; g_memo.kind = kind; g_memo.dev = dev; g_memo.pl = pl;
        g_memo.epoch = g_evict_epoch.load(std::memory_order_relaxed);
    };
    if (it != g_plans.end()) { it->second->stamp = ++g_plan_stamp; *out = it->second; remember(it->second); return SB200_OK; }
    if (g_plans.size() > 4096) {
        // unbounded descriptors (e.g. sliding regions): evict the plans that have not been looked up for a long time. A call
        // holds at most a handful of plans, all of them stamped within its last few lookups, so plans in use are never freed
        // (round 1 dropped the whole cache here, which could leave a caller with a dangling plan).
        g_evict_epoch.fetch_add(1, std::memory_order_relaxed);
        const unsigned long long keep_from = g_plan_stamp > 1024 ? g_plan_stamp - 1024 : 0;
        for (auto jt = g_plans.begin(); jt != g_plans.end();) {
            if (jt->second->stamp < keep_from) { free_plan(jt->second); jt = g_plans.erase(jt); }
            else ++jt;
        }
    }
    Plan* pl = new Plan();
    pl->d = *d;
    pl->key = key;
    DevDesc& p = pl->dd;
    memset(&p, 0, sizeof(p));
    p.ndim = d->ndim; p.L = d->noffsets; p.R = d->radius; p.reducer = d->reducer;
    long long ss = 1, ds = 1;
    bool has_region = false;
    for (int a = 0; a < d->ndim; a++) has_region |= d->region_lo[a] != 0 || d->region_hi[a] != 0;
    for (int a = 0; a < 3; a++) {
        const bool in = a < d->ndim;
        p.size[a] = in ? d->size[a] : 1;
        p.sext[a] = in ? d->src_ext[a] : 1;
        p.sstr[a] = in ? ss : 0; p.dstr[a] = in ? ds : 0;
        if (in) { ss *= d->src_ext[a]; ds *= d->dst_ext[a]; }
        p.soff[a] = in ? d->src_off[a] : 0; p.doff[a] = in ? d->dst_off[a] : 0;
        p.bc[a] = in ? d->boundary[a] : SB200_REMOVE;
        p.lo[a] = (in && has_region) ? d->region_lo[a] : 0;
        p.n[a] = in ? (has_region ? d->region_hi[a] - d->region_lo[a] : d->size[a]) : 1;
        if (in && d->boundary[a] != d->boundary[0]) pl->uniform_bc = false;
    }
    p.padbits = d->padval_bits; p.born = d->born_mask; p.survive = d->survive_mask; p.alpha = d->alpha;
    p.scatter_op = d->scatter_op; p.scatter_rule = d->scatter_rule; p.flags = d->flags;
    if (kind != PK_HALO) {
        const size_t ob = sizeof(int32_t) * 3 * d->noffsets;
        SB_CUDA(cudaMalloc(&pl->offs_dev, ob));
        SB_CUDA(cudaMemcpy(pl->offs_dev, d->offsets_host, ob, cudaMemcpyHostToDevice));
        p.offs = pl->offs_dev;
        if (d->weights_host) {
            const size_t wb = elsize(d->eltype) * d->noffsets;
            SB_CUDA(cudaMalloc(&pl->weights_dev, wb));
            SB_CUDA(cudaMemcpy(pl->weights_dev, d->weights_host, wb, cudaMemcpyHostToDevice));
            p.weights = pl->weights_dev;
        }
        // keep host copies alive inside the plan (callers may free theirs)
        int32_t* oh = new int32_t[3 * d->noffsets];
        memcpy(oh, d->offsets_host, ob);
        pl->d.offsets_host = oh;
        if (d->weights_host) {
            char* wh = new char[elsize(d->eltype) * d->noffsets];
            memcpy(wh, d->weights_host, elsize(d->eltype) * d->noffsets);
            pl->d.weights_host = wh;
        }
        // recognise named shapes (specialised kernels key on them)
        std::vector<int> gen;
        int used_dims = 1;
        for (int k = 0; k < d->noffsets; k++)
            for (int a = 0; a < 3; a++) if (d->offsets_host[3 * k + a] != 0) used_dims = std::max(used_dims, a + 1);
        for (int N = used_dims; N <= d->ndim && pl->shape_tag < 0; N++)
            for (int shape = SB200_WINDOW; shape <= SB200_ORDINAL && pl->shape_tag < 0; shape++) {
                if (shape == SB200_ANNULUS) continue;
                if (gen_offsets(shape, d->radius, 0, N, gen) != SB200_OK) continue;
                if ((int)gen.size() == 3 * d->noffsets && memcmp(gen.data(), d->offsets_host, ob) == 0) { pl->shape_tag = shape; pl->shape_ndim = N; }
            }
    }
    if (kind == PK_SCATTER) {
        // Static fold order per destination-column residue: sort k by (pass of its source column,
        // source row ascending == o0 descending, k). See generic.cu scatter_generic.
        const int S = 2 * d->radius + 1, L = d->noffsets;
        std::vector<int> order(S * L);
        for (int c = 0; c < S; c++) {
            std::vector<std::array<long long, 3>> keys(L);
            for (int k = 0; k < L; k++) {
                const int o0 = d->offsets_host[3 * k], o1 = d->offsets_host[3 * k + 1];
                const long long pass = (((c - o1) % S) + S) % S + 1;
                keys[k] = {pass, -(long long)o0, (long long)k};
            }
            std::vector<int> idx(L);
            for (int k = 0; k < L; k++) idx[k] = k;
            std::sort(idx.begin(), idx.end(), [&](int a, int b) { return keys[a] < keys[b]; });
            for (int q = 0; q < L; q++) order[c * L + q] = idx[q];
        }
        SB_CUDA(cudaMalloc(&pl->scatter_order_dev, sizeof(int) * S * L));
        SB_CUDA(cudaMemcpy(pl->scatter_order_dev, order.data(), sizeof(int) * S * L, cudaMemcpyHostToDevice));
    }
    pl->stamp = ++g_plan_stamp;
    g_plans[key] = pl;
    remember(pl);
    *out = pl;
    return SB200_OK;
}

// Would a sweep with this descriptor's SB200_FLAG_*_STEP flag run as ONE multi-generation launch (16-byte aligned parents assumed)?
bool multistep_accepts(const sb200_desc* d) {
    Plan* pl = nullptr;
    if (get_plan(d, PK_GATHER, &pl) != SB200_OK) return false;
    const int gens = SB200_FLAG_GENS_OF(d->flags);
    if (gens == 2) return life2_accepts(*d, *pl) || diffusion2_accepts(*d, *pl);
    if (gens > 2) return life_multi_accepts(*d, *pl, gens);
    return false;
}

int do_gather(const sb200_desc* d, const void* src, void* dst, cudaStream_t st) {
    if (!src || !dst) { set_error("NULL data pointer"); return SB200_EINVAL; }
    if (src == dst) { set_error("source and dest must not alias (the reference keeps distinct buffers)"); return SB200_EINVAL; }
    Plan* pl = nullptr;
    int rc = get_plan(d, PK_GATHER, &pl);
    if (rc) return rc;
    // The multi-generation kernels need 16-byte aligned parents; anything else would fall through to a single-generation
    // kernel with the flag ignored, so it is refused here (the header's promise: never a silent single sweep).
    if ((d->flags & (SB200_FLAG_DOUBLE_STEP | SB200_FLAG_QUAD_STEP | SB200_FLAG_OCT_STEP)) && (((uintptr_t)src | (uintptr_t)dst) & 15)) {
        set_error("SB200_FLAG_*_STEP: source and dest parents must be 16-byte aligned");
        return SB200_EUNSUPPORTED;
    }
    const int gens = SB200_FLAG_GENS_OF(d->flags);
    if (gens > 2 && !life_multi_accepts(*d, *pl, gens)) {
        set_error("SB200_FLAG_GENS(%d): only B3/S23 Life / Moore(1) on an unpadded Bool or UInt8 grid, Wrap on axis 0, width %% 32 == 0%s",
                  gens, gens > 4 ? " (and a library built with -DSB200_LB_ONE_HALO_LANE=1, the default)" : "");
        return SB200_EUNSUPPORTED;
    }
    if ((d->flags & (SB200_FLAG_SRC_BITS | SB200_FLAG_DST_BITS)) && !life_multi_accepts(*d, *pl, gens)) {
        set_error("SB200_FLAG_SRC_BITS / _DST_BITS: packed state is read and written by the bit-sliced Life kernel only (B3/S23, Moore(1), "
                  "SB200_FLAG_GENS(2 .. 8) — a packed source also one generation —, axis 0 and the packed parents' axis-0 extents multiples of 128 cells, Wrap on axis 0)");
        return SB200_EUNSUPPORTED;
    }
    if (gens == 2 && !life2_accepts(*d, *pl) && !diffusion2_accepts(*d, *pl)) {
        set_error("SB200_FLAG_DOUBLE_STEP: only Life / Moore(1) on an unpadded Bool or UInt8 grid with Wrap on axis 0, or "
                  "Diffusion / VonNeumann(1,3) on an unpadded Float32 / Float64 grid with Wrap on axes 0 and 1");
        return SB200_EUNSUPPORTED;
    }
    g_mirror = MirrorReq();
    if (d->mirror_parent && d->mirror_hi > d->mirror_lo) {
        const int last = d->ndim - 1;
        const long long lo = pl->dd.lo[last], hi = lo + pl->dd.n[last];
        if (d->mirror_lo < lo || d->mirror_hi > hi) { set_error("mirror planes [%lld,%lld) lie outside the output region", (long long)d->mirror_lo, (long long)d->mirror_hi); return SB200_EINVAL; }
        for (int a = 0; a < last; a++)
            if (d->dst_off[a] != 0 || pl->dd.lo[a] != 0 || pl->dd.n[a] != d->size[a]) { set_error("mirror needs whole, unpadded dest planes"); return SB200_EUNSUPPORTED; }
        g_mirror.ptr = d->mirror_parent; g_mirror.lo = d->mirror_lo; g_mirror.hi = d->mirror_hi;
    }
    rc = -1;
    if (!(d->flags & SB200_FLAG_FORCE_GENERIC)) {
        rc = try_life_swar(*pl, src, dst, st);
        if (rc < 0) rc = try_diffusion3d(*pl, src, dst, st);
        if (rc < 0) rc = try_small2d(*pl, src, dst, st);   // radius-1 shapes on grids that fit the L2 (small2d.cu)
        if (rc < 0) rc = try_tile2d(*pl, src, dst, st);
        if (rc < 0) rc = try_gather_stream(*pl, src, dst, st);
        if (rc < 0) rc = try_box3d(*pl, src, dst, st);
        if (rc < 0) rc = try_gather_stream3d(*pl, src, dst, st);
    }
    if (rc < 0) rc = launch_generic_gather(*pl, src, dst, st);
    if (rc == SB200_OK && g_mirror.ptr && !g_mirror.honoured) {
        // no fused store in the kernel that ran: copy the planes after the sweep (same stream)
        const int last = d->ndim - 1;
        size_t plane = elsize(d->out_eltype);
        for (int a = 0; a < last; a++) plane *= (size_t)d->dst_ext[a];
        const char* from = (const char*)dst + (size_t)(g_mirror.lo + d->dst_off[last]) * plane;
        SB_CUDA(cudaMemcpyAsync(g_mirror.ptr, from, (size_t)(g_mirror.hi - g_mirror.lo) * plane, cudaMemcpyDeviceToDevice, st));
    }
    g_mirror = MirrorReq();
    return rc;
}

static int do_halo(const sb200_desc* d, void* parent, cudaStream_t st) {
    if (!parent) { set_error("NULL data pointer"); return SB200_EINVAL; }
    bool any = false;
    for (int a = 0; a < d->ndim && a < 3; a++) any |= d->src_off[a] > 0 && d->boundary[a] != SB200_USE;
    if (!any) return SB200_OK;
    Plan* pl = nullptr;
    sb200_desc c = *d;  // the halo plan does not depend on the stencil table / reducer / region
    c.noffsets = 0; c.offsets_host = nullptr; c.weights_host = nullptr; c.reducer = 0; c.out_eltype = c.eltype;
    c.flags = 0; c.alpha = 0; c.born_mask = c.survive_mask = 0; c.scatter_op = c.scatter_rule = 0;
    memset(c.region_lo, 0, sizeof(c.region_lo)); memset(c.region_hi, 0, sizeof(c.region_hi));
    memset(c.dst_ext, 0, sizeof(c.dst_ext)); memset(c.dst_off, 0, sizeof(c.dst_off));
    int rc = get_plan(&c, PK_HALO, &pl);
    if (rc) return rc;
    return launch_update_halo(*pl, parent, st);
}

static bool needs_halo(const sb200_desc* d) {
    for (int a = 0; a < d->ndim && a < 3; a++)
        if (d->src_off[a] > 0 && d->boundary[a] != SB200_USE) return true;
    return false;
}

// ---- peer-memory helpers ----
__global__ void push_planes_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16,
                                   const unsigned char* __restrict__ srcb, unsigned char* __restrict__ dstb, size_t tail0,
                                   size_t bytes, uint32_t* flag, uint32_t value, unsigned int* done_ctr) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
    for (size_t i = tail0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < bytes; i += (size_t)gridDim.x * blockDim.x)
        dstb[i] = srcb[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done_ctr, 1u);
        if (prev == gridDim.x - 1) {  // last block: every block's stores are fenced -> publish
            *done_ctr = 0;
            __threadfence_system();
            if (flag) {
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
            }
        }
    }
}

__global__ void signal_flag_kernel(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// Acquire spin with a deadline: a neighbour that never publishes (dead process, desynchronised schedule) must not hang the GPU
// for good. After `timeout_ns` the kernel traps: the stream's next synchronisation returns a launch failure the caller sees.
// (The slab plans use plan_wait_kernel, which reports through an error word instead: csrc/slab_plan.cu.)
__global__ void wait_flag_kernel(const uint32_t* flag, uint32_t value, unsigned long long timeout_ns) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t v;
    for (unsigned it = 0;; it++) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - value) >= 0) return;
        __nanosleep(200);
        if ((it & 1023) == 1023) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) __trap();
        }
    }
}

// ---- multi-array gather: dest = t_1 + t_2 + ... over the logical box of the dest parent ----
// MODE 0: d = c*s   MODE 1: d = d + s   MODE 2: d = d + c*s   (every operation rounded separately)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) combine_kernel(T* __restrict__ d, const T* __restrict__ s, T c, long long n0, long long n1,
                                                      long long n2, long long e0, long long e1, int o0, int o1, int o2) {
    const long long total = n0 * n1 * n2;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const long long i0 = id % n0, r = id / n0, i1 = r % n1, i2 = r / n1;
        const long long idx = (i0 + o0) + e0 * ((i1 + o1) + e1 * (i2 + o2));
        const T x = s[idx];
        if (MODE == 0) d[idx] = mul_rn(c, x);
        else if (MODE == 1) d[idx] = add_rn(d[idx], x);
        else d[idx] = add_rn(d[idx], mul_rn(c, x));
    }
}
template <typename T> static int launch_combine(int mode, void* d, const void* s, double c, const sb200_desc* ds, cudaStream_t st) {
    const long long n0 = ds->size[0], n1 = ds->ndim > 1 ? ds->size[1] : 1, n2 = ds->ndim > 2 ? ds->size[2] : 1;
    const long long e0 = ds->dst_ext[0], e1 = ds->ndim > 1 ? ds->dst_ext[1] : 1;
    const int o0 = ds->dst_off[0], o1 = ds->ndim > 1 ? ds->dst_off[1] : 0, o2 = ds->ndim > 2 ? ds->dst_off[2] : 0;
    const long long total = n0 * n1 * n2;
    if (total == 0) return SB200_OK;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16);
    if (mode == 0) combine_kernel<T, 0><<<blocks, 256, 0, st>>>((T*)d, (const T*)s, (T)c, n0, n1, n2, e0, e1, o0, o1, o2);
    else if (mode == 1) combine_kernel<T, 1><<<blocks, 256, 0, st>>>((T*)d, (const T*)s, (T)c, n0, n1, n2, e0, e1, o0, o1, o2);
    else combine_kernel<T, 2><<<blocks, 256, 0, st>>>((T*)d, (const T*)s, (T)c, n0, n1, n2, e0, e1, o0, o1, o2);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}
// library-owned scratch of sb200_gather_multi: one buffer per (thread, stream) — a single buffer shared by calls on different
// streams would be a cross-stream race
struct MultiScratch { void* p = nullptr; size_t bytes = 0; };
static thread_local std::unordered_map<cudaStream_t, MultiScratch> g_multi_scratch;

static thread_local unsigned int* g_done_ctr = nullptr;  // 64 counters, one per in-flight push
static thread_local unsigned g_push_seq = 0;

}  // namespace sb

using namespace sb;

namespace sb {
// Relative launch times by generations per launch (tools/life_gens_probe.py, r02w, Life 16384^2 with 0/1 cells, us per launch: 94.9
// (one generation, life_tma_kernel), 107.7, 110.4, 108.6, 101.2, 111.3, 124.2, 137.8 (two .. eight, life_bit_kernel<G>): eight
// generations per launch is the best rate, 15.6 Tcell-updates/s; diffusion: the two-step kernel runs at 1.45x the single-step rate).
const double kLifeLaunchCost[kMaxGens + 1] = {0, 1.00, 1.13, 1.16, 1.14, 1.07, 1.17, 1.31, 1.45};
// packed -> packed launches, us (r02y): 30.5, 37.6, 45.7, 52.4, 62.2, 72.4, 86.6, 98.9 — six generations per launch is the best rate
// (22.2 Tcell-updates/s; 21.7 for seven and eight)
const double kLifePackedLaunchCost[kMaxGens + 1] = {0, 30.5, 37.6, 45.7, 52.4, 62.2, 72.4, 86.6, 98.9};
}  // namespace sb

extern "C" {

int32_t sb200_version(void) { return SB200_VERSION; }
const char* sb200_last_error(void) { return g_err; }
const char* sb200_last_kernel(void) { return g_kernel; }
int64_t sb200_launch_count(int32_t reset) {
    const long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int32_t sb200_stencil_offsets(int32_t shape, int32_t radius, int32_t inner_radius, int32_t ndim, int32_t* out,
                              int32_t cap, int32_t* count) {
    if (!count || ndim < 1 || ndim > 3 || radius < 0 || radius > 64) { set_error("bad arguments to sb200_stencil_offsets"); return SB200_EINVAL; }
    std::vector<int> gen;
    const int rc = gen_offsets(shape, radius, inner_radius, ndim, gen);
    if (rc) { set_error("unknown stencil shape %d", shape); return rc; }
    const int L = (int)gen.size() / 3;
    if (out)
        for (int k = 0; k < L && k < cap; k++) { out[3 * k] = gen[3 * k]; out[3 * k + 1] = gen[3 * k + 1]; out[3 * k + 2] = gen[3 * k + 2]; }
    *count = L;
    return SB200_OK;
}

int32_t sb200_out_eltype(int32_t reducer, int32_t eltype, int32_t* out) {
    if (!out) { set_error("NULL out"); return SB200_EINVAL; }
    return out_eltype_of(reducer, eltype, out);
}
size_t sb200_sizeof(int32_t eltype) { return elsize(eltype); }

int32_t sb200_gather(const sb200_desc* d, const void* src, void* dst, void* stream) {
    return do_gather(d, src, dst, (cudaStream_t)stream);
}
int32_t sb200_update_halo(const sb200_desc* d, void* parent, void* stream) {
    const int rc = validate(d, PK_HALO);
    if (rc) return rc;
    return do_halo(d, parent, (cudaStream_t)stream);
}
int32_t sb200_scatter(const sb200_desc* d, const void* src, void* dst, void* stream) {
    if (!src || !dst) { set_error("NULL data pointer"); return SB200_EINVAL; }
    Plan* pl = nullptr;
    int rc = get_plan(d, PK_SCATTER, &pl);
    if (rc) return rc;
    if (!(d->flags & SB200_FLAG_FORCE_GENERIC)) {
        rc = try_scatter_fast(*pl, src, dst, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    return launch_generic_scatter(*pl, src, dst, (cudaStream_t)stream);
}

static size_t parent_bytes(const sb200_desc* d, bool src);
int32_t sb200_gather_multi(const sb200_term* terms, int32_t nterms, void* dst, void* scratch, void* stream) {
    if (!terms || nterms < 1 || nterms > SB200_MAX_TERMS || !dst) { set_error("sb200_gather_multi: 1..%d terms and a dest are required", SB200_MAX_TERMS); return SB200_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const sb200_desc* d0 = terms[0].desc;
    for (int j = 0; j < nterms; j++) {
        const sb200_desc* d = terms[j].desc;
        if (!d || !terms[j].src_parent) { set_error("sb200_gather_multi: term %d has no descriptor / source", j); return SB200_EINVAL; }
        if (d->eltype != d0->eltype || d->out_eltype != d->eltype || (d->eltype != SB200_F32 && d->eltype != SB200_F64)) {
            set_error("sb200_gather_multi supports Float32 / Float64 arguments of one element type");
            return SB200_EUNSUPPORTED;
        }
        if (d->ndim != d0->ndim) { set_error("Source array sizes must match (dimension of argument %d)", j); return SB200_ESIZE; }
        for (int a = 0; a < d->ndim; a++)
            if (d->size[a] != d0->size[a] || d->dst_ext[a] != d0->dst_ext[a] || d->dst_off[a] != d0->dst_off[a]) {
                set_error("Source array sizes must match. Found a different size / dest layout for argument %d on axis %d", j, a);  // _checksizes
                return SB200_ESIZE;
            }
    }
    // One pass over all arguments where the kernel of multi_tile.cu takes the combination (opt-in, SB200_MULTI_SINGLE_PASS=1; 2-D
    // Float32 / Float64, the reducer menu of the streaming kernels, R <= 4): every source is read once, the dest written once, no
    // scratch parent.
    {
        Plan* pls[SB200_MAX_TERMS];
        sb200_desc dj[SB200_MAX_TERMS];
        bool ok = true;
        for (int j = 0; j < nterms && ok; j++) {
            dj[j] = *terms[j].desc;
            memset(dj[j].region_lo, 0, sizeof(dj[j].region_lo)); memset(dj[j].region_hi, 0, sizeof(dj[j].region_hi));
            ok = get_plan(&dj[j], PK_GATHER, &pls[j]) == SB200_OK;
        }
        if (ok && try_multi_tile2d(pls, terms, nterms, dst, st, true) == SB200_OK) {
            for (int j = 0; j < nterms; j++) {
                int rc;
                if (needs_halo(&dj[j]) && (rc = do_halo(&dj[j], const_cast<void*>(terms[j].src_parent), st))) return rc;   // update_boundary!(source)
            }
            return try_multi_tile2d(pls, terms, nterms, dst, st, false);
        }
    }
    const size_t db = parent_bytes(d0, false);
    const bool need_scratch = nterms > 1 || terms[0].has_coef;
    if (need_scratch && !scratch) {
        MultiScratch& ms = g_multi_scratch[st];
        if (ms.bytes < db) {
            if (ms.p) { SB_CUDA(cudaStreamSynchronize(st)); cudaFree(ms.p); }   // the previous call on this stream may still use it
            ms.p = nullptr; ms.bytes = 0;
            SB_CUDA(cudaMalloc(&ms.p, db));
            ms.bytes = db;
        }
        scratch = ms.p;
    }
    int rc;
    for (int j = 0; j < nterms; j++) {
        sb200_desc d = *terms[j].desc;
        memset(d.region_lo, 0, sizeof(d.region_lo)); memset(d.region_hi, 0, sizeof(d.region_hi));
        void* src = const_cast<void*>(terms[j].src_parent);
        if (needs_halo(&d) && (rc = do_halo(&d, src, st))) return rc;   // update_boundary!(source), src/gatherstencil.jl:93-95
        const bool direct = j == 0 && !terms[0].has_coef;
        if ((rc = do_gather(&d, src, direct ? dst : scratch, st))) return rc;
        if (direct) continue;
        const int mode = j == 0 ? 0 : (terms[j].has_coef ? 2 : 1);
        rc = d.eltype == SB200_F32 ? launch_combine<float>(mode, dst, scratch, terms[j].coef, &d, st)
                                   : launch_combine<double>(mode, dst, scratch, terms[j].coef, &d, st);
        if (rc) return rc;
    }
    return SB200_OK;
}

static const double* const kLifeCost = sb::kLifeLaunchCost;
static const double kDiffCost[kMaxGens + 1] = {0, 1.00, 1.38, 0, 0, 0, 0, 0, 0};

// Split nsteps generations into launches of the allowed sizes (ok_size[1] is always true): least total cost with an odd / even
// number of launches for an odd / even step count (the final state must land in the buffer sb200_iterate's contract names).
// The bulk of a long run is launches of the cheapest size per generation; a dynamic programme over (generations, launch-count
// parity) covers the last <= 72 generations. cnt[g] = launches of g generations.
static void split_steps(int nsteps, const bool* ok_size, const double* cost, int* cnt) {
    constexpr int MAXG = kMaxGens;
    for (int g = 0; g <= MAXG; g++) cnt[g] = 0;
    cnt[1] = nsteps;
    int bulk = 1;
    for (int g = 2; g <= MAXG; g++) if (ok_size[g] && cost[g] / g < cost[bulk] / bulk) bulk = g;
    if (bulk == 1) return;
    constexpr int TAIL = 64;
    const int nb0 = nsteps > TAIL ? (nsteps - TAIL + bulk - 1) / bulk : 0;   // launches of `bulk` before the tail
    double dp[TAIL + MAXG + 1][2];
    int from[TAIL + MAXG + 1][2];
    const int tmax = std::min(nsteps, TAIL + MAXG);
    for (int n = 0; n <= tmax; n++) dp[n][0] = dp[n][1] = 1e300;
    dp[0][0] = 0;
    for (int n = 1; n <= tmax; n++)
        for (int q = 0; q < 2; q++)
            for (int g = 1; g <= MAXG && g <= n; g++)
                if ((g == 1 || ok_size[g]) && dp[n - g][q ^ 1] + cost[g] < dp[n][q]) { dp[n][q] = dp[n - g][q ^ 1] + cost[g]; from[n][q] = g; }
    // the parity of the tail's launch count depends on the number of bulk launches: try nb0 and nb0 - 1
    double best = 1e300;
    int best_nb = -1;
    for (int nb = nb0; nb >= 0 && nb >= nb0 - 1; nb--) {
        const int tail = nsteps - nb * bulk;
        if (tail > tmax) break;
        const double c = nb * cost[bulk] + dp[tail][(nsteps ^ nb) & 1];
        if (c < best) { best = c; best_nb = nb; }
    }
    if (best_nb < 0 || best > 1e299) return;
    cnt[1] = 0;
    cnt[bulk] = best_nb;
    int q = (nsteps ^ best_nb) & 1;
    for (int n = nsteps - best_nb * bulk; n > 0; q ^= 1) { const int g = from[n][q]; cnt[g]++; n -= g; }
}

// Test hook (tests/test_host_api.py): the split sb200_iterate would use for `nsteps` with the sizes in `size_mask` (bit g = launches
// of g generations allowed) and the Life (1) or diffusion (0) cost table. out[0 .. 8].
int32_t sb200_debug_split_steps(int32_t nsteps, int32_t size_mask, int32_t life, int32_t* out) {
    if (nsteps < 0 || !out) return SB200_EINVAL;
    bool ok[kMaxGens + 1];
    for (int g = 0; g <= kMaxGens; g++) ok[g] = g == 1 || ((size_mask >> g) & 1);
    int cnt[kMaxGens + 1];
    split_steps(nsteps, ok, life ? kLifeCost : kDiffCost, cnt);
    for (int g = 0; g <= kMaxGens; g++) out[g] = cnt[g];
    return SB200_OK;
}

// sb200_iterate schedules two diffusion steps per launch by itself when this is true (SB200_DIFFUSION_DOUBLE_STEP overrides).
static constexpr bool kDiffusionDoubleStepDefault = true;

int32_t sb200_iterate(const sb200_desc* d, void* buf_a, void* buf_b, int32_t nsteps, void* stream) {
    if (nsteps < 0) { set_error("negative step count"); return SB200_EINVAL; }
    if (!d) { set_error("descriptor is NULL"); return SB200_EINVAL; }
    for (int a = 0; a < d->ndim && a < 3; a++)
        if (d->src_ext[a] != d->dst_ext[a] || d->src_off[a] != d->dst_off[a]) {
            set_error("source and dest arrays must be the same size (src/array.jl:574-575)");
            return SB200_ESIZE;
        }
    if (d->eltype != d->out_eltype) { set_error("iterated sweeps need a reducer that preserves the element type"); return SB200_EUNSUPPORTED; }
    void *s = buf_a, *t = buf_b;
    const bool halo = needs_halo(d);
    // After the first Life step on UInt8 the source is the kernel's own 0/1 output: later steps may skip the
    // "cell != 0" normalisation (the Remove padval is normalised separately, and rings only exist on axes the
    // packed kernel does not accept).
    sb200_desc later = *d;
    if (d->reducer == SB200_LIFE && d->eltype == SB200_U8) later.flags |= SB200_FLAG_CELLS_01;
    // Several generations per launch where a kernel supports it (Life: 2 .. 8 with the intermediate generations in
    // registers, life.cu; Diffusion: 2, stream3d2.cu). The run is split into launches of 1 .. maxg generations with the least
    // estimated time (kLaunchCost below: a launch of up to 4 Life generations costs about one trip through HBM, larger ones are
    // ALU-bound) and an odd / even number of launches for an odd / even step count, so that the final state lands in the buffer
    // the contract names: long runs are launches of 8 plus a fix-up, 20 steps are 5 + 5 + 5 + 5. Small launches go first.
    // SB200_NO_DOUBLE_STEP / SB200_NO_QUAD_STEP / SB200_OCT_STEP=0 / SB200_DIFFUSION_DOUBLE_STEP=0 cap the size (A/B runs),
    // SB200_POW2_STEPS=1 keeps the round-2 sizes 1 / 2 / 4 / 8.
    // The multi-generation kernels need 16-byte aligned parents: other pointers run one generation per launch.
    const bool aligned16 = ((((uintptr_t)buf_a | (uintptr_t)buf_b) & 15) == 0);
    const bool life = d->reducer == SB200_LIFE;
    const char* e_d2 = getenv("SB200_DIFFUSION_DOUBLE_STEP");
    const bool diff2 = d->reducer == SB200_DIFFUSION && (e_d2 ? atoi(e_d2) != 0 : kDiffusionDoubleStepDefault);
    constexpr int MAXG = kMaxGens;
    bool ok_size[MAXG + 1] = {false, true, false, false, false, false, false, false, false};
    if (aligned16 && (life || diff2) && nsteps >= 4 && !halo && !(d->flags & SB200_FLAG_STEP_MASK) && !getenv("SB200_NO_DOUBLE_STEP")) {
        auto accepts = [&](int gens) {
            sb200_desc probe = *d;
            probe.flags |= SB200_FLAG_GENS(gens);
            return multistep_accepts(&probe);
        };
        const bool pow2 = getenv("SB200_POW2_STEPS") && atoi(getenv("SB200_POW2_STEPS")) != 0;
        int cap = life ? MAXG : 2;
        if (getenv("SB200_NO_QUAD_STEP")) cap = 2;
        else if (getenv("SB200_OCT_STEP") && atoi(getenv("SB200_OCT_STEP")) == 0) cap = std::min(cap, 4);
        ok_size[2] = accepts(2);
        for (int g = 3; g <= cap && ok_size[2]; g++) ok_size[g] = !(pow2 && (g & (g - 1))) && accepts(g);
    }
    // Life runs keep their state PACKED (one bit per cell, SB200_FLAG_SRC_BITS / _DST_BITS) between the first and the last launch:
    // the first launch reads the bytes of buf_a and writes bits, the last one reads bits and writes the bytes of the buffer the
    // contract names, every launch in between is packed -> packed (no pack / unpack instructions, 1 / 8 of the memory traffic).
    // The packed grids live inside the two buffers themselves (each holds eight of them; regions 0 and 1 = the two halves of a
    // buffer are used), always in the buffer that is neither being read as bytes nor about to receive the final bytes:
    //   final state in buf_a:  a(bytes) -> b.0 -> b.1 -> b.0 ... -> a(bytes)
    //   final state in buf_b:  a(bytes) -> b.0 -> a.0 -> a.1 -> a.0 ... -> b(bytes)
    // so the launch count carries no parity constraint: launches of the best size plus one remainder. The content of the other
    // buffer after the call is unspecified (as it already is with several generations per launch).
    // SB200_LIFE_PACKED=0 turns the packed runs off, =1 forces them for grids of any size. Default: grids above 4 Mi cells — smaller ones
    // are launch-bound and replay byte launches as CUDA graphs below (tools/life_small_probe.py, r02ah, 1000 generations: 2048^2 2148
    // Gcell-updates/s with graphs against 1612 packed, 1024^2 599 against 602; 4096^2 5429 against 6095 packed, 8192^2 11 339 against 13 948).
    {
        long long ncells = 1, region_full = 1;
        for (int a = 0; a < d->ndim; a++) {
            ncells *= d->size[a];
            region_full &= d->region_lo[a] == 0 && (d->region_hi[a] == 0 || d->region_hi[a] == d->size[a]);
        }
        const char* e_pk = getenv("SB200_LIFE_PACKED");
        const bool want_packed = e_pk ? atoi(e_pk) != 0 : ncells > (4LL << 20);
        if (life && ok_size[8] && want_packed && region_full && nsteps >= 12 && !(d->flags & (SB200_FLAG_SRC_BITS | SB200_FLAG_DST_BITS))) {
            sb200_desc probe = *d;
            probe.flags |= SB200_FLAG_GENS(8) | SB200_FLAG_SRC_BITS | SB200_FLAG_DST_BITS;
            if (multistep_accepts(&probe)) {
                const size_t half = ((size_t)d->src_ext[0] * (size_t)d->src_ext[1] / 2) & ~(size_t)15;
                char *A0 = (char*)buf_a, *B0 = (char*)buf_b;
                const bool final_in_a = (nsteps & 1) == 0;
                // L = the fewest launches of <= kLifePackedBulkGens generations (at least two; three when the final bytes go to buf_b, whose first
                // packed grid must be dead by then), the generations spread evenly over them, smaller launches first
                const char* e_g = getenv("SB200_LIFE_PACKED_GENS");   // A/B: the largest launch size of a packed run (default: the best measured)
                const int gmax = e_g ? std::min(8, std::max(2, atoi(e_g))) : kLifePackedBulkGens;
                const int L = std::max((nsteps + gmax - 1) / gmax, final_in_a ? 2 : 3);
                const int base = nsteps / L, extra = nsteps % L;   // L - extra launches of `base`, then `extra` of base + 1
                const void* from = buf_a;
                int region = 0;           // region of the buffer that holds the packed state being read next
                bool in_a = false;        // ... and whether that buffer is buf_a
                for (int j = 0; j < L; j++) {
                    const int g = j < L - extra ? base : base + 1;
                    sb200_desc cur = *d;
                    cur.flags |= SB200_FLAG_GENS(g);
                    void* to;
                    if (j == 0) {                       // bytes of buf_a -> packed, region 0 of buf_b
                        cur.flags |= SB200_FLAG_DST_BITS;
                        to = B0; in_a = false; region = 0;
                    } else if (j == L - 1) {            // packed -> bytes of the final buffer (the packed source is in the other one)
                        cur.flags |= SB200_FLAG_SRC_BITS;
                        to = final_in_a ? buf_a : buf_b;
                    } else {
                        cur.flags |= SB200_FLAG_SRC_BITS | SB200_FLAG_DST_BITS;
                        if (in_a == final_in_a) {       // the packed state sits in the final buffer: hop to the other one (second launch only)
                            in_a = !in_a; region = 0;
                        } else {
                            region ^= 1;
                        }
                        to = (in_a ? A0 : B0) + (size_t)region * half;
                    }
                    const int rc = do_gather(&cur, from, to, (cudaStream_t)stream);
                    if (rc) return rc;
                    from = to;
                }
                return SB200_OK;
            }
        }
    }
    int cnt[kMaxGens + 1];   // launches of 1 .. 8 generations
    split_steps(nsteps, ok_size, life ? kLifeCost : kDiffCost, cnt);
    // One launch of the loop body: [ring refresh] + sweep from -> to (gens generations).
    bool fresh = true;   // the next launch is the first one of the call (UInt8 cells not yet known to be 0/1)
    auto body = [&](int gens, void* from, void* to) -> int {
        int rc;
        if (halo && (rc = sb200_update_halo(d, from, stream))) return rc;
        sb200_desc cur = fresh ? *d : later;
        if (gens > 1) cur.flags |= SB200_FLAG_GENS(gens);
        rc = do_gather(&cur, from, to, (cudaStream_t)stream);
        if (rc == SB200_OK) fresh = false;
        return rc;
    };
    // Small grids are launch-bound (a 1000 x 1000 sweep takes ~3 us of GPU time): once the plans exist, a run of equal
    // launches is captured ONCE as a CUDA graph of GRAPH_CHUNK launches and replayed, so the per-launch CPU + driver cost
    // is paid per chunk. Capture needs a real stream (not the legacy default stream) and an even chunk (same buffer roles).
    long long cells = 1;
    for (int a = 0; a < d->ndim; a++) cells *= d->size[a];
    constexpr int GRAPH_CHUNK = 32;
    bool want_graph = stream != nullptr && cells <= (4LL << 20) && !getenv("SB200_NO_GRAPH");
    for (int per = 1; per <= MAXG; per++) {
        int left = cnt[per];
        int issued = 0;   // direct launches of this size so far (the second one runs with the steady-state descriptor: its plan exists)
        while (left > 0) {
            if (want_graph && issued >= 2 && !fresh && left >= 2 * GRAPH_CHUNK) {
                cudaStream_t cs = (cudaStream_t)stream;
                cudaGraph_t graph = nullptr;
                cudaGraphExec_t exec = nullptr;
                bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                int rc = SB200_OK;
                if (ok) {
                    void *cs_ = s, *ct_ = t;
                    for (int j = 0; j < GRAPH_CHUNK && rc == SB200_OK; j++) {
                        rc = body(per, cs_, ct_);
                        void* tmp = cs_; cs_ = ct_; ct_ = tmp;
                    }
                    ok = cudaStreamEndCapture(cs, &graph) == cudaSuccess && rc == SB200_OK && graph != nullptr;
                }
                if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
                if (ok) {
                    const int chunks = left / GRAPH_CHUNK;
                    for (int c = 0; c < chunks && ok; c++) ok = cudaGraphLaunch(exec, cs) == cudaSuccess;
                    if (ok) {
                        left -= chunks * GRAPH_CHUNK;                       // the buffer roles are unchanged after an even number of launches
                        count_launch(chunks * GRAPH_CHUNK - GRAPH_CHUNK);   // the captured launches were counted once already
                    }
                }
                if (exec) cudaGraphExecDestroy(exec);
                if (graph) cudaGraphDestroy(graph);
                if (ok) continue;
                cudaGetLastError();   // capture unavailable (e.g. an enclosing capture): plain launches from here on
                want_graph = false;   // nothing was enqueued
                continue;
            }
            int rc;
            if ((rc = body(per, s, t))) return rc;
            issued++;
            left--;
            void* tmp = s; s = t; t = tmp;
        }
    }
    return SB200_OK;
}

// ---- host-buffer entry points ----
static size_t parent_bytes(const sb200_desc* d, bool src) {
    size_t n = 1;
    for (int a = 0; a < d->ndim; a++) n *= (size_t)(src ? d->src_ext[a] : d->dst_ext[a]);
    return n * elsize(src ? d->eltype : d->out_eltype);
}

struct HostScratch {
    void* a = nullptr; size_t a_bytes = 0;
    void* b = nullptr; size_t b_bytes = 0;
    cudaStream_t st[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[128];
    bool init = false;
};
static thread_local HostScratch g_hs;

static int ensure_scratch(size_t ab, size_t bb) {
    if (!g_hs.init) {
        for (int i = 0; i < 3; i++) SB_CUDA(cudaStreamCreateWithFlags(&g_hs.st[i], cudaStreamNonBlocking));
        for (int i = 0; i < 128; i++) SB_CUDA(cudaEventCreateWithFlags(&g_hs.ev[i], cudaEventDisableTiming));
        g_hs.init = true;
    }
    if (g_hs.a_bytes < ab) { if (g_hs.a) cudaFree(g_hs.a); g_hs.a = nullptr; g_hs.a_bytes = 0; SB_CUDA(cudaMalloc(&g_hs.a, ab)); g_hs.a_bytes = ab; }
    if (g_hs.b_bytes < bb) { if (g_hs.b) cudaFree(g_hs.b); g_hs.b = nullptr; g_hs.b_bytes = 0; SB_CUDA(cudaMalloc(&g_hs.b, bb)); g_hs.b_bytes = bb; }
    return SB200_OK;
}

int32_t sb200_gather_host(const sb200_desc* d, const void* src_host, void* dst_host) {
    int rc = validate(d, PK_GATHER);
    if (rc) return rc;
    if (!src_host || !dst_host) { set_error("NULL data pointer"); return SB200_EINVAL; }
    const size_t sb_ = parent_bytes(d, true), db = parent_bytes(d, false);
    if ((rc = ensure_scratch(sb_, db))) return rc;
    const int last = d->ndim - 1;
    bool has_region = false;
    for (int a = 0; a < d->ndim; a++) has_region |= d->region_lo[a] != 0 || d->region_hi[a] != 0;
    // Chunk the slowest axis so H2D, the sweep and D2H of neighbouring chunks overlap (3 streams).
    // Chunking needs each output chunk to depend only on a contiguous band of source planes, which holds
    // when the slowest axis is not wrapped/reflected on the fly and there is no ring to refresh.
    const long long nlast = d->size[last];
    size_t plane_src = elsize(d->eltype), plane_dst = elsize(d->out_eltype);
    for (int a = 0; a < last; a++) { plane_src *= (size_t)d->src_ext[a]; plane_dst *= (size_t)d->dst_ext[a]; }
    const bool wrap_last = d->src_off[last] == 0 && d->boundary[last] == SB200_WRAP;
    const bool chunkable = !has_region && !needs_halo(d) && d->ndim >= 2 && nlast >= 64 &&
                           sb_ >= (size_t)(8u << 20) && d->dst_off[last] == 0 && d->dst_ext[last] == d->size[last];
    if (!chunkable) {
        cudaStream_t st = g_hs.st[0];
        SB_CUDA(cudaMemcpyAsync(g_hs.a, src_host, sb_, cudaMemcpyHostToDevice, st));
        if (needs_halo(d) && (rc = do_halo(d, g_hs.a, st))) return rc;
        if (db != 0 && (d->dst_off[0] | d->dst_off[1] | d->dst_off[2]))  // keep the dest ring as the caller had it
            SB_CUDA(cudaMemcpyAsync(g_hs.b, dst_host, db, cudaMemcpyHostToDevice, st));
        if ((rc = do_gather(d, g_hs.a, g_hs.b, st))) return rc;
        SB_CUDA(cudaMemcpyAsync(dst_host, g_hs.b, db, cudaMemcpyDeviceToHost, st));
        if (needs_halo(d)) SB_CUDA(cudaMemcpyAsync(const_cast<void*>(src_host), g_hs.a, sb_, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        return SB200_OK;
    }
    int nchunks = 16;
    if (const char* e = getenv("SB200_HOST_CHUNKS")) nchunks = std::max(1, std::min(64, atoi(e)));  // tuning knob (events: 2 per chunk)
    if (nlast / nchunks < 2 * (long long)d->radius + 1) nchunks = 1;
    const int R = d->radius, off = d->src_off[last];
    long long sent_hi = 0;  // source planes [0, sent_hi) of the parent are on the device
    const long long src_planes = d->src_ext[last];
    // Wrap on the slowest axis: the first chunk also reads the last R planes (Reflect mirrors into planes the
    // chunk owns anyway, Remove reads padval).
    if (wrap_last && nchunks > 1)
        SB_CUDA(cudaMemcpyAsync((char*)g_hs.a + (src_planes - R) * plane_src, (const char*)src_host + (src_planes - R) * plane_src,
                                (size_t)R * plane_src, cudaMemcpyHostToDevice, g_hs.st[0]));
    for (int c = 0; c < nchunks; c++) {
        const long long lo = nlast * c / nchunks, hi = nlast * (c + 1) / nchunks;
        // source planes this chunk reads: parent planes [lo+off-R, hi+off+R) clipped
        long long need_hi = std::min(src_planes, hi + off + R);
        if (c == nchunks - 1) need_hi = src_planes;
        cudaStream_t sh = g_hs.st[0], sk = g_hs.st[1], sd = g_hs.st[2];
        if (need_hi > sent_hi) {
            SB_CUDA(cudaMemcpyAsync((char*)g_hs.a + sent_hi * plane_src, (const char*)src_host + sent_hi * plane_src,
                                    (size_t)(need_hi - sent_hi) * plane_src, cudaMemcpyHostToDevice, sh));
            sent_hi = need_hi;
        }
        SB_CUDA(cudaEventRecord(g_hs.ev[2 * c], sh));
        SB_CUDA(cudaStreamWaitEvent(sk, g_hs.ev[2 * c], 0));
        sb200_desc cd = *d;
        for (int a = 0; a < d->ndim; a++) { cd.region_lo[a] = 0; cd.region_hi[a] = d->size[a]; }
        cd.region_lo[last] = lo; cd.region_hi[last] = hi;
        if ((rc = do_gather(&cd, g_hs.a, g_hs.b, sk))) return rc;
        SB_CUDA(cudaEventRecord(g_hs.ev[2 * c + 1], sk));
        SB_CUDA(cudaStreamWaitEvent(sd, g_hs.ev[2 * c + 1], 0));
        SB_CUDA(cudaMemcpyAsync((char*)dst_host + lo * plane_dst, (char*)g_hs.b + lo * plane_dst,
                                (size_t)(hi - lo) * plane_dst, cudaMemcpyDeviceToHost, sd));
    }
    SB_CUDA(cudaStreamSynchronize(g_hs.st[2]));
    SB_CUDA(cudaStreamSynchronize(g_hs.st[1]));
    SB_CUDA(cudaStreamSynchronize(g_hs.st[0]));
    return SB200_OK;
}

int32_t sb200_iterate_host(const sb200_desc* d, void* state_host, int32_t nsteps) {
    int rc = validate(d, PK_GATHER);
    if (rc) return rc;
    if (!state_host) { set_error("NULL data pointer"); return SB200_EINVAL; }
    const size_t sb_ = parent_bytes(d, true);
    if ((rc = ensure_scratch(sb_, sb_))) return rc;
    cudaStream_t st = g_hs.st[0];
    SB_CUDA(cudaMemcpyAsync(g_hs.a, state_host, sb_, cudaMemcpyHostToDevice, st));
    if (d->src_off[0] | d->src_off[1] | d->src_off[2]) SB_CUDA(cudaMemcpyAsync(g_hs.b, g_hs.a, sb_, cudaMemcpyDeviceToDevice, st));
    if ((rc = sb200_iterate(d, g_hs.a, g_hs.b, nsteps, st))) return rc;
    SB_CUDA(cudaMemcpyAsync(state_host, (nsteps % 2 == 0) ? g_hs.a : g_hs.b, sb_, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    return SB200_OK;
}

int32_t sb200_shutdown(void) {
    cudaDeviceSynchronize();
    {
        std::lock_guard<std::mutex> lock(g_plan_mu);
        g_evict_epoch.fetch_add(1, std::memory_order_relaxed);
        for (auto& kv : g_plans) free_plan(kv.second);
        g_plans.clear();
    }
    if (g_hs.init) {
        for (int i = 0; i < 3; i++) cudaStreamDestroy(g_hs.st[i]);
        for (int i = 0; i < 128; i++) cudaEventDestroy(g_hs.ev[i]);
        g_hs.init = false;
    }
    if (g_hs.a) cudaFree(g_hs.a);
    if (g_hs.b) cudaFree(g_hs.b);
    g_hs.a = g_hs.b = nullptr; g_hs.a_bytes = g_hs.b_bytes = 0;
    for (auto& kv : g_multi_scratch) if (kv.second.p) cudaFree(kv.second.p);
    g_multi_scratch.clear();
    if (g_done_ctr) cudaFree(g_done_ctr);
    g_done_ctr = nullptr;
    cudaGetLastError();
    return SB200_OK;
}

// ---- memory helpers ----
int32_t sb200_device_count(int32_t* n) { if (!n) return SB200_EINVAL; int c = 0; SB_CUDA(cudaGetDeviceCount(&c)); *n = c; return SB200_OK; }
int32_t sb200_set_device(int32_t dev) { SB_CUDA(cudaSetDevice(dev)); return SB200_OK; }
int32_t sb200_malloc(void** p, size_t bytes) { if (!p) return SB200_EINVAL; SB_CUDA(cudaMalloc(p, bytes)); return SB200_OK; }
int32_t sb200_free(void* p) { SB_CUDA(cudaFree(p)); return SB200_OK; }
int32_t sb200_malloc_host(void** p, size_t bytes) { if (!p) return SB200_EINVAL; SB_CUDA(cudaMallocHost(p, bytes)); return SB200_OK; }
int32_t sb200_free_host(void* p) { SB_CUDA(cudaFreeHost(p)); return SB200_OK; }
int32_t sb200_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) { SB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream)); return SB200_OK; }
int32_t sb200_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) { SB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream)); return SB200_OK; }
int32_t sb200_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) { SB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream)); return SB200_OK; }
int32_t sb200_memset(void* p, int32_t byte, size_t bytes, void* stream) { SB_CUDA(cudaMemsetAsync(p, byte, bytes, (cudaStream_t)stream)); return SB200_OK; }
int32_t sb200_stream_sync(void* stream) { SB_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return SB200_OK; }

// ---- peer memory ----
int32_t sb200_ipc_export(void* p, void* handle64) {
    if (!p || !handle64) return SB200_EINVAL;
    cudaIpcMemHandle_t h;
    SB_CUDA(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return SB200_OK;
}
int32_t sb200_ipc_import(const void* handle64, void** p) {
    if (!p || !handle64) return SB200_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SB_CUDA(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess));
    return SB200_OK;
}
int32_t sb200_ipc_close(void* p) { SB_CUDA(cudaIpcCloseMemHandle(p)); return SB200_OK; }

int32_t sb200_push_planes(const void* src, void* peer_dst, size_t bytes, uint32_t* peer_flag, uint32_t value, void* stream) {
    if (!src || !peer_dst) { set_error("NULL data pointer"); return SB200_EINVAL; }
    if (!g_done_ctr) {
        SB_CUDA(cudaMalloc(&g_done_ctr, 64 * sizeof(unsigned int)));
        SB_CUDA(cudaMemset(g_done_ctr, 0, 64 * sizeof(unsigned int)));
    }
    const bool aligned = (((uintptr_t)src | (uintptr_t)peer_dst) & 15) == 0;
    const size_t n16 = aligned ? bytes / 16 : 0;
    const size_t tail0 = n16 * 16;
    size_t blocks = (std::max<size_t>(n16, 1) + 255) / 256;
    blocks = std::min<size_t>(blocks, (size_t)num_sms() * 4);
    push_planes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)src, (uint4*)peer_dst, n16, (const unsigned char*)src, (unsigned char*)peer_dst, tail0, bytes,
        peer_flag, value, g_done_ctr + (g_push_seq++ % 64));
    SB_LAUNCH_CHECK();
    return SB200_OK;
}

int32_t sb200_signal_flag(uint32_t* flag, uint32_t value, void* stream) {
    if (!flag) return SB200_EINVAL;
    signal_flag_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}

int32_t sb200_wait_flag(const uint32_t* flag, uint32_t value, void* stream) {
    if (!flag) return SB200_EINVAL;
    static const unsigned long long timeout_ns =
        (getenv("SB200_WAIT_TIMEOUT_MS") ? (unsigned long long)std::max(1, atoi(getenv("SB200_WAIT_TIMEOUT_MS"))) : 30000ull) * 1000000ull;
    wait_flag_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value, timeout_ns);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}

}  // extern "C"
