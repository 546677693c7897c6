// slab_plan.cu — sb200_plan_*: slab-partitioned iterated sweeps behind the C ABI.
//
// The reference loop is `A = gatherstencil!(f, A::SwitchingStencilArray)` (src/gatherstencil.jl:77-83) called n times on
// one array in one process; the reference has no multi-device form. A plan runs the same loop over an array split into slabs
// along its last axis. It owns the slab parents (double buffer), the mailboxes (landing slots for the neighbours' ghost
// planes + flag words), the streams and the events, and interprets the schedule of slab_sched.h.
//
// Exchange protocol (both forms): a slab's boundary planes are written into the NEIGHBOUR's landing slot — by the boundary
// sweep itself (sb200_desc.mirror_*, stores over NVLink as the planes are produced) or by a peer copy — then published;
// the receiver waits, copies the slot into the ghost planes of its own parent and (end slabs of a Remove / Reflect axis)
// re-imposes the boundary. Two slots per side are used alternately; no credits are needed: a slab can only publish
// exchange c + 2 after it has received its neighbour's exchange c + 1, which that neighbour published after it had emptied
// slot c (its JOIN / blocking PULL precedes its next boundary sweep on the same stream).
//   single-process form: publish = cudaEventRecord on the sender's stream, wait = cudaStreamWaitEvent (no spinning).
//   rank form: publish = system-scope release store of the exchange number into the neighbour's flag word, wait = an
//              acquire spin with a timeout (a dead neighbour sets the plan's error word instead of hanging the GPU).
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "slab_sched.h"

namespace sb {

constexpr size_t MB_FLAGS = 256;   // bytes reserved for the flag words at the start of a mailbox

// Acquire spin on a flag word in local memory until *flag >= value (wrap-safe) or `timeout_ns` has passed; a timeout is
// recorded in the plan's host-mapped error word and the stream continues (with stale ghosts — the run is void, but the
// GPU is not hung and sb200_plan_sync reports it).
__global__ void plan_wait_kernel(const uint32_t* flag, uint32_t value, unsigned long long timeout_ns, volatile int* err) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t v;
    for (unsigned it = 0;; it++) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - value) >= 0) return;
        __nanosleep(100);
        if ((it & 1023) == 1023) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) { *err = 1; __threadfence_system(); return; }
        }
    }
}

// ---- flag form: the whole exchange of one slab in TWO launches instead of ten stream operations ----
// (measured r02d, Life 16384^2 per GPU, two GPUs: peer copy x2, signal x2, wait x2, ghost copy x2 cost ~43 us per exchange,
// 6 % of a 32-generation cycle)
struct PlanXfer {
    const uint4* src[2];   // [0]: my top owned planes, [1]: my bottom owned planes            (push)   /  the two landing slots (pull)
    uint4* dst[2];         // [0]: upper neighbour's slot (side 0), [1]: lower neighbour's (1) (push)   /  my bottom / top ghost planes (pull)
    uint32_t* flag[2];     // push: the neighbours' flag words; pull: my own two flag words
    size_t n16;            // 16-byte words per zone (zones are whole planes of a 16-byte aligned parent; else the copy-engine path is used)
};

// Copies both boundary zones into the neighbours' landing slots (128-bit stores over NVLink), then the LAST block to finish
// publishes the exchange number to both neighbours with system-scope release stores.
__global__ void __launch_bounds__(256) plan_push_kernel(PlanXfer x, uint32_t value, unsigned int* done_ctr) {
    for (int z = 0; z < 2; z++) {
        if (!x.dst[z]) continue;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < x.n16; i += (size_t)gridDim.x * blockDim.x) x.dst[z][i] = x.src[z][i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done_ctr, 1u);
        if (prev == gridDim.x - 1) {   // every block's stores are fenced
            *done_ctr = 0;
            __threadfence_system();
            for (int z = 0; z < 2; z++)
                if (x.flag[z]) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(x.flag[z]), "r"(value) : "memory");
        }
    }
}

// Waits (acquire, with the plan's deadline) until both neighbours have published `value`, then moves the landing slots into the
// ghost planes. Every block waits by itself; after thread 0 has seen the flag, every thread re-reads it with acquire semantics
// before it touches the slot.
__global__ void __launch_bounds__(256) plan_pull_kernel(PlanXfer x, uint32_t value, unsigned long long timeout_ns, volatile int* err) {
    __shared__ int failed;
    if (threadIdx.x == 0) {
        failed = 0;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int z = 0; z < 2 && !failed; z++) {
            if (!x.flag[z]) continue;
            for (unsigned it = 0;; it++) {
                uint32_t v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(x.flag[z]) : "memory");
                if ((int32_t)(v - value) >= 0) break;
                __nanosleep(100);
                if ((it & 1023) == 1023) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    if (t - t0 > timeout_ns) { *err = 1; __threadfence_system(); failed = 1; break; }
                }
            }
        }
    }
    __syncthreads();
    if (failed) return;
    for (int z = 0; z < 2; z++) {
        if (!x.flag[z]) continue;
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(x.flag[z]) : "memory");
        (void)v;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < x.n16; i += (size_t)gridDim.x * blockDim.x) x.dst[z][i] = x.src[z][i];
    }
}

__global__ void plan_signal2_kernel(uint32_t* f0, uint32_t* f1, uint32_t value) {
    __threadfence_system();
    if (f0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f0), "r"(value) : "memory");
    if (f1) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f1), "r"(value) : "memory");
}

// dst plane (d0 + i) <- src plane (s0 + sstep * i), i = 0 .. nplanes-1, inside one parent (Reflect ends: sstep = -1).
__global__ void plan_planes_kernel(unsigned char* base, size_t plane_bytes, long long d0, long long s0, int sstep, int nplanes) {
    const bool vec = (plane_bytes % 16 == 0) && (((uintptr_t)base & 15) == 0);
    for (int i = blockIdx.y; i < nplanes; i += gridDim.y) {
        unsigned char* d = base + (size_t)(d0 + i) * plane_bytes;
        const unsigned char* s = base + (size_t)(s0 + (long long)sstep * i) * plane_bytes;
        if (vec) {
            const size_t n16 = plane_bytes / 16;
            for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n16; j += (size_t)gridDim.x * blockDim.x)
                reinterpret_cast<uint4*>(d)[j] = reinterpret_cast<const uint4*>(s)[j];
        } else {
            for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < plane_bytes; j += (size_t)gridDim.x * blockDim.x) d[j] = s[j];
        }
    }
}

// Fill `count` elements of `es` bytes with the low bytes of `bits` (Remove ends: ghost planes <- padval).
__global__ void plan_fill_kernel(unsigned char* p, size_t count, int es, unsigned long long bits) {
    for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < count; j += (size_t)gridDim.x * blockDim.x) {
        if (es == 1) p[j] = (unsigned char)bits;
        else if (es == 4) reinterpret_cast<uint32_t*>(p)[j] = (uint32_t)bits;
        else reinterpret_cast<unsigned long long*>(p)[j] = bits;
    }
}

struct Slab {
    int dev = 0;
    long long lo = 0, hi = 0, n = 0, ext = 0;   // owned global planes [lo, hi); parent = [G | n | G]
    void* buf[2] = {nullptr, nullptr};
    void* pk[2] = {nullptr, nullptr};             // packed parents (Life plans that run packed): ext rows of plane_bytes / 8
    int cur = 0;
    unsigned char* mailbox = nullptr;             // MB_FLAGS + 4 landing slots: (side 0 = from below, side 1 = from above) x parity
    unsigned char* peer_down = nullptr;           // the lower neighbour's mailbox (as addressable from this slab's device)
    unsigned char* peer_up = nullptr;
    bool peer_down_ipc = false, peer_up_ipc = false;
    int down = -1, up = -1;                       // neighbour slab index (single-process form) or rank; -1 = none (array end)
    cudaStream_t compute = nullptr, comm = nullptr;
    cudaEvent_t ev_boundary = nullptr, ev_done = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_sent[2] = {nullptr, nullptr};  // single-process form: exchange of parity p published
    uint32_t seq = 0;                             // exchanges published so far
    bool push_published = false;                  // the PUSH just issued also published the exchange (fused kernel): SIGNAL only counts
};

}  // namespace sb

using namespace sb;

struct sb200_plan {
    sb200_desc g;                       // the undivided array (tables owned below)
    std::vector<int32_t> offsets;
    std::vector<unsigned char> weights;
    int ndim = 0, R = 1, G = 1, k = 1;
    size_t plane_bytes = 0, es = 0;
    bool rank_form = false, use_flags = false, overlap = false, split_wrap = true;
    bool fused_xfer = false;            // flag form with 16-byte aligned zones: push + publish / wait + ghost copy as one kernel each
    bool packed = false;                // Life: sweeps and exchanges of a plan_iterate call run on packed parents (one bit per cell)
    bool run_packed = false;            // ... the state the interpreter is working on right now is the packed one
    bool sweep_to_packed = false;       // ... the sweep being issued converts: byte source -> packed dest (first sweep of a call)
    bool sweep_to_bytes = false;        // ... packed source -> byte dest (last sweep of a call)
    int rank = 0, world = 1;
    int later_flags = 0;
    std::vector<Slab> slabs;
    SlabSched* sched = nullptr;
    volatile int* err_host = nullptr;   // host-mapped error word (ghost exchange timed out)
    int* err_dev = nullptr;
    unsigned long long timeout_ns = 30ull * 1000000000ull;
    long long steps = 0, launches = 0, exchanges = 0;
    int max_gens = 1;
    bool connected = false;
};

namespace sb {

static size_t slot_off(const sb200_plan* p, int side, int parity) { return MB_FLAGS + (size_t)(2 * side + parity) * p->plane_bytes * p->G; }

struct DevGuard {
    int prev = 0;
    DevGuard() { cudaGetDevice(&prev); }
    ~DevGuard() { cudaSetDevice(prev); }
};

// descriptor of the sweep of parent planes [lo, hi) (absolute) of a slab with `ext` planes
static sb200_desc sweep_desc(const sb200_plan* p, long long ext, long long lo, long long hi, int gens, bool first) {
    sb200_desc d = p->g;
    const int last = p->ndim - 1;
    d.size[last] = d.src_ext[last] = d.dst_ext[last] = ext;
    d.boundary[last] = SB200_WRAP;   // never exercised: the output region stays R * gens planes inside the parent
    for (int a = 0; a < 3; a++) { d.region_lo[a] = 0; d.region_hi[a] = a < p->ndim ? d.size[a] : 0; }
    d.region_lo[last] = lo; d.region_hi[last] = hi;
    const bool src_bits = p->run_packed, dst_bits = (p->run_packed && !p->sweep_to_bytes) || p->sweep_to_packed;
    d.flags = (first ? 0 : p->later_flags) | SB200_FLAG_GENS(gens) | (src_bits ? SB200_FLAG_SRC_BITS : 0) | (dst_bits ? SB200_FLAG_DST_BITS : 0);
    d.mirror_parent = nullptr; d.mirror_lo = d.mirror_hi = 0;
    d.offsets_host = p->offsets.data();
    d.weights_host = p->weights.empty() ? nullptr : p->weights.data();
    return d;
}

static long long abs_plane(long long v, long long ext) { return v >= 0 ? v : ext + v; }

static void split_last(long long n, int world, int r, long long* lo, long long* hi) {
    const long long base = n / world, rem = n % world;
    *lo = r * base + std::min<long long>(r, rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}

// SB200_POW2_STEPS=1 (or any of the round-2 caps SB200_NO_QUAD_STEP / SB200_OCT_STEP=0 / SB200_NO_DOUBLE_STEP): launches of 8 / 4 / 2
// generations only, as in round 2; otherwise every size the bit-sliced kernel has, the one with the best measured rate first
// (life_bulk_gens(), common.cuh).
static bool life_pow2_only() {
    return (getenv("SB200_POW2_STEPS") && atoi(getenv("SB200_POW2_STEPS")) != 0) || getenv("SB200_NO_QUAD_STEP") || getenv("SB200_NO_DOUBLE_STEP") ||
           (getenv("SB200_OCT_STEP") && atoi(getenv("SB200_OCT_STEP")) == 0);
}

static int plan_common_init(sb200_plan* p, const sb200_desc* g, int ghost, int plan_flags, int nslabs_total) {
    if (!g) { set_error("descriptor is NULL"); return SB200_EINVAL; }
    if (g->struct_size != (int)sizeof(sb200_desc)) { set_error("sb200_desc.struct_size %d != %zu", g->struct_size, sizeof(sb200_desc)); return SB200_EINVAL; }
    if (g->ndim < 2 || g->ndim > 3) { set_error("slab plans need a 2-D or 3-D array (the last axis is split)"); return SB200_EUNSUPPORTED; }
    if (g->eltype != g->out_eltype) { set_error("iterated sweeps need a reducer that preserves the element type"); return SB200_EUNSUPPORTED; }
    if (g->noffsets < 1 || g->noffsets > SB200_MAX_OFFSETS || !g->offsets_host) { set_error("offset table missing"); return SB200_EINVAL; }
    for (int a = 0; a < g->ndim; a++) {
        if (g->src_off[a] || g->dst_off[a] || g->src_ext[a] != g->size[a] || g->dst_ext[a] != g->size[a]) {
            set_error("slab plans take unpadded arrays (Conditional padding): ghost planes are the plan's own ring");
            return SB200_EUNSUPPORTED;
        }
    }
    const int last = g->ndim - 1;
    const int bc = g->boundary[last];
    if (bc != SB200_WRAP && bc != SB200_REMOVE && bc != SB200_REFLECT) { set_error("the split axis needs Wrap, Remove or Reflect"); return SB200_EUNSUPPORTED; }
    p->g = *g;
    p->offsets.assign(g->offsets_host, g->offsets_host + 3 * (size_t)g->noffsets);
    p->es = elsize(g->eltype);
    if (!p->es) { set_error("unknown eltype %d", g->eltype); return SB200_EUNSUPPORTED; }
    if (g->weights_host) p->weights.assign((const unsigned char*)g->weights_host, (const unsigned char*)g->weights_host + p->es * g->noffsets);
    p->ndim = g->ndim;
    p->R = std::max(1, (int)g->radius);
    int G = ghost;
    // Library defaults (measured on 2 GPUs, r02i): Life ~128 ghost rows — an exchange costs ~26 us whatever its size, the rows a
    // wide halo recomputes are 0.8 % of a 16384-row slab; diffusion 4 planes (two double sweeps per cycle; 8 measured the same).
    // Life cycles are whole launches of the kernel's best size b = life_bulk_gens(): the multiples of b next to 128 and 32 (b = 8: 128
    // and 32; b = 7: 126 and 28).
    if (G <= 0) {
        const long long n_est = g->size[g->ndim - 1] / nslabs_total;
        // (a plan that will run packed — the static part of plan_make_sched's decision — is made of packed launches)
        const bool will_pack = g->reducer == SB200_LIFE && bc == SB200_WRAP && g->size[0] % 128 == 0 && !(plan_flags & SB200_PLAN_OVERLAP_ON) &&
                               !(plan_flags & SB200_PLAN_SINGLE_STEP) && !(getenv("SB200_LIFE_PACKED") && atoi(getenv("SB200_LIFE_PACKED")) == 0);
        const int b = life_pow2_only() ? 8 : will_pack ? kLifePackedBulkGens : life_bulk_gens();
        G = g->reducer == SB200_LIFE ? (n_est >= 1024 ? (128 + b / 2) / b * b : (32 + b / 2) / b * b) * p->R
                                     : (g->reducer == SB200_DIFFUSION ? 4 * p->R : p->R);
        while (G > p->R && G > n_est) G /= 2;
        G = std::max(p->R, G / p->R * p->R);
    }
    if (G < p->R || G % p->R) { set_error("ghost thickness must be a positive multiple of the radius"); return SB200_EINVAL; }
    p->G = G;
    p->k = G / p->R;
    p->plane_bytes = p->es;
    for (int a = 0; a < last; a++) p->plane_bytes *= (size_t)g->size[a];
    p->split_wrap = bc == SB200_WRAP;
    const long long n_min = g->size[last] / nslabs_total;
    if (n_min < G) { set_error("slab of %lld planes is thinner than the ghost zone (%d)", n_min, G); return SB200_ESIZE; }
    p->later_flags = (g->reducer == SB200_LIFE && g->eltype == SB200_U8) ? SB200_FLAG_CELLS_01 : 0;
    // Overlap (boundary sweeps first, exchange under the interior sweep) pays for the 3-D planes (ghost zones of MiBs); for 2-D
    // rows three thin multi-generation launches cost more than the exchange they hide (r02i: Life G = 64 0.923 with, G = 32 0.961 without)
    p->overlap = (plan_flags & SB200_PLAN_OVERLAP_ON) ? true : (plan_flags & SB200_PLAN_OVERLAP_OFF) ? false
                 : (g->ndim == 3 && p->plane_bytes * (size_t)G >= ((size_t)1 << 20));
    p->fused_xfer = p->plane_bytes % 16 == 0 && !(getenv("SB200_PLAN_FUSED_XFER") && atoi(getenv("SB200_PLAN_FUSED_XFER")) == 0);
    if (const char* e = getenv("SB200_WAIT_TIMEOUT_MS")) p->timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
    return SB200_OK;
}

static int slab_alloc(sb200_plan* p, Slab& s) {
    SB_CUDA(cudaSetDevice(s.dev));
    s.n = s.hi - s.lo;
    s.ext = s.n + 2 * p->G;
    const size_t bytes = (size_t)s.ext * p->plane_bytes;
    for (int b = 0; b < 2; b++) {
        SB_CUDA(cudaMalloc(&s.buf[b], bytes));
        SB_CUDA(cudaMemset(s.buf[b], 0, bytes));
    }
    const size_t mb = MB_FLAGS + 4 * (size_t)p->G * p->plane_bytes;
    SB_CUDA(cudaMalloc((void**)&s.mailbox, mb));
    SB_CUDA(cudaMemset(s.mailbox, 0, MB_FLAGS));
    SB_CUDA(cudaStreamCreateWithFlags(&s.compute, cudaStreamNonBlocking));
    SB_CUDA(cudaStreamCreateWithFlags(&s.comm, cudaStreamNonBlocking));
    SB_CUDA(cudaEventCreateWithFlags(&s.ev_boundary, cudaEventDisableTiming));
    SB_CUDA(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    SB_CUDA(cudaEventCreate(&s.ev_t0));
    SB_CUDA(cudaEventCreate(&s.ev_t1));
    for (int q = 0; q < 2; q++) SB_CUDA(cudaEventCreateWithFlags(&s.ev_sent[q], cudaEventDisableTiming));
    SB_CUDA(cudaDeviceSynchronize());
    return SB200_OK;
}

// Largest generations-per-launch every sweep of a cycle supports + the scheduler. Needs the slab sizes (all of them in the
// single-process form; the two possible sizes of an even split in the rank form) so that every rank derives the same answer.
static int plan_make_sched(sb200_plan* p, const std::vector<long long>& exts, int plan_flags) {
    SlabSchedCfg c;
    c.R = p->R; c.G = p->G; c.split_wrap = p->split_wrap; c.overlap = p->overlap;
    c.n_min = *std::min_element(exts.begin(), exts.end()) - 2 * p->G;
    int cand = 1;
    if (!(plan_flags & SB200_PLAN_SINGLE_STEP) && p->split_wrap) {   // Remove / Reflect ends are re-imposed after every generation
        if (p->g.reducer == SB200_LIFE && !getenv("SB200_NO_DOUBLE_STEP")) {
            cand = 2;
            if (!getenv("SB200_NO_QUAD_STEP")) cand = 4;
            if (cand == 4 && !(getenv("SB200_OCT_STEP") && atoi(getenv("SB200_OCT_STEP")) == 0)) cand = 8;
        } else if (p->g.reducer == SB200_DIFFUSION) {
            const char* e = getenv("SB200_DIFFUSION_DOUBLE_STEP");
            cand = (e && atoi(e) == 0) ? 1 : 2;
        }
    }
    while (cand > p->k) cand >>= 1;
    const sb200_plan* cp = p;
    const std::vector<long long> ex = exts;
    c.accept = [cp, ex](long long lo, long long hi, int gens) {
        for (long long ext : ex) {
            const long long a = abs_plane(lo, ext), b = abs_plane(hi, ext);
            if (b <= a) continue;
            const sb200_desc d = sweep_desc(cp, ext, a, b, gens, false);
            if (!multistep_accepts(&d)) return false;
        }
        return true;
    };
    // the largest candidate that the full-width first sweep of a cycle accepts
    int mg = 1;
    for (int m = std::max(cand, 1); m > 1; m >>= 1)
        if (c.accept((long long)p->R * m, -(long long)p->R * m, m)) { mg = m; break; }
    if (mg == 8 && p->g.reducer == SB200_LIFE && !life_pow2_only()) {
        // every size the bit-sliced kernel has, best rate first (the scheduler takes the first one that fits the cycle's room)
        std::vector<int> order = {8, 7, 6, 5, 4, 3, 2};
        std::stable_sort(order.begin(), order.end(), [](int a, int b) { return kLifeLaunchCost[a] / a < kLifeLaunchCost[b] / b; });
        for (int m : order)
            if (m <= p->k && c.accept((long long)p->R * m, -(long long)p->R * m, m)) c.sizes.push_back(m);
    }
    // Packed runs: Life on a ring (Remove / Reflect ends would need packed end fills), rows of whole 128-cell groups, and the
    // full-width sweep of every size in use accepted with packed parents. SB200_LIFE_PACKED=0 turns them off.
    p->packed = false;
    if (mg >= 2 && p->g.reducer == SB200_LIFE && p->split_wrap && p->plane_bytes % 16 == 0 && !p->overlap &&
        !(getenv("SB200_LIFE_PACKED") && atoi(getenv("SB200_LIFE_PACKED")) == 0)) {
        p->run_packed = true;   // sweep_desc adds the packed flags
        bool ok = true;
        for (int m = 2; m <= mg && ok; m++)
            if (c.sizes.empty() ? (m & (m - 1)) == 0 : std::find(c.sizes.begin(), c.sizes.end(), m) != c.sizes.end())
                ok = c.accept((long long)p->R * m, -(long long)p->R * m, m);
        p->run_packed = false;
        p->packed = ok;
        if (ok && !c.sizes.empty())   // packed launches have their own best size
            std::stable_sort(c.sizes.begin(), c.sizes.end(), [](int a, int b) { return kLifePackedLaunchCost[a] / a < kLifePackedLaunchCost[b] / b; });
    }
    if (p->packed) {
        DevGuard guard;
        for (Slab& s : p->slabs) {
            SB_CUDA(cudaSetDevice(s.dev));
            const size_t bytes = (size_t)s.ext * p->plane_bytes / 8;
            for (int b = 0; b < 2; b++) {
                SB_CUDA(cudaMalloc(&s.pk[b], bytes));
                SB_CUDA(cudaMemset(s.pk[b], 0, bytes));
            }
        }
    }
    c.max_gens = mg;
    p->max_gens = mg;
    p->sched = new SlabSched(c);
    return SB200_OK;
}

static void slab_free(Slab& s) {
    cudaSetDevice(s.dev);
    if (s.peer_down_ipc && s.peer_down) cudaIpcCloseMemHandle(s.peer_down);
    if (s.peer_up_ipc && s.peer_up && s.peer_up != s.peer_down) cudaIpcCloseMemHandle(s.peer_up);
    for (int b = 0; b < 2; b++) if (s.buf[b]) cudaFree(s.buf[b]);
    for (int b = 0; b < 2; b++) if (s.pk[b]) cudaFree(s.pk[b]);
    if (s.mailbox) cudaFree(s.mailbox);
    if (s.compute) cudaStreamDestroy(s.compute);
    if (s.comm) cudaStreamDestroy(s.comm);
    for (cudaEvent_t e : {s.ev_boundary, s.ev_done, s.ev_t0, s.ev_t1, s.ev_sent[0], s.ev_sent[1]}) if (e) cudaEventDestroy(e);
}

static void plan_free(sb200_plan* p) {
    if (!p) return;
    DevGuard guard;
    for (Slab& s : p->slabs) slab_free(s);
    if (p->err_host) cudaFreeHost((void*)p->err_host);
    delete p->sched;
    delete p;
}

static int plan_error_word(sb200_plan* p) {
    void* h = nullptr;
    SB_CUDA(cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(h, 0, 64);
    p->err_host = (volatile int*)h;
    return SB200_OK;
}

// ---- the interpreter: one op on one slab ----
static int end_fill(sb200_plan* p, Slab& s, void* buf, cudaStream_t st) {
    if (p->split_wrap) return SB200_OK;
    const int G = p->G;
    unsigned char* base = (unsigned char*)buf;
    const bool first = s.down < 0, lastslab = s.up < 0;
    const int bc = p->g.boundary[p->ndim - 1];
    if (bc == SB200_REMOVE) {
        const size_t count = (size_t)G * p->plane_bytes / p->es;
        const unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, (size_t)num_sms() * 8);
        if (first) { plan_fill_kernel<<<blocks, 256, 0, st>>>(base, count, (int)p->es, p->g.padval_bits); SB_LAUNCH_CHECK(); }
        if (lastslab) { plan_fill_kernel<<<blocks, 256, 0, st>>>(base + (size_t)(G + s.n) * p->plane_bytes, count, (int)p->es, p->g.padval_bits); SB_LAUNCH_CHECK(); }
    } else {   // Reflect: i < 0 -> -i ; i >= s -> 2 (s - 1) - i, without repeating the edge (src/array.jl:154-166)
        const unsigned bx = (unsigned)std::min<size_t>((p->plane_bytes / 16 + 255) / 256 + 1, 64);
        const dim3 grid(bx, (unsigned)std::min(G, 64));
        if (first) { plan_planes_kernel<<<grid, 256, 0, st>>>(base, p->plane_bytes, 0, 2 * G, -1, G); SB_LAUNCH_CHECK(); }
        if (lastslab) { plan_planes_kernel<<<grid, 256, 0, st>>>(base, p->plane_bytes, G + s.n, G + s.n - 2, -1, G); SB_LAUNCH_CHECK(); }
    }
    return SB200_OK;
}

static int exec_op(sb200_plan* p, Slab& s, const sb200_slab_op& o) {
    SB_CUDA(cudaSetDevice(s.dev));
    const int G = p->G;
    const size_t pb = p->run_packed ? p->plane_bytes / 8 : p->plane_bytes, gb = pb * (size_t)G;
    void* cur = p->run_packed ? s.pk[s.cur] : s.buf[s.cur];
    void* nxt = p->run_packed ? s.pk[1 - s.cur] : s.buf[1 - s.cur];
    switch (o.kind) {
    case SB200_SLAB_SWEEP: {
        const long long lo = abs_plane(o.lo, s.ext), hi = abs_plane(o.hi, s.ext);
        if (hi <= lo) return SB200_OK;
        sb200_desc d = sweep_desc(p, s.ext, lo, hi, o.gens, o.first != 0);
        const int par = (s.seq + 1) & 1;
        if (o.mirror == SB200_SLAB_MIRROR_DOWN && s.peer_down) {   // my bottom owned planes arrive at the lower neighbour from above: side 1
            d.mirror_parent = s.peer_down + slot_off(p, 1, par); d.mirror_lo = G; d.mirror_hi = 2 * G;
        } else if (o.mirror == SB200_SLAB_MIRROR_UP && s.peer_up) {   // my top owned planes arrive at the upper neighbour from below: side 0
            d.mirror_parent = s.peer_up + slot_off(p, 0, par); d.mirror_lo = s.n; d.mirror_hi = s.n + G;
        }
        // a converting sweep reads one representation and writes the other (same double-buffer index)
        void* from = cur;
        void* to = p->sweep_to_packed ? s.pk[1 - s.cur] : p->sweep_to_bytes ? s.buf[1 - s.cur] : nxt;
        const int rc = do_gather(&d, from, to, s.compute);
        if (rc) return rc;
        p->launches++;
        return SB200_OK;
    }
    case SB200_SLAB_PUSH: {
        void* b = o.buf == SB200_SLAB_CUR ? cur : nxt;
        const int par = (s.seq + 1) & 1;
        if (p->use_flags && p->fused_xfer && (s.peer_up || s.peer_down)) {   // copy + publish in one kernel
            PlanXfer x;
            x.src[0] = (const uint4*)((unsigned char*)b + (size_t)s.n * pb); x.dst[0] = s.peer_up ? (uint4*)(s.peer_up + slot_off(p, 0, par)) : nullptr;
            x.src[1] = (const uint4*)((unsigned char*)b + (size_t)G * pb);   x.dst[1] = s.peer_down ? (uint4*)(s.peer_down + slot_off(p, 1, par)) : nullptr;
            x.flag[0] = s.peer_up ? (uint32_t*)s.peer_up + 0 : nullptr;
            x.flag[1] = s.peer_down ? (uint32_t*)s.peer_down + 1 : nullptr;
            x.n16 = gb / 16;
            const unsigned blocks = (unsigned)std::min<size_t>(std::max<size_t>(x.n16 / 1024, 1), (size_t)num_sms());
            plan_push_kernel<<<blocks, 256, 0, s.compute>>>(x, s.seq + 1, (unsigned int*)(s.mailbox + 128));
            SB_LAUNCH_CHECK();
            s.push_published = true;
            return SB200_OK;
        }
        if (s.peer_up) SB_CUDA(cudaMemcpyAsync(s.peer_up + slot_off(p, 0, par), (unsigned char*)b + (size_t)s.n * pb, gb, cudaMemcpyDefault, s.compute));
        if (s.peer_down) SB_CUDA(cudaMemcpyAsync(s.peer_down + slot_off(p, 1, par), (unsigned char*)b + (size_t)G * pb, gb, cudaMemcpyDefault, s.compute));
        return SB200_OK;
    }
    case SB200_SLAB_SIGNAL: {
        s.seq++;
        if (s.push_published) {
            s.push_published = false;
        } else if (p->use_flags) {
            if (s.peer_up || s.peer_down) {
                plan_signal2_kernel<<<1, 1, 0, s.compute>>>(s.peer_up ? (uint32_t*)s.peer_up + 0 : nullptr, s.peer_down ? (uint32_t*)s.peer_down + 1 : nullptr, s.seq);
                SB_LAUNCH_CHECK();
            }
        } else {
            SB_CUDA(cudaEventRecord(s.ev_sent[s.seq & 1], s.compute));
        }
        if (&s == &p->slabs[0]) p->exchanges++;
        return SB200_OK;
    }
    case SB200_SLAB_PULL: {
        void* b = o.buf == SB200_SLAB_CUR ? cur : nxt;
        cudaStream_t st = o.async ? s.comm : s.compute;
        if (o.async) {
            SB_CUDA(cudaEventRecord(s.ev_boundary, s.compute));
            SB_CUDA(cudaStreamWaitEvent(st, s.ev_boundary, 0));
        }
        const int par = s.seq & 1;
        if (p->use_flags && p->fused_xfer && (s.down >= 0 || s.up >= 0)) {   // wait + ghost copy in one kernel
            PlanXfer x;
            x.src[0] = (const uint4*)(s.mailbox + slot_off(p, 0, par)); x.dst[0] = (uint4*)b;
            x.src[1] = (const uint4*)(s.mailbox + slot_off(p, 1, par)); x.dst[1] = (uint4*)((unsigned char*)b + (size_t)(G + s.n) * pb);
            x.flag[0] = s.down >= 0 ? (uint32_t*)s.mailbox + 0 : nullptr;
            x.flag[1] = s.up >= 0 ? (uint32_t*)s.mailbox + 1 : nullptr;
            x.n16 = gb / 16;
            const unsigned blocks = (unsigned)std::min<size_t>(std::max<size_t>(x.n16 / 1024, 1), (size_t)num_sms());
            plan_pull_kernel<<<blocks, 256, 0, st>>>(x, s.seq, p->timeout_ns, p->err_dev);
            SB_LAUNCH_CHECK();
            const int rc = end_fill(p, s, b, st);
            if (rc) return rc;
            if (o.async) SB_CUDA(cudaEventRecord(s.ev_done, st));
            return SB200_OK;
        }
        if (s.down >= 0) {   // planes from below -> my bottom ghost
            if (p->use_flags) { plan_wait_kernel<<<1, 1, 0, st>>>((const uint32_t*)s.mailbox + 0, s.seq, p->timeout_ns, p->err_dev); SB_LAUNCH_CHECK(); }
            else SB_CUDA(cudaStreamWaitEvent(st, p->slabs[s.down].ev_sent[par], 0));
            SB_CUDA(cudaMemcpyAsync(b, s.mailbox + slot_off(p, 0, par), gb, cudaMemcpyDeviceToDevice, st));
        }
        if (s.up >= 0) {     // planes from above -> my top ghost
            if (p->use_flags) { plan_wait_kernel<<<1, 1, 0, st>>>((const uint32_t*)s.mailbox + 1, s.seq, p->timeout_ns, p->err_dev); SB_LAUNCH_CHECK(); }
            else SB_CUDA(cudaStreamWaitEvent(st, p->slabs[s.up].ev_sent[par], 0));
            SB_CUDA(cudaMemcpyAsync((unsigned char*)b + (size_t)(G + s.n) * pb, s.mailbox + slot_off(p, 1, par), gb, cudaMemcpyDeviceToDevice, st));
        }
        const int rc = end_fill(p, s, b, st);
        if (rc) return rc;
        if (o.async) SB_CUDA(cudaEventRecord(s.ev_done, st));
        return SB200_OK;
    }
    case SB200_SLAB_JOIN:
        SB_CUDA(cudaStreamWaitEvent(s.compute, s.ev_done, 0));
        return SB200_OK;
    case SB200_SLAB_ENDFILL:
        return end_fill(p, s, o.buf == SB200_SLAB_CUR ? cur : nxt, s.compute);
    case SB200_SLAB_SWAP:
        s.cur = 1 - s.cur;
        return SB200_OK;
    default:
        set_error("unknown slab op %d", o.kind);
        return SB200_EINVAL;
    }
}

static int plan_run(sb200_plan* p, int nsteps) {
    if (nsteps < 0) { set_error("negative step count"); return SB200_EINVAL; }
    if (p->rank_form && p->world > 1 && !p->connected) { set_error("sb200_plan_connect has not been called"); return SB200_EINVAL; }
    std::vector<sb200_slab_op> ops;
    p->run_packed = p->packed;   // the scheduler's acceptance probes see the descriptors most sweeps will use
    p->sched->plan(nsteps, ops);
    p->run_packed = false;
    DevGuard guard;
    // Packed calls: the first sweep of >= 2 generations reads the byte parent and writes the packed one, every op after it works on
    // the packed state (sweeps, pushes, pulls: rows of plane_bytes / 8), and the LAST sweep of the call writes the byte parent
    // again — no conversion passes. Sweeps in front of the converting one (single generations) stay byte -> byte; a call whose
    // sweeps leave no room for both conversions runs on the bytes.
    int conv_in = -1, conv_out = -1;
    if (p->packed) {
        for (int i = 0; i < (int)ops.size(); i++)
            if (ops[i].kind == SB200_SLAB_SWEEP) {
                if (conv_in < 0 && ops[i].gens >= 2) conv_in = i;
                conv_out = i;
            }
        // boundary-first cycles issue three sweeps per step (overlap): packed plans have overlap off, so sweeps and steps coincide
        if (conv_in < 0 || conv_out <= conv_in) conv_in = conv_out = -1;
    }
    int rc = SB200_OK;
    for (int i = 0; i < (int)ops.size() && !rc; i++) {
        const sb200_slab_op& o = ops[i];
        p->sweep_to_packed = i == conv_in;
        p->sweep_to_bytes = i == conv_out;
        // op by op over all slabs: in the event-ordered form a slab's PULL must be enqueued after its neighbours' SIGNAL
        for (Slab& s : p->slabs)
            if ((rc = exec_op(p, s, o))) break;
        if (i == conv_in) p->run_packed = true;     // from here on the current state is the packed one
        if (i == conv_out) p->run_packed = false;
    }
    p->run_packed = p->sweep_to_packed = p->sweep_to_bytes = false;
    if (rc) return rc;
    p->steps += nsteps;
    return SB200_OK;
}

}  // namespace sb

extern "C" {

int32_t sb200_plan_create(const sb200_desc* global, int32_t nslabs, const int32_t* devices, int32_t ghost, int32_t plan_flags,
                          sb200_plan** out) {
    if (!out) { set_error("NULL out"); return SB200_EINVAL; }
    *out = nullptr;
    if (nslabs < 1 || nslabs > 64) { set_error("1..64 slabs"); return SB200_EINVAL; }
    int ndev = 0;
    SB_CUDA(cudaGetDeviceCount(&ndev));
    sb200_plan* p = new sb200_plan();
    int rc = plan_common_init(p, global, ghost, plan_flags, nslabs);
    if (rc) { delete p; return rc; }
    p->rank_form = false;
    p->use_flags = (plan_flags & SB200_PLAN_FLAGS_SYNC) != 0;
    p->world = nslabs;
    DevGuard guard;
    if ((rc = plan_error_word(p))) { plan_free(p); return rc; }
    p->slabs.resize(nslabs);
    const int last = p->ndim - 1;
    std::vector<long long> exts;
    for (int i = 0; i < nslabs; i++) {
        Slab& s = p->slabs[i];
        s.dev = devices ? devices[i] : i % ndev;
        if (s.dev < 0 || s.dev >= ndev) { set_error("device %d of slab %d does not exist (%d devices)", s.dev, i, ndev); plan_free(p); return SB200_EINVAL; }
        split_last(global->size[last], nslabs, i, &s.lo, &s.hi);
        s.up = (p->split_wrap || i < nslabs - 1) ? (i + 1) % nslabs : -1;
        s.down = (p->split_wrap || i > 0) ? (i + nslabs - 1) % nslabs : -1;
        if ((rc = slab_alloc(p, s))) { plan_free(p); return rc; }
        exts.push_back(s.ext);
    }
    // peer access between the devices of neighbouring slabs (the mirror stores and peer copies go over NVLink)
    for (int i = 0; i < nslabs; i++) {
        Slab& s = p->slabs[i];
        for (int nb : {s.up, s.down}) {
            if (nb < 0 || p->slabs[nb].dev == s.dev) continue;
            int can = 0;
            SB_CUDA(cudaDeviceCanAccessPeer(&can, s.dev, p->slabs[nb].dev));
            if (!can) { set_error("device %d cannot access device %d (no peer access)", s.dev, p->slabs[nb].dev); plan_free(p); return SB200_EUNSUPPORTED; }
            cudaSetDevice(s.dev);
            const cudaError_t e = cudaDeviceEnablePeerAccess(p->slabs[nb].dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); plan_free(p); return SB200_ECUDA; }
            cudaGetLastError();
        }
        s.peer_up = s.up >= 0 ? p->slabs[s.up].mailbox : nullptr;
        s.peer_down = s.down >= 0 ? p->slabs[s.down].mailbox : nullptr;
    }
    {   // device alias of the error word (portable mapped host memory: one pointer under UVA)
        void* dptr = nullptr;
        cudaSetDevice(p->slabs[0].dev);
        if (cudaHostGetDevicePointer(&dptr, (void*)p->err_host, 0) != cudaSuccess) { set_error("cudaHostGetDevicePointer failed"); plan_free(p); return SB200_ECUDA; }
        p->err_dev = (int*)dptr;
    }
    cudaSetDevice(p->slabs[0].dev);
    if ((rc = plan_make_sched(p, exts, plan_flags))) { plan_free(p); return rc; }
    p->connected = true;
    *out = p;
    return SB200_OK;
}

int32_t sb200_plan_create_rank(const sb200_desc* global, int32_t rank, int32_t world, int32_t ghost, int32_t plan_flags, sb200_plan** out) {
    if (!out) { set_error("NULL out"); return SB200_EINVAL; }
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) { set_error("bad rank %d / world %d", rank, world); return SB200_EINVAL; }
    sb200_plan* p = new sb200_plan();
    int rc = plan_common_init(p, global, ghost, plan_flags, world);
    if (rc) { delete p; return rc; }
    p->rank_form = true;
    p->use_flags = true;
    p->rank = rank; p->world = world;
    if ((rc = plan_error_word(p))) { plan_free(p); return rc; }
    p->slabs.resize(1);
    Slab& s = p->slabs[0];
    SB_CUDA(cudaGetDevice(&s.dev));
    const int last = p->ndim - 1;
    split_last(global->size[last], world, rank, &s.lo, &s.hi);
    s.up = (p->split_wrap || rank < world - 1) ? (rank + 1) % world : -1;
    s.down = (p->split_wrap || rank > 0) ? (rank + world - 1) % world : -1;
    if ((rc = slab_alloc(p, s))) { plan_free(p); return rc; }
    void* dptr = nullptr;
    if (cudaHostGetDevicePointer(&dptr, (void*)p->err_host, 0) != cudaSuccess) { set_error("cudaHostGetDevicePointer failed"); plan_free(p); return SB200_ECUDA; }
    p->err_dev = (int*)dptr;
    // every rank must derive the same schedule: probe both slab sizes of the even split
    std::vector<long long> exts;
    const long long base = global->size[last] / world;
    exts.push_back(base + 2 * p->G);
    if (global->size[last] % world) exts.push_back(base + 1 + 2 * p->G);
    if ((rc = plan_make_sched(p, exts, plan_flags))) { plan_free(p); return rc; }
    if (world == 1) {   // my own neighbour (ring) or none
        s.peer_up = s.up >= 0 ? s.mailbox : nullptr;
        s.peer_down = s.down >= 0 ? s.mailbox : nullptr;
        p->connected = true;
    }
    *out = p;
    return SB200_OK;
}

int32_t sb200_plan_ipc_handle(sb200_plan* p, void* handle64) {
    if (!p || !handle64 || !p->rank_form) { set_error("sb200_plan_ipc_handle: rank-form plan and a 64-byte buffer required"); return SB200_EINVAL; }
    cudaIpcMemHandle_t h;
    SB_CUDA(cudaIpcGetMemHandle(&h, p->slabs[0].mailbox));
    memcpy(handle64, &h, 64);
    return SB200_OK;
}

int32_t sb200_plan_connect(sb200_plan* p, const void* handles) {
    if (!p || !handles || !p->rank_form) { set_error("sb200_plan_connect: rank-form plan and world x 64 bytes of handles required"); return SB200_EINVAL; }
    if (p->connected) return SB200_OK;
    Slab& s = p->slabs[0];
    DevGuard guard;
    SB_CUDA(cudaSetDevice(s.dev));
    auto open = [&](int r, unsigned char** outp, bool* ipc) -> int {
        if (r == p->rank) { *outp = s.mailbox; *ipc = false; return SB200_OK; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)handles + 64 * (size_t)r, 64);
        void* ptr = nullptr;
        SB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        *outp = (unsigned char*)ptr; *ipc = true;
        return SB200_OK;
    };
    int rc;
    if (s.down >= 0 && (rc = open(s.down, &s.peer_down, &s.peer_down_ipc))) return rc;
    if (s.up >= 0) {
        if (s.up == s.down && s.peer_down) { s.peer_up = s.peer_down; s.peer_up_ipc = false; }   // two ranks on a ring: one mapping
        else if ((rc = open(s.up, &s.peer_up, &s.peer_up_ipc))) return rc;
    }
    p->connected = true;
    return SB200_OK;
}

int32_t sb200_plan_nslabs(const sb200_plan* p, int32_t* n) {
    if (!p || !n) return SB200_EINVAL;
    *n = (int32_t)p->slabs.size();
    return SB200_OK;
}

int32_t sb200_plan_slab(sb200_plan* p, int32_t i, int64_t* lo, int64_t* hi, int32_t* device, void** owned) {
    if (!p || i < 0 || i >= (int)p->slabs.size()) { set_error("no such slab"); return SB200_EINVAL; }
    const Slab& s = p->slabs[i];
    if (lo) *lo = s.lo;
    if (hi) *hi = s.hi;
    if (device) *device = s.dev;
    if (owned) *owned = (unsigned char*)s.buf[s.cur] + (size_t)p->G * p->plane_bytes;
    return SB200_OK;
}

static int plan_copy_host(sb200_plan* p, void* host, bool load) {
    if (!p || !host) { set_error("NULL plan / pointer"); return SB200_EINVAL; }
    DevGuard guard;
    const long long lo0 = p->slabs[0].lo;
    for (Slab& s : p->slabs) {
        SB_CUDA(cudaSetDevice(s.dev));
        SB_CUDA(cudaStreamSynchronize(s.compute));
        unsigned char* dev = (unsigned char*)s.buf[s.cur] + (size_t)p->G * p->plane_bytes;
        unsigned char* h = (unsigned char*)host + (size_t)(s.lo - lo0) * p->plane_bytes;
        const size_t bytes = (size_t)s.n * p->plane_bytes;
        if (load) SB_CUDA(cudaMemcpyAsync(dev, h, bytes, cudaMemcpyHostToDevice, s.compute));
        else SB_CUDA(cudaMemcpyAsync(h, dev, bytes, cudaMemcpyDeviceToHost, s.compute));
    }
    for (Slab& s : p->slabs) {
        SB_CUDA(cudaSetDevice(s.dev));
        SB_CUDA(cudaStreamSynchronize(s.compute));
    }
    if (load) { p->sched->since = p->k; p->sched->nsweeps = 0; }   // ghosts are stale for the new state; cells not known to be 0/1
    return SB200_OK;
}
int32_t sb200_plan_load_host(sb200_plan* p, const void* state_host) { return plan_copy_host(p, const_cast<void*>(state_host), true); }
int32_t sb200_plan_store_host(sb200_plan* p, void* state_host) { return plan_copy_host(p, state_host, false); }

int32_t sb200_plan_mark_dirty(sb200_plan* p) {
    if (!p) { set_error("NULL plan"); return SB200_EINVAL; }
    p->sched->since = p->k;
    p->sched->nsweeps = 0;
    return SB200_OK;
}

int32_t sb200_plan_iterate(sb200_plan* p, int32_t nsteps) {
    if (!p) { set_error("NULL plan"); return SB200_EINVAL; }
    return plan_run(p, nsteps);
}

int32_t sb200_plan_sync(sb200_plan* p) {
    if (!p) { set_error("NULL plan"); return SB200_EINVAL; }
    DevGuard guard;
    for (Slab& s : p->slabs) {
        SB_CUDA(cudaSetDevice(s.dev));
        SB_CUDA(cudaStreamSynchronize(s.compute));
        SB_CUDA(cudaStreamSynchronize(s.comm));
    }
    if (p->err_host && *p->err_host) {
        set_error("ghost exchange timed out after %llu ms: a neighbouring rank did not publish its planes (dead or desynchronised peer); the state is invalid",
                  p->timeout_ns / 1000000ull);
        return SB200_ECUDA;
    }
    return SB200_OK;
}

int32_t sb200_plan_iterate_timed(sb200_plan* p, int32_t nsteps, float* ms) {
    if (!p || !ms) { set_error("NULL plan / out"); return SB200_EINVAL; }
    DevGuard guard;
    int rc;
    for (Slab& s : p->slabs) { SB_CUDA(cudaSetDevice(s.dev)); SB_CUDA(cudaEventRecord(s.ev_t0, s.compute)); }
    if ((rc = plan_run(p, nsteps))) return rc;
    for (Slab& s : p->slabs) { SB_CUDA(cudaSetDevice(s.dev)); SB_CUDA(cudaEventRecord(s.ev_t1, s.compute)); }
    if ((rc = sb200_plan_sync(p))) return rc;
    float worst = 0.f;
    for (Slab& s : p->slabs) {
        float t = 0.f;
        SB_CUDA(cudaSetDevice(s.dev));
        SB_CUDA(cudaEventElapsedTime(&t, s.ev_t0, s.ev_t1));
        worst = std::max(worst, t);
    }
    *ms = worst;
    return SB200_OK;
}

int32_t sb200_plan_stats(const sb200_plan* p, int64_t out[8]) {
    if (!p || !out) return SB200_EINVAL;
    out[0] = p->steps; out[1] = p->launches; out[2] = p->exchanges; out[3] = p->G; out[4] = p->k;
    out[5] = p->overlap ? 1 : 0; out[6] = p->max_gens; out[7] = p->use_flags ? 2 : 1;
    return SB200_OK;
}

int32_t sb200_plan_destroy(sb200_plan* p) {
    if (!p) return SB200_OK;
    DevGuard guard;
    for (Slab& s : p->slabs) {
        cudaSetDevice(s.dev);
        cudaStreamSynchronize(s.compute);
        cudaStreamSynchronize(s.comm);
    }
    plan_free(p);
    return SB200_OK;
}

int32_t sb200_slab_schedule(int32_t radius, int32_t ghost, int64_t n_min, int32_t split_wrap, int32_t overlap, int32_t max_gens,
                            int32_t min_planes_multi, int32_t since, int32_t first_sweep, int32_t nsteps, sb200_slab_op* ops, int32_t cap,
                            int32_t* count, int32_t* since_out) {
    if (radius < 1 || ghost < radius || ghost % radius || nsteps < 0 || !count) { set_error("bad arguments to sb200_slab_schedule"); return SB200_EINVAL; }
    SlabSchedCfg c;
    c.R = radius; c.G = ghost; c.split_wrap = split_wrap != 0; c.overlap = overlap != 0; c.max_gens = std::max(1, (int)max_gens);
    if (max_gens < 0) {   // -mask: bit g set = launches of g generations, preference order of the Life plans
        c.max_gens = 1;
        for (int m : {7, 8, 6, 5, 4, 3, 2})
            if ((-max_gens >> m) & 1) { c.sizes.push_back(m); c.max_gens = std::max(c.max_gens, m); }
    }
    c.n_min = n_min;
    const long long ext = n_min + 2 * (long long)ghost;
    c.accept = [ext, min_planes_multi](long long lo, long long hi, int) { return abs_plane(hi, ext) - abs_plane(lo, ext) >= min_planes_multi; };
    SlabSched s(c);
    s.since = since;
    s.nsweeps = first_sweep ? 0 : 1;
    std::vector<sb200_slab_op> v;
    s.plan(nsteps, v);
    *count = (int32_t)v.size();
    if (ops)
        for (int i = 0; i < (int)v.size() && i < cap; i++) ops[i] = v[i];
    if (since_out) *since_out = s.since;
    return SB200_OK;
}

}  // extern "C"
