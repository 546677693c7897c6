/*
 * stencils_b200.h — C ABI of libstencils_b200.so
 *
 * Drop-in boundary for the per-cell sweep of rafaqz/Stencils.jl (reference v0.3.6).
 * Every entry point replaces one mutating Julia entry point of the reference; the
 * allocating wrappers (`gatherstencil`/`mapstencil`) stay on the host side (Julia shim
 * `julia/StencilsB200.jl`, Python mirror `stencils.jl_b200/`).
 *
 *   sb200_gather       <- gatherstencil!(f, dest, source, args...)      src/gatherstencil.jl:89-103
 *                         + gatherstencil_kernel!                        src/gatherstencil.jl:105-109
 *                         + stencil/neighbors/getneighbor/bounded_index  src/array.jl:22-30,68-188
 *   sb200_update_halo  <- update_boundary!(A)                            src/array.jl:195-239
 *   sb200_scatter      <- scatterstencil!(f, op, dest, source)           src/scatterstencil.jl:36-112
 *   sb200_iterate      <- loop of gatherstencil!(f, A::SwitchingStencilArray) + switch(A)
 *                                                                        src/gatherstencil.jl:77-83, src/array.jl:610-611
 *   sb200_plan_*       <- the same loop over an array split into slabs across the GPUs of one box
 *   sb200_gather_multi <- gatherstencil!(f, dest, A1, A2, ...) extra-array forms   src/gatherstencil.jl:84-88,112-113
 *   sb200_stencil_offsets <- offsets(::Type{<:Stencil})                  src/stencils/{*}.jl
 *   sb200_out_eltype   <- _return_type                                   src/gatherstencil.jl:41-59
 *
 * Conventions
 *   - Arrays are Julia column-major: axis 0 here is Julia dim 1 (contiguous). Offsets are
 *     (o0,o1,o2) with o0 along the contiguous axis. Indices in this ABI are 0-based.
 *   - All data pointers are DEVICE pointers unless the name ends in `_host`.
 *   - Calls are stream-ordered and non-blocking; the caller keeps every buffer alive until
 *     the stream has drained. `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - No exceptions cross the ABI. Every function returns an sb200_status; the thread-local
 *     message is read with sb200_last_error().
 *   - An unsupported reducer / eltype / combination returns SB200_EUNSUPPORTED. There is no
 *     CPU fallback anywhere in this library.
 */
#ifndef STENCILS_B200_H
#define STENCILS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_VERSION 100 /* 0.1.0 */
#define SB200_MAX_DIMS 3
#define SB200_MAX_OFFSETS 1024

typedef enum sb200_status {
    SB200_OK = 0,
    SB200_EINVAL = 1,       /* malformed descriptor / NULL pointer */
    SB200_EUNSUPPORTED = 2, /* reducer/eltype/shape not implemented: shim raises ArgumentError */
    SB200_ESIZE = 3,        /* _checksizes (src/gatherstencil.jl:118-124), radius-vs-axis (src/array.jl:451-453) */
    SB200_ECUDA = 4,        /* CUDA runtime/driver error, message holds cudaGetErrorString */
    SB200_ENOMEM = 5
} sb200_status;

/* Element types (Julia names). Bool is one byte holding 0/1. */
typedef enum sb200_eltype {
    SB200_BOOL = 0,
    SB200_U8 = 1,
    SB200_I32 = 2,
    SB200_I64 = 3,
    SB200_F32 = 4,
    SB200_F64 = 5
} sb200_eltype;

/* Boundary conditions, src/boundary.jl:17,27-32,42,52 */
typedef enum sb200_boundary {
    SB200_REMOVE = 0, /* out-of-bounds neighbours read `padval` */
    SB200_WRAP = 1,   /* single wrap: i<0 -> i+s, i>=s -> i-s   (src/array.jl:167-179) */
    SB200_REFLECT = 2,/* mirror without repeating the edge       (src/array.jl:154-166) */
    SB200_USE = 3     /* read the existing halo ring as it is (needs src_off >= radius) */
} sb200_boundary;

/* Named stencil shapes, src/stencils/{window,moore,vonneumman,shapes}.jl */
typedef enum sb200_shape {
    SB200_WINDOW = 0,
    SB200_MOORE = 1,
    SB200_VONNEUMANN = 2,
    SB200_CROSS = 3,
    SB200_ANGLEDCROSS = 4,
    SB200_FORWARDSLASH = 5,
    SB200_BACKSLASH = 6,
    SB200_CIRCLE = 7,
    SB200_VERTICAL = 8,
    SB200_HORIZONTAL = 9,
    SB200_DIAMOND = 10,
    SB200_ANNULUS = 11, /* uses inner_radius */
    SB200_CARDINAL = 12,
    SB200_ORDINAL = 13
    /* Positional / NamedStencil / Rectangle pass their offset table directly. */
} sb200_shape;

/* The function `f` applied to the filled stencil (fixed menu; anything else is unsupported). */
typedef enum sb200_reducer {
    SB200_SUM = 0,       /* sum(hood): strict left fold in offset order, seeded by the first value */
    SB200_MEAN = 1,      /* sum(hood) / L */
    SB200_MIN = 2,       /* minimum(hood): Julia min (NaN-propagating, -0.0 < +0.0) */
    SB200_MAX = 3,       /* maximum(hood) */
    SB200_KERNELDOT = 4, /* kernelproduct, src/stencils/kernel.jl:34-43: acc=0; acc += v_k*w_k, unfused */
    SB200_LIFE = 5,      /* s = sum of (neighbour != 0); out = ((centre!=0 ? survive : born) >> s) & 1 */
    SB200_DIFFUSION = 6  /* centre + alpha*(sum(hood) - L*centre), every operation rounded separately */
} sb200_reducer;

/* `op` of scatterstencil!, src/scatterstencil.jl:76-112 */
typedef enum sb200_scatter_op { SB200_OP_ADD = 0, SB200_OP_MAX = 1, SB200_OP_MIN = 2 } sb200_scatter_op;

/* Fixed menu for the user function of scatterstencil! (value sent to offset k). */
typedef enum sb200_scatter_rule {
    SB200_SCATTER_WEIGHTS = 0,        /* val_k = w_k                 (test/array.jl:391-393) */
    SB200_SCATTER_CENTER_WEIGHTS = 1  /* val_k = centre * w_k        (test/array.jl:412-415 with w=1) */
} sb200_scatter_rule;

/*
 * Sweep descriptor. Plain data, filled by the host shim from a StencilArray
 * (src/array.jl:441-468): parent array, stencil, boundary, padding.
 *
 * Per axis a (0 = contiguous):
 *   size[a]     logical size of the StencilArray (`size(A)`, src/array.jl:379,470-473)
 *   src_ext[a]  extent of the source parent array            (= size + 2R for Halo padding)
 *   src_off[a]  parent index of logical index 0              (= R for Halo: add_halo, src/array.jl:367-377; else 0)
 *   dst_ext/dst_off the same for the destination (plain dest array: ext=size, off=0;
 *               Switching+Halo dest: the padded twin buffer, src/gatherstencil.jl:77-83)
 *   boundary[a] boundary condition on this axis. The reference has one condition for all axes;
 *               the per-axis form is what the slab decomposition needs (ghost planes = USE on the
 *               split axis). An axis with src_off>0 is read straight through its ring (Halo
 *               semantics, src/array.jl:91-99); an axis with src_off==0 resolves out-of-bounds
 *               neighbours on the fly (Conditional semantics, src/array.jl:101-138).
 */
typedef struct sb200_desc {
    int32_t struct_size; /* = sizeof(sb200_desc) */
    int32_t ndim;        /* array dimensionality 1..3 */
    int64_t size[SB200_MAX_DIMS];
    int64_t src_ext[SB200_MAX_DIMS];
    int64_t dst_ext[SB200_MAX_DIMS];
    int32_t src_off[SB200_MAX_DIMS];
    int32_t dst_off[SB200_MAX_DIMS];
    int32_t boundary[SB200_MAX_DIMS];
    int32_t eltype;      /* sb200_eltype of the source */
    int32_t out_eltype;  /* sb200_eltype of the destination; must equal sb200_out_eltype() */
    uint64_t padval_bits;/* Remove(padval): raw bits of one `eltype` value in the low bytes */
    int32_t radius;      /* stencil radius R = max |offset| */
    int32_t noffsets;    /* L */
    const int32_t* offsets_host; /* [L][3] ordered offset table (host memory), unused axes = 0 */
    int32_t reducer;     /* sb200_reducer */
    int32_t scatter_op;  /* sb200_scatter_op   (sb200_scatter only) */
    int32_t scatter_rule;/* sb200_scatter_rule (sb200_scatter only) */
    uint32_t born_mask;  /* LIFE: bit s set -> dead cell with s live neighbours is born   (B3 = 1<<3) */
    uint32_t survive_mask;/* LIFE: bit s set -> live cell with s live neighbours survives (S23 = 0b1100) */
    int32_t reserved0;
    const void* weights_host; /* KERNELDOT / scatter: L values of `eltype` (host memory) */
    double alpha;        /* DIFFUSION coefficient, converted to `eltype` before use */
    /* Output sub-range [region_lo, region_hi) in logical coordinates; all-zero = whole array.
       Used for boundary-first / interior overlap in the slab iterator and for chunked host sweeps. */
    int64_t region_lo[SB200_MAX_DIMS];
    int64_t region_hi[SB200_MAX_DIMS];
    int32_t flags;       /* SB200_FLAG_* */
    int32_t reserved1;
    /* Fused ghost-plane push (slab-partitioned runs): output planes [mirror_lo, mirror_hi) of the LAST axis (logical
       coordinates) are ALSO stored, by the sweep kernel itself, to `mirror_parent` — plane mirror_lo first, planes
       packed with the dest parent's plane pitch. `mirror_parent` is normally a neighbour GPU's landing slot opened with
       sb200_ipc_import, so the boundary planes cross NVLink as they are produced; publish them with sb200_signal_flag.
       NULL = off. Kernels without the fused store fall back to a stream-ordered copy inside the call. */
    void* mirror_parent;
    int64_t mirror_lo;
    int64_t mirror_hi;
} sb200_desc;

#define SB200_FLAG_FORCE_GENERIC 1 /* bypass specialised kernels (testing: generic vs fast parity) */
#define SB200_FLAG_ZERO_DEST 2     /* scatter: treat dest as zero-filled (Switching forms, src/scatterstencil.jl:119,130) */
#define SB200_FLAG_NO_TMA 4        /* testing: use the non-TMA variant of a specialised kernel */
#define SB200_FLAG_CELLS_01 8      /* LIFE on UInt8: the caller guarantees every source cell is 0 or 1 */
#define SB200_FLAG_ALLOW_FMA 128  /* KERNELDOT: the caller allows acc = fma(v_k, w_k, acc) (one rounding per tap) instead of the
                                     reference's separately rounded multiply and add (src/stencils/kernel.jl:37-43). A permission,
                                     not a request: kernels without a contracted variant ignore it and stay bit-exact. Honoured by the
                                     streaming Window(1..3) Float32 / Float64 kernels (7 x 7 Float32: FP32-issue bound, ~2x faster) */
#define SB200_FLAG_QUAD_STEP 32   /* dest = f(f(f(f(src)))): four generations per launch (B3/S23 Life, axis 0 a multiple of 32
                                     cells, otherwise as SB200_FLAG_DOUBLE_STEP); the bit-sliced kernel */
#define SB200_FLAG_OCT_STEP 64    /* eight generations per launch (as SB200_FLAG_QUAD_STEP; a library built with
                                     -DSB200_LB_ONE_HALO_LANE=0 answers SB200_EUNSUPPORTED) */
#define SB200_FLAG_DOUBLE_STEP 16 /* dest = f(f(src)): two sweeps fused in one launch, the intermediate state never touches
                                     memory. LIFE: Moore(1), unpadded, Wrap on axis 0. DIFFUSION: VonNeumann(1,3), unpadded
                                     Float32 / Float64, Wrap on axes 0 and 1, axis 2 Wrap or an output region two planes
                                     inside the parent. Anything else: SB200_EUNSUPPORTED (never a silent single sweep).
                                     sb200_iterate uses it by itself where it applies. */
/* The three *_STEP bits together are a 3-bit field (flags >> 4) & 7 that names the generations of one launch: 1, 2 and 4 are the
   flags above (2, 4, 8 generations), and the combinations 3, 5, 6, 7 mean that many generations (B3/S23 Life through the bit-sliced
   kernel only; everything else answers SB200_EUNSUPPORTED). sb200_iterate uses them to split a step count into equal launches with
   the launch-count parity the buffer contract needs (20 steps = 5 + 5 + 5 + 5 instead of 8 + 8 + 2 + 2). */
#define SB200_FLAG_SRC_BITS 256  /* LIFE with SB200_FLAG_GENS(2 .. 8) (a packed source: 1 .. 8), B3/S23, axis 0 a multiple of 128 cells: the SOURCE parent holds one BIT per
                                     cell — row r starts at byte r * src_ext[0] / 8 and cell c of the row is bit c % 8 of its byte c / 8 (the
                                     layout of a Julia BitMatrix whose first dimension is a multiple of 64). eltype / out_eltype keep
                                     naming the unpacked type (UInt8 / Bool), the extents stay in cells. */
#define SB200_FLAG_DST_BITS 512  /* the same for the DEST parent. sb200_iterate and the slab plans keep the state of a Life run packed
                                     between its first and its last launch by themselves (pack and unpack are half of the instructions of
                                     a byte-to-byte launch); a binding that keeps a BitMatrix can use the flags directly. */
#define SB200_FLAG_STEP_MASK 112
#define SB200_FLAG_GENS(n) ((n) == 2 ? 16 : (n) == 4 ? 32 : (n) == 8 ? 64 : ((n) == 3 || ((n) >= 5 && (n) <= 7)) ? ((n) << 4) : 0)
#define SB200_FLAG_GENS_OF(flags) ((((flags) >> 4) & 7) == 0 ? 1 : (((flags) >> 4) & 7) == 1 ? 2 : (((flags) >> 4) & 7) == 2 ? 4 : \
                                   (((flags) >> 4) & 7) == 4 ? 8 : (((flags) >> 4) & 7))

/* ---- library ---- */
int32_t sb200_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* sb200_last_error(void);
/* Name of the kernel variant the last successful sweep on this thread dispatched to (diagnostics). */
const char* sb200_last_kernel(void);
/* Number of kernel launches issued by this library on this thread since the last reset. */
int64_t sb200_launch_count(int32_t reset);

/* ---- stencil algebra (host only) ---- */
/* Ordered offsets of a named shape: box (-R:R)^ndim iterated with axis 0 fastest, filtered by the shape
   predicate (src/stencils/{*}.jl). Writes up to cap triples; *count receives L. */
int32_t sb200_stencil_offsets(int32_t shape, int32_t radius, int32_t inner_radius, int32_t ndim,
                              int32_t* out, int32_t cap, int32_t* count);
/* Result element type of reducer applied to a stencil of `eltype` (src/gatherstencil.jl:41-59). */
int32_t sb200_out_eltype(int32_t reducer, int32_t eltype, int32_t* out);
size_t sb200_sizeof(int32_t eltype);

/* ---- sweeps (device pointers) ---- */
int32_t sb200_gather(const sb200_desc* d, const void* src_parent, void* dst_parent, void* stream);
int32_t sb200_update_halo(const sb200_desc* d, void* src_parent, void* stream);
int32_t sb200_scatter(const sb200_desc* d, const void* src_parent, void* dst_parent, void* stream);
/*
 * Multi-array gather: gatherstencil!(f, dest, A1, A2, ...) (src/gatherstencil.jl:84-88, 95, 107, 112-113; reference
 * tests test/array.jl:312-383) for user functions of the form
 *     f(hood_1, hood_2, ...) = c_1*g_1(hood_1) + c_2*g_2(hood_2) + ...
 * evaluated left to right, every multiplication and addition rounded separately (what Julia evaluates for
 * `center(a) + 0.1 * sum(neighbors(b))`). Each argument carries its own sweep descriptor (stencil table, boundary,
 * padding; reducer g_j from the menu). `center(hood)` is the one-offset table {0}; a plain array argument
 * (indexed, not stencilled: _getarg, src/gatherstencil.jl:112-113) is the same with Conditional padding.
 * All descriptors must agree on size, dest layout and element type (Float32 / Float64), out_eltype == eltype.
 * `scratch` holds one dest-parent-sized temporary (NULL = library-owned, grown on demand).
 */
typedef struct sb200_term {
    const sb200_desc* desc;  /* sweep of this argument (region fields are ignored) */
    const void* src_parent;  /* device pointer */
    int32_t has_coef;        /* 0: term = g(hood); 1: term = coef * g(hood) */
    int32_t reserved;
    double coef;             /* converted to the element type before use */
} sb200_term;
#define SB200_MAX_TERMS 8
int32_t sb200_gather_multi(const sb200_term* terms, int32_t nterms, void* dst_parent, void* scratch, void* stream);

/* nsteps x { update_halo(src) if the source has a ring and boundary != USE; gather(src->dst); swap }.
   buf_a holds the state on entry; the final state is in buf_a if nsteps is even, else buf_b. The contents of the OTHER buffer
   after the call are unspecified: where a kernel advances several generations per launch (SB200_FLAG_GENS: Life, Diffusion) the
   intermediate states never exist in memory, and B3/S23 Life runs of >= 12 generations on grids above 4 Mi cells (axis 0 a
   multiple of 128) keep their state one bit per cell between the first and the last launch (SB200_FLAG_SRC_BITS / _DST_BITS),
   in packed grids placed inside the two buffers themselves — nothing is allocated. SB200_LIFE_PACKED=0 / 1 turns the packed
   runs off / on for every grid size. */
int32_t sb200_iterate(const sb200_desc* d, void* buf_a, void* buf_b, int32_t nsteps, void* stream);

/* ---- host-buffer entry points (what a StencilArray over a CPU Array lowers to) ---- */
/* H2D(src) -> [update_halo] -> gather -> D2H(dst), pipelined over row chunks on internal streams; blocking. */
int32_t sb200_gather_host(const sb200_desc* d, const void* src_parent_host, void* dst_parent_host);
/* H2D once, nsteps sweeps resident in HBM, D2H once; blocking. state_host is the source parent on
   entry and receives the final source parent (SwitchingStencilArray loop). */
int32_t sb200_iterate_host(const sb200_desc* d, void* state_parent_host, int32_t nsteps);

/* ---- device memory helpers so a shim needs no CUDA binding of its own ---- */
int32_t sb200_device_count(int32_t* n);
int32_t sb200_set_device(int32_t dev);
int32_t sb200_malloc(void** p, size_t bytes);
int32_t sb200_free(void* p);
int32_t sb200_malloc_host(void** p, size_t bytes); /* pinned */
int32_t sb200_free_host(void* p);
int32_t sb200_memcpy_h2d(void* dst, const void* src_host, size_t bytes, void* stream);
int32_t sb200_memcpy_d2h(void* dst_host, const void* src, size_t bytes, void* stream);
int32_t sb200_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int32_t sb200_memset(void* p, int32_t byte, size_t bytes, void* stream);
int32_t sb200_stream_sync(void* stream);

/* ---- peer-memory ghost exchange for slab-partitioned iterated runs (one process per GPU) ---- */
/* Export / import a CUDA IPC handle (64 bytes) for a buffer allocated with sb200_malloc. */
int32_t sb200_ipc_export(void* p, void* handle64);
int32_t sb200_ipc_import(const void* handle64, void** p);
int32_t sb200_ipc_close(void* p);
/* Copy `nplanes` trailing-axis planes of my buffer into a peer's buffer (P2P store over NVLink),
   then publish `value` to a flag word in the peer's memory with system-scope release semantics. */
int32_t sb200_push_planes(const void* src, void* peer_dst, size_t bytes, uint32_t* peer_flag, uint32_t value,
                          void* stream);
/* Stream-ordered system-scope release store of `value` to a flag word (normally in a peer's memory): everything the
   stream wrote before — including the fused mirror stores of a sweep — is visible to a peer that acquires the flag. */
int32_t sb200_signal_flag(uint32_t* peer_flag, uint32_t value, void* stream);
/* Stream-ordered wait until *flag >= value (acquire). A flag that is not published within SB200_WAIT_TIMEOUT_MS (default
   30 s) traps the kernel: the stream's next synchronisation fails instead of the GPU hanging on a dead neighbour. */
int32_t sb200_wait_flag(const uint32_t* flag, uint32_t value, void* stream);

/* Releases everything the library keeps between calls on the CURRENT device and thread: cached plans (device offset / weight
   tables), the host-buffer scratch of sb200_gather_host / sb200_iterate_host, the scratch of sb200_gather_multi. The caller
   guarantees that no call of the library is in flight. */
int32_t sb200_shutdown(void);
/* Diagnostics / tests: how sb200_iterate splits `nsteps` generations into launches when the sizes in `size_mask` (bit g set =
   launches of g generations are available, g = 2 .. 8; single generations always are) can be used, with the relative launch
   times of Life (life != 0) or of the two-step diffusion kernel. out[g] = launches of g generations, g = 0 .. 8. Host only. */
int32_t sb200_debug_split_steps(int32_t nsteps, int32_t size_mask, int32_t life, int32_t* out);

/* ---- slab-partitioned iterated sweeps: the multi-GPU form of the SwitchingStencilArray loop ----
 *
 *   loop of  A = gatherstencil!(f, A::SwitchingStencilArray)             src/gatherstencil.jl:77-83, src/array.jl:610-611
 *
 * with A split into slabs along its LAST axis, one slab per GPU (the reference itself has no multi-device
 * path). A plan owns, per slab: two parents [G ghost planes | owned planes | G ghost planes] (the double buffer of the
 * SwitchingStencilArray), a mailbox (landing slots for the neighbours' planes + flags), two streams, events; and the cycle
 * schedule: ghosts are exchanged every G / radius generations, the boundary planes of the last sweep of a cycle are stored
 * straight into the neighbour's mailbox over NVLink by the sweep kernel and pulled into the ghost zones on a side stream
 * while the interior sweep runs. Results are bit-identical to sb200_iterate on the undivided array for any number of
 * slabs and any G. B3/S23 Life plans on a ring whose rows are multiples of 128 cells run PACKED inside a sb200_plan_iterate
 * call (the first sweep reads the byte parents and writes one bit per cell, sweeps and ghost exchanges work on packed rows, the
 * last sweep writes the byte parents again): sb200_plan_slab / _load_host / _store_host always see bytes.
 *
 * `global` describes the UNDIVIDED array: unpadded (src_off = dst_off = 0, ext = size), out_eltype == eltype, boundary
 * per axis (Wrap / Remove / Reflect on the split axis), region / mirror fields unused.
 *
 * Two ways to build a plan:
 *   sb200_plan_create       one process drives every GPU (what a Julia session does): slab i lives on devices[i]; a device
 *                           may appear more than once (several slabs on one GPU — how the 1-GPU tests exercise the exchange).
 *                           Neighbours are ordered by CUDA events.
 *   sb200_plan_create_rank  one process per GPU (torchrun / MPI): this process owns slab `rank` of `world` on its current
 *                           device; exchange sb200_plan_ipc_handle() results between the processes by any means and hand
 *                           all of them to sb200_plan_connect(). Neighbours are ordered by system-scope flags in peer memory.
 */
typedef struct sb200_plan sb200_plan;

#define SB200_PLAN_OVERLAP_OFF 1   /* never overlap the exchange with the interior sweep */
#define SB200_PLAN_OVERLAP_ON 2    /* always (default: when a ghost zone is >= 1 MiB) */
#define SB200_PLAN_SINGLE_STEP 4   /* one generation per launch only */
#define SB200_PLAN_FLAGS_SYNC 8    /* sb200_plan_create: order neighbours with the flag protocol of the rank form (testing) */

int32_t sb200_plan_create(const sb200_desc* global, int32_t nslabs, const int32_t* devices, int32_t ghost,
                          int32_t plan_flags, sb200_plan** out);
int32_t sb200_plan_create_rank(const sb200_desc* global, int32_t rank, int32_t world, int32_t ghost, int32_t plan_flags,
                               sb200_plan** out);
int32_t sb200_plan_ipc_handle(sb200_plan* p, void* handle64);
/* handles: world x 64 bytes, handles[r] from rank r */
int32_t sb200_plan_connect(sb200_plan* p, const void* handles);

/* Slab i of THIS plan (rank form: i = 0): owned planes [*lo, *hi) of the global last axis, the device it lives on and
   the device pointer of its first owned plane in the CURRENT state (valid until the next sb200_plan_iterate). */
int32_t sb200_plan_nslabs(const sb200_plan* p, int32_t* n);
int32_t sb200_plan_slab(sb200_plan* p, int32_t i, int64_t* lo, int64_t* hi, int32_t* device, void** owned);
/* The caller has written the owned planes through the pointers of sb200_plan_slab (device-side initialisation): the ghost
   planes are stale and UInt8 Life cells are no longer known to be 0/1. (A new plan starts in this state.) */
int32_t sb200_plan_mark_dirty(sb200_plan* p);
/* Copy the state in / out. `state_host` addresses the planes this plan owns, contiguous: the whole array in the
   single-process form, the rank's slab in the rank form. Blocking. */
int32_t sb200_plan_load_host(sb200_plan* p, const void* state_host);
int32_t sb200_plan_store_host(sb200_plan* p, void* state_host);
/* Enqueue nsteps generations on the plan's streams (returns at once); sb200_plan_sync waits for them and reports a
   ghost exchange that timed out (a dead neighbour) as SB200_ECUDA instead of hanging the GPU. */
int32_t sb200_plan_iterate(sb200_plan* p, int32_t nsteps);
int32_t sb200_plan_sync(sb200_plan* p);
/* sb200_plan_iterate bracketed by CUDA events on every slab's compute stream + sb200_plan_sync; *ms = the slowest slab. */
int32_t sb200_plan_iterate_timed(sb200_plan* p, int32_t nsteps, float* ms);
/* out[0] generations done, [1] sweep launches, [2] ghost exchanges, [3] ghost planes G, [4] generations per exchange,
   [5] 1 if exchanges overlap the interior sweep, [6] largest generations per launch in use, [7] 1 = event-ordered, 2 = flags. */
int32_t sb200_plan_stats(const sb200_plan* p, int64_t out[8]);
int32_t sb200_plan_destroy(sb200_plan* p);

/* The schedule a plan runs, as data (diagnostics + CPU tests: the list is interpreted there with a CPU sweep plugged in).
   Planes: >= 0 counted from the start of a slab's parent, < 0 from its end (ext + value). */
enum { SB200_SLAB_SWEEP = 1, SB200_SLAB_PUSH = 2, SB200_SLAB_SIGNAL = 3, SB200_SLAB_PULL = 4, SB200_SLAB_JOIN = 5,
       SB200_SLAB_ENDFILL = 6, SB200_SLAB_SWAP = 7 };
enum { SB200_SLAB_CUR = 0, SB200_SLAB_NXT = 1 };
enum { SB200_SLAB_MIRROR_DOWN = 1, SB200_SLAB_MIRROR_UP = 2 };
typedef struct sb200_slab_op {
    int32_t kind;   /* SB200_SLAB_* */
    int32_t gens;   /* SWEEP: generations in this launch */
    int32_t mirror; /* SWEEP: also store owned planes [G, 2G) into the lower neighbour's slot (DOWN) / [n, n+G) into the upper's (UP) */
    int32_t buf;    /* PUSH / PULL / ENDFILL: SB200_SLAB_CUR or SB200_SLAB_NXT */
    int32_t async;  /* PULL: on the side stream, overlapping the next sweep, closed by JOIN */
    int32_t first;  /* SWEEP: the very first sweep of the plan (UInt8 Life cells not yet known to be 0/1) */
    int64_t lo, hi; /* SWEEP: parent planes [lo, hi) of the split axis (signed encoding) */
} sb200_slab_op;
/* Ops of `nsteps` generations for a plan with radius R, G ghost planes, `n_min` owned planes in its thinnest slab, starting
   `since` generations after an exchange (G / R = ghosts stale). Launches of `gens` > 1 generations are taken when gens <=
   max_gens and the sweep covers at least `min_planes_multi` planes. *since_out receives the state after the steps. */
int32_t sb200_slab_schedule(int32_t radius, int32_t ghost, int64_t n_min, int32_t split_wrap, int32_t overlap,
                            int32_t max_gens, int32_t min_planes_multi, int32_t since, int32_t first_sweep, int32_t nsteps,
                            sb200_slab_op* ops, int32_t cap, int32_t* count, int32_t* since_out);

#ifdef __cplusplus
}
#endif
#endif /* STENCILS_B200_H */
