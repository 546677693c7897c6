/*
 * stencils_oracle.c — CPU restatement of the reference's per-cell sweep.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for libstencils_b200.so. It is NOT part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it. It restates, in plain C, the algorithm of rafaqz/Stencils.jl v0.3.6 for the hot path
 * (the reference is pure Julia and `julia` is not available in this image, so the reference
 * itself cannot be compiled or run here; see DESIGN.md "Oracle").
 *
 * Pinning: every golden vector the reference's own tests hold for this path
 * (test/array.jl, test/stencils.jl) is checked against this oracle in tests/test_oracle_golden.py,
 * and an independent NumPy restatement (oracle/np_restatement.py) cross-checks it on random inputs.
 * Arithmetic that lives in third-party Julia packages (StaticArrays `_mapreduce` left fold,
 * Statistics/StaticArrays `mean` = sum / length) is restated from their published source; results the
 * reference tests do not pin (Float32, Bool/UInt8 sums, min/max, Life, diffusion, 3-D sweeps) are
 * "parity unpinned" and this restatement + DESIGN.md is the spec for them.
 *
 * Compile: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off matters: Julia never contracts a*b+c into an FMA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/stencils_b200.h" /* descriptor layout + enums only (the data format of the boundary) */

#define ORC_OK 0
#define ORC_EINVAL 1
#define ORC_EUNSUPPORTED 2
#define ORC_ESIZE 3

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Offset tables — src/stencils/window.jl:4-8, moore.jl:5-18, vonneumman.jl:5-15, shapes.jl:2-175.
 * All of them iterate CartesianIndices((-R:R)^N), i.e. axis 0 fastest, and keep the cells that pass
 * the shape predicate. (Window in the reference emits 2-tuples regardless of N, which is only right
 * for N<=2; the N-D box here is identical for N=1,2 — SURVEY Appendix A.)
 * ---------------------------------------------------------------------------------------------- */
static int shape_keep(int shape, int R, int RI, int N, const int* t) {
    int manh = 0, maxabs = 0, zeros = 0, sq = 0;
    for (int a = 0; a < N; a++) {
        int v = t[a] < 0 ? -t[a] : t[a];
        manh += v;
        if (v > maxabs) maxabs = v;
        if (t[a] == 0) zeros++;
        sq += t[a] * t[a];
    }
    switch (shape) {
    case SB200_WINDOW: return 1;
    case SB200_MOORE: return manh != 0;                       /* moore.jl: skip the middle position */
    case SB200_VONNEUMANN: return manh >= 1 && manh <= R;     /* vonneumman.jl:10 */
    case SB200_CROSS: return zeros >= N - 1;                  /* shapes.jl:6 */
    case SB200_ANGLEDCROSS: {                                 /* shapes.jl:17-22 */
        int m = 0;
        for (int a = 1; a < N; a++) m += (abs(t[a]) == abs(t[0]));
        return m == N - 1;
    }
    case SB200_FORWARDSLASH: {                                /* shapes.jl:32-37 */
        int m = 0;
        for (int a = 1; a < N; a++) m += (t[a] == -t[0]);
        return m == N - 1;
    }
    case SB200_BACKSLASH: {                                   /* shapes.jl:47-52 */
        int m = 0;
        for (int a = 1; a < N; a++) m += (t[a] == t[0]);
        return m == N - 1;
    }
    case SB200_CIRCLE: return sqrt((double)sq) < R + 0.5;     /* shapes.jl:62 */
    case SB200_VERTICAL: return (N > 1 && t[1] == 0) || (N == 1 && t[0] == 0); /* shapes.jl:74 */
    case SB200_HORIZONTAL: return N > 1 && t[0] == 0;         /* shapes.jl:86 */
    case SB200_DIAMOND: return manh <= R;                     /* shapes.jl:99 */
    case SB200_ANNULUS: {                                     /* shapes.jl:143-147 */
        double dist = sqrt((double)sq);
        return dist < R + 0.5 && dist >= RI + 0.5;
    }
    case SB200_CARDINAL: return manh == R && maxabs == R;     /* shapes.jl:158 */
    case SB200_ORDINAL: return manh == R * N && maxabs == R;  /* shapes.jl:170 */
    default: return -1;
    }
}

int orc_stencil_offsets(int shape, int R, int RI, int N, int32_t* out, int cap, int32_t* count) {
    if (N < 1 || N > 3 || R < 0 || !count) return ORC_EINVAL;
    int D = 2 * R + 1, total = 1, n = 0;
    for (int a = 0; a < N; a++) total *= D;
    for (int lin = 0; lin < total; lin++) {
        int t[3] = {0, 0, 0}, rem = lin;
        for (int a = 0; a < N; a++) { t[a] = rem % D - R; rem /= D; }
        int keep = shape_keep(shape, R, RI, N, t);
        if (keep < 0) return ORC_EUNSUPPORTED;
        if (keep) {
            if (out && n < cap) { out[3 * n] = t[0]; out[3 * n + 1] = t[1]; out[3 * n + 2] = t[2]; }
            n++;
        }
    }
    *count = n;
    return ORC_OK;
}

/* _return_type for the reducer menu (src/gatherstencil.jl:41-59 applied to the named reducers).
 * sum: StaticArrays `_mapreduce(identity, +, ...)` seeded by Base.reduce_first(+, v1): Bool -> Int,
 * everything else keeps its type. mean: sum / length -> Float64 for integers. */
int orc_out_eltype(int reducer, int eltype, int32_t* out) {
    if (eltype < SB200_BOOL || eltype > SB200_F64) return ORC_EUNSUPPORTED;
    int isf = eltype == SB200_F32 || eltype == SB200_F64;
    switch (reducer) {
    case SB200_SUM: *out = eltype == SB200_BOOL ? SB200_I64 : eltype; return ORC_OK;
    case SB200_MEAN: *out = isf ? eltype : SB200_F64; return ORC_OK;
    case SB200_MIN: case SB200_MAX: case SB200_LIFE: *out = eltype; return ORC_OK;
    case SB200_KERNELDOT:
        if (eltype == SB200_BOOL || eltype == SB200_U8) return ORC_EUNSUPPORTED;
        *out = eltype; return ORC_OK;
    case SB200_DIFFUSION:
        if (!isf) return ORC_EUNSUPPORTED;
        *out = eltype; return ORC_OK;
    default: return ORC_EUNSUPPORTED;
    }
}

static size_t elsize(int t) {
    switch (t) { case SB200_BOOL: case SB200_U8: return 1; case SB200_I32: case SB200_F32: return 4; default: return 8; }
}

/* bounded_index, src/array.jl:146-179 (0-based). Returns -1 when the neighbour is out of bounds under
 * Remove (caller substitutes padval, src/array.jl:133-135). */
static inline int64_t bounded(int64_t j, int64_t s, int bc) {
    if (j >= 0 && j < s) return j;
    switch (bc) {
    case SB200_WRAP: return j < 0 ? j + s : j - s;            /* i<1 ? i+s : i>s ? i-s : i */
    case SB200_REFLECT: return j < 0 ? -j : 2 * (s - 1) - j;  /* i<1 ? 2-i : i>s ? 2s-i : i (1-based) */
    default: return -1;
    }
}

static int check_desc(const sb200_desc* d) {
    if (!d || d->ndim < 1 || d->ndim > 3 || d->noffsets < 1 || !d->offsets_host) return ORC_EINVAL;
    for (int a = 0; a < d->ndim; a++) {
        if (d->size[a] < 1) return ORC_ESIZE;
        if (d->src_off[a] > 0) {
            if (d->src_off[a] < d->radius) return ORC_ESIZE;
            if (d->src_ext[a] < d->size[a] + d->src_off[a] + d->radius) return ORC_ESIZE;
        } else {
            if (d->boundary[a] == SB200_USE) return ORC_EUNSUPPORTED; /* Use + Conditional: no getneighbor method */
            if (d->src_ext[a] != d->size[a]) return ORC_ESIZE;
            /* radius-vs-axis check, src/array.jl:451-453 */
            if (d->radius >= d->size[a]) {
                int used = 0;
                for (int k = 0; k < d->noffsets; k++) used |= d->offsets_host[3 * k + a] != 0;
                if (used) return ORC_ESIZE;
            }
        }
        if (d->dst_ext[a] < d->size[a] + d->dst_off[a]) return ORC_ESIZE;
    }
    return ORC_OK;
}

/* Julia max/min for floats: NaN-propagating, -0.0 < +0.0 (Base math: `max(x::T,y::T)`). */
#define JL_FMAX(T, x, y) ((isnan(x) || isnan(y)) ? (T)NAN : ((x) > (y) ? (x) : ((x) < (y) ? (y) : (signbit(x) ? (y) : (x)))))
#define JL_FMIN(T, x, y) ((isnan(x) || isnan(y)) ? (T)NAN : ((x) < (y) ? (x) : ((x) > (y) ? (y) : (signbit(x) ? (x) : (y)))))
#define JL_IMAX(T, x, y) ((x) > (y) ? (x) : (y))
#define JL_IMIN(T, x, y) ((x) < (y) ? (x) : (y))

/* Type-generic sweep bodies: one instantiation per element type. */
#define T uint8_t
#define TNAME u8
#define IS_FLOAT 0
#define IS_BOOL 0
#include "sweep_body.inc"
#undef T
#undef TNAME
#undef IS_FLOAT
#undef IS_BOOL

#define T uint8_t
#define TNAME b8
#define IS_FLOAT 0
#define IS_BOOL 1
#include "sweep_body.inc"
#undef T
#undef TNAME
#undef IS_FLOAT
#undef IS_BOOL

#define T int32_t
#define TNAME i32
#define IS_FLOAT 0
#define IS_BOOL 0
#include "sweep_body.inc"
#undef T
#undef TNAME
#undef IS_FLOAT
#undef IS_BOOL

#define T int64_t
#define TNAME i64
#define IS_FLOAT 0
#define IS_BOOL 0
#include "sweep_body.inc"
#undef T
#undef TNAME
#undef IS_FLOAT
#undef IS_BOOL

#define T float
#define TNAME f32
#define IS_FLOAT 1
#define IS_BOOL 0
#include "sweep_body.inc"
#undef T
#undef TNAME
#undef IS_FLOAT
#undef IS_BOOL

#define T double
#define TNAME f64
#define IS_FLOAT 1
#define IS_BOOL 0
#include "sweep_body.inc"
#undef T
#undef TNAME
#undef IS_FLOAT
#undef IS_BOOL

/* gatherstencil!(f, dest, source): src/gatherstencil.jl:89-109. The caller is responsible for calling
 * orc_update_halo first when the source has a ring (as gatherstencil! does, :93). */
int orc_gather(const sb200_desc* d, const void* src, void* dst) {
    int rc = check_desc(d);
    if (rc) return rc;
    int32_t want;
    rc = orc_out_eltype(d->reducer, d->eltype, &want);
    if (rc) return rc;
    if (want != d->out_eltype) return ORC_EINVAL;
    switch (d->eltype) {
    case SB200_BOOL: return orc_gather_b8(d, src, dst);
    case SB200_U8: return orc_gather_u8(d, src, dst);
    case SB200_I32: return orc_gather_i32(d, src, dst);
    case SB200_I64: return orc_gather_i64(d, src, dst);
    case SB200_F32: return orc_gather_f32(d, src, dst);
    case SB200_F64: return orc_gather_f64(d, src, dst);
    }
    return ORC_EUNSUPPORTED;
}

/* update_boundary!(A): src/array.jl:195-239. */
int orc_update_halo(const sb200_desc* d, void* parent) {
    int rc = check_desc(d);
    if (rc) return rc;
    switch (d->eltype) {
    case SB200_BOOL: case SB200_U8: return orc_halo_u8(d, parent);
    case SB200_I32: return orc_halo_i32(d, parent);
    case SB200_I64: return orc_halo_i64(d, parent);
    case SB200_F32: return orc_halo_f32(d, parent);
    case SB200_F64: return orc_halo_f64(d, parent);
    }
    return ORC_EUNSUPPORTED;
}

/* scatterstencil!(f, op, dest, source): src/scatterstencil.jl:36-112 (2-D only, same eltype). */
int orc_scatter(const sb200_desc* d, const void* src, void* dst) {
    int rc = check_desc(d);
    if (rc) return rc;
    if (d->ndim != 2) return ORC_EUNSUPPORTED;
    if (d->out_eltype != d->eltype) return ORC_EINVAL;
    if (!d->weights_host) return ORC_EINVAL;
    switch (d->eltype) {
    case SB200_I32: return orc_scatter_i32(d, src, dst);
    case SB200_I64: return orc_scatter_i64(d, src, dst);
    case SB200_F32: return orc_scatter_f32(d, src, dst);
    case SB200_F64: return orc_scatter_f64(d, src, dst);
    }
    return ORC_EUNSUPPORTED;
}

/* Loop of gatherstencil!(f, A::SwitchingStencilArray) + switch (src/gatherstencil.jl:77-83,
 * src/array.jl:610-611). Final state is in a when nsteps is even, else in b. */
int orc_iterate(const sb200_desc* d, void* a, void* b, int nsteps) {
    void *s = a, *t = b;
    for (int i = 0; i < nsteps; i++) {
        int need_halo = 0;
        for (int ax = 0; ax < d->ndim; ax++) need_halo |= (d->src_off[ax] > 0 && d->boundary[ax] != SB200_USE);
        int rc;
        if (need_halo && (rc = orc_update_halo(d, s))) return rc;
        if ((rc = orc_gather(d, s, t))) return rc;
        void* tmp = s; s = t; t = tmp;
    }
    return ORC_OK;
}

size_t orc_sizeof(int eltype) { return elsize(eltype); }
