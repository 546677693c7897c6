// slab_sched.h — the cycle schedule of slab-partitioned iterated sweeps, pure host logic, no CUDA.
//
// The array is split into slabs along its last axis; every slab keeps G ghost planes per side inside its parent
// [G ghost | n owned | G ghost] and ghosts are exchanged every k = G / R steps ("wide halo"). After an exchange the ghost
// planes are exact; a sweep that advances the state to s generations after the exchange writes the planes that are still
// exact, [R s, ext - R s) of the parent, so after k generations exactly the owned planes are valid again. A launch may
// advance m = 1 .. 8 generations (SB200_FLAG_GENS) while the cycle has room: the sizes in `sizes`, tried in that order (Life
// prefers 7: the bit-sliced kernel's best rate, so its default cycle is 126 = 18 x 7 generations), or 8 / 4 / 2 below max_gens.
//
// On the last sweep of a cycle the planes the neighbours need are computed first (two thin boundary sweeps that also
// store them into the neighbours' landing slots: sb200_desc.mirror_*), published (SIGNAL), and pulled into the ghost zones
// of the new state on a side stream (PULL async) while the interior sweep runs; JOIN closes the cycle. Without overlap
// the exchange is PUSH + SIGNAL + PULL in front of the first sweep of the next cycle.
//
// The schedule is a list of rank-independent ops (planes counted from the start of the parent when >= 0, from its end
// when < 0), so that (a) every rank derives the same list without talking to the others, (b) the executor
// (slab_plan.cu) stays a dumb interpreter, and (c) the list can be interpreted on the CPU with a reference sweep plugged in
// (tests/test_slab_schedule.py) — which is how this logic is tested without a GPU.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>
#include "../../include/stencils_b200.h"

namespace sb {

struct SlabSchedCfg {
    int R = 1, G = 1;
    bool split_wrap = true;   // Wrap on the split axis (ring); else Remove / Reflect ends re-imposed after every sweep
    bool overlap = false;     // boundary-first + async pull on the last sweep of a cycle
    int max_gens = 1;         // largest generations-per-launch the reducer / layout supports
    std::vector<int> sizes;   // generations per launch in order of preference (without 1); empty: max_gens, max_gens / 2, ..., 2
    // accept(lo, hi, gens): may the sweep of parent planes [lo, hi) (signed encoding) run `gens` generations in one launch on
    // EVERY slab? (gens == 1 is always accepted.)
    std::function<bool(long long, long long, int)> accept;
    long long n_min = 0;      // owned planes of the thinnest slab
};

struct SlabSched {
    SlabSchedCfg c;
    int k = 1;
    int since = 0;            // generations since the last exchange (k = ghosts stale)
    long long nsweeps = 0;    // sweeps issued so far (the first one may not assume 0/1 Life cells)

    explicit SlabSched(const SlabSchedCfg& cfg) : c(cfg), k(cfg.G / (cfg.R > 0 ? cfg.R : 1)), since(k) {}

    static sb200_slab_op op(int kind, int gens = 0, int mirror = 0, int buf = 0, int async = 0, long long lo = 0, long long hi = 0) {
        sb200_slab_op o;
        o.kind = kind; o.gens = gens; o.mirror = mirror; o.buf = buf; o.async = async; o.first = 0; o.lo = lo; o.hi = hi;
        return o;
    }

    bool ok(long long lo, long long hi, int m) const { return m == 1 || (c.accept && c.accept(lo, hi, m)); }

    // Appends the ops of `nsteps` generations.
    void plan(int nsteps, std::vector<sb200_slab_op>& out) {
        const int R = c.R, G = c.G;
        int left = nsteps;
        while (left > 0) {
            if (since >= k) {   // ghosts are stale: blocking exchange of the current state
                out.push_back(op(SB200_SLAB_PUSH, 0, 0, SB200_SLAB_CUR));
                out.push_back(op(SB200_SLAB_SIGNAL));
                out.push_back(op(SB200_SLAB_PULL, 0, 0, SB200_SLAB_CUR, 0));
                since = 0;
            }
            const int room = k - since;
            int m = 1;
            auto fits = [&](int cand) {
                if (cand <= 1 || cand > left || cand > room) return false;
                const long long s = since + cand;
                return ok((long long)R * s, -(long long)R * s, cand);
            };
            if (c.sizes.empty()) {
                for (int cand = c.max_gens; cand > 1; cand >>= 1)
                    if (fits(cand)) { m = cand; break; }
            } else {
                for (int cand : c.sizes)
                    if (fits(cand)) { m = cand; break; }
            }
            const long long s = since + m;
            const long long lo = (long long)R * s, hi = -(long long)R * s;   // parent planes [R s, ext - R s)
            const bool last = s == k;
            // boundary sweeps of G + 1 planes: a Reflect end mirrors planes G + 1 .. 2 G of the new state
            const long long b = G + 1;
            if (last && c.overlap && c.n_min >= 2 * (long long)G + 2 && ok(lo, lo + b, m) && ok(hi - b, hi, m) && ok(lo + b, hi - b, m)) {
                sb200_slab_op a0 = op(SB200_SLAB_SWEEP, m, SB200_SLAB_MIRROR_DOWN, 0, 0, lo, lo + b);
                sb200_slab_op a1 = op(SB200_SLAB_SWEEP, m, SB200_SLAB_MIRROR_UP, 0, 0, hi - b, hi);
                sb200_slab_op a2 = op(SB200_SLAB_SWEEP, m, 0, 0, 0, lo + b, hi - b);
                a0.first = a1.first = a2.first = nsweeps == 0;
                out.push_back(a0);
                out.push_back(a1);
                out.push_back(op(SB200_SLAB_SIGNAL));
                out.push_back(op(SB200_SLAB_PULL, 0, 0, SB200_SLAB_NXT, 1));
                out.push_back(a2);
                out.push_back(op(SB200_SLAB_JOIN));
                out.push_back(op(SB200_SLAB_SWAP));
                since = 0;
            } else {
                sb200_slab_op a = op(SB200_SLAB_SWEEP, m, 0, 0, 0, lo, hi);
                a.first = nsweeps == 0;
                out.push_back(a);
                if (!c.split_wrap) out.push_back(op(SB200_SLAB_ENDFILL, 0, 0, SB200_SLAB_NXT));
                out.push_back(op(SB200_SLAB_SWAP));
                since = (int)s;
            }
            nsweeps++;
            left -= m;
        }
    }
};

}  // namespace sb
