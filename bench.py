#!/usr/bin/env python
"""bench.py — headline benchmark of the stencil sweep (BASELINE.json metric: Gcell-updates/s and fraction of the
HBM roofline per stencil config).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--strong]
                    [--workload life|mean|mean1000|mean_halo|kernel|kernel_fma|circle|positional|scatter|window3d|diffusion]
    python bench.py --impl reference ...        # the reference algorithm on the host cores (CPU oracle)

One "step" is one sweep of the hot path over the whole grid. The default workload is BASELINE.json configs[1]: Game of Life,
Moore(1), UInt8 16384x16384, Wrap, SwitchingStencilArray iterated (sb200_iterate). Timing: W warm-up steps, then >= 10
repetitions of the K-step region (>= 50 ms in total), each bracketed by CUDA events on the launching stream; `value` uses the
median repetition. For the iterated workloads (life, diffusion) K is rounded UP to whole exchange cycles
of the slab plans (Life 126 generations, diffusion 4; >= 4 cycles) at EVERY N, N = 1 included, so that a scaling series times
one schedule; `steps_timed` says what ran and `k_step_calls` carries the rate of sb200_iterate calls of exactly K generations.

With N > 1 (torchrun, one rank per GPU) every rank owns a 16384x16384 slab of a 16384 x (16384*N) torus (weak scaling; --strong
splits the one-GPU grid instead) and runs it through the C-ABI slab plan (sb200_plan_create_rank / _connect / _iterate_timed:
ghost rows over NVLink peer memory). The timed step count is rounded up to whole exchange cycles (>= 4 per repetition), times
are the max over ranks. BASELINE configs[4] (3-D diffusion 1024^3 per GPU, the configuration the north star names for
scaling) is measured the same way beside the headline at every N: `c5_diffusion`. Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full summary, if any."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload, {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def ncu_pipes(workload):
    """Pipe / issue utilisation of the dominant kernel from the same committed ncu capture (what `binding` refers to)."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        e = json.load(open(p)).get(workload, {})
        return {k: e[k] for k in ("kernel", "pipe_alu_pct", "pipe_fma_pct", "issue_active_pct", "dram_pct_of_peak") if k in e} or None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ synthetic inputs
def synth(shape, dtype, seed):
    from stencils_b200.synth import synth_np
    return synth_np(shape, dtype, seed)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.stop, self.thr, self.max_mhz = [], set(), threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self.thr = threading.Thread(target=self._run, daemon=True)
            self.thr.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.thr:
            self.thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ workloads
def workloads():
    """name -> spec. bytes_per_cell is the algorithmic figure of SURVEY §8(d): each cell read once + written once."""
    return {
        "life": dict(desc="Game of Life: Moore(1), UInt8 16384x16384, Wrap, SwitchingStencilArray (BASELINE configs[1])",
                     shape=(16384, 16384), dtype=np.uint8, bytes_per_cell=2, iterated=True, seed=0x5EED0002),
        "mean": dict(desc="mapstencil(mean, Window(1)) Float64 16384x16384, Remove(0) (configs[0] at roofline size)",
                     shape=(16384, 16384), dtype=np.float64, bytes_per_cell=16, iterated=False, seed=0x5EED0001),
        "mean1000": dict(desc="mapstencil(mean, Window(1)) Float64 1000x1000, Remove(0) (configs[0], README size)",
                         shape=(1000, 1000), dtype=np.float64, bytes_per_cell=16, iterated=False, seed=0x5EED0001),
        "mean_halo": dict(desc="mapstencil(mean, Window(1)) Float64 16384x16384, Remove(0), padding=Halo{:out} (ring refresh + "
                               "sweep per call; rows of the padded parent are not 16-byte multiples)",
                          shape=(16384, 16384), dtype=np.float64, bytes_per_cell=16, iterated=False, seed=0x5EED0001),
        "kernel": dict(desc="kernelproduct, Kernel(Window(3), 7x7 Float32) 16384x16384, Remove(0)/Conditional (configs[2])",
                       shape=(16384, 16384), dtype=np.float32, bytes_per_cell=8, iterated=False, seed=0x5EED0003),
        "kernel_fma": dict(desc="the same with SB200_FLAG_ALLOW_FMA: acc = fma(v, w, acc), one rounding per tap (opt-in; the north star "
                                "allows 2 ulp; max ulp against the bit-exact kernel over the whole grid is reported)",
                           shape=(16384, 16384), dtype=np.float32, bytes_per_cell=8, iterated=False, seed=0x5EED0003),
        "circle": dict(desc="maximum over Circle(4), Float32 32768x32768, Remove(0) (configs[3]a)",
                       shape=(32768, 32768), dtype=np.float32, bytes_per_cell=8, iterated=False, seed=0x5EED0004),
        "scatter": dict(desc="scatterstencil!(+) Positional((-1,1),(-2,-1),(1,0),(-2,2)), val=centre*w, Float32 32768x32768 (configs[3]b)",
                        shape=(32768, 32768), dtype=np.float32, bytes_per_cell=12, iterated=False, seed=0x5EED0004),
        "positional": dict(desc="mapstencil(sum, Positional((-1,1),(-2,-1),(1,0),(-2,2))) Float32 16384x16384, Wrap "
                                "(run-time offset table; README.md:102 stencil)",
                           shape=(16384, 16384), dtype=np.float32, bytes_per_cell=8, iterated=False, seed=0x5EED0006),
        "window3d": dict(desc="mapstencil(mean, Window(1,3)) Float32 768^3, Wrap (27-point 3-D table at run time)",
                         shape=(768, 768, 768), dtype=np.float32, bytes_per_cell=8, iterated=False, seed=0x5EED0007),
        "diffusion": dict(desc="3-D diffusion: VonNeumann(1,3) Float32 1024^3, Wrap, iterated (configs[4])",
                          shape=(1024, 1024, 1024), dtype=np.float32, bytes_per_cell=8, iterated=True, seed=0x5EED0005),
    }


def make_sweep(name, spec, torch, sb, shape=None):
    """Returns (state dict, run(nsteps) -> None stream-ordered, cells_per_step)."""
    shape = tuple(shape or spec["shape"])
    dev = torch.device("cuda", torch.cuda.current_device())
    from stencils_b200.synth import synth_torch
    src = synth_torch(shape, spec["dtype"], spec["seed"], dev)  # bit-identical to synth_np (tests/test_gpu_api.py)
    cells = int(np.prod(shape))
    st = {"name": name}
    if name == "life":
        S = sb.SwitchingStencilArray(src, sb.Moore(1), boundary=sb.Wrap())
        st["S"] = S

        def run(n):
            st["S"] = sb.iterate_(sb.Life(), st["S"], n)
    elif name == "diffusion":
        S = sb.SwitchingStencilArray(src, sb.VonNeumann(1, 3), boundary=sb.Wrap())
        st["S"] = S

        def run(n):
            st["S"] = sb.iterate_(sb.Diffusion(0.1), st["S"], n)
    elif name in ("mean", "mean1000"):
        a = sb.StencilArray(src, sb.Window(1), boundary=sb.Remove(0.0))
        dst = sb.colmajor_empty(shape, torch.float64, dev)

        def run(n):
            for _ in range(n):
                sb.mapstencil_(sb.mean, dst, a)
    elif name == "mean_halo":
        a = sb.StencilArray(src, sb.Window(1), boundary=sb.Remove(0.0), padding=sb.Halo("out"))
        dst = sb.colmajor_empty(shape, torch.float64, dev)

        def run(n):
            for _ in range(n):
                sb.mapstencil_(sb.mean, dst, a)
    elif name in ("kernel", "kernel_fma"):
        from stencils_b200 import _abi as A_
        w = synth((7, 7), np.float32, 0x5EED1003)
        w = (w / w.sum(dtype=np.float32)).astype(np.float32)
        a = sb.StencilArray(src, sb.Kernel(sb.Window(3), w), boundary=sb.Remove(np.float32(0)))
        dst = sb.colmajor_empty(shape, torch.float32, dev)
        fl = A_.FLAG_ALLOW_FMA if name == "kernel_fma" else 0
        if fl:   # accuracy of the contracted fold against the bit-exact kernel of the same library, whole grid
            exact = sb.colmajor_empty(shape, torch.float32, dev)
            sb.mapstencil_(sb.kernelproduct, exact, a)
            sb.mapstencil_(sb.kernelproduct, dst, a, flags=fl)
            ia, ib = exact.T.contiguous().view(torch.int32).to(torch.int64), dst.T.contiguous().view(torch.int32).to(torch.int64)
            st["max_ulp_vs_exact_kernel"] = int((ia - ib).abs().max().item())   # all values are positive: bit patterns are ordered
            del exact, ia, ib

        def run(n):
            for _ in range(n):
                sb.mapstencil_(sb.kernelproduct, dst, a, flags=fl)
    elif name == "circle":
        a = sb.StencilArray(src, sb.Circle(4), boundary=sb.Remove(np.float32(0)))
        dst = sb.colmajor_empty(shape, torch.float32, dev)

        def run(n):
            for _ in range(n):
                sb.mapstencil_(sb.maximum, dst, a)
    elif name == "positional":
        a = sb.StencilArray(src, sb.Positional((-1, 1), (-2, -1), (1, 0), (-2, 2)), boundary=sb.Wrap())
        dst = sb.colmajor_empty(shape, torch.float32, dev)

        def run(n):
            for _ in range(n):
                sb.mapstencil_(sb.sum, dst, a)
    elif name == "window3d":
        a = sb.StencilArray(src, sb.Window(1, 3), boundary=sb.Wrap())
        dst = sb.colmajor_empty(shape, torch.float32, dev)

        def run(n):
            for _ in range(n):
                sb.mapstencil_(sb.mean, dst, a)
    elif name == "scatter":
        a = sb.StencilArray(src, sb.Positional((-1, 1), (-2, -1), (1, 0), (-2, 2)), boundary=sb.Remove(np.float32(0)))
        dst = sb.colmajor_empty(shape, torch.float32, dev)
        dst.zero_()
        rule = sb.ScatterCenterWeights(np.array([0.4, 0.3, 0.2, 0.1], dtype=np.float32))
        import operator

        def run(n):
            for _ in range(n):
                sb.scatterstencil_(rule, operator.add, dst, a)
    else:
        raise SystemExit(f"unknown workload {name}")
    return st, run, cells


def time_reps(torch, run, steps, warmup, min_reps=10, min_total_ms=50.0, max_reps=200):
    """SURVEY 8d timing: `warmup` untimed steps, then R >= 10 repetitions of the K-step region, each bracketed by CUDA events
    on the launching stream (R is raised until the timed regions add up to >= 50 ms). Returns the list of per-repetition
    times in ms. Nothing but the K steps of a repetition sits between its two events."""
    run(max(warmup, 1))
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    def one():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(steps)
        e1.record(stream)
        return e0, e1

    e0, e1 = one()   # pilot repetition (untimed): sizes R
    torch.cuda.synchronize()
    pilot = max(e0.elapsed_time(e1), 1e-3)
    reps = int(min(max_reps, max(min_reps, -(-1.3 * min_total_ms // pilot))))   # 1.3: the pilot repetition runs cold
    evs = [one() for _ in range(reps)]
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def rep_stats(ms_list, steps):
    a = np.asarray(ms_list, dtype=np.float64)
    return {"reps": int(a.size), "steps_per_rep": int(steps), "rep_ms_median": float(np.median(a)), "rep_ms_min": float(a.min()),
            "rep_ms_max": float(a.max()), "timed_region_ms": float(a.sum()),
            "how": "CUDA events on the launching stream around every repetition of the K-step region; value uses the median repetition"}


# ------------------------------------------------------------------------------------------------ CPU baseline / reference arm
def cpu_threads():
    """All host cores for the CPU arm: torchrun exports OMP_NUM_THREADS=1 to its workers, which silently ran the round-1
    reference arm on ONE core at N > 1 — the count is set explicitly here."""
    from oracle import oracle as orc
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except Exception:
        pass
    orc.set_threads(n)
    return orc.threads()


def cpu_life_step_factory(rows, spec):
    """Reference algorithm (CPU oracle = restatement of Stencils.jl's CPU path) on a bounded sample: a
    16384 x rows torus of the same synthetic field; per-cell work is identical to the full grid."""
    from oracle import np_restatement as npr
    from oracle import oracle as orc
    from stencils_b200 import _abi as A
    from stencils_b200._desc import build_desc
    cores = cpu_threads()
    shape = (spec["shape"][0], rows)
    a = np.asfortranarray(synth(shape, np.uint8, spec["seed"]))
    b = np.zeros_like(a, order="F")
    h = build_desc(size=shape, eltype=A.U8, out_eltype=A.U8, offsets=npr.offsets("Moore", 1, 2), radius=1,
                   boundary=A.WRAP, reducer=A.LIFE)
    bufs = [a, b]

    def step():
        orc.gather(h, bufs[0], bufs[1])
        bufs.reverse()
    return step, shape[0] * rows, cores


def cpu_baseline(spec, budget_s=12.0):
    step, cells, cores = cpu_life_step_factory(1024, spec)
    step()  # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    n = 0
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 2000:
            break
    return {"value": cells * n / el / 1e9, "unit": "Gcell-updates/s", "cores": cores, "kind": "port",
            "sample": f"Life on a 16384x1024 torus slab of the same field, {n} generations in {el:.1f} s, "
                      f"OpenMP static over columns, {cores} threads (CPU oracle: C restatement of Stencils.jl's CPU path; "
                      "the Julia reference itself cannot run in this image)"}


def run_reference(args, rank):
    if rank != 0:
        return
    spec = workloads()["life"]
    # size the sample so that (warmup + steps) generations take about 100 s
    step, cells, cores = cpu_life_step_factory(256, spec)
    step()
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    rate = cells * 3 / (time.perf_counter() - t0)
    total = max(args.steps + args.warmup, 1)
    rows = int(rate * 100.0 / total / spec["shape"][0])
    rows = max(64, min(spec["shape"][1], rows // 64 * 64))
    step, cells, cores = cpu_life_step_factory(rows, spec)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    val = cells * args.steps / el / 1e9
    sample = (f"each step = one Life generation on a 16384x{rows} torus slab of the synthetic field "
              f"(bounded sample of the 16384x16384 grid), {cores} OpenMP threads")
    line = {"impl": "reference", "metric": "gcell_updates_per_s", "value": val, "unit": "Gcell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": spec["desc"], "sample": sample},
            "cpu_baseline": {"value": val, "unit": "Gcell-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ slab-partitioned runs (N > 1)
def slab_case(workload):
    from stencils_b200 import _abi as A
    from stencils_b200.stencils import Moore, VonNeumann
    if workload == "life":
        return dict(st=Moore(1), reducer=A.LIFE, kw=dict(born_mask=1 << 3, survive_mask=0b1100), eltype=A.U8, R=1, ghost=0,
                    bcs=(A.WRAP, A.WRAP))   # ghost = 0: the library's default (128 rows for slabs of >= 1024 rows)
    if workload == "diffusion":
        return dict(st=VonNeumann(1, 3), reducer=A.DIFFUSION, kw=dict(alpha=0.1), eltype=A.F32, R=1, ghost=0, bcs=(A.WRAP, A.WRAP, A.WRAP))
    raise SystemExit(f"workload {workload} is not an iterated (slab-partitioned) configuration")


def plan_cycle(workload):
    """Generations per ghost exchange of the slab plans' library defaults (csrc/slab_plan.cu: Life 126 rows = 21 packed launches of six
    generations, diffusion 4 planes); bench_slabs reads the same number from the plan's own stats at N > 1."""
    return {"life": 126, "diffusion": 4}[workload]


def bench_slabs(torch, dist, workload, spec, steps, warmup, strong, min_reps=10, min_total_ms=50.0):
    """One slab per rank through the C-ABI slab plan (sb200_plan_create_rank / _connect / _iterate_timed; csrc/slab_plan.cu).
    Weak scaling: every rank owns spec['shape'], the global last axis is world x as long; strong: spec['shape'] is split.
    The timed region is whole exchange cycles: the step count of a repetition is rounded UP to a multiple of the steps per
    exchange and to at least four cycles, so every repetition contains >= 4 ghost exchanges whatever --steps says.
    Returns a dict (value over ALL ranks, per-repetition times = max over ranks)."""
    from stencils_b200 import _abi as A
    from stencils_b200.slab import SlabPlan
    from stencils_b200.synth import synth_torch
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    c = slab_case(workload)
    if os.environ.get("SB200_PLAN_GHOST"):   # A/B knob: ghost planes per side (a multiple of the radius)
        c["ghost"] = int(os.environ["SB200_PLAN_GHOST"])
    shape = tuple(spec["shape"])
    if strong:
        if shape[-1] % world:
            raise SystemExit(f"--strong needs the last axis ({shape[-1]}) to be a multiple of the number of GPUs ({world})")
        gshape, local = shape, shape[:-1] + (shape[-1] // world,)
    else:
        gshape, local = shape[:-1] + (shape[-1] * world,), shape
    cells_local = int(np.prod(local))
    exchange = os.environ.get("SB200_EXCHANGE", "auto")
    pflags = {"1": A.PLAN_OVERLAP_ON, "0": A.PLAN_OVERLAP_OFF}.get(os.environ.get("SB200_OVERLAP", ""), 0)
    try:
        if exchange == "nccl":
            raise A.SB200Error(A.ECUDA, "SB200_EXCHANGE=nccl: the NCCL fallback was requested")
        plan = SlabPlan(gshape, offsets=c["st"].offsets(), radius=c["R"], reducer=c["reducer"], boundary=c["bcs"], eltype=c["eltype"],
                        ghost=c["ghost"], rank=rank, world=world, reducer_kwargs=c["kw"], plan_flags=pflags)
    except A.SB200Error as ex:   # raised on every rank together (the constructor votes): CUDA IPC is not available between the ranks
        return bench_slabs_nccl(torch, dist, workload, spec, c, local, gshape, steps, warmup, strong, min_reps, min_total_ms, repr(ex))
    try:
        lo, hi, _, ptr = plan.slab(0)
        assert hi - lo == local[-1]
        # rank r's slab is planes [lo, hi) of the global field -> linear index offset lo * cells per plane
        field = synth_torch(local, spec["dtype"], spec["seed"], dev, lo=lo * int(np.prod(local[:-1])))
        lib = A.lib()
        A.check(lib.sb200_memcpy_d2d(ptr, field.data_ptr(), cells_local * field.element_size(), None))
        A.check(lib.sb200_stream_sync(None))
        del field
        plan.mark_dirty()
        st_ = plan.stats()
        c["ghost"] = st_["ghost_planes"]            # what the library chose when asked for its default
        k = st_["generations_per_exchange"]
        steps_timed = max(-(-steps // k) * k, 4 * k)
        plan.iterate(max(warmup, k))
        plan.sync()
        kernel = lib.sb200_last_kernel().decode()
        dist.barrier()
        pilot = torch.tensor([plan.iterate_timed(steps_timed)], device=dev, dtype=torch.float64)
        dist.all_reduce(pilot, op=dist.ReduceOp.MAX)
        reps = int(min(200, max(min_reps, -(-min_total_ms // max(float(pilot.item()), 1e-3)))))
        st0 = plan.stats()
        lib.sb200_launch_count(1)
        times = []
        for _ in range(reps):
            dist.barrier()
            times.append(plan.iterate_timed(steps_timed))
        launches = lib.sb200_launch_count(1)
        st1 = plan.stats()
        t = torch.tensor(times, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # every repetition: the slowest rank
        times = [float(v) for v in t.tolist()]
        kernel = lib.sb200_last_kernel().decode() or kernel
        dist.barrier()
    finally:
        plan.close()
    med = float(np.median(times))
    value = cells_local * world * steps_timed / (med * 1e-3) / 1e9
    ex_per_rep = (st1["exchanges"] - st0["exchanges"]) / reps
    how = ("boundary planes of the last sweep of a cycle are written into the neighbour's mailbox over NVLink by the sweep itself "
           "(sb200_desc.mirror_*; a stream-ordered peer copy inside the call for kernels without the fused store), published with a "
           "system-scope release flag, pulled into the ghost planes on a side stream under the interior sweep"
           if st1["overlap"] else
           "peer copies of the boundary planes into the neighbour's mailbox over NVLink + system-scope release / acquire flags in "
           "front of the first sweep of every cycle (ghost zones < 1 MiB: overlap does not pay)")
    return {"value": value, "unit": "Gcell-updates/s", "ms_per_step": med / steps_timed, "steps_timed": steps_timed,
            "exchanges_in_timed_region": ex_per_rep, "steps_per_exchange": k, "ghost_planes": c["ghost"],
            "generations_per_launch_max": st1["max_generations_per_launch"], "launches": int(launches), "launches_per_rep": launches / reps,
            "timing": rep_stats(times, steps_timed), "kernel": kernel, "exchange": how, "sync": st1["sync"],
            "grid_per_gpu": list(local), "global_grid": list(gshape), "scaling": "strong" if strong else "weak",
            "cells_total": cells_local * world, "api": "sb200_plan_create_rank / sb200_plan_connect / sb200_plan_iterate_timed (C ABI)",
            "exchange_requested": exchange}


def bench_slabs_nccl(torch, dist, workload, spec, c, local, gshape, steps, warmup, strong, min_reps, min_total_ms, why):
    """Fallback when the ranks cannot open each other's memory (no CUDA IPC): the round-1 Python slab iterator with NCCL
    send / recv for the ghost planes (stencils_b200.slab.SlabIterator). Same rounding of the timed step count to whole exchange
    cycles; the line carries `warning` so that the fallback is visible."""
    from stencils_b200 import _abi as A
    from stencils_b200.slab import SlabIterator
    from stencils_b200.synth import synth_torch
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    cells_local = int(np.prod(local))
    lo = rank * local[-1]
    field = synth_torch(local, spec["dtype"], spec["seed"], dev, lo=lo * int(np.prod(local[:-1])))
    t = field.permute(*reversed(range(len(local)))).contiguous()
    del field
    ghost = c["ghost"] or (32 if workload == "life" else 4)
    it = SlabIterator(t, offsets=c["st"].offsets(), radius=c["R"], reducer=c["reducer"], boundary=c["bcs"], eltype=c["eltype"], ghost=ghost,
                      rank=rank, world=world, reducer_kwargs=c["kw"], exchange="nccl")
    del t
    lib = A.lib()
    k = ghost // c["R"]
    steps_timed = max(-(-steps // k) * k, 4 * k)
    it.step(max(warmup, k))
    torch.cuda.synchronize()

    def one():
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        it.step(steps_timed)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    pilot = torch.tensor([one()], device=dev, dtype=torch.float64)
    dist.all_reduce(pilot, op=dist.ReduceOp.MAX)
    reps = int(min(200, max(min_reps, -(-min_total_ms // max(float(pilot.item()), 1e-3)))))
    lib.sb200_launch_count(1)
    times = [one() for _ in range(reps)]
    launches = lib.sb200_launch_count(1)
    tt = torch.tensor(times, device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    times = [float(v) for v in tt.tolist()]
    kernel = lib.sb200_last_kernel().decode()
    dist.barrier()
    it.close()
    med = float(np.median(times))
    return {"value": cells_local * world * steps_timed / (med * 1e-3) / 1e9, "unit": "Gcell-updates/s", "ms_per_step": med / steps_timed,
            "steps_timed": steps_timed, "exchanges_in_timed_region": steps_timed / k, "steps_per_exchange": k, "ghost_planes": ghost,
            "generations_per_launch_max": None, "launches": int(launches), "launches_per_rep": launches / reps,
            "timing": rep_stats(times, steps_timed), "kernel": kernel, "sync": "nccl",
            "exchange": "NCCL send / recv of the ghost planes (fallback: CUDA IPC peer access is not available between the ranks)",
            "grid_per_gpu": list(local), "global_grid": list(gshape), "scaling": "strong" if strong else "weak",
            "cells_total": cells_local * world, "api": "stencils_b200.slab.SlabIterator over sb200_gather (Python orchestration, NCCL exchange)",
            "exchange_requested": os.environ.get("SB200_EXCHANGE", "auto"), "warning": "peer-memory exchange unavailable, NCCL fallback: " + why}


# ------------------------------------------------------------------------------------------------ main
def roofline_of(workload, spec, value_per_gpu, kernel, peak, peak_src, sweeps_per_launch, ms_total, launches):
    """roofline object of the dominant kernel. `achieved` counts ALGORITHMIC bytes (each cell read once + written
    once per sweep); kernels that fuse several sweeps per launch beat the one-sweep HBM roofline (frac > 1), so the line also
    names the unit that actually binds them and the measured DRAM fraction."""
    traffic = ncu_traffic(workload)
    achieved = value_per_gpu * spec["bytes_per_cell"]
    dram_frac = None
    if traffic and launches and ms_total:
        dram_frac = traffic * launches / (ms_total * 1e-3) / 1e9 / peak
    binding = {"life": "alu", "kernel": "fp32_issue", "kernel_fma": "fp32_issue", "diffusion": "issue", "circle": "alu", "window3d": "fp32_issue"}.get(workload, "hbm")
    if workload == "life" and "life_bit" not in kernel:
        binding = "hbm"
    if workload == "diffusion" and "stream3d2" not in kernel:
        binding = "hbm"
    return {"bound": "hbm", "binding": binding, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "dram_frac": dram_frac, "traffic": traffic, "binding_util_ncu": ncu_pipes(workload),
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel, ncu --set full capture "
                              "committed as profiles/ncu_summary.json (not re-measured in this run)" if traffic else None,
            "peak_source": peak_src, "kernel": kernel, "algorithmic_bytes_per_cell": spec["bytes_per_cell"],
            "sweeps_per_launch": sweeps_per_launch,
            "note": ("frac counts the ALGORITHMIC bytes of every sweep (read once + write once per cell-update); a launch that fuses "
                     "several generations moves the grid through HBM once for all of them, so frac > 1 means the one-sweep HBM roofline "
                     "is beaten by temporal fusion and `binding` names the unit that limits the kernel instead (ncu: profiles/); "
                     "dram_frac = measured DRAM bytes per launch (`traffic`) x launches / time / peak"),
            "how": "algorithmic bytes per launch / mean launch duration (CUDA events on the launching stream around back-to-back launches)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="life")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configs / cpu baseline / e2e legs")
    ap.add_argument("--strong", action="store_true",
                    help="N > 1: split the ONE-GPU grid over the ranks (strong scaling) instead of one full grid per rank (weak, default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = workloads()[args.workload]
    if args.steps is None:
        args.steps = 1000 if args.workload == "life" else (100 if spec["iterated"] else 20)
    if args.warmup is None:
        args.warmup = 10 if spec["iterated"] else 3
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import stencils_b200 as sb
    from stencils_b200 import _abi as A
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = A.lib()
    peak, peak_src = measured_peak()
    dtype_name = {"uint8": "u8", "float64": "f64", "float32": "f32"}[np.dtype(spec["dtype"]).name]
    warnings = []

    if world > 1:
        if not spec["iterated"]:
            raise SystemExit("--gpus N > 1 runs the iterated (slab-partitioned) configurations: --workload life | diffusion")
        with ClockSampler(local_rank) as cs:
            res = bench_slabs(torch, dist, args.workload, spec, args.steps, args.warmup, args.strong)
        value, ms_per_step, kernel = res["value"], res["ms_per_step"], res["kernel"]
        steps_timed, launches, timing = res["steps_timed"], res["launches"], res["timing"]
        cells_total = res["cells_total"]
        extra_cfg = {k: res[k] for k in ("ghost_planes", "steps_per_exchange", "exchange", "sync", "global_grid", "api")}
        if res.get("warning"):
            warnings.append(res["warning"])
        sweeps_per_launch = steps_timed / max(res["launches_per_rep"], 1)
    else:
        st, run, cells_total = make_sweep(args.workload, spec, torch, sb)
        run(1)
        torch.cuda.synchronize()
        # Iterated workloads: a repetition is ONE sb200_iterate call whose step count is --steps rounded UP to whole exchange
        # cycles of the slab plans (>= 4 cycles), exactly as bench_slabs rounds it at N > 1 — every N of a scaling series then
        # times the same schedule (same generations per launch, same number of launches per step). The rate of calls of exactly
        # --steps generations is reported beside it (`k_step_calls`).
        steps_timed = args.steps
        if spec["iterated"]:
            cyc = plan_cycle(args.workload)
            steps_timed = max(-(-args.steps // cyc) * cyc, 4 * cyc)
        with ClockSampler(local_rank) as cs:
            times = time_reps(torch, run, steps_timed, args.warmup)
        kernel = lib.sb200_last_kernel().decode()  # the kernel of the timed region (iterated runs fuse generations)
        timing = rep_stats(times, steps_timed)
        lib.sb200_launch_count(1)
        run(steps_timed)                           # one more (untimed) repetition: its launches, counted by the library
        launches_per_rep = lib.sb200_launch_count(1)
        torch.cuda.synchronize()
        launches = int(launches_per_rep * len(times))
        ms_per_step = timing["rep_ms_median"] / steps_timed
        value = cells_total * steps_timed / (timing["rep_ms_median"] * 1e-3) / 1e9
        extra_cfg = {}
        if steps_timed != args.steps:
            extra_cfg["steps_rounding"] = (f"--steps {args.steps} rounded up to {steps_timed} = whole exchange cycles of the slab plans ({cyc} generations, "
                                           f">= 4 cycles) in ONE sb200_iterate call per repetition: the schedule every N of the scaling series times")
        sweeps_per_launch = steps_timed / max(launches_per_rep, 1e-9) if spec["iterated"] else 1.0

    line = {
        "metric": "gcell_updates_per_s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if (args.strong and world > 1) else "weak",
        "vs_baseline": None, "dtype": dtype_name,
        "data": "synthetic (splitmix64 hash of the linear index, SURVEY 8d)",
        "steps_timed": steps_timed, "timing": timing,
        "config": {"workload": spec["desc"],
                   "grid_per_gpu": list(spec["shape"][:-1]) + [spec["shape"][-1] // world if (args.strong and world > 1) else spec["shape"][-1]],
                   "parallelism": f"slab{world}",
                   "l2": "state per GPU (>= 256 MiB) is larger than the 126 MB L2; no flush needed", **extra_cfg},
        "roofline": roofline_of(args.workload, spec, value / max(world, 1), kernel, peak, peak_src, sweeps_per_launch,
                                timing["timed_region_ms"], launches if world == 1 else None),
        "gpu_launches": int(launches),
        "clocks": cs.summary(),
    }
    if world > 1:
        line["exchanges_in_timed_region"] = res["exchanges_in_timed_region"]
        line["config"]["steps_rounding"] = (f"--steps {args.steps} rounded up to {steps_timed} = whole exchange cycles (>= 4) per repetition")

    # ---- BASELINE configs[4] (C5) beside the headline at every N: 3-D diffusion 1024^3 per GPU, weak scaling ----
    if not args.no_extras and args.workload == "life":
        try:
            if world > 1:
                torch.cuda.empty_cache()
                sp5 = workloads()["diffusion"]
                r5 = bench_slabs(torch, dist, "diffusion", sp5, 100, 8, args.strong)
                r5["workload"] = sp5["desc"]
                r5["roofline_frac"] = r5["value"] / world * sp5["bytes_per_cell"] / peak
                line["c5_diffusion"] = r5
        except Exception as ex:  # pragma: no cover
            line["c5_diffusion"] = {"error": repr(ex)}

    if world > 1 and not args.no_extras and args.workload == "life":
        # e2e at N GPUs: every rank pushes its own slab through the host-buffer entry point concurrently
        # (one generation per call: H2D + sweep + D2H inside every step); whole-job value = sum over ranks
        try:
            e = e2e_host(torch, sb, lib, args.workload, spec, barrier=dist.barrier)
            v = torch.tensor([e["value"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(v)
            e.update(value=float(v.item()), h2d_bytes_per_step=e["h2d_bytes_per_step"] * world,
                     d2h_bytes_per_step=e["d2h_bytes_per_step"] * world,
                     how=e["how"] + f"; {world} ranks concurrently, one slab each, values summed")
            line["e2e"] = e
        except Exception as ex:  # pragma: no cover
            line["e2e"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_extras:
        # ---- e2e: the reference-facing call with HOST buffers, copies inside the timed region ----
        try:
            line["e2e"] = e2e_host(torch, sb, lib, args.workload, spec)
            if line["e2e"] and line["e2e"].get("iterated_call"):
                # what a SwitchingStencilArray user over a host Array hits: one copy in, 100 generations, one copy out (sb200_iterate_host)
                line["e2e_iterated"] = dict(line["e2e"]["iterated_call"], how="iterate_(Life(), SwitchingStencilArray(host array), 100) -> "
                                            "sb200_iterate_host; wall clock around the blocking call, pinned host buffer")
        except Exception as e:  # pragma: no cover
            line["e2e"] = {"error": repr(e)}
        if spec["iterated"] and steps_timed != args.steps:
            # calls of exactly --steps generations: a short call holds fewer full-size launches (20 Life steps = 4 + 4 + 6 + 6)
            try:
                tk = time_reps(torch, run, args.steps, 3)
                lib.sb200_launch_count(1)
                run(args.steps)
                lk = lib.sb200_launch_count(1)
                torch.cuda.synchronize()
                line["k_step_calls"] = {"steps_per_call": args.steps, "value": cells_total * args.steps / (float(np.median(tk)) * 1e-3) / 1e9,
                                        "unit": "Gcell-updates/s", "launches_per_call": int(lk), "timing": rep_stats(tk, args.steps)}
            except Exception as e:  # pragma: no cover
                line["k_step_calls"] = {"error": repr(e)}
        # ---- the other BASELINE configs, same measurement, for context ----
        del st, run
        torch.cuda.empty_cache()
        also = {}
        for name in ("mean", "mean_halo", "mean1000", "kernel", "kernel_fma", "circle", "positional", "scatter", "window3d", "diffusion"):
            if name == args.workload:
                continue
            try:
                sp = workloads()[name]
                st2, run2, cells2 = make_sweep(name, sp, torch, sb)
                k = 5 if not sp["iterated"] else 100
                if name == "mean1000":
                    k = 200
                run2(1)
                torch.cuda.synchronize()
                with ClockSampler(local_rank) as cs2:
                    t2 = time_reps(torch, run2, k, 3)
                kn = lib.sb200_last_kernel().decode()  # the kernel of the timed region (iterated runs may fuse steps)
                med2 = float(np.median(t2))
                v = cells2 * k / (med2 * 1e-3) / 1e9
                also[name] = {"value": v, "unit": "Gcell-updates/s", "ms_per_step": med2 / k, "kernel": kn,
                              "roofline_frac": v * sp["bytes_per_cell"] / peak, "workload": sp["desc"],
                              "roofline_frac_best_rep": cells2 * k / (min(t2) * 1e-3) / 1e9 * sp["bytes_per_cell"] / peak,
                              "timing": {kk: vv for kk, vv in rep_stats(t2, k).items() if kk != "how"}, "clocks": cs2.summary()}
                if "max_ulp_vs_exact_kernel" in st2:
                    also[name]["max_ulp_vs_exact_kernel"] = st2["max_ulp_vs_exact_kernel"]
                if name == "mean1000":
                    also[name]["note"] = "8 MB grid: L2-resident and launch-bound by construction (the README's own benchmark size)"
                del st2, run2
                torch.cuda.empty_cache()
            except Exception as e:  # pragma: no cover
                also[name] = {"error": repr(e)}
        line["other_configs"] = also
        if "diffusion" in also and "value" in also["diffusion"]:
            d5 = dict(also["diffusion"])
            d5.update(scaling="weak", grid_per_gpu=[1024, 1024, 1024], exchanges_in_timed_region=0,
                      api="sb200_iterate on the undivided array (the 1-GPU base of the weak-scaling series)")
            line["c5_diffusion"] = d5
        try:
            line["cpu_baseline"] = cpu_baseline(spec) if args.workload == "life" else None
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"error": repr(e)}
    if warnings:
        line["warnings"] = warnings
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def e2e_host(torch, sb, lib, workload, spec, barrier=None):
    """Same metric through the C-ABI entry point a host-array StencilArray lowers to (sb200_gather_host):
    pinned host source -> HBM -> sweep -> pinned host dest, every step."""
    from stencils_b200.array import _torch_dtype
    shape = spec["shape"]
    if workload != "life":
        return None
    host = torch.empty(tuple(reversed(shape)), dtype=_torch_dtype(spec["dtype"]), pin_memory=True)
    host.copy_(torch.from_numpy(np.ascontiguousarray(synth(shape, spec["dtype"], spec["seed"]).T)))
    out = torch.empty_like(host, pin_memory=True)
    a = sb.StencilArray(host.numpy().T, sb.Moore(1), boundary=sb.Wrap())
    dst = out.numpy().T
    sb.mapstencil_(sb.Life(), dst, a)  # warm-up (allocates the device scratch)
    sb.mapstencil_(sb.Life(), dst, a)
    n = 5
    if barrier:
        barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        sb.mapstencil_(sb.Life(), dst, a)  # blocking call: returns when dst is on the host
    el = time.perf_counter() - t0
    nbytes = int(np.prod(shape)) * np.dtype(spec["dtype"]).itemsize
    # informational: the iterated form of the same host-buffer API (SwitchingStencilArray over a host array ->
    # sb200_iterate_host): ONE copy in, 100 generations resident in HBM, ONE copy out
    iterated = None
    if barrier is None:
        S = sb.SwitchingStencilArray(host.numpy().T, sb.Moore(1), boundary=sb.Wrap())
        sb.iterate_(sb.Life(), S, 2)
        t1 = time.perf_counter()
        sb.iterate_(sb.Life(), S, 100)
        el2 = time.perf_counter() - t1
        iterated = {"value": int(np.prod(shape)) * 100 / el2 / 1e9, "unit": "Gcell-updates/s", "generations_per_call": 100,
                    "h2d_bytes_per_call": nbytes, "d2h_bytes_per_call": nbytes}
    return {"iterated_call": iterated, "value": int(np.prod(shape)) * n / el / 1e9, "unit": "Gcell-updates/s", "h2d_bytes_per_step": nbytes,
            "d2h_bytes_per_step": nbytes, "steps": n,
            "how": "mapstencil_(Life(), dest, StencilArray(host array)) -> sb200_gather_host; wall clock around "
                   "blocking calls, pinned host buffers, H2D + sweep + D2H inside every step"}


if __name__ == "__main__":
    main()
