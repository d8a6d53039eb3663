#!/usr/bin/env python
"""bench.py — alert cells / second to a converged, quorum-decided cut (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nodes n] [--workload c5|c2|c3]

One "step" = one pass of the hot path over one synthetic alert batch:
    filter -> per-receiver cut detection (subject-bucketed kernel) -> implicit invalidation -> per-node proposal
    fingerprints -> fast-round vote tally [-> NCCL histogram all-reduce when sharded] -> decision on the host.
Default workload (N=1): BASELINE config 5 — 1,000,000 virtual nodes, K=10, H=9, L=4, a 1 % churn batch (5,000 crashes
+ 5,000 joins, ~10^5 alert cells) — on however many GPUs --gpus names; receivers are sharded by ring-0 range, the
cluster size stays fixed ("scaling": "strong").

`value`  : device time only, cell arrays resident in HBM when the timed region starts (CUDA events on the library's
           streams, summed over the calls of a step, max over ranks).  The epoch reset between steps (clear() / new
           FastPaxos, the reference's decideViewChange) is outside the timed region.
`e2e`    : the same through the host-facing C ABI: host arrays in, H2D copies, epoch reset, kernels, decision read
           back — wall clock.
--impl reference : the reference's own CPU path.  The reference is Java and cannot be built or run in this image (no
           JDK), so this times oracle/'s literal C++ restatement of it (kind "port") on all host cores, on a bounded
           sample of the same workload, extrapolated linearly (see `sample`).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K, H, L = 10, 9, 4          # Cluster.java:72-74
METRIC = "alert_cells_per_sec_to_converged_cut"
UNIT = "cells/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--nodes", type=int, default=1_000_000)
    p.add_argument("--workload", default="c5", choices=["c5", "c2", "c3"])
    p.add_argument("--kernel", default="auto", choices=["auto", "bucketed", "sweep"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--emulate-shard", type=int, default=0,
                   help="tuning aid: run rank 0's shard of a G-way run on ONE GPU without NCCL (no decision is reached)")
    p.add_argument("--cpu-seconds", type=float, default=20.0, help="budget for the CPU baseline sample")
    return p.parse_args()


def workload_name(args, A, S):
    n = args.nodes
    if args.workload == "c5":
        return "C5 %d-node K=10 H=9 L=4, 1%% churn batch (%d DOWN + %d UP subjects, %d cells)" % (n, n // 200, n // 200, A)
    if args.workload == "c2":
        return "C2 %d-node K=10 H=9 L=4, 1%% simultaneous crash (%d subjects, %d cells)" % (n, S, A)
    return "C3 %d-node K=10 H=9 L=4, 5%% correlated one-way partition (%d subjects, %d cells)" % (n, S, A)


def make_batch(args, W, obs, joiner_obs, ring0):
    n = args.nodes
    if args.workload == "c5":
        return W.c5_churn(obs, joiner_obs, n, n // 200, n // 200)
    if args.workload == "c2":
        return W.c2_simultaneous_crash(obs, n, 0.01)
    return W.c3_correlated_partition(obs, ring0, n, 0.05)


def n_joiners(args):
    return args.nodes // 200 if args.workload == "c5" else 0


# --------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows, self.proc, self.dev, self.t_begin = [], None, device_index, 0.0
        self.first = threading.Event()

    def start(self):
        """Start ONE polling nvidia-smi and wait until it delivers its first row: its start-up (NVML init over every GPU of
        the box) must not overlap the timed region, or it slows the rank that shares the GPU with it."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            self.first.wait(10.0)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
            self.first.set()

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        timed = [r for t, r in self.rows if t >= self.t_begin]
        for r in (timed if timed else [r for _, r in self.rows]):
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(args, gpus):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu --set full
    capture of this very configuration (profiles/traffic.json); None if that configuration was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("%s:%d:%d" % (args.workload, args.nodes, gpus))
    except Exception:
        return None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------------------
# CPU legs (oracle): cpu_baseline at N=1 and --impl reference
# --------------------------------------------------------------------------------------------------------------
class CpuProblem:
    """The workload inside the oracle (built once; construction is outside every timed region)."""

    def __init__(self, args):
        from oracle import oracle_py as orc
        from rapid_b200 import workloads as W
        self.orc, self.args = orc, args
        n, nj = args.nodes, n_joiners(args)
        t0 = time.time()
        hb, off, ports = W.packed_endpoints(0, n + nj)
        self.u = orc.Universe()
        tags = self.u.add_bulk(hb, off, ports)
        hi, lo = W.node_ids(0, n)
        self.view = orc.MembershipView(self.u, K, tags[:n], hi, lo)
        self.cfg = self.view.getCurrentConfigurationId()
        obs = lambda ids: self.view.tables(ids)[0]
        joiner_obs = np.asarray([self.view.getExpectedObserversOf(n + j) for j in range(nj)], np.int32).reshape(nj, K)
        ring0 = np.asarray(self.view.getRing(0), np.int32) if args.workload == "c3" else None
        self.b = make_batch(args, W, obs, joiner_obs, ring0)
        self.A = len(self.b)
        self.cfgs = np.full(self.A, self.cfg, np.int64)
        self.live = int(n - int(self.b.blocked.sum()))
        self.setup_s = time.time() - t0
        self.Rs, self.Vs = None, None       # sample sizes, fixed by the first measurement

    def measure(self, threads, budget_s):
        """Time the literal C++ restatement on a bounded sample; value = cells/s for the WHOLE cluster (extrapolated)."""
        orc, b, n = self.orc, self.b, self.args.nodes
        # apply: R_s receivers x the full batch.  First measurement: a probe of 2 receivers per thread sizes the sample so that it
        # costs ~budget/4 (one more run); later measurements reuse the size.
        def run_apply(Rs):
            sim = orc.ClusterSim(self.view, K, H, L, Rs)
            out = sim.apply_batch(b.src, b.dst, b.ring, b.status, self.cfgs, threads=threads)
            return out, sim.last_seconds
        if self.Rs:
            Rs = self.Rs
            (o_len, o_ann, o_ids, o_off), t_apply = run_apply(Rs)
        else:
            Rs = max(1, min(threads * 2, self.live))
            (o_len, o_ann, o_ids, o_off), t_apply = run_apply(Rs)
            want = int(Rs * (budget_s / 4) / max(t_apply, 1e-3)) // threads * threads
            want = max(Rs, min(want, 64 * threads, max(threads, self.live // threads * threads)))
            if want >= 2 * Rs:
                Rs = want
                (o_len, o_ann, o_ids, o_off), t_apply = run_apply(Rs)
        self.Rs = Rs
        assert (o_len == len(b.expected_cut)).all(), "oracle did not converge to the expected cut"
        apply_whole = t_apply * self.live / Rs
        # tally: `threads` FastPaxos instances x V_s of the `live` identical votes
        prop = o_ids[o_off[0]: o_off[1]]
        def run_tally(Vs):
            senders = np.arange(Vs, dtype=np.int32)
            return orc.sim_tally(self.u, self.cfg, n, threads, senders, np.full(Vs, self.cfg, np.int64), np.zeros(Vs, np.int32),
                                 np.array([0, len(prop)], np.int32), prop, threads=threads)[3]
        if self.Vs:
            Vs = self.Vs
            t_tally = run_tally(Vs)
        else:
            Vs = min(64, self.live)
            t_tally = run_tally(Vs)
            want = max(Vs, min(int(Vs * (budget_s / 4) / max(t_tally, 1e-3)), self.live))
            if want >= 2 * Vs:
                Vs = want
                t_tally = run_tally(Vs)
        self.Vs = Vs
        tally_whole = t_tally * (self.live / threads) * (self.live / Vs)
        whole = apply_whole + tally_whole
        return {
            "value": self.A / whole, "unit": UNIT, "cores": threads, "kind": "port",
            "apply_only_value": self.A / apply_whole,
            "sample": ("literal C++ restatement of MultiNodeCutDetector / MembershipService batch handler / FastPaxos tally "
                       "(oracle/, g++ -O2; the Java reference cannot run here: no JDK).  apply: %d of %d live virtual nodes x "
                       "the full %d-cell batch on %d threads = %.3f s; tally: %d nodes x %d of %d votes (each vote re-hashes "
                       "the %d-endpoint proposal list like List.hashCode) = %.3f s; both extrapolated linearly to all %d "
                       "nodes and votes (whole job %.3g s, of which tally %.3g s)"
                       % (Rs, self.live, self.A, threads, t_apply, threads, Vs, self.live, len(prop), t_tally, self.live,
                          whole, tally_whole)),
            "setup_s": round(self.setup_s, 1),
        }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as orc
    orc.build()
    threads = max(1, orc.hardware_threads())
    prob = CpuProblem(args)
    log("[reference] setup %.1fs (n=%d, cells=%d), %d threads" % (prob.setup_s, args.nodes, prob.A, threads))
    n_iter = args.warmup + args.steps
    per = max(1.0, min(args.cpu_seconds, 150.0 / max(1, n_iter)))      # the whole run stays within a few minutes
    vals = []
    t_start = time.time()
    for i in range(n_iter):
        d = prob.measure(threads, per)
        if i >= args.warmup:
            vals.append(d)
        if time.time() - t_start > 200 and vals:
            break
    v = float(np.mean([x["value"] for x in vals]))
    d = dict(vals[-1])
    d["value"] = v
    A = prob.A
    S = len(np.unique(prob.b.dst))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": 1e3 * A / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic",
        "config": {"workload": workload_name(args, A, S), "nodes": args.nodes, "cells": A, "subjects": S, "K": K, "H": H, "L": L},
        "cpu_baseline": d,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import rapid_b200 as rb
    from rapid_b200 import _native as N
    from rapid_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log("note: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus))
    G = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (librapid_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if G > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, nj = args.nodes, n_joiners(args)
    t0 = time.time()
    hb, off, ports = W.packed_endpoints(0, n + nj)
    view = rb.MembershipView.from_packed(K, hb[: off[n]], off[: n + 1], ports[:n], device=local)
    if nj:
        jh = hb[off[n]: off[n + nj]]
        first = C.c_int32(0)
        N.check(N.lib().rapid_view_register_joiners(view._h, nj, N.ptr(np.ascontiguousarray(jh)),
                                                    N.ptr(np.ascontiguousarray(off[n:] - off[n])),
                                                    N.ptr(np.ascontiguousarray(ports[n:])), C.byref(first)))
        assert first.value == n
    obs, _ = view.tables()
    ring0 = view.getRing(0)
    joiner_obs = view.joinerTables() if nj else np.zeros((0, K), np.int32)
    b = make_batch(args, W, obs, joiner_obs, ring0)
    A = len(b)
    S = len(np.unique(b.dst))
    hi, lo = W.node_ids(0, n)
    cfg = view.getCurrentConfigurationId(hi, lo)
    Gs = args.emulate_shard if (args.emulate_shard > 1 and G == 1) else G      # sharding arithmetic only
    begin = rank * n // Gs
    R = (rank + 1) * n // Gs - begin
    emulated = Gs != G
    blocked = W.blocked_by_receiver(b.blocked, ring0, begin, R)
    cl = rb.VirtualCluster(view, H, L, n_receivers=R, receiver_begin=begin, kernel=args.kernel, max_subjects=S + 64)
    fp = rb.FastPaxos(cfg, n, sender_capacity=n, device=local)
    comm = None
    if G > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.from_numpy(rb.NcclComm.unique_id()))
        dist.broadcast(uid, 0)
        comm = rb.NcclComm(rank, G, uid.cpu().numpy(), local)
    want = rb.proposal_fingerprint(b.expected_cut)
    log("[rank %d] setup %.1fs: n=%d receivers=[%d,%d) cells=%d subjects=%d" % (rank, time.time() - t0, n, begin, begin + R, A, S))

    # device-resident inputs (torch only holds the memory)
    d_dst = torch.from_numpy(b.dst).cuda()
    d_ring = torch.from_numpy(b.ring).cuda()
    d_status = torch.from_numpy(b.status).cuda()
    d_blocked = torch.from_numpy(blocked).cuda()
    dl = N.Delivery()
    dl.flags = N.DELIVERY_BLOCKED
    dl.blocked = d_blocked.data_ptr()
    lib = N.lib()

    phases = [0.0, 0.0, 0.0]     # apply call, its dominant kernel, tally call (device ms, summed)

    def step_device():
        cl.clear()
        fp.reset(cfg)
        N.check(lib.rapid_cd_apply_batch_dev(cl._h, cfg, A, None, d_dst.data_ptr(), d_ring.data_ptr(),
                                             d_status.data_ptr(), None, C.byref(dl)))
        res = fp.tallyCluster(cl, comm)
        tot, main = cl.lastDeviceMs()
        phases[0] += tot; phases[1] += main; phases[2] += fp.lastDeviceMs()
        return res, tot + fp.lastDeviceMs(), main, cl.lastPath()[1] + fp.lastLaunches()

    def step_host():
        cl.clear()
        fp.reset(cfg)
        cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, read_outputs=False)
        return fp.tallyCluster(cl, comm)

    def barrier():
        torch.cuda.synchronize()
        if G > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()             # polls from here on; samples taken from the start of the timed region are reported
    for _ in range(max(3, args.warmup)):
        res, _, _, _ = step_device()
    assert emulated or (res.decided and (res.hash, res.hash2, res.length) == (want[0], want[1], len(b.expected_cut))), \
        "decision differs from the expected cut: %r" % (res,)

    import gc
    gc.collect()
    gc.disable()                   # no collector pauses inside the timed loops (one late rank stalls the whole all-reduce)
    clocks.mark_begin()
    barrier()                      # nothing rank-specific between the barrier and the first timed step: a rank that starts late
    phases[:] = [0.0, 0.0, 0.0]    # makes every other rank wait for it inside the first all-reduce
    dev_ms, main_ms, launches = 0.0, 0.0, 0
    per_step = []
    w0 = time.perf_counter()
    for _ in range(args.steps):
        t_step = time.perf_counter()
        res, ms, mk, nl = step_device()
        dev_ms += ms; main_ms += mk; launches += nl
        per_step.append((time.perf_counter() - t_step) * 1e3)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    log("[rank %d] per step: apply %.3f ms (dominant kernel %.3f ms), tally %.3f ms; host wall per step median %.3f max %.3f ms" % (
        rank, phases[0] / args.steps, phases[1] / args.steps, phases[2] / args.steps, float(np.median(per_step)), max(per_step)))
    # end-to-end through the host-facing ABI (host arrays, H2D, reset, kernels, decision back)
    for _ in range(2):
        step_host()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_host()
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / args.steps
    gc.enable()
    clk = clocks.stop() if rank == 0 else None
    assert emulated or (res.decided and res.hash == want[0])

    # per-rank breakdown (diagnosis of a late rank: every other rank's wait for it shows up in their tally time)
    mine = torch.tensor([float(np.median(per_step)), max(per_step), phases[0] / args.steps, phases[2] / args.steps],
                        dtype=torch.float64, device="cuda")
    if G > 1:
        allr = [torch.zeros_like(mine) for _ in range(G)]
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    per_rank = [[round(float(x), 4) for x in r.cpu()] for r in allr]
    t = torch.tensor([dev_ms, main_ms, e2e_ms, wall_ms], dtype=torch.float64, device="cuda")
    if G > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, main_ms, e2e_ms, wall_ms = [float(x) for x in t.cpu()]
    ms_per_step = dev_ms / args.steps
    main_per = main_ms / args.steps

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        fresh = True   # every step starts a new epoch: all subjects are new, their state is written, never read
        alg = (2 if fresh else 4) * S * R + 5 * R      # mask bytes written + rflags/blocked read (fresh subjects: no per-receiver partials)
        achieved = alg / (main_per * 1e-3) / 1e9 if main_per > 0 else 0.0
        line = {
            "metric": METRIC, "value": A / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": G, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": workload_name(args, A, S), "nodes": n, "cells": A, "alert_messages": b.n_messages(),
                       "subjects": S, "K": K, "H": H, "L": L, "receivers_per_gpu": R,
                       "parallelism": "virtual nodes sharded by ring-0 range x%d; one NCCL histogram all-reduce" % G,
                       "l2": "per-step state %.1f GB per GPU >> 126 MB L2 (inputs larger than L2, no flush)" % (2 * S * R / 1e9),
                       "timing": "CUDA events on the library streams per call, summed per step, max over ranks; "
                                 "epoch reset between steps excluded",
                       "kernel_path": {1: "sweep", 2: "bucketed-uniform", 3: "bucketed-generic"}.get(cl.lastPath()[0])},
            "wall_ms_per_step_incl_reset": wall_ms / args.steps,
            "per_rank_ms": {"host_wall_per_step_median": [r[0] for r in per_rank], "host_wall_per_step_max": [r[1] for r in per_rank],
                            "apply_device": [r[2] for r in per_rank], "tally_device_incl_allreduce_wait": [r[3] for r in per_rank]},
            "clocks": clk,
            "e2e": {"value": A / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(A * 6 + R), "d2h_bytes_per_step": 64 + 36},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_apply_uniform", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args, G), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg), "kernel_ms": main_per,
                         "kernel_share_of_step": main_per / ms_per_step if ms_per_step else None},
        }
    if G > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_cpu_baseline and G == 1:
            try:
                from oracle import oracle_py as orc
                orc.build()
                line["cpu_baseline"] = CpuProblem(args).measure(max(1, orc.hardware_threads()), args.cpu_seconds)
            except Exception as e:       # the baseline is a reported extra; never lose the GPU line over it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        emit(line)
    if G > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """the ONE JSON line, on the real stdout"""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    args = parse()
    # Libraries print to stdout behind our back (NCCL's "NCCL version ..." banner): keep the real stdout for the JSON
    # line only and point fd 1 at stderr for everything else.
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
