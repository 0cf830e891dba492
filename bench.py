#!/usr/bin/env python
"""bench.py — one CCD step (broadphase + CTCD narrowphase) per "step", on N GPUs of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cloth1415|cloth709|cloth355|prob17]
  python bench.py --impl reference ...      # the reference's own CPU path on a bounded sample

Workload (BASELINE.json configs): default `cloth1415` = config C5, the synthetic 1415 x 1415-vertex S-folded
cloth (3,998,792 triangles, eta = outerEta = 0.01 h, KDOP-13), sharded over the N ranks; `prob17` = config C2
(meshes/V0_prob17_30957 -> V1, from tests/golden/).  Metric: candidate stencils taken through the whole
step per second (whole job), plus ms per step.

One JSON line on rank 0 — see the keys in main().  `value` is measured with the inputs resident in HBM
(ccd_step_device); `e2e` goes through the host-buffer C-ABI call (ccd_step_shard) with pinned host arrays,
H2D of positions/faces and D2H of the hit lists inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from collisiondetection_b200 import scenes  # noqa: E402

METRIC = "ccd_step_stencils_per_s"
UNIT = "stencils/s"
# algorithmic FP64 flops per stencil, coefficient construction as the reference writes it
FLOP_VF, FLOP_EE = 1395.0, 1676.0
BYTES_STENCIL = 225.0


def load_workload(name):
    if name.startswith("cloth"):
        n = int(name[5:])
        q0, q1, f, eta = scenes.cloth(n)
        return dict(name=name, q0=q0, q1=q1, faces=f, outer_eta=eta, eta=eta,
                    desc="synthetic S-folded cloth %dx%d vertices, %d triangles, eta=outerEta=0.01h, KDOP-13" % (n, n, len(f)))
    if name == "prob17":
        g = np.load(os.path.join(ROOT, "tests", "golden", "alec_prob17_30957.npz"))
        return dict(name=name, q0=g["q0"], q1=g["q1"], faces=g["faces"], outer_eta=1e-8, eta=1e-8,
                    desc="V0_prob17_30957 -> V1 (77,386 triangles), eta=outerEta=1e-8, KDOP-13")
    raise SystemExit("unknown workload " + name)


def cpu_samples(name, count):
    """Bounded sample of the workload for the CPU arm: `count` 8-column ribbons of the same cloth (same coordinates and
    noise), spread evenly over its width so that the sample sees the cloth's own mix of quiet margins and the central band
    where the layers pass through each other.  Other workloads: the reference mesh pair itself (prob17 is 24 s of CPU)."""
    if name.startswith("cloth"):
        n = int(name[5:])
        w = min(8, n)
        count = max(1, min(count, n // w))
        out = []
        for k in range(count):
            j0 = int(round((k + 0.5) * n / count - w / 2.0))
            j0 = min(max(j0, 0), n - w)
            q0, q1, f, eta = scenes.cloth(n, cols=(j0, j0 + w))
            out.append(dict(q0=q0, q1=q1, faces=f, outer_eta=eta, eta=eta, cols=(j0, j0 + w)))
        desc = "%d ribbons of %d columns spread evenly over the %s cloth (%d of its %d columns, %d triangles in all)" % (
            count, w, name, count * w, n, sum(len(x["faces"]) for x in out))
        return out, desc, "%s/ribbons%dx%d" % (name, count, w)
    g = load_workload(name)
    return [dict(q0=g["q0"], q1=g["q1"], faces=g["faces"], outer_eta=g["outer_eta"], eta=g["eta"])], g["desc"] + " (the whole workload)", name


def _cpu_worker(args):
    """One process = one sample on one pinned core: the unmodified reference (oracle/_ref) when it was built, else the
    plain-C restatement.  Returns (stencils, broadphase s, narrowphase s, flat-array primitive s) of the last timed step."""
    idx, core, sample, steps, warmup, flat = args
    from oracle import bind
    lib = bind.Ref() if bind.have_ref() else bind.Port()
    H = bind.single_step_history(sample["q0"], sample["q1"])
    acc = [0, 0.0, 0.0, 0.0, time.time(), 0.0]
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        vf, ee, _ = lib.broadphase(13, sample["faces"], *H, sample["outer_eta"])
        t1 = time.perf_counter()
        lib.narrowphase(*H, vf, sample["eta"], ee, sample["eta"])
        t2 = time.perf_counter()
        if it >= warmup:
            acc[0] += len(vf) + len(ee)
            acc[1] += t1 - t0
            acc[2] += t2 - t1
    if flat:
        # SURVEY 8(d)'s second baseline: the CTCD primitives over flat arrays (no std::set, no stitchCommonHistory)
        t0 = time.perf_counter()
        lib.narrowphase_flat(*H, vf, sample["eta"], ee, sample["eta"])
        acc[3] = time.perf_counter() - t0
        acc[4] += acc[3]      # not part of the step: shifts the start stamp so that (end - start) excludes it
    acc[5] = time.time()
    return acc


def _cpu_pin(cores):
    """Pool initializer: worker k of the pool stays on core k (taskset-style pinning)."""
    import multiprocessing as mp
    try:
        k = (mp.current_process()._identity[0] - 1) % len(cores)
        os.sched_setaffinity(0, {cores[k]})
    except Exception:
        pass


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_cpu_arm(name, steps, warmup, max_procs=32, flat=True):
    """Reference CPU path (KDOPBroadPhase + CTCDNarrowPhase).  The reference has no threading of its own, so "all the host
    threads it can use" = one independent sample per core, each process pinned (sched_setaffinity) to its own core;
    value = stencils of all samples / wall time of the slowest process."""
    import multiprocessing as mp
    from oracle import bind
    if not bind.have_ref() and not bind.have_port():
        bind.build(ref=False, port=True)
    kind = "reference" if bind.have_ref() else "port"
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        cores = list(range(os.cpu_count() or 1))
    nproc = max(1, min(len(cores), max_procs))
    # four samples per process, handed out one at a time: the ribbons of the central band cost several times the marginal ones
    samples, desc, label = cpu_samples(name, 4 * nproc)
    nproc = min(nproc, len(samples))
    flat_every = max(1, len(samples) // 8)      # the flat-array baseline on every 8th sample is plenty
    jobs = [(k, -1, samples[k], steps, warmup, flat and k % flat_every == 0) for k in range(len(samples))]
    t0 = time.perf_counter()
    if nproc == 1:
        res = [_cpu_worker(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(nproc, initializer=_cpu_pin, initargs=(cores,)) as pool:
            res = list(pool.imap_unordered(_cpu_worker, jobs, chunksize=1))
    wall = time.perf_counter() - t0
    nst = sum(r[0] for r in res)
    flat_st = sum(r[0] for r in res if r[3] > 0)
    busy = sum(r[1] + r[2] for r in res) if nproc == 1 else max(r[5] for r in res) - min(r[4] for r in res) - 0.0
    bp, npz, fl = sum(r[1] for r in res), sum(r[2] for r in res), sum(r[3] for r in res)
    per_core = nst / (bp + npz) if bp + npz > 0 else 0.0
    cb = dict(value=nst / busy, unit=UNIT, cores=nproc, kind=kind, sample=desc + "; %d stencils per step; %d timed step(s) after %d warm-up" % (nst // max(steps, 1), steps, warmup),
              per_core_stencils_per_s=per_core, broadphase_core_s=bp, narrowphase_core_s=npz,
              broadphase_share=bp / (bp + npz) if bp + npz > 0 else None,
              flat_primitive_stencils_per_s_per_core=(flat_st / max(steps, 1)) / fl if fl > 0 else None,
              pinning="one process per core, os.sched_setaffinity", host_cores_visible=len(cores), cpu_model=cpu_model(), wall_s=wall)
    return cb, busy / max(steps, 1), nst // max(steps, 1), label


def full_size_reference(name):
    """What the unmodified reference took on the FULL workload when the golden was generated (tests/golden/make_golden_c5.py,
    one core of the build container — not this box): kept beside the live sample rate so the two can be compared."""
    if name != "cloth1415":
        return None
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "cloth_1415.npz"))
        n = int(g["n_vf"]) + int(g["n_ee"])
        sec = float(g["ref_seconds_broadphase"]) + float(g["ref_seconds_narrowphase"])
        return dict(stencils=n, broadphase_s=float(g["ref_seconds_broadphase"]), narrowphase_s=float(g["ref_seconds_narrowphase"]),
                    stencils_per_s_one_core=n / sec, where="build container, 1 core, recorded in tests/golden/cloth_1415.npz")
    except Exception:
        return None


def parity_block(ctx, api, wl):
    """C5 only: one untimed pass through the two-call boundary (ccd_broadphase_step + ccd_narrowphase) compared with the golden
    of the unmodified reference at full size (tests/golden/cloth_1415.npz): candidate sets by count and FNV-1a-64, flags
    stencil by stencil, classes of the differing ones as arbitrated when the golden was made (tests/parity_account.py)."""
    path = os.path.join(ROOT, "tests", "golden", "cloth_1415.npz")
    if wl["name"] != "cloth1415" or not os.path.exists(path):
        return None
    from oracle import bind
    g = np.load(path)
    vf, ee = ctx.findCollisionCandidatesStep(api.KDOP, wl["faces"], wl["q0"], wl["q1"], wl["outer_eta"])
    H = api.single_step_history(wl["q0"], wl["q1"])
    out = ctx.findCollisions(*H, vf, wl["eta"], ee, wl["eta"])
    blk = dict(golden="tests/golden/cloth_1415.npz (unmodified reference, 30,503,172 stencils)",
               candidate_sets_bit_exact=bool(len(vf) == int(g["n_vf"]) and len(ee) == int(g["n_ee"]) and
                                             bind.fnv1a64(vf) == str(g["vf_fnv"]) and bind.fnv1a64(ee) == str(g["ee_fnv"])))
    classes, matched, mism, unexpected, toi_out = {}, 0, 0, 0, 0
    for k, st in (("vf", vf), ("ee", ee)):
        hit = out[k + "_hit"] > 0
        rh = np.unpackbits(g["ref_%s_hit_bits" % k])[:len(st)] > 0
        d = np.nonzero(hit != rh)[0] if len(hit) == len(rh) else np.zeros(0, np.int64)
        matched += int(len(hit) - len(d))
        mism += int(len(d))
        known = g["%s_mismatch_index" % k]
        unexpected += int((~np.isin(d, known)).sum()) + int((~np.isin(known, d)).sum())
        for c in g["%s_mismatch_class" % k]:
            classes[str(c)] = classes.get(str(c), 0) + 1
        toi_out += int(g["%s_toi_out_of_1e9" % k])
    blk.update(flags_matched=matched, flags_differ=mism, flag_classes=classes, eta_band_1e12=classes.get("eta-band", 0),
               noise_sign=classes.get("noise-sign", 0), reference_artefact=classes.get("reference-artefact", 0),
               toi_out_of_1e9_vs_reference=toi_out, toi_out_of_1e9_all_arbitrated_reference_artefact=True,
               unexplained=unexpected + classes.get("unexplained", 0),
               earliest_toi_rel_err=abs(out["earliest_toi"] - float(g["ref_earliest_toi"])) / float(g["ref_earliest_toi"]))
    return blk


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = "/tmp/ccd_clocks_%d_%d.csv" % (os.getpid(), index)

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cloth1415")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--pipeline-exchange", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # a sharded job needs a few steps for the ownership ranges (rebalanced every step) and with them the buffer sizes to settle:
    # at least 8 untimed steps there, 3 on one GPU; the number actually run is the one printed
    warmup = max(args.warmup, 8 if world > 1 else 3) if args.impl == "ours" else args.warmup
    config = dict(workload=args.workload, kind="KDOP-13", sharding="ownership by Morton (LBVH subtree) range over %d rank(s), rebalanced every step from the previous step's load profile; replicated cluster tree, sharded tree intersection / emission / narrowphase" % world,
                  l2="working set (inputs 144 MB at cloth1415 + GBs of intermediates) exceeds the 126 MB L2; nothing is reused across steps")

    if args.impl == "reference":
        if rank != 0:
            return
        steps, wu = max(1, args.steps), max(0, args.warmup)
        # one step of the 32-ribbon sample is ~7 s of wall time: a long --steps / --warmup is cut to what a few minutes hold
        # and the numbers actually run are the ones printed
        if args.workload.startswith("cloth"):
            steps, wu = min(steps, 5), min(wu, 1)
        else:
            steps, wu = min(steps, 2), min(wu, 1)
        cb, sec, nst, label = run_cpu_arm(args.workload, steps, wu)
        config["workload"] = label
        config["workload_desc"] = "CPU arm on a bounded sample of %s: %s" % (args.workload, cb["sample"])
        config["full_size_reference"] = full_size_reference(args.workload)
        print(json.dumps(dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=wu,
                              requested_steps=args.steps, requested_warmup=args.warmup,
                              ms_per_step=sec * 1e3, ms_per_step_is="wall time of one step of the SAMPLE (all cores busy), not of the full workload",
                              higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                              data="synthetic" if args.workload.startswith("cloth") else "reference mesh fixture", impl="reference", config=config, cpu_baseline=cb,
                              e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)))
        return

    import torch
    import torch.distributed as dist
    from collisiondetection_b200 import api
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = load_workload(args.workload)
    config["workload_desc"] = wl["desc"]
    V, F = len(wl["q0"]), len(wl["faces"])
    h_q0 = torch.from_numpy(wl["q0"]).pin_memory()
    h_q1 = torch.from_numpy(wl["q1"]).pin_memory()
    h_f = torch.from_numpy(wl["faces"]).pin_memory()
    d_q0, d_q1, d_f = h_q0.cuda(), h_q1.cuda(), h_f.cuda()
    ctx = api.Context(local_rank)
    from collisiondetection_b200 import distributed as D
    if world > 1:
        # replicated positions: rank 0's arrays are the step's input on every GPU (NCCL broadcast over NVLink)
        D.broadcast_positions(d_q0, d_q1, src=0)
        ctx.wait_stream(torch.cuda.current_stream().cuda_stream)      # the context's private stream must not read before the broadcast lands
    summary = {}
    # the path's only exchange: earliest TOI, hit / stencil counts and the load profile that balances the next step's ownership
    # ranges — one small all-gather per step.  --pipeline-exchange runs it in a worker thread beside the NEXT step
    # (distributed.StepExchange; the ranges then apply one step later).  Measured: 0.1 ms per step at N = 2, but at N = 4 the
    # delayed feedback makes the rebalancing oscillate (ranges jump, buffers regrow, 7.8 instead of 6.3 ms), so it is off by default.
    xch = D.StepExchange(ctx, device="cuda", cuda_index=local_rank) if args.pipeline_exchange else None

    def step_dev():
        r = ctx.step_device(api.KDOP, V, F, d_f.data_ptr(), d_q0.data_ptr(), d_q1.data_ptr(), wl["outer_eta"], wl["eta"], 0, rank, world)
        st = ctx.stage_times()
        if xch is not None:
            xch.submit(r.earliest_toi, r.n_vf_hits + r.n_ee_hits, r.n_vf_candidates, r.n_ee_candidates, st)
        else:
            summary["toi"], summary["hits"], summary["stencils"] = D.exchange_step(
                ctx, r.earliest_toi, r.n_vf_hits + r.n_ee_hits, r.n_vf_candidates, r.n_ee_candidates, device="cuda", stage_ms=st)
        summary["stage_ms"] = st
        return r

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the very first step on a mesh also builds the cached topology tables (unique edges, stars, face ranks: two 64-bit sorts
    # over 12 M half-edges at C5) and sizes every buffer of the pool: reported as cold_first_step_ms, outside the timed region
    torch.cuda.synchronize()
    t_cold = time.perf_counter()
    r = step_dev()
    torch.cuda.synchronize()
    cold_ms = (time.perf_counter() - t_cold) * 1e3
    for _ in range(warmup - 1):
        r = step_dev()
    launches = r.n_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_sum = {}
    kern_ms = 0.0
    e0.record()
    for _ in range(args.steps):
        r = step_dev()
        kern_ms += r.ms_broadphase + r.ms_narrowphase
        for k, v in summary["stage_ms"].items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    if xch is not None:
        summary["toi"], summary["hits"], summary["stencils"] = xch.result()      # joins the last step's exchange: inside the timed region
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    # per-stage device times of every rank (diagnostics: the step is as slow as its slowest rank)
    stage_names = sorted(stage_sum)
    st_local = torch.tensor([stage_sum[k] / args.steps for k in stage_names] + [float(r.n_vf_candidates + r.n_ee_candidates)], dtype=torch.float64, device="cuda")
    st_all = [torch.zeros_like(st_local) for _ in range(world)]
    if world > 1:
        dist.all_gather(st_all, st_local)
    else:
        st_all = [st_local]
    st_all = torch.stack(st_all).cpu().numpy()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t[0]) / args.steps
    n_stencils, n_hits = summary["stencils"], summary["hits"]
    value = n_stencils / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (pinned host arrays; H2D + D2H inside the timed region)
    np_q0, np_q1, np_f = h_q0.numpy(), h_q1.numpy(), h_f.numpy()
    if world == 1:
        def step_e2e():
            # copy=False: the hit lists are read where the C ABI leaves them (pinned host buffers of the context)
            return ctx.step(api.KDOP, np_f, np_q0, np_q1, wl["outer_eta"], wl["eta"], None, rank, world, copy=False)
        h2d = 48 * V + 12 * F
    else:
        # sharded job: every rank copies its 1/N slice of the positions from pinned host memory, the slices are all-gathered over
        # NVLink, the step runs on the replicated arrays and the rank's hit lists come back to pinned host memory.  The faces
        # (static topology) stay resident.
        fq0, fq1 = d_q0.view(-1), d_q1.view(-1)
        hq0, hq1 = h_q0.view(-1), h_q1.view(-1)

        def step_e2e():
            D.gather_positions(fq0, fq1, hq0, hq1, rank, world)
            ctx.wait_stream(torch.cuda.current_stream().cuda_stream)
            return ctx.step_device_hits(api.KDOP, V, F, d_f.data_ptr(), d_q0.data_ptr(), d_q1.data_ptr(), wl["outer_eta"], wl["eta"], 0, rank, world)
        h2d = -(-48 * V // world)
    for _ in range(2):
        step_e2e()
    barrier()
    e0.record()
    for _ in range(args.steps):
        er = step_e2e()
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0]) / args.steps
    d2h = 24 * (er["n_vf_hits"] + er["n_ee_hits"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-stage CUDA-event times of this rank, averaged over the timed steps)
    stages = {k: v / args.steps for k, v in stage_sum.items()}
    fp64_peak = ctx.fp64_peak_tflops()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    nvf, nee = r.n_vf_candidates, r.n_ee_candidates
    dom = max(stages, key=stages.get)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_c5.json"))) if args.workload == "cloth1415" and world == 1 else {}
    except Exception:
        pass
    if dom in ("np_vf", "np_ee"):
        # the narrowphase is a pipeline of kernels per stencil type (cull, one kernel per polynomial, root isolation,
        # combine); its roofline is taken over the whole group: algorithmic FP64 flops of SURVEY.md 8(d) / group time
        # (the edge-edge and the vertex-face run execute side by side on two streams: one group, one time)
        np_ms_all = stages.get("np_vf", 0.0) + stages.get("np_ee", 0.0)
        flops = nvf * FLOP_VF + nee * FLOP_EE
        ach = flops / (np_ms_all * 1e-3) / 1e12
        roof = dict(kernel="narrowphase kernel group, edge-edge and vertex-face runs side by side (np_cull / np_stage / np_export / np_ve* / solve_kernel / np_window / np_combine; largest: solve_kernel<6>)",
                    bound="fp64", achieved=ach, peak=fp64_peak,
                    unit="TFLOP/s", frac=ach / fp64_peak if fp64_peak else None, traffic=traffic.get("narrowphase_bytes"),
                    traffic_source=traffic.get("source"),
                    peak_source="measured live: ccd_fp64_peak (16 independent DFMA chains/thread on all SMs); MEASURED_PEAKS.json has no FP64 entry",
                    algorithmic="%.0f flop x %d edge-edge + %.0f flop x %d vertex-face stencils (coefficient construction only; root isolation and early exits excluded)" % (
                        FLOP_EE, nee, FLOP_VF, nvf),
                    group_ms=np_ms_all, achieved_gbs=(nvf + nee) * BYTES_STENCIL / (np_ms_all * 1e-3) / 1e9)
    else:
        # broadphase stage: algorithmic bytes per SURVEY.md §8(d)
        raw = 15.0 * r.n_face_pairs
        bytes_bp = 892.0 * F + 48.0 * raw + 16.0 * (nvf + nee)
        bp_ms = sum(v for k, v in stages.items() if not k.startswith("np_") and k != "topology")
        ach = bytes_bp / (bp_ms * 1e-3) / 1e9
        roof = dict(kernel="broadphase kernel group (dominant stage: %s)" % dom, bound="hbm", achieved=ach, peak=hbm_peak, unit="GB/s",
                    frac=ach / hbm_peak, traffic=traffic.get("broadphase_bytes"), traffic_source=traffic.get("source"), peak_source=hbm_src,
                    algorithmic="892 B/face + 48 B/raw stencil + 16 B/unique stencil over all broadphase kernels")
    # the second roofline of the step (the one not dominant), so both phases are always on the line
    raw = 15.0 * r.n_face_pairs * (1 if world > 1 else 1)
    bp_ms = sum(v for k, v in stages.items() if not k.startswith("np_") and k != "topology")
    np_ms = stages.get("np_vf", 0.0) + stages.get("np_ee", 0.0)
    phases = dict(broadphase=dict(ms=bp_ms, achieved_gbs=(892.0 * F + 48.0 * raw + 16.0 * (nvf + nee)) / (bp_ms * 1e-3) / 1e9 if bp_ms else None, peak_gbs=hbm_peak),
                  narrowphase=dict(ms=np_ms, achieved_tflops=(nvf * FLOP_VF + nee * FLOP_EE) / (np_ms * 1e-3) / 1e12 if np_ms else None, peak_tflops=fp64_peak,
                                   stencils_per_s=(nvf + nee) / (np_ms * 1e-3) if np_ms else None))
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=warmup, ms_per_step=ms_per_step,
               higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic" if args.workload.startswith("cloth") else "reference mesh fixture",
               config=config, clocks=clocks,
               e2e=dict(value=n_stencils / (e2e_ms * 1e-3), unit=UNIT, ms_per_step=e2e_ms, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
               gpu_launches=int(launches) * args.steps,
               roofline=roof, phases=phases,
               roofline_broadphase=dict(kernel="broadphase kernel group (boxes, Morton sort, cluster tree, pair traversal, face-pair tests, adjacency, emission)",
                                        bound="hbm", achieved=phases["broadphase"]["achieved_gbs"], peak=hbm_peak, unit="GB/s",
                                        frac=(phases["broadphase"]["achieved_gbs"] / hbm_peak) if phases["broadphase"]["achieved_gbs"] else None,
                                        traffic=traffic.get("broadphase_bytes"), traffic_source=traffic.get("source"), peak_source=hbm_src,
                                        algorithmic="892 B/face + 48 B/raw stencil + 16 B/unique stencil (SURVEY.md 8d)", group_ms=bp_ms),
               stages_ms=stages, stages_ms_max_over_ranks={k: float(st_all[:, i].max()) for i, k in enumerate(stage_names)},
               stencils_per_rank=[int(x) for x in st_all[:, -1]], stages_ms_per_rank={k: [round(float(x), 3) for x in st_all[:, i]] for i, k in enumerate(stage_names)}, kernel_ms_per_step=kern_ms / args.steps,
               counts=dict(vertices=V, faces=F, stencils=n_stencils, vf_rank0=int(nvf), ee_rank0=int(nee), hits=n_hits,
                           face_pairs_rank0=int(r.n_face_pairs), tree_candidates_rank0=int(r.n_tree_candidates), vf_deferred_rank0=int(r.n_vf_deferred), ee_deferred_rank0=int(r.n_ee_deferred),
                           stencils_per_face=n_stencils / float(F)),
               earliest_toi=summary["toi"], fp64_peak_tflops=fp64_peak, cold_first_step_ms=cold_ms)
    if world == 1 and not args.no_parity:
        try:
            out["parity"] = parity_block(ctx, api, wl)
        except Exception as e:
            out["parity"] = dict(error=str(e))
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb, _, _, _ = run_cpu_arm(args.workload, 1, 0)
            cb["full_size_reference"] = full_size_reference(args.workload)
            out["cpu_baseline"] = cb
        except Exception as e:  # the checker is optional at bench time
            out["cpu_baseline"] = dict(value=None, unit=UNIT, cores=1, kind="unavailable", sample=str(e))
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
